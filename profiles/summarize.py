"""Turns gpurun_out/*.ncu-rep and launch lists into the tracked summaries under profiles/ (run in the authoring container).

    python profiles/summarize.py full  gpurun_out/prof_window_r1_final.ncu-rep  profiles/r1_window_kernel_ncu_full.md
    python profiles/summarize.py list  gpurun_out/launches_r1_final.csv         profiles/r1_launches.md
    python profiles/summarize.py inputs gpurun_out/prof_window_r2.ncu-rep       profiles/roofline_inputs.json
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full summary: `{rep}`", "", f"kernel: `{vals[col['Kernel Name']]}`", "", "| metric | unit | value |", "|---|---|---|"]
    for k in KEYS:
        if k in col:
            lines.append(f"| {k} | {units[col[k]]} | {vals[col[k]]} |")
    lines += ["", "## warp stall reasons (cycles per issued instruction, smsp__average_warps_issue_stalled_*_per_issue_active)", "",
              "| reason | ratio |", "|---|---|"]
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            lines.append(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {vals[col[h]]} |")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    shdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(shdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0
    tot = sum(f(r, "# Samples") for r in data) or 1.0
    lines += ["", "## sampled stall share over all SASS instructions (source page)", "", "| stall | share |", "|---|---|"]
    for k in [h for h in shdr if h.startswith("stall_") and "Not Issued" not in h]:
        lines.append(f"| {k} | {100 * sum(f(r, k) for r in data) / tot:.1f}% |")
    ops = collections.Counter()
    for r in data:
        toks = r[ix["Source"]].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        ops[op.split(".")[0]] += f(r, "Instructions Executed")
    lines += ["", "## executed warp-instructions by opcode (top 14)", "", "| opcode | executed | share |", "|---|---|---|"]
    n = sum(ops.values()) or 1.0
    for op, c in ops.most_common(14):
        lines.append(f"| {op} | {c:.4g} | {100 * c / n:.1f}% |")
    open(out, "w").write("\n".join(lines) + "\n")


def launch_list(path, out):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v_ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v_ms
    tot = sum(a[1] for a in agg.values()) or 1.0
    lines = [f"# kernel launch list (ncu --metrics gpu__time_duration.sum --clock-control none): `{path}`", "",
             "Per-launch times under ncu are serialised and cold-cache: compare SHARES, not absolutes.", "",
             "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for name, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{name}` | {c} | {ms:.3f} | {100 * ms / tot:.1f}% |")
    open(out, "w").write("\n".join(lines) + "\n")


def roofline_inputs(rep, out):
    """profiles/roofline_inputs.json from an `ncu --set full` capture of ekf_window_split_kernel<128> (one launch of the bench workload:
    200 IMU steps + 25 updates per filter).  Stamped with the source hash of the library in the tree, which must be the one profiled:
    bench.py reports the profiler-derived fields only while the loaded library has the same hash."""
    import json
    import os
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}

    def val(k, scale=None):
        v = float(vals[col[k]].replace(",", ""))
        u = units[col[k]].lower()
        if scale == "bytes":
            v *= {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        return v
    assert "ekf_window_split_kernel<128" in vals[col["Kernel Name"]].replace("(int)", ""), vals[col["Kernel Name"]]
    filters = int(val("launch__grid_size")) * 128
    rd, wr = val("dram__bytes_read.sum", "bytes"), val("dram__bytes_write.sum", "bytes")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    shdr, data = srows[1], srows[2:]
    ix = {h: i for i, h in enumerate(shdr)}
    flop = 0.0
    for r in data:
        toks = r[ix["Source"]].split()
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
        if op in ("DFMA", "DMUL", "DADD"):
            flop += (2.0 if op == "DFMA" else 1.0) * float(r[ix["Predicated-On Thread Instructions Executed"]])
    here = os.path.dirname(os.path.abspath(__file__))
    try:
        old = json.load(open(out))
    except Exception:
        old = {}
    new = {"kernel": "ekf_window_split_kernel<128>",
           "source": f"{os.path.basename(rep)} (ncu --set full --clock-control none, one launch: 200 IMU steps + 25 updates per filter); summary in "
                     "profiles/r2_window_kernel_ncu_full.md",
           "srchash": open(os.path.join(here, "..", "fbus_ekf_b200", "libfbus_ekf.so.srchash")).read().strip(),
           "measured_at_filters": filters,
           "dram__bytes_read.sum": rd, "dram__bytes_write.sum": wr,
           "ekf_window_dram_bytes_per_filter_per_launch": (rd + wr) / filters,
           "fp64_flop_executed_per_filter_per_launch": flop / filters,
           "fp64_pipe_busy_pct_ncu": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
           "issue_active_pct_ncu": val("smsp__issue_active.avg.pct_of_peak_sustained_active")}
    for k in ("refract_kernel", "solve_to_det_kernel"):
        if k in old:
            new[k] = old[k]
    json.dump(new, open(out, "w"), indent=1)
    print(json.dumps(new, indent=1))


if __name__ == "__main__":
    {"full": full, "list": launch_list, "inputs": roofline_inputs}[sys.argv[1]](sys.argv[2], sys.argv[3])
