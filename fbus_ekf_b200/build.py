"""In-tree build of the CUDA library (libfbus_ekf.so) for sm_100a with nvcc.  No JIT cache: the built
.so sits next to this file so that it travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libfbus_ekf.so")
SOURCES = ["fbus_capi.cu"]
DEPS = ["fbus_capi.cu", "fbus_kernels.cuh", "fbus_kernel_split.cuh", "fbus_kernel_lane.cuh", "fbus_tmem.cuh", "fbus_math.cuh", "fbus_refract.cuh", "fbus_host_consts.hpp",
        os.path.join("..", "..", "include", "fbus_ekf.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-cudart", "static"]


HASH_FILE = OUT + ".srchash"
LAST_MODE = "not built in this process"  # "compiled" | "up to date (source hash matches)" | "prebuilt, nvcc missing"


def nvcc_version() -> str:
    try:
        out = subprocess.run([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "--version"], capture_output=True, text=True).stdout
        return out.strip().splitlines()[-1] if out.strip() else "unknown"
    except OSError:
        return "nvcc not found"


def source_hash() -> str:
    """content hash of every source the library is built from (mtimes do not survive the copy to the GPU box)"""
    import hashlib
    h = hashlib.sha256()
    for d in DEPS:
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def built_with() -> str:
    """compiler that produced the library on disk (recorded next to it at build time)"""
    try:
        return open(OUT + ".nvcc").read().strip()
    except OSError:
        return "unknown"


def needs_build() -> bool:
    if not os.path.exists(OUT) or not os.path.exists(HASH_FILE):
        return True
    return open(HASH_FILE).read().strip() != source_hash()


def build_variant(win_bs: int, suffix: str, extra=()) -> str:
    """experiment builds: libfbus_ekf_<suffix>.so with a different CTA size / macros for the window kernel"""
    out = os.path.join(HERE, f"libfbus_ekf_{suffix}.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [f"-DFBUS_WIN_BS={win_bs}"] + list(extra) + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed")
    for line in (res.stdout + res.stderr).splitlines():
        if "ekf_window" in line or ("spill" in line and "3" in line[:0]):
            pass
    import re
    m = re.search(r"ekf_window_\w*kernel.*?\n.*?\n\s*(\d+ bytes stack frame, \d+ bytes spill stores, \d+ bytes spill loads)", res.stdout + res.stderr, re.S)
    print(suffix, m.group(1) if m else "")
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    global LAST_MODE
    if not force and not needs_build():
        LAST_MODE = "up to date (source hash matches)"
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True)
    except FileNotFoundError:
        if os.path.exists(OUT):  # no nvcc on this machine: keep the prebuilt library that travelled with the repo -- loudly
            sys.stderr.write("fbus_ekf_b200.build: nvcc not found and the sources changed since libfbus_ekf.so was built; "
                             "using the STALE prebuilt library\n")
            LAST_MODE = "prebuilt and STALE, nvcc missing"
            return OUT
        raise
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libfbus_ekf.so")
    with open(HASH_FILE, "w") as f:
        f.write(source_hash())
    with open(OUT + ".nvcc", "w") as f:
        f.write(nvcc_version())
    LAST_MODE = "compiled"
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    build(force=True, verbose=True)
