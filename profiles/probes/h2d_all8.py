"""Concurrent pinned host -> device copy bandwidth on 1/2/4/8 GPUs of one box: the hardware cap of the host-stream `e2e` path
(every rank copies its own per-filter streams from host memory every step).  Launch under torchrun, one rank per GPU:

    for n in 1 2 4 8; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
        --master-port 29511 profiles/probes/h2d_all8.py; done

Each rank: one pinned buffer (1 GiB by default), cudaMemcpyAsync to its GPU, 10 repetitions after a barrier; prints the per-rank
rates, the aggregate, the NUMA node of each GPU and the host's NUMA layout."""
import os
import sys
import time

import torch
import torch.distributed as dist


def numa_of(local):
    try:
        pr = torch.cuda.get_device_properties(local)
        bus = f"{int(pr.pci_domain_id):04x}:{int(pr.pci_bus_id):02x}:{int(pr.pci_device_id):02x}.0"
        return bus, int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
    except Exception as e:
        return str(e), None


def main():
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = int(os.environ.get("H2D_BYTES", 1 << 30))
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.fill_(rank + 1)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rate = torch.tensor([nbytes * reps / dt / 1e9], dtype=torch.float64, device="cuda")
    rates = [torch.zeros_like(rate) for _ in range(world)]
    if world > 1:
        dist.all_gather(rates, rate)
    else:
        rates = [rate]
    bus, node = numa_of(local)
    info = [None] * world
    if world > 1:
        dist.all_gather_object(info, (bus, node))
    else:
        info = [(bus, node)]
    if rank == 0:
        nodes = sorted(x for x in os.listdir("/sys/devices/system/node") if x.startswith("node") and x[4:].isdigit())
        r = [float(x.item()) for x in rates]
        print(f"n_gpus {world}: per-rank GB/s {[round(x, 1) for x in r]} aggregate {sum(r):.1f} GB/s; GPUs (pci, numa): {info}; "
              f"host NUMA nodes: {len(nodes)}; cpus: {os.cpu_count()}", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
