#!/usr/bin/env python
"""Bare pinned-memory D2H / H2D bandwidth with ALL ranks copying at once -- the control for bench.py's e2e scaling
(VERDICT r1 weak #5): what the box's host<->device paths deliver when N GPUs copy simultaneously, nothing else running.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_bw_ranks.py
Rank 0 prints one JSON line: per-rank and aggregate GB/s for D2H alone, H2D alone and both directions at once."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
aff = "unchanged"
if world > 1:
    import bench
    aff = bench.bind_to_gpu_numa_node(local)
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 30
d_out, d_in = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
h_out, h_in = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(d2h, h2d, reps=6):
    def once():
        if d2h:
            with torch.cuda.stream(s1):
                h_out.copy_(d_out, non_blocking=True)
        if h2d:
            with torch.cuda.stream(s2):
                d_in.copy_(h_in, non_blocking=True)
    once()
    barrier()
    t = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)     # slowest rank
    return n / float(dt.item()) / 1e9


res = {"n_gpus": world, "bytes_per_copy": n, "cpu_affinity_rank0": aff}
for name, a, b in (("d2h_alone", True, False), ("h2d_alone", False, True), ("both_directions", True, True)):
    per = run(a, b)
    res[name] = {"GBps_per_rank_per_direction": per, "GBps_aggregate_per_direction": per * world}
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
