"""Diagnostic: pinned host <-> device copy bandwidth with N ranks copying AT THE SAME TIME, one
direction and both at once -- the platform ceiling under the host-buffer arm of bench.py (`e2e`),
where every rank moves its particles' state through the one host's memory system.

    python scripts/pcie_bw.py                                            # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/pcie_bw.py                            # N GPUs at once

Prints one JSON line: per-rank and aggregate GB/s for h2d, d2h and duplex."""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n = 1 << 29  # 512 MiB per direction and rank
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def timed(fn, reps=5):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)  # the slowest rank defines the aggregate
    return float(dt.item())


def h2d():
    d_a.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_b, non_blocking=True)


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


out = {"ranks": world, "bytes_per_rank_and_direction": n}
for name, fn, nbytes in (("h2d", h2d, n), ("d2h", d2h, n), ("duplex", both, 2 * n)):
    secs = timed(fn)
    out[name] = {"per_rank_gbs": nbytes / secs / 1e9, "aggregate_gbs": world * nbytes / secs / 1e9}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
