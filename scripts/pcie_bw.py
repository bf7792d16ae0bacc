"""Diagnostic: pinned host <-> device copy bandwidth, one direction and both at once."""
import torch

n = 1 << 29  # 512 MiB
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    d_a.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_b, non_blocking=True)


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


for name, fn, nbytes in (("h2d", h2d, n), ("d2h", d2h, n), ("both", both, 2 * n)):
    ms = timed(fn)
    print(f"{name}: {nbytes / ms / 1e6:.1f} GB/s total ({ms:.2f} ms)")
