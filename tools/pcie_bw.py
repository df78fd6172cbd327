"""Development aid: pinned host <-> device copy bandwidth on this box (the ceiling of bench.py's e2e leg)."""
import time, torch
n = 17_200_800
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
m = 14_790_468
h_out = torch.empty(m, dtype=torch.uint8).pin_memory(); d_out = torch.empty(m, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, iters=200):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(iters):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t) / iters
for name, a, b in (("H2D only", 1, 0), ("D2H only", 0, 1), ("both", 1, 1)):
    dt = run(a, b)
    print(f"{name:9s} {dt * 1e6:7.1f} us/iter  H2D {a * n / dt / 1e9:6.1f} GB/s  D2H {b * m / dt / 1e9:6.1f} GB/s")
