"""Development aid: host topology of a multi-GPU box and the aggregate pinned-copy bandwidth with all GPUs busy.
    python tools/topo_probe.py            (prints topology, then runs one copy process per GPU concurrently)"""
import os, subprocess, sys, time

def child(idx, pin):
    import torch
    if pin:
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
            allowed = os.sched_getaffinity(0)
            use = cpus & allowed
            if use:
                os.sched_setaffinity(0, use)
        except Exception as e:
            print("pin failed", e)
    torch.cuda.set_device(idx)
    n, m = 17_200_800, 14_790_468
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    h_out = torch.empty(m, dtype=torch.uint8).pin_memory(); d_out = torch.empty(m, dtype=torch.uint8, device="cuda")
    h_in.fill_(1); h_out.fill_(0)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(iters):
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(iters):
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); return (time.perf_counter() - t) / iters
    run(50)
    # crude start alignment: wait for the next multiple of 5 s
    time.sleep(5 - time.time() % 5)
    dt = run(1500)
    print(f"gpu {idx} pin={pin} affinity={len(os.sched_getaffinity(0))} cpus: {dt * 1e6:7.1f} us/iter  H2D {n / dt / 1e9:5.1f} D2H {m / dt / 1e9:5.1f} GB/s", flush=True)

if len(sys.argv) > 2 and sys.argv[1] == "child":
    child(int(sys.argv[2]), int(sys.argv[3])); sys.exit(0)

for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"], ["numactl", "-H"]):
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout
        print("$", " ".join(cmd)); print("\n".join(out.splitlines()[:40]))
    except Exception as e:
        print("$", " ".join(cmd), "->", e)
print("sched_getaffinity:", sorted(os.sched_getaffinity(0)))
import torch
ng = torch.cuda.device_count()
for pin in (0, 1):
    ps = [subprocess.Popen([sys.executable, __file__, "child", str(i), str(pin)]) for i in range(ng)]
    for p in ps: p.wait()
