"""Host<->device copy bandwidth of this box with pinned buffers: H2D alone, D2H alone, both at once
(the floor under bench.py's e2e figure: 4.8 GB in + 4.8 GB out per step)."""
import time, torch
n = 1_200_000_000                      # 4.8 GB of float32
h_in = torch.empty(n, dtype=torch.float32, pin_memory=True); h_in.fill_(1.0)
h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
d_a = torch.empty(n, dtype=torch.float32, device="cuda"); d_b = torch.ones(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def run(h2d, d2h, reps=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

run(True, True, 1)
for name, a, b in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
    t = run(a, b)
    print(f"{name}: {t * 1e3:.1f} ms per 4.8 GB direction -> {4.8 / t:.1f} GB/s per direction")
