"""GPU-box experiment: what H2D / D2H of the bench's buffers cost by themselves (pinned host memory, torch copies)."""
import torch, time
n = 10_000_000
h = torch.empty((n, 4), dtype=torch.float32).pin_memory()
d = torch.empty((n, 4), dtype=torch.float32, device="cuda")
s = torch.empty(n, dtype=torch.float32, device="cuda")
hs = torch.empty(n, dtype=torch.float32).pin_memory()
for name, fn, nbytes in (("H2D 160 MB", lambda: d.copy_(h, non_blocking=True), 160e6), ("D2H 40 MB", lambda: hs.copy_(s, non_blocking=True), 40e6)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%s: %.3f ms = %.1f GB/s" % (name, ms, nbytes / ms / 1e6))
