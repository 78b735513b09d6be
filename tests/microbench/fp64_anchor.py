"""Independent FP64 anchors next to the library's own DFMA / DMMA probes: cuBLAS DGEMM through
torch.matmul (8192^3 and 4096^3), best of 5 and sustained over ~2 s.  Run under gpurun."""
import json
import time

import torch

out = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["dgemm_%d_burst_tflops" % n] = 2.0 * n ** 3 / best / 1e9
    reps = max(3, int(2000.0 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    out["dgemm_%d_sustained_tflops" % n] = 2.0 * n ** 3 * reps / e0.elapsed_time(e1) / 1e9
print(json.dumps(out))
