"""Measures the FP64 DFMA / DMMA peaks and an HBM copy on the GPU box (used for the roofline)."""
import ctypes as C
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesian_quadrature_optimization_b200 import _cabi

lib = _cabi.load()
torch.cuda.set_device(0)
lib.bqo_set_device(0)
out = {}
ws = torch.empty(int(lib.bqo_fp64_peak_probe_workspace_bytes()), dtype=torch.uint8, device="cuda")
for kind, name in [(0, "dfma"), (1, "dmma")]:
    ms, fl = C.c_double(), C.c_double()
    for iters in (20000, 200000):
        rc = lib.bqo_fp64_peak_probe(kind, iters, ws.data_ptr(), C.byref(ms), C.byref(fl))
        assert rc == 0, rc
        out["%s_tflops_%d" % (name, iters)] = fl.value / (ms.value * 1e-3) / 1e12
a = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
b = torch.empty_like(a)
ms = C.c_double()
rc = lib.bqo_hbm_copy_probe(b.data_ptr(), a.data_ptr(), a.numel() * 8, 5, C.byref(ms))
assert rc == 0
out["hbm_copy_gbs"] = 2 * a.numel() * 8 / (ms.value * 1e-3) / 1e9
print(json.dumps(out))
