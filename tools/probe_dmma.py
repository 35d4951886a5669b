"""Measured FP64 FMA rate of the CUDA-core path (DFMA) and the tensor-core path (DMMA m8n8k4)."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roboticmanipulators_b200 import _lib
lib = _lib.load()
sink = torch.zeros(8, dtype=torch.float64, device="cuda")
for name, iters in (("rk_probe_dfma", 300000), ("rk_probe_dmma", 100000)):
    fn = getattr(lib, name)
    total = C.c_int64(0)
    _lib.check(fn(2000, _lib.ptr(sink), C.byref(total), _lib.stream_ptr()))
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(fn(iters, _lib.ptr(sink), C.byref(total), _lib.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        best = max(best, total.value / (e0.elapsed_time(e1) * 1e-3))
    print(f"{name}: {best:.4g} FMA/s = {2 * best / 1e12:.2f} TFLOP/s")
