import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roboticmanipulators_b200 import _lib
lib = _lib.load()
n = 1 << 26
inp = torch.rand((12, n), dtype=torch.float64, device="cuda")
out = torch.empty((49, n), dtype=torch.float64, device="cuda")
for nin, nout in ((12, 49), (0, 49), (12, 0)):
    for cs in (1,):
        for bps in (4, 8):
            best = 1e9
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(lib.rk_probe_stream(_lib.ptr(inp), nin, _lib.ptr(out), max(nout,1), n, n, 1, cs, bps, _lib.stream_ptr()))
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"in {nin} out {nout} cs {cs} ctas/sm {bps}: {best:.3f} ms  {(nin+nout)*8*n/best/1e6:.0f} GB/s")
