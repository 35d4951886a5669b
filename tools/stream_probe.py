"""What can a 6-planes-in / 53-planes-out SOA access pattern reach on this GPU?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roboticmanipulators_b200 import _lib
lib = _lib.load()
n = 1 << 26
inp = torch.rand((6, n), dtype=torch.float64, device="cuda")
out = torch.empty((53, n), dtype=torch.float64, device="cuda")
for nin, nout in ((6, 53), (0, 53), (6, 6), (1, 1)):
    for vec in (1, 2):
        for cs in (0, 1):
            for bps in (4, 8, 16):
                best = 1e9
                for _ in range(4):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    _lib.check(lib.rk_probe_stream(_lib.ptr(inp), nin, _lib.ptr(out), nout, n, n, vec, cs, bps, _lib.stream_ptr()))
                    e1.record(); torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                print(f"in {nin} out {nout} vec {vec} cs {cs} ctas/sm {bps}: {best:.3f} ms  {(nin+nout)*8*n/best/1e6:.0f} GB/s")
