"""Where the host time of one batched call goes (run on a GPU box)."""
import sys, time
sys.path.insert(0, '.')
import torch
from roboticmanipulators_b200 import _batch, _lib
from roboticmanipulators_b200.robots.fanuc import Fanuc165F
from roboticmanipulators_b200.robots.robot import launch, _raw_stream
r = Fanuc165F(); lib = _lib.load()
q = torch.zeros((10000, 6), dtype=torch.float64, device="cuda") + 0.1
def t(fn, n=2000):
    for _ in range(50): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    dt = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    return round(dt, 2)
print("full fk_jac_ik_batch aos", t(lambda: r.fk_jac_ik_batch(q, layout="aos")))
print("full fk_jac_ik_batch soa", t(lambda: r.fk_jac_ik_batch(q.t().contiguous().t(), layout="soa")))
qs = q.t().contiguous().t()
print("full fk_jac_ik_batch soa (prebuilt view)", t(lambda: r.fk_jac_ik_batch(qs, layout="soa")))
print("full fk batch aos", t(lambda: r.forward_kinematics_batch(q, layout="aos")))
print("prep_input", t(lambda: _batch.prep_input(q, 6, "aos", q.dtype)))
print("4 x alloc_view", t(lambda: [_batch.alloc_view(10000, sh, "aos", q.dtype, q.device) for sh in ((4, 4), (6, 6), (), (8, 6))]))
print("torch.empty uint8", t(lambda: torch.empty((10000, 8), dtype=torch.uint8, device=q.device)))
print("cached_struct", t(lambda: r._cached_struct()))
print("raw_stream+current_device", t(lambda: _raw_stream(torch.cuda.current_device())))
Ts, Js, det, qstore = [_batch.alloc_view(10000, sh, "aos", q.dtype, q.device) for sh in ((4, 4), (6, 6), (), (8, 6))]
vstore = torch.empty((10000, 8), dtype=torch.uint8, device=q.device)
st = r._cached_struct(); fn = lib.rk_fanuc_fk_jac_ik8_f64
print("ctypes call only (2 launches, aos)", t(lambda: fn(st, q.data_ptr(), 10000, 10000, 0, Ts.data_ptr(), Js.data_ptr(), det.data_ptr(), qstore.data_ptr(), vstore.data_ptr(), _raw_stream(0))))
fn1 = lib.rk_fanuc_fk_jac_f64
print("ctypes call only (1 launch fk aos)", t(lambda: fn1(st, q.data_ptr(), 10000, 10000, 0, Ts.data_ptr(), Js.data_ptr(), det.data_ptr(), _raw_stream(0))))
Ts, Js, det, qstore = [_batch.alloc_view(10000, sh, "soa", q.dtype, q.device) for sh in ((4, 4), (6, 6), (), (8, 6))]
vs2 = torch.empty((8, 10000), dtype=torch.uint8, device=q.device)
print("ctypes call only (fused soa, 1 launch)", t(lambda: fn(st, qs.data_ptr(), 10000, 10000, 1, Ts.data_ptr(), Js.data_ptr(), det.data_ptr(), qstore.data_ptr(), vs2.data_ptr(), _raw_stream(0))))
print("view bool", t(lambda: vstore.view(torch.bool)))
