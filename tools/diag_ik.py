"""Diagnostic (GPU box): worst err / tolerance of the CUDA IK against the oracle under the
conditioning model of tests/ikcond.py, FP64 and FP32."""
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import np_oracle as O
import ikcond
from roboticmanipulators_b200.robots.fanuc import Fanuc165F
c = O.FanucConst(); robot = Fanuc165F()
def worst(qg, q_ref, v_ref, det, base, cond_ok, live=None):
    tol, well = ikcond.ik_tolerance(q_ref, det, base, cond_ok)
    err = ikcond.ik_diff(qg, q_ref.copy())
    ok = v_ref[..., None] & ~np.isnan(q_ref)
    if live is not None: ok &= live[..., None]
    r = np.where(ok, err / tol, 0)
    for j in range(6):
        print(f"   joint {j}: max err/tol {r[...,j].max():.3f}   (well-conditioned only: {np.where(well[...,j], r[...,j], 0).max():.3f}; raw err there {np.where(well[...,j]&ok[...,j], err[...,j], 0).max():.2e})")
    print("   well frac", well[ok].mean())
rng = np.random.default_rng(99)
q = rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(300000, 6))
T = O.fanuc_fk(q, c); T[::50, :3, 3] *= 1.6
q_ref, v_ref, det = ikcond.ik_details(T, O.CANONICAL_8, c)
for layout in ("soa", "aos"):
    qg, vg = robot.inverse_kinematics_batch(torch.tensor(T, device="cuda"), layout=layout)
    qg, vg = qg.cpu().numpy(), vg.cpu().numpy()
    print("f64", layout, "mask diffs", int((vg != v_ref).sum()))
    worst(qg, q_ref, v_ref, det, 1e-12, ikcond.COND_OK_F64)
rng = np.random.default_rng(4)
q = rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(200000, 6))
T = O.fanuc_fk(q, c); T[::100, :3, 3] *= 3.0
T32 = torch.tensor(T, device="cuda", dtype=torch.float32)
T64 = T32.cpu().numpy().astype(np.float64)
q_ref, v_ref, det = ikcond.ik_details(T64, O.CANONICAL_8, c)
for layout in ("soa", "aos"):
    qg, vg = robot.inverse_kinematics_batch(T32, layout=layout)
    qg, vg = qg.cpu().numpy().astype(np.float64), vg.cpu().numpy()
    clear = ikcond.mask_margin(det, c, 1e-3)
    print("f32", layout, "clear frac", clear.mean(), "mask diffs in clear", int((vg[clear] != v_ref[clear]).sum()), "overall agree", (vg == v_ref).mean())
    worst(qg, q_ref, v_ref & vg, det, 1e-5, ikcond.COND_OK_F32, live=clear)
