import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import np_oracle as O
import ikcond
from roboticmanipulators_b200.robots.fanuc import Fanuc165F
c = O.FanucConst(); robot = Fanuc165F()
np.set_printoptions(precision=5, suppress=True, linewidth=220)
pre = sys.argv[1] if len(sys.argv) > 1 else "f64"
if pre == "f64":
    rng = np.random.default_rng(99)
    q = rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(300000, 6))
    T = O.fanuc_fk(q, c); T[::50, :3, 3] *= 1.6
    for layout in ("soa", "aos"):
        qg, vg = robot.inverse_kinematics_batch(torch.tensor(T, device="cuda"), layout=layout)
        qg, vg = qg.cpu().numpy(), vg.cpu().numpy()
rng = np.random.default_rng(4)
q = rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(200000, 6))
T = O.fanuc_fk(q, c); T[::100, :3, 3] *= 3.0
T32 = torch.tensor(T, device="cuda", dtype=torch.float32)
T64 = T32.cpu().numpy().astype(np.float64)
res = {}
for layout in ("soa", "aos", "aos2", "soa2"):
    qg, vg = robot.inverse_kinematics_batch(T32, layout=layout[:3])
    res[layout] = (qg.cpu().numpy(), vg.cpu().numpy())
for a, b in (("soa", "aos"), ("aos", "aos2"), ("soa", "soa2")):
    qa, va = res[a]; qb, vb = res[b]
    diff = (np.nan_to_num(qa, nan=7.0) != np.nan_to_num(qb, nan=7.0)).any(axis=(1, 2)) | (va != vb).any(axis=1)
    idx = np.nonzero(diff)[0]
    print(a, "vs", b, ": poses that differ bitwise:", len(idx), idx[:30])
    big = [i for i in idx if np.nanmax(np.abs(qa[i] - qb[i])) > 1e-3 or (va[i] != vb[i]).any()]
    print("    with |diff| > 1e-3 or mask diff:", len(big), big[:30])
    for i in big[:3]:
        print("   pose", i, "lane", i % 32, "tile", i // 32)
        print("   ", a, va[i].astype(int), qa[i].ravel()[:18])
        print("   ", b, vb[i].astype(int), qb[i].ravel()[:18])
