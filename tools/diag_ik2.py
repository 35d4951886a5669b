import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import np_oracle as O
import ikcond
from roboticmanipulators_b200.robots.fanuc import Fanuc165F
c = O.FanucConst(); robot = Fanuc165F()
rng = np.random.default_rng(4)
q = rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(200000, 6))
T = O.fanuc_fk(q, c); T[::100, :3, 3] *= 3.0
T32 = torch.tensor(T, device="cuda", dtype=torch.float32)
qs, vs = robot.inverse_kinematics_batch(T32, layout="soa")
qa, va = robot.inverse_kinematics_batch(T32, layout="aos")
qs, vs, qa, va = qs.cpu().numpy(), vs.cpu().numpy(), qa.cpu().numpy(), va.cpu().numpy()
d = np.abs(qs - qa); d[..., [0, 3, 5]] = np.abs((qs - qa)[..., [0, 3, 5]] + np.pi) % (2 * np.pi) - np.pi
bad = np.nonzero((np.abs(d).max(axis=(1, 2)) > 1e-3) | (vs != va).any(axis=1))[0]
print("poses differing between FP32 SOA and AOS:", len(bad), bad[:40])
T64 = T32.cpu().numpy().astype(np.float64)
q_ref, v_ref, det = ikcond.ik_details(T64, O.CANONICAL_8, c)
np.set_printoptions(precision=4, suppress=True, linewidth=200)
for i in bad[:6]:
    print("pose", i, "lane", i % 32, "tile", i // 32, "unreach", det["unreachable"][i], "r11", det["r_11"][i], "r13", det["r_13"][i], "A", det["A"][i], "rxy", det["rxy"][i, 0])
    print(" soa valid", vs[i].astype(int), "aos valid", va[i].astype(int), "ref", v_ref[i].astype(int))
    for b in range(8):
        print("  b", b, "soa", qs[i, b], "aos", qa[i, b], "ref", q_ref[i, b])
# neighbours of a bad pose: is the whole tile off?
if len(bad):
    i = bad[0]; t0 = (i // 32) * 32
    print("tile diffs", [int(np.abs(d[j]).max() > 1e-3) for j in range(t0, t0 + 32)])
