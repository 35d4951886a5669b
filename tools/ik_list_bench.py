import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from roboticmanipulators_b200.robots.fanuc import Fanuc165F, ALL_12
robot = Fanuc165F()
n = 1 << 24
lim = torch.tensor(np.asarray(robot.qs_lim_rad), device="cuda")
q = lim[:, 0] + torch.rand((n, 6), device="cuda", dtype=torch.float64) * (lim[:, 1] - lim[:, 0])
T = robot.forward_kinematics_batch(q.t().contiguous().t(), layout="soa")
for name, br in (("single (1,1,0)", [(1, 1, 0)]), ("two", [(1, 1, 0), (-1, 1, 0)]), ("all 12", ALL_12), ("canonical 8", None)):
    for layout in ("soa", "aos"):
        Tin = T if layout == "soa" else T.contiguous()
        f = (lambda: robot.inverse_kinematics_batch(Tin, br, layout=layout)) if br else (lambda: robot.inverse_kinematics_batch(Tin, layout=layout))
        f(); f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"{name:16s} {layout}: {ms:.3f} ms per 2^24 poses = {n / ms / 1e6:.2f}e9 poses/s")
