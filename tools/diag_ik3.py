import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import np_oracle as O
from roboticmanipulators_b200.robots.fanuc import Fanuc165F
c = O.FanucConst(); robot = Fanuc165F()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rng = np.random.default_rng(4)
q = rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(n, 6))
T = O.fanuc_fk(q, c); T[::100, :3, 3] *= 3.0
for dt in (torch.float32, torch.float64):
    Td = torch.tensor(T, device="cuda", dtype=dt)
    for layout in ("aos", "soa"):
        q0, v0 = robot.inverse_kinematics_batch(Td, layout=layout)
        q0, v0 = q0.clone(), v0.clone()
        nbad = 0
        for r in range(reps):
            junk = torch.randn(1 << 22, device="cuda")        # churn the allocator / L2
            q1, v1 = robot.inverse_kinematics_batch(Td, layout=layout)
            same = torch.equal(q1.nan_to_num(7.0), q0.nan_to_num(7.0)) and torch.equal(v1, v0)
            if not same:
                nbad += 1
                if nbad == 1:
                    dd = ((q1.nan_to_num(7.0) != q0.nan_to_num(7.0)).any(dim=2).any(dim=1) | (v1 != v0).any(dim=1)).nonzero().flatten()
                    print("   differing poses", dd[:20].tolist(), "count", dd.numel())
        print(dt, layout, "nondeterministic repeats:", nbad, "of", reps)
