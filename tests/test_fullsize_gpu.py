"""Size-independent properties at BASELINE.json's full (per-GPU) sizes: the oracle cannot follow
there, so these check what must hold whatever the inputs are -- the fused call against the two calls,
FK -> IK -> FK closure, linearity of the calibration reduction, exact end points of trajectories."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


def _free_gb():
    torch.cuda.empty_cache()           # memory cached by earlier tests counts as free
    free, _ = torch.cuda.mem_get_info()
    return free / 2**30


def test_fused_roundtrip_at_chunk_size():
    """2^26 poses (the bench's per-step chunk of configs[1]/[2]): masks and angles of the fused kernel
    equal rk_fanuc_ik on the pose planes it wrote, bit for bit; wherever some branch is valid, one of
    them reproduces the joint vector (63 % of uniform samples, as the oracle finds on small sets);
    det J = 0 never appears as NaN; T's last row is exactly (0, 0, 0, 1)."""
    from roboticmanipulators_b200.robots.fanuc import Fanuc165F
    if _free_gb() < 120:
        pytest.skip("needs ~100 GB of free HBM")
    robot = Fanuc165F()
    n = 1 << 26
    g = torch.Generator(device="cuda").manual_seed(11)
    lim = torch.tensor(np.asarray(robot.qs_lim_rad), device="cuda")
    store = torch.rand((6, n), generator=g, device="cuda", dtype=torch.float64)
    store = lim[:, :1] + store * (lim[:, 1:] - lim[:, :1])
    T, J, det, qik, val = robot.fk_jac_ik_batch(store.t())
    assert bool(torch.isfinite(det).all())
    assert bool((T[:, 3, :3] == 0).all()) and bool((T[:, 3, 3] == 1).all())
    q2, v2 = robot.inverse_kinematics_batch(T)
    assert torch.equal(val, v2)
    assert torch.equal(qik.nan_to_num(7.0), q2.nan_to_num(7.0))
    del q2, v2, J
    frac_valid = val.any(dim=1).double().mean().item()
    assert 0.8 < frac_valid < 0.95           # 0.867: at least one branch within the joint limits
    # closure in chunks (keeps the temporaries small)
    found = 0
    step = 1 << 23
    for lo in range(0, n, step):
        d = (qik[lo:lo + step] - store[:, lo:lo + step].t()[:, None, :]).abs()
        d = torch.minimum(d, (d - 2 * np.pi).abs())
        best = torch.where(val[lo:lo + step], d.amax(dim=2), torch.full_like(d[..., 0], np.inf)).amin(dim=1)
        found += int((best < 1e-6).sum())
    assert 0.60 < found / n < 0.66


def test_calibration_reduction_is_linear_at_1e8():
    """configs[4]: 1e8 synthetic measurements.  H, g, r^T r of the whole set equal the sum over two
    halves (1e-12 relative: only the summation order differs), the result is deterministic, and the
    solution recovers the parameters the data were generated with."""
    import bench_workloads as W

    class A:
        batch = 100_000_000
        ld_pad = 0
    if _free_gb() < 40:
        pytest.skip("needs ~25 GB of free HBM")
    cw = W.CalibWorkload(A, 0, 1)
    mc = cw.mc
    p0 = np.zeros(18)
    full = mc.normal_equations(cw.q, cw.pts, cw.T_base, cw.T_tools, parameters=p0, process_group=False)
    again = mc.normal_equations(cw.q, cw.pts, cw.T_base, cw.T_tools, parameters=p0, process_group=False)
    assert torch.equal(full, again)
    h = cw.m // 2
    a = mc.normal_equations(cw.q[:h], cw.pts[:h], cw.T_base, cw.T_tools, parameters=p0, process_group=False)
    b = mc.normal_equations(cw.q[h:], cw.pts[h:], cw.T_base, cw.T_tools, parameters=p0, process_group=False)
    s = (a + b).cpu().numpy()
    f = full.cpu().numpy()
    assert np.max(np.abs(s - f) / np.maximum(np.abs(f), 1e-300 + np.abs(f).max() * 1e-9)) < 1e-10
    H, gv, rss = mc.unpack_normal_equations(f, 18)
    params = np.linalg.inv(H).dot(gv)
    # Gauss-Newton step from the nominal model towards the generating parameters (noise 0.044 mm on
    # 3e8 points; the linearisation error of a ~5 mm parameter change dominates)
    assert np.max(np.abs(params - cw.true_p)) < 0.05
    assert 0.04 < np.sqrt(rss / (3 * cw.m)) < 10.0


def test_trajectories_end_exactly_at_the_goal():
    """configs[3]: 2^17 trajectories x 1000 samples per call.  The last sample of every trajectory
    the reference can build equals the goal EXACTLY (np.linspace forces the end point and
    q_dec(t1) = q_f in the reference's arithmetic only up to rounding -- so the bound is 1e-12), the
    fused FK variant returns the same samples and counts as the plain sampler, and the FP32 variant
    the same counts."""
    from roboticmanipulators_b200.robots import RRRRobot
    from roboticmanipulators_b200.utils.trajectory_generator import TrajectoryGenerator
    if _free_gb() < 110:
        pytest.skip("needs ~90 GB of free HBM")
    nt, n = 1 << 17, 1000
    g = torch.Generator(device="cuda").manual_seed(5)
    q0 = (torch.rand((nt, 3), generator=g, device="cuda", dtype=torch.float64) * 2 - 1) * np.pi
    qf = (torch.rand((nt, 3), generator=g, device="cuda", dtype=torch.float64) * 2 - 1) * np.pi
    tg = TrajectoryGenerator(1, 20, 1, 10, control_freq=10)
    q, dq, ddq, t, L = tg.generate_p2p_trajectory_batch(q0, qf, n)
    assert int(L.min()) > 900 and int(L.max()) <= n
    last = q[torch.arange(nt, device="cuda"), (L - 1).long()]
    assert (last - qf).abs().max().item() < 1e-12
    assert bool((q[:, 0] == q0).all())
    assert bool((dq[torch.arange(nt, device="cuda"), (L - 1).long()].abs() < 1e-9).all())
    q2, dq2, ddq2, t2, L2, T, J = tg.generate_p2p_trajectory_batch(q0, qf, n, robot=RRRRobot())
    assert torch.equal(L, L2) and torch.equal(q, q2) and torch.equal(t, t2)
    del q2, dq2, ddq2, t2
    # FK of the last sample is the pose of the goal
    Tg = RRRRobot().forward_kinematics_batch(qf, layout="aos")
    Tl = T[torch.arange(nt, device="cuda"), (L - 1).long()]
    assert (Tl - Tg).abs().max().item() < 1e-11
    del T, J, q, dq, ddq, t
    torch.cuda.empty_cache()
    out32 = tg.generate_p2p_trajectory_batch(q0, qf, n, robot=RRRRobot(), dtype=torch.float32)
    assert torch.equal(out32[4], L)
