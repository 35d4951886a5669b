"""GPU parity: trapezoidal trajectories -- sample counts and values are bit-identical to the
reference (golden vectors) and to the oracle on random batches."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


def dev(x):
    return torch.tensor(np.asarray(x, dtype=np.float64), device="cuda")


def test_p2p_golden_bit_exact(golden):
    from roboticmanipulators_b200.utils.trajectory_generator import TrajectoryGenerator
    g = golden("trajectories.npz")
    for i in range(int(g["p2p_count"])):
        spec = g[f"p2p{i}_spec"]
        q0, qf, n, cf = spec[:3], spec[3:6], int(spec[6]), int(spec[7])
        tg = TrajectoryGenerator(1, 20, 1, 10, control_freq=cf)
        q, dq, ddq, t, L = tg.generate_p2p_trajectory_batch(dev(q0[None]), dev(qf[None]), n)
        L = int(L[0])
        if int(g[f"p2p{i}_error"]):
            assert L == 0
            with pytest.raises(ValueError):
                tg.generate_p2p_trajectory(list(q0), list(qf), n=n)
            continue
        ref = g[f"p2p{i}_q"]
        assert L == len(ref)
        assert np.array_equal(q[0, :L].cpu().numpy(), ref)
        assert np.array_equal(dq[0, :L].cpu().numpy(), g[f"p2p{i}_dq"])
        assert np.array_equal(ddq[0, :L].cpu().numpy(), g[f"p2p{i}_ddq"])
        assert np.array_equal(t[0, :L].cpu().numpy(), g[f"p2p{i}_t"])
        assert not q[0, L:].any() and not t[0, L:].any()
        # reference-shaped single call
        assert np.array_equal(tg.generate_p2p_trajectory(list(q0), list(qf), n=n), ref)


@pytest.mark.parametrize("cf", [0, 10])
def test_p2p_random_batch_vs_oracle(cf):
    from roboticmanipulators_b200.utils.trajectory_generator import TrajectoryGenerator
    rng = np.random.default_rng(3 + cf)
    N, n = 300, 257
    q0, qf = rng.uniform(-np.pi, np.pi, (N, 3)), rng.uniform(-np.pi, np.pi, (N, 3))
    qf[5] = q0[5]                       # zero-length move: the reference raises
    qf[6] = q0[6] + 1e-4                # too short for n samples
    tg = TrajectoryGenerator(1, 20, 1, 10, control_freq=cf)
    q, dq, ddq, t, L = tg.generate_p2p_trajectory_batch(dev(q0), dev(qf), n)
    rq, rdq, rddq, rt, rL = O.p2p_batch(q0, qf, n, 1, 20, cf)
    assert np.array_equal(L.cpu().numpy(), rL)
    assert rL[5] == 0
    assert np.array_equal(q.cpu().numpy(), rq)
    assert np.array_equal(dq.cpu().numpy(), rdq)
    assert np.array_equal(ddq.cpu().numpy(), rddq)
    assert np.array_equal(t.cpu().numpy(), rt)


def test_p2p_other_joint_counts():
    from roboticmanipulators_b200.utils.trajectory_generator import TrajectoryGenerator
    rng = np.random.default_rng(11)
    for J in (1, 2, 6, 8):
        q0, qf = rng.uniform(-2, 2, (40, J)), rng.uniform(-2, 2, (40, J))
        tg = TrajectoryGenerator(1.5, 7, 1, 10, control_freq=0)
        q, dq, ddq, t, L = tg.generate_p2p_trajectory_batch(dev(q0), dev(qf), 100)
        rq, rdq, rddq, rt, rL = O.p2p_batch(q0, qf, 100, 1.5, 7, 0)
        assert np.array_equal(L.cpu().numpy(), rL)
        assert np.array_equal(q.cpu().numpy(), rq) and np.array_equal(dq.cpu().numpy(), rdq)
        assert np.array_equal(ddq.cpu().numpy(), rddq) and np.array_equal(t.cpu().numpy(), rt)


def test_trapezoidal_generator_stateful_api(golden):
    from roboticmanipulators_b200.utils.trapezoidal_generator import TrapezoidalGenerator
    tg = TrapezoidalGenerator(1, 20, control_freq=10)
    t1, tau = tg.generate_coefficients(0.0, -2.3)
    assert (t1, tau) == (0.1, 2.3)
    tg.set_current_limits(-2.3 / tau, -2.3 / (tau * t1))
    tg.set_positions(0.0, -2.3)
    ref = O.trap_sample(0.0, -2.3 / tau, -2.3 / (tau * t1), t1, tau, 50)
    assert np.array_equal(tg.get_q(t1, tau, 50), ref[0])
    assert np.array_equal(tg.get_dq(t1, tau, 50), ref[1])
    assert np.array_equal(tg.get_ddq(t1, tau, 50), ref[2])
    assert np.array_equal(tg.get_t(t1, tau, 50), ref[3])
    assert len(ref[0]) == 49
    with pytest.raises(ValueError):
        tg.get_q(0.001, 5.0, 10)
    g = golden("trajectories.npz")
    cases = g["coef_cases"]
    for f in np.unique(cases[:, 2]):
        sel = cases[:, 2] == f
        for v, a in {(r[0], r[1]) for r in cases[sel]}:
            s2 = sel & (cases[:, 0] == v) & (cases[:, 1] == a)
            gen = TrapezoidalGenerator(v, a, control_freq=int(f))
            t1b, taub = gen.generate_coefficients_batch(dev(cases[s2, 3]), dev(cases[s2, 4]))
            assert np.array_equal(np.stack([t1b.cpu().numpy(), taub.cpu().numpy()], 1), g["coef_out"][s2])


def test_fused_rrr_fk_per_sample():
    from roboticmanipulators_b200.robots import RRRRobot
    from roboticmanipulators_b200.utils.trajectory_generator import TrajectoryGenerator
    rng = np.random.default_rng(2)
    N, n = 64, 200
    q0, qf = rng.uniform(-np.pi, np.pi, (N, 3)), rng.uniform(-np.pi, np.pi, (N, 3))
    tg = TrajectoryGenerator(1, 20, 1, 10, control_freq=10)
    q, dq, ddq, t, L, T, J = tg.generate_p2p_trajectory_batch(dev(q0), dev(qf), n, robot=RRRRobot())
    rq, rdq, rddq, rt, rL = O.p2p_batch(q0, qf, n, 1, 20, 10)
    assert np.array_equal(L.cpu().numpy(), rL) and np.array_equal(q.cpu().numpy(), rq)
    live = np.arange(n)[None, :] < rL[:, None]
    T_ref, J_ref = O.rrr_fk(rq), O.rrr_jacobian(rq)
    assert np.max(np.abs(T.cpu().numpy()[live] - T_ref[live])) < 1e-12
    assert np.max(np.abs(J.cpu().numpy()[live] - J_ref[live])) < 1e-12
    assert not T.cpu().numpy()[~live].any() and not J.cpu().numpy()[~live].any()


def test_lin_trajectory_ik_and_joint_velocities(golden):
    from roboticmanipulators_b200.robots import RRRRobot
    from roboticmanipulators_b200.utils.trajectory_generator import TrajectoryGenerator
    g = golden("trajectories.npz")
    robot = RRRRobot()
    for i in range(int(g["lin_count"])):
        spec = g[f"lin{i}_spec"]
        p0, pf, n, cf = spec[:3], spec[3:6], int(spec[6]), int(spec[7])
        tg = TrajectoryGenerator(1, 20, 1, 10, control_freq=cf)
        x, dx, t, q, dq, valid, L = tg.generate_lin_trajectory_batch(dev(p0[None]), dev(pf[None]), n, robot=robot)
        L = int(L[0])
        assert L == len(g[f"lin{i}_x"])
        assert np.array_equal(x[0, :L].cpu().numpy(), g[f"lin{i}_x"])
        assert np.array_equal(dx[0, :L].cpu().numpy(), g[f"lin{i}_dx"])
        assert np.array_equal(t[0, :L].cpu().numpy(), g[f"lin{i}_t"])
        assert np.max(np.abs(q[0, :L].cpu().numpy() - g[f"lin{i}_q"])) < 1e-12
        ref = g[f"lin{i}_dq"]
        ok = ~np.isnan(ref).any(-1)
        assert np.max(np.abs(dq[0, :L].cpu().numpy()[ok] - ref[ok])) < 1e-9 * max(1.0, np.abs(ref[ok]).max())
        xs, dxs, ts = tg.generate_lin_trajectory(list(p0), list(pf), n=n)
        assert np.array_equal(xs, g[f"lin{i}_x"]) and np.array_equal(ts, g[f"lin{i}_t"])
        qs = robot.move_via_points(xs)
        dqs = tg.get_dq_from_dx(qs, dxs)
        assert np.max(np.abs(dqs.T[ok] - ref[ok])) < 1e-8 * max(1.0, np.abs(ref[ok]).max())
    rng = np.random.default_rng(9)
    N, n = 200, 128
    c = np.array([0.0, 0.0, 1.0])
    dirs = rng.normal(size=(2, N, 3))
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    p0 = c + dirs[0] * rng.uniform(0.4, 1.8, (N, 1))
    pf = c + dirs[1] * rng.uniform(0.4, 1.8, (N, 1))
    tg = TrajectoryGenerator(1, 20, 1, 10, control_freq=0)
    x, dx, t, q, dq, valid, L = tg.generate_lin_trajectory_batch(dev(p0), dev(pf), n, robot=robot)
    rx, rdx, rt, rq, rdq, rvalid, rL = O.lin_rrr_batch(p0, pf, n, 1, 10, 0)
    assert np.array_equal(L.cpu().numpy(), rL) and np.array_equal(x.cpu().numpy(), rx)
    assert np.array_equal(valid.cpu().numpy(), rvalid)
    assert np.max(np.abs(q.cpu().numpy() - rq)) < 1e-9
    good = rvalid & np.isfinite(rdq).all(-1) & (np.abs(rdq).max(-1) < 1e3)
    assert np.max(np.abs(dq.cpu().numpy()[good] - rdq[good])) < 1e-7


def test_quintic_profiles(golden):
    from roboticmanipulators_b200.utils.polynomial_generator import PolynomialGenerator
    from roboticmanipulators_b200.utils.trajectory_generator import TrajectoryGenerator
    g = golden("polynomial.npz")
    coef, q, dq, ddq = PolynomialGenerator.profiles_batch(dev(g["q0"]), dev(g["qf"]), dev(g["tf"]), dev(g["t0"]), 64)
    scale = lambda a: np.maximum(1.0, np.abs(a).max(axis=-1, keepdims=True))  # noqa: E731
    assert np.max(np.abs(coef.cpu().numpy() - g["coefs"]) / scale(g["coefs"])) < 1e-11
    assert np.max(np.abs(q.cpu().numpy() - g["q"]) / scale(g["q"])) < 1e-11
    assert np.max(np.abs(dq.cpu().numpy() - g["dq"]) / scale(g["dq"])) < 1e-11
    assert np.max(np.abs(ddq.cpu().numpy() - g["ddq"]) / scale(g["ddq"])) < 1e-11
    assert np.allclose(coef[0].cpu().numpy(), [0, 0, 0, 10, -15, 6], atol=1e-12)   # reference test
    tg = TrajectoryGenerator(1, 20, 1, 10)
    q_0 = np.array([[[0.5, 0, 0], [-0.6, 0, 0], [0, 0, 0]]])
    q_f = np.array([[[1.57, 0, 0], [0.5, 0, 0], [-2.0, 0, 0]]])
    qb, _, _ = tg.generate_joint_poly_trajectory_batch(dev(q_0), dev(q_f), 2.0, 0.0, 50)
    assert np.max(np.abs(qb[0].cpu().numpy() - g["planning_qs"])) < 1e-12
    assert np.max(np.abs(tg.generate_joint_poly_trajectory(q_0[0], q_f[0], 2, n=50, plot=False) - g["planning_qs"])) < 1e-12


def test_p2p_large_golden_bit_exact(golden):
    """The same 480 real-reference moves (tests/golden/trajectories_large.npz) through the batched
    kernel, one launch per (n, control_freq) group: counts, times, q / dq / ddq bit-identical, the
    moves the reference raises on come back with length 0 and zero rows."""
    from roboticmanipulators_b200.robots import RRRRobot
    from roboticmanipulators_b200.utils.trajectory_generator import TrajectoryGenerator
    g = golden("trajectories_large.npz")
    for gi in range(int(g["groups"])):
        n, cf = (int(v) for v in g[f"g{gi}_spec"])
        tg = TrajectoryGenerator(1, 20, 1, 10, control_freq=cf)
        q, dq, ddq, t, L = tg.generate_p2p_trajectory_batch(dev(g[f"g{gi}_q0"]), dev(g[f"g{gi}_qf"]), n)
        assert np.array_equal(L.cpu().numpy(), g[f"g{gi}_len"]), (n, cf)
        for got, name in ((q, "q"), (dq, "dq"), (ddq, "ddq"), (t, "t")):
            assert np.array_equal(got.cpu().numpy(), g[f"g{gi}_{name}"]), (n, cf, name)
        # the fused FK variant walks the same profile
        q2, _, _, _, L2, T, J = tg.generate_p2p_trajectory_batch(dev(g[f"g{gi}_q0"]), dev(g[f"g{gi}_qf"]), n,
                                                                 robot=RRRRobot())
        assert torch.equal(q2, q) and torch.equal(L2, L)
