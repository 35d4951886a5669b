"""GPU parity: RRR / RR robots, generic chain interpreter, JacobianCalculator."""
import numpy as np
import pytest
import sympy as sp
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

TB = np.array([[0, -1, 0, 0.3], [1, 0, 0, -0.2], [0, 0, 1, 0.5], [0, 0, 0, 1.0]])
TT = O.chain("TxRyTz", [0.2, 0.3, -0.1])


def dev(x):
    return torch.tensor(np.asarray(x, dtype=np.float64), device="cuda")


@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_rrr_fk_jac_ik(golden, layout):
    from roboticmanipulators_b200.robots import RRRRobot
    rng = np.random.default_rng(0)
    q = rng.uniform(-np.pi, np.pi, size=(10000, 3))
    for robot, kw in ((RRRRobot(), {}), (RRRRobot(T_base=TB, T_tool=TT), dict(T_base=TB, T_tool=TT))):
        T, J = robot.forward_kinematics_batch(dev(q), jacobian=True, layout=layout)
        T_ref, J_ref = O.rrr_fk(q, **kw), O.rrr_jacobian(q, **kw)
        assert np.max(np.abs(T.cpu().numpy() - T_ref)) < 1e-12
        assert np.max(np.abs(J.cpu().numpy() - J_ref)) < 1e-12
        qb, vb = robot.inverse_kinematics_batch(dev(T_ref), layout=layout)
        for b, (m, k) in enumerate(((1, 1), (1, -1), (-1, 1), (-1, -1))):
            q_ref, v_ref = O.rrr_ik(T_ref, m, k, **kw)
            assert np.array_equal(vb[:, b].cpu().numpy(), v_ref)
            assert np.max(np.abs(qb[:, b].cpu().numpy() - q_ref)) < 1e-9
            assert np.percentile(np.abs(qb[:, b].cpu().numpy() - q_ref), 99) < 1e-12


def test_rrr_golden_and_scalar_api(golden):
    from roboticmanipulators_b200.robots import RRRRobot
    g = golden("rrr_rr.npz")
    robot = RRRRobot()
    T = robot.forward_kinematics_batch(dev(g["rrr_q"]))
    assert np.max(np.abs(T.cpu().numpy() - g["rrr_T"])) < 1e-13
    qb, _ = robot.inverse_kinematics_batch(dev(g["rrr_T"]), [tuple(b) for b in g["rrr_branches"]])
    assert np.max(np.abs(qb.cpu().numpy() - g["rrr_ik4"])) < 1e-9
    # SURVEY G5
    T = robot.forward_kinematics([np.pi / 2, 0, 0.2], plot=False)
    assert np.allclose(T[:3, 3], [0, 1.980066577841, 0.801330669205], atol=1e-11)
    assert np.allclose(robot.inverse_kinematics(T), [np.pi / 2, 0, 0.2], atol=1e-12)
    assert np.allclose(robot.inverse_kinematics(T, m=-1, k=1), [np.pi / 2, 0.2, -0.2], atol=1e-12)
    far = np.eye(4)
    far[0, 3] = 5.0
    assert not robot.inverse_kinematics(far).any()
    pts = np.array([[1.0, 0.0, 2.0], [0.5, 0.5, 1.2]])
    qs = robot.move_via_points(pts)
    assert qs.shape == (3, 2)
    Ts = robot.move_joints(qs.T)
    assert np.allclose(np.array(Ts)[:, :3, 3], pts, atol=1e-12)


@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_rr_fk_jac_ik(golden, layout):
    from roboticmanipulators_b200.robots import RRRobot
    g = golden("rrr_rr.npz")
    rng = np.random.default_rng(1)
    q = rng.uniform(-np.pi, np.pi, size=(5000, 2))
    for ls, kw in (((0.8, 0.8), {}), ((0.5, 1.25), dict(T_base=TB, T_tool=TT))):
        robot = RRRobot(lengths=ls, save=False, **kw)
        T, J = robot.forward_kinematics_batch(dev(q), jacobian=True, layout=layout)
        T_ref, J_ref = O.rr_fk(q, ls=ls, **kw), O.rr_jacobian(q, ls=ls, **kw)
        assert np.max(np.abs(T.cpu().numpy() - T_ref)) < 1e-12
        assert np.max(np.abs(J.cpu().numpy() - J_ref)) < 1e-12
        qb, vb = robot.inverse_kinematics_batch(dev(T_ref), layout=layout)
        for b, m in enumerate((1, -1)):
            q_ref, v_ref = O.rr_ik(T_ref, m, ls=ls, **kw)
            assert np.array_equal(vb[:, b].cpu().numpy(), v_ref)
            assert np.percentile(np.abs(qb[:, b].cpu().numpy() - q_ref), 99) < 1e-12
            assert np.max(np.abs(qb[:, b].cpu().numpy() - q_ref)) < 1e-7
    robot = RRRobot(lengths=(0.5, 1.25), save=False)
    T = robot.forward_kinematics_batch(dev(g["rr_l_q"]))
    assert np.max(np.abs(T.cpu().numpy() - g["rr_l_T"])) < 1e-13
    qb, _ = robot.inverse_kinematics_batch(dev(g["rr_l_T"]))
    assert np.max(np.abs(qb.cpu().numpy() - g["rr_l_ik2"])) < 1e-8
    assert np.allclose(robot.inverse_kinematics(g["rr_l_T"][0], m=1), g["rr_l_ik2"][0, 0], atol=1e-9)


def test_generic_chain_golden(golden):
    from roboticmanipulators_b200.utils.jacobians import JacobianCalculator
    from roboticmanipulators_b200.utils.robo_math import Transformation
    g = golden("generic_chain.npz")
    for c in range(int(g["n_cases"])):
        seq, ind = str(g[f"c{c}_seq"]), [int(i) for i in g[f"c{c}_ind"]]
        kw = dict(T_base=g["T_base"], T_tool=g["T_tool"]) if bool(g[f"c{c}_bt"]) else {}
        vals = g[f"c{c}_vals"]
        for layout in ("aos", "soa"):
            T = Transformation.batch(seq, dev(vals), layout=layout)
            assert np.max(np.abs(T.cpu().numpy() - g[f"c{c}_T"])) < 1e-13
        if any(t.endswith("i") and a for t, a in zip(O.tokens_from_sequence(seq), ind)):
            with pytest.raises(ValueError):
                JacobianCalculator(seq, ind, **kw).calculate_numerically_batch(dev(vals))
            continue
        jc = JacobianCalculator(seq, ind, **kw)
        for layout in ("aos", "soa"):
            J = jc.calculate_numerically_batch(dev(vals), layout=layout)
            ref = g[f"c{c}_J"]
            assert np.max(np.abs(J.cpu().numpy() - ref)) / max(1, np.abs(ref).max()) < 1e-13
    jc = JacobianCalculator("RzTxRzTx", [1, 0, 1, 0], simplify=False)       # tests/test_jacobians.py:59-75
    J = np.array(jc.calculate_numerically((0.0, 1.0, 0.0, 2.0)), dtype=float)
    assert np.allclose(J, [[0, 0], [3, 2], [0, 0], [0, 0], [0, 0], [1, 1]])
    with pytest.raises(ValueError):
        jc.calculate_numerically((0.0, 1.0))


def test_generic_chain_large_vs_oracle():
    from roboticmanipulators_b200.utils.jacobians import JacobianCalculator
    rng = np.random.default_rng(5)
    seq, ind = O.FANUC_SEQ, O.FANUC_IND
    vals = rng.uniform(-3, 3, size=(20000, 13))
    jc = JacobianCalculator(seq, ind, T_base=TB, T_tool=TT)
    T, J = jc.calculate_numerically_batch(dev(vals), with_pose=True)
    T_ref = TB @ O.chain(seq, vals) @ TT
    J_ref = O.jacobian_numerically(seq, ind, vals, TB, TT)
    assert np.max(np.abs(T.cpu().numpy() - T_ref)) / np.abs(T_ref).max() < 1e-13
    assert np.max(np.abs(J.cpu().numpy() - J_ref)) / np.abs(J_ref).max() < 1e-13
    # the specialised Fanuc kernel and the generic interpreter agree
    from roboticmanipulators_b200.robots import Fanuc165F
    robot = Fanuc165F(offsets=np.zeros(6), directions=np.ones(6))
    qm = rng.uniform(-3, 3, size=(5000, 6))
    Tf, Jf, _ = robot.forward_kinematics_batch(dev(qm), jacobian=True, measure=True)
    jc2 = JacobianCalculator(seq, ind)
    Tg, Jg = jc2.calculate_numerically_batch(dev(O.fanuc_values(qm, O.FanucConst().ls)), with_pose=True)
    assert (Tf - Tg).abs().max().item() < 1e-11
    assert (Jf - Jg).abs().max().item() < 1e-10


@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_arm_ik_special_targets(layout):
    """Targets on the special cases of rrr_robot.py:137-201 / rr_robot.py:147-195: on the base axis
    ("up position", q0 := 0), at and just beyond the reach limits (|arccos argument| = 1), the
    elbow fully folded, far away, the origin."""
    from roboticmanipulators_b200.robots import RRRRobot, RRRobot
    rrr, rr = RRRRobot(), RRRobot()
    l0, l1, l2 = [float(v) for v in rrr.ls]
    pts = [(0.0, 0.0, l0 + 0.7), (0.0, 0.0, l0 - 0.4), (1e-7, -1e-7, l0 + 1.0), (3e-6, 4e-6, l0 + 0.3),
           (l1 + l2, 0.0, l0), (0.0, l1 + l2, l0), ((l1 + l2) * (1 + 1e-15), 0.0, l0),
           ((l1 + l2) * (1 - 1e-15), 0.0, l0), (l1 + l2 + 1e-9, 0.0, l0), (abs(l1 - l2) + 1e-12, 0.0, l0),
           (0.0, 0.0, l0 + l1 + l2), (0.0, 0.0, l0), (5.0, 5.0, 5.0), (0.0, 0.0, 0.0), (-0.7, 0.0, l0 + 0.2),
           (0.3, -0.9, 0.1), (1.2, 0.4, 1.9)]
    T = np.tile(np.eye(4), (len(pts), 1, 1))
    T[:, :3, 3] = pts
    T = np.tile(T, (8, 1, 1))                          # more than one warp
    qb, vb = rrr.inverse_kinematics_batch(dev(T), layout=layout)
    with np.errstate(all="ignore"):
        for b, (m, k) in enumerate(((1, 1), (1, -1), (-1, 1), (-1, -1))):
            q_ref, v_ref = O.rrr_ik(T, m, k)
            assert np.array_equal(vb[:, b].cpu().numpy(), v_ref), (m, k)
            # the wrist folded onto the shoulder (l1 == l2, arccos argument exactly -1) leaves q1
            # arbitrary: the reference's value there is atan2(l2 sin(pi), 0) = atan2(1.2e-16, 0)
            folded = np.linalg.norm(T[:, :3, 3] - np.array([0.0, 0.0, l0]), axis=1) < 1e-6
            d = np.abs(qb[:, b].cpu().numpy() - q_ref)[~folded]
            assert np.nanmax(d) < 1e-6, (m, k)                                       # acos at +-1
    a0, a1 = [float(v) for v in rr._lengths()][:2]
    pts2 = [(a0 + a1, 0.0), (0.0, a0 + a1), ((a0 + a1) * (1 + 1e-15), 0.0), (a0 + a1 + 1e-9, 0.0),
            (abs(a0 - a1) + 1e-13, 0.0), (0.0, 0.0), (0.5, -0.3), (-1.1, 0.2), (9.0, 9.0), (1e-9, 1e-9)]
    T2 = np.tile(np.eye(4), (len(pts2), 1, 1))
    T2[:, :2, 3] = pts2
    T2 = np.tile(T2, (8, 1, 1))
    qb, vb = rr.inverse_kinematics_batch(dev(T2), layout=layout)
    with np.errstate(all="ignore"):
        for b, m in enumerate((1, -1)):
            q_ref, v_ref = O.rr_ik(T2, m, ls=(a0, a1))
            assert np.array_equal(vb[:, b].cpu().numpy(), v_ref), m
            folded = np.linalg.norm(T2[:, :2, 3], axis=1) < 1e-6                     # same corner
            assert np.nanmax(np.abs(qb[:, b].cpu().numpy() - q_ref)[~folded]) < 1e-6, m


@pytest.mark.parametrize("name", ["arms_base_tool.npz", "arms_base_tool_large.npz"])
def test_arms_base_tool_against_reference_golden(golden, name):
    """RRR / RR behind rotating T_base and T_tool against vectors from the real reference (57 joint
    vectors per arm, and 3001 from the same recipe)."""
    from roboticmanipulators_b200.robots import RRRRobot, RRRobot
    g = golden(name)
    Tb, Tt = g["T_base"], g["T_tool"]
    rrr = RRRRobot(T_base=Tb, T_tool=Tt)
    rr = RRRobot(T_base=Tb, T_tool=Tt, lengths=tuple(g["rr_lengths"]), save=False)
    for layout in ("aos", "soa"):
        T = rrr.forward_kinematics_batch(dev(g["rrr_q"]), layout=layout)
        assert np.max(np.abs(T.cpu().numpy() - g["rrr_T"])) < 1e-13
        qb, _ = rrr.inverse_kinematics_batch(dev(g["rrr_T"]), [tuple(b) for b in g["rrr_branches"]],
                                             layout=layout)
        assert np.max(np.abs(qb.cpu().numpy() - g["rrr_ik4"])) < 1e-9
        T = rr.forward_kinematics_batch(dev(g["rr_q"]), layout=layout)
        assert np.max(np.abs(T.cpu().numpy() - g["rr_T"])) < 1e-13
        qb, _ = rr.inverse_kinematics_batch(dev(g["rr_T"]), layout=layout)
        assert np.max(np.abs(qb.cpu().numpy() - g["rr_ik2"])) < 1e-9


def test_random_chains_against_the_real_reference(golden):
    """28 random token sequences from the REAL reference (tests/golden/generic_chain_random.npz): every
    structure compiles its own specialised kernel; product and Jacobian in both layouts, FP64 to
    1e-13 scaled and FP32 to 1e-5 scaled."""
    from roboticmanipulators_b200 import _lib
    from roboticmanipulators_b200.utils.jacobians import JacobianCalculator
    from roboticmanipulators_b200.utils.robo_math import Transformation
    g = golden("generic_chain_random.npz")
    for c in range(int(g["n_cases"])):
        seq, ind = str(g[f"c{c}_seq"]), [int(i) for i in g[f"c{c}_ind"]]
        kw = dict(T_base=g["T_base"], T_tool=g["T_tool"]) if bool(g[f"c{c}_bt"]) else {}
        vals, T_ref, J_ref = g[f"c{c}_vals"], g[f"c{c}_T"], g[f"c{c}_J"]
        jc = JacobianCalculator(seq, ind, **kw)
        for layout in ("aos", "soa"):
            T = Transformation.batch(seq, dev(vals), layout=layout)
            assert np.max(np.abs(T.cpu().numpy() - T_ref)) / max(1, np.abs(T_ref).max()) < 1e-13, seq
            J = jc.calculate_numerically_batch(dev(vals), layout=layout)
            assert np.max(np.abs(J.cpu().numpy() - J_ref)) / max(1, np.abs(J_ref).max()) < 1e-13, seq
        J32 = jc.calculate_numerically_batch(dev(vals).float(), layout="soa")
        assert J32.dtype == torch.float32
        assert np.max(np.abs(J32.double().cpu().numpy() - J_ref)) / max(1, np.abs(J_ref).max()) < 1e-5, seq
    assert _lib.load().rk_chain_jit_status() == b"ok"
