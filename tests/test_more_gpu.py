"""GPU parity, second helping: option flags, custom robot constants, constant tokens in the
generic chain, non-default streams, tool counts other than 3."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


def dev(x):
    return torch.tensor(np.asarray(x, dtype=np.float64), device="cuda")


def test_jacobian_user_angle_convention():
    from roboticmanipulators_b200.robots import Fanuc165F
    robot, c = Fanuc165F(), O.FanucConst()
    rng = np.random.default_rng(0)
    q = rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(2000, 6))
    _, Jm, dm = robot.forward_kinematics_batch(dev(q), jacobian=True, measure=True)
    _, Ju, du = robot.forward_kinematics_batch(dev(q), jacobian=True, measure=True, user_angles=True)
    dirs = torch.tensor(c.directions, device="cuda")
    assert torch.equal(Ju, Jm * dirs)                       # column i scaled by directions[i]
    assert torch.allclose(du, dm * float(np.prod(c.directions)), rtol=0, atol=0)
    # finite-difference check of d p / d q_user on a few poses
    h = 1e-6
    for i in range(3):
        for j in range(6):
            qp, qm = q[i].copy(), q[i].copy()
            qp[j] += h
            qm[j] -= h
            fd = (O.fanuc_fk(qp)[:3, 3] - O.fanuc_fk(qm)[:3, 3]) / (2 * h)
            assert np.max(np.abs(fd - Ju[i, :3, j].cpu().numpy())) < 1e-4


def test_custom_lengths_offsets_directions():
    from roboticmanipulators_b200.robots import Fanuc165F
    ls = (300.0, 350.0, 280.0, 900.0, 200.0, 1100.0, 180.0)
    off = np.array([0.1, -1.2, 1.4, 0.0, 0.2, -0.3])
    dirs = np.array([1, -1, 1, -1, 1, 1])
    robot = Fanuc165F(lengths=ls, offsets=off, directions=dirs)
    c = O.FanucConst(lengths=ls, offsets=off, directions=dirs)
    assert abs(float(robot.d) - c.d) == 0 and abs(float(robot.dq) - c.dq) == 0
    rng = np.random.default_rng(1)
    q = rng.uniform(-1.0, 1.0, size=(3000, 6))
    for layout in ("aos", "soa"):
        T, J, det = robot.forward_kinematics_batch(dev(q), jacobian=True, measure=True, layout=layout)
        T_ref, J_ref = O.fanuc_fk(q, c), O.fanuc_jacobian(q, c)
        assert np.max(np.abs(T.cpu().numpy() - T_ref) / np.maximum(1, np.abs(T_ref))) < 1e-12
        assert np.max(np.abs(J.cpu().numpy() - J_ref)) / np.abs(J_ref).max() < 1e-12
        scale = np.prod(np.linalg.norm(J_ref, axis=1), axis=1)
        assert np.all(np.abs(det.cpu().numpy() - np.linalg.det(J_ref)) <= 1e-12 * scale)
        qb, vb = robot.inverse_kinematics_batch(T, layout=layout)
        q_ref, v_ref = O.fanuc_ik_branches(T_ref, O.CANONICAL_8, c)
        assert np.array_equal(vb.cpu().numpy(), v_ref)
        d = qb.cpu().numpy() - q_ref
        d[..., [3, 5]] = (d[..., [3, 5]] + np.pi) % (2 * np.pi) - np.pi
        assert np.nanmax(np.abs(d)) < 1e-8


def test_chain_with_constant_tokens():
    from roboticmanipulators_b200.utils.robo_math import chain_batch
    rng = np.random.default_rng(2)
    tokens = O.tokens_from_sequence(O.FANUC_SEQ)
    ls = O.FanucConst().ls
    cst = [ls[0], 0, ls[1], ls[2], 0, ls[3], 0, ls[4], ls[5], 0, 0, 0, ls[6]]
    src = [-1, 0, -1, -1, 1, -1, 2, -1, -1, 3, 4, 5, -1]         # joint angles come from 6 columns
    qm = rng.uniform(-3, 3, size=(4000, 6))
    for layout in ("aos", "soa"):
        T, J = chain_batch(tokens, dev(qm), active=O.FANUC_IND, src=src, cst=cst, layout=layout)
        vals = O.fanuc_values(qm, ls)
        assert np.max(np.abs(T.cpu().numpy() - O.chain(tokens, vals))) / 3000 < 1e-13
        J_ref = O.jacobian_numerically(O.FANUC_SEQ, O.FANUC_IND, vals)
        assert np.max(np.abs(J.cpu().numpy() - J_ref)) / np.abs(J_ref).max() < 1e-13


def test_non_default_stream_and_reentrancy():
    from roboticmanipulators_b200.robots import Fanuc165F, RRRRobot
    robot, c = Fanuc165F(), O.FanucConst()
    rng = np.random.default_rng(3)
    q = rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(1 << 16, 6))
    streams = [torch.cuda.Stream() for _ in range(3)]
    outs = []
    qd = dev(q)
    torch.cuda.synchronize()
    for s in streams:
        with torch.cuda.stream(s):
            T, J, det = robot.forward_kinematics_batch(qd, jacobian=True, measure=True)
            qb, vb = robot.inverse_kinematics_batch(T)
            Tr = RRRRobot().forward_kinematics_batch(qd[:, :3].contiguous())
            outs.append((T, J, det, qb, vb, Tr))
    torch.cuda.synchronize()
    T_ref = O.fanuc_fk(q, c)
    for T, J, det, qb, vb, Tr in outs:
        assert np.max(np.abs(T.cpu().numpy() - T_ref) / np.maximum(1, np.abs(T_ref))) < 1e-12
        assert torch.equal(T, outs[0][0]) and torch.equal(qb.nan_to_num(), outs[0][3].nan_to_num())
        assert torch.equal(vb, outs[0][4])


def test_tool_base_other_tool_counts(golden):
    from test_host_logic_cpu import fanuc_calibrator
    from test_oracle_golden import _calib_model
    g = golden("calibration.npz")
    model = _calib_model(g)
    for ntools in (1, 2, 4):
        mc = fanuc_calibrator()
        mc.get_reduced_model()
        qs = g["qs"][:100]
        tools = [O.chain("TxTyTz", v) for v in ([400, -20, 60], [420, 55, -15], [410, -40, -45], [390, 10, 5])][:ntools]
        pts = np.concatenate([O.calib_pose(model, qs, np.zeros(18), g["T_base"], Tt)[:, :3, 3] for Tt in tools], axis=1)
        T_base, T_tools = mc.estimate_tool_base(qs, pts[:, :, None])
        A_sum, Adp = mc._last_tool_base_equations
        Ao, bo = O.calib_tool_base_equations(model, qs, pts, np.zeros(18))
        assert np.max(np.abs(A_sum - Ao)) / np.abs(Ao).max() < 1e-13
        assert np.max(np.abs(Adp - bo)) / np.abs(bo).max() < 1e-12
        assert len(T_tools) == ntools
        # the estimate is the oracle's (a small-angle linearisation: it does not reproduce the
        # generating base exactly, model_calibration.py:296-365)
        est_o, Tb_o, Tt_o, _, _ = O.calib_estimate_tool_base(model, qs, pts, np.zeros(18))
        assert np.max(np.abs(mc._base_est - est_o) / np.maximum(1, np.abs(est_o))) < 1e-8
        assert np.max(np.abs(np.array(T_base, dtype=float) - Tb_o)) < 1e-6
        ev = mc.evaluate(qs, pts[:, :, None], T_base, T_tools)
        evo = O.calib_evaluate(model, qs, pts, np.zeros(18), Tb_o, Tt_o)
        assert abs(ev["distance_rms"] - evo["distance_rms"]) < 1e-6


def test_ik_f32_with_base_and_tool(golden):
    from roboticmanipulators_b200.robots import Fanuc165F
    from test_fanuc_gpu import _check_ik_f32
    g = golden("fanuc_fk_ik.npz")
    robot = Fanuc165F(T_base=g["rotated_T_base"], T_tool=g["rotated_T_tool"])
    c = O.FanucConst(g["rotated_T_base"], g["rotated_T_tool"])
    rng = np.random.default_rng(5)
    q = rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(8192, 6))
    for layout in ("aos", "soa"):
        _check_ik_f32(robot, c, O.fanuc_fk(q, c), layout)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_no_out_of_range_writes(dtype, layout):
    """Ragged batch sizes (tiles cut short, TMA clipping, fallbacks for unaligned pitches), SOA
    planes with a padded leading dimension: every output sits inside a canary-filled buffer and the
    canaries must survive; results must equal the plain (unpadded) call."""
    import ctypes as C
    from roboticmanipulators_b200 import _lib
    from roboticmanipulators_b200.robots.fanuc import Fanuc165F
    lib = _lib.load()
    robot, c = Fanuc165F(), O.FanucConst()
    st = robot._struct()
    sfx = "f64" if dtype == torch.float64 else "f32"
    rng = np.random.default_rng(5)
    CAN = 12345.0
    for n, pad in ((1, 0), (31, 1), (32, 0), (33, 2), (100, 4), (1000, 0), (4097, 3), (4098, 6)):
        q = torch.tensor(rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(n, 6)), device="cuda", dtype=dtype)
        T_ref, J_ref, det_ref = robot.forward_kinematics_batch(q, jacobian=True, measure=True, layout=layout)
        q_ref, v_ref = robot.inverse_kinematics_batch(T_ref.contiguous(), layout=layout)
        ld = n + pad if layout == "soa" else n
        G = 64                                              # guard elements on both sides

        def guarded(ncomp, dt=dtype, fill=CAN):
            buf = torch.full((G + ncomp * ld + G,), fill, dtype=dt, device="cuda")
            return buf, buf[G:G + ncomp * ld]
        if layout == "soa":
            qs = torch.full((6, ld), CAN, dtype=dtype, device="cuda")
            qs[:, :n] = q.t()
        else:
            qs = q.contiguous()
        Tb, Tv = guarded(16)
        Jb, Jv = guarded(36)
        Db, Dv = guarded(1)
        flags = _lib.RK_SOA if layout == "soa" else _lib.RK_AOS
        _lib.check(getattr(lib, "rk_fanuc_fk_jac_" + sfx)(
            C.byref(st), _lib.ptr(qs), n, ld, flags, _lib.ptr(Tv), _lib.ptr(Jv), _lib.ptr(Dv), _lib.stream_ptr()))
        Qb, Qv = guarded(48)
        Vb, Vv = guarded(8, torch.uint8, 77)
        _lib.check(getattr(lib, "rk_fanuc_ik_" + sfx)(
            C.byref(st), _lib.ptr(Tv), n, ld, flags, None, 8, _lib.ptr(Qv), _lib.ptr(Vv), _lib.stream_ptr()))
        torch.cuda.synchronize()
        for buf, view, ncomp, can in ((Tb, Tv, 16, CAN), (Jb, Jv, 36, CAN), (Db, Dv, 1, CAN), (Qb, Qv, 48, CAN),
                                      (Vb, Vv, 8, 77)):
            assert bool((buf[:G] == can).all()) and bool((buf[-G:] == can).all()), (n, pad, ncomp)
            if layout == "soa" and pad:
                assert bool((view.view(ncomp, ld)[:, n:] == can).all()), (n, pad, ncomp)

        def logical(view, ncomp):
            return view.view(ncomp, ld)[:, :n].t() if layout == "soa" else view.view(n, ncomp)
        assert torch.equal(logical(Tv, 16), T_ref.reshape(n, 16))
        assert torch.equal(logical(Jv, 36), J_ref.reshape(n, 36))
        # the padded and the plain call may run different kernels (tile vs fallback): same masks,
        # angles equal to rounding (q3/q5 modulo 2 pi at the atan2 cut)
        assert torch.equal(logical(Vv, 8).bool(), v_ref.reshape(n, 8))
        dq = (logical(Qv, 48).reshape(n, 8, 6) - q_ref).double().cpu().numpy()
        dq[..., [3, 5]] = (dq[..., [3, 5]] + np.pi) % (2 * np.pi) - np.pi
        assert np.nanmax(np.abs(dq)) < (1e-9 if dtype == torch.float64 else 2e-2)


def test_custom_constants_against_reference_golden(golden):
    """Custom lengths / offsets / directions against vectors from the real reference
    (tests/golden/fanuc_custom.npz): FK, and all 12 IK branches with the same masks."""
    from roboticmanipulators_b200.robots import Fanuc165F
    g = golden("fanuc_custom.npz")
    robot = Fanuc165F(lengths=tuple(g["lengths"]), offsets=g["offsets"], directions=g["directions"])
    c_custom = O.FanucConst(lengths=tuple(g["lengths"]), offsets=g["offsets"], directions=g["directions"])
    assert float(robot.d) == float(g["d"]) and float(robot.dq) == float(g["dq"])
    Tg, ref = g["T"], g["ik12"]
    br = [tuple(int(v) for v in b) for b in g["branches12"]]
    for layout in ("aos", "soa"):
        T = robot.forward_kinematics_batch(dev(g["q"]), layout=layout).cpu().numpy()
        assert np.max(np.abs(T[:, :3, :3] - Tg[:, :3, :3])) < 1e-12
        assert np.max(np.abs(T[:, :3, 3] - Tg[:, :3, 3]) / np.maximum(1, np.abs(Tg[:, :3, 3]))) < 1e-12
        qg, vg = robot.inverse_kinematics_batch(dev(Tg), br, layout=layout)
        qg = qg.cpu().numpy()
        assert np.array_equal(np.isnan(qg), np.isnan(ref))
        import ikcond
        with np.errstate(all="ignore"):
            _, _, det = ikcond.ik_details(Tg, br, c_custom)
        tol, _ = ikcond.ik_tolerance(ref, det, base=2e-12)
        err = ikcond.ik_diff(qg, ref.copy())
        live = vg.cpu().numpy()[..., None] & ~np.isnan(ref)
        assert not np.any(live & (err > tol)), float(np.nanmax(np.where(live, err / tol, 0)))
        nz = lambda a: np.any(np.abs(np.nan_to_num(a, nan=1.0)) > 1e-12, axis=-1)  # noqa: E731
        assert np.array_equal(vg.cpu().numpy() & nz(qg), nz(ref))
        # the canonical 8 through the fused single call agree with the same vectors
        _, _, _, q8, v8 = robot.fk_jac_ik_batch(dev(g["q"]), layout=layout)
        q8 = q8.cpu().numpy()
        idx8 = [br.index(b) for b in O.CANONICAL_8]
        d8 = q8 - ref[:, idx8]
        d8[..., [3, 5]] = (d8[..., [3, 5]] + np.pi) % (2 * np.pi) - np.pi
        # IK of the pose the GPU computed (differs from the golden T by rounding): same masks except
        # where a joint sits within rounding of a limit, values to 1e-9
        same = v8.cpu().numpy() == nz(ref[:, idx8])
        assert same.mean() > 0.99
        assert np.nanmax(np.abs(d8[same & v8.cpu().numpy()])) < 1e-8


def test_soa_column_slice_of_wider_planes_is_safe():
    """Q[:, lo:hi].t() -- the per-rank shard of wider SOA planes -- has a plane
    pitch larger than its length.  The outputs are allocated dense, so the wrappers must not hand
    the input's pitch to the kernel as the output pitch: results equal those of a dense copy and
    nothing is written outside the result tensors (the caching allocator would place such writes in
    neighbouring live tensors, hence the sentinels around the calls)."""
    from roboticmanipulators_b200.robots import Fanuc165F, RRRRobot
    from roboticmanipulators_b200.utils.robo_math import Transformation
    robot, c = Fanuc165F(), O.FanucConst()
    rng = np.random.default_rng(11)
    N, lo, hi = 5000, 1237, 3290
    Q = dev(rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(N, 6)).T.copy())      # [6, N] planes
    view = Q[:, lo:hi].t()                                                      # [n, 6], pitch N > n
    dense = view.contiguous()
    sentinels = [torch.full((1 << 16,), 3.25, dtype=torch.float64, device="cuda") for _ in range(4)]
    T, J, det, qik, val = robot.fk_jac_ik_batch(view)
    T2, J2, det2 = robot.forward_kinematics_batch(view, jacobian=True, measure=True)
    q3, v3 = robot.inverse_kinematics_batch(T2)
    more = [torch.full((1 << 16,), 3.25, dtype=torch.float64, device="cuda") for _ in range(4)]
    Td, Jd, detd, qikd, vald = robot.fk_jac_ik_batch(dense, layout="aos")
    torch.cuda.synchronize()
    for a, b in ((T, Td), (J, Jd), (det, detd), (T2, Td), (J2, Jd), (det2, detd)):
        assert a.shape == b.shape and (a - b).abs().max().item() < 1e-9
    assert torch.equal(val, vald) and torch.equal(v3, vald)
    assert all(bool((s == 3.25).all()) for s in sentinels + more)
    # the same pattern through the arms and the generic chain
    rrr = RRRRobot()
    Q3 = dev(rng.uniform(-3, 3, size=(3, N)))
    Ta, Ja = rrr.forward_kinematics_batch(Q3[:, lo:hi].t(), jacobian=True)
    Tb, Jb = rrr.forward_kinematics_batch(Q3[:, lo:hi].t().contiguous(), jacobian=True, layout="aos")
    assert (Ta - Tb).abs().max().item() < 1e-12 and (Ja - Jb).abs().max().item() < 1e-12
    V = dev(rng.uniform(-2, 2, size=(4, N)))
    Tc = Transformation.batch("RzTxRzTx", V[:, lo:hi].t())
    Tdn = Transformation.batch("RzTxRzTx", V[:, lo:hi].t().contiguous(), layout="aos")
    torch.cuda.synchronize()
    assert (Tc - Tdn).abs().max().item() < 1e-12
    assert all(bool((s == 3.25).all()) for s in sentinels + more)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_host_stream_matches_device_batch(dtype):
    """Fanuc165F.fk_jac_ik_stream (rk_fanuc_fk_jac_ik8_host_*): host arrays streamed in chunks over
    three streams give bit-identical results to the device-resident call, for ragged sizes, chunk
    sizes that do and do not divide n, pinned and pageable inputs, output selection and SOA planes."""
    from roboticmanipulators_b200.robots import Fanuc165F
    robot, c = Fanuc165F(), O.FanucConst()
    rng = np.random.default_rng(21)
    for n, chunk in ((1, 64), (33, 32), (1000, 256), (5000, 4096), (70001, 1 << 14)):
        q = torch.tensor(rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(n, 6))).to(dtype)
        ref = robot.fk_jac_ik_batch(q.cuda(), layout="aos")
        names = ("T", "J", "det", "q_ik", "valid")
        for pinned in (True, False):
            hq = q.pin_memory() if pinned else q.numpy()
            res = robot.fk_jac_ik_stream(hq, chunk=chunk)
            for name, r in zip(names, ref):
                assert not res[name].is_cuda
                assert torch.equal(res[name], r.cpu()), (n, chunk, name)
        # output selection: only the IK answers travel; preallocated result buffers are filled in place
        outs = {"q_ik": torch.full((n, 8, 6), -5.0, dtype=dtype).pin_memory(),
                "valid": torch.zeros((n, 8), dtype=torch.bool).pin_memory()}
        res = robot.fk_jac_ik_stream(q.pin_memory(), want=("q_ik", "valid"), out=outs, chunk=chunk)
        assert set(res) == {"q_ik", "valid"} and res["q_ik"].data_ptr() == outs["q_ik"].data_ptr()
        assert torch.equal(res["q_ik"], ref[3].cpu()) and torch.equal(res["valid"], ref[4].cpu())
        res = robot.fk_jac_ik_stream(q.pin_memory(), want=("J", "det"), chunk=chunk)
        assert torch.equal(res["J"], ref[1].cpu()) and torch.equal(res["det"], ref[2].cpu())
        # component-major host planes
        planes = q.t().contiguous()
        ref_soa = robot.fk_jac_ik_batch(planes.cuda().t(), layout="soa")
        res = robot.fk_jac_ik_stream(planes.t(), layout="soa", chunk=chunk)
        for name, r in zip(names, ref_soa):
            # (the staged planes have the chunk's pitch, the reference call the batch's: for tiny or
            # odd n one of them takes the fused tile kernel and the other the fallback pair, which
            # agree to rounding, not bit for bit)
            assert res[name].shape == r.shape
            if name == "valid":
                assert torch.equal(res[name], r.cpu()), (n, chunk, name, "soa")
                continue
            d = (res[name].double() - r.cpu().double()).numpy()
            if name == "q_ik":
                d[..., [0, 3, 5]] = (d[..., [0, 3, 5]] + np.pi) % (2 * np.pi) - np.pi
            scale = np.maximum(1.0, np.abs(r.cpu().double().numpy()))
            assert np.max(np.abs(d) / scale) < (1e-9 if dtype == torch.float64 else 2e-3), (n, chunk, name)
    with pytest.raises(Exception):
        robot.fk_jac_ik_stream(q.cuda())


def test_concurrent_host_threads_and_streams():
    """Four host threads, each on its own CUDA stream, call the library at the same time (Fanuc fused
    call, a generic chain -- first use compiles its kernel under the cache lock --, the calibration
    evaluation with its per-(device, stream) workspace); every thread must get the serial answers."""
    import threading
    from roboticmanipulators_b200.robots.fanuc import Fanuc165F
    from roboticmanipulators_b200.utils.jacobians import JacobianCalculator
    rng = np.random.default_rng(41)
    c = O.FanucConst()
    robot = Fanuc165F()
    n = 30000
    qs = [torch.tensor(rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(n, 6)), device="cuda") for _ in range(4)]
    seq, ind = "RzTyRxTzRyTx", [1, 0, 1, 1, 1, 0]          # a structure no other test has compiled yet
    vals = [torch.tensor(rng.uniform(-2, 2, size=(n, 6)), device="cuda") for _ in range(4)]
    # one calibrator shared by all threads: its workspace (ticket counter + CTA partials) is per stream
    from test_host_logic_cpu import fanuc_calibrator
    from test_oracle_golden import T_BASE_0, T_TOOLS_0
    mc = fanuc_calibrator()
    mc.get_reduced_model()
    m = 200000
    cq = [torch.tensor(rng.uniform(-1.0, 1.0, size=(m, 6)), device="cuda") for _ in range(4)]
    cp = [torch.tensor(rng.uniform(-1500, 1500, size=(m, 9)), device="cuda") for _ in range(4)]
    serial = []
    for k in range(4):
        sums = mc.residual_sums(cq[k], cp[k], T_BASE_0, T_TOOLS_0, process_group=False)
        serial.append((robot.fk_jac_ik_batch(qs[k]), sums.clone()))
    torch.cuda.synchronize()
    results, errors = [None] * 4, []

    def work(k):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                out = None
                for _ in range(5):
                    out = robot.fk_jac_ik_batch(qs[k])
                jc = JacobianCalculator(seq, ind)
                T, J = jc.calculate_numerically_batch(vals[k], with_pose=True)
                dq, ok = jc.dq_from_dx_batch(vals[k], vals[k][:, :4], rows=(0, 1, 2, 4))
                sums = None
                for _ in range(5):
                    sums = mc.residual_sums(cq[k], cp[k], T_BASE_0, T_TOOLS_0, process_group=False).clone()
            st.synchronize()
            results[k] = (out, (T, J, dq, ok), sums)
        except Exception as e:                               # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(4):
        for a, b in zip(results[k][0], serial[k][0]):
            assert torch.equal(a, b)
        jc = JacobianCalculator(seq, ind)
        T, J = jc.calculate_numerically_batch(vals[k], with_pose=True)
        dq, ok = jc.dq_from_dx_batch(vals[k], vals[k][:, :4], rows=(0, 1, 2, 4))
        for a, b in zip(results[k][1], (T, J, dq, ok)):
            assert torch.equal(a, b)
        assert torch.equal(results[k][2], serial[k][1])


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_ik_host_stream_matches_device_batch(dtype):
    """Fanuc165F.inverse_kinematics_stream (rk_fanuc_ik8_host_*): targets on the host, streamed in
    chunks, give the bits of inverse_kinematics_batch on the device -- ragged sizes, chunk sizes that
    do not divide n, pinned and pageable inputs, preallocated outputs, unreachable targets."""
    from roboticmanipulators_b200.robots import Fanuc165F
    robot, c = Fanuc165F(), O.FanucConst()
    rng = np.random.default_rng(33)
    for n, chunk in ((1, 64), (33, 32), (1000, 256), (70001, 1 << 14)):
        q = torch.tensor(rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(n, 6))).to(dtype)
        T = robot.forward_kinematics_batch(q.cuda(), layout="aos")
        T[::7, :3, 3] *= 3.0                                  # mostly out of reach: all-zero answers, valid = 0
        q_ref, v_ref = robot.inverse_kinematics_batch(T, layout="aos")
        for hT in (T.cpu().pin_memory(), T.cpu().numpy()):
            q_ik, valid = robot.inverse_kinematics_stream(hT, chunk=chunk)
            assert not q_ik.is_cuda and q_ik.shape == (n, 8, 6) and valid.dtype == torch.bool
            assert torch.equal(q_ik, q_ref.cpu()) and torch.equal(valid, v_ref.cpu()), (n, chunk)
        outs = {"q_ik": torch.full((n, 8, 6), -5.0, dtype=dtype).pin_memory(),
                "valid": torch.ones((n, 8), dtype=torch.bool).pin_memory()}
        q_ik, valid = robot.inverse_kinematics_stream(T.cpu().pin_memory(), out=outs, chunk=chunk)
        assert q_ik.data_ptr() == outs["q_ik"].data_ptr()
        assert torch.equal(q_ik, q_ref.cpu()) and torch.equal(valid, v_ref.cpu())
        assert float(valid[::7].float().mean()) < 0.5 * max(float(valid.float().mean()), 1e-9) or n < 100
    with pytest.raises(Exception):
        robot.inverse_kinematics_stream(T)                    # a CUDA tensor is not a host array
