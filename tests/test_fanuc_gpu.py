"""GPU parity: Fanuc165F kernels (through the C ABI) vs the NumPy oracle and the golden
vectors of the real reference."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
import ikcond

pytestmark = pytest.mark.gpu

TB = np.array([[1, 0, 0, 500], [0, 1, 0, 300], [0, 0, 1, -100], [0, 0, 0, 1.0]])
TT = np.array([[1, 0, 0, 100], [0, 1, 0, 0], [0, 0, 1, -300], [0, 0, 0, 1.0]])


def robots(golden):
    from roboticmanipulators_b200.robots.fanuc import Fanuc165F
    g = golden("fanuc_fk_ik.npz")
    return [
        ("default", Fanuc165F(), O.FanucConst()),
        ("translated", Fanuc165F(T_base=TB, T_tool=TT), O.FanucConst(TB, TT)),
        ("rotated", Fanuc165F(T_base=g["rotated_T_base"], T_tool=g["rotated_T_tool"]),
         O.FanucConst(g["rotated_T_base"], g["rotated_T_tool"])),
    ]


def rand_q(n, seed=0):
    c = O.FanucConst()
    rng = np.random.default_rng(seed)
    return rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(n, 6))


def scaled_err(a, b, scale):
    return float(np.max(np.abs(a - b) / scale))


def check_fk_jac(T, J, det, q, c, tol):
    T_ref = O.fanuc_fk(q, c)
    J_ref = O.fanuc_jacobian(q, c)
    # parity metric of SURVEY section 7: per-block scaled max error
    assert np.max(np.abs(T[:, :3, :3] - T_ref[:, :3, :3])) < tol
    pscale = np.maximum(1.0, np.abs(T_ref[:, :3, 3]).max(axis=1))[:, None]
    assert scaled_err(T[:, :3, 3], T_ref[:, :3, 3], pscale) < tol
    assert np.array_equal(T[:, 3], np.tile([0, 0, 0, 1.0], (len(q), 1)))
    vscale = np.maximum(1.0, np.abs(J_ref[:, :3]).max(axis=(1, 2)))[:, None, None]
    assert scaled_err(J[:, :3], J_ref[:, :3], vscale) < tol
    assert np.max(np.abs(J[:, 3:] - J_ref[:, 3:])) < tol
    det_ref = np.linalg.det(J_ref)
    colscale = np.prod(np.linalg.norm(J_ref, axis=1), axis=1)
    assert np.all(np.abs(det - det_ref) <= tol * np.maximum(colscale, 1.0))


@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_fk_jac_f64_parity(golden, layout):
    q = rand_q(10000)                       # BASELINE config 1: 1e4 uniform-in-limit configs
    for name, robot, c in robots(golden):
        qt = torch.tensor(q, device="cuda")
        T, J, det = robot.forward_kinematics_batch(qt, jacobian=True, measure=True, layout=layout)
        assert T.shape == (10000, 4, 4) and J.shape == (10000, 6, 6)
        check_fk_jac(T.cpu().numpy(), J.cpu().numpy(), det.cpu().numpy(), q, c, 1e-12)


def test_fk_ragged_and_empty(golden):
    _, robot, c = robots(golden)[2]
    for n in (1, 2, 31, 33, 127, 129, 1000):
        q = rand_q(n, seed=n)
        for layout in ("aos", "soa"):
            T, J, det = robot.forward_kinematics_batch(torch.tensor(q, device="cuda"),
                                                       jacobian=True, measure=True, layout=layout)
            check_fk_jac(T.cpu().numpy(), J.cpu().numpy(), det.cpu().numpy(), q, c, 1e-12)
    T = robot.forward_kinematics_batch(torch.empty((0, 6), dtype=torch.float64, device="cuda"))
    assert T.shape == (0, 4, 4)


def test_fk_soa_view_is_zero_copy(golden):
    _, robot, c = robots(golden)[0]
    q = rand_q(4096, seed=3)
    store = torch.tensor(q.T.copy(), device="cuda")        # [6, N] planes
    T, J, det = robot.forward_kinematics_batch(store.t(), jacobian=True, measure=True)
    assert T.stride(0) == 1                                 # SOA result view
    check_fk_jac(T.cpu().numpy(), J.cpu().numpy(), det.cpu().numpy(), q, c, 1e-12)


def test_fk_jac_f32_parity(golden):
    q = rand_q(10000, seed=5)
    for name, robot, c in robots(golden):
        for layout in ("aos", "soa"):
            qt = torch.tensor(q, device="cuda", dtype=torch.float32)
            T, J, det = robot.forward_kinematics_batch(qt, jacobian=True, measure=True, layout=layout)
            assert T.dtype == torch.float32
            q32 = qt.cpu().numpy().astype(np.float64)
            T_ref, J_ref = O.fanuc_fk(q32, c), O.fanuc_jacobian(q32, c)
            T, J = T.cpu().numpy().astype(np.float64), J.cpu().numpy().astype(np.float64)
            # north_star: 1e-5 in FP32 (scaled per block)
            assert np.max(np.abs(T[:, :3, :3] - T_ref[:, :3, :3])) < 1e-5
            ps = np.maximum(1.0, np.abs(T_ref[:, :3, 3]).max(axis=1))[:, None]
            assert scaled_err(T[:, :3, 3], T_ref[:, :3, 3], ps) < 1e-5
            vs = np.maximum(1.0, np.abs(J_ref[:, :3]).max(axis=(1, 2)))[:, None, None]
            assert scaled_err(J[:, :3], J_ref[:, :3], vs) < 1e-5
            assert np.max(np.abs(J[:, 3:] - J_ref[:, 3:])) < 1e-5


def test_fk_golden_reference(golden):
    g = golden("fanuc_fk_ik.npz")
    for name, robot, c in robots(golden):
        q, Tg = g[f"{name}_q"], g[f"{name}_T"]
        T = robot.forward_kinematics_batch(torch.tensor(q, device="cuda")).cpu().numpy()
        assert np.max(np.abs(T[:, :3, :3] - Tg[:, :3, :3])) < 1e-12
        assert np.max(np.abs(T[:, :3, 3] - Tg[:, :3, 3]) / np.maximum(1, np.abs(Tg[:, :3, 3]))) < 1e-12


def test_jacobian_golden_and_singular_configs(golden):
    from roboticmanipulators_b200.robots.fanuc import Fanuc165F
    g = golden("fanuc_jacobian.npz")
    # goldens are in MODEL angles; a robot with zero offsets and unit directions maps q = q'
    robot = Fanuc165F(offsets=np.zeros(6), directions=np.ones(6))
    qm = g["qm"]
    J, det = robot.jacobian_batch(torch.tensor(qm, device="cuda"))
    J, det = J.cpu().numpy(), det.cpu().numpy()
    Jg = g["default_J"]
    assert np.max(np.abs(J[:, 3:] - Jg[:, 3:])) < 1e-12
    assert np.max(np.abs(J[:, :3] - Jg[:, :3]) / np.maximum(1, np.abs(Jg[:, :3]).max())) < 1e-12
    assert abs(det[0] - 270295653.8623566) / 270295653.8623566 < 1e-12
    scale = np.prod(np.linalg.norm(Jg[:6], axis=1), axis=1)
    assert np.all(np.abs(det[:6] - g["det_symbolic_first6"]) <= 1e-12 * scale)
    assert det[1] == 0.0                                                     # rotation singularity


def _cmp_ik(q, valid, q_ref, v_ref, tol=1e-12):
    assert np.array_equal(valid, v_ref)
    assert np.array_equal(np.isnan(q), np.isnan(q_ref))
    assert not np.any(q[~valid])                      # zeros sentinel on failure
    d = q - q_ref
    d[..., [3, 5]] = (d[..., [3, 5]] + np.pi) % (2 * np.pi) - np.pi   # atan2 cut at +-pi
    err = np.abs(d) / np.maximum(1.0, np.abs(q_ref))
    return float(np.nanmax(err)) if err.size else 0.0


@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_ik8_parity(golden, layout):
    """identical masks and branch order, values within 1e-12 max(1, |q|) -- every
    single value, with the tolerance growing only by the conditioning factor of tests/ikcond.py
    next to the elbow / wrist singularities."""
    q = rand_q(10000, seed=2)
    for name, robot, c in robots(golden):
        T = O.fanuc_fk(q, c)
        T[::100, :3, 3] *= 3.0                         # 1 % unreachable targets
        q_ref, v_ref, det = ikcond.ik_details(T, O.CANONICAL_8, c)
        qg, vg = robot.inverse_kinematics_batch(torch.tensor(T, device="cuda"), layout=layout)
        assert qg.shape == (10000, 8, 6) and vg.shape == (10000, 8)
        frac = ikcond.assert_ik_close(qg.cpu().numpy(), vg.cpu().numpy(), q_ref, v_ref, det, min_well=0.98)
        assert frac > 0.98, (name, frac)


def test_ik12_list_and_ragged(golden):
    from roboticmanipulators_b200.robots.fanuc import ALL_12
    for name, robot, c in robots(golden):
        for n in (1, 7, 8, 9, 250):
            q = rand_q(n, seed=10 + n)
            T = O.fanuc_fk(q, c)
            q_ref, v_ref, det = ikcond.ik_details(T, ALL_12, c)
            for layout in ("aos", "soa"):
                qg, vg = robot.inverse_kinematics_batch(torch.tensor(T, device="cuda"), ALL_12, layout=layout)
                ikcond.assert_ik_close(qg.cpu().numpy(), vg.cpu().numpy(), q_ref, v_ref, det, min_well=0.0)
            q8, v8 = robot.inverse_kinematics_batch(torch.tensor(T, device="cuda"))
            idx = [ALL_12.index(b) for b in O.CANONICAL_8]
            ikcond.assert_ik_close(q8.cpu().numpy(), v8.cpu().numpy(), q_ref[:, idx], v_ref[:, idx],
                                   {k: v[:, idx] for k, v in det.items()}, min_well=0.0)


def test_ik_golden_reference(golden):
    g = golden("fanuc_fk_ik.npz")
    br = [tuple(int(v) for v in b) for b in g["branches12"]]
    for name, robot, c in robots(golden):
        for key in ([name] if name != "default" else ["default", "extra"]):
            T, ref = g[f"{key}_T"], g[f"{key}_ik12"]
            qg, vg = robot.inverse_kinematics_batch(torch.tensor(T, device="cuda"), br)
            qg = qg.cpu().numpy()
            assert np.array_equal(np.isnan(qg), np.isnan(ref))
            # every value against the REAL reference's answer; the structured poses of this set sit
            # ON the wrist singularity / the r13 corner, hence the conditioned tolerance
            with np.errstate(all="ignore"):
                _, _, det = ikcond.ik_details(T, br, c)
            tol, _ = ikcond.ik_tolerance(ref, det, base=2e-12)
            err = ikcond.ik_diff(qg, ref.copy())
            live = vg.cpu().numpy()[..., None] & ~np.isnan(ref)
            assert not np.any(live & (err > tol)), float(np.nanmax(np.where(live, err / tol, 0)))
            # the reference's failure sentinel is the zero vector (rounding noise ~1e-17 counts as 0)
            nz = lambda a: np.any(np.abs(np.nan_to_num(a, nan=1.0)) > 1e-12, axis=-1)  # noqa: E731
            assert np.array_equal(vg.cpu().numpy() & nz(qg), nz(ref))


def _check_ik_f32(robot, c, T, layout):
    """FP32 bar (SURVEY section 7 / north_star): IDENTICAL masks for every (target, branch) whose
    threshold decisions are all >= 1e-3 away from flipping, and values within 1e-5 max(1, |q|)
    there (times the conditioning factor of tests/ikcond.py beyond cond = 10: FP32 carries 7 digits).  The oracle runs in FP64 on the SAME FP32-rounded
    poses, so the comparison isolates the kernel's arithmetic."""
    T32 = torch.tensor(T, device="cuda", dtype=torch.float32)
    qg, vg = robot.inverse_kinematics_batch(T32, layout=layout)
    qg, vg = qg.cpu().numpy().astype(np.float64), vg.cpu().numpy()
    T64 = T32.cpu().numpy().astype(np.float64)
    q_ref, v_ref, det = ikcond.ik_details(T64, O.CANONICAL_8, c)
    clear = ikcond.mask_margin(det, c, 1e-3)
    assert clear.mean() > 0.7
    assert np.array_equal(vg[clear], v_ref[clear]), int((vg[clear] != v_ref[clear]).sum())
    assert (vg == v_ref).mean() > 0.999
    assert not np.any(qg[~vg])
    tol, _ = ikcond.ik_tolerance(q_ref, det, base=1e-5, cond_ok=ikcond.COND_OK_F32)
    err = ikcond.ik_diff(qg, q_ref.copy())
    live = (clear & vg & v_ref)[..., None] & ~np.isnan(q_ref)
    assert live.sum() > 1000
    worst = float(np.max(np.where(live, err / tol, 0.0)))
    assert worst <= 1.0, worst


@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_ik_f32_masks(golden, layout):
    _, robot, c = robots(golden)[0]
    T = O.fanuc_fk(rand_q(20000, seed=4), c)
    T[::100, :3, 3] *= 3.0
    _check_ik_f32(robot, c, T, layout)


def test_scalar_api_matches_reference_tests(golden, capsys):
    """tests/test_robots.py of the reference, through the drop-in single-pose API."""
    from roboticmanipulators_b200.robots.fanuc import Fanuc165F
    robot = Fanuc165F()
    T = robot.forward_kinematics([0, 0, 0, 0, 0, 0], plot=False)
    assert np.allclose(T, [[1, 0, 0, 1807], [0, 1, 0, 0], [0, 0, 1, 1970], [0, 0, 0, 1]], atol=1e-9)
    T = robot.forward_kinematics([np.pi / 2, 0, 0, 0, 0, 0], plot=False)
    assert np.allclose(T, [[0, -1, 0, 0], [1, 0, 0, 1807], [0, 0, 1, 1970], [0, 0, 0, 1]], atol=1e-9)
    T = robot.forward_kinematics([1.570796, -0.28, 0.1, 0, 0.2, 0], plot=False)
    for m in (1, -1):
        qs = robot.inverse_kinematics(T, m=m)
        assert np.all(np.abs(robot.forward_kinematics(qs, plot=False) - T) <= robot.epsilon)
    assert not robot.inverse_kinematics(robot.forward_kinematics([0] * 6, plot=False), m=1).any()
    unreachable = np.array([[1, 0, 0, 100000], [0, 1, 0, 2882], [0, 0, 1, 895], [0, 0, 0, 1.0]])
    assert not robot.inverse_kinematics(unreachable, m=1).any()
    limit = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 100], [0, 0, 0, 1.0]])
    assert not robot.inverse_kinematics(limit, m=1).any()
    # fanuc_kinematics.py round trip with base/tool
    r2 = Fanuc165F(T_base=TB, T_tool=TT)
    qs = [-0.485, 0.5, -0.23, 0.23, 0.21, 0.213]
    T = r2.forward_kinematics(qs, plot=False)
    assert abs(T[0, 3] - 2281.1504233665910) < 1e-9
    assert np.allclose(r2.inverse_kinematics(T, m=-1, w=1), qs, atol=1e-12)
    assert not r2.inverse_kinematics(T, m=5, k=1, w=0).any()     # m defaults to 1 -> out of limits
    assert "[WARNING] m can only be -1 or 1" in capsys.readouterr().out


def test_large_batch_roundtrip_property():
    """Size-independent property at 2^22 poses: every valid k=+1 branch FK-round-trips."""
    from roboticmanipulators_b200.robots.fanuc import Fanuc165F
    robot = Fanuc165F()
    n = 1 << 22
    g = torch.Generator(device="cuda").manual_seed(1)
    lim = torch.tensor(np.asarray(robot.qs_lim_rad), device="cuda")
    store = torch.rand((6, n), generator=g, device="cuda", dtype=torch.float64)
    store = lim[:, :1] + store * (lim[:, 1:] - lim[:, :1])
    T = robot.forward_kinematics_batch(store.t())
    qb, vb = robot.inverse_kinematics_batch(T)
    assert 0.8 < vb.any(dim=1).double().mean().item() < 0.95
    for b in (0, 1, 4, 5):                                    # k = +1 branches
        ok = vb[:, b] & ~torch.isnan(qb[:, b]).any(dim=1)
        T2 = robot.forward_kinematics_batch(qb[:, b][ok].contiguous())
        err = (T2 - T[ok]).abs().amax(dim=(1, 2))
        assert err.median().item() < 1e-9
        # the only solutions that do not close the loop are the reference's wrist-singular
        # convention |r11 - 1| <= eps -> q3 = q4 = q5 = 0 (fanuc.py:283-287)
        bad = err > 1e-6
        assert bad.double().mean().item() < 0.01
        assert not qb[:, b][ok][bad][:, 3:].any()


def test_internal_atan2_accuracy():
    """The IK kernels' own FP64 atan2 vs libm over the whole plane, incl. axes and zeros."""
    import ctypes as C
    from roboticmanipulators_b200 import _lib
    rng = np.random.default_rng(0)
    n = 1 << 22
    y = rng.normal(size=n) * 10.0 ** rng.integers(-8, 8, size=n)
    x = rng.normal(size=n) * 10.0 ** rng.integers(-8, 8, size=n)
    special = np.array([0.0, -0.0, 1.0, -1.0, 1e-300, -1e-300, 1e300, np.inf, -np.inf, np.nan,
                        0.41421356237309503, 2.414213562373095, 1e-5])
    ys, xs = np.meshgrid(special, special)
    y[:ys.size], x[:xs.size] = ys.ravel(), xs.ravel()
    yt, xt = torch.tensor(y, device="cuda"), torch.tensor(x, device="cuda")
    out = torch.empty_like(yt)
    _lib.check(_lib.load().rk_probe_atan2(_lib.ptr(yt), _lib.ptr(xt), n, _lib.ptr(out), _lib.stream_ptr()))
    got, ref = out.cpu().numpy(), np.arctan2(y, x)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.array_equal(np.signbit(got[ok]), np.signbit(ref[ok]))
    err = np.abs(got[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-300)
    assert err.max() < 5e-16, err.max()                      # <= ~2 ulp


def test_internal_sincos_accuracy():
    """The kernels' own FP64 sincos vs libm: joint-angle range, large arguments, specials."""
    from roboticmanipulators_b200 import _lib
    rng = np.random.default_rng(1)
    n = 1 << 22
    x = np.concatenate([rng.uniform(-8, 8, n // 2), rng.uniform(-1e5, 1e5, n // 4),
                        rng.normal(size=n // 4) * 10.0 ** rng.integers(-12, 9, size=n // 4)])
    x[:12] = [0.0, -0.0, np.pi / 2, -np.pi / 2, np.pi, 1e5, -1e5, 1.0000001e5, 1e22, np.inf, -np.inf, np.nan]
    xt = torch.tensor(x, device="cuda")
    s, c = torch.empty_like(xt), torch.empty_like(xt)
    _lib.check(_lib.load().rk_probe_sincos(_lib.ptr(xt), n, _lib.ptr(s), _lib.ptr(c), _lib.stream_ptr()))
    s, c = s.cpu().numpy(), c.cpu().numpy()
    with np.errstate(invalid="ignore"):
        rs, rc = np.sin(x), np.cos(x)
    assert np.array_equal(np.isnan(s), np.isnan(rs)) and np.array_equal(np.isnan(c), np.isnan(rc))
    ok = ~np.isnan(rs)
    assert np.max(np.abs(s[ok] - rs[ok])) < 2.3e-16 and np.max(np.abs(c[ok] - rc[ok])) < 2.3e-16
    big = ok & (np.abs(rs) > 1e-3)
    assert np.max(np.abs(s[big] - rs[big]) / np.abs(rs[big])) < 4.5e-16
    big = ok & (np.abs(rc) > 1e-3)
    assert np.max(np.abs(c[big] - rc[big]) / np.abs(rc[big])) < 4.5e-16


def _special_poses():
    """Targets that hit every special case of fanuc.py:196-302: wrist singular (:272) exactly and
    within/around eps, |r13| <= eps (:279-282), wrist centre on the base axis (:253-255), exact
    zeros in R (axis-aligned tools), unreachable targets, zero and non-finite matrices."""
    c = O.FanucConst()
    rng = np.random.default_rng(7)
    off, dirs = np.asarray(c.offsets, float), np.asarray(c.directions, float)

    def user(qm):                                     # model angles -> user angles
        return (np.asarray(qm, float) - off) / dirs

    qs = []
    base = rng.uniform(-0.8, 0.8, size=(40, 6))
    for i, qm in enumerate(base):
        a = qm.copy()
        a[4] = [0.0, 1e-6, -3e-6, 4.4e-3, 4.6e-3, 1e-5, -2e-4, 3.2e-3][i % 8]   # q4' ~ 0: r11 ~ 1
        qs.append(user(a))
        b = qm.copy()
        b[5] = [np.pi / 2, -np.pi / 2, np.pi / 2 + 3e-6, np.pi / 2 - 8e-6][i % 4]   # r13 ~ 0
        qs.append(user(b))
        z = qm.copy()                                 # wrist centre on the base axis: rho3 = 0
        z[1] = rng.uniform(0.6, 1.2)
        rhs = -(c.ls[2] + c.ls[3] * np.cos(z[1])) / c.d
        q12 = c.dq + np.arccos(np.clip(rhs, -1, 1)) * (1 if i % 2 else -1)
        z[2] = q12 - z[1] + [0.0, 1e-9, -2e-9, 3e-7][i % 4]
        qs.append(user(z))
    T = O.fanuc_fk(np.array(qs), c)
    extra = []
    for ang in (0.0, np.pi / 2, np.pi, -np.pi / 2):   # axis-aligned tools: exact zeros in R
        for p in ((1500.0, 0.0, 1200.0), (0.0, 1400.0, 900.0), (1200.0, -700.0, 1500.0), (-900.0, 0.0, 800.0)):
            Tm = np.eye(4)
            ca, sa = np.round(np.cos(ang)), np.round(np.sin(ang))
            Tm[:3, :3] = [[ca, -sa, 0], [sa, ca, 0], [0, 0, 1]]
            Tm[:3, 3] = p
            extra.append(Tm)
            Tm2 = Tm.copy()
            Tm2[:3, :3] = Tm[:3, :3] @ np.array([[0, 0, 1.0], [0, 1, 0], [-1, 0, 0]])
            extra.append(Tm2)
    far = np.eye(4)
    far[:3, 3] = (1e5, -2e5, 3e5)
    onaxis = np.eye(4)
    onaxis[:3, 3] = (c.ls[6], 0.0, 2000.0)            # wrist centre exactly on the axis: atan2(0, 0)
    extra += [far, onaxis, np.zeros((4, 4))]
    bad = np.eye(4)
    bad[0, 3] = np.nan
    inf = np.eye(4)
    inf[1, 3] = np.inf
    extra += [bad, inf]
    return np.concatenate([T, np.array(extra)])


@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_ik8_special_poses(layout):
    """The thread-per-pose kernel defers measure-zero cases to the branch routine and handles
    the wrist singularity by a warp vote: both must reproduce the reference's answers."""
    from roboticmanipulators_b200.robots.fanuc import Fanuc165F
    c = O.FanucConst()
    T = _special_poses()
    reps = -(-4096 // len(T))
    Tb = np.tile(T, (reps, 1, 1))[:4096]              # even, > several warps: the TMA path in SOA
    with np.errstate(all="ignore"):
        q_ref, v_ref = O.fanuc_ik_branches(Tb, O.CANONICAL_8, c)
    qg, vg = Fanuc165F().inverse_kinematics_batch(torch.tensor(Tb, device="cuda"), layout=layout)
    qg, vg = qg.cpu().numpy(), vg.cpu().numpy()
    finite = np.isfinite(Tb).all(axis=(1, 2))
    assert np.array_equal(vg[finite], v_ref[finite])
    assert np.array_equal(np.isnan(qg[finite]), np.isnan(q_ref[finite]))
    assert not np.any(np.nan_to_num(qg[finite][~vg[finite]]))
    d = qg - q_ref
    d[..., [3, 5]] = (d[..., [3, 5]] + np.pi) % (2 * np.pi) - np.pi
    # the targets sit ON the ill-conditioned corners (q3, q5 amplify rounding by 1/|sin q4|, and
    # the reference itself divides by cos(q5) ~ 1e-5 at the r13 corner)
    ok = finite[:, None] & v_ref & ~np.isnan(q_ref).any(axis=-1)
    assert ok.sum() > 1000
    assert np.max(np.abs(d[..., :3][ok])) < 1e-9
    assert np.max(np.abs(d[..., 3:][ok])) < 1e-7


def test_ik8_masks_large_sample():
    """3e5 uniformly sampled poses (x 8 branches = 2.4e6 reachability + limit decisions): the
    validity masks of the thread-per-pose kernel must equal the oracle's exactly."""
    from roboticmanipulators_b200.robots.fanuc import Fanuc165F
    c = O.FanucConst()
    q = rand_q(300000, seed=99)
    T = O.fanuc_fk(q, c)
    T[::50, :3, 3] *= 1.6                              # 2 % far targets, some of them unreachable
    q_ref, v_ref, det = ikcond.ik_details(T, O.CANONICAL_8, c)
    Tt = torch.tensor(T, device="cuda")
    for layout in ("soa", "aos"):
        qg, vg = Fanuc165F().inverse_kinematics_batch(Tt, layout=layout)
        vg, qg = vg.cpu().numpy(), qg.cpu().numpy()
        assert np.array_equal(vg, v_ref), int((vg != v_ref).sum())
        frac = ikcond.assert_ik_close(qg, vg, q_ref, v_ref, det, min_well=0.98)
        assert frac > 0.98


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_fk_jac_ik8_single_call(golden, layout, dtype):
    """rk_fanuc_fk_jac_ik8 (one kernel for SOA): FK/J/det against the oracle, and the IK half
    bit-identical to rk_fanuc_ik applied to the pose it wrote (same per-pose code, the pose just
    stays in registers) -- ragged sizes included (spare lanes redo the last pose)."""
    f64 = dtype == torch.float64
    for n in (1, 31, 33, 1000, 10000):
        q = rand_q(n, seed=100 + n)
        for name, robot, c in robots(golden):
            qt = torch.tensor(q, device="cuda", dtype=dtype)
            if layout == "soa":
                qt = qt.t().contiguous().t()
            T, J, det, qik, val = robot.fk_jac_ik_batch(qt, layout=layout)
            assert T.shape == (n, 4, 4) and J.shape == (n, 6, 6) and qik.shape == (n, 8, 6)
            T2, J2, det2 = robot.forward_kinematics_batch(qt, jacobian=True, measure=True, layout=layout)
            if f64:
                check_fk_jac(T.cpu().numpy(), J.cpu().numpy(), det.cpu().numpy(), q, c, 1e-12)
            assert (T - T2).abs().max().item() <= (1e-12 if f64 else 2e-3) * 4000
            assert (J - J2).abs().max().item() <= (1e-12 if f64 else 2e-3) * 4000
            q2, v2 = robot.inverse_kinematics_batch(T, layout=layout)
            assert torch.equal(val, v2), name
            assert torch.equal(qik, q2), name
            if f64:
                q_ref, v_ref, det = ikcond.ik_details(T.cpu().numpy(), O.CANONICAL_8, c)
                ikcond.assert_ik_close(qik.cpu().numpy(), val.cpu().numpy(), q_ref, v_ref, det, min_well=0.0)


def test_fk_jac_ik8_no_out_of_range_writes():
    """The fused kernel stores without guards (ragged lanes redo the last pose): padding past n in
    every plane must stay untouched."""
    from roboticmanipulators_b200 import _lib
    from roboticmanipulators_b200.robots.fanuc import Fanuc165F
    import ctypes as C
    robot, lib = Fanuc165F(), _lib.load()
    for n, ld in ((1, 2), (33, 40), (1000, 1002), (4097, 4104)):
        q = torch.zeros((6, ld), dtype=torch.float64, device="cuda")
        q[:, :n] = torch.tensor(rand_q(n, seed=n).T, device="cuda")
        outs = [torch.full((r, ld), -7.0, dtype=torch.float64, device="cuda") for r in (16, 36, 1, 48)]
        valid = torch.full((8, ld), 9, dtype=torch.uint8, device="cuda")
        _lib.check(lib.rk_fanuc_fk_jac_ik8_f64(C.byref(robot._struct()), _lib.ptr(q), n, ld, _lib.RK_SOA,
                                               _lib.ptr(outs[0]), _lib.ptr(outs[1]), _lib.ptr(outs[2]),
                                               _lib.ptr(outs[3]), _lib.ptr(valid), _lib.stream_ptr()))
        torch.cuda.synchronize()
        for o in outs:
            assert torch.all(o[:, n:] == -7.0)
            assert not torch.any(o[:, :n] == -7.0)
        assert torch.all(valid[:, n:] == 9) and torch.all(valid[:, :n] <= 1)


def test_fk_jac_ik8_roundtrip_property_large():
    """2^22 poses through the single call: the IK of FK(q) contains q itself (some branch
    reproduces the joint vector) wherever a branch is valid, and masks match the two-call path."""
    from roboticmanipulators_b200.robots.fanuc import Fanuc165F
    robot = Fanuc165F()
    n = (1 << 22) + 5
    g = torch.Generator(device="cuda").manual_seed(3)
    lim = torch.tensor(np.asarray(robot.qs_lim_rad), device="cuda")
    store = torch.rand((6, n), generator=g, device="cuda", dtype=torch.float64)
    store = lim[:, :1] + store * (lim[:, 1:] - lim[:, :1])
    T, J, det, qik, val = robot.fk_jac_ik_batch(store.t())
    q2, v2 = robot.inverse_kinematics_batch(T)
    assert torch.equal(val, v2) and torch.equal(qik, q2)
    d = (qik - store.t()[:, None, :]).abs()
    d = torch.minimum(d, (d - 2 * np.pi).abs())           # q3/q5 are returned modulo 2 pi
    best = torch.where(val, d.amax(dim=2), torch.full_like(d[..., 0], np.inf)).amin(dim=1)
    found = best < 1e-6
    # the oracle finds q in 63 % of uniform samples (the rest: q0 = atan2(y, x) +- pi 'over the back'
    # configurations, which the reference's IK expresses through k = -1 with different q1, q2)
    assert 0.60 < found.double().mean().item() < 0.66
    assert best[found].median().item() < 1e-11


@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_config1_against_the_real_reference(golden, layout):
    """BASELINE config 1 at its stated size: 1e4 uniform-in-limits joint vectors (seed 0) whose FK,
    calculate_numerically Jacobian and 12-branch IK were computed by the REAL reference
    (tests/golden/fanuc_config1.npz, made by make_golden.py --only config1).  FK / J within 1e-12
    scaled, all 1.2e5 validity decisions identical, branch order identical, every IK value within
    its conditioned 1e-12 bound (2e-12 here: the reference's own rounding is part of the difference)."""
    from roboticmanipulators_b200.robots.fanuc import Fanuc165F, ALL_12
    g = golden("fanuc_config1.npz")
    q, T34, ik12, Jg = g["q"], g["T34"], g["ik12"], g["J"]
    n = len(q)
    assert n == 10000 and [tuple(int(v) for v in b) for b in g["branches12"]] == list(ALL_12)
    robot, c = Fanuc165F(), O.FanucConst()
    qt = torch.tensor(q, device="cuda")
    T, J, det = robot.forward_kinematics_batch(qt, jacobian=True, measure=True, layout=layout)
    T, J, det = T.cpu().numpy(), J.cpu().numpy(), det.cpu().numpy()
    assert np.max(np.abs(T[:, :3, :3] - T34[:, :, :3])) < 1e-12
    pscale = np.maximum(1.0, np.abs(T34[:, :, 3]).max(axis=1))[:, None]
    assert np.max(np.abs(T[:, :3, 3] - T34[:, :, 3]) / pscale) < 1e-12
    vscale = np.maximum(1.0, np.abs(Jg[:, :3]).max(axis=(1, 2)))[:, None, None]
    assert np.max(np.abs(J[:, :3] - Jg[:, :3]) / vscale) < 1e-12
    assert np.max(np.abs(J[:, 3:] - Jg[:, 3:])) < 1e-12
    colscale = np.prod(np.linalg.norm(Jg, axis=1), axis=1)
    assert np.all(np.abs(det - np.linalg.det(Jg)) <= 1e-12 * np.maximum(colscale, 1.0))
    # IK of the reference's own poses: all 12 (m, k, w) combinations
    Tg = np.zeros((n, 4, 4))
    Tg[:, :3], Tg[:, 3, 3] = T34, 1.0
    qik, vik = robot.inverse_kinematics_batch(torch.tensor(Tg, device="cuda"), ALL_12, layout=layout)
    qik, vik = qik.cpu().numpy(), vik.cpu().numpy()
    v_ref = np.any(ik12 != 0.0, axis=2)                # the reference's failure sentinel is the zero vector
    assert np.array_equal(vik, v_ref), int((vik != v_ref).sum())
    _, v_o, det_o = ikcond.ik_details(Tg, ALL_12, c)
    assert np.array_equal(v_o, v_ref)
    frac = ikcond.assert_ik_close(qik, vik, ik12, v_ref, det_o, base=2e-12, min_well=0.98)
    assert frac > 0.98
    # the canonical 8 through the fast kernels and through the fused single call on q
    idx8 = [list(ALL_12).index(b) for b in O.CANONICAL_8]
    det8 = {k: v[:, idx8] for k, v in det_o.items()}
    q8, v8 = robot.inverse_kinematics_batch(torch.tensor(Tg, device="cuda"), layout=layout)
    ikcond.assert_ik_close(q8.cpu().numpy(), v8.cpu().numpy(), ik12[:, idx8], v_ref[:, idx8], det8,
                           base=2e-12, min_well=0.98)


@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_general_robot_against_the_real_reference(golden, layout):
    """A robot with nothing at its default -- custom link lengths, joint offsets and directions,
    rotating base and tool -- on 4000 uniform-in-limits joint vectors whose FK and 12-branch IK came
    from the REAL reference (tests/golden/fanuc_general.npz, make_golden.py --only general): the
    kernels' general path (no sign-flip shortcuts for the user angles, T_base / T_tool products and
    their inverses) to the same gates as config 1, in FP64 and in FP32."""
    from roboticmanipulators_b200.robots.fanuc import Fanuc165F, ALL_12
    g = golden("fanuc_general.npz")
    q, T34, ik12 = g["q"], g["T34"], g["ik12"]
    n = len(q)
    robot = Fanuc165F(T_base=g["T_base"], T_tool=g["T_tool"], lengths=tuple(g["lengths"]), offsets=g["offsets"],
                      directions=g["directions"])
    c = O.FanucConst(T_base=g["T_base"], T_tool=g["T_tool"], lengths=g["lengths"], offsets=g["offsets"],
                     directions=g["directions"])
    T = robot.forward_kinematics_batch(torch.tensor(q, device="cuda"), layout=layout).cpu().numpy()
    assert np.max(np.abs(T[:, :3, :3] - T34[:, :, :3])) < 1e-12
    pscale = np.maximum(1.0, np.abs(T34[:, :, 3]).max(axis=1))[:, None]
    assert np.max(np.abs(T[:, :3, 3] - T34[:, :, 3]) / pscale) < 1e-12
    Tg = np.zeros((n, 4, 4))
    Tg[:, :3], Tg[:, 3, 3] = T34, 1.0
    v_ref = np.any(ik12 != 0.0, axis=2)
    _, v_o, det_o = ikcond.ik_details(Tg, ALL_12, c)
    assert np.array_equal(v_o, v_ref)
    qik, vik = robot.inverse_kinematics_batch(torch.tensor(Tg, device="cuda"), ALL_12, layout=layout)
    qik, vik = qik.cpu().numpy(), vik.cpu().numpy()
    assert np.array_equal(vik, v_ref), int((vik != v_ref).sum())
    assert ikcond.assert_ik_close(qik, vik, ik12, v_ref, det_o, base=2e-12, min_well=0.97) > 0.97
    # the canonical 8 through the pose kernels and the fused call on q
    idx8 = [list(ALL_12).index(b) for b in O.CANONICAL_8]
    det8 = {k: v[:, idx8] for k, v in det_o.items()}
    q8, v8 = robot.inverse_kinematics_batch(torch.tensor(Tg, device="cuda"), layout=layout)
    assert np.array_equal(v8.cpu().numpy(), v_ref[:, idx8])
    ikcond.assert_ik_close(q8.cpu().numpy(), v8.cpu().numpy(), ik12[:, idx8], v_ref[:, idx8], det8,
                           base=2e-12, min_well=0.97)
    # FP32: masks away from the thresholds, values to the conditioned 1e-5 gate
    _check_ik_f32(robot, c, Tg, layout)
