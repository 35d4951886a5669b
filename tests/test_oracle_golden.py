"""Pin the NumPy oracle against golden vectors produced by the REAL reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import np_oracle as O


def scaled(a, b, scale=None):
    a, b = np.asarray(a, float), np.asarray(b, float)
    s = max(1.0, float(np.max(np.abs(b)))) if scale is None else scale
    return float(np.max(np.abs(a - b))) / s


def test_elementary_tokens(golden):
    g = golden("generic_chain.npz")
    for tok in O.VALID_TOKENS:
        assert np.allclose(O.elem(tok, 0.7), g[f"elem_{tok}"], rtol=0, atol=1e-16), tok
    with pytest.raises(ValueError):
        O.elem("Qx", 0.1)
    with pytest.raises(ValueError):
        O.chain("TxTy", [1.0])


def test_generic_chain_and_jacobian(golden):
    g = golden("generic_chain.npz")
    for c in range(int(g["n_cases"])):
        seq, ind = str(g[f"c{c}_seq"]), list(g[f"c{c}_ind"])
        kw = dict(T_base=g["T_base"], T_tool=g["T_tool"]) if bool(g[f"c{c}_bt"]) else {}
        vals = g[f"c{c}_vals"]
        assert scaled(O.chain(seq, vals), g[f"c{c}_T"]) < 1e-14
        J = O.jacobian_numerically(seq, ind, vals, **kw)
        assert scaled(J, g[f"c{c}_J"]) < 1e-13
        # screw form == matrix-derivative form (tests/test_jacobians.py:10-57)
        assert scaled(O.jacobian_screw(seq, ind, vals, **kw), g[f"c{c}_J"]) < 1e-13
    J = O.jacobian_numerically("RzTxRzTx", [1, 0, 1, 0], [0.0, 1.0, 0.0, 2.0])
    assert np.allclose(J, [[0, 0], [3, 2], [0, 0], [0, 0], [0, 0], [1, 1]])  # test_jacobians.py:59-75
    assert np.allclose(J, g["test_numeric_values_J"])


def test_inverse_sequence():
    assert O.inverse_sequence("RzTzTxRyTxRyRyiTx") == "TxiRyRyiTxiRyiTxiTziRzi"
    v = np.array([0.3, 1.0, -2.0, 0.5])
    T = O.chain("RzTxRyiTz", v)
    Ti = O.chain(O.inverse_sequence("RzTxRyiTz"), v[::-1])
    assert np.allclose(T @ Ti, np.eye(4), atol=1e-15)


def test_fanuc_constants_and_fk(golden):
    g = golden("fanuc_fk_ik.npz")
    c = O.FanucConst()
    assert c.d == float(g["d"]) and c.dq == float(g["dq"])
    assert np.array_equal(c.lim, g["qs_lim_rad"])
    T = O.fanuc_fk(g["default_q"])
    # G0 (tests/test_robots.py:13-31)
    assert np.allclose(T[0], [[1, 0, 0, 1807], [0, 1, 0, 0], [0, 0, 1, 1970], [0, 0, 0, 1]], atol=1e-12)
    assert np.allclose(T[1], [[0, -1, 0, 0], [1, 0, 0, 1807], [0, 0, 1, 1970], [0, 0, 0, 1]], atol=1e-12)
    Tg = g["default_T"]
    assert scaled(T[:, :3, :3], Tg[:, :3, :3], 1.0) < 1e-14
    assert scaled(T[:, :3, 3], Tg[:, :3, 3]) < 1e-14
    for tag in ("translated", "rotated"):
        c = O.FanucConst(T_base=g[f"{tag}_T_base"], T_tool=g[f"{tag}_T_tool"])
        T = O.fanuc_fk(g[f"{tag}_q"], c)
        assert scaled(T[:, :3, :3], g[f"{tag}_T"][:, :3, :3], 1.0) < 1e-14
        assert scaled(T[:, :3, 3], g[f"{tag}_T"][:, :3, 3]) < 1e-14


def _check_ik(T, ik12, branches, c):
    for b, (m, k, w) in enumerate(branches):
        q, valid = O.fanuc_ik(T, int(m), int(k), int(w), c)
        ref = ik12[:, b]
        ref_valid = np.any(ref != 0, axis=-1) | np.any(np.isnan(ref), axis=-1)
        # the reference's failure sentinel is the zero vector; a valid all-zero
        # solution is indistinguishable, so compare values (zeros included)
        assert np.array_equal(valid | ~np.any(q != 0, -1), ref_valid | ~np.any(ref != 0, -1))
        assert np.array_equal(np.isnan(q), np.isnan(ref))
        d = q - ref
        # atan2 branch cut: when the reference's r12 is rounding noise around 0
        # with r13 < 0, q3/q5 land on +pi or -pi (same rotation); compare those
        # two joints modulo 2*pi
        d[:, [3, 5]] = (d[:, [3, 5]] + np.pi) % (2 * np.pi) - np.pi
        assert np.nanmax(np.abs(d)) < 2e-10, (m, k, w, np.nanmax(np.abs(d)))


def test_fanuc_ik_all_branches(golden):
    g = golden("fanuc_fk_ik.npz")
    br = g["branches12"]
    _check_ik(g["default_T"], g["default_ik12"], br, O.FanucConst())
    _check_ik(g["extra_T"], g["extra_ik12"], br, O.FanucConst())
    for tag in ("translated", "rotated"):
        c = O.FanucConst(T_base=g[f"{tag}_T_base"], T_tool=g[f"{tag}_T_tool"])
        _check_ik(g[f"{tag}_T"], g[f"{tag}_ik12"], br, c)
    # reference tests: unreachable and out-of-limit targets return zeros
    q, v = O.fanuc_ik(g["extra_T"][:2], 1, 1, 0)
    assert not v.any() and not q.any()


def test_fanuc_custom_constants(golden):
    """Non-stock lengths / offsets / directions (fanuc.py:72-96), run through the real reference."""
    g = golden("fanuc_custom.npz")
    c = O.FanucConst(lengths=g["lengths"], offsets=g["offsets"], directions=g["directions"])
    assert c.d == float(g["d"]) and c.dq == float(g["dq"])
    T = O.fanuc_fk(g["q"], c)
    assert scaled(T[:, :3, :3], g["T"][:, :3, :3], 1.0) < 1e-14
    assert scaled(T[:, :3, 3], g["T"][:, :3, 3]) < 1e-14
    _check_ik(g["T"], g["ik12"], g["branches12"], c)
    assert np.any(g["ik12"] != 0, axis=-1).mean() > 0.2      # the set does exercise valid branches


def test_fanuc_ik_roundtrip_k1():
    rng = np.random.default_rng(7)
    c = O.FanucConst()
    q = rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(500, 6))
    T = O.fanuc_fk(q)
    qb, vb = O.fanuc_ik_branches(T)
    assert vb.any(axis=1).mean() > 0.8   # measured 0.86: reference quirk, SURVEY A2
    for b, (m, k, w) in enumerate(O.CANONICAL_8):
        if k != 1:
            continue
        ok = vb[:, b] & ~np.isnan(qb[:, b]).any(-1)
        Tb = O.fanuc_fk(qb[ok, b])
        assert np.max(np.abs(Tb - T[ok])) < 1e-6


def test_fanuc_jacobian(golden):
    g = golden("fanuc_jacobian.npz")
    qm = g["qm"]
    J = O.fanuc_jacobian(qm, model_angles=True)
    Jg = g["default_J"]
    assert scaled(J[:, :3], Jg[:, :3]) < 1e-14
    assert scaled(J[:, 3:], Jg[:, 3:], 1.0) < 1e-14
    # G3 first row and determinant
    assert abs(J[0, 0, 0] - (-826.34434369832195)) < 1e-9
    det = O.jac_det(J)
    assert abs(det[0] - 270295653.8623566) / 270295653.8623566 < 1e-12
    # the reference's symbolic evalf().det(): scale-relative gate
    scale = np.prod(np.linalg.norm(J[:6], axis=1), axis=1)
    assert np.all(np.abs(det[:6] - g["det_symbolic_first6"]) <= 1e-12 * scale)
    n = int(g["rotated_n"])
    fk = golden("fanuc_fk_ik.npz")
    c = O.FanucConst(T_base=fk["rotated_T_base"], T_tool=fk["rotated_T_tool"])
    Jr = O.fanuc_jacobian(qm[:n], c, model_angles=True)
    assert scaled(Jr[:, :3], g["rotated_J"][:, :3]) < 1e-14
    assert scaled(Jr[:, 3:], g["rotated_J"][:, 3:], 1.0) < 1e-14


def test_rrr_rr(golden):
    g = golden("rrr_rr.npz")
    assert scaled(O.rrr_fk(g["rrr_q"]), g["rrr_T"]) < 1e-15
    for b, (m, k) in enumerate(g["rrr_branches"]):
        q, _ = O.rrr_ik(g["rrr_T"], int(m), int(k))
        assert np.max(np.abs(q - g["rrr_ik4"][:, b])) < 1e-9
    # SURVEY G5
    T = O.rrr_fk([np.pi / 2, 0, 0.2])
    assert np.allclose(T[:3, 3], [0, 1.980066577841, 0.801330669205], atol=1e-11)
    for nm, ls in (("rr", (0.8, 0.8)), ("rr_l", (0.5, 1.25))):
        assert scaled(O.rr_fk(g[f"{nm}_q"], ls=ls), g[f"{nm}_T"]) < 1e-15
        for b, m in enumerate((1, -1)):
            q, _ = O.rr_ik(g[f"{nm}_T"], m, ls=ls)
            assert np.max(np.abs(q - g[f"{nm}_ik2"][:, b])) < 1e-9


@pytest.mark.parametrize("name", ["arms_base_tool.npz", "arms_base_tool_large.npz"])
def test_arms_with_base_and_tool(golden, name):
    """RRR / RR behind rotating T_base and T_tool (robot.py:39-57), from the real reference (57
    joint vectors per arm, and the same recipe at 3001: make_golden.py --only armslarge)."""
    g = golden(name)
    Tb, Tt = g["T_base"], g["T_tool"]
    assert scaled(O.rrr_fk(g["rrr_q"], Tb, Tt), g["rrr_T"]) < 1e-14
    for b, (m, k) in enumerate(g["rrr_branches"]):
        q, _ = O.rrr_ik(g["rrr_T"], int(m), int(k), Tb, Tt)
        assert np.max(np.abs(q - g["rrr_ik4"][:, b])) < 1e-9
    ls = tuple(g["rr_lengths"])
    assert scaled(O.rr_fk(g["rr_q"], Tb, Tt, ls=ls), g["rr_T"]) < 1e-14
    for b, m in enumerate((1, -1)):
        q, _ = O.rr_ik(g["rr_T"], m, Tb, Tt, ls=ls)
        assert np.max(np.abs(q - g["rr_ik2"][:, b])) < 1e-9


def test_trapezoid_coefficients(golden):
    g = golden("trajectories.npz")
    for (v, a, f, q0, qf), ref in zip(g["coef_cases"], g["coef_out"]):
        assert O.trap_coefficients(q0, qf, v, a, int(f)) == tuple(ref)
    assert O.trap_coefficients(0, 5, 2, 2) == (1.0, 2.5)        # tests/test_trapezoidal_generator.py
    assert O.trap_coefficients(0, 5, 2, 2, 5) == (1.0, 2.6)


def test_p2p_trajectories_bit_exact(golden):
    g = golden("trajectories.npz")
    lengths = {}
    for i in range(int(g["p2p_count"])):
        spec = g[f"p2p{i}_spec"]
        q0, qf, n, cf = spec[:3], spec[3:6], int(spec[6]), int(spec[7])
        q, dq, ddq, t, L = O.p2p_batch(q0[None], qf[None], n, 1, 20, cf)
        if int(g[f"p2p{i}_error"]):
            assert L[0] == 0
            continue
        ref = g[f"p2p{i}_q"]
        assert L[0] == len(ref)
        assert np.array_equal(q[0, :L[0]], ref)
        assert np.array_equal(dq[0, :L[0]], g[f"p2p{i}_dq"])
        assert np.array_equal(ddq[0, :L[0]], g[f"p2p{i}_ddq"])
        assert np.array_equal(t[0, :L[0]], g[f"p2p{i}_t"])
        if i < 10:
            lengths[(n, cf)] = int(L[0])
    # variable-length outputs
    assert [lengths[(n, 10)] for n in (37, 50, 100, 1000, 1001)] == [35, 49, 99, 998, 999]
    assert lengths[(50, 0)] == 49 and lengths[(1000, 0)] == 999


def test_lin_trajectories(golden):
    g = golden("trajectories.npz")
    for i in range(int(g["lin_count"])):
        spec = g[f"lin{i}_spec"]
        p0, pf, n, cf = spec[:3], spec[3:6], int(spec[6]), int(spec[7])
        x, dx, t, q, dq, valid, L = O.lin_rrr_batch(p0[None], pf[None], n, 1, 10, cf)
        L = int(L[0])
        assert L == len(g[f"lin{i}_x"])
        assert np.array_equal(x[0, :L], g[f"lin{i}_x"])
        assert np.array_equal(dx[0, :L], g[f"lin{i}_dx"])
        assert np.array_equal(t[0, :L], g[f"lin{i}_t"])
        assert np.max(np.abs(q[0, :L] - g[f"lin{i}_q"])) < 1e-12
        ref_dq = g[f"lin{i}_dq"]
        ok = ~np.isnan(ref_dq).any(-1)
        assert np.max(np.abs(dq[0, :L][ok] - ref_dq[ok])) < 1e-9 * max(1, np.abs(ref_dq[ok]).max())
    assert len(golden("trajectories.npz")["lin0_x"]) == 48


def _calib_model(g):
    link = [('d_1', 670), ('d_2', 312), ('d_3', 1075), ('d_5', 225), ('d_4', 1280), ("d_6", 215)]
    return O.CalibModel(g["r_seq"], g["r_vars"], g["r_j_ind"], g["r_nom_par"], link,
                        offsets=[0.0, -np.pi / 2, np.pi / 2, 0.0, 0.0, 0.0],
                        directions=[1, 1, -1, -1, -1, -1])


T_BASE_0 = np.array([[1, 0, 0, -563.145620015451], [0, 1, 0, -559.28130331986],
                     [0, 0, 1, 554.98685687043], [0, 0, 0, 1]])
T_TOOLS_0 = [np.array([[1, 0, 0, 416.687065414667], [0, 1, 0, -17.0660344931335],
                       [0, 0, 1, 59.6890106614009], [0, 0, 0, 1]]),
             np.array([[1, 0, 0, 420.667529952858], [0, 1, 0, 57.1117905464326],
                       [0, 0, 1, -15.7733691549466], [0, 0, 0, 1]]),
             np.array([[1, 0, 0, 420.483792669375], [0, 1, 0, -43.7121193657286],
                       [0, 0, 1, -42.4708914215651], [0, 0, 0, 1]])]


def test_calibration_real_dataset(golden):
    g = golden("calibration.npz")
    model = _calib_model(g)
    assert model.r_seq == "RzTxTyRxRyRyTxRxRzRyRyTxTzRzRxRxTyTzRzRyRyTzRzRx"
    assert model.n_params == 18
    qs, pts = g["qs"], g["pts"]
    zero = np.zeros(18)
    keys = [str(k) for k in g["eval_keys"]]
    ev0 = O.calib_evaluate(model, qs, pts, zero, T_BASE_0, T_TOOLS_0)
    assert np.allclose([ev0[k] for k in keys], g["eval_initial"], rtol=1e-12)
    assert abs(ev0["distance_rms"] - 13.1540215624739) < 1e-10      # docs/GeometricCalibration.md:36
    assert abs(ev0["max_dist"] - 18.4905680970990) < 1e-10
    assert scaled(O.calib_pose(model, qs[:1], zero, T_BASE_0, T_TOOLS_0[0])[0], g["pose_row0"]) < 1e-14
    assert scaled(O.calib_jacobian(model, qs[:1], zero, T_BASE_0, T_TOOLS_0[0])[0], g["jac_row0"]) < 1e-14
    est, T_base, T_tools, A_sum, _ = O.calib_estimate_tool_base(model, qs, pts, zero)
    assert scaled(A_sum, g["A_sum"]) < 1e-13
    assert np.max(np.abs(est - g["base_est"]) / np.maximum(1, np.abs(g["base_est"]))) < 1e-9
    assert scaled(T_base, g["T_base"]) < 1e-10
    ev1 = O.calib_evaluate(model, qs, pts, zero, g["T_base"], list(g["T_tools"]))
    assert np.allclose([ev1[k] for k in keys], g["eval_after_base"], rtol=1e-11)
    params, H, gg = O.calib_estimate_parameters(model, qs, pts, zero, g["T_base"], list(g["T_tools"]))
    assert scaled(H, g["J_sum"]) < 1e-13
    assert np.max(np.abs(params - g["params"])) < 1e-9               # north_star: within 1e-9
    ev2 = O.calib_evaluate(model, qs, pts, g["params"], g["T_base"], list(g["T_tools"]))
    assert np.allclose([ev2[k] for k in keys], g["eval_after_params"], rtol=1e-11)
    assert abs(ev2["distance_rms"] - 0.531528594785394) < 1e-9      # docs/GeometricCalibration.md:51
    assert abs(ev2["max_dist"] - 1.32318162735269) < 1e-9


def test_polynomial_profiles(golden):
    g = golden("polynomial.npz")
    assert np.allclose(O.poly_coefficients([0, 0, 0], [1, 0, 0], 1.0), [0, 0, 0, 10, -15, 6])
    for i in range(len(g["tf"])):
        c = O.poly_coefficients(g["q0"][i], g["qf"][i], g["tf"][i], g["t0"][i])
        assert np.array_equal(c, g["coefs"][i])
        assert np.array_equal(O.poly_sample(c, 0, g["t0"][i], g["tf"][i], 64), g["q"][i])
        assert np.array_equal(O.poly_sample(c, 1, g["t0"][i], g["tf"][i], 64), g["dq"][i])
        assert np.array_equal(O.poly_sample(c, 2, g["t0"][i], g["tf"][i], 64), g["ddq"][i])


def test_config1_oracle_against_the_real_reference(golden):
    """BASELINE config 1 at full size (1e4 uniform-in-limits poses, seed 0) from the REAL reference:
    the oracle reproduces its FK and Jacobian to 1e-12 scaled, all 1.2e5 IK validity decisions of
    the 12 (m, k, w) combinations exactly, and every IK value within the conditioned 1e-12 bound
    of tests/ikcond.py (the gate the CUDA kernels are then held to against this same file)."""
    import ikcond
    g = golden("fanuc_config1.npz")
    q, T34, ik12, Jg = g["q"], g["T34"], g["ik12"], g["J"]
    assert q.shape == (10000, 6)
    c = O.FanucConst()
    rng = np.random.default_rng(0)                      # the recipe of make_golden.gen_config1
    assert np.array_equal(q, rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(10000, 6)))
    T = O.fanuc_fk(q, c)
    assert np.max(np.abs(T[:, :3, :3] - T34[:, :, :3])) < 1e-12
    assert np.max(np.abs(T[:, :3, 3] - T34[:, :, 3]) / np.maximum(1, np.abs(T34[:, :, 3]).max(1, keepdims=True))) < 1e-12
    J = O.fanuc_jacobian(q, c)
    assert np.max(np.abs(J[:, :3] - Jg[:, :3])) / np.abs(Jg[:, :3]).max() < 1e-12
    assert np.max(np.abs(J[:, 3:] - Jg[:, 3:])) < 1e-12
    Tg = np.zeros((10000, 4, 4))
    Tg[:, :3], Tg[:, 3, 3] = T34, 1.0
    br = [tuple(int(v) for v in b) for b in g["branches12"]]
    q_o, v_o, det = ikcond.ik_details(Tg, br, c)
    v_ref = np.any(ik12 != 0.0, axis=2)
    assert 0.3 < v_ref.mean() < 0.5
    frac = ikcond.assert_ik_close(ik12, v_ref, q_o, v_o, det)
    assert frac > 0.98


def test_general_robot_oracle_against_the_real_reference(golden):
    """A robot with NOTHING at its default (custom link lengths, joint offsets and directions,
    rotating base and tool transforms), 4000 uniform-in-limits joint vectors through the REAL
    reference (make_golden.py --only general): FK to 1e-12 scaled, all 4.8e4 validity decisions of
    the 12 (m, k, w) combinations exact, every IK value within the conditioned bound."""
    import ikcond
    g = golden("fanuc_general.npz")
    q, T34, ik12 = g["q"], g["T34"], g["ik12"]
    n = len(q)
    assert n == 4000
    c = O.FanucConst(T_base=g["T_base"], T_tool=g["T_tool"], lengths=g["lengths"], offsets=g["offsets"],
                     directions=g["directions"])
    T = O.fanuc_fk(q, c)
    assert np.max(np.abs(T[:, :3, :3] - T34[:, :, :3])) < 1e-12
    assert np.max(np.abs(T[:, :3, 3] - T34[:, :, 3]) / np.maximum(1, np.abs(T34[:, :, 3]).max(1, keepdims=True))) < 1e-12
    Tg = np.zeros((n, 4, 4))
    Tg[:, :3], Tg[:, 3, 3] = T34, 1.0
    br = [tuple(int(v) for v in b) for b in g["branches12"]]
    q_o, v_o, det = ikcond.ik_details(Tg, br, c)
    v_ref = np.any(ik12 != 0.0, axis=2)
    assert 0.05 < v_ref.mean() < 0.6
    frac = ikcond.assert_ik_close(ik12, v_ref, q_o, v_o, det)
    assert frac > 0.97


def test_p2p_trajectories_large_bit_exact(golden):
    """480 random point-to-point moves through the REAL reference (make_golden.py --only trajlarge),
    six (n, control_freq) groups, a third of the short-n ones raising there: the oracle reproduces
    every sample count, time and q / dq / ddq bit for bit."""
    g = golden("trajectories_large.npz")
    for gi in range(int(g["groups"])):
        n, cf = (int(v) for v in g[f"g{gi}_spec"])
        q, dq, ddq, t, L = O.p2p_batch(g[f"g{gi}_q0"], g[f"g{gi}_qf"], n, 1, 20, cf)
        assert np.array_equal(L, g[f"g{gi}_len"]), (n, cf)
        assert np.array_equal(q, g[f"g{gi}_q"]) and np.array_equal(dq, g[f"g{gi}_dq"])
        assert np.array_equal(ddq, g[f"g{gi}_ddq"]) and np.array_equal(t, g[f"g{gi}_t"])


def test_random_chains_against_the_real_reference(golden):
    """28 random token sequences (3 - 12 tokens, inverse tokens as passive elements, random joint
    masks, half of them behind base / tool) through the REAL reference (make_golden.py --only
    chainrandom): product and Jacobian (matrix-derivative and screw form) of the oracle."""
    g = golden("generic_chain_random.npz")
    for c in range(int(g["n_cases"])):
        seq, ind = str(g[f"c{c}_seq"]), [int(i) for i in g[f"c{c}_ind"]]
        kw = dict(T_base=g["T_base"], T_tool=g["T_tool"]) if bool(g[f"c{c}_bt"]) else {}
        vals = g[f"c{c}_vals"]
        assert scaled(O.chain(seq, vals), g[f"c{c}_T"]) < 1e-14, seq
        assert scaled(O.jacobian_numerically(seq, ind, vals, **kw), g[f"c{c}_J"]) < 1e-13, seq
        assert scaled(O.jacobian_screw(seq, ind, vals, **kw), g[f"c{c}_J"]) < 1e-13, seq
