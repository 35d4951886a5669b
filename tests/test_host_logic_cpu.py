"""CPU tests of the host-side logic: token algebra (the reference's tests/test_robo_math.py
cases), symbolic Jacobians (tests/test_jacobians.py), the model reducer against the real
reference's outputs, constants of the robots, trapezoid coefficients."""
import numpy as np
import pytest
import sympy as sp

from roboticmanipulators_b200.robots import Fanuc165F, RRRobot, RRRRobot
from roboticmanipulators_b200.utils.jacobians import JacobianCalculator
from roboticmanipulators_b200.utils.model_calibration import ModelCalibrator
from roboticmanipulators_b200.utils.polynomial_generator import PolynomialGenerator
from roboticmanipulators_b200.utils.robo_math import SymbolicTransformation as st
from roboticmanipulators_b200.utils.robo_math import TOKENS, Transformation as tf
from roboticmanipulators_b200.utils.trapezoidal_generator import TrapezoidalGenerator


def test_elementary_matrices_match_reference(golden):
    g = golden("generic_chain.npz")
    for tok in TOKENS:
        assert np.allclose(np.array(tf(tok, [0.7]).transformation, dtype=float),
                           g[f"elem_{tok}"], rtol=0, atol=1e-16), tok
    assert np.array_equal(np.array(tf.get_Rx(0), dtype=float), np.eye(4))
    assert np.allclose(np.array(tf.get_Rz_from_degrees(90), dtype=float),
                       [[0, -1, 0, 0], [1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], atol=1e-15)
    assert tf.get_Tx(3) == sp.Matrix([[1, 0, 0, 3], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    assert tf.get_Ty_inv(2)[1, 3] == -2
    T = tf("TxTyTz", [1.0, 1.2, 0.7]).transformation        # reference test_robo_math.py:198-208
    assert (T[0, 3], T[1, 3], T[2, 3]) == (1.0, 1.2, 0.7)
    with pytest.raises(ValueError):
        tf("TxTy", [1.0])
    with pytest.raises(ValueError):
        tf("Qx", [1.0])


def test_chain_against_reference_golden(golden):
    g = golden("generic_chain.npz")
    for c in range(int(g["n_cases"])):
        seq = str(g[f"c{c}_seq"])
        for v, T in zip(g[f"c{c}_vals"], g[f"c{c}_T"]):
            assert np.allclose(np.array(tf(seq, list(v)).transformation, dtype=float), T, atol=1e-13)


def test_symbolic_transformation_api():
    s = st("RzTx")
    assert s.variables == ["q_0", "d_0"]
    q, d = sp.symbols("q_0 d_0")
    assert sp.simplify(s.transformation[0, 3] - d * sp.cos(q)) == 0
    assert len(s.frames) == 2
    inv = s.inv()
    assert inv._seq == "TxiRzi" and inv.variables == ["d_0", "q_0"]
    assert s.inv().inv()._seq == "RzTx"
    assert np.allclose(np.array(sp.simplify(s.transformation * inv.transformation), dtype=float), np.eye(4))
    s2 = st("TzRx", ["l", "a"])
    s2.substitute([("l", 2.0)])
    assert s2.transformation[2, 3] == 2.0
    val = s2.evaluate_tuples([("a", 0.5)])
    assert abs(float(val[1, 1]) - np.cos(0.5)) < 1e-15
    assert (s * s2)._seq == "RzTxTzRx"
    assert st.get_Rx_inv("a") == st.get_Rx("a").T
    t = sp.Symbol("t")
    assert st("Rz", f_of_t=[True]).transformation[0, 0] == sp.cos(sp.Function("q_0")(t))
    with pytest.raises(ValueError):
        st("RzTx", ["only_one"])
    with pytest.raises(ValueError):
        st("Wz")
    assert st("RzTzTxRyTxRyRyiTx").inv()._seq == "TxiRyRyiTxiRyiTxiTziRzi"


def test_symbolic_jacobian_methods_agree():
    """tests/test_jacobians.py:10-57 of the reference."""
    for seq, ind in (("RzTx", [1, 0]), ("RzTxRzTx", [1, 0, 1, 0]), ("RzTxRzTxTyRz", [1, 1, 1, 1, 0, 1])):
        jc = JacobianCalculator(seq, ind, simplify=False)
        diff = sp.simplify(jc.calculate_scew() - jc.calculate_numeric())
        assert np.allclose(np.array(diff, dtype=float), 0)
    T_base = sp.Matrix([[1, 0, 0, 3], [0, 1, 0, 6], [0, 0, 1, -9], [0, 0, 0, 1]])
    T_tool = sp.Matrix([[0, 1, 0, 2], [1, 0, 0, -4], [0, 0, 1, 6], [0, 0, 0, 1]])
    jc = JacobianCalculator("RzTxRzTx", [1, 0, 1, 0], T_base=T_base, T_tool=T_tool)
    diff = sp.simplify(jc.calculate_scew() - jc.calculate_numeric())
    assert np.allclose(np.array(diff, dtype=float), 0)
    with pytest.raises(ValueError):
        JacobianCalculator("RzTx", [1])


def test_reduced_models_match_reference(golden):
    g = golden("reduced_models.npz")
    for nm, seq, ind in (("rr", "RzTxRzTx", [1, 0, 1, 0]),
                         ("rrr", "RzTzRyTxRyTx", [1, 0, 1, 0, 1, 0]),
                         ("fanuc13", "TzRzTzTxRyTxRyTzTxRxRyRxTx",
                          [0, 1, 0, 0, 1, 0, 1, 0, 0, 1, 1, 1, 0])):
        m = ModelCalibrator(seq, ind, None, None)
        m.verbose = False
        s, v = m.get_reduced_model()
        assert s == str(g[f"{nm}_seq"])
        assert v == list(g[f"{nm}_vars"])
        assert m._r_par == list(g[f"{nm}_par"])
        assert list(m._r_j_ind) == list(g[f"{nm}_j_ind"])
        assert [str(x) for x in m._r_nom_par] == list(g[f"{nm}_nom"])


def fanuc_calibrator():
    mc = ModelCalibrator(
        "TzRzTxRyTxRyTxTzRxRyRxTx", joint_indices=[0, 1, 0, 1, 0, 1, 0, 0, 1, 1, 1, 0],
        variables=['d_1', 'q_1', 'd_2', 'q_2', 'd_3', 'q_3', 'd_4', 'd_5', 'q_4', 'q_5', 'q_6', 'd_6'],
        link_lengths=[('d_1', 670), ('d_2', 312), ('d_3', 1075), ('d_5', 225), ('d_4', 1280), ("d_6", 215)],
        offsets=np.array([0.0, -np.pi / 2, np.pi / 2, 0.0, 0.0, 0.0]),
        directions=np.array([1, 1, -1, -1, -1, -1]))
    mc.verbose = False
    return mc


def test_fanuc_calibration_model_and_compile(golden):
    g = golden("calibration.npz")
    mc = fanuc_calibrator()
    assert mc.estimate_parameters(None, None) == ([], [])       # model not built yet
    s, v = mc.get_reduced_model()
    assert s == str(g["r_seq"]) == "RzTxTyRxRyRyTxRxRzRyRyTxTzRzRxRxTyTzRzRyRyTzRzRx"
    assert v == list(g["r_vars"]) and mc._r_par == list(g["r_par"])
    assert mc._to_base == ["Tz"] and mc._to_tool == ["Tx"]
    assert mc._n_params == 18
    m = mc._compile(np.zeros(18), np.eye(4), [np.eye(4)] * 3)
    assert (m.njoints, m.nparams, m.ntools) == (6, 18, 3)
    assert list(m.param_joint)[:18] == [0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4]
    assert list(m.joint_axis)[:6] == [2, 1, 1, 0, 1, 0]
    assert m.seg[0][3] == 312.0 and m.seg[2][3] == 1280.0 and m.seg[2][11] == 225.0
    # nominal token values feed _prepare_evaluation exactly like the reference
    vals = mc._prepare_evaluation(g["qs"][0])
    assert len(vals) == 24 and vals[1] == 312.0


def test_robot_constants_match_reference(golden):
    g = golden("fanuc_fk_ik.npz")
    r = Fanuc165F()
    assert float(r.d) == float(g["d"]) and float(r.dq) == float(g["dq"])
    assert np.array_equal(np.asarray(r.qs_lim_rad), g["qs_lim_rad"])
    assert r.epsilon == 1e-5
    assert r.T_base == sp.eye(4)
    st_ = r._struct()
    assert st_.has_base == 0 and st_.l[5] == 1280.0 and st_.directions[2] == -1.0
    r2 = Fanuc165F(T_base=g["rotated_T_base"], T_tool=g["rotated_T_tool"])
    s2 = r2._struct()
    assert s2.has_base == 1 and s2.has_tool == 1
    Tb = np.vstack([np.array(list(s2.T_base)).reshape(3, 4), [0, 0, 0, 1]])
    Tbi = np.vstack([np.array(list(s2.T_base_inv)).reshape(3, 4), [0, 0, 0, 1]])
    assert np.allclose(Tb @ Tbi, np.eye(4), atol=1e-12)
    assert RRRRobot.ls == (1.0, 1.0, 1.0) and RRRobot()._ls == (0.8, 0.8)
    assert RRRobot(lengths=(0.5, 1.25))._struct().l[1] == 1.25


def test_trapezoid_coefficients_host(golden):
    g = golden("trajectories.npz")
    for (v, a, f, q0, qf), ref in zip(g["coef_cases"], g["coef_out"]):
        tg = TrapezoidalGenerator(v, a, control_freq=int(f))
        assert tuple(float(x) for x in tg.generate_coefficients(q0, qf)) == tuple(ref)
    tg = TrapezoidalGenerator(dq_max=1, ddq_max=10)
    tg.set_limits(2, 2)
    assert np.allclose(tg.generate_coefficients(0, 5), (1.0, 2.5))     # reference test_generator
    tg.set_frequency(5)
    assert np.allclose(tg.generate_coefficients(0, 5), (1.0, 2.6))     # test_generator_frequency


def test_polynomial_generator():
    """tests/test_polynomial_generator.py of the reference."""
    pg = PolynomialGenerator()
    coefs = pg.generate_coefficients([0, 0, 0], [1, 0, 0], 1.0)
    assert np.allclose(coefs, [0, 0, 0, 10, -15, 6])
    assert np.allclose(pg.dpolynomial_from_coefs(0, coefs, 0, 1, 11), pg.polynomial_from_coefs(coefs, 0, 1, 11))


def test_calibration_mat_roundtrip(golden, tmp_path):
    from roboticmanipulators_b200.utils.calibration_data import load_calibration_mat, save_calibration_mat
    g = golden("calibration.npz")
    path = str(tmp_path / "calibration_dataset.mat")
    save_calibration_mat(path, g["qs"], g["pts"])
    qs, pts = load_calibration_mat(path, device=None)
    assert qs.shape == (240, 6) and pts.shape == (240, 9)
    assert np.array_equal(qs, g["qs"])
    assert np.max(np.abs(pts - g["pts"])) < 1e-9          # metres <-> millimetres round trip
    qt, pt = load_calibration_mat(path, device="cpu")
    assert qt.dtype == __import__("torch").float64 and tuple(pt.shape) == (240, 9)


def test_mat_loader_reproduces_the_reference_dataset(golden):
    """fanuc_calibration.py:13-20: `load_calibration_mat` on a MAT-v5 file with the reference's layout
    (q [240,6] rad, mA / mB / mC [240,3] metres) returns the arrays the golden run of the real
    reference used (calibration.npz qs / pts, captured from scipy.io.loadmat of the reference's own
    data/calibration_dataset.mat).  tests/golden/calibration_dataset_240.mat was written by
    save_calibration_mat from those arrays (metres = mm / 1000, hence 1 ulp); where the reference's
    own file is present (the build container) it must load bit for bit."""
    import os
    from roboticmanipulators_b200.utils.calibration_data import load_calibration_mat
    g = golden("calibration.npz")
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "calibration_dataset_240.mat")
    qs, pts = load_calibration_mat(here, device=None)
    assert qs.shape == (240, 6) and pts.shape == (240, 9)
    assert np.array_equal(qs, g["qs"])
    assert np.max(np.abs(pts - g["pts"]) / np.abs(g["pts"])) < 4e-16
    ref = "/root/reference/data/calibration_dataset.mat"
    if os.path.exists(ref):
        q2, p2 = load_calibration_mat(ref, device=None)
        assert np.array_equal(q2, g["qs"]) and np.array_equal(p2, g["pts"])


def test_chain_specialiser_source_compiles_for_sm100a(tmp_path):
    """The run-time specialiser (rk_chain_jit.cu) is exercised on the GPU box through NVRTC; here its
    generated source -- scalar and two-samples-per-thread, FP64 and FP32, derivative tokens and
    constant segments included -- is compiled with nvcc for sm_100a, so a generator bug shows up
    without a GPU."""
    import ctypes as C
    import shutil
    import subprocess
    from roboticmanipulators_b200 import _lib
    from roboticmanipulators_b200.utils.robo_math import make_chain_struct
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    lib = _lib.load()
    Tb = np.eye(4)
    Tb[:3, 3] = [0.1, 0.2, 0.3]
    cases = [
        (["Tz", "Rz", "Tx", "Ry", "Tx", "Ry", "Tx", "Rx", "Ry", "Rx", "Tx"],
         [0, 1, 0, 1, 0, 1, 0, 1, 1, 1, 0], None, None),
        (["Rz", "Tx", "Ryd", "Tzi", "Rxi", "Tyd", "Ty"], [1, 0, 0, 0, 0, 0, 1], Tb, Tb),
    ]
    for k, (tokens, active, Tb_, Tt_) in enumerate(cases):
        ch = make_chain_struct(tokens, active, T_base=Tb_, T_tool=Tt_)
        for f32 in (0, 1):
            for want in (1, 3, 1 | 4, 3 | 4) + ((8,) if not f32 else ()):
                need = lib.rk_chain_jit_source(C.byref(ch), len(tokens), _lib.RK_SOA, want, f32, None, 0)
                assert need > 0
                buf = C.create_string_buffer(need)
                assert lib.rk_chain_jit_source(C.byref(ch), len(tokens), _lib.RK_SOA, want, f32, buf, need) == need
                src = tmp_path / f"jit_{k}_{f32}_{want}.cu"
                src.write_bytes(buf.value)
                r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-cubin",
                                    "-o", str(src.with_suffix(".cubin")), str(src)], capture_output=True, text=True)
                assert r.returncode == 0, r.stderr[-2000:]
                if want & 4:
                    assert b"LD2(" in buf.value and b"ST2(" in buf.value


def test_get_angles_matches_reference_values():
    """Transformation.get_angles (robo_math.py:448-464): expected values produced by the reference
    itself for generic rotations, its gimbal branch (R[2,0] == 1), the R[2,2] == 0 fall-backs, and
    its unbound-y corner, which raises there and here."""
    cases = [
        ([[0.705957417269, -0.702419831123, -0.090722135383], [-0.192210119246, -0.313293905934, 0.930001182023], [-0.681673965369, -0.639103520058, -0.35618435618]],
         [-0.2658251885952269, -0.7500481162640394, -2.0792411469299323]),
        ([[-0.229094160879, 0.949412986178, 0.21478092822], [-0.936789830762, -0.274990703048, 0.216344461953], [0.264463000118, -0.151641336426, 0.95239919501]],
         [-1.8106412608984057, 0.2676470545929161, -0.15789498403708382]),
        ([[0.941832710923, 0.27957662208, -0.186515567772], [0.174901709248, 0.066162037069, 0.982360410925], [0.286985255261, -0.95784106053, 0.013415141665]],
         [0.18361198051204158, 0.29107821881298124, -1.556791639416813]),
        ([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0]],
         [0.0, 1.5707963267948966, 1.5707963267948966]),
        ([[0.0, -1.0, 0.0], [0.6, 0.0, -0.8], [0.8, 0.0, 0.6]],
         [1.5707963267948966, 0.9272952180016123, 0.0]),
        ([[0.6, 0.0, -0.8], [0.0, 1.0, 0.0], [0.8, 0.0, 0.6]],
         [0.0, 0.9272952180016123, 0.0]),
        ([[0.8, 0.0, 0.6], [0.6, 0.0, -0.8], [0.0, 1.0, 0.0]],
         [0.6435011087932844, -0.0, 1.5707963267948966]),
    ]
    for R, want in cases:
        got = tf.get_angles(np.array(R))
        assert np.allclose([float(v) for v in got], want, rtol=0, atol=1e-15), (R, got, want)
    with pytest.raises(UnboundLocalError):
        tf.get_angles(np.array([[0.6, -0.8, 0.0], [0.0, 0.0, -1.0], [0.8, 0.6, 0.0]]))
