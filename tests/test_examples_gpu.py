"""The example scripts (batched counterparts of the reference's fanuc_kinematics.py,
fanuc_calibration.py and rrr_robot_planning.py) run end to end and report sane numbers."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def test_fanuc_kinematics_example():
    import fanuc_kinematics_batch as ex
    out = ex.main(20000, verbose=False)
    assert out["max_roundtrip_error_mm"] < 1e-6
    assert 0.60 < out["original_q_recovered"] < 0.66   # as the reference's own IK (see the example's header)
    assert 2.0 < out["valid_solutions_per_pose"] <= 8.0


def test_fanuc_calibration_example():
    import fanuc_calibration_gpu as ex
    hist = ex.main(None, 1, refine=False, verbose=False)
    assert abs(hist[0]["distance_rms"] - 13.1540215624739) < 1e-9      # README.md:103 of the reference
    assert abs(hist[1]["distance_rms"] - 0.531528594785394) < 1e-8     # docs/GeometricCalibration.md:51


def test_rrr_planning_example():
    import rrr_planning_batch as ex
    out = ex.main(512, 50, verbose=False)
    assert out["p2p_goal_error"] < 0.2                 # the last sample is one control period short of q_f
    assert out["lin_tracking_error"] < 1e-9
    assert out["lin_ik_success"] > 0.9


def test_c_abi_demo_runs(tmp_path):
    """The plain-C program of examples/c_abi_demo.c (rk_fanuc_init + the fused SOA call + the
    host-array streaming call, no Python, no torch) builds with gcc and passes its own checks."""
    import shutil
    import subprocess
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("gcc / CUDA headers not available")
    exe = tmp_path / "c_abi_demo"
    libdir = os.path.join(ROOT, "roboticmanipulators_b200")
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                        os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L", libdir, "-lrk_b200",
                        "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    env = dict(os.environ, LD_LIBRARY_PATH=libdir + ":" + os.path.join(cuda, "lib64") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([str(exe), "50000"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "0 differing validity flags" in r.stdout


def test_rr_synthetic_calibration_example():
    """A model the structure-specialised kernels do not know (RR arm, 3 parameters) through the
    generic calibration kernels: one iteration recovers the 0.1 link error to the noise floor; a
    second one follows the reference's blend rule (model_calibration.py:442-446) exactly."""
    import numpy as np
    import rr_synthetic_calibration_gpu as ex
    sigma = 0.001
    hist, params = ex.main(50000, 1, sigma=sigma, verbose=False)
    floor = sigma * np.sqrt(3)                            # the tool / base fit was made before the link fit
    assert 0.95 * floor < hist[0]["distance_rms"] < 1.15 * floor
    assert abs(params[0] - 0.1) < 5e-3 and np.all(np.abs(params[1:]) < 1e-3)
    hist2, params2 = ex.main(50000, 2, sigma=sigma, verbose=False)
    assert abs(params2[0] - 0.95 * params[0]) < 1e-3      # (1 - s) pi + s dpi with s = 0.05, |dpi| << pi
    assert hist2[1]["distance_rms"] > hist2[0]["distance_rms"]
