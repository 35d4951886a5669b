#!/usr/bin/env python
"""Golden of ModelCalibrator.optimize (utils/model_calibration.py:585-608) from the REAL reference:
the full fanuc_calibration.py:11-96 sequence (estimate_tool_base, estimate_parameters, optimize,
evaluate) on the real 240-row data set.  TEST INFRASTRUCTURE; takes ~20-40 minutes on one core
(every cost evaluation is 720 SymPy chain products; scipy's 3-point differences need 37 of them per
iteration).  Writes tests/golden/calibration_optimize.npz.

  python tests/golden/make_optimize_golden.py
"""
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
    mg.prepare_reference()
    import scipy
    import scipy.io
    import sympy as sp
    import utils.model_calibration as mcmod
    from utils.model_calibration import ModelCalibrator
    mcmod.tqdm = lambda it, **k: it
    mat = scipy.io.loadmat('data/calibration_dataset.mat')
    qs = np.array(mat['q'], dtype=float)
    pts = np.concatenate([np.array(mat[k], dtype=float)[:, :, np.newaxis]
                          for k in ('mA', 'mB', 'mC')], axis=1) * 1000
    mc = ModelCalibrator("TzRzTxRyTxRyTxTzRxRyRxTx", joint_indices=[0, 1, 0, 1, 0, 1, 0, 0, 1, 1, 1, 0],
                         variables=['d_1', 'q_1', 'd_2', 'q_2', 'd_3', 'q_3', 'd_4', 'd_5', 'q_4', 'q_5', 'q_6', 'd_6'],
                         link_lengths=[('d_1', 670), ('d_2', 312), ('d_3', 1075), ('d_5', 225), ('d_4', 1280), ("d_6", 215)],
                         offsets=np.array([0.0, -np.pi / 2, np.pi / 2, 0.0, 0.0, 0.0]),
                         directions=np.array([1, 1, -1, -1, -1, -1]))
    old = sys.stdout
    sys.stdout = io.StringIO()
    mc.get_reduced_model()
    T_base, T_tools = mc.estimate_tool_base(qs, pts)
    mc.estimate_parameters(qs, pts, T_base, T_tools)
    p_before = np.array(mc._parameters, dtype=float)
    mc.optimize(qs, pts, T_base, T_tools)
    p_after = np.array(mc._parameters, dtype=float)
    ev = mc.evaluate(qs, pts, T_base, T_tools)
    log = sys.stdout.getvalue()
    sys.stdout = old
    keys = sorted(ev)
    np.savez_compressed(os.path.join(HERE, "calibration_optimize.npz"), params_before=p_before, params_after=p_after,
                        eval_keys=np.array(keys), eval_after_optimize=np.array([float(ev[k]) for k in keys]),
                        T_base=np.array(T_base, dtype=float), T_tools=np.array([np.array(t, dtype=float) for t in T_tools]),
                        scipy_version=np.array(scipy.__version__), log=np.array(log[-4000:]))
    print("optimize golden done:", dict(zip(keys, [float(ev[k]) for k in keys])))


if __name__ == "__main__":
    main()
