#!/usr/bin/env python
"""Counterpart of the reference's `fanuc_calibration.py`: the geometric calibration loop on a
calibration_dataset.mat file (q [m,6] rad; mA, mB, mC [m,3] metres), every pass over the
measurements on the GPU -- evaluate, estimate_tool_base, estimate_parameters (the J^T J / J^T r
reduction), evaluate -- and optionally the least-squares refinement (`optimize`).

  python examples/fanuc_calibration_gpu.py [path/to/calibration_dataset.mat] [iterations]

Without a path the 240-measurement file of the tests is used.
"""
import os
import sys

import numpy as np
import sympy as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from roboticmanipulators_b200.utils.calibration_data import load_calibration_mat
from roboticmanipulators_b200.utils.model_calibration import ModelCalibrator


def translation(x, y, z):
    return sp.Matrix([[1, 0, 0, x], [0, 1, 0, y], [0, 0, 1, z], [0, 0, 0, 1]])


def main(path=None, iterations=1, refine=False, verbose=True):
    path = path or os.path.join(ROOT, "tests", "golden", "calibration_dataset_240.mat")
    qs, pts = load_calibration_mat(path)                      # CUDA tensors, points in millimetres
    # nominal Fanuc model of the reference's script: 12 tokens, 6 joints
    mc = ModelCalibrator("TzRzTxRyTxRyTxTzRxRyRxTx",
                         joint_indices=[0, 1, 0, 1, 0, 1, 0, 0, 1, 1, 1, 0],
                         variables=["d_1", "q_1", "d_2", "q_2", "d_3", "q_3", "d_4", "d_5", "q_4", "q_5", "q_6", "d_6"],
                         link_lengths=[("d_1", 670), ("d_2", 312), ("d_3", 1075), ("d_5", 225), ("d_4", 1280),
                                       ("d_6", 215)],
                         offsets=np.array([0.0, -np.pi / 2, np.pi / 2, 0.0, 0.0, 0.0]),
                         directions=np.array([1, 1, -1, -1, -1, -1]))
    sequence, variables = mc.get_reduced_model()
    T_base = translation(-563.145620015451, -559.28130331986, 554.98685687043)
    T_tools = [translation(416.687065414667, -17.0660344931335, 59.6890106614009),
               translation(420.667529952858, 57.1117905464326, -15.7733691549466),
               translation(420.483792669375, -43.7121193657286, -42.4708914215651)]
    history = [mc.evaluate(qs, pts, T_base, T_tools)]
    for _ in range(iterations):
        T_base, T_tools = mc.estimate_tool_base(qs, pts)
        mc.estimate_parameters(qs, pts, T_base, T_tools)
        history.append(mc.evaluate(qs, pts, T_base, T_tools))
    if refine:
        mc.optimize(qs, pts, T_base, T_tools)
        history.append(mc.evaluate(qs, pts, T_base, T_tools))
    if verbose:
        print("reduced model:", "".join(sequence), f"({len(variables)} parameters)")
        for i, ev in enumerate(history):
            print(f"pass {i}: distance RMS {ev['distance_rms']:.6f} mm, max {ev['max_dist']:.6f} mm")
    return history


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
