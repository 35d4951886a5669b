#!/usr/bin/env python
"""Counterpart of the reference's `rr_synthetic_calibration.py` at batch scale: a planar RR arm whose
real first link is 2.1 (nominal 2.0) is observed through three markers on its end effector from a
tracker frame; N random joint vectors, Gaussian noise (sigma = 1 mm, seeded), then the reference's
loop -- estimate_tool_base, estimate_parameters, evaluate -- on the reduced model that
ModelCalibrator.get_reduced_model() derives for the chain RzTxRzTx.  Every pass over the N x 3
measurements is one kernel launch (the structure-generic calibration kernels: this model has 3
parameters, not the Fanuc's 18).

One iteration recovers the link error (0.1) to the noise floor (distance RMS = sigma sqrt(3)).  Further
iterations follow the reference's update rule to the letter (model_calibration.py:442-446): from the
second one on it blends the current ABSOLUTE parameters with the new CORRECTION,
pi <- (1 - s) pi + s dpi, which shrinks a converged estimate by the step s each time -- the reference's
own script shows the same drift.

  python examples/rr_synthetic_calibration_gpu.py [N] [iterations]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roboticmanipulators_b200.robots import RRRobot
from roboticmanipulators_b200.utils.model_calibration import ModelCalibrator


def translation(x, y, z):
    T = np.eye(4)
    T[:3, 3] = (x, y, z)
    return T


def main(n=200000, iterations=1, sigma=0.001, seed=0, verbose=True):
    dev = "cuda"
    E_T_p = [np.eye(4), translation(0.1, 0.1, 0.0), translation(0.1, -0.1, 0.0)]   # markers in the end frame
    F_T_B = translation(1.0, 0.5, 0.0)                                             # base in the tracker frame
    robot = RRRobot(lengths=(2.1, 1.0), save=False)                                # the REAL arm
    g = torch.Generator(device=dev).manual_seed(seed)
    qs = (torch.rand((n, 2), generator=g, device=dev, dtype=torch.float64) * 2 - 1) * np.pi
    B_T_E = robot.forward_kinematics_batch(qs)                                     # [n, 4, 4], one launch
    F = torch.tensor(F_T_B, device=dev)
    pts = torch.cat([(F @ B_T_E @ torch.tensor(E, device=dev))[:, :3, 3] for E in E_T_p], dim=1)
    pts = pts + sigma * torch.randn(pts.shape, generator=g, device=dev, dtype=torch.float64)

    mc = ModelCalibrator("RzTxRzTx", joint_indices=[1, 0, 1, 0], variables=["q_1", "L_1", "q_2", "L_2"],
                         link_lengths=[("L_1", 2.0), ("L_2", 1.0)], offsets=np.array([0.0, 0.0]),
                         directions=np.array([1, 1]), step=0.05)
    mc.verbose = False
    sequence, variables = mc.get_reduced_model()
    history = []
    for _ in range(iterations):
        T_base, T_tools = mc.estimate_tool_base(qs, pts)
        mc.estimate_parameters(qs, pts, T_base, T_tools)
        history.append(mc.evaluate(qs, pts, T_base, T_tools))
    if verbose:
        print("reduced model:", "".join(sequence), [str(v) for v in variables])
        for i, ev in enumerate(history):
            print(f"iteration {i}: distance RMS {ev['distance_rms']:.6f}, max {ev['max_dist']:.6f}")
        print("parameters:", np.round(np.asarray(mc._parameters, dtype=float), 6))
    return history, np.asarray(mc._parameters, dtype=float)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 200000, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
