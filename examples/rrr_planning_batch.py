#!/usr/bin/env python
"""Batched counterpart of the reference's `rrr_robot_planning.py`: for N (start, goal) pairs at
once -- equalised trapezoidal joint trajectories with the per-sample pose and Jacobian of the RRR
arm, and Cartesian straight lines with per-sample inverse kinematics and joint rates
dq = J(q)^-1 dx -- each family in one kernel launch.

  python examples/rrr_planning_batch.py [N]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roboticmanipulators_b200.robots import RRRRobot
from roboticmanipulators_b200.utils.trajectory_generator import TrajectoryGenerator


def main(n_traj=4096, n=50, seed=0, verbose=True):
    tg = TrajectoryGenerator(1, 20, 1, 10, control_freq=10)     # dq_max, ddq_max, dx_max, ddx_max
    robot = RRRRobot()
    rng = np.random.default_rng(seed)
    dev = "cuda"

    # joint space: q_0 -> q_f, all joints finish together; FK + Jacobian at every sample
    q_0 = torch.tensor(rng.uniform(-1.0, 1.0, size=(n_traj, 3)), device=dev)
    q_f = torch.tensor(rng.uniform(-2.5, 2.5, size=(n_traj, 3)), device=dev)
    q, dq, ddq, t, length, T, J = tg.generate_p2p_trajectory_batch(q_0, q_f, n, robot=robot)
    last = (length.long() - 1).clamp(min=0)
    rows = torch.arange(n_traj, device=dev)
    reach = (q[rows, last] - q_f).abs().max()

    # Cartesian space: straight line p_0 -> p_f, IK (m = k = 1) and joint rates per sample
    p_0 = torch.tensor(rng.uniform([0.6, -0.5, 0.4], [1.2, 0.5, 1.6], size=(n_traj, 3)), device=dev)
    p_f = torch.tensor(rng.uniform([0.6, -0.5, 0.4], [1.2, 0.5, 1.6], size=(n_traj, 3)), device=dev)
    x, dx, tx, qj, dqj, valid, len_x = tg.generate_lin_trajectory_batch(p_0, p_f, n, robot=robot)
    live = valid.bool() & (torch.arange(n, device=dev)[None, :] < len_x[:, None])
    # the joint path reproduces the line: FK(q) == x wherever the IK succeeded
    Tq = robot.forward_kinematics_batch(qj.reshape(-1, 3)).reshape(n_traj, n, 4, 4)
    line_err = ((Tq[:, :, :3, 3] - x).abs().amax(dim=2) * live).max()
    out = {"trajectories": n_traj, "samples": n, "p2p_mean_length": float(length.float().mean()),
           "p2p_goal_error": float(reach), "lin_mean_length": float(len_x.float().mean()),
           "lin_ik_success": float(live.float().sum() / max(1, int(len_x.sum()))),
           "lin_tracking_error": float(line_err), "max_joint_rate": float((dqj.abs() * live[..., None]).max())}
    if verbose:
        for k, v in out.items():
            print(f"{k:22s} {v}")
    return out


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 4096)
