#!/usr/bin/env python
"""Device-resident timing of the hot-path entry points that bench.py's workloads do not cover
(SURVEY section 8 rows A3-A9, A11 and the generic chain; the arms' IK is position-only and is
charged 24 B of input, the translation column it reads): RRR/RR FK+Jacobian and IK, the token
interpreter, the trapezoid / quintic samplers, Cartesian line + IK, estimate_tool_base's normal
equations and evaluate.  Inputs are synthetic and resident in HBM, every call goes through the
public Python API (one C-ABI launch), timing is CUDA events around `reps` calls after 2 warm-up
calls, outputs are larger than L2.  Prints ONE JSON object: {row: {ms, units_per_s,
algorithmic_bytes_per_unit, achieved_gbs, hbm_frac}}.

  python tools/bench_rows.py [--reps 5] > profiles/rNN_rows.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--scale", type=float, default=1.0, help="scale every batch size")
    ap.add_argument("--warmup", type=int, default=2, help="0 with --reps 1 under ncu: one launch per row")
    args = ap.parse_args()
    import torch
    from roboticmanipulators_b200 import _lib
    from roboticmanipulators_b200.robots import RRRRobot, RRRobot
    from roboticmanipulators_b200.robots.fanuc import Fanuc165F
    from roboticmanipulators_b200.utils.robo_math import Transformation
    from roboticmanipulators_b200.utils.jacobians import JacobianCalculator
    from roboticmanipulators_b200.utils.trajectory_generator import TrajectoryGenerator
    from roboticmanipulators_b200.utils.trapezoidal_generator import TrapezoidalGenerator
    from roboticmanipulators_b200.utils.polynomial_generator import PolynomialGenerator
    import bench_workloads as W
    _lib.load()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(11)
    f64 = torch.float64

    def rnd(shape, lo=-1.0, hi=1.0):
        return lo + (hi - lo) * torch.rand(shape, generator=g, device=dev, dtype=f64)

    out = {}

    def timed(name, units, nbytes, fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        gbs = units * nbytes / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "units": units, "units_per_s": units / (ms * 1e-3),
                     "algorithmic_bytes_per_unit": nbytes, "achieved_gbs": gbs, "hbm_frac": gbs / peak}
        torch.cuda.empty_cache()

    S = args.scale
    # --- A3-A5: RRR / RR arms (SOA views: [3, n] storage)
    n = int((1 << 25) * S)
    rrr, rr = RRRRobot(), RRRobot()
    q3 = rnd((3, n), -np.pi, np.pi).t()
    timed("rrr_fk_jac_soa", n, 24 + 128 + 144, lambda: rrr.forward_kinematics_batch(q3, jacobian=True, layout="soa"))
    q3a = q3.contiguous()
    timed("rrr_fk_jac_aos", n, 24 + 128 + 144, lambda: rrr.forward_kinematics_batch(q3a, jacobian=True, layout="aos"))
    T3 = rrr.forward_kinematics_batch(q3, layout="soa")
    timed("rrr_ik4_soa", n, 24 + 4 * 24 + 4, lambda: rrr.inverse_kinematics_batch(T3, layout="soa"))
    del T3
    q2 = rnd((2, n), -np.pi, np.pi).t()
    timed("rr_fk_jac_soa", n, 16 + 128 + 96, lambda: rr.forward_kinematics_batch(q2, jacobian=True, layout="soa"))
    T2 = rr.forward_kinematics_batch(q2, layout="soa")
    timed("rr_ik2_soa", n, 24 + 2 * 16 + 2, lambda: rr.inverse_kinematics_batch(T2, layout="soa"))
    # FP32 variants (float32 input -> FP32 kernels, half the bytes)
    f32 = torch.float32
    q3f = q3.t().to(f32).contiguous().t()
    timed("rrr_fk_jac_soa_f32", n, (24 + 128 + 144) // 2,
          lambda: rrr.forward_kinematics_batch(q3f, jacobian=True, layout="soa"))
    T3f = rrr.forward_kinematics_batch(q3f, layout="soa")
    timed("rrr_ik4_soa_f32", n, 12 + 4 * 12 + 4, lambda: rrr.inverse_kinematics_batch(T3f, layout="soa"))
    del T3f, q3f
    q2f = q2.t().to(f32).contiguous().t()
    timed("rr_fk_jac_soa_f32", n, (16 + 128 + 96) // 2,
          lambda: rr.forward_kinematics_batch(q2f, jacobian=True, layout="soa"))
    T2f = rr.forward_kinematics_batch(q2f, layout="soa")
    timed("rr_ik2_soa_f32", n, 12 + 2 * 8 + 2, lambda: rr.inverse_kinematics_batch(T2f, layout="soa"))
    del T2f, q2f
    del T2, q2, q3, q3a
    # --- A6/A7: generic token interpreter on the Fanuc sequence (13 tokens, 6 joints)
    n = int((1 << 24) * S)
    seq = "TzRzTzTxRyTxRyTzTxRxRyRxTx"
    ls = [346.0, 324.0, 312.0, 1075.0, 225.0, 1280.0, 215.0]
    qm = rnd((n, 6), -3.0, 3.0)
    vals = torch.empty((13, n), dtype=f64, device=dev)
    li = qi = 0
    for k in range(13):
        if seq[2 * k] == "T":
            vals[k] = ls[li]
            li += 1
        else:
            vals[k] = qm[:, qi]
            qi += 1
    vals = vals.t()
    jind = [1 if seq[2 * k] == "R" else 0 for k in range(13)]
    jc = JacobianCalculator(seq, jind)
    timed("chain_fk_jac_fanuc_seq_soa", n, 13 * 8 + 128 + 288,
          lambda: jc.calculate_numerically_batch(vals, layout="soa"))
    timed("chain_fk_fanuc_seq_soa", n, 13 * 8 + 128, lambda: Transformation.batch(seq, vals, layout="soa"))
    valsf = vals.t().to(f32).contiguous().t()
    timed("chain_fk_jac_fanuc_seq_soa_f32", n, (13 * 8 + 128 + 288) // 2,
          lambda: jc.calculate_numerically_batch(valsf, layout="soa"))
    timed("chain_fk_fanuc_seq_soa_f32", n, (13 * 8 + 128) // 2, lambda: Transformation.batch(seq, valsf, layout="soa"))
    del valsf
    # dq = J^-1 dx inside the kernel (TrajectoryGenerator.get_dq_from_dx for any robot)
    dx6 = rnd((6, n)).t()
    vals_a, dx6_a = vals.contiguous(), dx6.contiguous()     # this wrapper takes sample-major arrays
    timed("chain_dq_from_dx_6x6", n, 13 * 8 + 48 + 48 + 1, lambda: jc.dq_from_dx_batch(vals_a, dx6_a))
    del vals_a, dx6_a
    fan = Fanuc165F()
    qf = rnd((6, n), -1.0, 1.0).t()
    timed("fanuc_dq_from_dx", n, 48 + 48 + 48 + 1, lambda: fan.dq_from_dx_batch(qf, dx6, layout="soa"))
    del vals, qm, dx6, qf
    # --- A8/A9: trajectories
    tg = TrajectoryGenerator(1, 20, 1, 10, control_freq=10)
    nt, ns = int((1 << 15) * S), 1000
    a6, b6 = rnd((nt, 6), -np.pi, np.pi), rnd((nt, 6), -np.pi, np.pi)
    timed("trap_p2p_6joints", nt * ns, 3 * 48 + 8, lambda: tg.generate_p2p_trajectory_batch(a6, b6, ns))
    a3, b3 = a6[:, :3].contiguous(), b6[:, :3].contiguous()
    timed("rrr_p2p_fk_f32", nt * ns, (3 * 24 + 8 + 128 + 144) // 2 + 4,
          lambda: tg.generate_p2p_trajectory_batch(a3, b3, ns, robot=rrr, dtype=f32))
    p0 = torch.stack([rnd((nt,), 0.5, 1.2), rnd((nt,), -0.6, 0.6), rnd((nt,), 0.2, 1.0)], 1)
    pf = torch.stack([rnd((nt,), 0.5, 1.2), rnd((nt,), -0.6, 0.6), rnd((nt,), 0.2, 1.0)], 1)
    timed("rrr_lin_ik_dq", nt * ns, 24 + 24 + 8 + 24 + 24 + 1,
          lambda: tg.generate_lin_trajectory_batch(p0, pf, ns, robot=rrr))
    tz = TrapezoidalGenerator(1, 20, control_freq=10)
    n = int((1 << 26) * S)
    qa, qb = rnd((n,), -np.pi, np.pi), rnd((n,), -np.pi, np.pi)
    timed("trap_coefficients", n, 32, lambda: tz.generate_coefficients_batch(qa, qb))
    npf = int((1 << 16) * S)
    t1, tau = tz.generate_coefficients_batch(qa[:npf], qb[:npf])
    v, acc = torch.full((npf,), 1.0, dtype=f64, device=dev), torch.full((npf,), 20.0, dtype=f64, device=dev)
    timed("trap_sample", npf * ns, 32, lambda: TrapezoidalGenerator.sample_batch(qa[:npf], v, acc, t1, tau, ns))
    bi, bf = rnd((npf, 3)), rnd((npf, 3))
    timed("poly5_profiles", npf * ns, 24, lambda: PolynomialGenerator.profiles_batch(bi, bf, 2.0, 0.0, ns))
    del qa, qb
    # --- A10/A11: the rest of a calibration iteration
    class A:
        batch = int((1 << 23) * S)
        ld_pad = 0
    cw = W.CalibWorkload(A, 0, 1)
    m = cw.m
    # the C-ABI call itself (ModelCalibrator.estimate_tool_base adds a host solve + SymPy composition)
    import ctypes as C
    mc, lib = cw.mc, _lib.load()
    model = mc._compile(np.zeros(18), None, [np.eye(4)] * 3)
    tb_out = torch.empty(15 * 15 + 15, dtype=f64, device=dev)
    ws = mc._workspace(dev)
    timed("calib_tool_base_eq", m, 120, lambda: _lib.check(lib.rk_calib_tool_base_eq_f64(
        C.byref(model), _lib.ptr(cw.q), _lib.ptr(cw.pts), m, _lib.ptr(tb_out), _lib.ptr(ws), ws.numel(),
        _lib.stream_ptr(dev))))
    timed("calib_evaluate", m, 120, lambda: cw.mc.residual_sums(cw.q, cw.pts, cw.T_base, cw.T_tools, parameters=np.zeros(18), process_group=False))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
