#!/usr/bin/env python
"""Benchmark of the batched kinematics hot path (BASELINE.json metric:
"FK+Jacobian+IK poses/s (Fanuc165F, fp64)").

Headline step = one pass of the hot path over one batch of synthetic joint vectors resident in
HBM: rk_fanuc_fk_jac_ik8_f64 (pose + 6x6 Jacobian + det J, then all 8 Pieper branches + validity
of that pose; one fused kernel).  Default: 15 steps x 2^26 poses = 1.0e9 poses (BASELINE configs
[1]+[2] chunked for one GPU).  After the headline region the same run measures, and reports under
`e2e`, `workloads` and `latency_us`:
  * e2e       the public host-array API (Fanuc165F.fk_jac_ik_stream -> rk_fanuc_fk_jac_ik8_host_f64)
              on pinned host buffers, copies inside the timed region, next to a memcpy-only probe
              of the same bytes (the link roofline of that call);
  * workloads BASELINE configs[2] (IK only, 1 % unreachable targets), configs[3] (RRR trapezoid
              trajectories + FK + Jacobian per sample) and configs[4] (calibration J^T J / J^T r over
              1e8 measurements sharded over the ranks, NCCL all-reduce of the 190 packed doubles,
              18x18 solve), each with its own roofline and CPU baseline;
  * latency   wall time of the batched call at 1 / 100 / 10^4 poses.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload fk_jac_ik|fk_jac_ik_split|fk_jac|ik|calib|traj] [--batch B]
                  [--layout soa|aos] [--dtype f64|f32] [--no-extras] [--no-cpu]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "FK+Jacobian+IK poses/s (Fanuc165F, fp64)"
UNIT = "poses/s"
# identical in both arms (the driver compares the two `config` objects)
CONFIG = {"workload": "fanuc165f_fk_jac_det_ik8",
          "poses": "uniform in the joint limits (robots/fanuc.py:36-41), synthetic",
          "per_pose": "forward_kinematics + 6x6 Jacobian + det J + inverse_kinematics over the "
                      "canonical 8 (m, k, w) branches with validity", "dtype": "f64"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        rows = [l for (t, l) in self.lines if t0 <= t <= t1 + 0.06] or [l for _, l in self.lines]
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), smax.append(float(f[2])), power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------
# CPU arm: the oracle port (and, when baseline/_ref travelled, the real reference) on the box's
# host cores
# --------------------------------------------------------------------------------------
def _cpu_fk_jac_ik(args):
    seed, n = args
    from oracle import np_oracle as O
    c = O.FanucConst()
    rng = np.random.default_rng(seed)
    q = rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(n, 6))
    T = O.fanuc_fk(q, c)
    J = O.fanuc_jacobian(q, c)
    det = O.jac_det(J)
    qb, vb = O.fanuc_ik_branches(T, O.CANONICAL_8, c)
    return float(T[0, 0, 3] + det[0] + np.nansum(qb[0]))


def cpu_poses_per_s(n_total, procs):
    """Oracle FK + Jacobian + det + 8-branch IK over n_total poses on `procs` processes."""
    chunk = 20000
    tasks = [(i, min(chunk, n_total - i * chunk)) for i in range((n_total + chunk - 1) // chunk)]
    t0 = time.perf_counter()
    if procs <= 1:
        for t in tasks:
            _cpu_fk_jac_ik(t)
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            pool.map(_cpu_fk_jac_ik, tasks)
    dt = time.perf_counter() - t0
    return n_total / dt, dt


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the step on all host cores, same
    metric / unit / config.  The line's value is the NumPy restatement of the reference
    (oracle/np_oracle.py, pinned to the real reference by tests/golden: a vectorised port that is
    ~5e4 x FASTER than the original, i.e. the conservative baseline); when the unmodified reference
    travelled to this box (baseline/_ref, built by oracle/make_ref.py) a bounded sample of it is
    timed as well and reported under `cpu_baseline_reference`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 20000 * cores          # poses per step
    for _ in range(args.warmup):
        cpu_poses_per_s(min(sample, 20000 * cores), cores)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_poses_per_s(sample, cores)
        times.append(dt)
    total = sample * args.steps
    val = total / sum(times)
    desc = (f"{sample} uniform-in-limits poses per step x {args.steps} steps: oracle "
            f"fanuc_fk + fanuc_jacobian + det + 8-branch fanuc_ik, {cores} processes")
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": CONFIG,
        "run": {"batch_per_step": sample, "cpu_model": cpu_model(),
                "note": "value = NumPy restatement of the reference on all host cores"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:
        from oracle import make_ref
        if make_ref.available():
            ref = make_ref.time_reference(poses_per_core=8, procs=cores)
            ref["cpu_model"] = cpu_model()
            ref["port_speedup_over_reference"] = val / ref["value"]
            out["cpu_baseline_reference"] = ref
        else:
            out["cpu_baseline_reference"] = {"unavailable": "baseline/_ref not present on this box "
                                             "(built by oracle/make_ref.py where /root/reference exists)"}
    except Exception as e:  # the port's number above stands on its own
        out["cpu_baseline_reference"] = {"unavailable": f"{type(e).__name__}: {e}"}
    print(json.dumps(out), file=args._out, flush=True)


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def measure(wl, steps, warmup, dist, rank, sample_clocks=False, local=0):
    """W warm-up steps, then exactly `steps` timed steps between barrier + synchronize on both
    sides; CUDA events on the launching stream; max over ranks.  Returns (ms_total, clocks)."""
    import torch
    for _ in range(warmup):
        wl.step(-1)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler = ClockSampler(local) if sample_clocks and rank == 0 else None
    if sampler is not None:
        sampler.start()
        time.sleep(0.12)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()      # rank 0 slept while starting the clock sampler: line the ranks up again
        torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    for s in range(steps):
        wl.step(s)
    ev1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    if dist is not None:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler is not None else None
    return ms, clocks


def roofline_of(wl, kern, peak, peak_src, fp64):
    """Roofline record of the dominant kernel: HBM (algorithmic bytes / event time against the
    measured copy bandwidth) unless the workload declares the kernel FP64-bound, in which case the
    algorithmic flop against the DFMA rate measured in this run; both views are kept."""
    dom = max(kern, key=lambda k: kern[k])
    alg_bytes = wl.alg_bytes[dom] * wl.units_per_step
    achieved = alg_bytes / (kern[dom] * 1e-3) / 1e9
    src = wl.traffic_source.get(dom) if getattr(wl, "traffic_source", None) else None
    roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": wl.traffic.get(dom), "traffic_source": src,
            "peak_source": peak_src, "ms_per_launch": kern[dom],
            "algorithmic_bytes_per_unit": wl.alg_bytes[dom]}
    if getattr(wl, "alg_flop", None) and dom in wl.alg_flop and fp64:
        tf = wl.alg_flop[dom] * wl.units_per_step / (kern[dom] * 1e-3) / 1e12
        roof = {"bound": "fp64", "kernel": dom, "achieved": tf, "peak": fp64["tflops"],
                "unit": "TFLOP/s", "frac": tf / fp64["tflops"], "traffic": wl.traffic.get(dom),
                "traffic_source": src, "peak_source": "measured (rk_probe_dfma, this run)",
                "ms_per_launch": kern[dom], "algorithmic_flop_per_unit": wl.alg_flop[dom],
                "hbm_frac": achieved / peak}
    return roof


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=15)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fk_jac_ik",
                    choices=["fk_jac_ik", "fk_jac_ik_split", "fk_jac", "ik", "calib", "traj"])
    ap.add_argument("--batch", type=int, default=1 << 26, help="poses per GPU per step")
    ap.add_argument("--e2e-batch", type=int, default=1 << 21)
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--layout", default="soa", choices=["soa", "aos"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="FK/IK arithmetic type")
    ap.add_argument("--ld-pad", type=int, default=0, help="extra elements between SOA planes")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-extras", action="store_true",
                    help="headline + e2e only (skip the IK / trajectory / calibration workloads)")
    ap.add_argument("--cpu-sample", type=int, default=500000)
    args = ap.parse_args()
    # Libraries (NCCL's version banner, torchrun) may write to stdout; the contract is ONE JSON
    # line there, so everything else is sent to stderr and the line is written to the real fd.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    args._out = real_stdout
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from roboticmanipulators_b200 import _lib
    import bench_workloads as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.load()
    peak, peak_src = load_peaks()
    warm = max(args.warmup, 3)

    wl = W.make(args, rank, world)
    ms, clocks = measure(wl, args.steps, warm, dist, rank, sample_clocks=True, local=local)
    kern = wl.kernel_times_ms()          # {name: mean ms per launch} from events in the region
    fp64 = W.fp64_probe()
    e2e = wl.e2e(args, dist)             # host buffers, copies inside the timed region
    headline_is_fanuc = args.workload.startswith("fk") or args.workload == "ik"

    units = wl.units_per_step * world
    value = units * args.steps / (ms * 1e-3)
    out = {
        "metric": METRIC if headline_is_fanuc else wl.metric,
        "value": value, "unit": wl.unit, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": getattr(wl, "sfx", "f64"), "data": "synthetic",
        "config": CONFIG if args.workload == "fk_jac_ik" and args.dtype == "f64" else wl.config,
        "run": wl.config, "clocks": clocks,
        "roofline": roofline_of(wl, kern, peak, peak_src, fp64),
        "kernels": {k: {"ms_per_launch": v,
                        "units_per_s": wl.units_per_step / (v * 1e-3),
                        "achieved_gbs": wl.alg_bytes[k] * wl.units_per_step / (v * 1e-3) / 1e9,
                        "hbm_frac": wl.alg_bytes[k] * wl.units_per_step / (v * 1e-3) / 1e9 / peak,
                        "share_of_step": v / sum(kern.values())}
                    for k, v in kern.items()},
        "e2e": e2e,
        "gpu_launches": wl.launches_per_step * args.steps,
        "fp64": fp64,
    }
    if not args.no_cpu and rank == 0:
        procs = 1
        v, dt = cpu_poses_per_s(args.cpu_sample, procs) if args.workload.startswith("fk_jac_ik") \
            else wl.cpu(args)
        out["cpu_baseline"] = {"value": v, "unit": wl.unit, "cores": procs, "kind": "port",
                               "sample": wl.cpu_sample_desc(args, dt)}
    if hasattr(wl, "latency_us"):
        out["latency_us"] = wl.latency_us()
    del wl
    gc.collect()
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, same run, same box (each: W >= 3 warm-ups, 5 timed steps) ----
    if args.workload == "fk_jac_ik" and not args.no_extras:
        out["workloads"] = {}
        for name in ("ik", "traj", "calib"):
            a = argparse.Namespace(**vars(args))
            a.workload, a.batch, a.steps = name, 1 << 26, 5
            w2 = W.make(a, rank, world)
            ms2, _ = measure(w2, a.steps, 3, dist, rank)
            k2 = w2.kernel_times_ms()
            rec = {"metric": w2.metric, "unit": w2.unit,
                   "value": w2.units_per_step * world * a.steps / (ms2 * 1e-3),
                   "ms_per_step": ms2 / a.steps, "steps": a.steps, "warmup": 3, "config": w2.config,
                   "roofline": roofline_of(w2, k2, peak, peak_src, fp64),
                   "kernels_ms": k2, "gpu_launches": w2.launches_per_step * a.steps}
            if hasattr(w2, "extra_record"):
                rec.update(w2.extra_record(dist))
            if not args.no_cpu and rank == 0:
                v, dt = w2.cpu(a)
                rec["cpu_baseline"] = {"value": v, "unit": w2.unit, "cores": 1, "kind": "port",
                                       "sample": w2.cpu_sample_desc(a, dt)}
            out["workloads"][name] = rec
            del w2
            gc.collect()
            torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(out), file=args._out, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
