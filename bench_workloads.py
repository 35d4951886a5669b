"""Workloads timed by bench.py.  Each drives the C ABI (include/rk_b200.h) directly on
preallocated device buffers for the device-resident number, and the public Python API
with pinned host buffers for the end-to-end number."""
import ctypes as C
import os
import time

import numpy as np
import torch

from roboticmanipulators_b200 import _lib
from roboticmanipulators_b200.robots.fanuc import Fanuc165F


def ncu_traffic_per_unit():
    """DRAM bytes per unit of each kernel from the committed `ncu --set full` summaries
    (profiles/*_ncu_full_summary.json, written by tools/ncu_summary.py; the newest round wins).
    Returns {kernel key: (bytes per unit, "profiles/<file>")}: the bench line's `roofline.traffic`
    is a recorded capture (a number under ncu is never a bench value), `traffic_source` says which."""
    import glob
    import json
    import os
    out = {}
    root = os.path.dirname(os.path.abspath(__file__))
    for path in sorted(glob.glob(os.path.join(root, "profiles", "*_ncu_full_summary.json"))):
        try:
            for rec in json.load(open(path)):
                if "traffic_bytes_per_unit" not in rec:
                    continue
                for key, pat in (("fanuc_fk_jac", "fanuc_fk_jac_tma"), ("fanuc_ik8", "fanuc_ik8"),
                                 ("fanuc_fk_jac_ik8", "fanuc_fk_jac_ik8"),
                                 ("calib_normal_eq", "calib_normal_eq"), ("rrr_p2p_fk", "traj_kernel<3, 1>")):
                    if pat in rec["kernel"]:
                        out[key] = (rec["traffic_bytes_per_unit"], "profiles/" + os.path.basename(path))
        except (OSError, ValueError):
            pass
    return out


def traffic_for(units):
    t = ncu_traffic_per_unit()
    return ({k: v[0] * units for k, v in t.items()},
            {k: f"{v[1]} (ncu --set full dram__bytes_read+write per unit x units of this launch)"
             for k, v in t.items()})


def fp64_probe():
    """Measured FP64 FMA peak (dependent-free DFMA loop on all SMs)."""
    lib = _lib.load()
    sink = torch.zeros(8, dtype=torch.float64, device="cuda")
    total = C.c_int64(0)
    st = _lib.stream_ptr()
    _lib.check(lib.rk_probe_dfma(20000, _lib.ptr(sink), C.byref(total), st))
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.rk_probe_dfma(300000, _lib.ptr(sink), C.byref(total), st))
        e1.record()
        torch.cuda.synchronize()
        best = max(best, total.value / (e0.elapsed_time(e1) * 1e-3))
    return {"dfma_lane_per_s": best, "tflops": 2 * best / 1e12,
            "how": "rk_probe_dfma: 8 independent FMA chains/thread, 148x8 CTAs x 256, best of 3"}


def finish_e2e(rec, dist, unit):
    """seconds -> max over ranks; value = units of all ranks / that time."""
    world = 1
    if dist is not None:
        world = dist.get_world_size()
        t = torch.tensor([rec["seconds"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rec["seconds"] = float(t.item())
    return {"value": rec["units"] * world / rec["seconds"], "unit": unit,
            "h2d_bytes_per_step": rec["h2d"], "d2h_bytes_per_step": rec["d2h"],
            "batch_per_step": rec["batch"], "steps": rec["steps"], "seconds": rec["seconds"],
            "api": rec.get("api")}


class FanucWorkload:
    """fk+jac+det and ik8 over `batch` poses per step, SOA or AOS planes.  `fk_jac_ik` is the single
    call rk_fanuc_fk_jac_ik8 (one fused kernel for SOA); `fk_jac_ik_split` (and AOS) are the two calls
    rk_fanuc_fk_jac then rk_fanuc_ik on the pose planes K1 wrote."""
    unit = "poses/s"
    metric = "FK+Jacobian+IK poses/s (Fanuc165F, fp64)"

    def __init__(self, args, rank, world):
        self.args = args
        self.mode = "fk_jac_ik" if args.workload == "fk_jac_ik_split" else args.workload
        self.n = n = args.batch
        self.units_per_step = n
        self.robot = Fanuc165F()
        self.st = self.robot._struct()
        self.lib = _lib.load()
        self.soa = args.layout == "soa"
        self.flags = _lib.RK_SOA if self.soa else _lib.RK_AOS
        self.f32 = getattr(args, "dtype", "f64") == "f32"
        self.sfx = "f32" if self.f32 else "f64"
        self.fused = args.workload == "fk_jac_ik" and self.soa
        dev = torch.device("cuda")
        # distinct input batches (so 15 steps really are 1e9 different poses), capped at 48 GB
        nbuf = 1 if self.mode == "ik" else max(1, min(args.steps, int(48e9 // (48 * n))))
        lim = torch.tensor(np.asarray(self.robot.qs_lim_rad), device=dev)
        self.q = []
        for b in range(nbuf):
            g = torch.Generator(device=dev).manual_seed(1 + 1000 * rank + b)
            u = torch.rand((6, n), generator=g, device=dev, dtype=torch.float64)
            u = lim[:, :1] + u * (lim[:, 1:] - lim[:, :1])
            if self.soa and getattr(args, "ld_pad", 0):
                up = torch.empty((6, n + args.ld_pad), dtype=torch.float64, device=dev)
                up[:, :n] = u
                u = up
            if self.f32:
                u = u.float()
            self.q.append(u if self.soa else u.t().contiguous())
            del u
        f64 = torch.float32 if self.f32 else torch.float64
        self.ld = n + getattr(args, "ld_pad", 0) if self.soa else n
        shape = (lambda c: (c, self.ld)) if self.soa else (lambda c: (n, c))
        self.T = torch.empty(shape(16), dtype=f64, device=dev)
        self.J = torch.empty(shape(36), dtype=f64, device=dev) if self.mode != "ik" else None
        self.det = torch.empty(n, dtype=f64, device=dev) if self.mode != "ik" else None
        if self.mode != "fk_jac":
            self.qik = torch.empty(shape(48), dtype=f64, device=dev)
            self.valid = torch.empty(shape(8), dtype=torch.uint8, device=dev)
        if self.mode == "ik":      # targets = FK of the first batch, 1 % scaled out of reach (configs[2])
            self._k1(self.q[0], with_j=False)
            if self.soa:
                self.T[[3, 7, 11], ::100] *= 3.0
            else:
                self.T[::100, [3, 7, 11]] *= 3.0
            torch.cuda.synchronize()
        self.events = []
        self.alg_bytes = {"fanuc_fk_jac": 236, "fanuc_ik8": 264, "fanuc_fk_jac_ik8": 432} if self.f32 else \
            {"fanuc_fk_jac": 472, "fanuc_ik8": 520, "fanuc_fk_jac_ik8": 864}
        self.traffic, self.traffic_source = ({}, {}) if self.f32 else traffic_for(n)
        self.launches_per_step = 1 if self.fused else {"fk_jac_ik": 2, "fk_jac": 1, "ik": 1}[self.mode]
        self.config = {
            # same workload name as the reference arm (bench.py run_reference)
            "workload": {"fk_jac_ik": "fanuc165f_fk_jac_det_ik8",
                         "fk_jac": "fanuc165f_fk_jac_det", "ik": "fanuc165f_ik8_1pct_unreachable"}[self.mode],
            "calls_per_step": "rk_fanuc_fk_jac_ik8 (one fused kernel)" if self.fused else
                              {"fk_jac_ik": "rk_fanuc_fk_jac then rk_fanuc_ik on the planes it wrote",
                               "fk_jac": "rk_fanuc_fk_jac", "ik": "rk_fanuc_ik"}[self.mode],
            "baseline_config": "configs[1] (+configs[2] IK on the same poses), chunked: "
                               f"{args.steps} steps x {n} poses per GPU",
            "batch_per_gpu_per_step": n, "layout": args.layout, "dtype": self.sfx,
            "distinct_input_batches": nbuf,
            "l2": "inputs larger than L2 (each step streams >= 3.2 GB of fresh q and "
                  "overwrites >= 28 GB of outputs)",
            "sharding": "independent poses, no collective",
        }

    def _k1(self, q, with_j=True):
        st = _lib.stream_ptr()
        _lib.check(getattr(self.lib, "rk_fanuc_fk_jac_" + self.sfx)(
            C.byref(self.st), _lib.ptr(q), self.n, self.ld, self.flags, _lib.ptr(self.T),
            _lib.ptr(self.J) if with_j else None, _lib.ptr(self.det) if with_j else None, st))

    def _k2(self):
        st = _lib.stream_ptr()
        _lib.check(getattr(self.lib, "rk_fanuc_ik_" + self.sfx)(
            C.byref(self.st), _lib.ptr(self.T), self.n, self.ld, self.flags, None, 8,
            _lib.ptr(self.qik), _lib.ptr(self.valid), st))

    def _k12(self, q):
        st = _lib.stream_ptr()
        _lib.check(getattr(self.lib, "rk_fanuc_fk_jac_ik8_" + self.sfx)(
            C.byref(self.st), _lib.ptr(q), self.n, self.ld, self.flags, _lib.ptr(self.T),
            _lib.ptr(self.J), _lib.ptr(self.det), _lib.ptr(self.qik), _lib.ptr(self.valid), st))

    def step(self, s):
        timed = s >= 0
        q = self.q[max(s, 0) % len(self.q)]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timed else None
        if timed:
            ev[0].record()
        if self.fused:
            self._k12(q)
            if timed:
                ev[2].record()
                self.events.append(ev)
            return
        if self.mode != "ik":
            self._k1(q)
        if timed:
            ev[1].record()
        if self.mode != "fk_jac":
            self._k2()
        if timed:
            ev[2].record()
            self.events.append(ev)

    def kernel_times_ms(self):
        out = {}
        if self.fused:
            return {"fanuc_fk_jac_ik8": float(np.mean([e[0].elapsed_time(e[2]) for e in self.events]))}
        if self.mode != "ik":
            out["fanuc_fk_jac"] = float(np.mean([e[0].elapsed_time(e[1]) for e in self.events]))
        if self.mode != "fk_jac":
            out["fanuc_ik8"] = float(np.mean([e[1].elapsed_time(e[2]) for e in self.events]))
        return out

    def e2e(self, args, dist=None):
        """The public host-array API (Fanuc165F.fk_jac_ik_stream -> rk_fanuc_fk_jac_ik8_host_f64) on
        reference-shaped (pose-major) PINNED HOST buffers: every step uploads its joint vectors and
        downloads every result inside the timed region; wall clock between barriers, max over
        ranks.  Next to it the link roofline of that call: the same bytes moved by memcpy only
        (upload and download concurrently, all ranks at once)."""
        world = dist.get_world_size() if dist is not None else 1
        n = min(args.e2e_batch, self.n)
        steps = max(args.e2e_steps, 30)
        dt_t = torch.float32 if self.f32 else torch.float64
        esz = 4 if self.f32 else 8
        rng = np.random.default_rng(7 + 13 * int(os.environ.get("RANK", "0")))
        lim = np.asarray(self.robot.qs_lim_rad)
        hq = torch.empty((n, 6), dtype=dt_t, pin_memory=True)
        hq.copy_(torch.from_numpy(rng.uniform(lim[:, 0], lim[:, 1], size=(n, 6))))
        outs = {"T": torch.empty((n, 4, 4), dtype=dt_t, pin_memory=True),
                "J": torch.empty((n, 6, 6), dtype=dt_t, pin_memory=True),
                "det": torch.empty(n, dtype=dt_t, pin_memory=True),
                "q_ik": torch.empty((n, 8, 6), dtype=dt_t, pin_memory=True),
                "valid": torch.empty((n, 8), dtype=torch.uint8, pin_memory=True)}
        full = ("T", "J", "det", "q_ik", "valid")

        def timed(fn, reps):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            sec = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([sec], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sec = float(t.item())
                dist.barrier()
            return sec

        sec = timed(lambda: self.robot.fk_jac_ik_stream(hq, want=full, out=outs), steps)
        h2d = n * 6 * esz
        d2h = n * ((16 + 36 + 1 + 48) * esz + 8)
        # the same call shipping only what an IK user needs (q_ik + valid: 392 B per pose)
        sec_ik = timed(lambda: self.robot.fk_jac_ik_stream(hq, want=("q_ik", "valid"), out=outs), steps)
        # link roofline: memcpy only, same buffers, both directions at once
        dq = torch.empty((n, 6), dtype=dt_t, device="cuda")
        douts = {k: torch.empty(v.shape, dtype=v.dtype, device="cuda") for k, v in outs.items()}
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

        def link(names):
            with torch.cuda.stream(s_up):
                dq.copy_(hq, non_blocking=True)
            with torch.cuda.stream(s_dn):
                for k in names:
                    outs[k].copy_(douts[k], non_blocking=True)
            s_up.synchronize(), s_dn.synchronize()

        sec_link = timed(lambda: link(full), steps)
        sec_link_ik = timed(lambda: link(("q_ik", "valid")), steps)
        # BASELINE config 2 end to end: targets on the host -> the 8 canonical IK answers on the host
        # (Fanuc165F.inverse_kinematics_stream: 128 B up, 392 B down per target), and its link bound
        hT = torch.empty((n, 4, 4), dtype=dt_t, pin_memory=True)
        hT.copy_(self.robot.forward_kinematics_batch(hq.cuda(), layout="aos"))
        ik_out = {"q_ik": outs["q_ik"], "valid": outs["valid"]}
        sec_iks = timed(lambda: self.robot.inverse_kinematics_stream(hT, out=ik_out), steps)
        dT = torch.empty((n, 4, 4), dtype=dt_t, device="cuda")

        def link_iks():
            with torch.cuda.stream(s_up):
                dT.copy_(hT, non_blocking=True)
            with torch.cuda.stream(s_dn):
                for k in ("q_ik", "valid"):
                    outs[k].copy_(douts[k], non_blocking=True)
            s_up.synchronize(), s_dn.synchronize()

        sec_link_iks = timed(link_iks, steps)
        total = n * steps * world
        return {
            "value": total / sec, "unit": self.unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "batch_per_step": n, "steps": steps, "seconds": sec,
            "api": "Fanuc165F.fk_jac_ik_stream (rk_fanuc_fk_jac_ik8_host): pinned host arrays, chunks of "
                   "65536 poses on 3 streams (upload / kernels / download), every result shipped",
            "link": {"what": "memcpy only: the same pinned buffers, upload and download concurrently, all "
                             "ranks at once -- the roofline of the call above",
                     "value": total / sec_link, "unit": self.unit, "frac": sec_link / sec,
                     "gbs_per_gpu_both_ways": (h2d + d2h) * steps / sec_link / 1e9},
            "selected_outputs": {"want": ["q_ik", "valid"], "value": total / sec_ik,
                                 "d2h_bytes_per_step": n * (48 * esz + 8),
                                 "link_value": total / sec_link_ik, "frac": sec_link_ik / sec_ik},
            "ik_stream": {"api": "Fanuc165F.inverse_kinematics_stream (rk_fanuc_ik8_host): targets T on the "
                                 "host in, q_ik + valid on the host out",
                          "value": total / sec_iks, "unit": "targets/s", "h2d_bytes_per_step": n * 16 * esz,
                          "d2h_bytes_per_step": n * (48 * esz + 8), "link_value": total / sec_link_iks,
                          "frac": sec_link_iks / sec_iks},
        }

    def latency_us(self):
        """Wall time per call of the batched public API on device-resident inputs (launch-side cost:
        no synchronisation inside the loop), and of the scalar drop-in call (upload, kernel,
        download, synchronise)."""
        out = {}
        for n in (1, 100, 10000):
            q = torch.zeros((n, 6), dtype=torch.float64, device="cuda") + 0.1
            for fn_name, fn in (("fk_jac_ik_batch", lambda: self.robot.fk_jac_ik_batch(q, layout="aos")),
                                ("forward_kinematics_batch", lambda: self.robot.forward_kinematics_batch(q, layout="aos"))):
                for _ in range(20):
                    fn()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(200):
                    fn()
                t1 = time.perf_counter()
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                out[f"{fn_name}_{n}_poses"] = {"issue": round((t1 - t0) / 200 * 1e6, 2),
                                               "complete": round((t2 - t0) / 200 * 1e6, 2)}
        # scalar drop-in calls: every call is a full round trip (upload, kernel, download, sync), so a
        # GPU that idled through the CPU leg above shows its wake-up in the first calls: warm up, then
        # report the median of the per-call times
        def median_us(fn, warm=10, reps=50):
            for _ in range(warm):
                fn()
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            return round(float(np.median(ts)) * 1e6, 1)
        out["scalar_forward_kinematics"] = median_us(lambda: self.robot.forward_kinematics([0.1] * 6, plot=False))
        T = self.robot.forward_kinematics([0.1] * 6, plot=False)
        out["scalar_inverse_kinematics"] = median_us(lambda: self.robot.inverse_kinematics(T, m=-1, k=1, w=1))
        return out

    def cpu(self, args):
        """CPU leg of the non-headline Fanuc workloads (the headline uses bench.cpu_poses_per_s)."""
        from oracle import np_oracle as O
        c = O.FanucConst()
        rng = np.random.default_rng(0)
        n = 200000
        q = rng.uniform(c.lim[:, 0], c.lim[:, 1], size=(n, 6))
        if self.mode == "ik":
            T = O.fanuc_fk(q, c)
            T[::100, :3, 3] *= 3.0
            t0 = time.perf_counter()
            with np.errstate(all="ignore"):
                O.fanuc_ik_branches(T, O.CANONICAL_8, c)
        else:
            t0 = time.perf_counter()
            O.fanuc_fk(q, c), O.jac_det(O.fanuc_jacobian(q, c))
        dt = time.perf_counter() - t0
        self._cpu_n = n
        return n / dt, dt

    def cpu_sample_desc(self, args, dt):
        if getattr(self, "_cpu_n", None):
            what = "8-branch fanuc_ik (1 % unreachable)" if self.mode == "ik" else "fanuc_fk + fanuc_jacobian + det"
            return f"{self._cpu_n} poses, oracle {what} (NumPy restatement of the reference), 1 process, {dt:.1f} s"
        return (f"{args.cpu_sample} uniform-in-limits poses, oracle fanuc_fk + fanuc_jacobian + "
                f"det + 8-branch fanuc_ik (NumPy restatement of the reference), 1 process, "
                f"{dt:.1f} s")


class CalibWorkload:
    """BASELINE config 5: one calibration iteration = J^T J / J^T r reduction over the
    rank's shard of measurements (+ the 190-double all-reduce) + the 18x18 host solve."""
    unit = "measurements/s"
    metric = "calibration normal-equation measurements/s (Fanuc reduced model, fp64)"
    cpu = None

    def __init__(self, args, rank, world):
        import torch.distributed as dist
        from roboticmanipulators_b200.utils.model_calibration import ModelCalibrator
        from roboticmanipulators_b200.utils.robo_math import chain_batch
        self.args, self.world = args, world
        self.m = m = args.batch if args.batch != (1 << 26) else (100_000_000 // world)
        self.units_per_step = m
        mc = ModelCalibrator(
            "TzRzTxRyTxRyTxTzRxRyRxTx", [0, 1, 0, 1, 0, 1, 0, 0, 1, 1, 1, 0],
            ['d_1', 'q_1', 'd_2', 'q_2', 'd_3', 'q_3', 'd_4', 'd_5', 'q_4', 'q_5', 'q_6', 'd_6'],
            [('d_1', 670), ('d_2', 312), ('d_3', 1075), ('d_5', 225), ('d_4', 1280), ("d_6", 215)],
            offsets=np.array([0.0, -np.pi / 2, np.pi / 2, 0.0, 0.0, 0.0]),
            directions=np.array([1, 1, -1, -1, -1, -1]))
        mc.verbose = False
        mc.get_reduced_model()
        self.mc = mc
        # SURVEY G7 values: the calibrated base/tools and the parameter step of the real data set
        self.true_p = np.array([1.30610510, -3.29297131, 5.31380935e-4, -2.10034074e-4, 0.545044972,
                                2.54192873e-4, 6.39792394e-5, -4.30615152e-3, -0.272805100,
                                -5.07970965, 1.94468834e-3, -2.39723149e-3, -0.358522920,
                                0.197375190, 4.43141007e-4, 4.43439344e-4, -0.130112096,
                                1.01596659e-3])
        est = [-561.306307, -560.988903, 557.860232, -3.06694464e-3, 8.27619470e-3, 5.71090871e-3,
               417.073077, -17.1212195, 59.7437387, 421.054034, 56.9970287, -15.7799788,
               420.743893, -43.8505917, -42.4028750]
        Tb, Tt = mc._compose_base_tools(est)
        self.T_base = np.array(Tb, dtype=np.float64)
        self.T_tools = [np.array(t, dtype=np.float64) for t in Tt]
        dev = torch.device("cuda")
        g = torch.Generator(device=dev).manual_seed(4 + 1000 * rank)
        lim = torch.tensor(np.deg2rad(np.array(Fanuc165F.qs_lim_deg)), device=dev)
        self.q = torch.empty((m, 6), dtype=torch.float64, device=dev)
        self.pts = torch.empty((m, 9), dtype=torch.float64, device=dev)
        off = torch.tensor(mc.qs_offsets, device=dev)
        dirs = torch.tensor(np.asarray(mc.qs_directions, dtype=np.float64), device=dev)
        pvals = torch.tensor(mc._nominal_values(self.true_p), dtype=torch.float64, device=dev)
        chunk = 1 << 21
        jpos = [i for i, j in enumerate(mc._r_j_ind) if j]
        ppos = [i for i, j in enumerate(mc._r_j_ind) if not j]
        for s in range(0, m, chunk):
            e = min(m, s + chunk)
            u = torch.rand((e - s, 6), generator=g, device=dev, dtype=torch.float64)
            qc = lim[:, 0] + u * (lim[:, 1] - lim[:, 0])
            self.q[s:e] = qc
            vals = torch.empty((e - s, len(mc._r_tokens)), dtype=torch.float64, device=dev)
            vals[:, jpos] = qc * dirs + off
            vals[:, ppos] = pvals
            for t, Tt_ in enumerate(self.T_tools):
                T = chain_batch(mc._r_tokens, vals, T_base=self.T_base, T_tool=Tt_, layout="aos")
                noise = torch.randn((e - s, 3), generator=g, device=dev, dtype=torch.float64) * 0.0439
                self.pts[s:e, 3 * t:3 * t + 3] = T[:, :3, 3] + noise
        torch.cuda.synchronize()
        self.events, self.params, self.host_ms, self.iter_ms = [], None, [], []
        self.alg_bytes = {"calib_normal_eq": 120}
        # FP64-bound kernel: algorithmic work per measurement: 6 compressed
        # rows x 210 upper-triangular FMAs of [J|r]^T [J|r] + 1240 FMAs for the sincos, the chain walk,
        # the 18 parameter columns and the residuals = 2500 FMAs = 5000 flop
        self.alg_flop = {"calib_normal_eq": 5000.0}
        self.traffic, self.traffic_source = traffic_for(m)
        self.launches_per_step = 1
        self.config = {
            "workload": "fanuc_calibration_normal_equations",
            "baseline_config": "configs[4]: J^T J / J^T r over synthetic measurements "
                               "(calibration_dataset.mat shape), all-reduce of 190 doubles, 18x18 solve",
            "measurements_per_gpu": m, "tools": 3, "params": 18,
            "l2": "inputs larger than L2 (12 GB of q/pts per pass at 1e8 measurements)",
            "sharding": "measurements sharded across ranks; one NCCL all-reduce of the packed "
                        "normal equations per iteration",
        }

    def step(self, s):
        """One calibration iteration: the normal-equation kernel over this rank's shard, the NCCL
        all-reduce of the packed [H | g | r^T r] vector (190 doubles), the download and the 18x18
        solve with the reference's np.linalg.inv(H).dot(g) (model_calibration.py:427-430)."""
        timed = s >= 0
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timed else None
        if timed:
            ev[0].record()
        tl0 = time.perf_counter()
        out = self.mc.normal_equations(self.q, self.pts, self.T_base, self.T_tools,
                                       parameters=np.zeros(18), process_group=False)
        if timed:
            ev[1].record()
        t0 = time.perf_counter()
        self.launch_ms = getattr(self, "launch_ms", []) + [1e3 * (t0 - tl0)]
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(out)
        if timed:
            ev[2].record()
            self.events.append(ev)
        host = out.cpu().numpy()
        t1 = time.perf_counter()
        H, g, _ = self.mc.unpack_normal_equations(host, 18)
        self.params = np.linalg.inv(H).dot(g)
        if timed:
            self.host_ms.append((1e3 * (t1 - t0), 1e3 * (time.perf_counter() - t1)))
            self.iter_ms.append(1e3 * (time.perf_counter() - tl0))

    def kernel_times_ms(self):
        hm = np.mean(np.array(self.host_ms), axis=0)
        self.config["step_breakdown_ms"] = {
            "host_compile_and_launch": float(np.mean(self.launch_ms[-len(self.host_ms):])),
            "wait_kernel_allreduce_d2h": float(hm[0]), "unpack_and_18x18_solve": float(hm[1])}
        return {"calib_normal_eq": float(np.mean([e[0].elapsed_time(e[1]) for e in self.events]))}

    def extra_record(self, dist):
        """What north_star asks of config 5: the iteration over 1e8 measurements (all ranks) in
        wall-clock ms, with the all-reduce separated (CUDA events around the NCCL call on the
        kernel's stream; it includes the wait for the slowest rank's kernel)."""
        ar = float(np.mean([e[1].elapsed_time(e[2]) for e in self.events]))
        it = float(np.mean(self.iter_ms))
        k = float(np.mean([e[0].elapsed_time(e[1]) for e in self.events]))
        if dist is not None:
            t = torch.tensor([it, ar, k], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            it, ar, k = (float(x) for x in t.tolist())
        # a bare all-reduce of the same 190 doubles, ranks already in step (pure collective latency)
        bare = None
        if dist is not None:
            buf = torch.zeros(190, dtype=torch.float64, device="cuda")
            for _ in range(5):
                dist.all_reduce(buf)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                dist.all_reduce(buf)
            e1.record()
            torch.cuda.synchronize()
            bare = e0.elapsed_time(e1) / 50
        return {"iteration_ms_wall": it, "kernel_ms": k, "allreduce_ms_incl_rank_skew": ar,
                "allreduce_190_doubles_ms_bare": bare,
                "measurements_all_ranks": self.m * (dist.get_world_size() if dist is not None else 1),
                "target": "north_star: a calibration iteration on 1e8 measurements in < 100 ms",
                "target_met": it < 100.0,
                "solution_check_max_abs_error_vs_true_parameters": float(np.max(np.abs(self.params - self.true_p)))}

    def e2e(self, args, dist=None):
        """ModelCalibrator.estimate_parameters with HOST (pinned) q/pts: upload + reduce + solve."""
        n = min(self.m, 1 << 22)
        hq = self.q[:n].cpu().pin_memory()
        hp = self.pts[:n].cpu().pin_memory()
        steps = 3

        def one():
            dq = hq.to("cuda", non_blocking=True)
            dp = hp.to("cuda", non_blocking=True)
            self.mc._params_iteration = 0
            return self.mc.estimate_parameters(dq, dp, self.T_base, self.T_tools,
                                               process_group=False)
        one()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        dt = time.perf_counter() - t0
        return finish_e2e({"seconds": dt, "units": n * steps, "h2d": n * 120, "d2h": 190 * 8, "batch": n,
                           "steps": steps, "api": "ModelCalibrator.estimate_parameters on pinned host q / pts"},
                          dist, self.unit)

    def cpu(self, args):
        from oracle import np_oracle as O
        mc = self.mc
        lens = [('d_1', 670), ('d_2', 312), ('d_3', 1075), ('d_5', 225), ('d_4', 1280), ("d_6", 215)]
        model = O.CalibModel(mc._r_seq, mc._r_var, mc._r_j_ind, mc._r_nom_par, lens,
                             mc.qs_offsets, mc.qs_directions)
        n = 20000
        q, p = self.q[:n].cpu().numpy(), self.pts[:n].cpu().numpy()
        t0 = time.perf_counter()
        O.calib_estimate_parameters(model, q, p, np.zeros(18), self.T_base, self.T_tools)
        dt = time.perf_counter() - t0
        self._cpu_n = n
        return n / dt, dt

    def cpu_sample_desc(self, args, dt):
        return (f"{self._cpu_n} measurements x 3 tools, oracle calib_estimate_parameters "
                f"(NumPy restatement of model_calibration.py:381-430), 1 process, {dt:.1f} s")


class TrajWorkload:
    """BASELINE config 4: equalised trapezoidal joint trajectories fused with the RRR forward
    kinematics and 6x3 Jacobian of every sample."""
    unit = "samples/s"
    metric = "RRR trajectory samples/s (p2p trapezoid + FK + Jacobian per sample, fp64)"

    def __init__(self, args, rank, world):
        from roboticmanipulators_b200.robots import RRRRobot
        self.args = args
        self.ntraj = nt = args.batch if args.batch != (1 << 26) else (1 << 17)
        self.n = 1000
        self.units_per_step = nt * self.n
        self.robot = RRRRobot()
        self.st = self.robot._struct()
        self.lib = _lib.load()
        dev = torch.device("cuda")
        g = torch.Generator(device=dev).manual_seed(3 + 1000 * rank)
        self.q0 = (torch.rand((nt, 3), generator=g, device=dev, dtype=torch.float64) * 2 - 1) * np.pi
        self.qf = (torch.rand((nt, 3), generator=g, device=dev, dtype=torch.float64) * 2 - 1) * np.pi
        f64 = torch.float64
        self.q, self.dq, self.ddq = (torch.empty((nt, self.n, 3), dtype=f64, device=dev) for _ in range(3))
        self.t = torch.empty((nt, self.n), dtype=f64, device=dev)
        self.len = torch.empty(nt, dtype=torch.int32, device=dev)
        self.T = torch.empty((nt, self.n, 16), dtype=f64, device=dev)
        self.J = torch.empty((nt, self.n, 18), dtype=f64, device=dev)
        self.events = []
        self.alg_bytes = {"rrr_p2p_fk": 352}
        self.traffic, self.traffic_source = traffic_for(nt * self.n)
        self.launches_per_step = 1
        self.config = {
            "workload": "rrr_p2p_trapezoid_fk_jac",
            "baseline_config": f"configs[3] chunked: {nt} trajectories x 1000 steps per GPU per step "
                               "(dq_max=1, ddq_max=20, control_freq=10)",
            "trajectories_per_gpu_per_step": nt, "samples_per_trajectory": self.n,
            "l2": "outputs larger than L2 (46 GB written per step)",
            "sharding": "independent trajectories, no collective",
        }

    def step(self, s):
        timed = s >= 0
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if timed else None
        if timed:
            ev[0].record()
        _lib.check(self.lib.rk_rrr_p2p_fk_f64(
            C.byref(self.st), _lib.ptr(self.q0), _lib.ptr(self.qf), self.ntraj, self.n, 1.0, 20.0,
            10.0, _lib.ptr(self.q), _lib.ptr(self.dq), _lib.ptr(self.ddq), _lib.ptr(self.t),
            _lib.ptr(self.len), _lib.ptr(self.T), _lib.ptr(self.J), _lib.stream_ptr()))
        if timed:
            ev[1].record()
            self.events.append(ev)

    def kernel_times_ms(self):
        return {"rrr_p2p_fk": float(np.mean([e[0].elapsed_time(e[1]) for e in self.events]))}

    def e2e(self, args, dist=None):
        from roboticmanipulators_b200.utils.trajectory_generator import TrajectoryGenerator
        tg = TrajectoryGenerator(1, 20, 1, 10, control_freq=10)
        nt = min(self.ntraj, 2048)
        h0, hf = self.q0[:nt].cpu().pin_memory(), self.qf[:nt].cpu().pin_memory()
        hout = [torch.empty(shape, dtype=torch.float64).pin_memory() for shape in
                [(nt, self.n, 3)] * 3 + [(nt, self.n), (nt, self.n, 4, 4), (nt, self.n, 6, 3)]]
        hlen = torch.empty(nt, dtype=torch.int32).pin_memory()
        steps = 3

        def one():
            a, b = h0.to("cuda", non_blocking=True), hf.to("cuda", non_blocking=True)
            q, dq, ddq, t, L, T, J = tg.generate_p2p_trajectory_batch(a, b, self.n, robot=self.robot)
            for h, d in zip(hout, (q, dq, ddq, t, T, J)):
                h.copy_(d, non_blocking=True)
            hlen.copy_(L, non_blocking=True)
            torch.cuda.synchronize()
        one()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        dt = time.perf_counter() - t0
        return finish_e2e({"seconds": dt, "units": nt * self.n * steps, "h2d": nt * 48,
                           "d2h": nt * self.n * 352 + nt * 4, "batch": nt * self.n, "steps": steps,
                           "api": "TrajectoryGenerator.generate_p2p_trajectory_batch, results copied to pinned host"},
                          dist, self.unit)

    def cpu(self, args):
        from oracle import np_oracle as O
        nt = 300
        q0, qf = self.q0[:nt].cpu().numpy(), self.qf[:nt].cpu().numpy()
        t0 = time.perf_counter()
        q, dq, ddq, t, L = O.p2p_batch(q0, qf, self.n, 1, 20, 10)
        O.rrr_fk(q)
        O.rrr_jacobian(q)
        dt = time.perf_counter() - t0
        self._cpu_n = nt
        return nt * self.n / dt, dt

    def cpu_sample_desc(self, args, dt):
        return (f"{self._cpu_n} trajectories x 1000 samples: oracle p2p_batch + rrr_fk + rrr_jacobian "
                f"(NumPy restatement), 1 process, {dt:.1f} s")


def make(args, rank, world):
    if args.workload in ("fk_jac_ik", "fk_jac_ik_split", "fk_jac", "ik"):
        return FanucWorkload(args, rank, world)
    if args.workload == "calib":
        return CalibWorkload(args, rank, world)
    if args.workload == "traj":
        return TrajWorkload(args, rank, world)
    raise SystemExit(f"unknown workload {args.workload}")
