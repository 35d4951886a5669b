"""Workloads timed by bench.py.  Each drives the C ABI (include/rk_b200.h) directly on
preallocated device buffers for the device-resident number, and the public Python API
with pinned host buffers for the end-to-end number."""
import ctypes as C
import time

import numpy as np
import torch

from roboticmanipulators_b200 import _lib
from roboticmanipulators_b200.robots.fanuc import Fanuc165F


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


class FanucWorkload:
    """K1 fk+jac+det then K2 ik8 over `batch` poses per step, SOA or AOS planes."""
    unit = "poses/s"
    metric = "FK+Jacobian+IK poses/s (Fanuc165F, fp64)"
    cpu = None

    def __init__(self, args, rank, world):
        self.args = args
        self.mode = args.workload
        self.n = n = args.batch
        self.units_per_step = n
        self.robot = Fanuc165F()
        self.st = self.robot._struct()
        self.lib = _lib.load()
        self.soa = args.layout == "soa"
        self.flags = _lib.RK_SOA if self.soa else _lib.RK_AOS
        dev = torch.device("cuda")
        # distinct input batches (so 15 steps really are 1e9 different poses), capped at 48 GB
        nbuf = max(1, min(args.steps, int(48e9 // (48 * n))))
        lim = torch.tensor(np.asarray(self.robot.qs_lim_rad), device=dev)
        self.q = []
        for b in range(nbuf):
            g = torch.Generator(device=dev).manual_seed(1 + 1000 * rank + b)
            u = torch.rand((6, n), generator=g, device=dev, dtype=torch.float64)
            u = lim[:, :1] + u * (lim[:, 1:] - lim[:, :1])
            self.q.append(u if self.soa else u.t().contiguous())
            del u
        f64 = torch.float64
        shape = (lambda c: (c, n)) if self.soa else (lambda c: (n, c))
        self.T = torch.empty(shape(16), dtype=f64, device=dev)
        self.J = torch.empty(shape(36), dtype=f64, device=dev) if self.mode != "ik" else None
        self.det = torch.empty(n, dtype=f64, device=dev) if self.mode != "ik" else None
        if self.mode != "fk_jac":
            self.qik = torch.empty(shape(48), dtype=f64, device=dev)
            self.valid = torch.empty(shape(8), dtype=torch.uint8, device=dev)
        if self.mode == "ik":      # targets = FK of the first batch, 1 % scaled out of reach
            self._k1(self.q[0], with_j=False)
            torch.cuda.synchronize()
        self.events = []
        self.alg_bytes = {"fanuc_fk_jac": 472, "fanuc_ik8": 520}
        self.traffic = {}
        self.launches_per_step = {"fk_jac_ik": 2, "fk_jac": 1, "ik": 1}[self.mode]
        self.config = {
            "workload": {"fk_jac_ik": "fanuc165f_fk_jac_det_then_ik8",
                         "fk_jac": "fanuc165f_fk_jac_det", "ik": "fanuc165f_ik8"}[self.mode],
            "baseline_config": "configs[1] (+configs[2] IK on the same poses), chunked: "
                               f"{args.steps} steps x {n} poses per GPU",
            "batch_per_gpu_per_step": n, "layout": args.layout,
            "distinct_input_batches": nbuf,
            "l2": "inputs larger than L2 (each step streams >= 3.2 GB of fresh q and "
                  "overwrites >= 28 GB of outputs)",
            "sharding": "independent poses, no collective",
        }

    def _k1(self, q, with_j=True):
        st = _lib.stream_ptr()
        _lib.check(self.lib.rk_fanuc_fk_jac_f64(
            C.byref(self.st), _lib.ptr(q), self.n, self.n, self.flags, _lib.ptr(self.T),
            _lib.ptr(self.J) if with_j else None, _lib.ptr(self.det) if with_j else None, st))

    def _k2(self):
        st = _lib.stream_ptr()
        _lib.check(self.lib.rk_fanuc_ik_f64(
            C.byref(self.st), _lib.ptr(self.T), self.n, self.n, self.flags, None, 8,
            _lib.ptr(self.qik), _lib.ptr(self.valid), st))

    def step(self, s):
        timed = s >= 0
        q = self.q[max(s, 0) % len(self.q)]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timed else None
        if timed:
            ev[0].record()
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
        if self.mode != "ik":
            out["fanuc_fk_jac"] = float(np.mean([e[0].elapsed_time(e[1]) for e in self.events]))
        if self.mode != "fk_jac":
            out["fanuc_ik8"] = float(np.mean([e[1].elapsed_time(e[2]) for e in self.events]))
        return out

    def e2e(self, args):
        """Public API, reference-shaped (pose-major) PINNED HOST buffers; every step copies
        its inputs host->device and all results device->host inside the timed region."""
        n = min(args.e2e_batch, self.n)
        steps = max(3, min(args.steps, 10))
        rng = np.random.default_rng(7)
        lim = np.asarray(self.robot.qs_lim_rad)
        hq = torch.from_numpy(rng.uniform(lim[:, 0], lim[:, 1], size=(n, 6))).pin_memory()
        f64 = torch.float64
        hT = torch.empty((n, 4, 4), dtype=f64).pin_memory()
        hJ = torch.empty((n, 6, 6), dtype=f64).pin_memory()
        hdet = torch.empty(n, dtype=f64).pin_memory()
        hqik = torch.empty((n, 8, 6), dtype=f64).pin_memory()
        hval = torch.empty((n, 8), dtype=torch.bool).pin_memory()
        dq = torch.empty((n, 6), dtype=f64, device="cuda")
        h2d = hq.numel() * 8
        d2h = (hT.numel() + hJ.numel() + hdet.numel() + hqik.numel()) * 8 + hval.numel()

        def one():
            dq.copy_(hq, non_blocking=True)
            T, J, det = self.robot.forward_kinematics_batch(dq, jacobian=True, measure=True,
                                                            layout="aos")
            qik, val = self.robot.inverse_kinematics_batch(T, layout="aos")
            hT.copy_(T, non_blocking=True)
            hJ.copy_(J, non_blocking=True)
            hdet.copy_(det, non_blocking=True)
            hqik.copy_(qik, non_blocking=True)
            hval.copy_(val, non_blocking=True)
            torch.cuda.synchronize()          # the user reads the results

        for _ in range(2):
            one()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        dt = time.perf_counter() - t0
        return {"seconds": dt, "units": n * steps, "h2d": h2d, "d2h": d2h, "batch": n,
                "steps": steps}

    def cpu_sample_desc(self, args, dt):
        return (f"{args.cpu_sample} uniform-in-limits poses, oracle fanuc_fk + fanuc_jacobian + "
                f"det + 8-branch fanuc_ik (NumPy restatement of the reference), 1 process, "
                f"{dt:.1f} s")


def make(args, rank, world):
    if args.workload in ("fk_jac_ik", "fk_jac", "ik"):
        return FanucWorkload(args, rank, world)
    raise SystemExit(f"workload {args.workload} is not available in this build")
