#!/usr/bin/env python
"""Time the generic-chain rows (Fanuc sequence, 2^24 samples, SOA planes) under the current
RK_CHAIN_VEC2 / RK_CHAIN_JIT environment: one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from roboticmanipulators_b200.utils.jacobians import JacobianCalculator
    from roboticmanipulators_b200.utils.robo_math import Transformation
    dev = torch.device("cuda")
    n = 1 << 24
    seq = "TzRzTzTxRyTxRyTzTxRxRyRxTx"
    jind = [1 if seq[2 * k] == "R" else 0 for k in range(13)]
    jc = JacobianCalculator(seq, jind)
    out = {"RK_CHAIN_VEC2": os.environ.get("RK_CHAIN_VEC2", "default")}
    for name, dt, w in (("f64", torch.float64, 8), ("f32", torch.float32, 4)):
        vals = (6 * torch.rand((13, n), device=dev, dtype=dt) - 3).t()
        for row, nb, fn in (("T+J", (13 + 16 + 36) * w, lambda: jc.calculate_numerically_batch(vals, layout="soa")),
                            ("T", (13 + 16) * w, lambda: Transformation.batch(seq, vals, layout="soa"))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            out[f"{name} {row}"] = {"ms": round(ms, 4), "gbs": round(n * nb / ms / 1e6, 1)}
        del vals
    print(json.dumps(out))


if __name__ == "__main__":
    main()
