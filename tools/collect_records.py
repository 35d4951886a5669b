#!/usr/bin/env python
"""Copy the records of tools/run_round_records.sh from gpurun_out/ into profiles/ and derive the two
digests that need no GPU:

  profiles/<R>_fanuc_launches.csv / _launches_summary.json   per-kernel launch counts, mean time and
        share of all GPU time of the `ncu --metrics gpu__time_duration.sum` pass over the default
        bench command (cold-cache, serialised: shares, not absolutes)
  profiles/<R>_sass_summary.txt   cuobjdump -sass of librk_b200.so: per kernel the instruction count,
        the mnemonics that show the data path, registers / stack / shared memory

  python tools/collect_records.py r02
"""
import collections
import csv
import glob
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ["UTMALDG", "UTMASTG", "DMMA", "DFMA", "DMUL", "DADD", "MUFU.RCP64H", "MUFU.RSQ64H", "STG", "LDG",
             "STS", "LDS", "LDL", "STL", "FFMA", "SHFL"]


def launches(R):
    src = os.path.join(ROOT, "gpurun_out", f"{R}_launches.csv")
    if not os.path.exists(src):
        return
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    kn, gs, mv, mu = (hdr.index(c) for c in ("Kernel Name", "Grid Size", "Metric Value", "Metric Unit"))
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"^void ", "", r[kn].split("(")[0])
        us = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu], 1e-3)
        a = agg.setdefault((name, r[gs]), [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    out = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fanuc|calib|traj' "
                      "python bench.py --steps 2 --warmup 3 --batch 8388608 --no-cpu",
           "note": "per-launch times under ncu are cold-cache and serialised: shares, not absolutes. The fused "
                   "kernel at the largest grid is the headline step (8.39 M poses per launch); the 65536-pose "
                   "launches are the chunks of the host-streaming e2e call; the rest belong to the extra workloads",
           "kernels": [{"kernel": k, "grid": g, "launches": n, "mean_us": round(t / n, 2), "total_us": round(t, 1),
                        "share_of_all_gpu_time": round(t / total, 4)} for (k, g), (n, t) in agg.items()]}
    json.dump(out, open(os.path.join(ROOT, "profiles", f"{R}_fanuc_launches_summary.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "profiles", f"{R}_fanuc_launches.csv"), "w") as f:   # compact: one line per kernel/grid
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "launches", "mean_us", "total_us", "share"])
        for k in out["kernels"]:
            w.writerow([k["kernel"], k["grid"], k["launches"], k["mean_us"], k["total_us"], k["share_of_all_gpu_time"]])


def sass(R):
    so = os.path.join(ROOT, "roboticmanipulators_b200", "librk_b200.so")
    res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
    usage, cur = {}, None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = " ".join(re.findall(r"(?:REG|STACK|SHARED):\d+", line))
            cur = None
    text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    counts, order, cur = {}, [], None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["instr"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    counts[cur][mn] += 1
    names = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
    with open(os.path.join(ROOT, "profiles", f"{R}_sass_summary.txt"), "w") as f:
        f.write("# SASS summary of roboticmanipulators_b200/librk_b200.so (cuobjdump -sass / -res-usage, sm_100a cubins;\n"
                "# tools/collect_records.py).  Per kernel: instructions, the mnemonics that show the data path\n"
                "# (UTMALDG/UTMASTG = TMA tensor loads/stores, DMMA = FP64 tensor-core MMA, DFMA/DMUL/DADD = FP64 pipe,\n"
                "# LDL/STL = local memory -- what is left sits in the libm slow paths and loop-control spills), registers,\n"
                "# stack, static shared memory.  No UTC*MMA / LDTM: tcgen05 has no FP64 and these are per-pose chains.\n"
                "# The generic-chain kernels are generated at run time (NVRTC) and are not in this file: see\n"
                "# rk_chain_jit_source() and tests/test_host_logic_cpu.py for their source.\n\n")
        for mangled, name in zip(order, names):
            c = counts[mangled]
            short = re.sub(r"\(.*", "", name)
            parts = " ".join(f"{mn}={c[mn]}" for mn in MNEMONICS if c[mn])
            f.write(f"{short:92s} instr={c['instr']:5d} {parts}  [{usage.get(mangled, '')}]\n")


def main():
    R = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = os.path.join(ROOT, "gpurun_out")
    for path in sorted(glob.glob(os.path.join(out, f"{R}_bench_*.json")) + glob.glob(os.path.join(out, f"{R}_rows.json")) +
                       glob.glob(os.path.join(out, f"{R}_*_ncu_full_summary.json"))):
        if os.path.getsize(path) > 0:
            shutil.copy(path, os.path.join(ROOT, "profiles", os.path.basename(path)))
    launches(R)
    sass(R)


if __name__ == "__main__":
    main()
