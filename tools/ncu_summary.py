#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small JSON/markdown record for
profiles/.  usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep units_per_launch"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_ncu_peak",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9,
         "s": 1.0}


def main():
    rep, units = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, unit_row = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        rec = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")}
        for k, name in KEYS.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    rec[name] = float(r[i]) * SCALE.get(unit_row[i], 1.0)
                except ValueError:
                    pass
        stalls = {}
        for i, h in enumerate(hdr):
            if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls[h.split("stalled_")[1].replace("_per_issue_active.ratio", "")] = round(float(r[i]), 3)
                except ValueError:
                    pass
        rec["top_stalls"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
        if "dram_read" in rec and "dram_write" in rec:
            rec["dram_traffic"] = rec["dram_read"] + rec["dram_write"]
            rec["dram_gbs"] = rec["dram_traffic"] / rec["time"] / 1e9
            if units:
                rec["traffic_bytes_per_unit"] = rec["dram_traffic"] / units
                rec["warp_instructions_per_unit"] = rec.get("warp_instructions", 0) / units
        out.append(rec)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
