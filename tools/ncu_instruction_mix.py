#!/usr/bin/env python
"""Dynamic instruction mix of one kernel from an .ncu-rep captured with --import-source on
(`ncu -i rep --page source --csv`, read here -- no GPU): executed warp-instructions per unit by
opcode, their share of the stall samples, and the split by how often a static instruction runs per
unit (loop nesting).

  python tools/ncu_instruction_mix.py gpurun_out/fused_src.ncu-rep 8388608 > profiles/r02_fused_instruction_mix.json
"""
import collections
import csv
import io
import json
import subprocess
import sys


def opcode(src):
    parts = src.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    if op.startswith("IMAD"):
        return "IMAD.MOV" if "MOV" in op else "IMAD"
    return op.split(".")[0]


def main():
    rep, units = sys.argv[1], float(sys.argv[2])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    kernel, hdr, data = rows[0][1], rows[1], rows[2:]
    i_src, i_exec, i_samp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    per_warp = units / 32.0
    by_op, samples, by_mult = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    total = 0
    for r in data:
        try:
            e, s = int(r[i_exec]), int(r[i_samp])
        except (ValueError, IndexError):
            continue
        op = opcode(r[i_src].strip())
        by_op[op] += e
        samples[op] += s
        total += e
        by_mult[round(e / per_warp, 2)][op] += 1
    nsamp = sum(samples.values())
    out = {"kernel": kernel, "units": units, "warp_instructions_per_unit_warp": total / per_warp,
           "by_opcode": {op: {"per_unit": round(c / per_warp, 1), "share": round(c / total, 4),
                              "stall_sample_share": round(samples[op] / nsamp, 4)}
                         for op, c in by_op.most_common(30)},
           "by_executions_per_unit": {str(m): {"static_instructions": sum(c.values()),
                                               "per_unit": round(m * sum(c.values()), 1),
                                               "top": dict(c.most_common(10))}
                                      for m, c in sorted(by_mult.items(), reverse=True) if m * sum(c.values()) >= 5}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
