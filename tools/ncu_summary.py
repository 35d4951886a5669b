#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small JSON record for profiles/.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [units_per_launch | rows.json[:scale]]

Per kernel: duration, DRAM bytes / GB/s / % of ncu's peak, FP64-pipe utilisation (cycles active --
counts the 16-cycle DMMA properly -- and instructions), issue-slot utilisation, LSU data-pipe
utilisation, occupancy, registers, instructions per unit, the top stall reasons, and `binding`: the
busiest of those units, i.e. the roofline the kernel actually sits under ("report both
for every kernel, name the binding one").  With a rows.json (tools/bench_rows.py output) every row
is matched by kernel name (ROW_KERNELS) for its units and algorithmic bytes; launches that belong to
no row (set-up work of the script) are dropped."""
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
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_inst_pct",
    "TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "fp64_pipe_cycles_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_data_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
}
# bench_rows.py row -> (substring of the kernel name, which occurrence of that kernel)
ROW_KERNELS = [
    ("rrr_fk_jac_soa", "arm_fk_jac_kernel<double, 3, 1>", 0),
    ("rrr_fk_jac_aos", "arm_fk_jac_kernel<double, 3, 0>", 0),
    ("rrr_ik4_soa", "arm_ik_pose_kernel<double, 3, 1>", 0),
    ("rr_fk_jac_soa", "arm_fk_jac_kernel<double, 2, 1>", 0),
    ("rr_ik2_soa", "arm_ik_pose_kernel<double, 2, 1>", 0),
    ("chain_fk_jac_fanuc_seq_soa", "rk_chain_jit", 0),
    ("chain_fk_fanuc_seq_soa", "rk_chain_jit", 1),
    ("chain_fk_jac_fanuc_seq_soa_f32", "rk_chain_jit", 2),
    ("chain_fk_fanuc_seq_soa_f32", "rk_chain_jit", 3),
    ("rrr_fk_jac_soa_f32", "arm_fk_jac_soa2_kernel<3>", 0),
    ("rrr_ik4_soa_f32", "arm_ik_pose_kernel<float, 3, 1>", 0),
    ("rr_fk_jac_soa_f32", "arm_fk_jac_soa2_kernel<2>", 0),
    ("rr_ik2_soa_f32", "arm_ik_pose_kernel<float, 2, 1>", 0),
    ("chain_dq_from_dx_6x6", "rk_chain_jit", 4),
    ("fanuc_dq_from_dx", "fanuc_dq_from_dx_kernel<double", 0),
    ("rrr_p2p_fk_f32", "traj_kernel<3, 1, float", 0),
    ("trap_p2p_6joints", "traj_kernel<6, 0", 0),
    ("rrr_lin_ik_dq", "traj_kernel<3, 2", 0),
    ("trap_coefficients", "trap_coefficients_kernel", 0),
    ("trap_sample", "trap_sample_kernel", 0),
    ("poly5_profiles", "poly5_kernel", 0),
    ("calib_tool_base_eq", "calib_tool_base_stats_kernel", 0),
    ("calib_evaluate", "calib_eval_kernel", 0),
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9,
         "s": 1.0}


def main():
    rep = sys.argv[1]
    units, rows = None, None
    if len(sys.argv) > 2:
        arg = sys.argv[2]
        if arg.split(":")[0].endswith(".json"):
            path, _, scale = arg.partition(":")
            scale = float(scale) if scale else 1.0
            rows = {k: (v["units"] * scale, v["algorithmic_bytes_per_unit"])
                    for k, v in json.load(open(path)).items()}
        else:
            units = float(arg)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    table = list(csv.reader(io.StringIO(raw)))
    hdr, unit_row = table[0], table[1]
    out = []
    seen = {}
    for n, r in enumerate(table[2:]):
        rec = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")}
        occ = seen.get(rec["kernel"], 0)
        seen[rec["kernel"]] = occ + 1
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
        u, alg = units, None
        if rows is not None:
            hit = [row for row, pat, k in ROW_KERNELS if pat in rec["kernel"] and k == occ and row in rows]
            if not hit:
                continue
            rec["row"] = hit[0]
            u, alg = rows[hit[0]]
        if "dram_read" in rec and "dram_write" in rec:
            rec["dram_traffic"] = rec["dram_read"] + rec["dram_write"]
            rec["dram_gbs"] = rec["dram_traffic"] / rec["time"] / 1e9
            if u:
                rec["units"] = u
                rec["traffic_bytes_per_unit"] = rec["dram_traffic"] / u
                rec["warp_instructions_per_unit"] = rec.get("warp_instructions", 0) / u
                rec["thread_instructions_per_unit"] = 32 * rec.get("warp_instructions", 0) / u
            if alg:
                rec["algorithmic_bytes_per_unit"] = alg
        cand = {"hbm": rec.get("dram_pct_of_ncu_peak", 0.0),
                "fp64_pipe": max(rec.get("fp64_pipe_cycles_pct", 0.0), rec.get("fp64_pipe_inst_pct", 0.0)),
                "issue_slots": rec.get("issue_active_pct", 0.0),
                "lsu_data_pipe": rec.get("lsu_data_pipe_pct", 0.0)}
        rec["binding"] = max(cand, key=cand.get)
        rec["binding_pct"] = round(cand[rec["binding"]], 1)
        out.append(rec)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
