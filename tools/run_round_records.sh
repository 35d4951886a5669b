# Regenerates the bench records and ncu captures kept under profiles/ (run on a GPU box):
#   gpurun --timeout 1500 -- 'bash tools/run_round_records.sh r02'
# Every ncu capture follows a plain run of the same command that exited 0; numbers printed under ncu
# are never bench values.
R=${1:-r02}
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest.log 2>&1; tail -n 3 gpurun_out/${R}_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err || tail -n 5 gpurun_out/${R}_bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2> gpurun_out/${R}_bench_ref.err
for v in "split:--workload fk_jac_ik_split" "f32:--dtype f32" "aos:--layout aos" "ik_aos:--workload ik --layout aos" "ik_f32:--workload ik --dtype f32" "ik_f32_aos:--workload ik --dtype f32 --layout aos" "fk_jac:--workload fk_jac"; do
  name=${v%%:*}; flags=${v#*:}
  python bench.py $flags --no-cpu --no-extras --steps 10 > gpurun_out/${R}_bench_$name.json 2> gpurun_out/${R}_bench_$name.err || tail -n 3 gpurun_out/${R}_bench_$name.err
done
python tools/bench_rows.py > gpurun_out/${R}_rows.json 2> gpurun_out/${R}_rows.err
# launch list of the default command (kernel share of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fanuc|calib|traj' --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 3 --batch 8388608 --no-cpu > gpurun_out/${R}_ncu_launches.log 2>&1
# one full capture per kernel
# (gpurun copies back at most 64 MiB: every report is summarised on the box and only the two
# headline reports are kept)
cap() { # name, kernel regex, skip, bench flags, units per launch, keep
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/${R}_$1 -f python bench.py $4 --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/${R}_ncu_$1.log 2>&1
  tail -n 1 gpurun_out/${R}_ncu_$1.log
  python tools/ncu_summary.py gpurun_out/${R}_$1.ncu-rep $5 > gpurun_out/${R}_$1_ncu_full_summary.json
  [ "$6" = keep ] || rm -f gpurun_out/${R}_$1.ncu-rep
}
cap fused fanuc_fk_jac_ik8 3 "--batch 8388608" 8388608 keep
cap fk_jac fanuc_fk_jac_tma 3 "--workload fk_jac --batch 8388608" 8388608
cap ik fanuc_ik8_pose 3 "--workload ik --batch 8388608" 8388608
cap ik_aos fanuc_ik8_pose_aos 3 "--workload ik --layout aos --batch 8388608" 8388608
cap f32fused fanuc_fk_jac_ik8 3 "--dtype f32 --batch 8388608" 8388608
cap ik_f32 fanuc_ik8_pose 3 "--workload ik --dtype f32 --batch 8388608" 8388608
cap calib calib_normal 3 "--workload calib --batch 8388608" 8388608 keep
cap traj traj_kernel 3 "--workload traj --batch 32768" 32768000
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'arm_|rk_chain_jit|chain_kernel|traj_kernel|trap_|poly5|calib_|fanuc_dq' -o gpurun_out/${R}_rows -f python tools/bench_rows.py --warmup 0 --reps 1 --scale 0.25 > gpurun_out/${R}_ncu_rows.log 2>&1
tail -n 1 gpurun_out/${R}_ncu_rows.log
python tools/ncu_summary.py gpurun_out/${R}_rows.ncu-rep gpurun_out/${R}_rows.json:0.25 > gpurun_out/${R}_rows_ncu_full_summary.json
rm -f gpurun_out/${R}_rows.ncu-rep
# what ncu reports for the FP64 pipe at its true peak (the dependent-free DFMA probe): the scale of
# the fp64 percentages above
timeout 120 ncu --set full --clock-control none -k regex:dfma_probe -c 1 -o gpurun_out/${R}_dfma -f python -c "
import sys; sys.path.insert(0, '.')
import bench_workloads as W; print(W.fp64_probe())" > gpurun_out/${R}_ncu_dfma.log 2>&1
python tools/ncu_summary.py gpurun_out/${R}_dfma.ncu-rep > gpurun_out/${R}_dfma_ncu_full_summary.json; rm -f gpurun_out/${R}_dfma.ncu-rep
du -sh gpurun_out
# python tools/collect_records.py $R   # (here, after the call: copies the records into profiles/, derives the launch and SASS digests)
