# Regenerates the bench records and ncu summaries kept under profiles/ (run on a GPU box).
set -x
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --workload fk_jac_ik_split --no-cpu > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err
python bench.py --dtype f32 --no-cpu > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err
python bench.py --layout aos --no-cpu > gpurun_out/bench_aos.json 2> gpurun_out/bench_aos.err
python bench.py --workload traj > gpurun_out/bench_traj.json 2> gpurun_out/bench_traj.err
python bench.py --workload calib > gpurun_out/bench_calib.json 2> gpurun_out/bench_calib.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fanuc --csv --log-file gpurun_out/launches_fanuc.csv python bench.py --steps 2 --warmup 3 --batch 8388608 --no-cpu > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fanuc_fk_jac_ik8 -s 3 -c 1 -o gpurun_out/prof_fused -f python bench.py --steps 2 --warmup 3 --batch 8388608 --no-cpu > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fanuc -s 6 -c 2 -o gpurun_out/prof_fanuc -f python bench.py --workload fk_jac_ik_split --steps 2 --warmup 3 --batch 8388608 --no-cpu > gpurun_out/ncu2b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:calib_normal -s 3 -c 1 -o gpurun_out/prof_calib -f python bench.py --steps 2 --warmup 3 --batch 8388608 --no-cpu --workload calib > gpurun_out/ncu3.log 2>&1
python tools/bench_rows.py > gpurun_out/rows.json 2> gpurun_out/rows.err
tail -n 2 gpurun_out/ncu2.log; tail -n 2 gpurun_out/ncu3.log
# then, in the repo: python tools/ncu_summary.py gpurun_out/prof_fanuc.ncu-rep 8388608 > profiles/rNN_fanuc_ncu_full_summary.json
#                    python tools/ncu_summary.py gpurun_out/prof_fused.ncu-rep 8388608 > profiles/rNN_fused_ncu_full_summary.json
#                    python tools/ncu_summary.py gpurun_out/prof_calib.ncu-rep 8388608 > profiles/rNN_calib_ncu_full_summary.json
