# Partial refresh of the round records (run on a GPU box): the GPU test-suite, tools/bench_rows.py and
# its ncu pass -- the part of tools/run_round_records.sh that the non-headline kernels need.
#   gpurun --timeout 1500 -- 'bash tools/run_rows_records.sh'   then   python tools/collect_records.py r02
R=r02
python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest.log 2>&1; tail -n 3 gpurun_out/${R}_pytest.log
python tools/bench_rows.py > gpurun_out/${R}_rows.json 2> gpurun_out/${R}_rows.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'arm_|rk_chain_jit|chain_kernel|traj_kernel|trap_|poly5|calib_|fanuc_dq' -o gpurun_out/${R}_rows -f python tools/bench_rows.py --warmup 0 --reps 1 --scale 0.25 > gpurun_out/${R}_ncu_rows.log 2>&1
tail -n 1 gpurun_out/${R}_ncu_rows.log
python tools/ncu_summary.py gpurun_out/${R}_rows.ncu-rep gpurun_out/${R}_rows.json:0.25 > gpurun_out/${R}_rows_ncu_full_summary.json
rm -f gpurun_out/${R}_rows.ncu-rep
