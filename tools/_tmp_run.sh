python -m pytest tests/test_r2_gpu.py tests/test_arms_chain_gpu.py tests/test_traj_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/bench_rows.py > gpurun_out/rows_new.json 2> gpurun_out/rows_new.err; tail -3 gpurun_out/rows_new.err
