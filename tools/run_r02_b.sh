set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; tail -n 30 gpurun_out/r02b_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; tail -n 3 gpurun_out/r02b_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -n 5 gpurun_out/r02b_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err; tail -n 3 gpurun_out/r02b_bench_ref.err
