set -x
python tools/diag_ik4.py none 2>&1 | tail -n 12
python tools/diag_ik3.py 200000 20 2>&1 | tail -n 6
python -m pytest tests -m gpu -q 2>&1 | tail -n 15
python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --workload ik > gpurun_out/r02c_ik.json 2>> gpurun_out/r02c_bench.err
python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --workload ik --layout aos > gpurun_out/r02c_ik_aos.json 2>> gpurun_out/r02c_bench.err
python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --workload ik --layout aos --dtype f32 > gpurun_out/r02c_ik_aos_f32.json 2>> gpurun_out/r02c_bench.err
python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --workload ik --dtype f32 > gpurun_out/r02c_ik_f32.json 2>> gpurun_out/r02c_bench.err
python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --dtype f32 > gpurun_out/r02c_f32.json 2>> gpurun_out/r02c_bench.err
for f in bench ik ik_aos ik_aos_f32 ik_f32 f32; do python -c "
import json,sys; d=json.load(open('gpurun_out/r02c_$f.json')); print('$f', d['ms_per_step'], d['roofline']['frac'])"; done
