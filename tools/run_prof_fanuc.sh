set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --dtype f32 --no-cpu > gpurun_out/bench_f32.json 2>gpurun_out/bench_f32.err
python -c "
import json
d=json.load(open('gpurun_out/bench_f32.json')); print('f32', d['value'], {k:(v['ms_per_launch'],v['hbm_frac']) for k,v in d['kernels'].items()})"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fanuc --csv --log-file gpurun_out/launches_fanuc.csv python bench.py --steps 2 --warmup 3 --batch 8388608 --no-cpu > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fanuc -s 6 -c 2 -o gpurun_out/prof_fanuc -f python bench.py --steps 2 --warmup 3 --batch 8388608 --no-cpu > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
