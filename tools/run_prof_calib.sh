set -x
python bench.py --workload calib > gpurun_out/bench_calib.json 2> gpurun_out/bench_calib.err
python -c "
import json
d=json.load(open('gpurun_out/bench_calib.json')); print('calib', d['value'], d['ms_per_step'], d['roofline'], d.get('cpu_baseline'))"
ncu --set full --clock-control none --import-source on -k regex:calib_normal -s 3 -c 1 -o gpurun_out/prof_calib -f python bench.py --steps 2 --warmup 3 --batch 8388608 --no-cpu --workload calib > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
