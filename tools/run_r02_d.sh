set -x
python -m pytest tests/test_calibration_gpu.py tests/test_more_gpu.py -m gpu -q -x 2>&1 | tail -n 8
python bench.py --workload calib --steps 5 --warmup 3 --no-cpu > gpurun_out/r02d_calib.json 2> gpurun_out/r02d_calib.err
RK_CALIB_GENERIC=1 python bench.py --workload calib --steps 5 --warmup 3 --no-cpu > gpurun_out/r02d_calib_generic.json 2>> gpurun_out/r02d_calib.err
for f in calib calib_generic; do python -c "
import json,sys; d=json.load(open('gpurun_out/r02d_$f.json')); print('$f', d['ms_per_step'], d['roofline']['frac'], d['kernels'])"; done
ncu --set full --clock-control none --import-source on -k regex:calib_normal -s 3 -c 1 -o gpurun_out/r02d_calib -f python bench.py --steps 2 --warmup 3 --batch 8388608 --no-cpu --workload calib > gpurun_out/r02d_ncu.log 2>&1
tail -n 2 gpurun_out/r02d_ncu.log
