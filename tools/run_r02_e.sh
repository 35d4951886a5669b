set -x
python -m pytest tests -m gpu -q 2>&1 | tail -n 40
python tools/bench_rows.py > gpurun_out/r02e_rows.json 2> gpurun_out/r02e_rows.err; tail -n 3 gpurun_out/r02e_rows.err
python -c "
import json; d=json.load(open('gpurun_out/r02e_rows.json'))
for k,v in d.items(): print(k, round(v['ms'],3), round(v['hbm_frac'],3))"
