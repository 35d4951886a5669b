python -m pytest tests/test_r2_gpu.py tests/test_arms_chain_gpu.py tests/test_calibration_gpu.py -m gpu -q 2>&1 | tail -n 5
python -c "
import sys; sys.path.insert(0,'.')
from roboticmanipulators_b200 import _lib
import torch
from roboticmanipulators_b200.utils.robo_math import Transformation
v=torch.rand(1000,4,device='cuda',dtype=torch.float64)
T=Transformation.batch('RzTxRzTx', v); torch.cuda.synchronize()
print('jit status:', _lib.load().rk_chain_jit_status().decode()[:1500])
"
python tools/bench_rows.py 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
for k,v in d.items(): print(k, round(v['ms'],3), round(v['hbm_frac'],3))"
