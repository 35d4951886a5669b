N=${1:-8}
set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/probe_host_link.py > gpurun_out/r02_hostlink_${N}gpu.json 2> gpurun_out/r02_hostlink_${N}gpu.err; tail -n 3 gpurun_out/r02_hostlink_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; tail -n 5 gpurun_out/r02_bench_${N}gpu.err
nvidia-smi topo -m > gpurun_out/r02_topo_${N}gpu.txt 2>&1
lscpu | head -n 25 > gpurun_out/r02_lscpu_${N}gpu.txt; free -g >> gpurun_out/r02_lscpu_${N}gpu.txt
