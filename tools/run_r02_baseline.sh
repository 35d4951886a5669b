# Round-2 baseline evidence (run on a GPU box): ncu --set full of every kernel that round 1 left
# below 0.9 of its roofline, the host-link probe, and the plain timings of the same rows.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; tail -n 3 gpurun_out/r02a_pytest.log
python tools/probe_host_link.py > gpurun_out/r02a_hostlink.json 2> gpurun_out/r02a_hostlink.err
python tools/bench_rows.py > gpurun_out/r02a_rows.json 2> gpurun_out/r02a_rows.err
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'arm_|chain_fk|traj_kernel|trap_|poly5|calib_' -o gpurun_out/r02a_rows -f \
  python tools/bench_rows.py --warmup 0 --reps 1 --scale 0.25 > gpurun_out/r02a_rows_ncu.log 2>&1
tail -n 2 gpurun_out/r02a_rows_ncu.log
for v in "f32fused:--dtype f32" "ikaos:--workload ik --layout aos" "ikf32:--workload ik --dtype f32" "ikf32aos:--workload ik --dtype f32 --layout aos"; do
  name=${v%%:*}; flags=${v#*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fanuc -s 3 -c 1 \
    -o gpurun_out/r02a_$name -f python bench.py $flags --steps 2 --warmup 3 --batch 8388608 --no-cpu \
    > gpurun_out/r02a_$name.log 2>&1
  tail -n 1 gpurun_out/r02a_$name.log
done
ls -la gpurun_out | tail -n 20
