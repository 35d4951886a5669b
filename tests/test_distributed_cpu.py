"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous sharding and the
single all-reduce of the packed calibration normal equations followed by the redundant
18x18 solve on every rank.  The per-shard partial sums come from the oracle here (the GPU
kernel that produces them is covered by the -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_bounds_cover_everything():
    from roboticmanipulators_b200._shard import shard_bounds
    for n in (0, 1, 7, 8, 9, 240, 10**9 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import np_oracle as O
    from roboticmanipulators_b200._shard import shard_bounds
    from test_host_logic_cpu import fanuc_calibrator
    from test_oracle_golden import _calib_model
    g = np.load(os.path.join(ROOT, "tests", "golden", "calibration.npz"))
    mc = fanuc_calibrator()
    mc.get_reduced_model()
    model = _calib_model(g)
    lo, hi = shard_bounds(len(g["qs"]), rank, world)
    H, gv = O.calib_normal_equations(model, g["qs"][lo:hi], g["pts"][lo:hi], np.zeros(18),
                                     g["T_base"], list(g["T_tools"]))
    packed = torch.tensor(np.concatenate([H[np.triu_indices(18)], gv, [0.0]]))
    mc._allreduce(packed, dist.ReduceOp.SUM, None)           # the ONE collective of the path
    Hs, gs, _ = mc.unpack_normal_equations(packed.numpy(), 18)
    params = np.linalg.inv(Hs).dot(gs)
    # the max-type metrics of evaluate() reduce with MAX
    mx = torch.tensor([float(rank + 1)])
    mc._allreduce(mx, dist.ReduceOp.MAX, None)
    # evaluate() on sharded measurements: the residual sums AND the point count are summed over the
    # ranks, the maxima are MAX-reduced; every rank then reports the whole data set's RMS values
    diffs = []
    for j, Tt in enumerate(g["T_tools"]):
        T = O.calib_pose(model, g["qs"][lo:hi], g["params"], g["T_base"], Tt)
        diffs.append(g["pts"][lo:hi, 3 * j:3 * j + 3] - T[:, :3, 3])
    d = np.abs(np.concatenate(diffs))
    nrm = np.linalg.norm(d, axis=1)
    local = torch.tensor([(nrm * nrm).sum(), (d[:, 0]**2).sum(), (d[:, 1]**2).sum(), (d[:, 2]**2).sum(),
                          nrm.max(), d[:, 0].max(), d[:, 1].max(), d[:, 2].max(), float(len(d))])
    red = mc._reduce_residual_sums(local).numpy()
    metrics = mc._metrics(red, red[8])
    q.put((rank, params, float(np.abs(Hs - g["J_sum"]).max() / np.abs(g["J_sum"]).max()), float(mx), metrics))
    dist.destroy_process_group()


def test_two_rank_normal_equation_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.load(os.path.join(ROOT, "tests", "golden", "calibration.npz"))
    want = dict(zip([str(k) for k in g["eval_keys"]], g["eval_after_params"]))
    for rank, params, herr, mx, metrics in results:
        assert herr < 1e-12
        assert np.max(np.abs(params - g["params"])) < 1e-9     # every rank solves the same system
        assert mx == 2.0
        for k, v in want.items():                               # the reference's evaluate() on all 240 rows
            assert abs(metrics[k] - v) < 1e-9, (k, metrics[k], v)
    assert np.array_equal(results[0][1], results[1][1])        # bit-identical across ranks


def test_reference_arm_under_torchrun_prints_one_line():
    """`bench.py --impl reference` launched the way the driver launches it for N > 1 (torchrun, one
    rank per GPU): rank 0 alone times the CPU path and prints ONE JSON line with the contract's
    keys; the other ranks exit 0 without output.  Needs no GPU."""
    import json
    import subprocess
    port = _free_port()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["higher_is_better"] is True
    assert d["unit"] == "poses/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"] == "fanuc165f_fk_jac_det_ik8"
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"]
