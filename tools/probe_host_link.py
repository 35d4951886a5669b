#!/usr/bin/env python
"""Host <-> device link probe (what the end-to-end number of bench.py can reach at most).

Measures, on this rank's GPU, with pinned host buffers and cudaMemcpyAsync (torch copy_):
H2D and D2H bandwidth by size, both directions at once on two streams, the cost of splitting
a transfer in several buffers, pageable copies, and the cost of pinning.  Prints one JSON object.
Under torchrun every rank measures at the same time (that is the point: shared root ports / host
memory show up as a per-rank drop) and rank 0 prints the list.

  python tools/probe_host_link.py            # one GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/probe_host_link.py
"""
import json
import os
import subprocess
import time

import torch


def bw(fn, nbytes, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = max(best, nbytes / (time.perf_counter() - t0) / 1e9)
    return round(best, 2)


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"rank": rank, "cpu_count": os.cpu_count(),
           "affinity": len(os.sched_getaffinity(0))}
    t0 = time.perf_counter()
    big = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    out["pin_1GiB_s"] = round(time.perf_counter() - t0, 3)
    dbig = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    big2 = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    dbig2 = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    if dist is not None:
        dist.barrier()
    for mb in (1, 8, 64, 512, 1024):
        n = mb << 20
        out[f"h2d_{mb}MB"] = bw(lambda: dbig[:n].copy_(big[:n], non_blocking=True), n)
        out[f"d2h_{mb}MB"] = bw(lambda: big[:n].copy_(dbig[:n], non_blocking=True), n)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both(n):
        with torch.cuda.stream(s1):
            dbig[:n].copy_(big[:n], non_blocking=True)
        with torch.cuda.stream(s2):
            big2[:n].copy_(dbig2[:n], non_blocking=True)
    for mb in (64, 512):
        n = mb << 20
        out[f"bidir_{mb}MB_each_way_sum"] = bw(lambda: both(n), 2 * n)
    # 53 MB in, 856 MB out (the e2e step of bench.py at 2^20 poses) on two streams
    out["e2e_shape_50in_856out"] = bw(lambda: (both_shape(s1, s2, dbig, big, big2, dbig2)), (50 + 856) << 20)
    # one transfer split in 16 pieces on one stream
    n = 512 << 20
    piece = n // 16
    out["d2h_512MB_in_16_pieces"] = bw(
        lambda: [big[i * piece:(i + 1) * piece].copy_(dbig[i * piece:(i + 1) * piece], non_blocking=True)
                 for i in range(16)], n)
    page = torch.empty(256 << 20, dtype=torch.uint8)
    out["h2d_pageable_256MB"] = bw(lambda: dbig[:256 << 20].copy_(page), 256 << 20, reps=3)
    out["d2h_pageable_256MB"] = bw(lambda: page.copy_(dbig[:256 << 20]), 256 << 20, reps=3)
    # host memcpy rate (a user that gathers / scatters reference-shaped arrays on the host)
    t0 = time.perf_counter()
    big2.copy_(big)
    out["host_memcpy_1GiB_gbs"] = round((1 << 30) / (time.perf_counter() - t0) / 1e9, 2)
    if rank == 0:
        for cmd, key in ((["nvidia-smi", "topo", "-m"], "topo"), (["numactl", "--hardware"], "numa"),
                         (["lscpu"], "lscpu")):
            try:
                out[key] = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout[-3000:]
            except (OSError, subprocess.TimeoutExpired):
                out[key] = None
    if dist is not None:
        gathered = [None] * world
        dist.all_gather_object(gathered, out)
        dist.destroy_process_group()
        out = gathered
    if rank == 0:
        print(json.dumps(out, indent=1))


def both_shape(s1, s2, dbig, big, big2, dbig2):
    with torch.cuda.stream(s1):
        dbig[:50 << 20].copy_(big[:50 << 20], non_blocking=True)
    with torch.cuda.stream(s2):
        big2[:856 << 20].copy_(dbig2[:856 << 20], non_blocking=True)


if __name__ == "__main__":
    main()
