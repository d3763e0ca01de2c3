"""BASELINE configs[2]: CCX matrix of N events, row-block sharded over the ranks of one box.
    torchrun --nproc-per-node G experiments/ccx_multi_gpu.py [N]
Each rank computes the row block parallel.ccx_row_blocks gives it (equal pair counts), then the
blocks are all-gathered over NCCL; rank 0 checks the result against a single-GPU run of a
subset and prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from detex_b200 import parallel, synth  # noqa: E402
from detex_b200.engine import Engine  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local)
    X = synth.event_families(3003, max(1, N // 64), 64, 1000, 3, max_shift=100)[:N]
    blocks = parallel.ccx_row_blocks(N, world)
    b0, b1 = blocks[rank]
    for rep in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cc, lag, sub = eng.ccx(X, 3, row_begin=b0, row_end=b1, engine="tcgen05")
        t_local = time.perf_counter() - t0
        cc_all = parallel.gather_row_blocks(cc, blocks, N)
        lag_all = parallel.gather_row_blocks(lag, blocks, N)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt, t_local], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, t_local = float(t[0]), float(t[1])
    if rank == 0:
        # spot check: rows 0..7 and the last rows against a direct single-rank computation
        c0, l0, _ = eng.ccx(X, 3, row_begin=0, row_end=8, engine="tcgen05")
        ok = np.array_equal(l0[:, 8:], lag_all[:8, 8:]) and np.array_equal(c0[:, 8:], cc_all[:8, 8:])
        pairs = N * (N - 1) // 2
        print(json.dumps({"config": "configs[2] CCX %d events, row-block sharded" % N, "n_gpus": world,
                          "seconds_total": dt, "seconds_compute_max_rank": t_local, "pairs_per_s": pairs / dt,
                          "pair_lags_per_s": pairs * 1001 / dt, "blocks": blocks, "matches_single_rank": bool(ok)}))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
