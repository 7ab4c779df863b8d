"""Multi-GPU host logic on CPU: world_size-2 gloo run of the shard / broadcast / gather plumbing (gato_b200/sharding.py).
Each rank solves its contiguous row block with the oracle standing in for the per-GPU solver; the gathered result must equal
the single-process result bit-for-bit (solves are independent: no collective on the hot path)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from gato_b200.sharding import gather_rows, select_best, shard_range


def test_shard_range_partitions():
    for B in (1, 7, 512, 8192):
        for W in (1, 2, 3, 8):
            r = [shard_range(B, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == B and all(r[i][1] == r[i + 1][0] for i in range(W - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    assert select_best(np.array([3.0, np.nan, 1.0, 2.0])) == 2


def _worker(rank, world, port, out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gato_b200.sharding import broadcast_inputs
    from gato_b200.workloads import make_config
    from oracle.pyapi import Backend

    B = 5  # ragged split: 3 + 2
    w = make_config(1, B=B)
    xs = torch.from_numpy(w["xs"].copy()) if rank == 0 else torch.zeros(B, 14)
    ref = torch.from_numpy(w["ref"].copy()) if rank == 0 else torch.zeros(B, 48)
    broadcast_inputs(dist, [xs, ref], src=0)
    lo, hi = shard_range(B, rank, world)
    be = Backend("oracle", "iiwa14", 8)
    be.set_threads(1)
    o = be.solver(hi - lo, w["params"]).solve(w["xu"][lo:hi], xs.numpy()[lo:hi], ref.numpy()[lo:hi], w["dt"])
    xu_all = gather_rows(dist, torch.from_numpy(o["XU"]), B, world)
    merit_all = gather_rows(dist, torch.from_numpy(o["final_merit"]).reshape(-1, 1), B, world)
    if rank == 0:
        np.save(out, np.concatenate([xu_all.numpy(), merit_all.numpy()], axis=1))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_solve_equals_single_process(oracle_built, tmp_path):
    from gato_b200.workloads import make_config
    from oracle.pyapi import Backend

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    w = make_config(1, B=5)
    be = Backend("oracle", "iiwa14", 8)
    ref = be.solver(5, w["params"]).solve(w["xu"], w["xs"], w["ref"], w["dt"])
    assert np.array_equal(got[:, :-1], ref["XU"]) and np.array_equal(got[:, -1], ref["final_merit"])
