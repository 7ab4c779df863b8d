"""Host-side logic of the multi-GPU closed loop (gato_b200/sharding.py: ShardedMPC) on CPU: a world_size-2 gloo run -- broadcast of the
measurement, one all-gather of the shards' winner records, adoption of the global winner -- with the CPU oracle standing in for the per-GPU
engine (tests/mpc_checker_engine.py) must reproduce the single-process control loop over the whole hypothesis batch bit-for-bit."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

PLANT, N, B_TOTAL, DT, STEPS, TRUE_HYP = "iiwa14", 8, 6, 0.01, 4, 4


def _problem():
    from gato_b200.workloads import DEFAULT_SOLVER_PARAMS, figure8

    p = dict(DEFAULT_SOLVER_PARAMS)
    p.update(max_sqp_iters=1, max_pcg_iters=40, dt=DT)
    fext = np.zeros((B_TOTAL, 6), np.float32)
    fext[1:, 2] = [4.0, -6.0, 9.0, 2.0, -3.0]
    rho = np.logspace(-3, 0, B_TOTAL).astype(np.float32)
    return p, fext, rho, figure8(DT).reshape(-1, 6)


def _loop(engine_of, dist_mod, rank, world):
    """The closed loop of SURVEY.md section 8(d) cfg 5: the plant is hypothesis TRUE_HYP of the solver's own simulator."""
    from gato_b200.sharding import ShardedMPC
    from oracle.pyapi import Backend, dims

    p, fext, rho, fig = _problem()
    d = dims(PLANT, N)
    nx, nu = d["nx"], d["nu"]
    nb = B_TOTAL // world
    sl = slice(rank * nb, (rank + 1) * nb)
    be = Backend("oracle", PLANT, N)
    be.set_threads(1)
    s = be.solver(nb, p)
    s.set_batch("f_ext", fext[sl])
    s.set_batch("rho", rho[sl], True)
    s.reset("dual")
    plant = be.solver(B_TOTAL, p)  # the "real" robot: every rank simulates it identically (rank 0's copy is the one that counts)
    plant.set_batch("f_ext", fext)
    x = np.zeros(nx, np.float32)
    xu0 = np.zeros(d["traj"], np.float32)
    mpc = ShardedMPC(engine_of(s, d, nb, xu0), DT, dist_mod)
    out = [mpc.step(x, fig[:N].reshape(-1), None, None, 0.0, reset_rho=False)]
    for step in range(1, STEPS + 1):
        x_last, u_last = x.copy(), out[-1]["XU_best"][nx:nx + nu].copy()
        x = plant.sim_forward(x_last, u_last, DT)[TRUE_HYP].copy()
        out.append(mpc.step(x, fig[step:step + N].reshape(-1), x_last, u_last, DT))
    return np.array([o["best_id"] for o in out]), np.stack([o["XU_best"] for o in out]), np.array([o["best_error"] for o in out])


def _worker(rank, world, port, outfile):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpc_checker_engine import CheckerShardEngine

    ids, xus, errs = _loop(CheckerShardEngine, dist, rank, world)
    np.savez(f"{outfile}.{rank}.npz", ids=ids, xus=xus, errs=errs)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_closed_loop_equals_single_process(oracle_built, tmp_path):
    from mpc_checker_engine import CheckerShardEngine

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "loop")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    ids1, xus1, errs1 = _loop(CheckerShardEngine, None, 0, 1)
    assert TRUE_HYP in ids1[1:], "the scoring never identified the true hypothesis"
    for r in range(2):
        g = np.load(f"{out}.{r}.npz")
        assert np.array_equal(g["ids"], ids1) and np.array_equal(g["xus"], xus1) and np.array_equal(g["errs"], errs1), f"rank {r}"
