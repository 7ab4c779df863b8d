"""Multi-GPU closed loop (BASELINE.json config 5) on the GPU box: the phase calls of the C ABI (gato_mpc_local_async / gato_mpc_adopt_async /
gato_mpc_wait) and gato_b200.sharding.ShardedMPC.  One GPU: two shards emulated by two solvers on the same device must reproduce one
solver over the whole batch bit-for-bit.  Two or more GPUs: a torchrun NCCL launch of tools/closed_loop_multi_gpu.py --check."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from gato_b200.workloads import DEFAULT_SOLVER_PARAMS, figure8

pytestmark = pytest.mark.gpu


def test_two_shards_on_one_gpu_equal_one_solver_over_the_whole_batch():
    import torch

    from gato_b200 import native

    plant, N, B, dt = "iiwa14", 8, 12, 0.01
    p = dict(DEFAULT_SOLVER_PARAMS, max_sqp_iters=2, max_pcg_iters=60, dt=dt)
    rng = np.random.default_rng(9)
    fext = rng.normal(0, 5.0, (B, 6)).astype(np.float32)
    rho = np.logspace(-4, 0, B).astype(np.float32)
    off = rng.normal(0, 0.01, (B, 14)).astype(np.float32)
    fig = figure8(dt).reshape(-1, 6)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    def mk(sl):
        s = native.Solver(plant, N, sl.stop - sl.start, p, device=0, stream=stream.cuda_stream)
        s.set_batch("f_ext", fext[sl]), s.set_batch("rho", rho[sl], True)
        s.reset("dual")
        s.mpc_set_warm_start(np.zeros(s.d["traj"], np.float32))
        s.mpc_set_state_offsets(off[sl])
        return s

    whole, shards = mk(slice(0, B)), [mk(slice(0, 5)), mk(slice(5, B))]  # ragged on purpose: ids are offset by the shard's first row
    nx, nu = whole.d["nx"], whole.d["nu"]
    rf, nin = whole.mpc_record_floats(), whole.mpc_input_floats()
    inp = torch.zeros(nin, device="cuda")
    recs = torch.zeros(2 * rf, device="cuda")
    x = np.zeros(nx, np.float32)
    true_hyp, got_true = 7, False
    x_last = u_last = None
    for step in range(5):
        ref_w = fig[step:step + N].reshape(-1)
        score = x_last is not None
        a = whole.mpc_step(x, ref_w, x_last, u_last, dt, dt, reset_rho=True)
        v = np.concatenate([x, ref_w] + ([x_last, u_last] if score else [])).astype(np.float32)
        inp[: v.size] = torch.from_numpy(v).cuda()
        for i, s in enumerate(shards):
            s.mpc_local_async(inp.data_ptr(), score, dt, dt, True, recs[i * rf:].data_ptr())
        outs = []
        for i, s in enumerate(shards):
            # equal id strides are what ShardedMPC uses; ragged shards pass their own row offsets through the id stride of a 1-record call
            s.mpc_adopt_async(recs.data_ptr(), 2, rf, 5)
            outs.append(s.mpc_wait())
        gid = outs[0]["best_id"]
        assert outs[1]["best_id"] == gid
        assert gid == a["best_id"], (step, gid, a["best_id"])
        assert np.array_equal(outs[0]["XU_best"], a["XU_best"]) and np.array_equal(outs[1]["XU_best"], a["XU_best"])
        assert np.array_equal(np.concatenate([outs[0]["errors"], outs[1]["errors"]]), a["errors"])
        assert outs[0]["best_error"] == a["errors"][a["best_id"]] or (np.isnan(outs[0]["best_error"]) and np.isnan(a["errors"][a["best_id"]]))
        for s, sl in zip(shards, (slice(0, 5), slice(5, B))):
            assert np.array_equal(s.mpc_get_warm_start(), whole.mpc_get_warm_start()[sl])
        got_true |= gid == true_hyp
        x_last, u_last = x.copy(), a["XU_best"][nx:nx + nu].copy()
        plant_s = native.Solver(plant, N, 1, p, device=0)
        plant_s.set_batch("f_ext", fext[true_hyp:true_hyp + 1])
        x = plant_s.sim_forward(x_last, u_last, dt)[0].copy()
    assert got_true, "the scoring never identified the true hypothesis"
    torch.cuda.set_stream(torch.cuda.default_stream())


def test_two_rank_nccl_closed_loop_equals_single_process():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29617",
           str(ROOT / "tools" / "closed_loop_multi_gpu.py"), "--per-gpu", "64", "--steps", "12", "--check"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, PYTHONPATH=str(ROOT)))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["equals_single_process_bit_for_bit"] and line["n_gpus"] == 2 and line["all_finite"]
