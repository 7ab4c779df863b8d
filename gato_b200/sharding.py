"""Multi-GPU plumbing for closed-loop batches: the batch of independent solves shards across ranks with NO collective on the
hot path; torch.distributed (NCCL over NVLink on GPUs, gloo in CPU tests) is used only to broadcast the
shared inputs and to gather the shards' winners for best-trajectory selection (mpc_controller.py:240-242, 294-309).

`ShardedMPC` is the closed-loop control step of BASELINE.json config 5 (figure-8 tracking, iiwa14 N=32, 8192 sampled hypotheses
over 8 GPUs): per control step ONE broadcast of the measured state and reference window (227 floats), the local step of every
shard (`gato_mpc_local_async`: solve + hypothesis scoring on the device), ONE all-gather of the shards' winner records
(error, id and the 2.7 KB trajectory of each shard's best hypothesis: 21 KB in total at 8 ranks instead of the 21.8 MB of all
trajectories), and the adoption of the global winner on every rank (`gato_mpc_adopt_async`).  The engine behind it is anything
with the three phase calls: `NativeShardEngine` (the CUDA library), or a checker in the tests.
"""
import numpy as np


def shard_range(batch, rank, world):
    """Contiguous row block of `batch` solves owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_inputs(dist, tensors, src=0):
    """Broadcast x0 / reference window tensors from `src` to every rank (in place)."""
    for t in tensors:
        dist.broadcast(t, src=src)
    return tensors


def gather_rows(dist, local, batch, world):
    """All-gather per-rank row blocks (possibly ragged) into the full [batch, ...] tensor on every rank."""
    import torch

    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[:n] for o, n in zip(outs, sizes)], dim=0)


def select_best(merits):
    """Index of the best (lowest final merit, NaN-safe) solve — the selection the MPC harness makes per control step."""
    m = np.where(np.isnan(merits), np.inf, merits)
    return int(np.argmin(m))


class NativeShardEngine:
    """One rank's shard of the hypothesis batch on its GPU: gato_b200.native.Solver created on torch's current stream."""

    def __init__(self, solver):
        import torch

        self.s, self.device = solver, torch.device("cuda", torch.cuda.current_device())
        self.in_floats, self.rec_floats, self.batch = solver.mpc_input_floats(), solver.mpc_record_floats(), solver.B

    def local_async(self, inp, score, sim_dt, dt, reset_rho, rec):
        self.s.mpc_local_async(inp.data_ptr(), score, sim_dt, dt, reset_rho, rec.data_ptr())

    def adopt_async(self, recs, n, id_stride):
        self.s.mpc_adopt_async(recs.data_ptr(), n, self.rec_floats, id_stride)

    def wait(self):
        return self.s.mpc_wait()


class ShardedMPC:
    """Closed-loop control step over `dist` (an initialised torch.distributed module, or None for one rank).  Every rank owns `engine.batch`
    hypotheses (equal shards); rank 0 holds the measurement.  The result dictionary's best_id is the GLOBAL hypothesis index."""

    def __init__(self, engine, dt, dist=None):
        import torch

        self.e, self.dt, self.dist = engine, float(dt), dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        dev = engine.device
        self.inp = torch.zeros(engine.in_floats, dtype=torch.float32, device=dev)
        self.rec = torch.zeros(engine.rec_floats, dtype=torch.float32, device=dev)
        self.recs = torch.zeros(self.world * engine.rec_floats, dtype=torch.float32, device=dev)
        self.h_in = torch.zeros(engine.in_floats, dtype=torch.float32)
        if dev.type == "cuda":
            self.h_in = self.h_in.pin_memory()

    def step(self, x_curr, ref_window, x_last=None, u_last=None, sim_dt=0.0, reset_rho=True):
        import torch

        score = x_last is not None
        if self.rank == 0:
            parts = [np.asarray(x_curr, np.float32).ravel(), np.asarray(ref_window, np.float32).ravel()]
            if score:
                parts += [np.asarray(x_last, np.float32).ravel(), np.asarray(u_last, np.float32).ravel()]
            v = np.concatenate(parts)
            self.h_in[: v.size] = torch.from_numpy(v)
            self.inp.copy_(self.h_in, non_blocking=True)
        if self.dist is not None:
            self.dist.broadcast(self.inp, src=0)  # measured state, reference window, last state / control: 227 floats
        self.e.local_async(self.inp, score, sim_dt, self.dt, reset_rho, self.rec)
        if self.dist is not None:
            self.dist.all_gather_into_tensor(self.recs, self.rec)  # one winner record per shard
            recs = self.recs
        else:
            recs = self.rec
        self.e.adopt_async(recs, self.world, self.e.batch)
        return self.e.wait()
