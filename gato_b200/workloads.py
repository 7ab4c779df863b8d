"""Synthetic BSQP workloads: the five BASELINE.json configs (inputs per SURVEY.md §8(d)).

Input generators only (numpy, seeded); no solver code.  The figure-8 end-effector reference follows
the reference's `figure8()` (python/bsqp/common.py:10-46, parameters python/bsqp/config.py:12-19) and the
hyper-parameter defaults follow `DEFAULT_SOLVER_PARAMS` (python/bsqp/config.py:35-50).
"""
import numpy as np

NQ = {"indy7": 6, "iiwa14": 7}

# python/bsqp/config.py:35-50
DEFAULT_SOLVER_PARAMS = dict(max_sqp_iters=1, kkt_tol=0.001, max_pcg_iters=200, pcg_tol=1e-4, solve_ratio=1.0, mu=10.0, q_cost=2.0, qd_cost=1e-2, u_cost=2e-6, N_cost=50.0,
                             q_lim_cost=0.01, vel_lim_cost=0.0, ctrl_lim_cost=0.0, rho=0.01)
# python/bsqp/config.py:25
INDY7_READY = np.array([-1.096711, -0.09903229, 0.83125766, -0.10907673, 0.49704404, 0.01499449])


def traj_size(plant, N):
    nq = NQ[plant]
    return 3 * nq * N - nq


def figure8(dt, A_x=0.4, A_z=0.4, offset=(0.0, 0.5, 0.6), period=6, cycles=5, theta=np.pi / 4):
    """End-effector figure-8, [x,y,z,0,0,0] per timestep, flattened (python/bsqp/common.py:10-46)."""
    t = np.linspace(0, 2 * np.pi, int(period / dt))
    unrot = np.stack([offset[0] + A_x * np.sin(t), np.full_like(t, offset[1]), offset[2] + A_z * np.sin(2 * t) / 2 + A_z / 2])
    R = np.array([[np.cos(theta), -np.sin(theta), 0.0], [np.sin(theta), np.cos(theta), 0.0], [0.0, 0.0, 1.0]])
    rot = R @ unrot
    pts = np.zeros((t.size, 6))
    pts[:, :3] = rot.T
    return np.tile(pts.reshape(-1), int(cycles))


def warm_start(x0, N, nq):
    """Every knot = x0, u = 0 (python/bsqp/common.py:93-99).  x0: [B, 2nq] -> [B, traj]."""
    B = x0.shape[0]
    nx, nu = 2 * nq, nq
    xu = np.zeros((B, N, nx + nu), np.float32)
    xu[:, :, :nx] = x0[:, None, :]
    return xu.reshape(B, -1)[:, : (nx + nu) * N - nu].copy()


def make_config(cfg, B=None, N=None):
    """Return dict(plant, N, B, dt, params, xu, xs, ref, extra) for BASELINE.json config index 1..5.

    cfg may also be 'bench' = the headline shape (iiwa14, N=32, B=512) with config-2's input
    distribution and fixed iteration caps.
    """
    p = dict(DEFAULT_SOLVER_PARAMS)
    extra = {}
    if cfg == 1:
        plant, N_, B_, dt, seed = "iiwa14", 8, 1, 0.01, 0
        p.update(max_sqp_iters=5)
    elif cfg in (2, "bench"):
        plant, N_, B_, dt, seed = "iiwa14", 32, (128 if cfg == 2 else 512), 0.01, 1
        p.update(max_sqp_iters=4, max_pcg_iters=50, pcg_tol=-1.0)
    elif cfg == 3:
        plant, N_, B_, dt, seed = "indy7", 32, 512, 0.03125, 2
        p.update(max_sqp_iters=5, pcg_tol=1e-4)
    elif cfg == 4:
        plant, N_, B_, dt, seed = "iiwa14", 128, 1024, 0.01, 3
        p.update(max_sqp_iters=2, max_pcg_iters=200, pcg_tol=1e-5)
    elif cfg == 5:
        plant, N_, B_, dt, seed = "iiwa14", 32, 1024, 0.01, 4  # per-GPU shard of the 8192 batch
    else:
        raise ValueError(cfg)
    N_ = N or N_
    B_ = B or B_
    nq = NQ[plant]
    rng = np.random.default_rng(seed)
    p["dt"] = dt
    fig = figure8(dt).reshape(-1, 6)
    if cfg == 1:
        x0 = np.zeros((B_, 2 * nq), np.float32)
        ref = np.tile(fig[:N_].reshape(1, -1), (B_, 1))
    elif cfg in (2, "bench", 4):
        q = rng.uniform(-0.5, 0.5, (B_, nq))
        qd = rng.uniform(-0.1, 0.1, (B_, nq))
        x0 = np.concatenate([q, qd], 1).astype(np.float32)
        start = rng.integers(0, fig.shape[0] - N_ - 1, B_)
        ref = np.stack([fig[s : s + N_].reshape(-1) for s in start])
    elif cfg == 3:
        q = INDY7_READY[None, :] + rng.normal(0, 0.05, (B_, nq))
        x0 = np.concatenate([q, np.zeros((B_, nq))], 1).astype(np.float32)
        idx = rng.integers(0, fig.shape[0], B_)
        goal = fig[idx].copy()
        goal[:, :3] += rng.uniform(-0.1, 0.1, (B_, 3))
        ref = np.tile(goal[:, None, :], (1, N_, 1)).reshape(B_, -1)
    else:  # cfg 5: one true state + noise; per-solve rho log-spaced and mu in {1,10}
        state = np.zeros(2 * nq)
        x0 = (state[None, :] + rng.normal(0, 0.01, (B_, 2 * nq))).astype(np.float32)
        ref = np.tile(fig[:N_].reshape(1, -1), (B_, 1))
        extra["rho"] = np.logspace(-8, 1, B_).astype(np.float32)
        extra["mu"] = np.where(np.arange(B_) % 2 == 0, 1.0, 10.0).astype(np.float32)
    xu = warm_start(x0, N_, nq)
    return dict(plant=plant, N=N_, B=B_, dt=np.float32(dt), params=p, xu=xu.astype(np.float32), xs=x0.astype(np.float32), ref=ref.astype(np.float32), extra=extra)
