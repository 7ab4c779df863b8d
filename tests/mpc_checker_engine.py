"""TEST INFRASTRUCTURE: a CPU stand-in for one shard of gato_b200.sharding.ShardedMPC built on any solver object with solve / sim_forward /
reset (the CPU oracle in the gloo tests): the same three phases and the same winner-record layout as the C ABI's gato_mpc_local_async /
gato_mpc_adopt_async / gato_mpc_wait, composed the way the reference composes the step (python/bsqp/mpc_controller.py:233-253, 294-309)."""
import numpy as np
import torch


class CheckerShardEngine:
    def __init__(self, solver, d, batch, xu0):
        self.s, self.d, self.batch = solver, d, batch
        self.device = torch.device("cpu")
        self.in_floats = 2 * d["nx"] + 6 * d["N"] + d["nu"]
        self.rec_floats = (4 + d["traj"] + 3) // 4 * 4
        self.XU = np.tile(np.asarray(xu0, np.float32), (batch, 1))
        self.offsets = None

    def local_async(self, inp, score, sim_dt, dt, reset_rho, rec):
        d, B = self.d, self.batch
        v = inp.numpy()
        nx, nu, N = d["nx"], d["nu"], d["N"]
        x_curr, ref = v[:nx], v[nx:nx + 6 * N]
        xs = np.tile(x_curr, (B, 1))
        if self.offsets is not None:
            xs = (xs + self.offsets).astype(np.float32)
        self.XU[:, :nx] = xs
        if reset_rho:
            self.s.reset("rho")
        self.res = self.s.solve(self.XU, xs, np.tile(ref, (B, 1)), dt)
        self.XU[:, :] = self.res["XU"]
        self.errors, best = np.zeros(B), 0
        if score:
            x_next = self.s.sim_forward(v[nx + 6 * N:2 * nx + 6 * N], v[2 * nx + 6 * N:], sim_dt)
            self.errors = np.linalg.norm(x_next.astype(np.float64) - x_curr.astype(np.float64)[None, :], axis=1)
            best = int(np.argmin(self.errors))
        r = rec.numpy()
        r[:2] = np.array([self.errors[best]], np.float64).view(np.float32)
        r[2:3] = np.array([best], np.int32).view(np.float32)
        r[4:4 + d["traj"]] = self.XU[best]

    def adopt_async(self, recs, n, id_stride):
        r = recs.numpy().reshape(n, self.rec_floats)
        errs = np.array([r[i, :2].copy().view(np.float64)[0] for i in range(n)])
        win = int(np.argmin(errs))  # first minimum, first NaN wins
        self.best_id = win * id_stride + int(r[win, 2:3].copy().view(np.int32)[0])
        self.best_err = errs[win]
        self.XU[:, :] = r[win, 4:4 + self.d["traj"]]

    def wait(self):
        out = dict(self.res)
        out.update(best_id=self.best_id, best_error=self.best_err, errors=self.errors, XU_best=self.XU[0].copy())
        return out
