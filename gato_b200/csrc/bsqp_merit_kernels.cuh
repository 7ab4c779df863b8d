// k_merit_ls / k_sim_forward of the BSQP path (see bsqp_ctx.cuh for the kernel map)
#pragma once
#include "bsqp_ctx.cuh"
#include "items.cuh"
#include "rt_slots.cuh"

namespace gato {

// =====================================================================================================
// k_merit_ls: one CTA per solve, thread per (alpha, knot) — computeMeritBatchedKernel + lineSearchAndUpdateBatchedKernel
// NA = 8: merit at z + 2^-a dz for a = 0..7, then the line search.  NA = 1: merit at z (initial / final merit).
// =====================================================================================================
// SPLIT (small batches, where a launch lasts as long as one thread's instruction stream): two threads per (alpha, knot) -- one evaluates the
// forward dynamics and the defect, the other the tracking cost -- combined as fmaf(mu, defect, cost) exactly like the single-thread version.
#ifndef GATO_RT_MERIT_MIN_BLOCKS
#define GATO_RT_MERIT_MIN_BLOCKS 2  // table-driven instantiations (their per-thread state lives in local memory: fewer resident threads, more of it in L1)
#endif
template<class P, int NA, bool SPLIT = false>
__global__ void __launch_bounds__(SPLIT ? 512 : (NA == 1 ? 128 : 256), (NA == 1 || SPLIT) ? 1 : (is_rt_plant<P> ? GATO_RT_MERIT_MIN_BLOCKS : GATO_MERIT_MIN_BLOCKS)) k_merit_ls(Ctx c)
{
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ;
        const bool overlap = NA > 1 && (c.flags & F_OVERLAP) != 0;
        // the iteration that meets the early-exit test skips merit + line search (bsqp.cuh:165).  Launched after the PCG launch has completed the
        // test is final; overlapped with it (F_OVERLAP) only the earlier iterations are, and this iteration's test is decided before the line search
        if (NA > 1 && stopped_before(c, overlap ? c.it : c.it + 1)) return;
        extern __shared__ float smf[];  // [NA][N] per-knot merits, NA sums, (SPLIT: [NA][N] cost halves)
        const int               N = c.N, b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
        if (overlap) {
                // this solve's primal step comes from a PCG CTA that may still be running: wait for its hand-over flag
                if (tid == 0) wait_flag(c.pcg_done + (size_t)c.it * c.B + b);
                __syncthreads();
        }
        const int               traj = (NX + NU) * N - NU;
        float*                  mk = smf;
        float*                  msum = smf + NA * N;
        float*                  mcost = msum + NA;
        const float*            xu = c.xu + (size_t)b * traj;
        const float*            dz = c.dz + (size_t)b * traj;
        if (c.flags & F_MERIT) {
                const float mu = c.mu[b];
                const bool  zero_dz = (c.flags & F_ZERO_DZ) != 0;
                float       fext[6];
                sfor<0, 6>([&](auto ic) { fext[ic] = c.fext[6 * b + ic]; });
                const Items<P> it = make_items<P>(c);
                for (int w0 = tid; w0 < (SPLIT ? 2 : 1) * NA * N; w0 += T) {
                        const int   half = SPLIT ? w0 / (NA * N) : 0, w = SPLIT ? w0 % (NA * N) : w0;
                        const int   a = w / N, k = w % N;
                        const float alpha = (float)(1.0 / (double)(1 << a));
                        float       ref3[3];
                        sfor<0, 3>([&](auto ic) { ref3[ic] = c.ref[(size_t)b * 6 * N + 6 * k + ic]; });
                        float       xux[2 * NX + NU];
                        const float *xk = xu + (size_t)k * (NX + NU), *dk = dz + (size_t)k * (NX + NU);
                        float        m;
                        if (k < N - 1) {
                                if (zero_dz)  // dz == 0 (bsqp.cuh:112,180): z + 1*0 = z, skip the loads
                                        sfor<0, 2 * NX + NU>([&](auto ic) { xux[ic] = xk[ic]; });
                                else
                                        sfor<0, 2 * NX + NU>([&](auto ic) { xux[ic] = fmaf(alpha, dk[ic], xk[ic]); });
                                if constexpr (SPLIT)
                                        m = half == 0 ? it.merit_mid_cons(xux, fext, c.dt) : it.template tracking_cost<false>(xux, ref3, c.cs);
                                else
                                        m = it.merit_mid(xux, ref3, mu, fext, c.dt, c.cs);
                        } else {
                                float e0[NX];
                                if (zero_dz) {
                                        sfor<0, NX>([&](auto ic) { xux[ic] = xk[ic]; });
                                        sfor<0, NX>([&](auto ic) { e0[ic] = fabsf(xu[ic] - c.xs[(size_t)b * NX + ic]); });
                                } else {
                                        sfor<0, NX>([&](auto ic) { xux[ic] = fmaf(alpha, dk[ic], xk[ic]); });
                                        sfor<0, NX>([&](auto ic) { e0[ic] = fabsf(fmaf(alpha, dz[ic], xu[ic]) - c.xs[(size_t)b * NX + ic]); });
                                }
                                if constexpr (SPLIT)
                                        m = half == 0 ? it.merit_last_cons(e0) : it.template tracking_cost<true>(xux, ref3, c.cs);
                                else
                                        m = it.merit_last(xux, ref3, mu, e0, c.cs);
                        }
                        if (SPLIT && half == 1)
                                mcost[a * N + k] = m;
                        else
                                mk[a * N + k] = m;
                }
                __syncthreads();
                if constexpr (SPLIT) {
                        for (int w = tid; w < NA * N; w += T) mk[w] = fmaf(mu, mk[w], mcost[w]);
                        __syncthreads();
                }
                if (tid < NA) {
                        // the reference sums the knots with unordered float atomics (merit.cuh:88-91); here: ascending k
                        float s = 0.0f;
                        for (int k = 0; k < N; k++) s = s + mk[tid * N + k];
                        msum[tid] = s;
                        if (NA == 1)
                                c.merit_cur[b] = s;
                        else
                                c.merit[(size_t)b * NA + tid] = s;
                }
                __syncthreads();
        } else if (NA > 1) {
                if (tid < NA) msum[tid] = c.merit[(size_t)b * NA + tid];
                __syncthreads();
        }
        if constexpr (NA > 1) {
                if (!(c.flags & F_LS)) return;
                __shared__ float s_step;
                __shared__ int   s_ok;
                if (overlap) {
                        // the merits above were computed speculatively; the line search only runs if this iteration does not meet the exit test
                        __shared__ int s_stop;
                        if (tid == 0) s_stop = stop_decided_early(c, c.it) ? 1 : 0;
                        __syncthreads();
                        if (s_stop) return;
                }
                if (tid == 0) {
                        // first strict minimum over the 8 merits; NaN / >= 1e38 count as 1e38 at index 0 (line_search.cuh:23-55)
                        float best = 1e38f;
                        int   bi = 0;
                        {
                                float mer[NA];
                                int   idx[NA];
                                for (int i = 0; i < NA; i++) {
                                        float lm = 1e38f;
                                        int   li = 0;
                                        if (msum[i] < lm) {
                                                lm = msum[i];
                                                li = i;
                                        }
                                        mer[i] = lm, idx[i] = li;
                                }
                                for (int s = 1; s < NA; s *= 2)
                                        for (int t = 0; 2 * s * t + s < NA; t++) {
                                                const int index = 2 * s * t;
                                                if (mer[index + s] < mer[index]) {
                                                        mer[index] = mer[index + s];
                                                        idx[index] = idx[index + s];
                                                }
                                        }
                                best = mer[0], bi = idx[0];
                        }
                        const bool ok = best < c.merit_cur[b];
                        float      rho = c.rho[b];
                        if (c.adapt) {
                                const float d = c.drho[b];
                                const float mult = ok ? fminf(d / kRhoFactor, 1.0f / kRhoFactor) : fmaxf(d * kRhoFactor, kRhoFactor);
                                c.drho[b] = mult;
                                rho = fmaxf(rho * mult, kRhoMin);
                                rho = fminf(rho, kRhoMax);
                        }
                        float st;
                        if (!ok) {
                                if (rho > kRhoMax) rho = kRhoInit;
                                st = -1.0f;
                        } else {
                                st = (float)(1.0 / (double)(float)(1 << bi));
                                c.merit_cur[b] = best;
                        }
                        c.rho[b] = rho;
                        c.step[b] = st;
                        if (c.ls_merit_log) {
                                c.ls_merit_log[(size_t)c.it * c.B + b] = ok ? best : c.merit_cur[b];
                                c.ls_step_log[(size_t)c.it * c.B + b] = st;
                        }
                        s_step = st;
                        s_ok = ok ? 1 : 0;
                        // the merit buffer is zeroed by the reference here (line_search.cuh:30); ours is overwritten, not accumulated
                }
                __syncthreads();
                if (s_ok) {
                        const float st = s_step;
                        float*      xw = c.xu + (size_t)b * traj;
                        for (int i = tid; i < traj; i += T) xw[i] = fmaf(st, dz[i], xw[i]);
                }
        }
}

// =====================================================================================================
// k_sim_forward: thread per solve — simForwardBatchedKernel / sim_step (sim.cuh:16-49, integrator.cuh:191-209)
// =====================================================================================================
template<class P>
__global__ void __launch_bounds__(64) k_sim_forward(int B, float* xkp1, const float* xk, const float* uk, const float* fext, float dt, int model_slot)
{
        constexpr int NQ = P::NQ, NX = 2 * NQ;
        const int     b = blockIdx.x * blockDim.x + threadIdx.x;
        if (b >= B) return;
        float x[NX], u[NQ], fe[6], qn[NQ], qdn[NQ];
        sfor<0, NX>([&](auto ic) { x[ic] = xk[ic]; });
        sfor<0, NQ>([&](auto ic) { u[ic] = uk[ic]; });
        sfor<0, 6>([&](auto ic) { fe[ic] = fext[6 * b + ic]; });
        make_items_slot<P>(model_slot).sim_step(x, u, fe, dt, qn, qdn);
        sfor<0, NQ>([&](auto ic) {
                xkp1[(size_t)b * NX + ic] = qn[ic];
                xkp1[(size_t)b * NX + NQ + ic] = qdn[ic];
        });
}

// =====================================================================================================
// k_ee_pos: thread per joint configuration -- the forward kinematics the cost uses (end_effector_pose_inner, iiwa14_grid.cuh:2596), exposed for
// the Python wrapper's ee_pos (python/bsqp/interface.py:212-214, which calls pinocchio there)
// =====================================================================================================
template<class P>
__global__ void __launch_bounds__(64) k_ee_pos(int n, const float* q, float* ee, int model_slot)
{
        constexpr int NQ = P::NQ;
        const int     i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n) return;
        float qi[NQ], e[3];
        sfor<0, NQ>([&](auto ic) { qi[ic] = q[(size_t)i * NQ + ic]; });
        make_items_slot<P>(model_slot).ee_pos(qi, e);
        sfor<0, 3>([&](auto ic) { ee[(size_t)i * 3 + ic] = e[ic]; });
}

}  // namespace gato
