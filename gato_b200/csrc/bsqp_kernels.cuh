// CUDA kernels of the B200-native BSQP solve path (sm_100a).  One SQP iteration is four launches:
//
//   k_kkt       thread per work item          linearise dynamics + quadraticise cost      -> A,B,c,Q,R,q,r (HBM/L2)
//               (cost block of a knot / d-dq half / d-dqd half of a linearisation, one kind per warp)
//   k_schur     warp   per pair of knots      in-place Gauss-Jordan inverses (two matrices per pass), phi/theta/gamma
//                                             -> S, diag(P^-1), gamma, Q^-1, R^-1
//   k_pcg       CTA    per solve              each thread keeps its rows of S and P^-1 in registers; off-diagonal P^-1 blocks,
//                                             PCG, primal step dz, device-side convergence bookkeeping
//               (k_pcg_stream for horizons whose system does not fit the register file)
//   k_merit_ls  CTA    per solve              8 x N forward-dynamics merit evaluations (thread per (alpha, knot)),
//                                             deterministic knot-ordered sum, line search, trajectory/rho update
//
// replacing setupKKTSystemBatchedKernel (setup_kkt.cuh:15), formSchurSystemBatchedKernel1/2 (schur_linsys.cuh:14,214),
// solvePCGBatchedKernel (pcg.cuh:14), computeDzBatchedKernel (schur_linsys.cuh:316), computeMeritBatchedKernel
// (merit.cuh:17), lineSearchAndUpdateBatchedKernel (line_search.cuh:13) and the host bookkeeping of
// BSQP::solve (bsqp.cuh:133-176).  No host synchronisation happens inside a solve: the "enough solves converged"
// early exit (bsqp.cuh:165) is evaluated on the device from a per-iteration counter that every later kernel reads.
//
// Arithmetic is bit-identical to the CPU oracle (see rbd.cuh): same expression trees, explicit fmaf, the
// reference's reduction trees (linalg.cuh:175-221, 291-327) rebuilt per thread.
#pragma once
#include <cuda_runtime.h>

#include "items.cuh"

namespace gato {

#ifndef GATO_KKT_MIN_BLOCKS
#define GATO_KKT_MIN_BLOCKS 1
#endif
#ifndef GATO_MERIT_MIN_BLOCKS
#define GATO_MERIT_MIN_BLOCKS 2
#endif
constexpr int   kNumAlphas = 8;                                                         // settings.h:15
constexpr float kRhoInit = 1e-3f, kRhoFactor = 1.2f, kRhoMin = 1e-8f, kRhoMax = 10.0f;  // settings.h:18-21
constexpr int   kPcgRefThreads = 1024;  // settings.h:25 — fixes the shape of the reference's dot-product tree

struct Ctx {
        int   N, B, it, max_pcg, adapt, flags;
        float dt, thresh;
        Costs cs;
        float*       xu;
        const float *xs, *ref, *fext;
        float *      Q, *R, *q, *r, *A, *Bm, *c, *Qinv, *Rinv;  // KKT blocks, reference layout [b][k][elements]
        float *      S, *Pinv, *gamma, *lambda, *dz;            // Schur system, dual, primal step
        float*       Pmain;  // main (diagonal) blocks of P^-1 as k_schur produces them, packed [b][k][nx x nx row-major]; k_pcg builds the rest
        float *      rho, *drho, *merit, *merit_cur, *step;
        const float *mu, *pcg_tol;
        int*         conv;        // [B] "PCG performed 0 iterations" flags (bsqp.cuh:153)
        unsigned*    num_solved;  // [max_sqp_iters] #flagged solves after the PCG of iteration i
        int*         pcg_log;     // [max_sqp_iters][B]
        float *      ls_merit_log, *ls_step_log;  // [max_sqp_iters][B]
};

enum : int { F_K2 = 1, F_PCG = 2, F_DZ = 4, F_WRITE_P = 8, F_MERIT = 16, F_LS = 32, F_BOOK = 64, F_CHECK_STOP = 128, F_ZERO_DZ = 256 };

// true when an iteration j < upto already satisfied the early-exit test of bsqp.cuh:165
__device__ __forceinline__ bool stopped_before(const Ctx& c, int upto)
{
        if (!(c.flags & F_CHECK_STOP)) return false;
        bool s = false;
        for (int j = 0; j < upto; j++) s |= ((float)c.num_solved[j] >= c.thresh);
        return s;
}

// =====================================================================================================
// k_kkt: one thread per work item, three kinds of items in separate warps (blockIdx.y = kind) so that no warp diverges:
//   kind 0  cost blocks of knot k = 0..N-1 (Q,q,R,r); knot N-1 is the "terminal" item: Q_{N-1}, q_{N-1} evaluated at
//           x_{N-2} against ref_{N-1} (setup_kkt.cuh:83-100) and c_0 = x_0 - x_s
//   kind 1  linearised dynamics of knot k = 0..N-2, d/dq half: columns 0..nq-1 of A_k and the defect c_{k+1}
//   kind 2  d/dqd half: columns nq..nx-1 of A_k and B_k
// (The two dynamics halves repeat the M^-1 / RNEA prologue; splitting doubles the parallelism of what is a latency-bound
// kernel at batch 512.)  Results are transposed through shared memory so that HBM/L2 stores are coalesced per knot block.
// =====================================================================================================
template<class P>
__global__ void __launch_bounds__(32, GATO_KKT_MIN_BLOCKS) k_kkt(Ctx c)
{
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ, ST = 33;
        // staged floats per item: kind 0: Q (its nq x nq block and the lower diagonal; everything else in Q is a structural zero) | q |
        // R (diagonal) | r | c0;  kind 1: half of A | c;  kind 2: half of A | B  -- the largest.  26 KB per warp keeps 8 warps per SM.
        constexpr int ROWS = NX * NQ + NX * NU;
        static_assert(ROWS >= NQ * NQ + NQ + NX + NU + NU + NX, "kind 0 fits");
        if (stopped_before(c, c.it)) return;
        __shared__ float stage[ROWS * ST];
        const int        kind = blockIdx.y;
        const int        lane = threadIdx.x;
        const int        item0 = blockIdx.x * 32;
        const int        per = (kind == 0) ? c.N : c.N - 1;  // items per solve
        const int        total = c.B * per;
        if (item0 >= total) return;
        const int  item = item0 + lane;
        const bool valid = item < total;
        const int  b = valid ? item / per : 0, k = valid ? item % per : 0;
        const bool term = (kind == 0) && (k == c.N - 1);
        const int  traj = (NX + NU) * c.N - NU;
        const int  ks = term ? k - 1 : k;  // knot whose (x,u) this item evaluates
        float      xux[2 * NX + NU];
        {
                const float* src = c.xu + (size_t)b * traj + (size_t)ks * (NX + NU);
                sfor<0, 2 * NX + NU>([&](auto ic) { xux[ic] = src[ic]; });
        }
        // write staged rows [row0, row0+count) of every selected item to dst[(b*N + knot + koff)*stride + off + e]
        // rowbase[i] = global knot index (b * N + k) of the warp's i-th item, bit 30 set for the terminal item: written once, so that the
        // flushes below need no integer divisions
        __shared__ int rowbase[32];
        rowbase[lane] = (b * c.N + k) | (term ? (1 << 30) : 0);
        __syncwarp();
        // write staged rows [row0, row0+COUNT) of every selected item to dst[(b*N + knot + koff)*stride + off + e]; the (item, element) pairs are
        // flattened over the lanes so that every store instruction is full and consecutive lanes write consecutive addresses
        auto flush = [&](auto count_c, float* dst, int row0, int stride, int off, int koff, int which /*0 non-terminal, 1 terminal, 2 all*/) {
                constexpr int COUNT = decltype(count_c)::value;
                const int     nvalid = min(32, total - item0);
                for (int f = lane; f < nvalid * COUNT; f += 32) {
                        const int  i = f / COUNT, e = f - i * COUNT;
                        const int  rb = rowbase[i];
                        const bool ti = (rb >> 30) & 1;
                        if ((which == 0 && ti) || (which == 1 && !ti)) continue;
                        dst[((size_t)(rb & ~(1 << 30)) + koff) * stride + off + e] = stage[(row0 + e) * ST + i];
                }
        };
        if (kind == 0) {
                float ref3[3];
                sfor<0, 3>([&](auto ic) { ref3[ic] = c.ref[(size_t)b * 6 * c.N + 6 * k + ic]; });
                constexpr int rQ = 0, rQd = NQ * NQ, rq = rQd + NQ, rR = rq + NX, rr = rR + NU, rc0 = rr + NU;
                // Q = [[h h^T w + barrier terms, 0], [0, diag]], R = diag (plant cost Hessians, iiwa14_plant.cuh:400-450): only those entries
                // are staged; the indices are compile-time constants after inlining, so the stores of structural zeros fold away
                Items<P>::template cost_grad_hess<true>(
                    xux, ref3, c.cs,
                    [&](int e, float v) {
                            const int i = e / NX, j = e % NX;
                            if (i < NQ && j < NQ)
                                    stage[(rQ + i * NQ + j) * ST + lane] = v;
                            else if (i == j)
                                    stage[(rQd + i - NQ) * ST + lane] = v;
                    },
                    [&](int e, float v) { stage[(rq + e) * ST + lane] = v; },
                    [&](int e, float v) {
                            if (e / NU == e % NU) stage[(rR + e / NU) * ST + lane] = v;
                    },
                    [&](int e, float v) { stage[(rr + e) * ST + lane] = v; }, term, term || Items<P>::pos_form_b_for(c.N));
                if (term && valid) {
                        const float* x0 = c.xu + (size_t)b * traj;
                        sfor<0, NX>([&](auto ic) { stage[(rc0 + ic) * ST + lane] = x0[ic] - c.xs[(size_t)b * NX + ic]; });
                }
                __syncwarp();
                {  // Q and R: expand the staged entries, zeros elsewhere ((item, element) pairs flattened over the lanes like flush)
                        const int nvalid = min(32, total - item0);
                        for (int f = lane; f < nvalid * NX * NX; f += 32) {
                                const int i = f / (NX * NX), e = f - i * (NX * NX), r_ = e / NX, c_ = e - r_ * NX;
                                float     v = 0.0f;
                                if (r_ < NQ && c_ < NQ)
                                        v = stage[(rQ + r_ * NQ + c_) * ST + i];
                                else if (r_ == c_)
                                        v = stage[(rQd + r_ - NQ) * ST + i];
                                c.Q[(size_t)(rowbase[i] & ~(1 << 30)) * NX * NX + e] = v;
                        }
                        for (int f = lane; f < nvalid * NU * NU; f += 32) {
                                const int i = f / (NU * NU), e = f - i * (NU * NU), r_ = e / NU, c_ = e - r_ * NU;
                                const int rb = rowbase[i];
                                if ((rb >> 30) & 1) continue;  // the terminal item has no R
                                c.R[(size_t)rb * NU * NU + e] = (r_ == c_) ? stage[(rR + r_) * ST + i] : 0.0f;
                        }
                }
                flush(std::integral_constant<int, NX>{}, c.q, rq, NX, 0, 0, 2);
                flush(std::integral_constant<int, NU>{}, c.r, rr, NU, 0, 0, 0);
                flush(std::integral_constant<int, NX>{}, c.c, rc0, NX, 0, -(c.N - 1), 1);
        } else {
                float fext[6];
                sfor<0, 6>([&](auto ic) { fext[ic] = c.fext[6 * b + ic]; });
                constexpr int rA = 0, rX = NX * NQ;  // half of A (NX*NQ contiguous floats), then c (kind 1) or B (kind 2)
                if (kind == 1) {
                        Items<P>::template linearize_half_rolled<0>(
                            xux, fext, c.dt, [&](int e, float v) { stage[(rA + e) * ST + lane] = v; }, [&](int, float) {}, [&](int e, float v) { stage[(rX + e) * ST + lane] = v; });
                        __syncwarp();
                        flush(std::integral_constant<int, NX * NQ>{}, c.A, rA, NX * NX, 0, 0, 2);
                        flush(std::integral_constant<int, NX>{}, c.c, rX, NX, 0, 1, 2);
                } else {
                        Items<P>::template linearize_half_rolled<1>(
                            xux, fext, c.dt, [&](int e, float v) { stage[(rA + e - NX * NQ) * ST + lane] = v; }, [&](int e, float v) { stage[(rX + e) * ST + lane] = v; }, [&](int, float) {});
                        __syncwarp();
                        flush(std::integral_constant<int, NX * NQ>{}, c.A, rA, NX * NX, NX * NQ, 0, 2);
                        flush(std::integral_constant<int, NX * NU>{}, c.Bm, rX, NX * NU, 0, 0, 2);
                }
        }
}

// =====================================================================================================
// k_kkt_fine: the same work cut finer, for small batches (the MPC regime) where k_kkt's three kinds leave most of the GPU idle and the
// time of a launch is the latency of one thread's instruction stream: 2 + 2 nq kinds (blockIdx.y) --
//   kind 0       cost blocks (as k_kkt)                     kind 1            B_k and the defect c_{k+1}
//   kind 2+j     column j of A_k (d/dq_j)                   kind 2+nq+j       column nq+j of A_k (d/dqd_j)
// every dynamics kind repeats the prologue; a thread then runs about 40 % of the instructions of a k_kkt thread.
// =====================================================================================================
template<class P>
__global__ void __launch_bounds__(32) k_kkt_fine(Ctx c)
{
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ, ST = 33;
        constexpr int ROWS = NX * NU + NX;  // kind 1 stages the most: B | c
        static_assert(ROWS >= NQ * NQ + NQ + NX + NU + NU + NX, "kind 0 fits");
        if (stopped_before(c, c.it)) return;
        __shared__ float stage[ROWS * ST];
        const int        kind = blockIdx.y;
        const int        lane = threadIdx.x;
        const int        item0 = blockIdx.x * 32;
        const int        per = (kind == 0) ? c.N : c.N - 1;
        const int        total = c.B * per;
        if (item0 >= total) return;
        const int  item = item0 + lane;
        const bool valid = item < total;
        const int  b = valid ? item / per : 0, k = valid ? item % per : 0;
        const bool term = (kind == 0) && (k == c.N - 1);
        const int  traj = (NX + NU) * c.N - NU;
        const int  ks = term ? k - 1 : k;
        float      xux[2 * NX + NU];
        {
                const float* src = c.xu + (size_t)b * traj + (size_t)ks * (NX + NU);
                sfor<0, 2 * NX + NU>([&](auto ic) { xux[ic] = src[ic]; });
        }
        // rowbase[i] = global knot index (b * N + k) of the warp's i-th item, bit 30 set for the terminal item: written once, so that the
        // flushes below need no integer divisions
        __shared__ int rowbase[32];
        rowbase[lane] = (b * c.N + k) | (term ? (1 << 30) : 0);
        __syncwarp();
        // write staged rows [row0, row0+COUNT) of every selected item to dst[(b*N + knot + koff)*stride + off + e]; the (item, element) pairs are
        // flattened over the lanes so that every store instruction is full and consecutive lanes write consecutive addresses
        auto flush = [&](auto count_c, float* dst, int row0, int stride, int off, int koff, int which /*0 non-terminal, 1 terminal, 2 all*/) {
                constexpr int COUNT = decltype(count_c)::value;
                const int     nvalid = min(32, total - item0);
                for (int f = lane; f < nvalid * COUNT; f += 32) {
                        const int  i = f / COUNT, e = f - i * COUNT;
                        const int  rb = rowbase[i];
                        const bool ti = (rb >> 30) & 1;
                        if ((which == 0 && ti) || (which == 1 && !ti)) continue;
                        dst[((size_t)(rb & ~(1 << 30)) + koff) * stride + off + e] = stage[(row0 + e) * ST + i];
                }
        };
        if (kind == 0) {
                float ref3[3];
                sfor<0, 3>([&](auto ic) { ref3[ic] = c.ref[(size_t)b * 6 * c.N + 6 * k + ic]; });
                constexpr int rQ = 0, rQd = NQ * NQ, rq = rQd + NQ, rR = rq + NX, rr = rR + NU, rc0 = rr + NU;
                Items<P>::template cost_grad_hess<true>(
                    xux, ref3, c.cs,
                    [&](int e, float v) {
                            const int i = e / NX, j = e % NX;
                            if (i < NQ && j < NQ)
                                    stage[(rQ + i * NQ + j) * ST + lane] = v;
                            else if (i == j)
                                    stage[(rQd + i - NQ) * ST + lane] = v;
                    },
                    [&](int e, float v) { stage[(rq + e) * ST + lane] = v; },
                    [&](int e, float v) {
                            if (e / NU == e % NU) stage[(rR + e / NU) * ST + lane] = v;
                    },
                    [&](int e, float v) { stage[(rr + e) * ST + lane] = v; }, term, term || Items<P>::pos_form_b_for(c.N));
                if (term && valid) {
                        const float* x0 = c.xu + (size_t)b * traj;
                        sfor<0, NX>([&](auto ic) { stage[(rc0 + ic) * ST + lane] = x0[ic] - c.xs[(size_t)b * NX + ic]; });
                }
                __syncwarp();
                {  // Q and R: expand the staged entries, zeros elsewhere ((item, element) pairs flattened over the lanes like flush)
                        const int nvalid = min(32, total - item0);
                        for (int f = lane; f < nvalid * NX * NX; f += 32) {
                                const int i = f / (NX * NX), e = f - i * (NX * NX), r_ = e / NX, c_ = e - r_ * NX;
                                float     v = 0.0f;
                                if (r_ < NQ && c_ < NQ)
                                        v = stage[(rQ + r_ * NQ + c_) * ST + i];
                                else if (r_ == c_)
                                        v = stage[(rQd + r_ - NQ) * ST + i];
                                c.Q[(size_t)(rowbase[i] & ~(1 << 30)) * NX * NX + e] = v;
                        }
                        for (int f = lane; f < nvalid * NU * NU; f += 32) {
                                const int i = f / (NU * NU), e = f - i * (NU * NU), r_ = e / NU, c_ = e - r_ * NU;
                                const int rb = rowbase[i];
                                if ((rb >> 30) & 1) continue;  // the terminal item has no R
                                c.R[(size_t)rb * NU * NU + e] = (r_ == c_) ? stage[(rR + r_) * ST + i] : 0.0f;
                        }
                }
                flush(std::integral_constant<int, NX>{}, c.q, rq, NX, 0, 0, 2);
                flush(std::integral_constant<int, NU>{}, c.r, rr, NU, 0, 0, 0);
                flush(std::integral_constant<int, NX>{}, c.c, rc0, NX, 0, -(c.N - 1), 1);
                return;
        }
        float fext[6];
        sfor<0, 6>([&](auto ic) { fext[ic] = c.fext[6 * b + ic]; });
        typename Rbd<P>::DynState st;
        Rbd<P>::dyn_prologue(xux, xux + NQ, xux + NX, fext, st);
        if (kind == 1) {
                constexpr int rB = 0, rc = NX * NU;
                Items<P>::linearize_base(st, xux, c.dt, [&](int e, float v) { stage[(rB + e) * ST + lane] = v; }, [&](int e, float v) { stage[(rc + e) * ST + lane] = v; });
                __syncwarp();
                flush(std::integral_constant<int, NX * NU>{}, c.Bm, rB, NX * NU, 0, 0, 2);
                flush(std::integral_constant<int, NX>{}, c.c, rc, NX, 0, 1, 2);
                return;
        }
        const int col = kind - 2;  // column of A
        sfor<0, NX>([&](auto cc) {
                constexpr int cidx = cc;
                if (col == cidx) Items<P>::template linearize_column<cidx / NQ, cidx % NQ>(st, xux + NQ, c.dt, [&](int e, float v) { stage[(e - cidx * NX) * ST + lane] = v; });
        });
        __syncwarp();
        flush(std::integral_constant<int, NX>{}, c.A, 0, NX * NX, col * NX, 0, 2);
}

#include "bsqp_linalg_kernels.cuh"  // k_schur
#include "bsqp_pcg_kernels.cuh"     // k_pcg, k_pcg_stream

// =====================================================================================================
// k_merit_ls: one CTA per solve, thread per (alpha, knot) — computeMeritBatchedKernel + lineSearchAndUpdateBatchedKernel
// NA = 8: merit at z + 2^-a dz for a = 0..7, then the line search.  NA = 1: merit at z (initial / final merit).
// =====================================================================================================
// SPLIT (small batches, where a launch lasts as long as one thread's instruction stream): two threads per (alpha, knot) -- one evaluates the
// forward dynamics and the defect, the other the tracking cost -- combined as fmaf(mu, defect, cost) exactly like the single-thread version.
template<class P, int NA, bool SPLIT = false>
__global__ void __launch_bounds__(SPLIT ? 512 : (NA == 1 ? 128 : 256), (NA == 1 || SPLIT) ? 1 : GATO_MERIT_MIN_BLOCKS) k_merit_ls(Ctx c)
{
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ;
        if (NA > 1 && stopped_before(c, c.it + 1)) return;  // the iteration that meets the test skips merit + line search (bsqp.cuh:165)
        extern __shared__ float smf[];                       // [NA][N] per-knot merits, NA sums, (SPLIT: [NA][N] cost halves)
        const int               N = c.N, b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
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
                                        m = half == 0 ? Items<P>::merit_mid_cons(xux, fext, c.dt) : Items<P>::template tracking_cost<false>(xux, ref3, c.cs);
                                else
                                        m = Items<P>::merit_mid(xux, ref3, mu, fext, c.dt, c.cs);
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
                                        m = half == 0 ? Items<P>::merit_last_cons(e0) : Items<P>::template tracking_cost<true>(xux, ref3, c.cs);
                                else
                                        m = Items<P>::merit_last(xux, ref3, mu, e0, c.cs);
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
__global__ void __launch_bounds__(64) k_sim_forward(int B, float* xkp1, const float* xk, const float* uk, const float* fext, float dt)
{
        constexpr int NQ = P::NQ, NX = 2 * NQ;
        const int     b = blockIdx.x * blockDim.x + threadIdx.x;
        if (b >= B) return;
        float x[NX], u[NQ], fe[6], qdd[NQ], qn[NQ], qdn[NQ];
        sfor<0, NX>([&](auto ic) { x[ic] = xk[ic]; });
        sfor<0, NQ>([&](auto ic) { u[ic] = uk[ic]; });
        sfor<0, 6>([&](auto ic) { fe[ic] = fext[6 * b + ic]; });
        Rbd<P>::forward_dynamics(x, x + NQ, u, fe, qdd);
        Rbd<P>::integrate(x, x + NQ, qdd, dt, qn, qdn);
        sfor<0, NQ>([&](auto ic) {
                xkp1[(size_t)b * NX + ic] = qn[ic];
                xkp1[(size_t)b * NX + NQ + ic] = qdn[ic];
        });
}

}  // namespace gato
