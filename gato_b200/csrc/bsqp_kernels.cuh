// CUDA kernels of the B200-native BSQP solve path (sm_100a).  One SQP iteration is four launches:
//
//   k_kkt       thread per (solve, knot)      linearise dynamics + quadraticise cost      -> A,B,c,Q,R,q,r (HBM/L2)
//   k_schur     warp   per (solve, knot)      Gauss-Jordan inverses, phi/theta/gamma       -> S, diag(P^-1), gamma, Q^-1, R^-1
//   k_pcg       CTA    per solve              S,P^-1 resident in shared memory; off-diagonal P^-1 blocks, PCG,
//                                             primal step dz, device-side convergence bookkeeping
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
// k_kkt: one thread per (solve, knot).  Knot N-1 is the "terminal" item: Q_{N-1}, q_{N-1} evaluated at
// x_{N-2} against ref_{N-1} (setup_kkt.cuh:83-100) and c_0 = x_0 - x_s; it runs the same cost code as the
// other lanes, so a warp (= one solve when N = 32) does not diverge in the cost part.
// Results are transposed through shared memory so that HBM/L2 stores are coalesced per knot block.
// =====================================================================================================
template<class P>
__global__ void __launch_bounds__(32) k_kkt(Ctx c)
{
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ, ST = 33;
        constexpr int ROWS = NX * NX + NX * NU + NX;  // A | B | c  (the cost pass uses fewer rows)
        if (stopped_before(c, c.it)) return;
        __shared__ float stage[ROWS * ST];
        const int        lane = threadIdx.x;
        const int        item0 = blockIdx.x * 32;
        const int        total = c.B * c.N;
        const int        item = item0 + lane;
        const bool       valid = item < total;
        const int        b = valid ? item / c.N : 0, k = valid ? item % c.N : 0;
        const bool       term = (k == c.N - 1);
        const int        traj = (NX + NU) * c.N - NU;
        const int        ks = term ? k - 1 : k;  // knot whose (x,u) this item evaluates
        float            xux[2 * NX + NU];
        {
                const float* src = c.xu + (size_t)b * traj + (size_t)ks * (NX + NU);
                sfor<0, 2 * NX + NU>([&](auto ic) { xux[ic] = src[ic]; });
        }
        float fext[6], ref3[3];
        sfor<0, 6>([&](auto ic) { fext[ic] = c.fext[6 * b + ic]; });
        sfor<0, 3>([&](auto ic) { ref3[ic] = c.ref[(size_t)b * 6 * c.N + 6 * k + ic]; });

        // write staged rows [row0, row0+count) of every selected item to dst[(b*N + knot + koff)*count + e]
        auto flush = [&](float* dst, int row0, int count, int koff, int which /*0 mid items, 1 terminal items, 2 all*/) {
                const int nvalid = min(32, total - item0);
                for (int i = 0; i < nvalid; i++) {
                        const int  it_ = item0 + i, bi = it_ / c.N, ki = it_ % c.N;
                        const bool ti = (ki == c.N - 1);
                        if ((which == 0 && ti) || (which == 1 && !ti)) continue;
                        float* d = dst + ((size_t)bi * c.N + ki + koff) * count;
                        for (int e = lane; e < count; e += 32) d[e] = stage[(row0 + e) * ST + i];
                }
        };
        constexpr int rQ = 0, rq = NX * NX, rR = rq + NX, rr = rR + NU * NU, rc0 = rr + NU;
        static_assert(rc0 + NX <= ROWS, "stage too small");

        // ---- cost blocks ----
        Items<P>::template cost_grad_hess<true>(
            xux, ref3, c.cs, [&](int e, float v) { stage[(rQ + e) * ST + lane] = v; }, [&](int e, float v) { stage[(rq + e) * ST + lane] = v; },
            [&](int e, float v) { stage[(rR + e) * ST + lane] = v; }, [&](int e, float v) { stage[(rr + e) * ST + lane] = v; });
        if (term && valid) {
                const float* x0 = c.xu + (size_t)b * traj;
                sfor<0, NX>([&](auto ic) { stage[(rc0 + ic) * ST + lane] = x0[ic] - c.xs[(size_t)b * NX + ic]; });
        }
        __syncwarp();
        flush(c.Q, rQ, NX * NX, 0, 2);
        flush(c.q, rq, NX, 0, 2);
        flush(c.R, rR, NU * NU, 0, 0);
        flush(c.r, rr, NU, 0, 0);
        flush(c.c, rc0, NX, -(c.N - 1), 1);
        __syncwarp();

        // ---- linearised dynamics (the terminal lane idles) ----
        constexpr int rA = 0, rB = NX * NX, rc = rB + NX * NU;
        if (!term && valid) {
                Items<P>::linearize(
                    xux, fext, c.dt, [&](int e, float v) { stage[(rA + e) * ST + lane] = v; }, [&](int e, float v) { stage[(rB + e) * ST + lane] = v; },
                    [&](int e, float v) { stage[(rc + e) * ST + lane] = v; });
        }
        __syncwarp();
        flush(c.A, rA, NX * NX, 0, 0);
        flush(c.Bm, rB, NX * NU, 0, 0);
        flush(c.c, rc, NX, 1, 0);
}

// =====================================================================================================
// k_schur: one warp per (solve, knot)  — formSchurSystemBatchedKernel1 (schur_linsys.cuh:14-211)
// Inverses go to separate Qinv/Rinv buffers (the reference overwrites Q in place while neighbouring blocks
// still read it — the cross-block race noted in SURVEY.md §5 is not reproduced).
// =====================================================================================================
template<int NX, int NU>
struct SchurSmem {
        float M1[2 * NX * NX], M2[2 * NX * NX], M3[2 * NU * NU];  // [V | I] augmented, col-major dim x 2dim
        float A[NX * NX], Bm[NX * NU], phi[NX * NX], BR[NX * NU];
        float qk[NX], qk1[NX], rk[NU], g[NX], f1[NX], f2[NX], f3[NU];
};

// one pivot step of the reference's lock-step Gauss-Jordan (linalg.cuh:364-400, 457-519), split in three phases:
//   factors f[row] = col[row]/piv (or col[row]*(1/piv)),  M[row][col] = fmaf(-f[row], M[piv][col], M[row][col]),  pivot row /= piv.
// Lanes [l0, l0+nl) of the warp take part.
template<int DIM, bool RCP>
__device__ __forceinline__ void gj_factors(const float* M, int p, float* f, int lane, int l0)
{
        const int r = lane - l0;
        if (r >= 0 && r < DIM) {
                const float cv = M[p * DIM + r], pv = M[p * DIM + p];
                f[r] = RCP ? (cv * (1.0f / pv)) : (cv / pv);
        }
}
template<int DIM>
__device__ __forceinline__ void gj_update(float* M, int p, const float* f, int lane, int nl, int l0)
{
        if (lane < l0 || lane >= l0 + nl) return;
        for (int ind = lane - l0; ind < DIM * (DIM + 1); ind += nl) {
                const int row = ind % DIM, col = ind / DIM;
                if (row == p) continue;
                const float rowv = M[(p + col) * DIM + p];
                M[(p + col) * DIM + row] = fmaf(-f[row], rowv, M[(p + col) * DIM + row]);
        }
}
template<int DIM, bool RCP>
__device__ __forceinline__ void gj_pivot_row(float* M, int p, float pv, int lane, int nl, int l0)
{
        if (lane < l0 || lane >= l0 + nl) return;
        for (int col = lane - l0; col <= DIM; col += nl) {
                float& e = M[(p + col) * DIM + p];
                e = RCP ? (e * (1.0f / pv)) : (e / pv);
        }
}

template<class P>
__global__ void __launch_bounds__(128) k_schur(Ctx c)
{
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ, NX2 = NX * NX, NU2 = NU * NU, W = 3 * NX;
        if (stopped_before(c, c.it)) return;
        extern __shared__ float smem_raw[];
        using SM = SchurSmem<NX, NU>;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        SM&       s = reinterpret_cast<SM*>(smem_raw)[warp];
        const int item = blockIdx.x * (blockDim.x >> 5) + warp;
        if (item >= c.B * c.N) return;
        const int    b = item / c.N, k = item % c.N;
        const float  rho = c.rho[b];
        const size_t kb = (size_t)b * c.N;
        float*       Sb = c.S + kb * 3 * NX2;
        float*       Pb = c.Pinv + kb * 3 * NX2;
        float*       gam = c.gamma + (size_t)b * (c.N + 2) * NX;

        if (k < c.N - 1) {
                // rho is added as "v + rho" exactly like the reference; rho == 0 must still add (x + 0 = x)
                for (int i = lane; i < NX2; i += 32) {
                        const int r = i % NX, cc = i / NX;
                        float     v1 = c.Q[(kb + k) * NX2 + i], v2 = c.Q[(kb + k + 1) * NX2 + i];
                        if (r == cc && r < NX / 2) {
                                v1 = v1 + rho;
                                v2 = v2 + rho;
                        }
                        s.M1[i] = v1, s.M2[i] = v2;
                        s.M1[NX2 + i] = s.M2[NX2 + i] = (r == cc) ? 1.0f : 0.0f;
                        s.A[i] = c.A[(kb + k) * NX2 + i];
                }
                for (int i = lane; i < NU2; i += 32) {
                        s.M3[i] = c.R[(kb + k) * NU2 + i];
                        s.M3[NU2 + i] = (i % NU == i / NU) ? 1.0f : 0.0f;
                }
                for (int i = lane; i < NX * NU; i += 32) s.Bm[i] = c.Bm[(kb + k) * NX * NU + i];
                if (lane < NX) {
                        s.qk[lane] = c.q[(kb + k) * NX + lane];
                        s.qk1[lane] = c.q[(kb + k + 1) * NX + lane];
                        s.g[lane] = -1.0f * c.c[(kb + k + 1) * NX + lane];
                }
                if (lane < NU) s.rk[lane] = c.r[(kb + k) * NU + lane];
                __syncwarp();
                // ---- three inverses in lock step (division form) ----
                for (int p = 0; p < NX; p++) {
                        gj_factors<NX, false>(s.M1, p, s.f1, lane, 0);
                        gj_factors<NX, false>(s.M2, p, s.f2, lane, 16);
                        const float pv1 = s.M1[p * NX + p], pv2 = s.M2[p * NX + p];
                        float       pv3 = 0.0f;
                        if (p < NU) pv3 = s.M3[p * NU + p];
                        __syncwarp();
                        if (p < NU) gj_factors<NU, false>(s.M3, p, s.f3, lane, 0);
                        __syncwarp();
                        gj_update<NX>(s.M1, p, s.f1, lane, 32, 0);
                        gj_update<NX>(s.M2, p, s.f2, lane, 32, 0);
                        if (p < NU) gj_update<NU>(s.M3, p, s.f3, lane, 32, 0);
                        __syncwarp();
                        gj_pivot_row<NX, false>(s.M1, p, pv1, lane, 16, 0);
                        gj_pivot_row<NX, false>(s.M2, p, pv2, lane, 16, 16);
                        __syncwarp();
                        if (p < NU) gj_pivot_row<NU, false>(s.M3, p, pv3, lane, 32, 0);
                        __syncwarp();
                }
                const float *Qi = s.M1 + NX2, *Q1i = s.M2 + NX2, *Ri = s.M3 + NU2;
                for (int i = lane; i < NX2; i += 32) {
                        c.Qinv[(kb + k) * NX2 + i] = Qi[i];
                        if (k == c.N - 2) c.Qinv[(kb + k + 1) * NX2 + i] = Q1i[i];
                }
                for (int i = lane; i < NU2; i += 32) c.Rinv[(kb + k) * NU2 + i] = Ri[i];
                // ---- phi = A Qinv ; BR = B Rinv ----
                for (int i = lane; i < NX2; i += 32) {
                        const int y = i % NX, x = i / NX;
                        float     sum = 0.0f;
#pragma unroll
                        for (int j = 0; j < NX; j++) sum = fmaf(s.A[j * NX + y], Qi[x * NX + j], sum);
                        s.phi[i] = sum;
                }
                for (int i = lane; i < NX * NU; i += 32) {
                        const int y = i % NX, x = i / NX;
                        float     sum = 0.0f;
#pragma unroll
                        for (int j = 0; j < NU; j++) sum = fmaf(s.Bm[j * NX + y], Ri[x * NU + j], sum);
                        s.BR[i] = sum;
                }
                __syncwarp();
                // ---- theta = Q1inv + phi A^T + BR B^T  (kept in M1's left half: Qk no longer needed) ----
                float* theta = s.M1;
                for (int i = lane; i < NX2; i += 32) {
                        const int y = i % NX, x = i / NX;
                        float     s1 = 0.0f, s2 = 0.0f;
#pragma unroll
                        for (int j = 0; j < NX; j++) s1 = fmaf(s.phi[j * NX + y], s.A[j * NX + x], s1);
#pragma unroll
                        for (int j = 0; j < NU; j++) s2 = fmaf(s.BR[j * NX + y], s.Bm[j * NX + x], s2);
                        theta[i] = (Q1i[i] + s1) + s2;
                }
                // ---- gamma_{k+1} ----
                if (lane < NX) {
                        const int y = lane;
                        float     s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
                        for (int j = 0; j < NX; j++) s1 = fmaf(Q1i[j * NX + y], s.qk1[j], s1);
#pragma unroll
                        for (int j = 0; j < NX; j++) s2 = fmaf(s.phi[j * NX + y], s.qk[j], s2);
#pragma unroll
                        for (int j = 0; j < NU; j++) s3 = fmaf(s.BR[j * NX + y], s.rk[j], s3);
                        float g = s.g[y] + s1;
                        g = g + (-s2);
                        g = g + (-s3);
                        gam[(k + 2) * NX + y] = -1.0f * g;
                }
                __syncwarp();
                // ---- S blocks (row-major nx x 3nx block rows) ----
                float* Sright = Sb + (size_t)k * 3 * NX2 + 2 * NX;
                float* Sleft = Sb + (size_t)(k + 1) * 3 * NX2;
                float* Smain = Sleft + NX;
                for (int i = lane; i < NX2; i += 32) {
                        const int x = i % NX, y = i / NX, off = y * W + x;
                        Sright[off] = s.phi[i];
                        Sleft[off] = s.phi[x * NX + y];
                        Smain[off] = -theta[x * NX + y];
                }
                __syncwarp();
                // ---- (theta + rho I~)^-1, reciprocal form (linalg.cuh:364-400) ----
                for (int i = lane; i < NX2; i += 32) {
                        const int r = i % NX, cc = i / NX;
                        if (r == cc && r < NX / 2) theta[i] = theta[i] + rho;
                        s.M1[NX2 + i] = (r == cc) ? 1.0f : 0.0f;
                }
                __syncwarp();
                for (int p = 0; p < NX; p++) {
                        gj_factors<NX, true>(s.M1, p, s.f1, lane, 0);
                        const float pv = s.M1[p * NX + p];
                        __syncwarp();
                        gj_update<NX>(s.M1, p, s.f1, lane, 32, 0);
                        __syncwarp();
                        gj_pivot_row<NX, true>(s.M1, p, pv, lane, 32, 0);
                        __syncwarp();
                }
                float* Pmain = Pb + (size_t)(k + 1) * 3 * NX2 + NX;
                for (int i = lane; i < NX2; i += 32) {
                        const int x = i % NX, y = i / NX;
                        Pmain[y * W + x] = -s.M1[NX2 + x * NX + y];
                }
        } else {
                // ---- last knot's block handles Q_0 (schur_linsys.cuh:166-210) ----
                for (int i = lane; i < NX2; i += 32) {
                        const int r = i % NX, cc = i / NX;
                        float     v = c.Q[kb * NX2 + i];
                        if (r == cc && r < NX / 2) v = v + rho;
                        s.M1[i] = v;
                        s.M1[NX2 + i] = (r == cc) ? 1.0f : 0.0f;
                }
                if (lane < NX) {
                        s.qk[lane] = c.q[kb * NX + lane];
                        s.g[lane] = c.c[kb * NX + lane];
                }
                __syncwarp();
                float* P0 = Pb + NX;
                for (int i = lane; i < NX2; i += 32) {
                        const int x = i % NX, y = i / NX;
                        P0[y * W + x] = -s.M1[x * NX + y];
                }
                __syncwarp();
                for (int p = 0; p < NX; p++) {
                        gj_factors<NX, true>(s.M1, p, s.f1, lane, 0);
                        const float pv = s.M1[p * NX + p];
                        __syncwarp();
                        gj_update<NX>(s.M1, p, s.f1, lane, 32, 0);
                        __syncwarp();
                        gj_pivot_row<NX, true>(s.M1, p, pv, lane, 32, 0);
                        __syncwarp();
                }
                float* S0 = Sb + NX;
                for (int i = lane; i < NX2; i += 32) {
                        const int x = i % NX, y = i / NX;
                        S0[y * W + x] = -s.M1[NX2 + x * NX + y];
                }
                if (lane < NX) {
                        const int y = lane;
                        float     s1 = 0.0f;
#pragma unroll
                        for (int j = 0; j < NX; j++) s1 = fmaf(s.M1[NX2 + j * NX + y], s.qk[j], s1);
                        gam[NX + y] = s.g[y] + (-s1);
                }
        }
}

// =====================================================================================================
// k_pcg: one CTA per solve — formSchurSystemBatchedKernel2 + solvePCGBatchedKernel + computeDzBatchedKernel +
// the host convergence bookkeeping of bsqp.cuh:142-163, with S and P^-1 resident in shared memory
// (the reference re-reads both from global memory in every PCG iteration, pcg.cuh:100,119).
// =====================================================================================================
__device__ __forceinline__ float warp_tree(float v)  // __shfl_down tree 16,8,4,2,1 -> lane 0 (linalg.cuh:215)
{
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = v + __shfl_down_sync(0xffffffffu, v, off);
        return v;
}

// Row r of a block-tridiagonal matvec with the reference's reduction tree: 32 lane partials (columns l and l+32),
// then the shuffle tree — evaluated by ONE thread on registers.
template<int NX, int WP>
__device__ __forceinline__ float btd_row(const float* __restrict__ Mrow, const float* __restrict__ vec)
{
        constexpr int W = 3 * NX;
        float         m[WP], v[WP];
        // vectorised shared-memory loads: matrix row is 16B aligned (padded to WP), vector window is 8B aligned
#pragma unroll
        for (int i = 0; i < WP / 4; i++) {
                const float4 t = reinterpret_cast<const float4*>(Mrow)[i];
                m[4 * i] = t.x, m[4 * i + 1] = t.y, m[4 * i + 2] = t.z, m[4 * i + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < W / 2; i++) {
                const float2 t = reinterpret_cast<const float2*>(vec)[i];
                v[2 * i] = t.x, v[2 * i + 1] = t.y;
        }
        float p[32];
#pragma unroll
        for (int l = 0; l < 32; l++) {
                float s = fmaf(m[l], v[l], 0.0f);
                if (l + 32 < W) s = fmaf(m[l + 32], v[l + 32], s);
                p[l] = s;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                for (int l = 0; l < off; l++) p[l] = p[l] + p[l + off];
        }
        return p[0];
}

template<class P>
__global__ void __launch_bounds__(512) k_pcg(Ctx c)
{
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ, NX2 = NX * NX, W = 3 * NX, WP = (W + 3) / 4 * 4;
        if (stopped_before(c, c.it)) return;
        extern __shared__ __align__(16) float sm[];
        const int                             N = c.N, b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
        const int                             nrows = N * NX, n = (N + 2) * NX;
        float*                                sS = sm;                   // [nrows][WP]
        float*                                sP = sS + (size_t)nrows * WP;  // [nrows][WP]
        float*                                vAp = sP + (size_t)nrows * WP;
        float *                               vx = vAp + n, *vr = vx + n, *vz = vr + n, *vp = vz + n;
        float*                                scratch = vp + n;          // 32 warp sums + 4 scalars
        float*                                dzbuf = scratch + 40;      // 64 floats per warp for the dz phase
        float*                                scr2 = dzbuf + 64 * (T >> 5);  // K2 scratch: (N-1) * NX2
        const size_t                          kb = (size_t)b * N;
        const float*                          gS = c.S + kb * 3 * NX2;
        float*                                gP = c.Pinv + kb * 3 * NX2;

        // ---- load S and P^-1 (diagonal blocks; off-diagonals are built below) ----
        for (int i = tid; i < nrows * W; i += T) {
                const int r = i / W, cc = i % W;
                sS[r * WP + cc] = gS[i];
                sP[r * WP + cc] = gP[i];
        }
        if (WP > W)
                for (int i = tid; i < nrows * (WP - W); i += T) {
                        const int r = i / (WP - W), cc = W + i % (WP - W);
                        sS[r * WP + cc] = 0.0f;
                        sP[r * WP + cc] = 0.0f;
                }
        for (int i = tid; i < 5 * n; i += T) vAp[i] = 0.0f;
        __syncthreads();

        if (c.flags & F_K2) {
                // left_{k+1} = -(Theta_k * (phi_k * Theta_{k-1})), right_k = left_{k+1}^T   (schur_linsys.cuh:227-259)
                for (int i = tid; i < (N - 1) * NX2; i += T) {
                        const int k = i / NX2, e = i % NX2, y = e % NX, x = e / NX;
                        // scr(y,x) = sum_j phi(y,j) * tkm1(j,x);  phi = S left of row k+1, tkm1 = stored P main of row k
                        const float* ph = sS + (size_t)(k + 1) * NX * WP;
                        const float* tk1 = sP + (size_t)k * NX * WP + NX;
                        float        sum = 0.0f;
#pragma unroll
                        for (int j = 0; j < NX; j++) sum = fmaf(ph[y * WP + j], tk1[j * WP + x], sum);
                        scr2[i] = sum;  // col-major (y + NX*x)
                }
                __syncthreads();
                for (int i = tid; i < (N - 1) * NX2; i += T) {
                        const int    k = i / NX2, e = i % NX2, y = e % NX, x = e / NX;
                        const float* tk = sP + (size_t)(k + 1) * NX * WP + NX;
                        const float* sc = scr2 + (size_t)k * NX2;
                        float        sum = 0.0f;
#pragma unroll
                        for (int j = 0; j < NX; j++) sum = fmaf(tk[y * WP + j], sc[j + NX * x], sum);
                        // out(y,x): left block of row k+1 at (y,x), right block of row k at (x,y)
                        sP[((size_t)(k + 1) * NX + y) * WP + x] = -sum;
                        sP[((size_t)k * NX + x) * WP + 2 * NX + y] = -sum;
                }
                __syncthreads();
                if (c.flags & F_WRITE_P)
                        for (int i = tid; i < nrows * W; i += T) gP[i] = sP[(i / W) * WP + i % W];
        }

        int iters = 0;
        if (c.flags & F_PCG) {
                const float* gam = c.gamma + (size_t)b * n;
                float*       lam = c.lambda + (size_t)b * n;
                const float  eps = c.pcg_tol[b];
                const float  abs_tol = 1e-6f;
                const bool   skip = c.conv[b] != 0;  // pcg.cuh:29-32
                const int    warp = tid >> 5, lane = tid & 31, nwarps = T >> 5;
                float*       s_res = scratch + 32;

                // block::dot with the reference's 1024-thread geometry (linalg.cuh:291-327): result in s_res[0]
                auto dot = [&](const float* a, const float* bb) {
                        for (int vw = warp; vw < 32; vw += nwarps) {
                                const int vt = vw * 32 + lane;
                                float     s = 0.0f;
                                for (int i = vt; i < n; i += kPcgRefThreads) s = fmaf(a[i], bb[i], s);
                                s = warp_tree(s);
                                if (lane == 0) scratch[vw] = s;
                        }
                        __syncthreads();
                        if (warp == 0) {
                                float s = warp_tree(scratch[lane]);
                                if (lane == 0) s_res[0] = s;
                        }
                        __syncthreads();
                };
                auto matvec = [&](const float* M, const float* v, float* out) {
                        for (int r = tid; r < nrows; r += T) {
                                const int br = r / NX;
                                out[NX + r] = btd_row<NX, WP>(M + (size_t)r * WP, v + br * NX);
                        }
                };
                if (!skip) {
                        for (int i = tid; i < n; i += T) vx[i] = lam[i];
                        __syncthreads();
                        matvec(sS, vx, vr);
                        __syncthreads();
                        for (int i = tid; i < n; i += T) vr[i] = gam[i] - vr[i];
                        __syncthreads();
                        matvec(sP, vr, vz);
                        __syncthreads();
                        for (int i = tid; i < n; i += T) vp[i] = vz[i];
                        dot(vr, vz);
                        float rho = s_res[0];
                        __syncthreads();
                        if (!(fabsf(rho) < abs_tol)) {
                                const float rho_init = fabsf(rho);
                                for (int itn = 0; itn < c.max_pcg; itn++) {
                                        iters++;
                                        matvec(sS, vp, vAp);
                                        __syncthreads();
                                        dot(vp, vAp);
                                        const float alpha = rho / s_res[0];
                                        for (int j = tid; j < n; j += T) {
                                                vx[j] = fmaf(alpha, vp[j], vx[j]);
                                                vr[j] = fmaf(-alpha, vAp[j], vr[j]);
                                        }
                                        __syncthreads();
                                        matvec(sP, vr, vz);
                                        __syncthreads();
                                        dot(vr, vz);
                                        const float rho_new = s_res[0];
                                        if (fabsf(rho_new) < fmaf(eps, rho_init, abs_tol)) break;
                                        const float beta = rho_new / rho;
                                        rho = rho_new;
                                        for (int j = tid; j < n; j += T) vp[j] = fmaf(beta, vp[j], vz[j]);
                                        __syncthreads();
                                }
                                __syncthreads();
                                for (int i = tid; i < n; i += T) lam[i] = vx[i];
                        }
                }
                if (tid == 0) {
                        if (c.pcg_log) c.pcg_log[(size_t)c.it * c.B + b] = iters;
                        if (c.flags & F_BOOK) {
                                // bsqp.cuh:153-163: a solve is flagged once PCG performs no iteration; count flagged solves
                                int cv = c.conv[b];
                                if (iters == 0) cv = 1;
                                c.conv[b] = cv;
                                if (cv) atomicAdd(&c.num_solved[c.it], 1u);
                        }
                }
                __syncthreads();
        }

        if (c.flags & F_DZ) {
                // dz_x,k = -Qinv_k (q_k - lambda_k + A_k^T lambda_{k+1}), dz_u,k = -Rinv_k (r_k + B_k^T lambda_{k+1})
                // residuals are stored back into q, r   (schur_linsys.cuh:331-430).  One warp per knot.
                const float* lam = c.lambda + (size_t)b * n;
                const int    warp = tid >> 5, lane = tid & 31, nwarps = T >> 5;
                float*       wbuf = dzbuf + warp * 64;
                const int    traj = (NX + NU) * N - NU;
                for (int k = warp; k < N; k += nwarps) {
                        const float* lk = lam + (k + 1) * NX;
                        const float* lk1 = lam + (k + 2) * NX;
                        __syncwarp();
                        if (lane < NX) {
                                float scr = 0.0f;
                                if (k < N - 1) {
                                        const float* Ak = c.A + (kb + k) * NX2;
                                        float        sum = 0.0f;
#pragma unroll
                                        for (int j = 0; j < NX; j++) sum = fmaf(lk1[j], Ak[lane * NX + j], sum);
                                        scr = -sum;
                                }
                                scr = scr + lk[lane];
                                const float res = c.q[(kb + k) * NX + lane] - scr;
                                wbuf[lane] = res;
                        } else if (lane >= 16 && lane < 16 + NU && k < N - 1) {
                                const int    x = lane - 16;
                                const float* Bk = c.Bm + (kb + k) * NX * NU;
                                float        sum = 0.0f;
#pragma unroll
                                for (int j = 0; j < NX; j++) sum = fmaf(lk1[j], Bk[x * NX + j], sum);
                                const float su = c.r[(kb + k) * NU + x] - (-sum);
                                wbuf[32 + x] = su;
                        }
                        __syncwarp();
                        if (lane < NX) {
                                const float* Qi = c.Qinv + (kb + k) * NX2;
                                float        sum = 0.0f;
#pragma unroll
                                for (int j = 0; j < NX; j++) sum = fmaf(Qi[j * NX + lane], wbuf[j], sum);
                                c.dz[(size_t)b * traj + (size_t)k * (NX + NU) + lane] = -1.0f * sum;
                                c.q[(kb + k) * NX + lane] = wbuf[lane];
                        } else if (lane >= 16 && lane < 16 + NU) {
                                const int x = lane - 16;
                                if (k < N - 1) {
                                        const float* Ri = c.Rinv + (kb + k) * NU * NU;
                                        float        sum = 0.0f;
#pragma unroll
                                        for (int j = 0; j < NU; j++) sum = fmaf(Ri[j * NU + x], wbuf[32 + j], sum);
                                        c.dz[(size_t)b * traj + (size_t)k * (NX + NU) + NX + x] = -1.0f * sum;
                                        c.r[(kb + k) * NU + x] = wbuf[32 + x];
                                } else {
                                        c.r[(kb + k) * NU + x] = 0.0f;
                                }
                        }
                }
        }
}

// =====================================================================================================
// k_merit_ls: one CTA per solve, thread per (alpha, knot) — computeMeritBatchedKernel + lineSearchAndUpdateBatchedKernel
// NA = 8: merit at z + 2^-a dz for a = 0..7, then the line search.  NA = 1: merit at z (initial / final merit).
// =====================================================================================================
template<class P, int NA>
__global__ void __launch_bounds__(NA == 1 ? 128 : 256) k_merit_ls(Ctx c)
{
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ;
        if (NA > 1 && stopped_before(c, c.it + 1)) return;  // the iteration that meets the test skips merit + line search (bsqp.cuh:165)
        extern __shared__ float smf[];                       // [NA][N] per-knot merits, then NA sums
        const int               N = c.N, b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
        const int               traj = (NX + NU) * N - NU;
        float*                  mk = smf;
        float*                  msum = smf + NA * N;
        const float*            xu = c.xu + (size_t)b * traj;
        const float*            dz = c.dz + (size_t)b * traj;
        if (c.flags & F_MERIT) {
                const float mu = c.mu[b];
                const bool  zero_dz = (c.flags & F_ZERO_DZ) != 0;
                float       fext[6];
                sfor<0, 6>([&](auto ic) { fext[ic] = c.fext[6 * b + ic]; });
                for (int w = tid; w < NA * N; w += T) {
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
                                m = Items<P>::merit_last(xux, ref3, mu, e0, c.cs);
                        }
                        mk[a * N + k] = m;
                }
                __syncthreads();
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
