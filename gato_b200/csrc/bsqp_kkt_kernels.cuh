// k_kkt / k_kkt_fine of the BSQP path (see bsqp_ctx.cuh for the kernel map)
#pragma once
#include "bsqp_ctx.cuh"
#include "items.cuh"

namespace gato {

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

}  // namespace gato
