// k_kkt / k_kkt_fine of the BSQP path (see bsqp_ctx.cuh for the kernel map)
#pragma once
#include "bsqp_ctx.cuh"
#include "items.cuh"
#include "rt_slots.cuh"

namespace gato {

// What both kernels share: a warp of 32 work items of one kind, one item per lane.  Decodes the items, loads their (x, u, x_next), stages results
// transposed in shared memory (ROWS rows of 33 floats: row = element, column = lane) and flushes them with (item, element) pairs flattened over
// the lanes, so that every store instruction is full and consecutive lanes write consecutive addresses; and the cost item (kind 0).
template<class P, int ROWS>
struct KktWarp {
        static constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ, ST = 33;
        static_assert(ROWS >= NQ * NQ + NQ + NX + NU + NU + NX, "the cost item fits the staging buffer");
        float* stage;    // [ROWS * ST] shared
        int*   rowbase;  // [32] shared: global knot index (b * N + k) of the warp's i-th item, bit 30 set for the terminal item (no divisions in the flushes)
        int    lane, item0, total, b, k, traj;
        bool   valid, term;
        float  xux[2 * NX + NU];

        // false: the warp has no item (beyond the grid's last partial block)
        __device__ __forceinline__ bool init(const Ctx& c, float* stage_, int* rowbase_, bool cost_kind)
        {
                stage = stage_, rowbase = rowbase_;
                lane = threadIdx.x, item0 = blockIdx.x * 32;
                const int per = cost_kind ? c.N : c.N - 1;  // items per solve
                total = c.B * per;
                if (item0 >= total) return false;
                const int item = item0 + lane;
                valid = item < total;
                b = valid ? item / per : 0, k = valid ? item % per : 0;
                term = cost_kind && (k == c.N - 1);
                traj = (NX + NU) * c.N - NU;
                const int    ks = term ? k - 1 : k;  // knot whose (x,u) this item evaluates
                const float* src = c.xu + (size_t)b * traj + (size_t)ks * (NX + NU);
                sfor<0, 2 * NX + NU>([&](auto ic) { xux[ic] = src[ic]; });
                rowbase[lane] = (b * c.N + k) | (term ? (1 << 30) : 0);
                __syncwarp();
                return true;
        }
        __device__ __forceinline__ void put(int row, float v) const { stage[row * ST + lane] = v; }
        // staged rows [row0, row0+COUNT) of every selected item -> dst[(b*N + knot + koff)*stride + off + e]
        template<int COUNT>
        __device__ __forceinline__ void flush(float* dst, int row0, int stride, int off, int koff, int which /*0 non-terminal, 1 terminal, 2 all*/) const
        {
                const int nvalid = min(32, total - item0);
                for (int f = lane; f < nvalid * COUNT; f += 32) {
                        const int  i = f / COUNT, e = f - i * COUNT;
                        const int  rb = rowbase[i];
                        const bool ti = (rb >> 30) & 1;
                        if ((which == 0 && ti) || (which == 1 && !ti)) continue;
                        dst[((size_t)(rb & ~(1 << 30)) + koff) * stride + off + e] = stage[(row0 + e) * ST + i];
                }
        }
        // kind 0: cost blocks of knot k (Q, q, R, r); knot N-1 is the "terminal" item: Q_{N-1}, q_{N-1} evaluated at x_{N-2} against ref_{N-1}
        // (setup_kkt.cuh:83-100) and c_0 = x_0 - x_s
        __device__ __forceinline__ void cost_item(const Ctx& c) const
        {
                float ref3[3];
                sfor<0, 3>([&](auto ic) { ref3[ic] = c.ref[(size_t)b * 6 * c.N + 6 * k + ic]; });
                constexpr int rQ = 0, rQd = NQ * NQ, rq = rQd + NQ, rR = rq + NX, rr = rR + NU, rc0 = rr + NU;
                // Q = [[h h^T w + barrier terms, 0], [0, diag]], R = diag (plant cost Hessians, iiwa14_plant.cuh:400-450): only those entries
                // are staged; the indices are compile-time constants after inlining, so the stores of structural zeros fold away
                const Items<P> it = make_items<P>(c);
                it.template cost_grad_hess<true>(
                    xux, ref3, c.cs,
                    [&](int e, float v) {
                            const int i = e / NX, j = e % NX;
                            if (i < NQ && j < NQ)
                                    put(rQ + i * NQ + j, v);
                            else if (i == j)
                                    put(rQd + i - NQ, v);
                    },
                    [&](int e, float v) { put(rq + e, v); },
                    [&](int e, float v) {
                            if (e / NU == e % NU) put(rR + e / NU, v);
                    },
                    [&](int e, float v) { put(rr + e, v); }, term, term || Items<P>::pos_form_b_for(c.N));
                if (term && valid) {
                        const float* x0 = c.xu + (size_t)b * traj;
                        sfor<0, NX>([&](auto ic) { put(rc0 + ic, x0[ic] - c.xs[(size_t)b * NX + ic]); });
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
                flush<NX>(c.q, rq, NX, 0, 0, 2);
                flush<NU>(c.r, rr, NU, 0, 0, 0);
                flush<NX>(c.c, rc0, NX, 0, -(c.N - 1), 1);
        }
};

// =====================================================================================================
// k_kkt: one thread per work item, three kinds of items in separate warps (blockIdx.y = kind) so that no warp diverges:
//   kind 0  cost blocks of knot k = 0..N-1 (KktWarp::cost_item)
//   kind 1  linearised dynamics of knot k = 0..N-2, d/dq half: columns 0..nq-1 of A_k and the defect c_{k+1}
//   kind 2  d/dqd half: columns nq..nx-1 of A_k and B_k
// (The two dynamics halves repeat the M^-1 / RNEA prologue; splitting doubles the parallelism of what is a latency-bound
// kernel at batch 512.)  Results are transposed through shared memory so that HBM/L2 stores are coalesced per knot block.
// =====================================================================================================
template<class P>
__global__ void __launch_bounds__(32, GATO_KKT_MIN_BLOCKS) k_kkt(Ctx c)
{
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ;
        // staged floats per item: kind 0: Q (its nq x nq block and the lower diagonal; everything else in Q is a structural zero) | q |
        // R (diagonal) | r | c0;  kind 1: half of A | c;  kind 2: half of A | B  -- the largest.  26 KB per warp keeps 8 warps per SM.
        // The table-driven instantiations keep their state in local memory: they run better with twice the warps per multiprocessor and stage
        // one column (nx rows) at a time -- the cost item then sets the size.
        constexpr int ROWS = is_rt_plant<P> ? (NQ * NQ + NQ + NX + NU + NU + NX) : (NX * NQ + NX * NU);
        if (stopped_before(c, c.it)) return;
        __shared__ float stage[ROWS * KktWarp<P, ROWS>::ST];
        __shared__ int   rowbase[32];
        // (measured: dispatching the two long dynamics kinds before the short cost items is SLOWER, 94.6 vs 73.6 us at B = 512 -- eight dynamics warps
        // per multiprocessor streaming their code at the same time contend for instruction fetch; the cost warps in between stagger them)
        const int kind = blockIdx.y;
        KktWarp<P, ROWS> w;
        if (!w.init(c, stage, rowbase, kind == 0)) return;
        if (kind == 0) {
                w.cost_item(c);
                return;
        }
        float fext[6];
        sfor<0, 6>([&](auto ic) { fext[ic] = c.fext[6 * w.b + ic]; });
        constexpr int rA = 0, rX = NX * NQ;  // half of A (NX*NQ contiguous floats), then c (kind 1) or B (kind 2)
        const Items<P> it = make_items<P>(c);
        if constexpr (is_rt_plant<P>) {
                typename Items<P>::DynState st;
                it.prologue(w.xux, fext, st);
                const int half = kind - 1;
#pragma unroll 1
                for (int k = 0; k < NQ; k++) {
                        const int col = k + half * NQ;
                        if (half == 0)
                                it.template column<0>(k, st, w.xux + NQ, c.dt, [&](int e, float v) { w.put(e - col * NX, v); });
                        else
                                it.template column<1>(k, st, w.xux + NQ, c.dt, [&](int e, float v) { w.put(e - col * NX, v); });
                        __syncwarp();
                        w.template flush<NX>(c.A, 0, NX * NX, col * NX, 0, 2);
                        __syncwarp();
                }
                if (half == 0) {
                        it.defect(st, w.xux, c.dt, [&](int e, float v) { w.put(e, v); });
                        __syncwarp();
                        w.template flush<NX>(c.c, 0, NX, 0, 1, 2);
                } else {
#pragma unroll 1
                        for (int cb = 0; cb < NU; cb++) {
                                it.b_column(st, cb, c.dt, [&](int e, float v) { w.put(e - cb * NX, v); });
                                __syncwarp();
                                w.template flush<NX>(c.Bm, 0, NX * NU, cb * NX, 0, 2);
                                __syncwarp();
                        }
                }
                return;
        } else if (kind == 1) {
                it.template linearize_half_rolled<0>(
                    w.xux, fext, c.dt, [&](int e, float v) { w.put(rA + e, v); }, [&](int, float) {}, [&](int e, float v) { w.put(rX + e, v); });
                __syncwarp();
                w.template flush<NX * NQ>(c.A, rA, NX * NX, 0, 0, 2);
                w.template flush<NX>(c.c, rX, NX, 0, 1, 2);
        } else {
                it.template linearize_half_rolled<1>(
                    w.xux, fext, c.dt, [&](int e, float v) { w.put(rA + e - NX * NQ, v); }, [&](int e, float v) { w.put(rX + e, v); }, [&](int, float) {});
                __syncwarp();
                w.template flush<NX * NQ>(c.A, rA, NX * NX, NX * NQ, 0, 2);
                w.template flush<NX * NU>(c.Bm, rX, NX * NU, 0, 0, 2);
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
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ;
        constexpr int ROWS = NX * NU + NX;  // kind 1 stages the most: B | c
        if (stopped_before(c, c.it)) return;
        __shared__ float stage[ROWS * KktWarp<P, ROWS>::ST];
        __shared__ int   rowbase[32];
        const int        kind = blockIdx.y;
        KktWarp<P, ROWS> w;
        if (!w.init(c, stage, rowbase, kind == 0)) return;
        if (kind == 0) {
                w.cost_item(c);
                return;
        }
        float fext[6];
        sfor<0, 6>([&](auto ic) { fext[ic] = c.fext[6 * w.b + ic]; });
        const Items<P>              it = make_items<P>(c);
        typename Items<P>::DynState st;
        it.prologue(w.xux, fext, st);
        if (kind == 1) {
                constexpr int rB = 0, rc = NX * NU;
                it.linearize_base(st, w.xux, c.dt, [&](int e, float v) { w.put(rB + e, v); }, [&](int e, float v) { w.put(rc + e, v); });
                __syncwarp();
                w.template flush<NX * NU>(c.Bm, rB, NX * NU, 0, 0, 2);
                w.template flush<NX>(c.c, rc, NX, 0, 1, 2);
                return;
        }
        const int col = kind - 2;  // column of A
        it.linearize_column_any(col, st, w.xux + NQ, c.dt, [&](int e, float v) { w.put(e - col * NX, v); });
        __syncwarp();
        w.template flush<NX>(c.A, 0, NX * NX, col * NX, 0, 2);
}

}  // namespace gato
