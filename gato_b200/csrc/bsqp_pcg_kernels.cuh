// PCG kernels of the BSQP path (included by bsqp_kernels.cuh inside namespace gato):
//
//   k_pcg         CTA per solve, thread per matrix row: each thread keeps ITS ROW of S and of P^-1 (2 x 3nx floats) in registers for
//                 the whole solve; only the two shared vectors (p, r) live in shared memory.  The rows arrive through the TMA unit
//                 (cp.async.bulk + mbarrier), the row products run on Blackwell's packed FFMA2 / FADD2.  Builds the off-diagonal
//                 preconditioner blocks, runs PCG with the reference's reduction trees, recovers dz and does the convergence
//                 bookkeeping.  Replaces formSchurSystemBatchedKernel2 (schur_linsys.cuh:214-260), solvePCGBatchedKernel
//                 (pcg.cuh:14-148, which re-reads S and P^-1 from global memory every iteration), computeDzBatchedKernel
//                 (schur_linsys.cuh:316-431) and the host loop of bsqp.cuh:142-163.
//   k_pcg_stream  the same for horizons whose system does not fit the register file: rows streamed from L2 every iteration.
// -----------------------------------------------------------------------------------------------------
#pragma once
#include "bsqp_ctx.cuh"
#include "pcg_layout.h"
#include "sfor.h"

namespace gato {
__device__ __forceinline__ float warp_tree(float v)  // __shfl_down tree 16,8,4,2,1 -> lane 0 (linalg.cuh:215)
{
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = v + __shfl_down_sync(0xffffffffu, v, off);
        return v;
}

// The reference's row reduction (btdMatrixVectorProduct, linalg.cuh:197-216): lane l accumulates columns l and l+32,
// then the shuffle tree 16,8,4,2,1.  Evaluated depth-first by one thread: val(l, s) = val(l, 2s) + val(l+s, 2s).
template<int W, int L, int S>
__device__ __forceinline__ float row_tree(const float (&m)[W], const float (&v)[W])
{
        if constexpr (S == 32) {
                float p = fmaf(m[L], v[L], 0.0f);
                if constexpr (L + 32 < W) p = fmaf(m[L + 32], v[L + 32], p);
                return p;
        } else {
                const float lo = row_tree<W, L, 2 * S>(m, v);
                const float hi = row_tree<W, L + S, 2 * S>(m, v);
                return lo + hi;
        }
}

// asynchronous global -> shared copies (cp.async, LDGSTS): BYTES per copy = the largest of 16/8/4 that divides the per-knot block size, so
// that every solve's block is aligned to it
template<int BYTES>
__device__ __forceinline__ void cp_async_region(float* sdst, const float* gsrc, int nfloats, int tid, int nthreads)
{
        constexpr int  F = BYTES / 4;
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(sdst);
        for (int i = tid; i < nfloats / F; i += nthreads) {
                if constexpr (BYTES == 16)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * i), "l"(gsrc + 4 * i) : "memory");
                else if constexpr (BYTES == 8)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sbase + 8u * i), "l"(gsrc + 2 * i) : "memory");
                else
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sbase + 4u * i), "l"(gsrc + i) : "memory");
        }
}
constexpr int cp_bytes(int block_floats) { return (block_floats * 4) % 16 == 0 ? 16 : ((block_floats * 4) % 8 == 0 ? 8 : 4); }
// dz_x,k = -Qinv_k (q_k - lambda_k + A_k^T lambda_{k+1}), dz_u,k = -Rinv_k (r_k + B_k^T lambda_{k+1}); the bracketed residuals are
// stored back into q, r (computeDzBatchedKernel, schur_linsys.cuh:331-430).  One warp per knot; wbuf = 64 floats per warp.
template<int NX, int NU, int LDV = NX>
__device__ __forceinline__ void dz_phase(const Ctx& c, int b, int N, int n, int warp, int lane, int nwarps, float* dzbuf, const float* lam, const float* Ab, const float* Bb,
                                         const float* Qib, const float* Rib, const float* qb, const float* rb, int k_begin = 0, int k_end = -1)
{
        // knots k_begin .. k_end-1 (default: all).  lam: padded lambda, block of knot k_begin - 1 first; Ab, Bb, Qib, Rib, qb, rb: the A, B, Q^-1,
        // R^-1, q, r blocks of these knots (knot k at (k - k_begin) * block size) -- in global memory, or prefetched to shared memory.  The
        // residuals are written back to c.q / c.r in global memory either way.
        constexpr int NX2 = NX * NX;
        const size_t  kb = (size_t)b * N;
        float*        wbuf = dzbuf + warp * 64;
        const int     traj = (NX + NU) * N - NU;
        if (k_end < 0) k_end = N;
        lam -= (size_t)k_begin * LDV;
        Ab -= (size_t)k_begin * NX2, Bb -= (size_t)k_begin * NX * NU, Qib -= (size_t)k_begin * NX2, Rib -= (size_t)k_begin * NU * NU, qb -= (size_t)k_begin * NX, rb -= (size_t)k_begin * NU;
        for (int k = k_begin + warp; k < k_end; k += nwarps) {
                const float* lk = lam + (k + 1) * LDV;  // LDV: floats between consecutive blocks of lambda (NX: packed; kSlot: k_pcg's slotted vectors)
                const float* lk1 = lam + (k + 2) * LDV;
                __syncwarp();
                if (lane < NX) {
                        float scr = 0.0f;
                        if (k < N - 1) {
                                const float* Ak = Ab + (size_t)k * NX2;
                                float        sum = 0.0f;
#pragma unroll
                                for (int j = 0; j < NX; j++) sum = fmaf(lk1[j], Ak[lane * NX + j], sum);
                                scr = -sum;
                        }
                        scr = scr + lk[lane];
                        wbuf[lane] = qb[(size_t)k * NX + lane] - scr;
                } else if (lane >= 16 && lane < 16 + NU && k < N - 1) {
                        const int    x = lane - 16;
                        const float* Bk = Bb + (size_t)k * NX * NU;
                        float        sum = 0.0f;
#pragma unroll
                        for (int j = 0; j < NX; j++) sum = fmaf(lk1[j], Bk[x * NX + j], sum);
                        wbuf[32 + x] = rb[(size_t)k * NU + x] - (-sum);
                }
                __syncwarp();
                if (lane < NX) {
                        const float* Qi = Qib + (size_t)k * NX2;
                        float        sum = 0.0f;
#pragma unroll
                        for (int j = 0; j < NX; j++) sum = fmaf(Qi[j * NX + lane], wbuf[j], sum);
                        c.dz[(size_t)b * traj + (size_t)k * (NX + NU) + lane] = -1.0f * sum;
                        c.q[(kb + k) * NX + lane] = wbuf[lane];
                } else if (lane >= 16 && lane < 16 + NU) {
                        const int x = lane - 16;
                        if (k < N - 1) {
                                const float* Ri = Rib + (size_t)k * NU * NU;
                                float        sum = 0.0f;
#pragma unroll
                                for (int j = 0; j < NU; j++) sum = fmaf(Ri[j * NU + x], wbuf[32 + j], sum);
                                c.dz[(size_t)b * traj + (size_t)k * (NX + NU) + NX + x] = -1.0f * sum;
                                c.r[(kb + k) * NU + x] = wbuf[32 + x];
                        } else {
                                c.r[(kb + k) * NU + x] = 0.0f;
                        }
                }
                if (c.kkt_qmax) {
                        // telemetry the reference computes on the host and discards (bsqp.cuh:149-150): max |q residual| and max |c| over the solve's
                        // state entries; non-negative floats order like their bit patterns, so the maximum is taken on the bits
                        const bool     st = lane < NX;
                        const unsigned mq = __reduce_max_sync(0xffffffffu, st ? __float_as_uint(fabsf(wbuf[lane])) : 0u);
                        const unsigned mc = __reduce_max_sync(0xffffffffu, st ? __float_as_uint(fabsf(c.c[(kb + k) * NX + lane])) : 0u);
                        if (lane == 0) {
                                atomicMax(&c.kkt_qmax[(size_t)c.it * c.B + b], mq);
                                atomicMax(&c.kkt_cmax[(size_t)c.it * c.B + b], mc);
                        }
                }
        }
}

// The warp's 32 consecutive rows r0 .. r0+31 of a block-tridiagonal matrix (row-major, W floats per row; rows outside [0, nrows) do not
// exist) -> the warp's shared-memory tile, as coalesced 16-byte asynchronous copies (cp.async.cg: global -> shared without passing
// through registers or L1).  Thread-per-row loads straight from global memory are 168-byte strided: every 32-byte sector is requested
// four times and the kernel ends up bound by L1/L2 request throughput (ncu: L1/TEX 65 %, L2 56 %, DRAM 8 % -- the 148 solves in flight
// fit in L2).  r0 and nrows are even, so both ends of the copy are 16-byte aligned.
template<int W>
__device__ __forceinline__ void warp_tile_load(float* tile, const float* gM, int r0, int nrows, int lane)
{
        __syncwarp();  // every lane is done reading the previous contents of the tile
        const int lo = r0 > 0 ? r0 : 0, hi = (r0 + 32 < nrows) ? r0 + 32 : nrows;
        if (hi > lo) {
                const float*   src = gM + (size_t)lo * W;
                const unsigned dst = (unsigned)__cvta_generic_to_shared(tile + (lo - r0) * W);
                const int      nchunk = (hi - lo) * W / 4;
                for (int i = lane; i < nchunk; i += 32) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + 4 * i) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
}
// ---- Blackwell / Hopper asynchronous-copy plumbing (PTX): mbarrier + bulk copies through the TMA unit (cp.async.bulk, SASS UBLKCP) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void     mbar_init(unsigned long long* bar, unsigned arrivals)
{
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// make the barrier initialisation visible to the asynchronous proxy (the TMA unit completes transactions on it)
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// order this thread's generic-proxy accesses to shared memory before later asynchronous-proxy (bulk copy) writes to the same bytes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes)
{
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// one contiguous span global -> shared, completion (bytes) signalled on the mbarrier; addresses and size are multiples of 16
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar)
{
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
        unsigned done;
        do {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        } while (!done);
}

// shared-memory accesses of the PCG iteration on 32-bit shared-window addresses held in registers (the generic-pointer forms make the
// compiler re-derive the window base from SR_CgaCtaId inside the loop, on the critical path between the reduction stages)
__device__ __forceinline__ float4 lds128(unsigned a)
{
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
        return v;
}
__device__ __forceinline__ float2 lds64(unsigned a)
{
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
        return v;
}
__device__ __forceinline__ void sts32(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
// a shared address the compiler cannot see through: it stays in a register across the loop (otherwise `window base + constant` is re-derived)
__device__ __forceinline__ unsigned opaque_addr(unsigned a)
{
        unsigned r;
        asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(a));
        return r;
}

// (M v)[row] with the reference's reduction tree (btdMatrixVectorProduct, linalg.cuh:197-216: lane l accumulates columns l and l+32, then
// the shuffle tree 16, 8, 4, 2, 1), evaluated by ONE thread on PACKED pairs: Blackwell's FFMA2 / FADD2 (fma.rn.f32x2, add.rn.f32x2) work
// on two fp32 lanes per instruction, each rounded exactly like the scalar operation.  acc[j] = (leaf(2j), leaf(2j+1)); every level of the
// tree adds pair j+h to pair j; the last level adds the two halves of the remaining pair.  win: shared address of the window's first block.
template<int NX>
__device__ __forceinline__ float matvec_packed(const float2 (&M)[3 * NX / 2], unsigned win)
{
        constexpr int W = 3 * NX, NP = W / 2, PPS = NX / 2;  // pairs per row, pairs per slot
        static_assert(NX % 2 == 0 && W >= 32 && W <= 64 && NX <= kSlot, "row_tree geometry");
        float2 v[NP];
        sfor<0, 3>([&](auto sc) {
                constexpr int sl = sc;
                sfor<0, (NX + 3) / 4>([&](auto cc) {
                        constexpr int ch = cc;
                        if constexpr (2 * ch + 1 < PPS) {
                                const float4 t = lds128(win + 4u * (sl * kSlot + 4 * ch));
                                v[sl * PPS + 2 * ch] = make_float2(t.x, t.y);
                                v[sl * PPS + 2 * ch + 1] = make_float2(t.z, t.w);
                        } else {
                                v[sl * PPS + 2 * ch] = lds64(win + 4u * (sl * kSlot + 4 * ch));
                        }
                });
        });
        float2 acc[16];
        sfor<0, 16>([&](auto jc) { acc[jc] = __ffma2_rn(M[jc], v[jc], make_float2(0.0f, 0.0f)); });
        sfor<16, NP>([&](auto jc) { acc[jc - 16] = __ffma2_rn(M[jc], v[jc], acc[jc - 16]); });
        sfor<0, 8>([&](auto jc) { acc[jc] = __fadd2_rn(acc[jc], acc[jc + 8]); });
        sfor<0, 4>([&](auto jc) { acc[jc] = __fadd2_rn(acc[jc], acc[jc + 4]); });
        sfor<0, 2>([&](auto jc) { acc[jc] = __fadd2_rn(acc[jc], acc[jc + 2]); });
        const float2 u = __fadd2_rn(acc[0], acc[1]);
        return u.x + u.y;
}

#ifndef GATO_PCG_DOT_STAGE1
#define GATO_PCG_DOT_STAGE1 0
#endif
#ifndef GATO_PCG_DOT_STAGE2
#define GATO_PCG_DOT_STAGE2 0
#endif
// block::dot (linalg.cuh:291-327): thread i contributes a_i * b_i; warp tree (shfl_down 16, 8, 4, 2, 1); lane 0 -> scratch[warp]; then the
// same tree over the per-warp sums (lanes beyond the warp count hold +0.0f).  Both trees are evaluated by ONE lane from shared memory on
// packed adds instead of as five dependent shuffle round trips: every lane drops its product into the warp's 32-float row, lane 0 folds the
// row (offset 16: elements l and l+16 ... offset 1: the two halves of the last pair) and publishes the warp's sum.
//   tree32: the 32 floats at shared address a -> lane 0's value of the reference's shuffle tree
__device__ __forceinline__ float tree32(unsigned a)
{
        float4 e[8];
#pragma unroll
        for (int i = 0; i < 8; i++) e[i] = lds128(a + 16u * i);
        float2 t[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {  // offset 16: elements 4i..4i+3 with 16+4i..16+4i+3
                t[2 * i] = __fadd2_rn(make_float2(e[i].x, e[i].y), make_float2(e[i + 4].x, e[i + 4].y));
                t[2 * i + 1] = __fadd2_rn(make_float2(e[i].z, e[i].w), make_float2(e[i + 4].z, e[i + 4].w));
        }
#pragma unroll
        for (int i = 0; i < 4; i++) t[i] = __fadd2_rn(t[i], t[i + 4]);  // offset 8
        t[0] = __fadd2_rn(t[0], t[2]), t[1] = __fadd2_rn(t[1], t[3]);   // offset 4
        const float2 y = __fadd2_rn(t[0], t[1]);                        // offset 2
        return y.x + y.y;                                               // offset 1
}
//   tree16: the second stage over at most 16 per-warp sums: the offset-16 level only adds +0.0f, which is exact here -- a partial is never
//   -0.0f (each thread's term is fmaf(a, b, +0.0f), and sums of values that are not -0 are not -0)
__device__ __forceinline__ float tree16(unsigned a)
{
        const float4 p = lds128(a), q = lds128(a + 16u), r = lds128(a + 32u), w = lds128(a + 48u);  // entries >= #warps are +0.0f
        const float2 t01 = __fadd2_rn(make_float2(p.x, p.y), make_float2(r.x, r.y)), t23 = __fadd2_rn(make_float2(p.z, p.w), make_float2(r.z, r.w));  // offset 8
        const float2 u01 = __fadd2_rn(make_float2(q.x, q.y), make_float2(w.x, w.y)), u23 = __fadd2_rn(make_float2(q.z, q.w), make_float2(w.z, w.w));
        const float2 w01 = __fadd2_rn(t01, u01), w23 = __fadd2_rn(t23, u23);  // offset 4
        const float2 y = __fadd2_rn(w01, w23);                                // offset 2
        return y.x + y.y;                                                     // offset 1
}

// MAXT: the largest block the instantiation is launched with (480: iiwa14 up to N = 32; 512: everything else that fits a thread per padded index).
template<class P, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_pcg(Ctx c)
{
        // Thread t owns PADDED vector index i = t (so real warp w == virtual warp w of the reference's block::dot) and, for
        // NX <= i < NX + N*NX, matrix row r = i - NX: its rows of S and P^-1 and its elements of x, r, p live in registers, as float2 pairs.
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ, NX2 = NX * NX, W = 3 * NX, NP = W / 2, HP = NX / 2;
        pdl_launch_dependents();  // the merit / line-search launch of this iteration may fill SMs this grid leaves idle (its last, partial wave)
        if (stopped_before(c, c.it)) return;
        extern __shared__ __align__(16) float sm[];
        const int                             N = c.N, b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
        const int                             nrows = N * NX, n = (N + 2) * NX;
        const int                             warp = tid >> 5, lane = tid & 31, nwarps = T >> 5;
        unsigned long long*                   bars = reinterpret_cast<unsigned long long*>(sm);  // [0]: rows staged, [1]: dz operands staged
        float*                                vp = sm + 4;
        float*                                vr = vp + (N + 2) * kSlot;
        float*                                scratchA = vr + (N + 2) * kSlot;
        float*                                scratchB = scratchA + 16;
        float*                                prod = scratchB + 16;  // one 32-float row per warp: the lanes' dot-product terms
        float*                                dzbuf = prod + 32 * nwarps;
        float*                                stageA = dzbuf + 64 * nwarps;
        const int                             stageA_floats = 3 * N * NX2 > dz_stage_floats<NX, NU>(N) ? 3 * N * NX2 : dz_stage_floats<NX, NU>(N);
        float*                                mains = stageA + stageA_floats;
        const size_t                          kb = (size_t)b * N;
        const float*                          gS = c.S + kb * 3 * NX2;
        float*                                gP = c.Pinv + kb * 3 * NX2;
        const int                             row = tid - NX;
        const bool                            row_ok = (row >= 0) && (row < nrows);
        const int                             br = row_ok ? row / NX : 0, ry = row_ok ? row % NX : 0;
        const bool                            k2 = (c.flags & F_K2) != 0, need_rows = (c.flags & (F_K2 | F_PCG)) != 0;
        const bool                            in_vec = tid < n;
        const int                             own = (tid / NX) * kSlot + tid % NX;  // this thread's element of the slotted vectors

        if (tid == 0) {
                mbar_init(&bars[0], 1);
                mbar_init(&bars[1], 1);
                fence_mbar_init();
        }
        __syncthreads();
        // The solve's rows of S (and the packed main blocks of P^-1) arrive as two bulk copies through the TMA unit: one thread issues them,
        // the data lands in shared memory without passing through registers, and each thread then picks up its row with conflict-free
        // 8-byte shared loads (a row is 3*NX floats: consecutive rows start 2*(3*NX/2 mod 16) banks apart).
        if (tid == 0 && need_rows) {
                const unsigned bS = sizeof(float) * 3 * N * NX2, bP = sizeof(float) * N * NX2;
                mbar_arrive_expect_tx(&bars[0], bS + (k2 ? bP : 0u));
                bulk_g2s(stageA, gS, bS, &bars[0]);
                if (k2) bulk_g2s(mains, c.Pmain + kb * NX2, bP, &bars[0]);
        }
        for (int i = tid; i < 2 * (N + 2) * kSlot + 32 + 32 * nwarps; i += T) vp[i] = 0.0f;  // vp, vr (including the padding), both dot scratch rows, products

        float2 S2[NP], P2[NP];
        sfor<0, NP>([&](auto ic) { S2[ic] = make_float2(0.0f, 0.0f), P2[ic] = make_float2(0.0f, 0.0f); });
        if (need_rows) {
                static_assert(NX % 2 == 0, "blocks are float2-aligned");
                mbar_wait(&bars[0], 0);
                if (row_ok) {
                        const float2* s2 = reinterpret_cast<const float2*>(stageA + (size_t)row * W);
                        sfor<0, NP>([&](auto ic) { S2[ic] = s2[ic]; });
                        if (k2) {
                                // only the main block exists so far; the others hold the zeros that row 0's left and row N-1's right block keep
                                // (schur_linsys.cuh:227-259)
                                const float2* p2 = reinterpret_cast<const float2*>(mains + (size_t)row * NX);
                                sfor<0, HP>([&](auto ic) { P2[HP + ic] = p2[ic]; });
                        } else {
                                const float2* p2 = reinterpret_cast<const float2*>(gP + (size_t)row * W);  // stage test: complete rows, reference layout
                                sfor<0, NP>([&](auto ic) { P2[ic] = __ldcs(p2 + ic); });
                        }
                }
        }
        if (k2) {
                // left_{k+1} = -(Theta_k (phi_k Theta_{k-1})), right_k = left_{k+1}^T  (schur_linsys.cuh:227-259); Theta = the staged main blocks.
                // Each 14-term chain runs over j in the reference's order; two neighbouring columns x, x+1 share one packed FFMA2.
                __syncthreads();  // every row of S is in registers: stage A is free
                float* scr2 = stageA;
                if (row_ok && br >= 1) {
                        // scr(y, x) = sum_j phi(y, j) * Theta_{k-1}(j, x), k = br-1; phi row y = this thread's S left block
                        const float* tk1 = mains + (size_t)(br - 1) * NX2;
                        float2       acc[HP];
                        sfor<0, HP>([&](auto xc) { acc[xc] = make_float2(0.0f, 0.0f); });
                        sfor<0, NX>([&](auto jc) {
                                constexpr int j = jc;
                                const float   a = (j & 1) ? S2[j / 2].y : S2[j / 2].x;
                                const float2* t2 = reinterpret_cast<const float2*>(tk1 + j * NX);
                                sfor<0, HP>([&](auto xc) { acc[xc] = __ffma2_rn(make_float2(a, a), t2[xc], acc[xc]); });
                        });
                        float2* o2 = reinterpret_cast<float2*>(scr2 + (size_t)(br - 1) * NX2 + ry * NX);
                        sfor<0, HP>([&](auto xc) { o2[xc] = acc[xc]; });
                }
                __syncthreads();
                float2 outrow[HP];
                if (row_ok && br >= 1) {
                        // out(y, x) = sum_j Theta_k(y, j) * scr(j, x); Theta_k row y = this thread's P main block
                        const float* sc = scr2 + (size_t)(br - 1) * NX2;
                        sfor<0, HP>([&](auto xc) { outrow[xc] = make_float2(0.0f, 0.0f); });
                        sfor<0, NX>([&](auto jc) {
                                constexpr int j = jc;
                                const float   a = (j & 1) ? P2[HP + j / 2].y : P2[HP + j / 2].x;
                                const float2* t2 = reinterpret_cast<const float2*>(sc + j * NX);
                                sfor<0, HP>([&](auto xc) { outrow[xc] = __ffma2_rn(make_float2(a, a), t2[xc], outrow[xc]); });
                        });
                        sfor<0, HP>([&](auto xc) { P2[xc] = make_float2(-outrow[xc].x, -outrow[xc].y); });  // left block of this row
                }
                __syncthreads();  // everyone is done reading mains / scr2
                if (row_ok && br >= 1) {
                        float2* o2 = reinterpret_cast<float2*>(mains + (size_t)(br - 1) * NX2 + ry * NX);
                        sfor<0, HP>([&](auto xc) { o2[xc] = outrow[xc]; });
                }
                __syncthreads();
                if (row_ok && br < N - 1) {
                        // right block of row (br, x = ry): entry (x, y) = -out_k(y, x), k = br
                        const float* ok = mains + (size_t)br * NX2;
                        sfor<0, HP>([&](auto yc) { P2[2 * HP + yc] = make_float2(-ok[(2 * yc) * NX + ry], -ok[(2 * yc + 1) * NX + ry]); });
                }
                if ((c.flags & F_WRITE_P) && row_ok) {
                        float2* g2 = reinterpret_cast<float2*>(gP + (size_t)row * W);
                        sfor<0, NP>([&](auto ic) { g2[ic] = P2[ic]; });
                }
        }

        // The operands of the primal step (A, B, Q^-1, R^-1, q, r of this solve, 71 KB) were written two kernels ago and have left L2 by now:
        // read on demand, each of the dz phase's dependent steps would wait for HBM.  They are prefetched into stage A (free once the rows
        // are in registers and K2 is done) while the PCG iterations run -- six bulk copies through the TMA unit signalled on an mbarrier
        // (or, for horizons whose per-solve arrays are not 16-byte multiples, cp.async) -- and lambda is left in shared memory by the PCG phase.
        float* sA = stageA;
        float* sB = sA + pad4(N * NX2);
        float* sQi = sB + pad4(N * NX * NU);
        float* sRi = sQi + pad4(N * NX2);
        float* sq = sRi + pad4(N * NU * NU);
        float* sr = sq + pad4(N * NX);
        const bool bulk_dz = dz_bulk_ok<NX, NU>(N);
        if (c.flags & F_DZ) {
                fence_proxy_async();
                __syncthreads();  // all generic accesses to stage A are done and ordered before the asynchronous writes
                if (bulk_dz) {
                        if (tid == 0) {
                                const unsigned bA = 4u * N * NX2, bB = 4u * N * NX * NU, bR = 4u * N * NU * NU, bq = 4u * N * NX, br_ = 4u * N * NU;
                                mbar_arrive_expect_tx(&bars[1], 2 * bA + bB + bR + bq + br_);
                                bulk_g2s(sA, c.A + kb * NX2, bA, &bars[1]);
                                bulk_g2s(sB, c.Bm + kb * NX * NU, bB, &bars[1]);
                                bulk_g2s(sQi, c.Qinv + kb * NX2, bA, &bars[1]);
                                bulk_g2s(sRi, c.Rinv + kb * NU * NU, bR, &bars[1]);
                                bulk_g2s(sq, c.q + kb * NX, bq, &bars[1]);
                                bulk_g2s(sr, c.r + kb * NU, br_, &bars[1]);
                        }
                } else {
                        cp_async_region<cp_bytes(NX2)>(sA, c.A + kb * NX2, (N - 1) * NX2, tid, T);
                        cp_async_region<cp_bytes(NX * NU)>(sB, c.Bm + kb * NX * NU, (N - 1) * NX * NU, tid, T);
                        cp_async_region<cp_bytes(NX2)>(sQi, c.Qinv + kb * NX2, N * NX2, tid, T);
                        cp_async_region<cp_bytes(NU * NU)>(sRi, c.Rinv + kb * NU * NU, (N - 1) * NU * NU, tid, T);
                        cp_async_region<cp_bytes(NX)>(sq, c.q + kb * NX, N * NX, tid, T);
                        cp_async_region<cp_bytes(NU)>(sr, c.r + kb * NU, (N - 1) * NU, tid, T);
                        asm volatile("cp.async.commit_group;" ::: "memory");
                }
        } else {
                __syncthreads();
        }

        int iters = 0;
        if (c.flags & F_PCG) {
                const float* gam = c.gamma + (size_t)b * n;
                float*       lam = c.lambda + (size_t)b * n;
                const float  eps = c.pcg_tol[b];
                const float  abs_tol = 1e-6f;
                const bool   skip = c.conv[b] != 0;  // pcg.cuh:29-32
                const unsigned wp = opaque_addr(smem_u32(vp + br * kSlot)), wr = opaque_addr(smem_u32(vr + br * kSlot));  // this row's window: padded blocks br .. br+2
                const unsigned own_p = opaque_addr(smem_u32(vp + own)), own_r = opaque_addr(smem_u32(vr + own));
                const unsigned sA = opaque_addr(smem_u32(scratchA)), sB = opaque_addr(smem_u32(scratchB));
                const unsigned my_prod = smem_u32(prod + tid), row_prod = smem_u32(prod + 32 * warp);
                // dot, phase 1 (before the CTA barrier): the warp's partial sum -> scratch[warp]
                auto dot_partial = [&](float term, unsigned scratch) {
#if GATO_PCG_DOT_STAGE1 == 1
                        sts32(my_prod, term);
                        __syncwarp();
                        if (lane == 0) sts32(scratch + 4u * warp, tree32(row_prod));
                        __syncwarp();  // the row is free for the next product
#else
                        const float sum = warp_tree(term);
                        if (lane == 0) sts32(scratch + 4u * warp, sum);
#endif
                };
                // dot, phase 2 (after it): lane 0 folds the per-warp sums, the warp takes its value
                auto dot_final = [&](unsigned scratch) -> float {
#if GATO_PCG_DOT_STAGE2 == 1
                        float v = 0.0f;
                        if (lane == 0) v = tree16(scratch);
                        return __shfl_sync(0xffffffffu, v, 0);
#else
                        return tree16(scratch);
#endif
                };
                if (!skip) {
                        float x_i = in_vec ? lam[tid] : 0.0f;
                        if (in_vec) sts32(own_p, x_i);  // vp temporarily holds x for r = gamma - S x
                        __syncthreads();
                        float r_i = 0.0f, p_i = 0.0f, z_i = 0.0f;
                        {
                                const float sx = row_ok ? matvec_packed<NX>(S2, wp) : 0.0f;
                                r_i = in_vec ? (gam[tid] - sx) : 0.0f;
                                if (in_vec) sts32(own_r, r_i);
                        }
                        __syncthreads();
                        z_i = row_ok ? matvec_packed<NX>(P2, wr) : 0.0f;
                        p_i = z_i;
                        if (in_vec) sts32(own_p, p_i);
                        dot_partial(fmaf(r_i, z_i, 0.0f), sA);
                        __syncthreads();
                        float rho = dot_final(sA);
                        if (!(fabsf(rho) < abs_tol)) {
                                const float rho_init = fabsf(rho);
                                for (int itn = 0; itn < c.max_pcg; itn++) {
                                        iters++;
                                        const float Ap_i = row_ok ? matvec_packed<NX>(S2, wp) : 0.0f;
                                        dot_partial(fmaf(p_i, Ap_i, 0.0f), sB);
                                        __syncthreads();
                                        const float alpha = rho / dot_final(sB);
                                        x_i = fmaf(alpha, p_i, x_i);
                                        r_i = fmaf(-alpha, Ap_i, r_i);
                                        if (in_vec) sts32(own_r, r_i);
                                        __syncthreads();
                                        z_i = row_ok ? matvec_packed<NX>(P2, wr) : 0.0f;
                                        dot_partial(fmaf(r_i, z_i, 0.0f), sA);
                                        __syncthreads();
                                        const float rho_new = dot_final(sA);
                                        if (fabsf(rho_new) < fmaf(eps, rho_init, abs_tol)) break;
                                        const float beta = rho_new / rho;
                                        rho = rho_new;
                                        p_i = fmaf(beta, p_i, z_i);
                                        if (in_vec) sts32(own_p, p_i);
                                        __syncthreads();
                                }
                                if (in_vec) lam[tid] = x_i;
                        }
                }
                if (tid == 0) {
                        if (c.pcg_log) c.pcg_log[(size_t)c.it * c.B + b] = iters;
                        if (c.flags & F_BOOK) {
                                // bsqp.cuh:153-163: a solve is flagged once PCG performs no iteration; count flagged solves
                                int cv = c.conv[b];
                                if (iters == 0) cv = 1;
                                c.conv[b] = cv;
                                atomicAdd(cv ? &c.num_solved[c.it] : &c.num_unsolved[c.it], 1u);
                        }
                }
                __syncthreads();
        }

        if (c.flags & F_DZ) {
                // lambda of this solve (possibly just updated by threads of this CTA) -> shared memory (vp is free now)
                if (in_vec) vp[own] = c.lambda[(size_t)b * n + tid];
                if (bulk_dz)
                        mbar_wait(&bars[1], 0);
                else
                        asm volatile("cp.async.wait_all;" ::: "memory");
                __syncthreads();
                dz_phase<NX, NU, kSlot>(c, b, N, n, warp, lane, nwarps, dzbuf, vp, sA, sB, sQi, sRi, sq, sr);
        }
        if (c.flags & F_BOOK) {
                // hand solve b over to a merit / line-search CTA that may already be waiting for it: every write of this CTA, then the flag
                __syncthreads();
                if (tid == 0) {
                        __threadfence();
                        st_release_gpu(c.pcg_done + (size_t)c.it * c.B + b, 1u);
                }
        }
}


// -----------------------------------------------------------------------------------------------------
// k_pcg_stream: the same algorithm for horizons whose Schur system does not fit the register file ((N+2)*nx > 512, e.g. N = 128):
// 1024 threads (exactly the reference's PCG block, so thread t owns padded indices t, t+1024, ... like block::dot); the rows of S and
// P^-1 are streamed in every matvec (what the reference does for every N, pcg.cuh:100,119) -- from L2, which holds the systems of the
// 148 solves in flight -- through per-warp shared-memory tiles filled with coalesced asynchronous copies; the off-diagonal P^-1 blocks
// are built through shared memory and written back to global memory first.
// -----------------------------------------------------------------------------------------------------
// this lane's row (already in the warp's tile) times the padded shared-memory vector v
template<int NX>
__device__ __forceinline__ float tile_row_matvec(const float* trow, int row, const float* v)
{
        constexpr int W = 3 * NX;
        float         m[W], vv[W];
        const float2* m2 = reinterpret_cast<const float2*>(trow);
        const float2* v2 = reinterpret_cast<const float2*>(v + (row / NX) * NX);
#pragma unroll
        for (int i = 0; i < W / 2; i++) {
                const float2 a = m2[i], t = v2[i];
                m[2 * i] = a.x, m[2 * i + 1] = a.y;
                vv[2 * i] = t.x, vv[2 * i + 1] = t.y;
        }
        return row_tree<W, 0, 1>(m, vv);
}

template<class P, int RPT>
__global__ void __launch_bounds__(1024, 1) k_pcg_stream(Ctx c)
{
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ, NX2 = NX * NX, W = 3 * NX, T = 1024;
        if (stopped_before(c, c.it)) return;
        extern __shared__ __align__(16) float sm[];
        // tid is read through asm so that the compiler has no range for it: with the [0,1024) range nvcc 12.9 folds
        // sext((tid + c - NX) * W) into zext32(tid * W - NX * W) + c * W, wrong for tid < NX (seen in PTX; faulted at N = 128)
        int tid;
        asm("mov.u32 %0, %%tid.x;" : "=r"(tid));
        const int N = c.N, b = blockIdx.x;
        const int nrows = N * NX, n = (N + 2) * NX;
        const int warp = tid >> 5, lane = tid & 31, nwarps = T >> 5;
        float*                                vp = sm;
        float*                                vr = vp + n;
        float*                                scratchA = vr + n;
        float*                                scratchB = scratchA + 32;
        float*                                dzbuf = scratchB + 32;
        float*                                scr2 = dzbuf + 64 * nwarps;  // (N-1) * NX2 during K2; afterwards the 32 per-warp row tiles (32 x W floats each)
        float*                                tile = scr2 + (size_t)warp * 32 * W;
        const float*                          trow = tile + lane * W;
        const size_t                          kb = (size_t)b * N;
        const float*                          gS = c.S + kb * 3 * NX2;
        float*                                gP = c.Pinv + kb * 3 * NX2;
        for (int i = tid; i < 2 * n + 64; i += T) vp[i] = 0.0f;  // vp, vr and both dot scratch rows

        if (c.flags & F_K2) {
                // scr_k = phi_k Theta_{k-1};  out_k = Theta_k scr_k;  left_{k+1} = -out_k, right_k = -out_k^T   (schur_linsys.cuh:227-259)
                const float* gPm = c.Pmain + kb * NX2;  // main blocks as k_schur packed them
                for (int e = tid; e < (N - 1) * NX2; e += T) {
                        const int    k = e / NX2, y = (e % NX2) / NX, x = e % NX;
                        const float* ph = gS + (size_t)((k + 1) * NX + y) * W;  // S left block of row k+1, row y
                        const float* tk1 = gPm + (size_t)k * NX2;               // main block of row k
                        float        sum = 0.0f;
#pragma unroll
                        for (int j = 0; j < NX; j++) sum = fmaf(ph[j], tk1[j * NX + x], sum);
                        scr2[e] = sum;
                }
                // the streamed matvec reads complete rows in the reference layout: main blocks into place
                for (int e = tid; e < N * NX2; e += T) gP[(size_t)(e / NX) * W + NX + e % NX] = gPm[e];
                __syncthreads();
                for (int e = tid; e < (N - 1) * NX2; e += T) {
                        const int    k = e / NX2, y = (e % NX2) / NX, x = e % NX;
                        const float* tk = gPm + (size_t)(k + 1) * NX2 + y * NX;  // main block of row k+1, row y
                        const float* sc = scr2 + (size_t)k * NX2;
                        float        sum = 0.0f;
#pragma unroll
                        for (int j = 0; j < NX; j++) sum = fmaf(tk[j], sc[j * NX + x], sum);
                        gP[(size_t)((k + 1) * NX + y) * W + x] = -sum;
                        gP[(size_t)(k * NX + x) * W + 2 * NX + y] = -sum;
                }
                __threadfence_block();
        }
        __syncthreads();

        int iters = 0;
        if (c.flags & F_PCG) {
                const float* gam = c.gamma + (size_t)b * n;
                float*       lam = c.lambda + (size_t)b * n;
                const float* gPc = gP;
                const float  eps = c.pcg_tol[b];
                const float  abs_tol = 1e-6f;
                const bool   skip = c.conv[b] != 0;
                if (!skip) {
                        float x_[RPT], r_[RPT], p_[RPT], z_[RPT], Ap_[RPT];
                        // thread t owns padded indices t, t+1024, ... (the reference's block::dot geometry, linalg.cuh:306)
#pragma unroll
                        for (int j = 0; j < RPT; j++) {
                                const int i = tid + j * T;
                                x_[j] = (i < n) ? lam[i] : 0.0f;
                                if (i < n) vp[i] = x_[j];  // vp temporarily holds x for r = gamma - S x
                        }
                        __syncthreads();
#pragma unroll
                        for (int j = 0; j < RPT; j++) {
                                const int   i = tid + j * T;
                                const bool  rok = (i >= NX) && (i < NX + nrows);
                                warp_tile_load<W>(tile, gS, warp * 32 + j * T - NX, nrows, lane);
                                const float sx = rok ? tile_row_matvec<NX>(trow, i - NX, vp) : 0.0f;
                                r_[j] = (i < n) ? (gam[i] - sx) : 0.0f;
                                if (i < n) vr[i] = r_[j];
                        }
                        __syncthreads();
                        float prod = 0.0f;
#pragma unroll
                        for (int j = 0; j < RPT; j++) {
                                const int  i = tid + j * T;
                                const bool rok = (i >= NX) && (i < NX + nrows);
                                warp_tile_load<W>(tile, gPc, warp * 32 + j * T - NX, nrows, lane);
                                                z_[j] = rok ? tile_row_matvec<NX>(trow, i - NX, vr) : 0.0f;
                                p_[j] = z_[j];
                                prod = fmaf(r_[j], z_[j], prod);
                        }
                        __syncthreads();  // all reads of vp (as x) are done before it is overwritten with p
#pragma unroll
                        for (int j = 0; j < RPT; j++) {
                                const int i = tid + j * T;
                                if (i < n) vp[i] = p_[j];
                        }
                        {
                                const float s = warp_tree(prod);
                                if (lane == 0) scratchA[warp] = s;
                        }
                        __syncthreads();
                        float rho = __shfl_sync(0xffffffffu, warp_tree(scratchA[lane]), 0);
                        if (!(fabsf(rho) < abs_tol)) {
                                const float rho_init = fabsf(rho);
                                for (int itn = 0; itn < c.max_pcg; itn++) {
                                        iters++;
                                        prod = 0.0f;
#pragma unroll
                                        for (int j = 0; j < RPT; j++) {
                                                const int  i = tid + j * T;
                                                const bool rok = (i >= NX) && (i < NX + nrows);
                                                warp_tile_load<W>(tile, gS, warp * 32 + j * T - NX, nrows, lane);
                                                Ap_[j] = rok ? tile_row_matvec<NX>(trow, i - NX, vp) : 0.0f;
                                                prod = fmaf(p_[j], Ap_[j], prod);
                                        }
                                        {
                                                const float s = warp_tree(prod);
                                                if (lane == 0) scratchB[warp] = s;
                                        }
                                        __syncthreads();
                                        const float alpha = rho / __shfl_sync(0xffffffffu, warp_tree(scratchB[lane]), 0);
#pragma unroll
                                        for (int j = 0; j < RPT; j++) {
                                                const int i = tid + j * T;
                                                x_[j] = fmaf(alpha, p_[j], x_[j]);
                                                r_[j] = fmaf(-alpha, Ap_[j], r_[j]);
                                                if (i < n) vr[i] = r_[j];
                                        }
                                        __syncthreads();
                                        prod = 0.0f;
#pragma unroll
                                        for (int j = 0; j < RPT; j++) {
                                                const int  i = tid + j * T;
                                                const bool rok = (i >= NX) && (i < NX + nrows);
                                                warp_tile_load<W>(tile, gPc, warp * 32 + j * T - NX, nrows, lane);
                                                z_[j] = rok ? tile_row_matvec<NX>(trow, i - NX, vr) : 0.0f;
                                                prod = fmaf(r_[j], z_[j], prod);
                                        }
                                        {
                                                const float s = warp_tree(prod);
                                                if (lane == 0) scratchA[warp] = s;
                                        }
                                        __syncthreads();
                                        const float rho_new = __shfl_sync(0xffffffffu, warp_tree(scratchA[lane]), 0);
                                        if (fabsf(rho_new) < fmaf(eps, rho_init, abs_tol)) break;
                                        const float beta = rho_new / rho;
                                        rho = rho_new;
#pragma unroll
                                        for (int j = 0; j < RPT; j++) {
                                                const int i = tid + j * T;
                                                p_[j] = fmaf(beta, p_[j], z_[j]);
                                                if (i < n) vp[i] = p_[j];
                                        }
                                        __syncthreads();
                                }
#pragma unroll
                                for (int j = 0; j < RPT; j++) {
                                        const int i = tid + j * T;
                                        if (i < n) lam[i] = x_[j];
                                }
                        }
                }
                if (tid == 0) {
                        if (c.pcg_log) c.pcg_log[(size_t)c.it * c.B + b] = iters;
                        if (c.flags & F_BOOK) {
                                int cv = c.conv[b];
                                if (iters == 0) cv = 1;
                                c.conv[b] = cv;
                                atomicAdd(cv ? &c.num_solved[c.it] : &c.num_unsolved[c.it], 1u);
                        }
                }
                __syncthreads();
        }
        if (c.flags & F_DZ)
                dz_phase<NX, NU>(c, b, N, n, warp, lane, nwarps, dzbuf, c.lambda + (size_t)b * n, c.A + kb * NX2, c.Bm + kb * NX * NU, c.Qinv + kb * NX2, c.Rinv + kb * NU * NU, c.q + kb * NX,
                                 c.r + kb * NU);
}


// -----------------------------------------------------------------------------------------------------
// k_pcg_cluster: the register-resident PCG for horizons that do not fit one CTA's register file (N = 35 ... 144 for iiwa14, e.g. BASELINE
// config 4, N = 128): a THREAD-BLOCK CLUSTER of CL CTAs per solve.  CTA c keeps the rows of S and P^-1 of block rows [c*NB, (c+1)*NB) in
// registers exactly like k_pcg (one row per thread, packed FFMA2 row trees), so the whole system (588 KB at N = 128) stays on chip for all
// iterations -- k_pcg_stream re-reads it from L2 in every iteration, as the reference does from global memory (pcg.cuh:100,119).  What crosses
// CTAs goes through DISTRIBUTED SHARED MEMORY as asynchronous remote stores that complete a transaction on the RECEIVER's mbarrier
// (st.async ... mbarrier::complete_tx::bytes to mapa addresses): a CTA waits for its own inbox to fill, there is no cluster-wide barrier inside
// the iteration (measured: a cluster barrier costs about 490 cycles here and drags a gpu-scope memory barrier along; 23.6 -> 19.1 ms on config 4):
//   * the window halos: one block of Ap (or z) to each neighbour together with the dot-product terms; the neighbour then updates its copy of my
//     boundary block of r (or p) itself with the same fused multiply-add, so the vector updates need no exchange of their own;
//   * the dot products in the reference's block::dot geometry (linalg.cuh:291-327: 1024 virtual threads, virtual thread t accumulates elements
//     t and t + 1024 in that order, warp tree, tree over the 32 warp sums): every thread posts its term -- fmaf(a, b, 0) of element t, or the
//     operands (a, b) of element t + 1024 -- into the inbox of the CTA that reduces virtual warp t / 32 [inbox barrier]; the reducer warps finish
//     fmaf(a', b', first), run the shuffle tree and post the warp sum into every CTA's table [table barrier]; every thread folds the table.
//     (Posting every term to every CTA and reducing everything locally -- one hop instead of two -- was measured slower: 20.2 vs 16.2 ms in
//     k_pcg_cluster on config 4; the remote mbarrier transactions are the cost, not the hops.)
// The rows and the primal step's operands arrive through the TMA unit (cp.async.bulk + mbarrier).  P^-1 is read complete: its off-diagonal blocks
// are built beforehand by k_pcg_stream's K2 phase (once per SQP iteration).
// -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned cluster_rank()
{
        unsigned r;
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
        return r;
}
__device__ __forceinline__ unsigned mapa_cluster(unsigned laddr, unsigned rank)  // this CTA's shared address -> the same offset in CTA `rank` of the cluster
{
        unsigned r;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(laddr), "r"(rank));
        return r;
}
__device__ __forceinline__ void st_cluster(unsigned raddr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory"); }
// asynchronous store into another CTA's shared memory that completes 4 bytes of a transaction on THAT CTA's mbarrier: the receiver waits on its own
// barrier phase instead of a cluster-wide barrier (st.async ... mbarrier::complete_tx::bytes)
__device__ __forceinline__ void st_async_cluster(unsigned raddr, float v, unsigned rmbar)
{
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(raddr), "f"(v), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void cluster_sync()
{
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template<class P>
struct ClusterGeom {
        static constexpr int NX = 2 * P::NQ;
        static constexpr int NB = (P::NQ == 6) ? 40 : 32;  // block rows per CTA: NB * NX is a multiple of 32 (nq 6: 480 = 15 warps, 7: 448 = 14 warps, 8: 512 = 16 warps)
        static constexpr int T = NB * NX;
        static_assert(T % 32 == 0 && T <= 512, "whole warps, one row per thread");
        __host__ __device__ static constexpr int ctas(int N) { return (N + NB - 1) / NB; }
        // virtual warps of the reference's 1024-thread dot product reduced by one CTA, and floats of its inbox (first terms | a | b)
        __host__ __device__ static constexpr int vw_per_cta(int N) { return (32 + ctas(N) - 1) / ctas(N); }
        __host__ __device__ static constexpr bool supported(int N) { return (N + 2) * NX > 512 && (N + 2) * NX <= 2048 && ctas(N) <= 8; }
        __host__ __device__ static constexpr size_t smem_floats(int N)
        {
                const int stage = 3 * NB * NX * NX > dz_stage_floats<NX, P::NQ>(NB) ? 3 * NB * NX * NX : dz_stage_floats<NX, P::NQ>(NB);
                return 8 + 2 * (size_t)(NB + 2) * kSlot + 4 * 16 + 3 * 32 * (size_t)vw_per_cta(N) + 32 + 64 * (size_t)(T / 32) + (size_t)stage;
        }
};

template<class P>
__global__ void __launch_bounds__(ClusterGeom<P>::T, 1) k_pcg_cluster(Ctx c)
{
        using G = ClusterGeom<P>;
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ, NX2 = NX * NX, W = 3 * NX, NP = W / 2, NB = G::NB, T = G::T;
        if (stopped_before(c, c.it)) return;  // uniform over the grid: whole clusters leave together
        extern __shared__ __align__(16) float sm[];
        const int      N = c.N, n = (N + 2) * NX, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = T >> 5;
        const int      CL = G::ctas(N), VWPC = G::vw_per_cta(N);
        const unsigned rank = cluster_rank();
        const int      b = blockIdx.x / CL;
        const int      kb0 = (int)rank * NB;                          // first block row (knot) of this CTA
        const int      nbl = (N - kb0 < NB) ? (N - kb0) : NB;         // block rows it owns
        const int      nrows = nbl * NX;
        // shared memory: 4 mbarriers (rows staged, dz operands staged, dot / halo inbox complete, warp-sum table complete) | vp, vr: NB+2 slots (slot j = padded block kb0 + j: own blocks in 1..NB, halos in 0 and NB+1) |
        // halo inboxes (Ap / z from below and above) | dot inbox (first | a | b) | table of the 32 warp sums | dz scratch | stage
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm);
        float*              vp = sm + 8;
        float*              vr = vp + (NB + 2) * kSlot;
        float*              halo_in = vr + (NB + 2) * kSlot;  // [0] Ap from below, [1] Ap from above, [2] z from below, [3] z from above (16 floats each)
        float*              inbox = halo_in + 4 * 16;         // first[32 VWPC] | a[32 VWPC] | b[32 VWPC]
        float*              sums = inbox + 3 * 32 * VWPC;
        float*              dzbuf = sums + 32;
        float*              stage = dzbuf + 64 * nwarps;
        const size_t        kb = (size_t)b * N;
        const float*        gS = c.S + (kb + kb0) * 3 * NX2;
        const float*        gP = c.Pinv + (kb + kb0) * 3 * NX2;
        const bool          row_ok = tid < nrows;
        const int           brl = tid / NX;                       // local block row
        const int           i_glob = NX + kb0 * NX + tid;         // padded vector index of this thread's element
        const int           own = (brl + 1) * kSlot + tid % NX;   // its place in the slotted local vectors
        const bool          has_lo = rank > 0, has_hi = (int)rank + 1 < CL;

        // bytes that arrive in this CTA's inbox per dot product: one float per first element and two per second element of the virtual threads it
        // reduces (only real rows post), plus one boundary block from each neighbour; and 32 warp sums in its table
        unsigned in_bytes;
        {
                const int lo = (int)rank * VWPC * 32, hi = (lo + VWPC * 32 < 1024) ? lo + VWPC * 32 : 1024, r0 = NX, r1 = NX + N * NX;
                const int n1 = ((hi < r1 ? hi : r1) - (lo > r0 ? lo : r0)), n2 = ((hi + 1024 < r1 ? hi + 1024 : r1) - (lo + 1024 > r0 ? lo + 1024 : r0));
                in_bytes = 4u * (unsigned)(n1 > 0 ? n1 : 0) + 8u * (unsigned)(n2 > 0 ? n2 : 0) + 4u * NX * ((has_lo ? 1u : 0u) + (has_hi ? 1u : 0u));
        }
        if (tid == 0) {
                mbar_init(&bars[0], 1);
                mbar_init(&bars[1], 1);
                mbar_init(&bars[2], 1);
                mbar_init(&bars[3], 1);
                fence_mbar_init();
                mbar_arrive_expect_tx(&bars[2], in_bytes);  // armed for the first dot product
                mbar_arrive_expect_tx(&bars[3], 4u * 32u);
        }
        for (int i = tid; i < 2 * (NB + 2) * kSlot + 4 * 16 + 3 * 32 * VWPC + 32; i += T) vp[i] = 0.0f;  // vectors, halo inboxes, dot inbox, sums
        __syncthreads();
        cluster_sync();  // every CTA's shared memory is initialised before anyone posts into it

        float2 S2[NP], P2[NP];
        sfor<0, NP>([&](auto ic) { S2[ic] = make_float2(0.0f, 0.0f), P2[ic] = make_float2(0.0f, 0.0f); });
        if (c.flags & F_PCG) {
                // this CTA's rows of S, then of P^-1, through the TMA unit into the stage and from there into registers
                const unsigned bytes = sizeof(float) * (unsigned)nrows * W;
                if (tid == 0 && nrows > 0) {
                        mbar_arrive_expect_tx(&bars[0], bytes);
                        bulk_g2s(stage, gS, bytes, &bars[0]);
                }
                if (nrows > 0) mbar_wait(&bars[0], 0);
                if (row_ok) {
                        const float2* s2 = reinterpret_cast<const float2*>(stage + (size_t)tid * W);
                        sfor<0, NP>([&](auto ic) { S2[ic] = s2[ic]; });
                }
                fence_proxy_async();
                __syncthreads();
                if (tid == 0 && nrows > 0) {
                        mbar_arrive_expect_tx(&bars[0], bytes);
                        bulk_g2s(stage, gP, bytes, &bars[0]);
                }
                if (nrows > 0) mbar_wait(&bars[0], 1);
                if (row_ok) {
                        const float2* p2 = reinterpret_cast<const float2*>(stage + (size_t)tid * W);
                        sfor<0, NP>([&](auto ic) { P2[ic] = p2[ic]; });
                }
        }
        // the primal step's operands of this CTA's knots, prefetched into the stage while the iterations run (as in k_pcg)
        float*     sA = stage;
        float*     sB = sA + pad4(NB * NX2);
        float*     sQi = sB + pad4(NB * NX * NU);
        float*     sRi = sQi + pad4(NB * NX2);
        float*     sq = sRi + pad4(NB * NU * NU);
        float*     sr = sq + pad4(NB * NX);
        const bool bulk_dz = dz_bulk_ok<NX, NU>(nbl) && dz_bulk_ok<NX, NU>(kb0) && dz_bulk_ok<NX, NU>(N) && nbl > 0;
        if (c.flags & F_DZ) {
                fence_proxy_async();
                __syncthreads();
                if (bulk_dz && tid == 0) {
                        const unsigned bA = 4u * nbl * NX2, bB = 4u * nbl * NX * NU, bR = 4u * nbl * NU * NU, bq = 4u * nbl * NX, br_ = 4u * nbl * NU;
                        mbar_arrive_expect_tx(&bars[1], 2 * bA + bB + bR + bq + br_);
                        bulk_g2s(sA, c.A + (kb + kb0) * NX2, bA, &bars[1]);
                        bulk_g2s(sB, c.Bm + (kb + kb0) * NX * NU, bB, &bars[1]);
                        bulk_g2s(sQi, c.Qinv + (kb + kb0) * NX2, bA, &bars[1]);
                        bulk_g2s(sRi, c.Rinv + (kb + kb0) * NU * NU, bR, &bars[1]);
                        bulk_g2s(sq, c.q + (kb + kb0) * NX, bq, &bars[1]);
                        bulk_g2s(sr, c.r + (kb + kb0) * NU, br_, &bars[1]);
                }
        } else {
                __syncthreads();
        }

        const unsigned a_vp = smem_u32(vp), a_vr = smem_u32(vr), a_halo = smem_u32(halo_in), a_inbox = smem_u32(inbox), a_sums = smem_u32(sums);
        const unsigned a_mb_in = smem_u32(&bars[2]), a_mb_sum = smem_u32(&bars[3]);
        float          x_i = 0.0f;
        int            iters = 0;
        if (c.flags & F_PCG) {
                const float* gam = c.gamma + (size_t)b * n;
                float*       lam = c.lambda + (size_t)b * n;
                const float  eps = c.pcg_tol[b];
                const float  abs_tol = 1e-6f;
                const bool   skip = c.conv[b] != 0;  // pcg.cuh:29-32
                const unsigned wp = a_vp + 4u * (brl * kSlot), wr = a_vr + 4u * (brl * kSlot);  // window: slots brl .. brl+2
                const unsigned own_p = a_vp + 4u * own, own_r = a_vr + 4u * own;
                // where this thread's dot-product term goes: virtual thread t = i mod 1024 of virtual warp t / 32, reduced by CTA (t / 32) / VWPC
                const int      vt = i_glob & 1023, rc = (vt >> 5) / VWPC, q = vt - rc * VWPC * 32;
                const bool     second = i_glob >= 1024;
                const unsigned post_f = mapa_cluster(a_inbox + 4u * q, rc), post_a = mapa_cluster(a_inbox + 4u * (32 * VWPC + q), rc),
                               post_b = mapa_cluster(a_inbox + 4u * (64 * VWPC + q), rc), post_mb = mapa_cluster(a_mb_in, rc);
                unsigned       par_in = 0, par_sum = 0;  // phase parities of the two inbox barriers
                // the first NX threads post the first block to the CTA below, the next NX threads the last block to the CTA above
                const bool     send_lo = has_lo && tid < NX, send_hi = has_hi && tid >= NX && tid < 2 * NX;
                const int      hx = send_lo ? tid : tid - NX;  // element within the boundary block
                // boundary values live in other threads' registers: they go through a small exchange row
                float* xch = dzbuf;  // 2 x 16 floats: this CTA's first and last block of the freshly computed vector (Ap or z)
                auto   post_halo = [&](int kind) {
                        // kind 0: Ap, 1: z.  My first block is the "from above" halo of the CTA below; my last block the "from below" halo of the CTA above.
                        if (send_lo) st_async_cluster(mapa_cluster(a_halo + 4u * ((2 * kind + 1) * 16 + hx), rank - 1), xch[hx], mapa_cluster(a_mb_in, rank - 1));
                        if (send_hi) st_async_cluster(mapa_cluster(a_halo + 4u * ((2 * kind + 0) * 16 + hx), rank + 1), xch[16 + hx], mapa_cluster(a_mb_in, rank + 1));
                };
                auto   publish_boundary = [&](float v) {
                        if (row_ok && brl == 0) xch[tid] = v;
                        if (row_ok && brl == nbl - 1) xch[16 + tid - (nbl - 1) * NX] = v;
                };
                auto dot_post = [&](float a, float bb) {
                        if (row_ok) {
                                if (!second)
                                        st_async_cluster(post_f, fmaf(a, bb, 0.0f), post_mb);
                                else
                                        st_async_cluster(post_a, a, post_mb), st_async_cluster(post_b, bb, post_mb);
                        }
                };
                // once this CTA's inbox is complete: the reducer warps finish the virtual threads, run the warp tree and post the sums to every CTA;
                // then everybody waits for its table of the 32 warp sums.  One thread re-arms each barrier for the next dot product as soon as it has
                // seen the phase complete (nobody can post into the next phase before every CTA has finished this one).
                auto dot_reduce = [&]() {
                        mbar_wait(&bars[2], par_in);
                        par_in ^= 1u;
                        if (tid == 0) mbar_arrive_expect_tx(&bars[2], in_bytes);
                        for (int vw = warp; vw < VWPC; vw += nwarps) {
                                const int   gvw = (int)rank * VWPC + vw;
                                const float f = inbox[32 * vw + lane], a = inbox[32 * VWPC + 32 * vw + lane], bb = inbox[64 * VWPC + 32 * vw + lane];
                                const float sum = __shfl_sync(0xffffffffu, warp_tree(fmaf(a, bb, f)), 0);
                                if (gvw < 32 && lane < CL) st_async_cluster(mapa_cluster(a_sums + 4u * gvw, lane), sum, mapa_cluster(a_mb_sum, lane));  // lane l posts to CTA l
                        }
                        mbar_wait(&bars[3], par_sum);
                        par_sum ^= 1u;
                        if (tid == 0) mbar_arrive_expect_tx(&bars[3], 4u * 32u);
                };
                if (!skip) {
                        x_i = row_ok ? lam[i_glob] : 0.0f;
                        if (row_ok) sts32(own_p, x_i);  // vp temporarily holds x for r = gamma - S x; its halos come straight from global memory
                        if (tid < NX) {
                                vp[tid] = lam[kb0 * NX + tid];  // padded block kb0 (the block below; zeros for the first CTA)
                                vp[(nbl + 1) * kSlot + tid] = lam[(kb0 + nbl + 1) * NX + tid];
                        }
                        __syncthreads();
                        float r_i = 0.0f, p_i = 0.0f, z_i = 0.0f;
                        {
                                const float sx = row_ok ? matvec_packed<NX>(S2, wp) : 0.0f;
                                r_i = row_ok ? (gam[i_glob] - sx) : 0.0f;
                                if (row_ok) sts32(own_r, r_i);
                                __syncthreads();
                                // r halos: posted straight into the neighbours' vr halo slots
                                if (send_lo) st_cluster(mapa_cluster(a_vr + 4u * ((NB + 1) * kSlot + hx), rank - 1) , vr[kSlot + hx]);
                                if (send_hi) st_cluster(mapa_cluster(a_vr + 4u * hx, rank + 1), vr[nbl * kSlot + hx]);
                        }
                        cluster_sync();
                        // (only the last CTA can be partial, and it has no neighbour above: a CTA that receives an upper halo is full, slot NB+1)
                        z_i = row_ok ? matvec_packed<NX>(P2, wr) : 0.0f;
                        p_i = z_i;
                        __syncthreads();  // every row is done reading vp as x
                        if (row_ok) sts32(own_p, p_i);
                        publish_boundary(z_i);
                        dot_post(r_i, z_i);
                        __syncthreads();
                        post_halo(1);  // p = z: the neighbours' boundary blocks arrive like every later z halo
                        dot_reduce();
                        if (has_lo && tid < NX) vp[tid] = halo_in[2 * 16 + tid];
                        if (has_hi && tid >= NX && tid < 2 * NX) vp[(NB + 1) * kSlot + tid - NX] = halo_in[3 * 16 + tid - NX];
                        float rho = tree32(a_sums);
                        __syncthreads();
                        if (!(fabsf(rho) < abs_tol)) {
                                const float rho_init = fabsf(rho);
                                for (int itn = 0; itn < c.max_pcg; itn++) {
                                        iters++;
                                        const float Ap_i = row_ok ? matvec_packed<NX>(S2, wp) : 0.0f;
                                        publish_boundary(Ap_i);
                                        dot_post(p_i, Ap_i);
                                        __syncthreads();
                                        post_halo(0);
                                        dot_reduce();
                                        const float alpha = rho / tree32(a_sums);
                                        x_i = fmaf(alpha, p_i, x_i);
                                        r_i = fmaf(-alpha, Ap_i, r_i);
                                        if (row_ok) sts32(own_r, r_i);
                                        // my copies of the neighbours' boundary blocks of r, updated with the same fused multiply-add they use
                                        if (has_lo && tid < NX) vr[tid] = fmaf(-alpha, halo_in[0 * 16 + tid], vr[tid]);
                                        if (has_hi && tid >= NX && tid < 2 * NX) vr[(NB + 1) * kSlot + tid - NX] = fmaf(-alpha, halo_in[1 * 16 + tid - NX], vr[(NB + 1) * kSlot + tid - NX]);
                                        __syncthreads();
                                        z_i = row_ok ? matvec_packed<NX>(P2, wr) : 0.0f;
                                        publish_boundary(z_i);
                                        dot_post(r_i, z_i);
                                        __syncthreads();
                                        post_halo(1);
                                        dot_reduce();
                                        const float rho_new = tree32(a_sums);
#ifdef GATO_CLUSTER_EXTRA_SYNC
                                        for (int e_ = 0; e_ < GATO_CLUSTER_EXTRA_SYNC; e_++) cluster_sync();  // measurement only: the cost of one cluster barrier
#endif
                                        if (fabsf(rho_new) < fmaf(eps, rho_init, abs_tol)) break;
                                        const float beta = rho_new / rho;
                                        rho = rho_new;
                                        p_i = fmaf(beta, p_i, z_i);
                                        if (row_ok) sts32(own_p, p_i);
                                        if (has_lo && tid < NX) vp[tid] = fmaf(beta, vp[tid], halo_in[2 * 16 + tid]);
                                        if (has_hi && tid >= NX && tid < 2 * NX) vp[(NB + 1) * kSlot + tid - NX] = fmaf(beta, vp[(NB + 1) * kSlot + tid - NX], halo_in[3 * 16 + tid - NX]);
                                        __syncthreads();
                                }
                                if (row_ok) lam[i_glob] = x_i;
                        }
                }
                if (tid == 0 && rank == 0) {
                        if (c.pcg_log) c.pcg_log[(size_t)c.it * c.B + b] = iters;
                        if (c.flags & F_BOOK) {
                                int cv = c.conv[b];
                                if (iters == 0) cv = 1;
                                c.conv[b] = cv;
                                atomicAdd(cv ? &c.num_solved[c.it] : &c.num_unsolved[c.it], 1u);
                        }
                }
        }
        // lambda of the neighbouring block above is needed by the primal step: make this CTA's lambda visible, then everybody reads global memory
        __threadfence();
        cluster_sync();
        if (c.flags & F_DZ) {
                const float* lamg = c.lambda + (size_t)b * n;
                for (int i = tid; i < (nbl + 2) * NX; i += T) vp[(i / NX) * kSlot + i % NX] = lamg[kb0 * NX + i];  // padded blocks kb0 .. kb0+nbl+1
                if (bulk_dz) mbar_wait(&bars[1], 0);
                __syncthreads();
                if (nbl > 0) {
                        if (bulk_dz)
                                dz_phase<NX, NU, kSlot>(c, b, N, n, warp, lane, nwarps, dzbuf, vp, sA, sB, sQi, sRi, sq, sr, kb0, kb0 + nbl);
                        else
                                dz_phase<NX, NU, kSlot>(c, b, N, n, warp, lane, nwarps, dzbuf, vp, c.A + (kb + kb0) * NX2, c.Bm + (kb + kb0) * NX * NU, c.Qinv + (kb + kb0) * NX2,
                                                        c.Rinv + (kb + kb0) * NU * NU, c.q + (kb + kb0) * NX, c.r + (kb + kb0) * NU, kb0, kb0 + nbl);
                }
        }
        cluster_sync();  // nobody leaves while a neighbour may still post into its shared memory
}

}  // namespace gato
