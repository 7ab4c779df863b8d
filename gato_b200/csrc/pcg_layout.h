// Shared-memory layout of k_pcg, shared between the kernel (bsqp_pcg_kernels.cuh) and the host code that sizes its launch.
#pragma once
#include <cstddef>

namespace gato {

constexpr int pad4(int x) { return (x + 3) / 4 * 4; }

// floats between consecutive blocks of the shared PCG vectors: a multiple of 4 (every block 16-byte aligned: a row's window is 3 groups of
// LDS.128) such that the blocks of four consecutive block rows -- what one warp touches -- start 20 banks apart and never share a bank
// (16 would put blocks k and k+2 on the same banks).
constexpr int kSlot = 20;

// Shared-memory layout of k_pcg (floats unless noted); the host sizes the launch with pcg_smem_floats().
//   2 mbarriers (4 floats) | vp, vr: (N+2) blocks of kSlot floats each | scratchA(16) scratchB(16) | products (32 per warp) | dz scratch (64 per warp) |
//   stage A: the solve's S rows as the TMA unit delivers them (N*3*NX^2); once the rows are in registers: the K2 scratch ((N-1)*NX^2), then
//            the prefetched operands of the primal step (A, B, Q^-1, R^-1, q, r)
//   stage B: the packed main blocks of P^-1 (N*NX^2); after K2: the product blocks the right-hand P^-1 blocks are read from
template<int NX, int NU>
constexpr int dz_stage_floats(int N)
{
        return pad4(N * NX * NX) * 2 + pad4(N * NX * NU) + pad4(N * NU * NU) + pad4(N * NX) + pad4(N * NU);
}
template<int NX, int NU>
constexpr size_t pcg_smem_floats(int N, int nwarps)
{
        const int stageA = 3 * N * NX * NX > dz_stage_floats<NX, NU>(N) ? 3 * N * NX * NX : dz_stage_floats<NX, NU>(N);
        return 4 + 2 * (size_t)(N + 2) * kSlot + 32 + (32 + 64) * (size_t)nwarps + (size_t)stageA + (size_t)N * NX * NX;
}
// all six per-knot operand arrays of one solve can be moved by bulk copies (16-byte granularity) when N whole blocks of each are a
// multiple of 16 bytes: then every solve's first block is 16-byte aligned as well
template<int NX, int NU>
__host__ __device__ constexpr bool dz_bulk_ok(int N)
{
        return (N * NX * NX) % 4 == 0 && (N * NX * NU) % 4 == 0 && (N * NU * NU) % 4 == 0 && (N * NX) % 4 == 0 && (N * NU) % 4 == 0;
}

}  // namespace gato
