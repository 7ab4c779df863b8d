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

#include "costs.h"

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
        int   sms;  // multiprocessors of the device (launch heuristics)
        int   model_slot;  // run-time models: constant-memory slot of the robot's tables (rt_slots.cuh); unused by the compiled plants
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
        unsigned*    num_unsolved;  // [max_sqp_iters] #solves whose PCG of iteration i finished unflagged (overlapped launches decide the exit test early)
        unsigned*    pcg_done;    // [max_sqp_iters][B] set by k_pcg when everything solve b needs from iteration i's PCG launch is in memory
        int*         pcg_log;     // [max_sqp_iters][B]
        unsigned *   kkt_qmax, *kkt_cmax;  // optional [max_sqp_iters][B]: max |KKT residual q|, max |c| per solve as float bits (bsqp.cuh:149-150); may be null
        float *      ls_merit_log, *ls_step_log;  // [max_sqp_iters][B]
};

// F_OVERLAP: this launch may run beside the PCG launch of the same iteration (programmatic dependent launch): it waits per solve on pcg_done
// and decides the early-exit test from the running counters instead of relying on the stream order
enum : int { F_K2 = 1, F_PCG = 2, F_DZ = 4, F_WRITE_P = 8, F_MERIT = 16, F_LS = 32, F_BOOK = 64, F_CHECK_STOP = 128, F_ZERO_DZ = 256, F_OVERLAP = 512 };

// true when an iteration j < upto already satisfied the early-exit test of bsqp.cuh:165
__device__ __forceinline__ bool stopped_before(const Ctx& c, int upto)
{
        if (!(c.flags & F_CHECK_STOP)) return false;
        bool s = false;
        for (int j = 0; j < upto; j++) s |= ((float)c.num_solved[j] >= c.thresh);
        return s;
}

// ---- per-solve hand-over between overlapped launches ---------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p)
{
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// lets the next kernel in the stream start (if it was launched with programmatic stream serialization) once every CTA of this grid has
// passed this point or exited; a no-op otherwise
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// spin (one thread) until *flag is non-zero; gives up after about two seconds so that a protocol error cannot hang the GPU
__device__ __forceinline__ bool wait_flag(const unsigned* flag)
{
        const long long t0 = clock64();
        while (ld_acquire_gpu(flag) == 0u) {
                __nanosleep(200);
                if (clock64() - t0 > 4000000000LL) return false;
        }
        return true;
}
// The early-exit test of bsqp.cuh:165 for iteration `it` while its PCG launch may still be running: true / false as soon as the running counters
// settle it (stop once enough solves are flagged; no stop once so many are unflagged that the threshold cannot be met any more).
__device__ __forceinline__ bool stop_decided_early(const Ctx& c, int it)
{
        const long long t0 = clock64();
        for (;;) {
                const unsigned s = ld_acquire_gpu(c.num_solved + it), u = ld_acquire_gpu(c.num_unsolved + it);
                if ((float)s >= c.thresh) return true;
                if ((float)((unsigned)c.B - u) < c.thresh) return false;
                __nanosleep(500);
                if (clock64() - t0 > 4000000000LL) return true;
        }
}

}  // namespace gato
