// Kernel launchers shared between the host translation unit (gato_b200.cu) and the kernel translation units
// (tu_*.cu).  The kernels are heavy templates (fully unrolled rigid-body dynamics); each (group, plant) is compiled
// in its own translation unit so that nvcc --threads builds them in parallel.  The host TU only sees declarations.
#pragma once
#include <cuda_runtime.h>

#include "bsqp_ctx.cuh"
#include "pcg_layout.h"
#include "rbd.cuh"  // plant tags (Iiwa14, Indy7)

namespace gato {

#ifndef GATO_SCHUR_WARPS
#define GATO_SCHUR_WARPS 4
#endif
constexpr int kSchurWarps = GATO_SCHUR_WARPS;

template<class P>
void enqueue_kkt(const Ctx& c, cudaStream_t st);
template<class P>
void enqueue_schur(const Ctx& c, size_t smem, cudaStream_t st);
// rpt = 0: register-resident k_pcg with `threads` threads; rpt = 1..4: k_pcg_stream<rpt> with 1024 threads
template<class P>
void enqueue_pcg(const Ctx& c, int rpt, int threads, size_t smem, cudaStream_t st);
// opt in to the device's maximum dynamic shared memory for every linear-algebra kernel (the attribute is global per kernel and device)
template<class P>
cudaError_t configure_linalg(int device);
template<class P>
size_t schur_smem_bytes();
// num_alphas = 1 (initial / final merit) or kNumAlphas (merit + line search)
template<class P>
void enqueue_merit(const Ctx& c, int num_alphas, cudaStream_t st);
// end-effector position (forward kinematics) of n joint configurations: q[n][nq] -> ee[n][3]
template<class P>
void enqueue_ee_pos(int n, const float* q, float* ee, cudaStream_t st);
template<class P>
void enqueue_sim_forward(int B, float* xkp1, const float* xk, const float* uk, const float* fext, float dt, cudaStream_t st);

}  // namespace gato
