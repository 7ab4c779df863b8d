// Kernel launchers shared between the host translation unit (gato_b200.cu) and the kernel translation units
// (tu_*.cu).  The kernels are heavy templates (fully unrolled rigid-body dynamics); each (group, plant) is compiled
// in its own translation unit so that nvcc --threads builds them in parallel.  The host TU only sees declarations.
#pragma once
#include <cuda_runtime.h>

#include "bsqp_ctx.cuh"
#include "pcg_layout.h"
#include "rbd_rt.cuh"  // plant tags (Iiwa14, Indy7, RtPlant<NQ>), RtModel

namespace gato {

#ifndef GATO_SCHUR_WARPS
#define GATO_SCHUR_WARPS 4
#endif
constexpr int kSchurWarps = GATO_SCHUR_WARPS;

template<class P>
void enqueue_kkt(const Ctx& c, cudaStream_t st);
template<class P>
void enqueue_schur(const Ctx& c, size_t smem, cudaStream_t st);
// rpt = 0: register-resident k_pcg with `threads` threads; rpt = 1..4: k_pcg_stream<rpt> with 1024 threads, or -- cluster = true -- the
// cluster kernel k_pcg_cluster (after k_pcg_stream's K2 phase when F_K2 is set).  Returns the number of kernels launched.
template<class P>
int enqueue_pcg(const Ctx& c, int rpt, int threads, size_t smem, bool cluster, cudaStream_t st);
// horizons the cluster kernel serves: too long for one CTA's registers, at most 2048 padded vector entries and 8 CTAs per solve
template<class P>
bool pcg_cluster_supported(int N);
// opt in to the device's maximum dynamic shared memory for every linear-algebra kernel (the attribute is global per kernel and device)
template<class P>
cudaError_t configure_linalg(int device);
template<class P>
size_t schur_smem_bytes();
// num_alphas = 1 (initial / final merit) or kNumAlphas (merit + line search).  c.flags & F_OVERLAP: launched with programmatic stream
// serialization so that its CTAs may start while the preceding k_pcg launch drains (they wait per solve on the hand-over flags).
// Returns false if the launch did not use the overlapped form (small batches use the split kernel).
template<class P>
bool enqueue_merit(const Ctx& c, int num_alphas, cudaStream_t st);
// end-effector position (forward kinematics) of n joint configurations: q[n][nq] -> ee[n][3]
template<class P>
void enqueue_ee_pos(int n, const float* q, float* ee, int model_slot, cudaStream_t st);
template<class P>
void enqueue_sim_forward(int B, float* xkp1, const float* xk, const float* uk, const float* fext, float dt, int model_slot, cudaStream_t st);
// run-time models: copy the tables into constant-memory slot `slot` of the current device, once per kernel group and nq (every translation
// unit holds its own copy of the slots)
template<class P>
cudaError_t upload_rt_model_kkt(int slot, const RtModel& m);
template<class P>
cudaError_t upload_rt_model_merit(int slot, const RtModel& m);

}  // namespace gato
