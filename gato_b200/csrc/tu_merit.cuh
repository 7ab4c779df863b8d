// body of the merit / line-search / sim_forward translation units; GATO_TU_PLANT selects the plant
#include "launchers.h"
#include "bsqp_merit_kernels.cuh"
namespace gato {
namespace {
#ifndef GATO_RT_MERIT_THREADS
#define GATO_RT_MERIT_THREADS 256
#endif
inline int merit_threads(int na, int N)
{
        const int cap = na == 1 ? 128 : (is_rt_plant<GATO_TU_PLANT> ? GATO_RT_MERIT_THREADS : 256);
        int       t = na * N;
        t = (t + 31) / 32 * 32;
        return t > cap ? cap : t;
}
}  // namespace
template<>
bool enqueue_merit<GATO_TU_PLANT>(const Ctx& c, int na, cudaStream_t st)
{
        const size_t smem = sizeof(float) * (size_t)(na * c.N + na);
        // small batches (one CTA per SM at most) with room for two threads per (alpha, knot) in one block: the split kernel
        if (na == kNumAlphas && c.B <= c.sms && 2 * na * c.N <= 512) {
                const int threads = (2 * na * c.N + 31) / 32 * 32;
                Ctx       k = c;
                k.flags &= ~F_OVERLAP;
                k_merit_ls<GATO_TU_PLANT, kNumAlphas, true><<<c.B, threads, smem + sizeof(float) * (size_t)(na * c.N), st>>>(k);
                return false;
        }
        if (na == 1) {
                k_merit_ls<GATO_TU_PLANT, 1><<<c.B, merit_threads(1, c.N), smem, st>>>(c);
                return false;
        }
        if (c.flags & F_OVERLAP) {
                // programmatic dependent launch: the grid may start once every CTA of the preceding k_pcg grid has started (k_pcg triggers at its
                // top), i.e. during k_pcg's last, partial wave; per-solve ordering comes from the hand-over flags
                cudaLaunchConfig_t  cfg{};
                cudaLaunchAttribute at[1];
                cfg.gridDim = dim3((unsigned)c.B), cfg.blockDim = dim3((unsigned)merit_threads(kNumAlphas, c.N)), cfg.dynamicSmemBytes = smem, cfg.stream = st;
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at, cfg.numAttrs = 1;
                cudaLaunchKernelEx(&cfg, k_merit_ls<GATO_TU_PLANT, kNumAlphas, false>, c);
                return true;
        }
        k_merit_ls<GATO_TU_PLANT, kNumAlphas><<<c.B, merit_threads(kNumAlphas, c.N), smem, st>>>(c);
        return false;
}
template<>
void enqueue_ee_pos<GATO_TU_PLANT>(int n, const float* q, float* ee, int model_slot, cudaStream_t st)
{
        k_ee_pos<GATO_TU_PLANT><<<(n + 63) / 64, 64, 0, st>>>(n, q, ee, model_slot);
}
template<>
void enqueue_sim_forward<GATO_TU_PLANT>(int B, float* xkp1, const float* xk, const float* uk, const float* fext, float dt, int model_slot, cudaStream_t st)
{
        const int T = 64, G = (B + T - 1) / T;
        k_sim_forward<GATO_TU_PLANT><<<G, T, 0, st>>>(B, xkp1, xk, uk, fext, dt, model_slot);
}
#ifdef GATO_RT_TU
template<>
cudaError_t upload_rt_model_merit<GATO_TU_PLANT>(int slot, const RtModel& m)
{
        return cudaMemcpyToSymbol(g_rt_models, &m, sizeof(RtModel), sizeof(RtModel) * (size_t)slot, cudaMemcpyHostToDevice);
}
#endif
}  // namespace gato
