// body of the k_kkt translation units; GATO_TU_PLANT selects the plant
#include "launchers.h"
#include "bsqp_kkt_kernels.cuh"
namespace gato {
template<>
void enqueue_kkt<GATO_TU_PLANT>(const Ctx& c, cudaStream_t st)
{
        const int items = c.B * c.N;  // kind 0 has B*N items, the dynamics kinds have B*(N-1)
        const int warps = (items + 31) / 32;
        constexpr int kFineKinds = 2 + 4 * GATO_TU_PLANT::NQ / 2;  // 2 + 2 nq
        // small grids (the MPC regime): one column of the linearisation per thread, as long as all its warps are resident at once
        // (8 warps of 255 registers per SM)
        if (warps * kFineKinds <= c.sms * 8)
                k_kkt_fine<GATO_TU_PLANT><<<dim3(warps, kFineKinds), 32, 0, st>>>(c);
        else
                k_kkt<GATO_TU_PLANT><<<dim3(warps, 3), 32, 0, st>>>(c);
}
#ifdef GATO_RT_TU
template<>
cudaError_t upload_rt_model_kkt<GATO_TU_PLANT>(int slot, const RtModel& m)
{
        return cudaMemcpyToSymbol(g_rt_models, &m, sizeof(RtModel), sizeof(RtModel) * (size_t)slot, cudaMemcpyHostToDevice);
}
#endif
}  // namespace gato
