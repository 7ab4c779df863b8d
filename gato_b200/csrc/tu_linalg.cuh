// body of the Schur / PCG translation units; GATO_TU_PLANT selects the plant
#include "launchers.h"
namespace gato {
template<>
size_t schur_smem_bytes<GATO_TU_PLANT>()
{
        return sizeof(SchurSmem<2 * GATO_TU_PLANT::NQ, GATO_TU_PLANT::NQ>) * kSchurWarps;
}
template<>
void enqueue_schur<GATO_TU_PLANT>(const Ctx& c, size_t smem, cudaStream_t st)
{
        const int items = c.B * ((c.N + 1) / 2);  // one warp per pair of knots
        k_schur<GATO_TU_PLANT><<<(items + kSchurWarps - 1) / kSchurWarps, kSchurWarps * 32, smem, st>>>(c);
}
template<>
void enqueue_pcg<GATO_TU_PLANT>(const Ctx& c, int rpt, int threads, size_t smem, cudaStream_t st)
{
        using P = GATO_TU_PLANT;
        switch (rpt) {
                case 0:
                        if (threads <= 480)
                                k_pcg<P, 480><<<c.B, threads, smem, st>>>(c);
                        else
                                k_pcg<P, 512><<<c.B, threads, smem, st>>>(c);
                        break;
                case 1: k_pcg_stream<P, 1><<<c.B, 1024, smem, st>>>(c); break;
                case 2: k_pcg_stream<P, 2><<<c.B, 1024, smem, st>>>(c); break;
                case 3: k_pcg_stream<P, 3><<<c.B, 1024, smem, st>>>(c); break;
                default: k_pcg_stream<P, 4><<<c.B, 1024, smem, st>>>(c); break;
        }
}
template<>
cudaError_t configure_linalg<GATO_TU_PLANT>(int rpt, size_t smem_pcg, size_t smem_schur)
{
        using P = GATO_TU_PLANT;
        cudaError_t e = cudaFuncSetAttribute(k_schur<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_schur);
        if (e != cudaSuccess) return e;
        switch (rpt) {
                case 0:
                        e = cudaFuncSetAttribute(k_pcg<P, 480>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pcg);
                        if (e != cudaSuccess) return e;
                        return cudaFuncSetAttribute(k_pcg<P, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pcg);
                case 1: return cudaFuncSetAttribute(k_pcg_stream<P, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pcg);
                case 2: return cudaFuncSetAttribute(k_pcg_stream<P, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pcg);
                case 3: return cudaFuncSetAttribute(k_pcg_stream<P, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pcg);
                default: return cudaFuncSetAttribute(k_pcg_stream<P, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pcg);
        }
}
}  // namespace gato
