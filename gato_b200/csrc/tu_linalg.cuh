// body of the Schur / PCG translation units; GATO_TU_PLANT selects the plant
#include "launchers.h"
#include "bsqp_linalg_kernels.cuh"
#include "bsqp_pcg_kernels.cuh"
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
namespace {
// cudaFuncAttributeMaxDynamicSharedMemorySize is global per kernel and device: it is always raised to the device's opt-in maximum (minus
// the kernel's static shared memory), never to what one solver happens to need, so that solvers with different horizons can coexist.
template<class K>
cudaError_t opt_in_max_smem(K kernel, int device)
{
        int         maxopt = 0;
        cudaError_t e = cudaDeviceGetAttribute(&maxopt, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
        if (e != cudaSuccess) return e;
        cudaFuncAttributes fa{};
        e = cudaFuncGetAttributes(&fa, kernel);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxopt - (int)fa.sharedSizeBytes);
}
}  // namespace
template<>
cudaError_t configure_linalg<GATO_TU_PLANT>(int device)
{
        using P = GATO_TU_PLANT;
        cudaError_t e = opt_in_max_smem(k_schur<P>, device);
        if (e == cudaSuccess) e = opt_in_max_smem(k_pcg<P, 480>, device);
        if (e == cudaSuccess) e = opt_in_max_smem(k_pcg<P, 512>, device);
        if (e == cudaSuccess) e = opt_in_max_smem(k_pcg_stream<P, 1>, device);
        if (e == cudaSuccess) e = opt_in_max_smem(k_pcg_stream<P, 2>, device);
        if (e == cudaSuccess) e = opt_in_max_smem(k_pcg_stream<P, 3>, device);
        if (e == cudaSuccess) e = opt_in_max_smem(k_pcg_stream<P, 4>, device);
        return e;
}
}  // namespace gato
