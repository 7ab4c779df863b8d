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
namespace {
// CL CTAs per solve as one thread-block cluster
void launch_pcg_cluster(const Ctx& c, cudaStream_t st)
{
        using P = GATO_TU_PLANT;
        using G = ClusterGeom<P>;
        const int           cl = G::ctas(c.N);
        cudaLaunchConfig_t  cfg{};
        cudaLaunchAttribute at[1];
        cfg.gridDim = dim3((unsigned)(c.B * cl)), cfg.blockDim = dim3(G::T), cfg.dynamicSmemBytes = sizeof(float) * G::smem_floats(c.N), cfg.stream = st;
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)cl, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
        cfg.attrs = at, cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k_pcg_cluster<P>, c);
}
}  // namespace
template<>
bool pcg_cluster_supported<GATO_TU_PLANT>(int N)
{
        return ClusterGeom<GATO_TU_PLANT>::supported(N);
}
template<>
int enqueue_pcg<GATO_TU_PLANT>(const Ctx& c, int rpt, int threads, size_t smem, bool cluster, cudaStream_t st)
{
        using P = GATO_TU_PLANT;
        if (cluster && rpt > 0) {
                // off-diagonal blocks of P^-1 by the streaming kernel's K2 phase (complete rows in global memory), everything else by the cluster kernel
                int n = 0;
                if (c.flags & F_K2) {
                        Ctx k2 = c;
                        k2.flags = c.flags & (F_K2 | F_WRITE_P | F_CHECK_STOP);
                        n += enqueue_pcg<P>(k2, rpt, threads, smem, false, st);
                }
                if (c.flags & (F_PCG | F_DZ)) {
                        Ctx k = c;
                        k.flags = c.flags & ~(F_K2 | F_WRITE_P);
                        launch_pcg_cluster(k, st);
                        n++;
                }
                return n;
        }
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
        return 1;
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
        if (e == cudaSuccess) e = opt_in_max_smem(k_pcg_cluster<P>, device);
        return e;
}
}  // namespace gato
