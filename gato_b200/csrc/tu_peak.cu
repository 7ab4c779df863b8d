// FP32 CUDA-core peak of the device this library runs on, measured ("measure an FMA micro-benchmark on the
// box"): the denominator of the FP32 roofline fractions bench.py reports.  Two variants: scalar FFMA (one fused multiply-add per lane
// and instruction) and Blackwell's packed FFMA2 (fma.rn.f32x2: two per lane and instruction), which is what the linear-algebra kernels
// of this library issue.  MEASURED_PEAKS.json (driver-written) has no FP32 CUDA-core number.
#include <cuda_runtime.h>

#include "../../include/gato_b200.h"

namespace {

template<bool PACKED>
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float a, float b)
{
        // 8 independent accumulator pairs per thread: enough ILP to cover the 4-cycle pipe with 8 resident warps per scheduler
        float2 acc[8];
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = make_float2((float)threadIdx.x + i, (float)i);
        const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
        for (int it = 0; it < iters; it++) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                        if constexpr (PACKED) {
                                acc[i] = __ffma2_rn(acc[i], a2, b2);
                        } else {
                                acc[i].x = fmaf(acc[i].x, a, b);
                                acc[i].y = fmaf(acc[i].y, a, b);
                        }
                }
        }
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) s += acc[i].x + acc[i].y;
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int gato_measure_fp32_peak(int device, int packed, double* tflops)
{
        if (!tflops) return GATO_ERR_ARG;
        if (cudaSetDevice(device) != cudaSuccess) return GATO_ERR_CUDA;
        cudaDeviceProp prop{};
        if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GATO_ERR_CUDA;
        const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 8192;
        float*    out = nullptr;
        if (cudaMalloc((void**)&out, sizeof(float) * blocks * threads) != cudaSuccess) return GATO_ERR_CUDA;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0), cudaEventCreate(&e1);
        double best = 0.0;
        for (int rep = 0; rep < 6; rep++) {  // the first repetitions warm the clocks up
                cudaEventRecord(e0);
                if (packed)
                        k_fma_peak<true><<<blocks, threads>>>(out, iters, 1.0000001f, 1e-7f);
                else
                        k_fma_peak<false><<<blocks, threads>>>(out, iters, 1.0000001f, 1e-7f);
                cudaEventRecord(e1);
                if (cudaEventSynchronize(e1) != cudaSuccess) break;
                float ms = 0.0f;
                cudaEventElapsedTime(&ms, e0, e1);
                const double flop = 2.0 * 16.0 * (double)iters * (double)blocks * threads;  // 16 fused multiply-adds per thread and iteration
                if (rep >= 2 && ms > 0.0f) best = best > flop / (ms * 1e-3) / 1e12 ? best : flop / (ms * 1e-3) / 1e12;
        }
        const cudaError_t e = cudaGetLastError();
        cudaEventDestroy(e0), cudaEventDestroy(e1);
        cudaFree(out);
        if (e != cudaSuccess || best == 0.0) return GATO_ERR_CUDA;
        *tflops = best;
        return GATO_OK;
}
