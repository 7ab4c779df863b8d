// Shim for gato/utils/cuda.cuh:7-65 (gpuErrchk, L2 persisting-cache helpers).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#ifndef NDEBUG
inline void gpuAssert(cudaError_t code, const char* file, int line, bool abort = true)
{
        if (code != cudaSuccess) {
                fprintf(stderr, "GPUassert: %s %s %d\n", cudaGetErrorString(code), file, line);
                if (abort) exit(code);
        }
}
#define gpuErrchk(ans) { gpuAssert((ans), __FILE__, __LINE__); }
#else
#define gpuErrchk(ans) ans
#endif
// The B200 solver keeps its working set on chip / in the 126 MB L2 by construction; the persisting-L2 carve-out the
// reference asks for (bindings.cu:45) is accepted and ignored.
inline void setL2PersistingAccess(float, bool = false) {}
inline void resetL2PersistingAccess() {}
