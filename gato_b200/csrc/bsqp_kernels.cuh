// All kernels of the BSQP path (see bsqp_ctx.cuh for the kernel map).  The translation units include only the group they instantiate.
#pragma once
#include "bsqp_kkt_kernels.cuh"
#include "bsqp_linalg_kernels.cuh"  // k_schur
#include "bsqp_pcg_kernels.cuh"     // k_pcg, k_pcg_stream
#include "bsqp_merit_kernels.cuh"
