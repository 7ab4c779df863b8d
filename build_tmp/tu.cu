#include "../gato_b200/csrc/bsqp_kernels.cuh"
namespace gato { template __global__ void k_schur<Iiwa14>(Ctx); template __global__ void k_schur<Indy7>(Ctx);}
