#include "../gato_b200/csrc/bsqp_kernels.cuh"
namespace gato { template __global__ void k_pcg2<Iiwa14>(Ctx); }
