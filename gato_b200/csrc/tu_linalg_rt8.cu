// linear-algebra kernels (k_schur, k_pcg, k_pcg_stream, k_pcg_cluster) for nx = 16: no compiled plant has 8 joints, the run-time models instantiate them
#define GATO_TU_PLANT RtPlant<8>
#include "tu_linalg.cuh"
