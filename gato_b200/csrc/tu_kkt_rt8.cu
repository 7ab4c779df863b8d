// run-time models with nq = 8 (rt_model.h): table-driven k_kkt / k_kkt_fine
#define GATO_RT_TU 1
#define GATO_TU_PLANT RtPlant<8>
#include "tu_kkt.cuh"
