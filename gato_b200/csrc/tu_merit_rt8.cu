// run-time models with nq = 8 (rt_model.h): table-driven merit / line search / sim_forward / ee_pos
#define GATO_RT_TU 1
#define GATO_TU_PLANT RtPlant<8>
#include "tu_merit.cuh"
