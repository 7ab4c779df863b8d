#define GATO_TU_PLANT Iiwa14
#include "tu_kkt.cuh"
