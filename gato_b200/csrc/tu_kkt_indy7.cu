#define GATO_TU_PLANT Indy7
#include "tu_kkt.cuh"
