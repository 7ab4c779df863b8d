#define GATO_TU_PLANT Indy7
#include "tu_merit.cuh"
