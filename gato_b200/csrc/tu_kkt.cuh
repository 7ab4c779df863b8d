// body of the k_kkt translation units; GATO_TU_PLANT selects the plant
#include "launchers.h"
namespace gato {
template<>
void enqueue_kkt<GATO_TU_PLANT>(const Ctx& c, cudaStream_t st)
{
        const int items = c.B * c.N;  // kind 0 has B*N items, kinds 1 and 2 have B*(N-1)
        k_kkt<GATO_TU_PLANT><<<dim3((items + 31) / 32, 3), 32, 0, st>>>(c);
}
}  // namespace gato
