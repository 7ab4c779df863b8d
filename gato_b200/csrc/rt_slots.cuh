// Where the kernels get their Items<P> from.  Compiled plants: an empty object.  Run-time models (translation units built with GATO_RT_TU):
// a reference into this translation unit's constant-memory copy of the model tables, slot = plant id - GATO_PLANT_MODEL0; the host side
// fills the slot (upload_rt_model_* in tu_kkt.cuh / tu_merit.cuh) before the first launch that uses it.
#pragma once
#include "bsqp_ctx.cuh"
#include "items_rt.cuh"

namespace gato {

#ifdef GATO_RT_TU
static __constant__ RtModel g_rt_models[kRtSlots];
#endif

template<class P>
__device__ __forceinline__ Items<P> make_items_slot(int slot)
{
        if constexpr (is_rt_plant<P>) {
#ifdef GATO_RT_TU
                return Items<P>(g_rt_models[slot]);
#else
                static_assert(!is_rt_plant<P>, "run-time models are instantiated in the GATO_RT_TU translation units only");
#endif
        } else {
                return Items<P>();
        }
}
template<class P>
__device__ __forceinline__ Items<P> make_items(const Ctx& c)
{
        return make_items_slot<P>(c.model_slot);
}

}  // namespace gato
