// compile-time loops: sfor<I, N>(f) calls f(integral_constant<int, i>) for i = I .. N-1; sfor_down runs N-1 .. I
#pragma once
#include <type_traits>

#if defined(__CUDACC__)
#define GATO_HD __host__ __device__ __forceinline__
#else
#define GATO_HD inline __attribute__((always_inline))
#endif

namespace gato {

template<int I, int N, class F>
GATO_HD void sfor(F&& f)
{
        if constexpr (I < N) {
                f(std::integral_constant<int, I>{});
                sfor<I + 1, N>(f);
        }
}
template<int I, int N, class F>
GATO_HD void sfor_down(F&& f)  // I = N-1 ... 0
{
        if constexpr (N > I) {
                f(std::integral_constant<int, N - 1>{});
                sfor_down<I, N - 1>(f);
        }
}

}  // namespace gato
