// Run-time models use the linear-algebra kernels of the compiled plant with the same nq (k_schur / k_pcg* depend on the plant only through
// its dimensions): forwarders instead of a second instantiation.
#include "launchers.h"
namespace gato {
#define GATO_FORWARD_LINALG(RT, CP)                                                                                    \
        template<>                                                                                                     \
        size_t schur_smem_bytes<RT>()                                                                                  \
        {                                                                                                              \
                return schur_smem_bytes<CP>();                                                                         \
        }                                                                                                              \
        template<>                                                                                                     \
        void enqueue_schur<RT>(const Ctx& c, size_t smem, cudaStream_t st)                                             \
        {                                                                                                              \
                enqueue_schur<CP>(c, smem, st);                                                                        \
        }                                                                                                              \
        template<>                                                                                                     \
        int enqueue_pcg<RT>(const Ctx& c, int rpt, int threads, size_t smem, bool cluster, cudaStream_t st)            \
        {                                                                                                              \
                return enqueue_pcg<CP>(c, rpt, threads, smem, cluster, st);                                            \
        }                                                                                                              \
        template<>                                                                                                     \
        bool pcg_cluster_supported<RT>(int N)                                                                          \
        {                                                                                                              \
                return pcg_cluster_supported<CP>(N);                                                                   \
        }                                                                                                              \
        template<>                                                                                                     \
        cudaError_t configure_linalg<RT>(int device)                                                                   \
        {                                                                                                              \
                return configure_linalg<CP>(device);                                                                   \
        }
GATO_FORWARD_LINALG(RtPlant<6>, Indy7)
GATO_FORWARD_LINALG(RtPlant<7>, Iiwa14)
}  // namespace gato
