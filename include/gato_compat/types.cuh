// Shim for gato/types.cuh:13-59: the PODs callers of BSQP<T,B>::solve see.
#pragma once
#include <cstdint>
#include <vector>
#include "settings.h"
#include "constants.h"
#include "utils/cuda.cuh"
using namespace sqp;
using namespace gato::constants;

template<typename T, uint32_t BatchSize>
struct ProblemInputs {
        T     timestep;
        T*    d_x_s_batch;
        T*    d_reference_traj_batch;
        void* d_GRiD_mem;
};
template<uint32_t BatchSize>
struct PCGStats {
        double           solve_time_us;
        std::vector<int> num_iterations;
        std::vector<int> converged;
        PCGStats() : solve_time_us(0), num_iterations(BatchSize, 0), converged(BatchSize, 0) {}
};
template<typename T, uint32_t BatchSize>
struct LineSearchStats {
        std::vector<T> min_merit;
        std::vector<T> step_size;
        LineSearchStats() : min_merit(BatchSize, 0.0), step_size(BatchSize, 0.0) {}
};
template<typename T, uint32_t BatchSize>
struct SQPStats {
        double                                     solve_time_us;
        std::vector<int>                           sqp_iterations;
        std::vector<int>                           kkt_converged;
        std::vector<PCGStats<BatchSize>>           pcg_stats;
        std::vector<LineSearchStats<T, BatchSize>> line_search_stats;
        SQPStats() : solve_time_us(0), sqp_iterations(BatchSize, 0), kkt_converged(BatchSize, 0) {}
};
