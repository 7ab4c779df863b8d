// cost weights of the tracking problem, in the reference constructor's order (bsqp.cuh:43)
#pragma once
namespace gato {
struct Costs {
        float q_cost, qd_cost, u_cost, N_cost, q_lim_cost, vel_lim_cost, ctrl_lim_cost;
};
}  // namespace gato
