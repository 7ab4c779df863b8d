// Source-compatibility shim: the reference's solver class BSQP<T,BatchSize> (gato/bsqp/bsqp.cuh:20-353) implemented
// over the gato_b200 C ABI, so that callers written against the reference — examples/bsqp.cu:23,63 and
// python/bindings.cu:10-209 — compile unchanged with  -I include/gato_compat  and link  -lgato_b200.
// Same constructor scalars, setters, ownership (caller owns d_xu / x_s / reference buffers), synchronous solve.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../types.cuh"
#include "../../gato_b200.h"

template<typename T, uint32_t BatchSize>
class BSQP {
        static_assert(sizeof(T) == sizeof(float), "gato_b200 is an fp32 solver");

      public:
        BSQP() : BSQP(0.01, 5, 0.0001, 100, 1e-5, 1.0, 10.0, 1.0, 1e-3, 1e-6, 50.0, 1e-3, 0.0, 0.0, 1e-3) {}  // bsqp.cuh:24-28
        BSQP(T dt, uint32_t max_sqp_iters, T kkt_tol, uint32_t max_pcg_iters, T pcg_tol, T solve_ratio, T mu, T q_cost, T qd_cost, T u_cost, T N_cost, T q_lim_cost, T vel_lim_cost,
             T ctrl_lim_cost, T rho)
        {
                gato_params p{(float)dt, max_sqp_iters, (float)kkt_tol, max_pcg_iters, (float)pcg_tol, (float)solve_ratio, (float)mu, (float)q_cost, (float)qd_cost, (float)u_cost,
                              (float)N_cost, (float)q_lim_cost, (float)vel_lim_cost, (float)ctrl_lim_cost, (float)rho};
                int         dev = 0;
                cudaGetDevice(&dev);
                // the reference launches on the legacy default stream of the current device; keep that contract
                if (gato_create(&s_, GATO_COMPAT_PLANT, KNOT_POINTS, BatchSize, dev, (void*)cudaStreamLegacy, &p) != GATO_OK) {
                        fprintf(stderr, "BSQP: %s\n", gato_last_error(nullptr));
                        abort();
                }
        }
        ~BSQP() { gato_destroy(s_); }
        BSQP(const BSQP&) = delete;
        BSQP& operator=(const BSQP&) = delete;

        void set_f_ext_batch(T* h) { gato_set_batch(s_, GATO_F_EXT, h, 0); }
        void set_rho_penalty_batch(const T* h, bool set_as_reset_default = true) { gato_set_batch(s_, GATO_RHO, h, set_as_reset_default); }
        void set_drho_batch(const T* h, bool set_as_reset_default = true) { gato_set_batch(s_, GATO_DRHO, h, set_as_reset_default); }
        void set_mu_batch(const T* h) { gato_set_batch(s_, GATO_MU, h, 0); }
        void set_pcg_tol_batch(const T* h) { gato_set_batch(s_, GATO_PCG_TOL, h, 0); }
        void reset_dual() { gato_reset(s_, GATO_RESET_DUAL); }
        void reset_rho() { gato_reset(s_, GATO_RESET_RHO); }
        void set_rho_adaptation(bool enabled) { gato_set_rho_adaptation(s_, enabled); }
        void sim_forward(T* d_xkp1_batch, T* d_xk, T* d_uk, T dt) { gato_sim_forward(s_, d_xkp1_batch, d_xk, d_uk, dt); }
        void copy_final_merit_to_host(T* h_out) { gato_get_merits(s_, h_out, nullptr); }
        void copy_initial_merit0_to_host(T* h_out) { gato_get_merits(s_, nullptr, h_out); }

        SQPStats<T, BatchSize> solve(T* d_xu_traj_batch, ProblemInputs<T, BatchSize> inputs)
        {
                SQPStats<T, BatchSize> out;
                gato_stats             st{};
                if (gato_solve(s_, d_xu_traj_batch, inputs.d_x_s_batch, inputs.d_reference_traj_batch, inputs.timestep, &st) != GATO_OK) {
                        fprintf(stderr, "BSQP::solve: %s\n", gato_last_error(s_));
                        return out;
                }
                out.solve_time_us = st.solve_time_us;
                for (uint32_t b = 0; b < BatchSize; b++) {
                        out.sqp_iterations[b] = st.sqp_iters[b];
                        out.kkt_converged[b] = st.kkt_converged[b];
                }
                for (int i = 0; i < st.n_pcg; i++) {
                        PCGStats<BatchSize> p;
                        for (uint32_t b = 0; b < BatchSize; b++) p.num_iterations[b] = st.pcg_iters[(size_t)i * BatchSize + b];
                        out.pcg_stats.push_back(p);
                }
                for (int i = 0; i < st.n_ls; i++) {
                        LineSearchStats<T, BatchSize> l;
                        for (uint32_t b = 0; b < BatchSize; b++) {
                                l.min_merit[b] = st.ls_min_merit[(size_t)i * BatchSize + b];
                                l.step_size[b] = st.ls_step_size[(size_t)i * BatchSize + b];
                        }
                        out.line_search_stats.push_back(l);
                }
                return out;
        }

      private:
        gato_solver* s_ = nullptr;
};
