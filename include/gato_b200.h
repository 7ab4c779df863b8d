/* gato_b200 — C ABI of the B200-native batched SQP (BSQP) solve path.
 *
 * Drop-in boundary for A2R-Lab/GATO's solver object.  Each entry point names the reference interface it
 * replaces (paths relative to the reference repository root):
 *
 *   gato_create / gato_destroy      BSQP<T,BatchSize>::BSQP(...) / ~BSQP()        gato/bsqp/bsqp.cuh:24-61, 200-297
 *                                   (plant and KNOT_POINTS, compile-time macros in the reference —
 *                                    CMakeLists.txt:71-75 — are runtime selectors here; BatchSize is runtime)
 *   gato_solve                      BSQP::solve(T* d_xu, ProblemInputs)            gato/bsqp/bsqp.cuh:103-197
 *   gato_solve_host                 PyBSQP::solve (H2D, solve, D2H, stats)         python/bindings.cu:68-148
 *   gato_set_batch                  set_f_ext_batch / set_rho_penalty_batch / set_drho_batch /
 *                                   set_mu_batch / set_pcg_tol_batch               gato/bsqp/bsqp.cuh:63-79
 *   gato_reset                      reset_dual / reset_rho                         gato/bsqp/bsqp.cuh:81-87
 *   gato_set_rho_adaptation         set_rho_adaptation                             gato/bsqp/bsqp.cuh:89
 *   gato_sim_forward(_host)         BSQP::sim_forward / PyBSQP::sim_forward        gato/bsqp/bsqp.cuh:91, python/bindings.cu:180-194
 *   gato_get_merits                 copy_final_merit_to_host / copy_initial_merit0_to_host  gato/bsqp/bsqp.cuh:93-101
 *   gato_stats (struct)             SQPStats / PCGStats / LineSearchStats          gato/types.cuh:23-59
 *
 * Conventions: fp32; all layouts are the reference's external layouts: xu[B][(nx+nu)N-nu],
 * x_s[B][nx], ref[B][6N].  Pointers named d_* are DEVICE pointers on the solver's device, h_* are HOST pointers.
 * Every function returns 0 on success or a negative gato_status; gato_last_error() gives the message.
 * There is NO CPU fallback: if no CUDA device / kernel image is available the calls fail with GATO_ERR_CUDA.
 */
#ifndef GATO_B200_H
#define GATO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gato_solver gato_solver;

enum gato_plant { GATO_PLANT_INDY7 = 0, GATO_PLANT_IIWA14 = 1 };
enum gato_status { GATO_OK = 0, GATO_ERR_ARG = -1, GATO_ERR_CUDA = -2, GATO_ERR_UNSUPPORTED = -3 };
enum gato_batch_field { GATO_F_EXT = 0, GATO_RHO = 1, GATO_DRHO = 2, GATO_MU = 3, GATO_PCG_TOL = 4 };
enum gato_reset_field { GATO_RESET_DUAL = 0, GATO_RESET_RHO = 1 };

/* Constructor scalars, in the reference's constructor order (bsqp.cuh:43). */
typedef struct gato_params {
        float    dt;
        uint32_t max_sqp_iters;
        float    kkt_tol;
        uint32_t max_pcg_iters;
        float    pcg_tol;
        float    solve_ratio;
        float    mu;
        float    q_cost, qd_cost, u_cost, N_cost, q_lim_cost, vel_lim_cost, ctrl_lim_cost;
        float    rho;
} gato_params;

/* Per-solve statistics; buffers are owned by the solver and valid until the next gato_solve*/
typedef struct gato_stats {
        double         solve_time_us;   /* host wall clock around the solve, like bsqp.cuh:109,185-190 */
        float          device_time_ms;  /* CUDA-event time of the device work on the solver stream */
        int32_t        batch;
        int32_t        n_pcg;           /* PCG solves performed (= outer iterations executed) */
        int32_t        n_ls;            /* line searches performed (n_pcg or n_pcg-1, bsqp.cuh:165 vs :176) */
        const int32_t* sqp_iters;       /* [B] */
        const int32_t* kkt_converged;   /* [B] */
        const int32_t* pcg_iters;       /* [n_pcg][B] */
        const float*   ls_min_merit;    /* [n_ls][B] */
        const float*   ls_step_size;    /* [n_ls][B], -1 = rejected */
        const float*   final_merit;     /* [B] */
        const float*   initial_merit;   /* [B] */
} gato_stats;

/* Run-time robot models: the constants the reference bakes into GRiD-generated source per robot
 * (init_XImats, gato/dynamics/iiwa14/iiwa14_grid.cuh:1211-2087; the sin/cos assignment patterns of load_update_XImats_helpers :2212-2293 and
 * load_update_XmatsHom_helpers :2365-2448; the limits of iiwa14_plant.cuh:36-70), as DATA.  A registered model gets a plant id >= GATO_PLANT_MODEL0
 * that every entry point taking `plant` accepts (gato_create, gato_dims, gato_stage_*); its solves run the table-driven kernels.
 * Robots covered: fixed-base serial chains of z-axis revolute joints with nq = 6, 7 or 8.
 *   X[36 j + 6 c + r]      constant part of the 6x6 Pluecker transform of joint j (column-major); top-right 3x3 block zero, bottom-right = top-left
 *   I[36 j + 6 c + r]      spatial inertia of link j
 *   Xhom / dXhom[16 j + 4 c + r]   constant parts of the 4x4 homogeneous transform of joint j and of its derivative w.r.t. q_j
 *   *_trig                 entry idx (into the flattened array above) = (float)(coef * (double)t[k]), t[k < nq] = sin(q_k), t[k >= nq] = cos(q_{k-nq})
 *   *_limit                symmetric limits +-L; the reference's margin (JOINT_LIMIT_MARGIN = -0.1) is applied by the library
 *   style                  limit-barrier terms of the cost Hessian: 1 = iiwa14_plant.cuh:399-420 (second derivatives), 0 = indy7_plant.cuh:385-415 */
#define GATO_MODEL_MAX_NQ 8
#define GATO_MODEL_MAX_TRIG 16 /* per joint */
#define GATO_PLANT_MODEL0 2
typedef struct gato_trig {
        int32_t idx, k;
        double  coef;
} gato_trig;
typedef struct gato_model {
        char      name[32];
        int32_t   nq, style;
        double    X[36 * GATO_MODEL_MAX_NQ], I[36 * GATO_MODEL_MAX_NQ], Xhom[16 * GATO_MODEL_MAX_NQ], dXhom[16 * GATO_MODEL_MAX_NQ];
        double    joint_limit[GATO_MODEL_MAX_NQ], vel_limit[GATO_MODEL_MAX_NQ], ctrl_limit[GATO_MODEL_MAX_NQ];
        int32_t   n_x_trig, n_xh_trig, n_dxh_trig;
        gato_trig x_trig[GATO_MODEL_MAX_TRIG * GATO_MODEL_MAX_NQ], xh_trig[8 * GATO_MODEL_MAX_NQ], dxh_trig[8 * GATO_MODEL_MAX_NQ];
} gato_model;
int gato_model_builtin(int plant /* GATO_PLANT_INDY7 | GATO_PLANT_IIWA14 */, gato_model* out); /* the compiled robots' tables as a model */
int gato_model_load(const char* path, gato_model* out); /* text file written by gato_model_save */
int gato_model_save(const gato_model* m, const char* path);
int gato_model_register(const gato_model* m); /* plant id >= GATO_PLANT_MODEL0, or a negative gato_status (gato_last_error(NULL) says why) */

/* stream: a cudaStream_t (as void*) or NULL for a solver-owned non-blocking stream. */
int  gato_create(gato_solver** out, int plant, int knot_points, int batch, int device, void* stream, const gato_params* params);
void gato_destroy(gato_solver* s);
const char* gato_last_error(const gato_solver* s); /* s may be NULL: last creation error */

int gato_set_batch(gato_solver* s, int field, const float* h_values, int set_as_reset_default);
int gato_reset(gato_solver* s, int field);
int gato_set_rho_adaptation(gato_solver* s, int enabled);

/* Solve in place on device memory (synchronous, like the reference). */
int gato_solve(gato_solver* s, float* d_xu, const float* d_x_s, const float* d_ref, float timestep, gato_stats* stats);
/* Same through host buffers: H2D of xu/x_s/ref, solve, D2H of xu. */
int gato_solve_host(gato_solver* s, float* h_xu, const float* h_x_s, const float* h_ref, float timestep, gato_stats* stats);
/* Asynchronous pair for multi-GPU drivers: enqueue on the solver stream, then collect. */
int gato_solve_async(gato_solver* s, float* d_xu, const float* d_x_s, const float* d_ref, float timestep);
int gato_solve_wait(gato_solver* s, gato_stats* stats);

int gato_sim_forward(gato_solver* s, float* d_xkp1 /*[B][nx]*/, const float* d_xk /*[nx]*/, const float* d_uk /*[nu]*/, float dt);
int gato_sim_forward_host(gato_solver* s, float* h_xkp1, const float* h_xk, const float* h_uk, float dt);
int gato_get_merits(gato_solver* s, float* h_final /*[B] or NULL*/, float* h_initial /*[B] or NULL*/);

/* Closed-loop MPC step on the device: what python/bsqp/mpc_controller.py:233-253,294-309 does per control
 * step through four host round trips, as ONE stream of device work with one synchronisation:
 *   x_s[b] = x_curr (+ x_offset[b] if set), ref[b] = ref_window, XU[b][0:nx] = x_s[b]          (mpc_controller.py:238-242)
 *   reset_rho (optional, :245), solve (:248),
 *   x_next[b] = sim_forward(x_last, u_last, sim_dt) under hypothesis b's f_ext; err[b] = ||x_next[b] - x_curr||_2 in float64
 *   with numpy's summation order; best = np.argmin (first minimum; the first NaN wins) (:298-303)   (skipped when h_x_last == NULL: best = 0)
 *   XU[:] = XU[best]  (:252-253) -- the batch of warm starts stays resident on the device between steps.
 * The warm-start batch is solver-owned: seed it with gato_mpc_set_warm_start before the first step. */
typedef struct gato_mpc_out {
        int32_t       best_id;
        double        best_error;
        const double* errors;   /* [B], solver-owned, valid until the next step; all zero when scoring was skipped */
        const float*  xu_best;  /* [traj], solver-owned pinned buffer: the selected trajectory */
} gato_mpc_out;
#define GATO_MPC_RESET_RHO 1
int gato_mpc_set_warm_start(gato_solver* s, const float* h_xu /* [traj] tiled over the batch, or [B][traj] if per_solve */, int per_solve);
int gato_mpc_set_state_offsets(gato_solver* s, const float* h_x_offset /* [B][nx] or NULL to clear */);
int gato_mpc_step(gato_solver* s, const float* h_x_curr /*[nx]*/, const float* h_ref_window /*[6N]*/, const float* h_x_last /*[nx] or NULL*/,
                  const float* h_u_last /*[nu] or NULL*/, float sim_dt, float timestep, int flags, gato_mpc_out* out, gato_stats* stats);
int gato_mpc_get_warm_start(gato_solver* s, float* h_xu /*[B][traj]*/);

/* The same step cut at the points where the ranks of a multi-GPU closed loop exchange data (the batch of hypotheses sharded over GPUs,
 * BASELINE.json config 5; mpc_controller.py:233-253, 294-309).  All three calls only enqueue on the solver's stream; pointers are DEVICE pointers:
 *   gato_mpc_local_async   d_in = [x_curr nx | ref window 6N | x_last nx | u_last nu] (e.g. just broadcast by NCCL); prepare, (reset_rho,) solve,
 *                          score this shard's hypotheses (if score != 0); writes the shard's winner record to d_record:
 *                          floats [0..1] error (one double) | [2] local id (int32) | [3] unused | [4 .. 4+traj) the winner's trajectory
 *   gato_mpc_adopt_async   d_records = n_records records, record_stride floats apart, in shard order (e.g. just all-gathered by NCCL): the global
 *                          winner by np.argmin semantics over the concatenated batch (first minimum, first NaN wins) becomes every local warm
 *                          start; its global id is shard * id_stride + local id
 *   gato_mpc_wait          synchronise and collect: out->best_id is the GLOBAL id, out->errors this shard's errors
 * gato_mpc_step is local + adopt (one record) + wait. */
int gato_mpc_record_floats(const gato_solver* s);
int gato_mpc_input_floats(const gato_solver* s);
int gato_mpc_local_async(gato_solver* s, const float* d_in, int score, float sim_dt, float timestep, int flags, float* d_record);
int gato_mpc_adopt_async(gato_solver* s, const float* d_records, int n_records, int record_stride, int id_stride);
int gato_mpc_wait(gato_solver* s, gato_mpc_out* out, gato_stats* stats);

/* Introspection */
int  gato_dims(int plant, int knot_points, int* nx, int* nu, int* traj_size);
long gato_kernel_launches(const gato_solver* s); /* kernels launched by this solver so far */
int  gato_get_device_pointers(gato_solver* s, float** d_xu, float** d_x_s, float** d_ref); /* solver-owned staging buffers used by *_host calls */
/* Per-kernel device timing (measurement aid, off by default): when enabled, every kernel launch of a solve is bracketed by CUDA
 * events on the solver's stream.  gato_get_kernel_times returns, for the LAST completed solve, the summed duration and the number
 * of launches per kernel class: 0 k_kkt, 1 k_schur, 2 k_pcg, 3 k_merit_ls<8> (merit + line search), 4 k_merit_ls<1>. */
#define GATO_NUM_KERNEL_CLASSES 5
int gato_set_kernel_timing(gato_solver* s, int enable);
int gato_get_kernel_times(gato_solver* s, float* total_ms /* [5] */, int* launches /* [5] */);
/* every launch of the last completed solve in launch order: class id and duration; returns the number of launches written (<= cap) or < 0 */
int gato_get_launch_times(gato_solver* s, int* kernel_class, float* ms, int cap);

/* End-effector position of n joint configurations by the solver's own forward kinematics (the one its tracking cost uses):
 * q[n][nq] -> ee[n][3].  Replaces the pinocchio call of python/bsqp/interface.py:212-214 (BSQP.ee_pos). */
int gato_ee_pos(gato_solver* s, const float* h_q, int n, float* h_ee);
/* Optional telemetry: the KKT residual norms the reference computes on the host in every SQP iteration and then discards
 * (gato/bsqp/bsqp.cuh:149-150): q_max[i][b] = max |q residual| and c_max[i][b] = max |c| over the state entries of solve b after the primal
 * step of iteration i.  Off by default (no cost); gato_get_kkt_residuals copies [n_pcg][B] floats each for the last completed solve and
 * returns n_pcg (or a negative status). */
int gato_set_kkt_residual_log(gato_solver* s, int enable);
int gato_get_kkt_residuals(gato_solver* s, float* h_q_max, float* h_c_max);

/* Measured FP32 CUDA-core peak of `device` in TFLOP/s (the denominator of the FP32 roofline fractions bench.py
 * reports; MEASURED_PEAKS.json has no FP32 number): a register-resident fused-multiply-add loop at full occupancy, packed = 0 scalar FFMA,
 * packed = 1 Blackwell's two-wide FFMA2 (what the linear-algebra kernels of this library issue).  Measurement aid, not on the solve path. */
int gato_measure_fp32_peak(int device, int packed, double* tflops);

/* Stage-level entry points (host buffers; used by the parity tests, same layouts as the reference kernels'
 * global buffers — setup_kkt.cuh:15, schur_linsys.cuh:14/214/316, pcg.cuh:14, merit.cuh:17, line_search.cuh:13). */
int gato_stage_kkt(int plant, int N, int B, const float* xu, const float* xs, const float* ref, const float* fext, float dt, const float* cost7, float* Q, float* R, float* q, float* r, float* A,
                   float* Bm, float* c);
int gato_stage_schur(int plant, int N, int B, float* Q, float* R, const float* q, const float* r, const float* A, const float* Bm, const float* c, const float* rho, float* S, float* Pinv,
                     float* gamma);
int gato_stage_pcg(int plant, int N, int B, const float* S, const float* Pinv, const float* gamma, float* lambda, const float* eps, int max_iters, const int* kkt_conv, int* iters);
int gato_stage_dz(int plant, int N, int B, const float* lambda, const float* Qinv, const float* Rinv, float* q, float* r, const float* A, const float* Bm, float* dz);
int gato_stage_merit(int plant, int N, int B, const float* xu, const float* dz, const float* xs, const float* ref, const float* mu, const float* fext, float dt, const float* cost7, int num_alphas,
                     float* merit);
int gato_stage_linesearch(int plant, int N, int B, float* xu, const float* dz, const float* merit8, float* merit_init, float* step, float* rho, float* drho, int adapt);

#ifdef __cplusplus
}
#endif
#endif /* GATO_B200_H */
