/* TEST INFRASTRUCTURE — CPU oracle for the BSQP solve path (NOT part of the product).
 *
 * A host-only fp32 restatement of the reference's algorithm (A2R-Lab/GATO, /root/reference/gato),
 * stage by stage, citing the reference file:line each function follows (see bsqp_oracle.cpp).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (gato_b200/) never includes, links or calls anything in oracle/.
 *
 * Parity status: PINNED against outputs of the reference itself — the unmodified reference headers
 * are compiled for sm_100 by oracle/build_ref.sh (oracle/ref_harness.cu) and run on a B200 by
 * oracle/gen_golden.py; the resulting vectors live in tests/golden/ (see tests/golden/README.md).
 * The reference ships no tests/golden vectors of its own.
 *
 * All entry points mirror oracle/ref_harness.cu's gref_* signatures so one test driver serves both.
 */
#ifndef GATO_BSQP_ORACLE_H
#define GATO_BSQP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* plant: 0 = indy7 (nq 6), 1 = iiwa14 (nq 7), >= 2 = a robot registered with gato_oracle_register_model */
typedef struct gato_oracle gato_oracle;

/* params15 = BSQP ctor order (bsqp.cuh:43): dt, max_sqp_iters, kkt_tol, max_pcg_iters, pcg_tol, solve_ratio, mu,
 * q_cost, qd_cost, u_cost, N_cost, q_lim_cost, vel_lim_cost, ctrl_lim_cost, rho */
gato_oracle* gato_oracle_create(int plant, int N, int B, const float* params15);
void         gato_oracle_destroy(gato_oracle*);
/* which: 0 f_ext[B*6], 1 rho[B], 2 drho[B], 3 mu[B], 4 pcg_tol[B]   (bsqp.cuh:63-79) */
int  gato_oracle_set_batch(gato_oracle*, int which, const float* h, int set_default);
/* which: 0 dual (bsqp.cuh:81), 1 rho (bsqp.cuh:83-87) */
int  gato_oracle_reset(gato_oracle*, int which);
void gato_oracle_set_rho_adaptation(gato_oracle*, int on);
int  gato_oracle_solve(gato_oracle*, float* xu, const float* xs, const float* ref, float dt, int* sqp_iters, int* kkt_conv, int* n_pcg, int* n_ls, int* pcg_iters, float* ls_min_merit,
                       float* ls_step, int cap_iters, float* final_merit, float* initial_merit, double* solve_time_us);
int  gato_oracle_sim_forward(gato_oracle*, const float* xk, const float* uk, float dt, float* xkp1);
void gato_oracle_set_threads(int n);
/* a robot from data tables (see bsqp_oracle.cpp); returns its plant id (>= 2) or -1 */
int  gato_oracle_register_model(int nq, int style, const double* table, const double* limits, int nxt, const int* xt_idx, const double* xt_coef, const int* xt_k, int nxht,
                                const int* xht_idx, const double* xht_coef, const int* xht_k, int ndxht, const int* dxht_idx, const double* dxht_coef, const int* dxht_k);

/* stateless per-stage entry points (same buffers/layouts as the reference kernels) */
int gato_oracle_stage_kkt(int plant, int N, int B, const float* xu, const float* xs, const float* ref, const float* fext, float dt, const float* cost7, float* Q, float* R, float* q, float* r,
                          float* A, float* Bm, float* c);
int gato_oracle_stage_schur(int plant, int N, int B, float* Q, float* R, const float* q, const float* r, const float* A, const float* Bm, const float* c, const float* rho, float* S,
                            float* Pinv, float* gamma);
int gato_oracle_stage_pcg(int plant, int N, int B, const float* S, const float* Pinv, const float* gamma, float* lambda, const float* eps, int max_iters, const int* kkt_conv, int* iters);
int gato_oracle_stage_dz(int plant, int N, int B, const float* lambda, const float* Qinv, const float* Rinv, float* q, float* r, const float* A, const float* Bm, float* dz);
int gato_oracle_stage_merit(int plant, int N, int B, const float* xu, const float* dz, const float* xs, const float* ref, const float* mu, const float* fext, float dt, const float* cost7,
                            int num_alphas, float* merit);
int gato_oracle_stage_linesearch(int plant, int N, int B, float* xu, const float* dz, const float* merit8, float* merit_init, float* step, float* rho, float* drho, int adapt);
int gato_oracle_dyn_dump(int plant, int n, const float* x, const float* u, const float* fext, float* qdd, float* dqdd, float* ee, float* dee);
/* libdevice-equivalent scalar functions, exposed for tests */
float gato_oracle_sinf(float x);
float gato_oracle_cosf(float x);
float gato_oracle_logf(float x);

#ifdef __cplusplus
}
#endif
#endif
