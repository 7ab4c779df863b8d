// Host side of the B200-native BSQP solver: the C ABI declared in include/gato_b200.h.
// Mirrors the reference's solver object BSQP<T,BatchSize> (gato/bsqp/bsqp.cuh:20-353): same constructor
// scalars, same persistent state (lambda and rho survive solves, drho is reset after each solve), same
// statistics — but the whole solve is enqueued on one stream without any host synchronisation inside
// (the reference blocks on a device->host copy in every SQP iteration, bsqp.cuh:133-137).
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gato_b200.h"
#include "launchers.h"

using namespace gato;

namespace {

thread_local std::string g_create_error;

#define CUDA_TRY(s, expr)                                                                                         \
        do {                                                                                                      \
                cudaError_t e__ = (expr);                                                                         \
                if (e__ != cudaSuccess) {                                                                         \
                        char buf__[512];                                                                          \
                        snprintf(buf__, sizeof(buf__), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
                        (s)->err = buf__;                                                                         \
                        return GATO_ERR_CUDA;                                                                     \
                }                                                                                                 \
        } while (0)

struct Dims {
        int nq, nx, nu, N, traj, vecp, brow;
        Dims(int nq_, int N_) : nq(nq_), nx(2 * nq_), nu(nq_), N(N_), traj(3 * nq_ * N_ - nq_), vecp((N_ + 2) * 2 * nq_), brow(3 * 4 * nq_ * nq_) {}
};

template<typename T>
struct DevArr {
        T*     p = nullptr;
        size_t n = 0;
        cudaError_t alloc(size_t count)
        {
                n = count;
                cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T));
                if (e != cudaSuccess) return e;
                return cudaMemset(p, 0, std::max<size_t>(count, 1) * sizeof(T));
        }
        void release()
        {
                if (p) cudaFree(p);
                p = nullptr;
        }
};

// ---- run-time robot models (include/gato_b200.h: gato_model_*) ------------------------------------------------------------------
// Registered models live for the life of the process; plant id = GATO_PLANT_MODEL0 + index = constant-memory slot + GATO_PLANT_MODEL0.
struct RegisteredModel {
        gato_model desc;
        RtModel    rt;
};
std::mutex                   g_model_mu;
std::vector<RegisteredModel> g_models;

// joints of a plant id, or 0 if the id names neither a compiled plant nor a registered model
int plant_nq(int plant)
{
        if (plant == GATO_PLANT_INDY7) return 6;
        if (plant == GATO_PLANT_IIWA14) return 7;
        std::lock_guard<std::mutex> lk(g_model_mu);
        const int                   i = plant - GATO_PLANT_MODEL0;
        return (i >= 0 && i < (int)g_models.size()) ? g_models[i].rt.nq : 0;
}
// f(plant tag): the compiled plants and the table-driven instantiations for nq = 6 / 7 / 8
template<class F>
auto with_plant(int plant, int nq, F&& f)
{
        if (plant == GATO_PLANT_INDY7) return f(Indy7{});
        if (plant == GATO_PLANT_IIWA14) return f(Iiwa14{});
        if (nq == 6) return f(RtPlant<6>{});
        if (nq == 7) return f(RtPlant<7>{});
        return f(RtPlant<8>{});
}

}  // namespace

struct gato_solver {
        int          plant, N, B, device;
        int          model_slot = 0;  // run-time models: plant - GATO_PLANT_MODEL0
        Dims         d;
        gato_params  prm;
        bool         adapt_rho = true;
        cudaStream_t stream = nullptr;
        bool         own_stream = false;
        // side stream for the initial merit, which only the first line search needs: it runs beside the first KKT / Schur / PCG kernels
        cudaStream_t side = nullptr;
        cudaEvent_t  ev_fork = nullptr, ev_join = nullptr;
        cudaEvent_t  ev0 = nullptr, ev1 = nullptr;
        std::string  err;
        long         launches = 0;
        // closed-loop MPC step state (allocated on first use)
        DevArr<float>  mpc_xu, mpc_xs, mpc_ref, mpc_off, mpc_in, mpc_xnext, mpc_rec;
        DevArr<double> mpc_err, mpc_gerr;
        double*        h_mpc_gerr = nullptr;  // pinned: the global winner's error
        bool           mpc_scored = false;
        DevArr<int>    mpc_best;
        bool           mpc_has_off = false;
        float*         h_mpc_in = nullptr;    // pinned: x_curr | ref window | x_last | u_last
        float*         h_mpc_best = nullptr;  // pinned: selected trajectory
        double*        h_mpc_err = nullptr;   // pinned: [B] errors
        int*           h_mpc_id = nullptr;    // pinned: best id
        // optional per-kernel timing: events e[0] K0 e[1] K1 ... on the stream; tick_class[i] = class of the kernel after e[i]
        bool                     timing = false;
        std::vector<cudaEvent_t> tick_ev;
        std::vector<int>         tick_class;
        size_t                   n_ticks = 0;
        int          max_it;  // allocation size of the per-iteration logs: max(max_sqp_iters, 1)
        int          n_it;    // SQP iterations enqueued per solve: max_sqp_iters (may be 0, bsqp.cuh:121)
        // device state
        DevArr<float>    Q, R, q, r, A, Bm, c, Qinv, Rinv, S, Pinv, Pmain, gamma, lambda, dz;
        DevArr<float>    rho, drho, mu, pcg_tol, fext, merit, merit_cur, merit0, step, ls_merit_log, ls_step_log;
        DevArr<int>      conv, pcg_log;
        // optional KKT residual telemetry (gato_set_kkt_residual_log): max |q residual| / max |c| per solve and iteration, as float bits
        DevArr<unsigned> kkt_qmax, kkt_cmax;
        float *          h_kkt_qmax = nullptr, *h_kkt_cmax = nullptr;
        bool             kkt_log = false;
        DevArr<float>    ee_q, ee_out;  // staging of gato_ee_pos
        // CUDA graphs of the solve's launch sequence for small batches (the MPC regime is launch-latency bound): one instantiated graph per
        // distinct (xu, x_s, ref pointers, dt, switches); replayed while the caller keeps passing the same device buffers
        struct GraphEntry {
                float*          xu;
                const float *   xs, *ref;
                float           dt;
                int             adapt, kkt_log;
                long            launches;
                cudaGraphExec_t exec;
        };
        std::vector<GraphEntry> graphs;
        bool                    use_graph = false;
        int                     sms = 148;  // multiprocessors of the device
        DevArr<unsigned> num_solved, num_unsolved, pcg_done;
        bool             overlap = true;  // merit / line search overlapped with the tail of k_pcg (GATO_NO_OVERLAP=1 switches it off)
        DevArr<float>    st_xu, st_xs, st_ref, st_xkp1, st_xk, st_uk;  // staging for *_host calls
        // reset defaults of rho / drho (bsqp.cuh:48-58, 84-87, 189), resident on the device so that resets are device-to-device copies
        DevArr<float> rho_init, drho_init;
        // pinned result buffers
        int *     h_pcg_log = nullptr, *h_conv = nullptr;
        unsigned* h_num_solved = nullptr;
        float *   h_ls_merit = nullptr, *h_ls_step = nullptr, *h_final = nullptr, *h_initial = nullptr;
        std::vector<int32_t> h_sqp_iters;
        std::chrono::high_resolution_clock::time_point t_start;
        bool                                           pending = false;
        size_t                                         smem_pcg = 0, smem_schur = 0;
        int                                            pcg_threads = 0, pcg_rpt = 0;
        bool                                           pcg_cluster = false;  // long horizons: thread-block cluster per solve (k_pcg_cluster) instead of streaming

        gato_solver(int plant_, int nq_, int N_, int B_, int dev)
            : plant(plant_), N(N_), B(B_), device(dev), model_slot(plant_ >= GATO_PLANT_MODEL0 ? plant_ - GATO_PLANT_MODEL0 : 0), d(nq_, N_), prm{}, max_it(1), n_it(1)
        {
        }
};

namespace {

template<class P>
int pcg_threads(int N)
{
        return ((N + 2) * 2 * P::NQ + 31) / 32 * 32;  // one thread per padded vector index (rows are indices nx .. nx + N*nx)
}
template<class P>
size_t pcg_smem_bytes(int N, int threads)
{
        return sizeof(float) * pcg_smem_floats<2 * P::NQ, P::NQ>(N, threads / 32);
}
template<class P>
int configure_kernels(gato_solver* s)
{
        s->pcg_threads = pcg_threads<P>(s->N);
        s->pcg_rpt = 0;  // 0: register-resident kernel (one thread per padded index); >0: streaming kernel, indices per thread
        if (s->pcg_threads > 512) {
                const int n = (s->N + 2) * 2 * P::NQ;
                s->pcg_rpt = (n + 1023) / 1024;
                if (s->pcg_rpt > 4) {
                        s->err = "knot_points too large (the streaming PCG kernel supports (N + 2) * nx <= 4096)";
                        return GATO_ERR_UNSUPPORTED;
                }
                s->pcg_threads = 1024;
                // GATO_PCG_NO_CLUSTER=1 keeps the streaming kernel (A/B measurements)
                s->pcg_cluster = pcg_cluster_supported<P>(s->N) && !getenv("GATO_PCG_NO_CLUSTER");
                // vectors | dot scratch | dz scratch | max(K2 scratch, 32 per-warp tiles of 32 rows x 3nx floats)
                s->smem_pcg = sizeof(float) * ((size_t)2 * n + 64 + 64 * 32 + std::max((size_t)(s->N - 1) * 4 * P::NQ * P::NQ, (size_t)32 * 32 * 6 * P::NQ));
        } else {
                s->smem_pcg = pcg_smem_bytes<P>(s->N, s->pcg_threads);
        }
        s->smem_schur = schur_smem_bytes<P>();
        int maxsm = 0;
        CUDA_TRY(s, cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device));
        if (s->smem_pcg > (size_t)maxsm) {
                s->err = "knot_points too large for the PCG kernel's shared memory on this device";
                return GATO_ERR_UNSUPPORTED;
        }
        CUDA_TRY(s, configure_linalg<P>(s->device));
        return GATO_OK;
}

Ctx make_ctx(gato_solver* s, float* d_xu, const float* d_xs, const float* d_ref, float dt)
{
        Ctx c{};
        c.sms = s->sms;
        c.model_slot = s->model_slot;
        c.N = s->N, c.B = s->B, c.it = 0, c.max_pcg = (int)s->prm.max_pcg_iters, c.adapt = s->adapt_rho ? 1 : 0, c.flags = 0;
        c.dt = dt;
        c.thresh = (float)(uint32_t)s->B * s->prm.solve_ratio;
        c.cs = Costs{s->prm.q_cost, s->prm.qd_cost, s->prm.u_cost, s->prm.N_cost, s->prm.q_lim_cost, s->prm.vel_lim_cost, s->prm.ctrl_lim_cost};
        c.xu = d_xu, c.xs = d_xs, c.ref = d_ref, c.fext = s->fext.p;
        c.Q = s->Q.p, c.R = s->R.p, c.q = s->q.p, c.r = s->r.p, c.A = s->A.p, c.Bm = s->Bm.p, c.c = s->c.p, c.Qinv = s->Qinv.p, c.Rinv = s->Rinv.p;
        c.S = s->S.p, c.Pinv = s->Pinv.p, c.Pmain = s->Pmain.p, c.gamma = s->gamma.p, c.lambda = s->lambda.p, c.dz = s->dz.p;
        c.rho = s->rho.p, c.drho = s->drho.p, c.merit = s->merit.p, c.merit_cur = s->merit_cur.p, c.step = s->step.p;
        c.mu = s->mu.p, c.pcg_tol = s->pcg_tol.p;
        c.kkt_qmax = s->kkt_log ? s->kkt_qmax.p : nullptr, c.kkt_cmax = s->kkt_log ? s->kkt_cmax.p : nullptr;
        c.num_unsolved = s->num_unsolved.p, c.pcg_done = s->pcg_done.p;
        c.conv = s->conv.p, c.num_solved = s->num_solved.p, c.pcg_log = s->pcg_log.p, c.ls_merit_log = s->ls_merit_log.p, c.ls_step_log = s->ls_step_log.p;
        return c;
}

// timing: record an event before a kernel of class `cls` (cls < 0: closing event)
void tick(gato_solver* s, int cls)
{
        if (!s->timing) return;
        if (s->n_ticks == s->tick_ev.size()) {
                cudaEvent_t e;
                if (cudaEventCreate(&e) != cudaSuccess) return;
                s->tick_ev.push_back(e);
                s->tick_class.push_back(-1);
        }
        s->tick_class[s->n_ticks] = cls;
        cudaEventRecord(s->tick_ev[s->n_ticks++], s->stream);
}
template<class P>
void launch_kkt(gato_solver* s, const Ctx& c)
{
        tick(s, 0);
        enqueue_kkt<P>(c, s->stream);
        s->launches++;
}
template<class P>
void launch_schur(gato_solver* s, const Ctx& c)
{
        tick(s, 1);
        enqueue_schur<P>(c, s->smem_schur, s->stream);
        s->launches++;
}
template<class P>
void launch_pcg(gato_solver* s, const Ctx& c)
{
        tick(s, 2);
        s->launches += enqueue_pcg<P>(c, s->pcg_rpt, s->pcg_threads, s->smem_pcg, s->pcg_cluster, s->stream);
}
template<class P, int NA>
void launch_merit(gato_solver* s, const Ctx& c)
{
        tick(s, NA == 1 ? 4 : 3);
        enqueue_merit<P>(c, NA, s->stream);
        s->launches++;
}

// BSQP::solve (bsqp.cuh:103-197) as one stream of launches
template<class P>
int enqueue_solve(gato_solver* s, float* d_xu, const float* d_xs, const float* d_ref, float dt)
{
        const int B = s->B;
        Ctx       c = make_ctx(s, d_xu, d_xs, d_ref, dt);
        s->n_ticks = 0;
        CUDA_TRY(s, cudaMemsetAsync(s->conv.p, 0, sizeof(int) * B, s->stream));
        CUDA_TRY(s, cudaMemsetAsync(s->num_solved.p, 0, sizeof(unsigned) * s->max_it, s->stream));
        CUDA_TRY(s, cudaMemsetAsync(s->num_unsolved.p, 0, sizeof(unsigned) * s->max_it, s->stream));
        CUDA_TRY(s, cudaMemsetAsync(s->pcg_done.p, 0, sizeof(unsigned) * (size_t)s->max_it * B, s->stream));
        CUDA_TRY(s, cudaMemsetAsync(s->pcg_log.p, 0, sizeof(int) * (size_t)s->max_it * B, s->stream));
        if (s->kkt_log) {
                CUDA_TRY(s, cudaMemsetAsync(s->kkt_qmax.p, 0, sizeof(unsigned) * (size_t)s->max_it * B, s->stream));
                CUDA_TRY(s, cudaMemsetAsync(s->kkt_cmax.p, 0, sizeof(unsigned) * (size_t)s->max_it * B, s->stream));
        }
        // initial merit (dz = 0, alpha = 1)  bsqp.cuh:116-118.  Nothing before the first line search depends on it, so it is enqueued on a
        // side stream and joined there (with the per-kernel timing on it stays in line so that the event brackets remain meaningful).
        c.flags = F_MERIT | F_ZERO_DZ;
        const bool forked = !s->timing && s->side;
        if (forked) {
                CUDA_TRY(s, cudaEventRecord(s->ev_fork, s->stream));
                CUDA_TRY(s, cudaStreamWaitEvent(s->side, s->ev_fork, 0));
                enqueue_merit<P>(c, 1, s->side);
                s->launches++;
                CUDA_TRY(s, cudaMemcpyAsync(s->merit0.p, s->merit_cur.p, sizeof(float) * B, cudaMemcpyDeviceToDevice, s->side));
                CUDA_TRY(s, cudaEventRecord(s->ev_join, s->side));
        } else {
                launch_merit<P, 1>(s, c);
                tick(s, -1);
                CUDA_TRY(s, cudaMemcpyAsync(s->merit0.p, s->merit_cur.p, sizeof(float) * B, cudaMemcpyDeviceToDevice, s->stream));
        }
        bool joined = !forked;
        for (int it = 0; it < s->n_it; it++) {
                c.it = it;
                c.flags = F_CHECK_STOP;
                launch_kkt<P>(s, c);
                launch_schur<P>(s, c);
                if (!joined) {
                        // the initial merit (side stream, beside k_kkt and k_schur) is joined ahead of the PCG launch: nothing may sit between
                        // k_pcg and the line-search launch that overlaps its tail
                        CUDA_TRY(s, cudaStreamWaitEvent(s->stream, s->ev_join, 0));
                        joined = true;
                }
                c.flags = F_CHECK_STOP | F_K2 | F_PCG | F_DZ | F_BOOK;
                launch_pcg<P>(s, c);
                // the register-resident k_pcg hands solves over one by one: the line search may start under its last wave (not with per-kernel timing,
                // whose events serialise the launches)
                const bool ov = s->overlap && !s->timing && s->pcg_rpt == 0 && !s->use_graph;  // (graphs are for batches that have no wave tail to fill)
                c.flags = F_CHECK_STOP | F_MERIT | F_LS | (ov ? F_OVERLAP : 0);
                launch_merit<P, kNumAlphas>(s, c);
        }
        // Final merit on the updated trajectory (bsqp.cuh:180-182): no launch needed.  merit_cur[b] already holds it bit-for-bit: it is the
        // merit of the last accepted line-search candidate, evaluated by k_merit_ls<8> at exactly the trajectory fmaf(step, dz, xu) that the
        // line search then stored (same expression, same knot-ordered sum), or -- if no step was ever accepted -- the initial merit of the
        // unchanged trajectory.  (The reference re-evaluates it with a fresh kernel.)
        if (!joined) CUDA_TRY(s, cudaStreamWaitEvent(s->stream, s->ev_join, 0));  // max_sqp_iters = 0: only the merits are evaluated
        tick(s, -1);
        CUDA_TRY(s, cudaGetLastError());
        // results -> pinned host buffers
        CUDA_TRY(s, cudaMemcpyAsync(s->h_num_solved, s->num_solved.p, sizeof(unsigned) * s->max_it, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->h_pcg_log, s->pcg_log.p, sizeof(int) * (size_t)s->max_it * B, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->h_ls_merit, s->ls_merit_log.p, sizeof(float) * (size_t)s->max_it * B, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->h_ls_step, s->ls_step_log.p, sizeof(float) * (size_t)s->max_it * B, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->h_conv, s->conv.p, sizeof(int) * B, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->h_final, s->merit_cur.p, sizeof(float) * B, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->h_initial, s->merit0.p, sizeof(float) * B, cudaMemcpyDeviceToHost, s->stream));
        if (s->kkt_log) {
                CUDA_TRY(s, cudaMemcpyAsync(s->h_kkt_qmax, s->kkt_qmax.p, sizeof(unsigned) * (size_t)s->max_it * B, cudaMemcpyDeviceToHost, s->stream));
                CUDA_TRY(s, cudaMemcpyAsync(s->h_kkt_cmax, s->kkt_cmax.p, sizeof(unsigned) * (size_t)s->max_it * B, cudaMemcpyDeviceToHost, s->stream));
        }
        // drho is reset after every solve (bsqp.cuh:189); lambda and rho persist
        CUDA_TRY(s, cudaMemcpyAsync(s->drho.p, s->drho_init.p, sizeof(float) * B, cudaMemcpyDeviceToDevice, s->stream));
        return GATO_OK;
}

int enqueue_direct(gato_solver* s, float* d_xu, const float* d_xs, const float* d_ref, float dt)
{
        return with_plant(s->plant, s->d.nq, [&](auto P) { return enqueue_solve<decltype(P)>(s, d_xu, d_xs, d_ref, dt); });
}

// Small batches: the same sequence (memsets, 4 * iterations + 1 kernels, the side-stream fork / join, the copies of the statistics to pinned
// memory) captured once into a CUDA graph and replayed with one launch.  Any failure of the capture falls back to direct enqueueing.
int dispatch_enqueue(gato_solver* s, float* d_xu, const float* d_xs, const float* d_ref, float dt)
{
        if (!s->use_graph || s->timing) return enqueue_direct(s, d_xu, d_xs, d_ref, dt);
        const int adapt = s->adapt_rho ? 1 : 0, klog = s->kkt_log ? 1 : 0;
        for (auto& e : s->graphs)
                if (e.xu == d_xu && e.xs == d_xs && e.ref == d_ref && e.dt == dt && e.adapt == adapt && e.kkt_log == klog) {
                        CUDA_TRY(s, cudaGraphLaunch(e.exec, s->stream));
                        s->launches += e.launches;
                        return GATO_OK;
                }
        const long l0 = s->launches;
        if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
                cudaGetLastError();
                s->use_graph = false;
                return enqueue_direct(s, d_xu, d_xs, d_ref, dt);
        }
        const int       rc = enqueue_direct(s, d_xu, d_xs, d_ref, dt);
        cudaGraph_t     graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaError_t     ce = cudaStreamEndCapture(s->stream, &graph);
        if (rc == GATO_OK && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (rc != GATO_OK || ce != cudaSuccess || !exec) {
                cudaGetLastError();
                s->use_graph = false;  // not capturable here: enqueue directly from now on
                s->launches = l0;
                return enqueue_direct(s, d_xu, d_xs, d_ref, dt);
        }
        if (s->graphs.size() >= 4) {
                cudaGraphExecDestroy(s->graphs.front().exec);
                s->graphs.erase(s->graphs.begin());
        }
        s->graphs.push_back({d_xu, d_xs, d_ref, dt, adapt, klog, s->launches - l0, exec});
        CUDA_TRY(s, cudaGraphLaunch(exec, s->stream));
        return GATO_OK;
}

int fill_stats(gato_solver* s, gato_stats* st)
{
        const int B = s->B;
        // outer iterations executed: up to and including the first one whose count met the early-exit test (bsqp.cuh:165)
        const float thresh = (float)(uint32_t)B * s->prm.solve_ratio;
        int         n_pcg = s->n_it, n_ls = s->n_it;
        for (int i = 0; i < s->n_it; i++)
                if ((float)s->h_num_solved[i] >= thresh) {
                        n_pcg = i + 1;
                        n_ls = i;
                        break;
                }
        for (int b = 0; b < B; b++) s->h_sqp_iters[b] = n_pcg;  // every solve is counted once per outer iteration (bsqp.cuh:153-162)
        if (st) {
                st->batch = B;
                st->n_pcg = n_pcg, st->n_ls = n_ls;
                st->sqp_iters = s->h_sqp_iters.data();
                st->kkt_converged = s->h_conv;
                st->pcg_iters = s->h_pcg_log;
                st->ls_min_merit = s->h_ls_merit;
                st->ls_step_size = s->h_ls_step;
                st->final_merit = s->h_final;
                st->initial_merit = s->h_initial;
        }
        return GATO_OK;
}

int check_dev(gato_solver* s)
{
        CUDA_TRY(s, cudaSetDevice(s->device));
        return GATO_OK;
}

}  // namespace

extern "C" {

int gato_dims(int plant, int N, int* nx, int* nu, int* traj)
{
        const int nq = plant_nq(plant);
        if (!nq || N < 3) return GATO_ERR_ARG;
        Dims d(nq, N);
        if (nx) *nx = d.nx;
        if (nu) *nu = d.nu;
        if (traj) *traj = d.traj;
        return GATO_OK;
}

const char* gato_last_error(const gato_solver* s) { return s ? s->err.c_str() : g_create_error.c_str(); }
long        gato_kernel_launches(const gato_solver* s) { return s ? s->launches : 0; }

int gato_get_launch_times(gato_solver* s, int* kernel_class, float* ms, int cap)
{
        if (!s || !kernel_class || !ms || cap < 0) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        if (s->pending) {
                s->err = "gato_get_launch_times: a solve is still pending (call gato_solve_wait first)";
                return GATO_ERR_ARG;
        }
        int n = 0;
        for (size_t i = 0; i + 1 < s->n_ticks && n < cap; i++) {
                if (s->tick_class[i] < 0) continue;
                float t = 0.0f;
                CUDA_TRY(s, cudaEventElapsedTime(&t, s->tick_ev[i], s->tick_ev[i + 1]));
                kernel_class[n] = s->tick_class[i];
                ms[n++] = t;
        }
        return n;
}

int gato_set_kernel_timing(gato_solver* s, int enable)
{
        if (!s) return GATO_ERR_ARG;
        s->timing = enable != 0;
        s->n_ticks = 0;
        return GATO_OK;
}

int gato_get_kernel_times(gato_solver* s, float* total_ms, int* launches)
{
        if (!s || !total_ms || !launches) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        if (s->pending) {
                s->err = "gato_get_kernel_times: a solve is still pending (call gato_solve_wait first)";
                return GATO_ERR_ARG;
        }
        for (int i = 0; i < GATO_NUM_KERNEL_CLASSES; i++) total_ms[i] = 0.0f, launches[i] = 0;
        for (size_t i = 0; i + 1 < s->n_ticks; i++) {
                const int cls = s->tick_class[i];
                if (cls < 0) continue;
                float ms = 0.0f;
                CUDA_TRY(s, cudaEventElapsedTime(&ms, s->tick_ev[i], s->tick_ev[i + 1]));
                total_ms[cls] += ms;
                launches[cls]++;
        }
        return GATO_OK;
}


// ---- run-time robot models --------------------------------------------------------------------------------------------------
extern "C++" {
namespace {
template<class P>
void fill_builtin(gato_model* m, const char* name, int style)
{
        memset(m, 0, sizeof(*m));
        snprintf(m->name, sizeof(m->name), "%s", name);
        constexpr int nq = P::NQ;
        m->nq = nq, m->style = style;
        for (int i = 0; i < 36 * nq; i++) m->X[i] = P::TABLE[i], m->I[i] = P::TABLE[36 * nq + i];
        for (int i = 0; i < 16 * nq; i++) m->Xhom[i] = P::TABLE[72 * nq + i], m->dXhom[i] = P::TABLE[72 * nq + 16 * nq + i];
        for (int j = 0; j < nq; j++) m->joint_limit[j] = P::JL[j], m->vel_limit[j] = P::VL[j], m->ctrl_limit[j] = P::CL[j];
        m->n_x_trig = P::NXT, m->n_xh_trig = P::NXHT, m->n_dxh_trig = P::NDXHT;
        for (int i = 0; i < P::NXT; i++) m->x_trig[i] = gato_trig{P::XT[i].idx, P::XT[i].k, P::XT[i].coef};
        for (int i = 0; i < P::NXHT; i++) m->xh_trig[i] = gato_trig{P::XHT[i].idx, P::XHT[i].k, P::XHT[i].coef};
        for (int i = 0; i < P::NDXHT; i++) m->dxh_trig[i] = gato_trig{P::DXHT[i].idx, P::DXHT[i].k, P::DXHT[i].coef};
}
}  // namespace
}  // extern "C++"

int gato_model_builtin(int plant, gato_model* out)
{
        if (!out) return GATO_ERR_ARG;
        if (plant == GATO_PLANT_IIWA14)
                fill_builtin<Iiwa14>(out, "iiwa14", 1);
        else if (plant == GATO_PLANT_INDY7)
                fill_builtin<Indy7>(out, "indy7", 0);
        else
                return GATO_ERR_ARG;
        return GATO_OK;
}

// Text format: "gato_model 1", then `key values...` records; numbers with 17 significant digits so that a save / load round trip is exact.
int gato_model_save(const gato_model* m, const char* path)
{
        if (!m || !path || m->nq < 1 || m->nq > GATO_MODEL_MAX_NQ) return GATO_ERR_ARG;
        FILE* f = fopen(path, "w");
        if (!f) {
                g_create_error = std::string("cannot write ") + path;
                return GATO_ERR_ARG;
        }
        const int nq = m->nq;
        fprintf(f, "gato_model 1\nname %s\nnq %d\nstyle %d\n", m->name[0] ? m->name : "robot", nq, m->style);
        auto vec = [&](const char* key, const double* v, int n) {
                fprintf(f, "%s", key);
                for (int i = 0; i < n; i++) fprintf(f, " %.17g", v[i]);
                fprintf(f, "\n");
        };
        vec("joint_limit", m->joint_limit, nq), vec("vel_limit", m->vel_limit, nq), vec("ctrl_limit", m->ctrl_limit, nq);
        for (int j = 0; j < nq; j++) {
                char key[32];
                snprintf(key, sizeof(key), "X %d", j), vec(key, m->X + 36 * j, 36);
                snprintf(key, sizeof(key), "I %d", j), vec(key, m->I + 36 * j, 36);
                snprintf(key, sizeof(key), "Xhom %d", j), vec(key, m->Xhom + 16 * j, 16);
                snprintf(key, sizeof(key), "dXhom %d", j), vec(key, m->dXhom + 16 * j, 16);
        }
        auto trig = [&](const char* key, const gato_trig* t, int n) {
                fprintf(f, "%s %d\n", key, n);
                for (int i = 0; i < n; i++) fprintf(f, "%d %.17g %d\n", t[i].idx, t[i].coef, t[i].k);
        };
        trig("x_trig", m->x_trig, m->n_x_trig), trig("xh_trig", m->xh_trig, m->n_xh_trig), trig("dxh_trig", m->dxh_trig, m->n_dxh_trig);
        fprintf(f, "end\n");
        const bool ok = !ferror(f);
        fclose(f);
        return ok ? GATO_OK : GATO_ERR_ARG;
}

int gato_model_load(const char* path, gato_model* out)
{
        if (!path || !out) return GATO_ERR_ARG;
        FILE* f = fopen(path, "r");
        if (!f) {
                g_create_error = std::string("cannot read ") + path;
                return GATO_ERR_ARG;
        }
        memset(out, 0, sizeof(*out));
        auto bad = [&](const std::string& why) {
                g_create_error = std::string(path) + ": " + why;
                fclose(f);
                return GATO_ERR_ARG;
        };
        char key[64];
        int  version = 0;
        if (fscanf(f, "%63s %d", key, &version) != 2 || strcmp(key, "gato_model") != 0 || version != 1) return bad("not a gato_model version 1 file");
        bool ended = false;
        auto vec = [&](double* v, int n) {
                for (int i = 0; i < n; i++)
                        if (fscanf(f, "%lf", v + i) != 1) return false;
                return true;
        };
        auto joint = [&](int& j) { return fscanf(f, "%d", &j) == 1 && j >= 0 && j < out->nq; };
        auto trig = [&](gato_trig* t, int32_t& n, int cap) {
                int cnt = 0;
                if (fscanf(f, "%d", &cnt) != 1 || cnt < 0 || cnt > cap) return false;
                n = cnt;
                for (int i = 0; i < cnt; i++) {
                        int idx, k;
                        double coef;
                        if (fscanf(f, "%d %lf %d", &idx, &coef, &k) != 3) return false;
                        t[i] = gato_trig{idx, k, coef};
                }
                return true;
        };
        while (!ended && fscanf(f, "%63s", key) == 1) {
                const std::string k = key;
                int               j = 0;
                if (k == "end") {
                        ended = true;
                } else if (k == "name") {
                        char nm[64];
                        if (fscanf(f, "%63s", nm) != 1) return bad("name");
                        snprintf(out->name, sizeof(out->name), "%.31s", nm);
                } else if (k == "nq") {
                        if (fscanf(f, "%d", &out->nq) != 1 || out->nq < 1 || out->nq > GATO_MODEL_MAX_NQ) return bad("nq out of range");
                } else if (k == "style") {
                        if (fscanf(f, "%d", &out->style) != 1) return bad("style");
                } else if (!out->nq) {
                        return bad("nq must come before the tables");
                } else if (k == "joint_limit") {
                        if (!vec(out->joint_limit, out->nq)) return bad(k);
                } else if (k == "vel_limit") {
                        if (!vec(out->vel_limit, out->nq)) return bad(k);
                } else if (k == "ctrl_limit") {
                        if (!vec(out->ctrl_limit, out->nq)) return bad(k);
                } else if (k == "X") {
                        if (!joint(j) || !vec(out->X + 36 * j, 36)) return bad(k);
                } else if (k == "I") {
                        if (!joint(j) || !vec(out->I + 36 * j, 36)) return bad(k);
                } else if (k == "Xhom") {
                        if (!joint(j) || !vec(out->Xhom + 16 * j, 16)) return bad(k);
                } else if (k == "dXhom") {
                        if (!joint(j) || !vec(out->dXhom + 16 * j, 16)) return bad(k);
                } else if (k == "x_trig") {
                        if (!trig(out->x_trig, out->n_x_trig, GATO_MODEL_MAX_TRIG * GATO_MODEL_MAX_NQ)) return bad(k);
                } else if (k == "xh_trig") {
                        if (!trig(out->xh_trig, out->n_xh_trig, 8 * GATO_MODEL_MAX_NQ)) return bad(k);
                } else if (k == "dxh_trig") {
                        if (!trig(out->dxh_trig, out->n_dxh_trig, 8 * GATO_MODEL_MAX_NQ)) return bad(k);
                } else {
                        return bad("unknown record '" + k + "'");
                }
        }
        if (!ended) return bad("truncated (no 'end' record)");
        fclose(f);
        return GATO_OK;
}

int gato_model_register(const gato_model* m)
{
        if (!m) return GATO_ERR_ARG;
        RegisteredModel r;
        r.desc = *m;
        if (const char* why = rt_model_from_desc(*m, r.rt)) {
                g_create_error = std::string("gato_model_register: ") + why;
                return GATO_ERR_UNSUPPORTED;
        }
        std::lock_guard<std::mutex> lk(g_model_mu);
        // registering the same tables again returns the id they already have
        for (size_t i = 0; i < g_models.size(); i++)
                if (memcmp(&g_models[i].rt, &r.rt, sizeof(RtModel)) == 0) return GATO_PLANT_MODEL0 + (int)i;
        if ((int)g_models.size() >= kRtSlots) {
                g_create_error = "gato_model_register: all model slots are in use";
                return GATO_ERR_UNSUPPORTED;
        }
        g_models.push_back(r);
        return GATO_PLANT_MODEL0 + (int)g_models.size() - 1;
}

int gato_create(gato_solver** out, int plant, int N, int B, int device, void* stream, const gato_params* prm)
{
        const int nq = plant_nq(plant);
        if (!out || !prm || !nq || N < 3 || B < 1) {
                g_create_error = "invalid argument";
                return GATO_ERR_ARG;
        }
        gato_solver* s = new gato_solver(plant, nq, N, B, device);
        s->prm = *prm;
        s->max_it = (int)std::max<uint32_t>(prm->max_sqp_iters, 1u);
        s->n_it = (int)prm->max_sqp_iters;
        auto fail = [&](int rc) {
                g_create_error = s->err;
                gato_destroy(s);
                *out = nullptr;
                return rc;
        };
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
                s->err = "no CUDA device available (gato_b200 has no CPU fallback)";
                return fail(GATO_ERR_CUDA);
        }
        if (check_dev(s)) return fail(GATO_ERR_CUDA);
        if (stream) {
                s->stream = (cudaStream_t)stream;
        } else {
                if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) {
                        s->err = "cudaStreamCreate failed";
                        return fail(GATO_ERR_CUDA);
                }
                s->own_stream = true;
        }
        if (cudaDeviceGetAttribute(&s->sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || s->sms < 1) {
                s->err = "cudaDeviceGetAttribute(multiProcessorCount) failed";
                return fail(GATO_ERR_CUDA);
        }
        int rc = with_plant(plant, nq, [&](auto P) { return configure_kernels<decltype(P)>(s); });
        if (rc) return fail(rc);
        if (plant >= GATO_PLANT_MODEL0) {
                // the robot's tables -> this device's constant-memory slot of both table-driven kernel groups (idempotent: a slot never changes)
                RtModel rt;
                {
                        std::lock_guard<std::mutex> lk(g_model_mu);
                        rt = g_models[s->model_slot].rt;
                }
                const cudaError_t ue = with_plant(plant, nq, [&](auto P) -> cudaError_t {
                        using Pl = decltype(P);
                        if constexpr (is_rt_plant<Pl>) {
                                const cudaError_t e1 = upload_rt_model_kkt<Pl>(s->model_slot, rt);
                                return e1 != cudaSuccess ? e1 : upload_rt_model_merit<Pl>(s->model_slot, rt);
                        } else {
                                return cudaSuccess;
                        }
                });
                if (ue != cudaSuccess) {
                        s->err = std::string("uploading the robot model failed: ") + cudaGetErrorString(ue);
                        return fail(GATO_ERR_CUDA);
                }
        }
        const Dims&  d = s->d;
        const size_t b = B, it = s->max_it;
        cudaError_t  e = cudaSuccess;
        auto         A = [&](auto& arr, size_t n) {
                if (e == cudaSuccess) e = arr.alloc(n);
        };
        A(s->Q, b * d.nx * d.nx * N), A(s->R, b * d.nu * d.nu * N), A(s->q, b * d.nx * N), A(s->r, b * d.nu * N), A(s->A, b * d.nx * d.nx * N), A(s->Bm, b * d.nx * d.nu * N);
        A(s->c, b * d.nx * N), A(s->Qinv, b * d.nx * d.nx * N), A(s->Rinv, b * d.nu * d.nu * N);
        A(s->S, b * d.brow * N), A(s->Pinv, b * d.brow * N), A(s->Pmain, b * d.nx * d.nx * N), A(s->gamma, b * d.vecp), A(s->lambda, b * d.vecp), A(s->dz, b * d.traj);
        A(s->rho, b), A(s->drho, b), A(s->rho_init, b), A(s->drho_init, b), A(s->mu, b), A(s->pcg_tol, b), A(s->fext, 6 * b), A(s->merit, kNumAlphas * b), A(s->merit_cur, b), A(s->merit0, b), A(s->step, b);
        A(s->ls_merit_log, it * b), A(s->ls_step_log, it * b), A(s->conv, b), A(s->pcg_log, it * b), A(s->num_solved, it), A(s->num_unsolved, it), A(s->pcg_done, it * b);
        A(s->st_xu, b * d.traj), A(s->st_xs, b * d.nx), A(s->st_ref, b * 6 * N), A(s->st_xkp1, b * d.nx), A(s->st_xk, d.nx), A(s->st_uk, d.nu);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&s->h_pcg_log, sizeof(int) * it * b);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&s->h_conv, sizeof(int) * b);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&s->h_num_solved, sizeof(unsigned) * it);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&s->h_ls_merit, sizeof(float) * it * b);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&s->h_ls_step, sizeof(float) * it * b);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&s->h_final, sizeof(float) * b);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&s->h_initial, sizeof(float) * b);
        if (e == cudaSuccess) e = cudaEventCreate(&s->ev0);
        if (e == cudaSuccess) e = cudaEventCreate(&s->ev1);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming);
        if (e != cudaSuccess) {
                s->err = std::string("allocation failed: ") + cudaGetErrorString(e);
                return fail(GATO_ERR_CUDA);
        }
        s->h_sqp_iters.assign(B, 0);
        s->overlap = !getenv("GATO_NO_OVERLAP");
        // graphs for the regime where the launches, not the kernels, set the latency (the split line-search kernel's regime: no overlapped launch inside)
        s->use_graph = B <= s->sms && !getenv("GATO_NO_GRAPH");
        // per-batch hyper-parameters  (bsqp.cuh:48-58)
        std::vector<float> rho0(B, prm->rho), drho0(B, 1.0f), mu(B, prm->mu), tol(B, prm->pcg_tol);
        cudaMemcpy(s->rho.p, rho0.data(), sizeof(float) * B, cudaMemcpyHostToDevice);
        cudaMemcpy(s->drho.p, drho0.data(), sizeof(float) * B, cudaMemcpyHostToDevice);
        cudaMemcpy(s->rho_init.p, rho0.data(), sizeof(float) * B, cudaMemcpyHostToDevice);
        cudaMemcpy(s->drho_init.p, drho0.data(), sizeof(float) * B, cudaMemcpyHostToDevice);
        cudaMemcpy(s->mu.p, mu.data(), sizeof(float) * B, cudaMemcpyHostToDevice);
        cudaMemcpy(s->pcg_tol.p, tol.data(), sizeof(float) * B, cudaMemcpyHostToDevice);
        if (cudaDeviceSynchronize() != cudaSuccess) {
                s->err = "initialisation failed";
                return fail(GATO_ERR_CUDA);
        }
        *out = s;
        return GATO_OK;
}

void gato_destroy(gato_solver* s)
{
        if (!s) return;
        cudaSetDevice(s->device);
        if (s->stream) cudaStreamSynchronize(s->stream);
        for (auto* a : {&s->Q, &s->R, &s->q, &s->r, &s->A, &s->Bm, &s->c, &s->Qinv, &s->Rinv, &s->S, &s->Pinv, &s->Pmain, &s->gamma, &s->lambda, &s->dz, &s->rho, &s->drho, &s->rho_init, &s->drho_init, &s->mu, &s->pcg_tol, &s->fext,
                        &s->merit, &s->merit_cur, &s->merit0, &s->step, &s->ls_merit_log, &s->ls_step_log, &s->st_xu, &s->st_xs, &s->st_ref, &s->st_xkp1, &s->st_xk, &s->st_uk})
                a->release();
        s->conv.release(), s->pcg_log.release(), s->num_solved.release(), s->num_unsolved.release(), s->pcg_done.release(), s->kkt_qmax.release(), s->kkt_cmax.release(), s->ee_q.release(), s->ee_out.release();
        if (s->h_kkt_qmax) cudaFreeHost(s->h_kkt_qmax);
        if (s->h_kkt_cmax) cudaFreeHost(s->h_kkt_cmax);
        for (void* p : {(void*)s->h_pcg_log, (void*)s->h_conv, (void*)s->h_num_solved, (void*)s->h_ls_merit, (void*)s->h_ls_step, (void*)s->h_final, (void*)s->h_initial})
                if (p) cudaFreeHost(p);
        for (auto* a : {&s->mpc_xu, &s->mpc_xs, &s->mpc_ref, &s->mpc_off, &s->mpc_in, &s->mpc_xnext, &s->mpc_rec}) a->release();
        s->mpc_err.release(), s->mpc_best.release(), s->mpc_gerr.release();
        for (void* p : {(void*)s->h_mpc_in, (void*)s->h_mpc_best, (void*)s->h_mpc_err, (void*)s->h_mpc_id, (void*)s->h_mpc_gerr})
                if (p) cudaFreeHost(p);
        for (cudaEvent_t e : s->tick_ev) cudaEventDestroy(e);
        for (auto& e : s->graphs) cudaGraphExecDestroy(e.exec);
        if (s->side) cudaStreamSynchronize(s->side), cudaStreamDestroy(s->side);
        if (s->ev_fork) cudaEventDestroy(s->ev_fork);
        if (s->ev_join) cudaEventDestroy(s->ev_join);
        if (s->ev0) cudaEventDestroy(s->ev0);
        if (s->ev1) cudaEventDestroy(s->ev1);
        if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
        delete s;
}

int gato_set_batch(gato_solver* s, int field, const float* h, int set_default)
{
        if (!s || !h) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        const size_t B = s->B;
        float *      dst = nullptr, *dflt = nullptr;
        size_t       n = B;
        switch (field) {
                case GATO_F_EXT: dst = s->fext.p, n = 6 * B; break;
                case GATO_RHO: dst = s->rho.p, dflt = s->rho_init.p; break;
                case GATO_DRHO: dst = s->drho.p, dflt = s->drho_init.p; break;
                case GATO_MU: dst = s->mu.p; break;
                case GATO_PCG_TOL: dst = s->pcg_tol.p; break;
                default: s->err = "unknown batch field"; return GATO_ERR_ARG;
        }
        CUDA_TRY(s, cudaMemcpyAsync(dst, h, sizeof(float) * n, cudaMemcpyHostToDevice, s->stream));
        if (set_default && dflt) CUDA_TRY(s, cudaMemcpyAsync(dflt, dst, sizeof(float) * n, cudaMemcpyDeviceToDevice, s->stream));
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        return GATO_OK;
}

int gato_reset(gato_solver* s, int field)
{
        if (!s) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        if (field == GATO_RESET_DUAL) {
                CUDA_TRY(s, cudaMemsetAsync(s->lambda.p, 0, sizeof(float) * s->lambda.n, s->stream));
        } else if (field == GATO_RESET_RHO) {
                CUDA_TRY(s, cudaMemcpyAsync(s->rho.p, s->rho_init.p, sizeof(float) * s->B, cudaMemcpyDeviceToDevice, s->stream));
                CUDA_TRY(s, cudaMemcpyAsync(s->drho.p, s->drho_init.p, sizeof(float) * s->B, cudaMemcpyDeviceToDevice, s->stream));
        } else {
                s->err = "unknown reset field";
                return GATO_ERR_ARG;
        }
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        return GATO_OK;
}

int gato_set_rho_adaptation(gato_solver* s, int enabled)
{
        if (!s) return GATO_ERR_ARG;
        s->adapt_rho = enabled != 0;
        return GATO_OK;
}

int gato_solve_async(gato_solver* s, float* d_xu, const float* d_xs, const float* d_ref, float dt)
{
        if (!s || !d_xu || !d_xs || !d_ref) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        s->t_start = std::chrono::high_resolution_clock::now();
        CUDA_TRY(s, cudaEventRecord(s->ev0, s->stream));
        int rc = dispatch_enqueue(s, d_xu, d_xs, d_ref, dt);
        if (rc) return rc;
        CUDA_TRY(s, cudaEventRecord(s->ev1, s->stream));
        s->pending = true;
        return GATO_OK;
}

int gato_solve_wait(gato_solver* s, gato_stats* st)
{
        if (!s) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        s->pending = false;
        fill_stats(s, st);
        if (st) {
                st->solve_time_us = std::chrono::duration<double, std::micro>(std::chrono::high_resolution_clock::now() - s->t_start).count();
                float ms = 0;
                CUDA_TRY(s, cudaEventElapsedTime(&ms, s->ev0, s->ev1));
                st->device_time_ms = ms;
        }
        return GATO_OK;
}

int gato_solve(gato_solver* s, float* d_xu, const float* d_xs, const float* d_ref, float dt, gato_stats* st)
{
        int rc = gato_solve_async(s, d_xu, d_xs, d_ref, dt);
        if (rc) return rc;
        return gato_solve_wait(s, st);
}

int gato_solve_host(gato_solver* s, float* h_xu, const float* h_xs, const float* h_ref, float dt, gato_stats* st)
{
        if (!s || !h_xu || !h_xs || !h_ref) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        const Dims& d = s->d;
        const auto  t0 = std::chrono::high_resolution_clock::now();
        CUDA_TRY(s, cudaMemcpyAsync(s->st_xu.p, h_xu, sizeof(float) * (size_t)s->B * d.traj, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->st_xs.p, h_xs, sizeof(float) * (size_t)s->B * d.nx, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->st_ref.p, h_ref, sizeof(float) * (size_t)s->B * 6 * s->N, cudaMemcpyHostToDevice, s->stream));
        int rc = gato_solve_async(s, s->st_xu.p, s->st_xs.p, s->st_ref.p, dt);
        if (rc) return rc;
        CUDA_TRY(s, cudaMemcpyAsync(h_xu, s->st_xu.p, sizeof(float) * (size_t)s->B * d.traj, cudaMemcpyDeviceToHost, s->stream));
        rc = gato_solve_wait(s, st);
        if (st) st->solve_time_us = std::chrono::duration<double, std::micro>(std::chrono::high_resolution_clock::now() - t0).count();
        return rc;
}

int gato_get_device_pointers(gato_solver* s, float** d_xu, float** d_xs, float** d_ref)
{
        if (!s) return GATO_ERR_ARG;
        if (d_xu) *d_xu = s->st_xu.p;
        if (d_xs) *d_xs = s->st_xs.p;
        if (d_ref) *d_ref = s->st_ref.p;
        return GATO_OK;
}

int gato_get_merits(gato_solver* s, float* h_final, float* h_initial)
{
        if (!s) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        if (s->pending) {
                s->err = "gato_get_merits: a solve is still pending (call gato_solve_wait first)";
                return GATO_ERR_ARG;
        }
        // ordered after everything enqueued on the solver's (non-blocking) stream
        if (h_final) CUDA_TRY(s, cudaMemcpyAsync(h_final, s->merit_cur.p, sizeof(float) * s->B, cudaMemcpyDeviceToHost, s->stream));
        if (h_initial) CUDA_TRY(s, cudaMemcpyAsync(h_initial, s->merit0.p, sizeof(float) * s->B, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        return GATO_OK;
}

int gato_set_kkt_residual_log(gato_solver* s, int enable)
{
        if (!s) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        if (s->pending) {
                s->err = "gato_set_kkt_residual_log: a solve is still pending (call gato_solve_wait first)";
                return GATO_ERR_ARG;
        }
        if (enable && !s->kkt_qmax.p) {
                const size_t n = (size_t)s->max_it * s->B;
                CUDA_TRY(s, s->kkt_qmax.alloc(n));
                CUDA_TRY(s, s->kkt_cmax.alloc(n));
                CUDA_TRY(s, cudaMallocHost((void**)&s->h_kkt_qmax, sizeof(float) * n));
                CUDA_TRY(s, cudaMallocHost((void**)&s->h_kkt_cmax, sizeof(float) * n));
        }
        s->kkt_log = enable != 0;
        return GATO_OK;
}

int gato_get_kkt_residuals(gato_solver* s, float* h_q_max, float* h_c_max)
{
        if (!s) return GATO_ERR_ARG;
        if (s->pending || !s->kkt_log) {
                s->err = s->pending ? "gato_get_kkt_residuals: a solve is still pending" : "gato_get_kkt_residuals: enable the log with gato_set_kkt_residual_log first";
                return GATO_ERR_ARG;
        }
        // rows of the last completed solve: one per PCG solve performed
        const float thresh = (float)(uint32_t)s->B * s->prm.solve_ratio;
        int         n_pcg = s->n_it;
        for (int i = 0; i < s->n_it; i++)
                if ((float)s->h_num_solved[i] >= thresh) {
                        n_pcg = i + 1;
                        break;
                }
        const size_t n = (size_t)n_pcg * s->B;
        if (h_q_max) memcpy(h_q_max, s->h_kkt_qmax, sizeof(float) * n);  // the maxima were taken on the bit patterns of non-negative floats
        if (h_c_max) memcpy(h_c_max, s->h_kkt_cmax, sizeof(float) * n);
        return n_pcg;
}

int gato_ee_pos(gato_solver* s, const float* h_q, int n, float* h_ee)
{
        if (!s || !h_q || !h_ee || n < 1) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        const int nq = s->d.nq;
        if (s->ee_q.n < (size_t)n * nq) {
                s->ee_q.release(), s->ee_out.release();
                CUDA_TRY(s, s->ee_q.alloc((size_t)n * nq));
                CUDA_TRY(s, s->ee_out.alloc((size_t)n * 3));
        }
        CUDA_TRY(s, cudaMemcpyAsync(s->ee_q.p, h_q, sizeof(float) * (size_t)n * nq, cudaMemcpyHostToDevice, s->stream));
        with_plant(s->plant, nq, [&](auto P) { enqueue_ee_pos<decltype(P)>(n, s->ee_q.p, s->ee_out.p, s->model_slot, s->stream); });
        s->launches++;
        CUDA_TRY(s, cudaGetLastError());
        CUDA_TRY(s, cudaMemcpyAsync(h_ee, s->ee_out.p, sizeof(float) * (size_t)n * 3, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        return GATO_OK;
}

int gato_sim_forward(gato_solver* s, float* d_xkp1, const float* d_xk, const float* d_uk, float dt)
{
        if (!s || !d_xkp1 || !d_xk || !d_uk) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        with_plant(s->plant, s->d.nq, [&](auto P) { enqueue_sim_forward<decltype(P)>(s->B, d_xkp1, d_xk, d_uk, s->fext.p, dt, s->model_slot, s->stream); });
        s->launches++;
        CUDA_TRY(s, cudaGetLastError());
        return GATO_OK;
}

int gato_sim_forward_host(gato_solver* s, float* h_xkp1, const float* h_xk, const float* h_uk, float dt)
{
        if (!s || !h_xkp1 || !h_xk || !h_uk) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        CUDA_TRY(s, cudaMemcpyAsync(s->st_xk.p, h_xk, sizeof(float) * s->d.nx, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->st_uk.p, h_uk, sizeof(float) * s->d.nu, cudaMemcpyHostToDevice, s->stream));
        int rc = gato_sim_forward(s, s->st_xkp1.p, s->st_xk.p, s->st_uk.p, dt);
        if (rc) return rc;
        CUDA_TRY(s, cudaMemcpyAsync(h_xkp1, s->st_xkp1.p, sizeof(float) * (size_t)s->B * s->d.nx, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        return GATO_OK;
}

// ------------------------------------------------------------------------------------------------
// closed-loop MPC step (python/bsqp/mpc_controller.py:233-253, 294-309)
// ------------------------------------------------------------------------------------------------
namespace {
// in = x_curr[nx] | ref[6N] | x_last[nx] | u_last[nu]
__global__ void k_mpc_prepare(int B, int nx, int N, int traj, const float* __restrict__ in, const float* __restrict__ off, float* xs, float* ref, float* xu)
{
        const int b = blockIdx.x;
        for (int i = threadIdx.x; i < nx; i += blockDim.x) {
                float v = in[i];
                if (off) v = v + off[(size_t)b * nx + i];
                xs[(size_t)b * nx + i] = v;
                xu[(size_t)b * traj + i] = v;
        }
        for (int i = threadIdx.x; i < 6 * N; i += blockDim.x) ref[(size_t)b * 6 * N + i] = in[nx + i];
}
// err[b] = sqrt(sum_i (double(xnext[b][i]) - double(x_curr[i]))^2) with numpy's pairwise order for a contiguous row of n < 128 doubles
// (eight running sums over the first 8*(n/8) entries, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail in order);
// best = first index of the minimum (np.argmin).  One CTA.
__global__ void k_mpc_score(int B, int nx, const float* __restrict__ xnext, const float* __restrict__ in, double* err, int* best)
{
        __shared__ double s_val[256];
        __shared__ int    s_idx[256];
        double            bv = 0.0;
        int               bi = -1;
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
                const float* x = xnext + (size_t)b * nx;
                double       res;
                auto         sq = [&](int i) {
                        const double d = (double)x[i] - (double)in[i];
                        return d * d;
                };
                if (nx < 8) {
                        res = 0.0;
                        for (int i = 0; i < nx; i++) res += sq(i);
                } else {
                        double r[8];
                        for (int j = 0; j < 8; j++) r[j] = sq(j);
                        int i = 8;
                        for (; i < nx - (nx % 8); i += 8)
                                for (int j = 0; j < 8; j++) r[j] += sq(i + j);
                        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
                        for (; i < nx; i++) res += sq(i);
                }
                const double e = sqrt(res);
                err[b] = e;
                // np.argmin: the first minimum; a NaN counts as smaller than everything (the first NaN wins).  ascending b per thread.
                if (bi < 0 || (bv == bv && (e != e || e < bv))) bv = e, bi = b;
        }
        s_val[threadIdx.x] = bv;
        s_idx[threadIdx.x] = bi;
        __syncthreads();
        if (threadIdx.x == 0) {
                double v = 0.0;
                int    id = -1;
                for (int t = 0; t < (int)blockDim.x; t++) {
                        if (s_idx[t] < 0) continue;
                        const double x = s_val[t];
                        const bool   xn = x != x, vn = v != v;
                        bool         take;
                        if (id < 0)
                                take = true;
                        else if (xn || vn)
                                take = xn && (!vn || s_idx[t] < id);  // NaN beats numbers; among NaNs the smallest index
                        else
                                take = x < v || (x == v && s_idx[t] < id);
                        if (take) v = x, id = s_idx[t];
                }
                best[0] = id < 0 ? 0 : id;
        }
}
// Winner record of one shard (floats): [0..1] error (double) | [2] local id (int) | [3] unused | [4 .. 4+traj) the winner's trajectory
__host__ __device__ inline int mpc_record_floats(int traj) { return (4 + traj + 3) / 4 * 4; }
__global__ void k_mpc_record(int traj, const float* __restrict__ xu, const double* __restrict__ err, const int* __restrict__ best, float* rec)
{
        const int id = best[0];
        if (threadIdx.x == 0) {
                *reinterpret_cast<double*>(rec) = err[id];
                reinterpret_cast<int*>(rec)[2] = id;
                rec[3] = 0.0f;
        }
        const float* src = xu + (size_t)id * traj;
        for (int i = threadIdx.x; i < traj; i += blockDim.x) rec[4 + i] = src[i];
}
// Global winner among n shard records (np.argmin over the concatenated error vector: the first minimum, the first NaN beats every number;
// each record already is its shard's first minimum, so ties go to the lower shard), then XU[:] = winner (mpc_controller.py:252-253).
// Block B exports the winner: trajectory, global id (= shard * id_stride + local id) and error.
__global__ void k_mpc_select_adopt(int B, int traj, float* xu, const float* __restrict__ recs, int n, int stride, int id_stride, float* best_out, int* gid_out, double* gerr_out)
{
        __shared__ int s_win;
        if (threadIdx.x == 0) {
                int    win = 0;
                double v = *reinterpret_cast<const double*>(recs);
                for (int r = 1; r < n; r++) {
                        const double x = *reinterpret_cast<const double*>(recs + (size_t)r * stride);
                        if (v == v && (x != x || x < v)) v = x, win = r;  // v is NaN: nothing later can win
                }
                s_win = win;
        }
        __syncthreads();
        const float* rec = recs + (size_t)s_win * stride;
        const int    b = blockIdx.x;
        if (b == B) {
                for (int i = threadIdx.x; i < traj; i += blockDim.x) best_out[i] = rec[4 + i];
                if (threadIdx.x == 0) {
                        gid_out[0] = s_win * id_stride + reinterpret_cast<const int*>(rec)[2];
                        gerr_out[0] = *reinterpret_cast<const double*>(rec);
                }
                return;
        }
        float* d = xu + (size_t)b * traj;
        for (int i = threadIdx.x; i < traj; i += blockDim.x) d[i] = rec[4 + i];
}

int mpc_ensure(gato_solver* s)
{
        if (s->mpc_xu.p) return GATO_OK;
        const Dims&  d = s->d;
        const size_t B = (size_t)s->B;
        CUDA_TRY(s, s->mpc_xu.alloc(B * d.traj));
        CUDA_TRY(s, s->mpc_xs.alloc(B * d.nx));
        CUDA_TRY(s, s->mpc_ref.alloc(B * 6 * d.N));
        CUDA_TRY(s, s->mpc_off.alloc(B * d.nx));
        CUDA_TRY(s, s->mpc_in.alloc(2 * d.nx + 6 * d.N + d.nu));
        CUDA_TRY(s, s->mpc_xnext.alloc(B * d.nx + d.traj));  // x_next batch, then the exported winner
        CUDA_TRY(s, s->mpc_err.alloc(B));
        CUDA_TRY(s, s->mpc_best.alloc(2));  // [0] local winner, [1] global winner id
        CUDA_TRY(s, s->mpc_rec.alloc(mpc_record_floats(d.traj)));
        CUDA_TRY(s, s->mpc_gerr.alloc(1));
        CUDA_TRY(s, cudaMallocHost((void**)&s->h_mpc_in, sizeof(float) * (2 * d.nx + 6 * d.N + d.nu)));
        CUDA_TRY(s, cudaMallocHost((void**)&s->h_mpc_best, sizeof(float) * d.traj));
        CUDA_TRY(s, cudaMallocHost((void**)&s->h_mpc_err, sizeof(double) * B));
        CUDA_TRY(s, cudaMallocHost((void**)&s->h_mpc_id, sizeof(int)));
        CUDA_TRY(s, cudaMallocHost((void**)&s->h_mpc_gerr, sizeof(double)));
        return GATO_OK;
}
}  // namespace

int gato_mpc_set_warm_start(gato_solver* s, const float* h_xu, int per_solve)
{
        if (!s || !h_xu) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        if (int rc = mpc_ensure(s)) return rc;
        const size_t traj = s->d.traj;
        if (per_solve) {
                CUDA_TRY(s, cudaMemcpyAsync(s->mpc_xu.p, h_xu, sizeof(float) * s->B * traj, cudaMemcpyHostToDevice, s->stream));
        } else {
                for (int b = 0; b < s->B; b++) CUDA_TRY(s, cudaMemcpyAsync(s->mpc_xu.p + b * traj, h_xu, sizeof(float) * traj, cudaMemcpyHostToDevice, s->stream));
        }
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        return GATO_OK;
}

int gato_mpc_set_state_offsets(gato_solver* s, const float* h_x_offset)
{
        if (!s) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        if (int rc = mpc_ensure(s)) return rc;
        s->mpc_has_off = h_x_offset != nullptr;
        if (h_x_offset) {
                CUDA_TRY(s, cudaMemcpyAsync(s->mpc_off.p, h_x_offset, sizeof(float) * (size_t)s->B * s->d.nx, cudaMemcpyHostToDevice, s->stream));
                CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        }
        return GATO_OK;
}

int gato_mpc_get_warm_start(gato_solver* s, float* h_xu)
{
        if (!s || !h_xu) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        if (!s->mpc_xu.p) {
                s->err = "gato_mpc_get_warm_start: no warm start has been set";
                return GATO_ERR_ARG;
        }
        CUDA_TRY(s, cudaMemcpyAsync(h_xu, s->mpc_xu.p, sizeof(float) * (size_t)s->B * s->d.traj, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        return GATO_OK;
}

int gato_mpc_record_floats(const gato_solver* s) { return s ? mpc_record_floats(s->d.traj) : GATO_ERR_ARG; }
int gato_mpc_input_floats(const gato_solver* s) { return s ? 2 * s->d.nx + 6 * s->d.N + s->d.nu : GATO_ERR_ARG; }

int gato_mpc_local_async(gato_solver* s, const float* d_in, int score, float sim_dt, float timestep, int flags, float* d_record)
{
        if (!s || !d_in || !d_record) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        if (!s->mpc_xu.p) {
                s->err = "gato_mpc_step: call gato_mpc_set_warm_start first";
                return GATO_ERR_ARG;
        }
        const Dims& d = s->d;
        const int   B = s->B;
        s->t_start = std::chrono::high_resolution_clock::now();
        CUDA_TRY(s, cudaEventRecord(s->ev0, s->stream));
        k_mpc_prepare<<<B, 64, 0, s->stream>>>(B, d.nx, d.N, d.traj, d_in, s->mpc_has_off ? s->mpc_off.p : nullptr, s->mpc_xs.p, s->mpc_ref.p, s->mpc_xu.p);
        s->launches++;
        if (flags & GATO_MPC_RESET_RHO) {
                CUDA_TRY(s, cudaMemcpyAsync(s->rho.p, s->rho_init.p, sizeof(float) * B, cudaMemcpyDeviceToDevice, s->stream));
                CUDA_TRY(s, cudaMemcpyAsync(s->drho.p, s->drho_init.p, sizeof(float) * B, cudaMemcpyDeviceToDevice, s->stream));
        }
        if (int rc = dispatch_enqueue(s, s->mpc_xu.p, s->mpc_xs.p, s->mpc_ref.p, timestep)) return rc;
        if (score) {
                const float* d_xl = d_in + d.nx + 6 * d.N;
                with_plant(s->plant, d.nq, [&](auto P) { enqueue_sim_forward<decltype(P)>(B, s->mpc_xnext.p, d_xl, d_xl + d.nx, s->fext.p, sim_dt, s->model_slot, s->stream); });
                k_mpc_score<<<1, 256, 0, s->stream>>>(B, d.nx, s->mpc_xnext.p, d_in, s->mpc_err.p, s->mpc_best.p);
                s->launches += 2;
        } else {
                CUDA_TRY(s, cudaMemsetAsync(s->mpc_best.p, 0, sizeof(int), s->stream));
                CUDA_TRY(s, cudaMemsetAsync(s->mpc_err.p, 0, sizeof(double) * B, s->stream));
        }
        k_mpc_record<<<1, 128, 0, s->stream>>>(d.traj, s->mpc_xu.p, s->mpc_err.p, s->mpc_best.p, d_record);
        s->launches++;
        CUDA_TRY(s, cudaGetLastError());
        s->mpc_scored = score != 0;
        s->pending = true;
        return GATO_OK;
}

int gato_mpc_adopt_async(gato_solver* s, const float* d_records, int n_records, int record_stride, int id_stride)
{
        if (!s || !d_records || n_records < 1 || record_stride < mpc_record_floats(s->d.traj)) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        const Dims& d = s->d;
        const int   B = s->B;
        float*      d_best_out = s->mpc_xnext.p + (size_t)B * d.nx;
        k_mpc_select_adopt<<<B + 1, 128, 0, s->stream>>>(B, d.traj, s->mpc_xu.p, d_records, n_records, record_stride, id_stride, d_best_out, s->mpc_best.p + 1, s->mpc_gerr.p);
        s->launches++;
        CUDA_TRY(s, cudaEventRecord(s->ev1, s->stream));
        CUDA_TRY(s, cudaGetLastError());
        CUDA_TRY(s, cudaMemcpyAsync(s->h_mpc_best, d_best_out, sizeof(float) * d.traj, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->h_mpc_err, s->mpc_err.p, sizeof(double) * B, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->h_mpc_id, s->mpc_best.p + 1, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(s, cudaMemcpyAsync(s->h_mpc_gerr, s->mpc_gerr.p, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        return GATO_OK;
}

int gato_mpc_wait(gato_solver* s, gato_mpc_out* out, gato_stats* st)
{
        if (!s || !out) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        s->pending = false;
        fill_stats(s, st);
        if (st) {
                st->solve_time_us = std::chrono::duration<double, std::micro>(std::chrono::high_resolution_clock::now() - s->t_start).count();
                float ms = 0;
                CUDA_TRY(s, cudaEventElapsedTime(&ms, s->ev0, s->ev1));
                st->device_time_ms = ms;
        }
        out->best_id = *s->h_mpc_id;
        out->best_error = *s->h_mpc_gerr;
        out->errors = s->h_mpc_err;
        out->xu_best = s->h_mpc_best;
        return GATO_OK;
}

int gato_mpc_step(gato_solver* s, const float* h_x_curr, const float* h_ref_window, const float* h_x_last, const float* h_u_last, float sim_dt, float timestep, int flags,
                  gato_mpc_out* out, gato_stats* st)
{
        if (!s || !h_x_curr || !h_ref_window || !out || ((h_x_last == nullptr) != (h_u_last == nullptr))) return GATO_ERR_ARG;
        if (check_dev(s)) return GATO_ERR_CUDA;
        if (!s->mpc_xu.p) {
                s->err = "gato_mpc_step: call gato_mpc_set_warm_start first";
                return GATO_ERR_ARG;
        }
        const Dims& d = s->d;
        const int   nin = 2 * d.nx + 6 * d.N + d.nu;
        const bool  score = h_x_last != nullptr;
        memcpy(s->h_mpc_in, h_x_curr, sizeof(float) * d.nx);
        memcpy(s->h_mpc_in + d.nx, h_ref_window, sizeof(float) * 6 * d.N);
        if (score) {
                memcpy(s->h_mpc_in + d.nx + 6 * d.N, h_x_last, sizeof(float) * d.nx);
                memcpy(s->h_mpc_in + 2 * d.nx + 6 * d.N, h_u_last, sizeof(float) * d.nu);
        }
        CUDA_TRY(s, cudaMemcpyAsync(s->mpc_in.p, s->h_mpc_in, sizeof(float) * nin, cudaMemcpyHostToDevice, s->stream));
        // one shard: the local winner is the global one
        if (int rc = gato_mpc_local_async(s, s->mpc_in.p, score ? 1 : 0, sim_dt, timestep, flags, s->mpc_rec.p)) return rc;
        if (int rc = gato_mpc_adopt_async(s, s->mpc_rec.p, 1, mpc_record_floats(d.traj), s->B)) return rc;
        return gato_mpc_wait(s, out, st);
}

// ------------------------------------------------------------------------------------------------
// stage-level entry points: run ONE product kernel (or kernel phase) on host buffers, for parity tests
// ------------------------------------------------------------------------------------------------
namespace {
struct StageSolver {
        gato_solver* s = nullptr;
        StageSolver(int plant, int N, int B, const float* cost7, int max_pcg = 0)
        {
                gato_params p{};
                p.dt = 0.01f, p.max_sqp_iters = 1, p.max_pcg_iters = (uint32_t)max_pcg, p.solve_ratio = 1.0f, p.mu = 10.0f, p.rho = 0.0f;
                if (cost7) p.q_cost = cost7[0], p.qd_cost = cost7[1], p.u_cost = cost7[2], p.N_cost = cost7[3], p.q_lim_cost = cost7[4], p.vel_lim_cost = cost7[5], p.ctrl_lim_cost = cost7[6];
                gato_create(&s, plant, N, B, 0, nullptr, &p);
        }
        ~StageSolver() { gato_destroy(s); }
        bool up(DevArr<float>& a, const float* h) { return cudaMemcpyAsync(a.p, h, sizeof(float) * a.n, cudaMemcpyHostToDevice, s->stream) == cudaSuccess; }
        bool down(float* h, DevArr<float>& a) { return cudaMemcpyAsync(h, a.p, sizeof(float) * a.n, cudaMemcpyDeviceToHost, s->stream) == cudaSuccess; }
        int  finish()
        {
                cudaError_t e = cudaStreamSynchronize(s->stream);
                if (e == cudaSuccess) e = cudaGetLastError();
                if (e != cudaSuccess) {
                        fprintf(stderr, "gato stage: %s\n", cudaGetErrorString(e));
                        return GATO_ERR_CUDA;
                }
                return GATO_OK;
        }
};
#define PLANT_CALL(plant, fn, ...) with_plant((plant), s->d.nq, [&](auto P_) { fn<decltype(P_)>(__VA_ARGS__); })
}  // namespace

int gato_stage_kkt(int plant, int N, int B, const float* xu, const float* xs, const float* ref, const float* fext, float dt, const float* cost7, float* Q, float* R, float* q, float* r, float* A,
                   float* Bm, float* c)
{
        StageSolver t(plant, N, B, cost7);
        if (!t.s) return GATO_ERR_CUDA;
        gato_solver* s = t.s;
        t.up(s->st_xu, xu), t.up(s->st_xs, xs), t.up(s->st_ref, ref), t.up(s->fext, fext);
        Ctx ctx = make_ctx(s, s->st_xu.p, s->st_xs.p, s->st_ref.p, dt);
        PLANT_CALL(plant, launch_kkt, s, ctx);
        t.down(Q, s->Q), t.down(R, s->R), t.down(q, s->q), t.down(r, s->r), t.down(A, s->A), t.down(Bm, s->Bm), t.down(c, s->c);
        return t.finish();
}

int gato_stage_schur(int plant, int N, int B, float* Q, float* R, const float* q, const float* r, const float* A, const float* Bm, const float* c, const float* rho, float* S, float* Pinv,
                     float* gamma)
{
        StageSolver t(plant, N, B, nullptr);
        if (!t.s) return GATO_ERR_CUDA;
        gato_solver* s = t.s;
        t.up(s->Q, Q), t.up(s->R, R), t.up(s->q, q), t.up(s->r, r), t.up(s->A, A), t.up(s->Bm, Bm), t.up(s->c, c), t.up(s->rho, rho);
        Ctx ctx = make_ctx(s, s->st_xu.p, s->st_xs.p, s->st_ref.p, 0.01f);
        PLANT_CALL(plant, launch_schur, s, ctx);
        ctx.flags = F_K2 | F_WRITE_P;
        PLANT_CALL(plant, launch_pcg, s, ctx);
        t.down(S, s->S), t.down(Pinv, s->Pinv), t.down(gamma, s->gamma), t.down(Q, s->Qinv), t.down(R, s->Rinv);
        return t.finish();
}

int gato_stage_pcg(int plant, int N, int B, const float* S, const float* Pinv, const float* gamma, float* lambda, const float* eps, int max_iters, const int* kkt_conv, int* iters)
{
        StageSolver t(plant, N, B, nullptr, max_iters);
        if (!t.s) return GATO_ERR_CUDA;
        gato_solver* s = t.s;
        t.up(s->S, S), t.up(s->Pinv, Pinv), t.up(s->gamma, gamma), t.up(s->lambda, lambda), t.up(s->pcg_tol, eps);
        cudaMemcpyAsync(s->conv.p, kkt_conv, sizeof(int) * B, cudaMemcpyHostToDevice, s->stream);
        Ctx ctx = make_ctx(s, s->st_xu.p, s->st_xs.p, s->st_ref.p, 0.01f);
        ctx.flags = F_PCG;
        PLANT_CALL(plant, launch_pcg, s, ctx);
        t.down(lambda, s->lambda);
        cudaMemcpyAsync(iters, s->pcg_log.p, sizeof(int) * B, cudaMemcpyDeviceToHost, s->stream);
        return t.finish();
}

int gato_stage_dz(int plant, int N, int B, const float* lambda, const float* Qinv, const float* Rinv, float* q, float* r, const float* A, const float* Bm, float* dz)
{
        StageSolver t(plant, N, B, nullptr);
        if (!t.s) return GATO_ERR_CUDA;
        gato_solver* s = t.s;
        t.up(s->lambda, lambda), t.up(s->Qinv, Qinv), t.up(s->Rinv, Rinv), t.up(s->q, q), t.up(s->r, r), t.up(s->A, A), t.up(s->Bm, Bm);
        Ctx ctx = make_ctx(s, s->st_xu.p, s->st_xs.p, s->st_ref.p, 0.01f);
        ctx.flags = F_DZ;
        PLANT_CALL(plant, launch_pcg, s, ctx);
        t.down(dz, s->dz), t.down(q, s->q), t.down(r, s->r);
        return t.finish();
}

int gato_stage_merit(int plant, int N, int B, const float* xu, const float* dz, const float* xs, const float* ref, const float* mu, const float* fext, float dt, const float* cost7, int num_alphas,
                     float* merit)
{
        StageSolver t(plant, N, B, cost7);
        if (!t.s) return GATO_ERR_CUDA;
        gato_solver* s = t.s;
        t.up(s->st_xu, xu), t.up(s->dz, dz), t.up(s->st_xs, xs), t.up(s->st_ref, ref), t.up(s->mu, mu), t.up(s->fext, fext);
        Ctx ctx = make_ctx(s, s->st_xu.p, s->st_xs.p, s->st_ref.p, dt);
        ctx.flags = F_MERIT;
        ctx.ls_merit_log = nullptr;
        if (num_alphas == 1) {
                with_plant(plant, s->d.nq, [&](auto P) { launch_merit<decltype(P), 1>(s, ctx); });
                cudaMemcpyAsync(merit, s->merit_cur.p, sizeof(float) * B, cudaMemcpyDeviceToHost, s->stream);
        } else {
                with_plant(plant, s->d.nq, [&](auto P) { launch_merit<decltype(P), kNumAlphas>(s, ctx); });
                t.down(merit, s->merit);
        }
        return t.finish();
}

int gato_stage_linesearch(int plant, int N, int B, float* xu, const float* dz, const float* merit8, float* merit_init, float* step, float* rho, float* drho, int adapt)
{
        StageSolver t(plant, N, B, nullptr);
        if (!t.s) return GATO_ERR_CUDA;
        gato_solver* s = t.s;
        s->adapt_rho = adapt != 0;
        t.up(s->st_xu, xu), t.up(s->dz, dz), t.up(s->merit, merit8), t.up(s->merit_cur, merit_init), t.up(s->rho, rho), t.up(s->drho, drho);
        Ctx ctx = make_ctx(s, s->st_xu.p, s->st_xs.p, s->st_ref.p, 0.01f);
        ctx.flags = F_LS;
        ctx.ls_merit_log = nullptr;
        with_plant(plant, s->d.nq, [&](auto P) { launch_merit<decltype(P), kNumAlphas>(s, ctx); });
        t.down(xu, s->st_xu), t.down(merit_init, s->merit_cur), t.down(step, s->step), t.down(rho, s->rho), t.down(drho, s->drho);
        return t.finish();
}

}  // extern "C"
