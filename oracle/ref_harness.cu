// TEST INFRASTRUCTURE — reference harness (NOT part of the product).
//
// Compiles the UNMODIFIED reference headers where they lie under /root/reference/gato
// (never copied into this repo) into a small C-ABI shared library, one per
// (plant, KNOT_POINTS, math-mode), so that golden vectors can be minted on the B200 box:
//   * whole-solve entry points driving BSQP<float,B>::solve exactly like
//     /root/reference/python/bindings.cu:68-148 does, and
//   * per-stage entry points that call the reference's own host launchers
//     (setupKKTSystemBatched setup_kkt.cuh:130, formSchurSystemBatched schur_linsys.cuh:295,
//      solvePCGBatched pcg.cuh:157, computeDzBatched schur_linsys.cuh:445,
//      computeMeritBatched merit.cuh:103, lineSearchAndUpdateBatched line_search.cuh:100,
//      simForwardBatched sim.cuh:66).
// Built by oracle/build_ref.sh into oracle/_ref/ (git-ignored, travels with gpurun).
// Only tests/, bench.py (reference rows) and oracle/gen_golden.py load it.
//
// Compile-time: -DKNOT_POINTS=<N> and one of -DPLANT_IIWA14=1 / -DPLANT_INDY7=1,
// -DGREF_BATCHES="1,64,128" (comma list of BatchSize instantiations).

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#include "bsqp/bsqp.cuh"
#include "types.cuh"
#include "utils/cuda.cuh"

#ifndef GREF_BATCHES
#define GREF_BATCHES 1
#endif

namespace {

constexpr uint32_t kNX = STATE_SIZE;
constexpr uint32_t kNU = CONTROL_SIZE;
constexpr uint32_t kN = KNOT_POINTS;

#define CK(x)                                                                                  \
        do {                                                                                   \
                cudaError_t e_ = (x);                                                          \
                if (e_ != cudaSuccess) {                                                       \
                        fprintf(stderr, "gref CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
                        return -1;                                                             \
                }                                                                              \
        } while (0)

struct DevBuf {
        float* p = nullptr;
        size_t n = 0;
        DevBuf() {}
        explicit DevBuf(size_t count) { alloc(count); }
        void alloc(size_t count)
        {
                n = count;
                cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(float));
                cudaMemset(p, 0, std::max<size_t>(count, 1) * sizeof(float));
        }
        ~DevBuf()
        {
                if (p) cudaFree(p);
        }
        int up(const float* h)
        {
                CK(cudaMemcpy(p, h, n * sizeof(float), cudaMemcpyHostToDevice));
                return 0;
        }
        int down(float* h)
        {
                CK(cudaMemcpy(h, p, n * sizeof(float), cudaMemcpyDeviceToHost));
                return 0;
        }
        DevBuf(const DevBuf&) = delete;
        DevBuf& operator=(const DevBuf&) = delete;
};

// ---------------------------------------------------------------------------------------------
// whole-solve object (type-erased over BatchSize)
// ---------------------------------------------------------------------------------------------
struct ISolver {
        virtual ~ISolver() {}
        virtual uint32_t batch() const = 0;
        virtual int      set_batch(int which, const float* h, int set_default) = 0;
        virtual int      reset(int which) = 0;
        virtual void     set_rho_adaptation(int on) = 0;
        virtual int      solve(float* h_xu, const float* h_xs, const float* h_ref, float dt, int* sqp_iters, int* kkt_conv, int* n_pcg, int* n_ls, int* pcg_iters, float* ls_min_merit,
                               float* ls_step, int cap_iters, float* final_merit, float* initial_merit, double* solve_time_us, float* event_ms) = 0;
        virtual int      solve_timed(const float* h_xu, const float* h_xs, const float* h_ref, float dt, int reps, int reset_between, float* event_ms_each, double* ref_us_each) = 0;
        virtual int      sim_forward(const float* h_xk, const float* h_uk, float dt, float* h_xkp1) = 0;
};

template<uint32_t B>
struct Solver final : ISolver {
        BSQP<float, B>* s;
        DevBuf          d_xu, d_xs, d_ref, d_xkp1, d_xk, d_uk;
        cudaEvent_t     e0, e1;

        explicit Solver(const float* p)
            : d_xu(size_t(TRAJ_SIZE) * B), d_xs(size_t(kNX) * B), d_ref(size_t(REFERENCE_TRAJ_SIZE) * B), d_xkp1(size_t(kNX) * B), d_xk(kNX), d_uk(kNU)
        {
                // ctor order: bsqp.cuh:43
                s = new BSQP<float, B>(p[0], (uint32_t)p[1], p[2], (uint32_t)p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10], p[11], p[12], p[13], p[14]);
                setL2PersistingAccess(1.0);  // bindings.cu:45
                cudaEventCreate(&e0);
                cudaEventCreate(&e1);
        }
        ~Solver() override
        {
                delete s;
                cudaEventDestroy(e0);
                cudaEventDestroy(e1);
        }
        uint32_t batch() const override { return B; }

        int set_batch(int which, const float* h, int set_default) override
        {
                switch (which) {
                        case 0: s->set_f_ext_batch(const_cast<float*>(h)); break;
                        case 1: s->set_rho_penalty_batch(h, set_default != 0); break;
                        case 2: s->set_drho_batch(h, set_default != 0); break;
                        case 3: s->set_mu_batch(h); break;
                        case 4: s->set_pcg_tol_batch(h); break;
                        default: return -2;
                }
                CK(cudaDeviceSynchronize());
                return 0;
        }
        int reset(int which) override
        {
                if (which == 0)
                        s->reset_dual();
                else if (which == 1)
                        s->reset_rho();
                else
                        return -2;
                CK(cudaDeviceSynchronize());
                return 0;
        }
        void set_rho_adaptation(int on) override { s->set_rho_adaptation(on != 0); }

        int solve(float* h_xu, const float* h_xs, const float* h_ref, float dt, int* sqp_iters, int* kkt_conv, int* n_pcg, int* n_ls, int* pcg_iters, float* ls_min_merit, float* ls_step,
                  int cap_iters, float* final_merit, float* initial_merit, double* solve_time_us, float* event_ms) override
        {
                if (d_xu.up(h_xu) || d_xs.up(h_xs) || d_ref.up(h_ref)) return -1;
                ProblemInputs<float, B> in;
                in.timestep = dt;
                in.d_x_s_batch = d_xs.p;
                in.d_reference_traj_batch = d_ref.p;
                in.d_GRiD_mem = nullptr;
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0));
                SQPStats<float, B> st = s->solve(d_xu.p, in);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (event_ms) *event_ms = ms;
                if (solve_time_us) *solve_time_us = st.solve_time_us;
                if (d_xu.down(h_xu)) return -1;
                s->copy_final_merit_to_host(final_merit);
                s->copy_initial_merit0_to_host(initial_merit);
                for (uint32_t b = 0; b < B; b++) {
                        sqp_iters[b] = st.sqp_iterations[b];
                        kkt_conv[b] = st.kkt_converged[b];
                }
                *n_pcg = (int)st.pcg_stats.size();
                *n_ls = (int)st.line_search_stats.size();
                for (int i = 0; i < *n_pcg && i < cap_iters; i++)
                        for (uint32_t b = 0; b < B; b++) pcg_iters[size_t(i) * B + b] = st.pcg_stats[i].num_iterations[b];
                for (int i = 0; i < *n_ls && i < cap_iters; i++)
                        for (uint32_t b = 0; b < B; b++) {
                                ls_min_merit[size_t(i) * B + b] = st.line_search_stats[i].min_merit[b];
                                ls_step[size_t(i) * B + b] = st.line_search_stats[i].step_size[b];
                        }
                return 0;
        }

        // Repeated solves of the same problem for timing: inputs are re-uploaded OUTSIDE the timed
        // region, the timed region is BSQP::solve only (== the reference's own sqp_time_us window,
        // bsqp.cuh:109-190). reset_between: 1 -> reset_dual + reset_rho before each rep (cold start).
        int solve_timed(const float* h_xu, const float* h_xs, const float* h_ref, float dt, int reps, int reset_between, float* event_ms_each, double* ref_us_each) override
        {
                if (d_xs.up(h_xs) || d_ref.up(h_ref)) return -1;
                ProblemInputs<float, B> in;
                in.timestep = dt;
                in.d_x_s_batch = d_xs.p;
                in.d_reference_traj_batch = d_ref.p;
                in.d_GRiD_mem = nullptr;
                for (int r = 0; r < reps; r++) {
                        if (d_xu.up(h_xu)) return -1;
                        if (reset_between) {
                                s->reset_dual();
                                s->reset_rho();
                        }
                        CK(cudaDeviceSynchronize());
                        CK(cudaEventRecord(e0));
                        SQPStats<float, B> st = s->solve(d_xu.p, in);
                        CK(cudaEventRecord(e1));
                        CK(cudaEventSynchronize(e1));
                        float ms = 0;
                        CK(cudaEventElapsedTime(&ms, e0, e1));
                        event_ms_each[r] = ms;
                        ref_us_each[r] = st.solve_time_us;
                }
                return 0;
        }

        int sim_forward(const float* h_xk, const float* h_uk, float dt, float* h_xkp1) override
        {
                // bindings.cu:180-194
                if (d_xk.up(h_xk) || d_uk.up(h_uk)) return -1;
                s->sim_forward(d_xkp1.p, d_xk.p, d_uk.p, dt);
                CK(cudaDeviceSynchronize());
                return d_xkp1.down(h_xkp1);
        }
};

// ---------------------------------------------------------------------------------------------
// per-stage drivers (templated on BatchSize because the reference launchers are)
// ---------------------------------------------------------------------------------------------
void* g_grid_mem = nullptr;
void* grid_mem()
{
        if (!g_grid_mem) g_grid_mem = gato::plant::initializeDynamicsConstMem<float>();
        return g_grid_mem;
}

template<uint32_t B>
int stage_kkt(const float* xu, const float* xs, const float* ref, const float* fext, float dt, const float* cost7, float* Q, float* R, float* q, float* r, float* A, float* Bm, float* c)
{
        DevBuf dxu(size_t(TRAJ_SIZE) * B), dxs(size_t(kNX) * B), dref(size_t(REFERENCE_TRAJ_SIZE) * B), dfe(6 * size_t(B));
        DevBuf dQ(size_t(STATE_SQ_P_KNOTS) * B), dR(size_t(CONTROL_SQ_P_KNOTS) * B), dq(size_t(STATE_P_KNOTS) * B), dr(size_t(CONTROL_P_KNOTS) * B), dA(size_t(STATE_SQ_P_KNOTS) * B),
            dB(size_t(STATE_P_CONTROL_P_KNOTS) * B), dc(size_t(STATE_P_KNOTS) * B);
        if (dxu.up(xu) || dxs.up(xs) || dref.up(ref) || dfe.up(fext)) return -1;
        KKTSystem<float, B> k{dQ.p, dR.p, dq.p, dr.p, dA.p, dB.p, dc.p};
        ProblemInputs<float, B> in;
        in.timestep = dt;
        in.d_x_s_batch = dxs.p;
        in.d_reference_traj_batch = dref.p;
        in.d_GRiD_mem = nullptr;
        setupKKTSystemBatched<float, B>(k, in, dxu.p, dfe.p, grid_mem(), cost7[0], cost7[1], cost7[2], cost7[3], cost7[4], cost7[5], cost7[6]);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        if (dQ.down(Q) || dR.down(R) || dq.down(q) || dr.down(r) || dA.down(A) || dB.down(Bm) || dc.down(c)) return -1;
        return 0;
}

template<uint32_t B>
int stage_schur(float* Q, float* R, const float* q, const float* r, const float* A, const float* Bm, const float* c, const float* rho, float* S, float* Pinv, float* gamma)
{
        DevBuf dQ(size_t(STATE_SQ_P_KNOTS) * B), dR(size_t(CONTROL_SQ_P_KNOTS) * B), dq(size_t(STATE_P_KNOTS) * B), dr(size_t(CONTROL_P_KNOTS) * B), dA(size_t(STATE_SQ_P_KNOTS) * B),
            dB(size_t(STATE_P_CONTROL_P_KNOTS) * B), dc(size_t(STATE_P_KNOTS) * B), drho(B);
        DevBuf dS(size_t(B3D_MATRIX_SIZE_PADDED) * B), dP(size_t(B3D_MATRIX_SIZE_PADDED) * B), dg(size_t(VEC_SIZE_PADDED) * B);
        if (dQ.up(Q) || dR.up(R) || dq.up(q) || dr.up(r) || dA.up(A) || dB.up(Bm) || dc.up(c) || drho.up(rho)) return -1;
        KKTSystem<float, B>   k{dQ.p, dR.p, dq.p, dr.p, dA.p, dB.p, dc.p};
        SchurSystem<float, B> sc{dS.p, dP.p, dg.p};
        formSchurSystemBatched<float, B>(sc, k, drho.p);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        if (dS.down(S) || dP.down(Pinv) || dg.down(gamma) || dQ.down(Q) || dR.down(R)) return -1;
        return 0;
}

template<uint32_t B>
int stage_pcg(const float* S, const float* Pinv, const float* gamma, float* lambda, const float* eps, int max_iters, const int* kkt_conv, int* iters)
{
        DevBuf dS(size_t(B3D_MATRIX_SIZE_PADDED) * B), dP(size_t(B3D_MATRIX_SIZE_PADDED) * B), dg(size_t(VEC_SIZE_PADDED) * B), dl(size_t(VEC_SIZE_PADDED) * B), de(B);
        if (dS.up(S) || dP.up(Pinv) || dg.up(gamma) || dl.up(lambda) || de.up(eps)) return -1;
        int32_t*  dconv;
        uint32_t* dit;
        CK(cudaMalloc(&dconv, sizeof(int32_t) * B));
        CK(cudaMalloc(&dit, sizeof(uint32_t) * B));
        CK(cudaMemcpy(dconv, kkt_conv, sizeof(int32_t) * B, cudaMemcpyHostToDevice));
        CK(cudaMemset(dit, 0, sizeof(uint32_t) * B));
        SchurSystem<float, B> sc{dS.p, dP.p, dg.p};
        solvePCGBatched<float, B>(dl.p, sc, de.p, (uint32_t)max_iters, dconv, dit);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(iters, dit, sizeof(uint32_t) * B, cudaMemcpyDeviceToHost));
        cudaFree(dconv);
        cudaFree(dit);
        return dl.down(lambda);
}

template<uint32_t B>
int stage_dz(const float* lambda, const float* Qinv, const float* Rinv, float* q, float* r, const float* A, const float* Bm, float* dz)
{
        DevBuf dQ(size_t(STATE_SQ_P_KNOTS) * B), dR(size_t(CONTROL_SQ_P_KNOTS) * B), dq(size_t(STATE_P_KNOTS) * B), dr(size_t(CONTROL_P_KNOTS) * B), dA(size_t(STATE_SQ_P_KNOTS) * B),
            dB(size_t(STATE_P_CONTROL_P_KNOTS) * B), dc(1), dl(size_t(VEC_SIZE_PADDED) * B), ddz(size_t(TRAJ_SIZE) * B);
        if (dQ.up(Qinv) || dR.up(Rinv) || dq.up(q) || dr.up(r) || dA.up(A) || dB.up(Bm) || dl.up(lambda)) return -1;
        KKTSystem<float, B> k{dQ.p, dR.p, dq.p, dr.p, dA.p, dB.p, dc.p};
        computeDzBatched<float, B>(ddz.p, dl.p, k);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        if (ddz.down(dz) || dq.down(q) || dr.down(r)) return -1;
        return 0;
}

template<uint32_t B>
int stage_merit(const float* xu, const float* dz, const float* xs, const float* ref, const float* mu, const float* fext, float dt, const float* cost7, int num_alphas, float* merit)
{
        DevBuf dxu(size_t(TRAJ_SIZE) * B), ddz(size_t(TRAJ_SIZE) * B), dxs(size_t(kNX) * B), dref(size_t(REFERENCE_TRAJ_SIZE) * B), dmu(B), dfe(6 * size_t(B)), dm(size_t(NUM_ALPHAS) * B);
        if (dxu.up(xu) || ddz.up(dz) || dxs.up(xs) || dref.up(ref) || dmu.up(mu) || dfe.up(fext)) return -1;
        ProblemInputs<float, B> in;
        in.timestep = dt;
        in.d_x_s_batch = dxs.p;
        in.d_reference_traj_batch = dref.p;
        in.d_GRiD_mem = nullptr;
        if (getenv("GREF_MERIT_EXTRA_SMEM")) {
                // The reference's own launcher (merit.cuh:110-141) requests too little dynamic shared memory for indy7 (and for iiwa14 at
                // N = 128): the kernel writes past it and faults on B200.  With this switch the harness launches the SAME, unmodified kernel
                // with the launcher's grid and arguments but with extra shared memory, so that the stage can be compared at all.
                const int    na = num_alphas == 1 ? 1 : NUM_ALPHAS;
                const size_t smem = getComputeMeritBatchedSMemSize<float>() + (size_t)atoi(getenv("GREF_MERIT_EXTRA_SMEM"));
                CK(cudaFuncSetAttribute(computeMeritBatchedKernel<float, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                CK(cudaMemset(dm.p, 0, sizeof(float) * B * na));
                computeMeritBatchedKernel<float, B><<<dim3(KNOT_POINTS, B, na), dim3(grid::SUGGESTED_THREADS), smem>>>(
                    dm.p, ddz.p, dxu.p, in.d_x_s_batch, in.d_reference_traj_batch, grid_mem(), dmu.p, dfe.p, in.timestep, cost7[0], cost7[1], cost7[2], cost7[3], cost7[4], cost7[5], cost7[6]);
        } else if (num_alphas == 1)
                computeMeritBatched<float, B, 1>(dm.p, ddz.p, dxu.p, dfe.p, in, dmu.p, grid_mem(), cost7[0], cost7[1], cost7[2], cost7[3], cost7[4], cost7[5], cost7[6]);
        else
                computeMeritBatched<float, B, NUM_ALPHAS>(dm.p, ddz.p, dxu.p, dfe.p, in, dmu.p, grid_mem(), cost7[0], cost7[1], cost7[2], cost7[3], cost7[4], cost7[5], cost7[6]);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        dm.n = size_t(num_alphas == 1 ? 1 : NUM_ALPHAS) * B;
        return dm.down(merit);
}

template<uint32_t B>
int stage_linesearch(float* xu, const float* dz, const float* merit8, float* merit_init, float* step, float* rho, float* drho, int adapt)
{
        DevBuf dxu(size_t(TRAJ_SIZE) * B), ddz(size_t(TRAJ_SIZE) * B), dm(size_t(NUM_ALPHAS) * B), dmi(B), dst(B), drh(B), ddr(B);
        if (dxu.up(xu) || ddz.up(dz) || dm.up(merit8) || dmi.up(merit_init) || drh.up(rho) || ddr.up(drho)) return -1;
        lineSearchAndUpdateBatched<float, B, NUM_ALPHAS>(dxu.p, ddz.p, dm.p, dmi.p, dst.p, drh.p, ddr.p, adapt);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        if (dxu.down(xu) || dmi.down(merit_init) || dst.down(step) || drh.down(rho) || ddr.down(drho)) return -1;
        return 0;
}

// Direct dump of the plant-level dynamics the KKT/merit kernels are built from (debug pin for the
// oracle's RBD restatement): one block per sample.
//   qdd[nq], dqdd[nq*(2nq+nq)] (d qdd/dq | d qdd/dqd | d qdd/du = Minv), ee[6], dee[6*nq]
__global__ void dynDumpKernel(const float* d_x, const float* d_u, const float* d_fext, void* d_grid, float* o_qdd, float* o_dqdd, float* o_ee, float* o_dee)
{
        extern __shared__ float sm[];
        constexpr int       nq = kNX / 2;
        float*              s_q = sm;            // nx
        float*              s_u = s_q + kNX;     // nu
        float*              s_qdd = s_u + kNU;   // nq
        float*              s_dqdd = s_qdd + nq; // nq*3nq
        float*              s_ee = s_dqdd + nq * 3 * nq;
        float*              s_dee = s_ee + 6;
        float*              s_temp = s_dee + 6 * nq;
        const int           b = blockIdx.x;
        for (int i = threadIdx.x; i < (int)kNX; i += blockDim.x) s_q[i] = d_x[b * kNX + i];
        for (int i = threadIdx.x; i < (int)kNU; i += blockDim.x) s_u[i] = d_u[b * kNU + i];
        __syncthreads();
        float* fe = const_cast<float*>(d_fext) + 6 * b;
        gato::plant::forwardDynamicsAndGradient<float>(s_dqdd, s_qdd, s_q, s_q + nq, s_u, s_temp, d_grid, fe);
        __syncthreads();
#if defined(PLANT_IIWA14)
        grid::end_effector_pose_device<float>(s_ee, s_q, s_temp, (grid::robotModel<float>*)d_grid);
        __syncthreads();
        grid::end_effector_pose_gradient_device<float>(s_dee, s_q, s_temp, (grid::robotModel<float>*)d_grid);
#else
        grid::end_effector_positions_device<float>(s_ee, s_q, s_temp, (grid::robotModel<float>*)d_grid);
        __syncthreads();
        grid::end_effector_positions_gradient_device<float>(s_dee, s_q, s_temp, (grid::robotModel<float>*)d_grid);
#endif
        __syncthreads();
        for (int i = threadIdx.x; i < nq; i += blockDim.x) o_qdd[b * nq + i] = s_qdd[i];
        for (int i = threadIdx.x; i < nq * 3 * nq; i += blockDim.x) o_dqdd[b * nq * 3 * nq + i] = s_dqdd[i];
        for (int i = threadIdx.x; i < 6; i += blockDim.x) o_ee[b * 6 + i] = s_ee[i];
        for (int i = threadIdx.x; i < 6 * nq; i += blockDim.x) o_dee[b * 6 * nq + i] = s_dee[i];
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------

extern "C" {

int gref_info(int* out6)
{
        out6[0] = (int)kN;
        out6[1] = (int)kNX;
        out6[2] = (int)kNU;
#if defined(PLANT_IIWA14)
        out6[3] = 1;
#else
        out6[3] = 0;
#endif
        out6[4] = (int)TRAJ_SIZE;
#if defined(GREF_FAST_MATH)
        out6[5] = 1;
#else
        out6[5] = 0;
#endif
        return 0;
}

static const uint32_t kBatches[] = {GREF_BATCHES};

int gref_batches(int* out, int cap)
{
        int n = (int)(sizeof(kBatches) / sizeof(kBatches[0]));
        for (int i = 0; i < n && i < cap; i++) out[i] = (int)kBatches[i];
        return n;
}

}  // extern "C"

// dispatch helper: expands the comma list at compile time through a variadic template
template<uint32_t... Bs>
struct BatchList {};

template<typename F>
int dispatch(uint32_t, BatchList<>, F&&)
{
        fprintf(stderr, "gref: batch size not instantiated in this build\n");
        return -3;
}
template<uint32_t B0, uint32_t... Bs, typename F>
int dispatch(uint32_t b, BatchList<B0, Bs...>, F&& f)
{
        if (b == B0) return f(std::integral_constant<uint32_t, B0>{});
        return dispatch(b, BatchList<Bs...>{}, f);
}
using AllBatches = BatchList<GREF_BATCHES>;

extern "C" {

void* gref_create(int B, const float* params15)
{
        ISolver* out = nullptr;
        dispatch((uint32_t)B, AllBatches{}, [&](auto bc) {
                out = new Solver<decltype(bc)::value>(params15);
                return 0;
        });
        return out;
}
void gref_destroy(void* h) { delete (ISolver*)h; }
int  gref_set_batch(void* h, int which, const float* v, int set_default) { return ((ISolver*)h)->set_batch(which, v, set_default); }
int  gref_reset(void* h, int which) { return ((ISolver*)h)->reset(which); }
void gref_set_rho_adaptation(void* h, int on) { ((ISolver*)h)->set_rho_adaptation(on); }
int  gref_solve(void* h, float* xu, const float* xs, const float* ref, float dt, int* sqp_iters, int* kkt_conv, int* n_pcg, int* n_ls, int* pcg_iters, float* ls_min_merit, float* ls_step, int cap_iters,
                float* final_merit, float* initial_merit, double* solve_time_us, float* event_ms)
{
        return ((ISolver*)h)->solve(xu, xs, ref, dt, sqp_iters, kkt_conv, n_pcg, n_ls, pcg_iters, ls_min_merit, ls_step, cap_iters, final_merit, initial_merit, solve_time_us, event_ms);
}
int gref_solve_timed(void* h, const float* xu, const float* xs, const float* ref, float dt, int reps, int reset_between, float* event_ms_each, double* ref_us_each)
{
        return ((ISolver*)h)->solve_timed(xu, xs, ref, dt, reps, reset_between, event_ms_each, ref_us_each);
}
int gref_sim_forward(void* h, const float* xk, const float* uk, float dt, float* xkp1) { return ((ISolver*)h)->sim_forward(xk, uk, dt, xkp1); }

int gref_stage_kkt(int B, const float* xu, const float* xs, const float* ref, const float* fext, float dt, const float* cost7, float* Q, float* R, float* q, float* r, float* A, float* Bm, float* c)
{
        return dispatch((uint32_t)B, AllBatches{}, [&](auto bc) { return stage_kkt<decltype(bc)::value>(xu, xs, ref, fext, dt, cost7, Q, R, q, r, A, Bm, c); });
}
int gref_stage_schur(int B, float* Q, float* R, const float* q, const float* r, const float* A, const float* Bm, const float* c, const float* rho, float* S, float* Pinv, float* gamma)
{
        return dispatch((uint32_t)B, AllBatches{}, [&](auto bc) { return stage_schur<decltype(bc)::value>(Q, R, q, r, A, Bm, c, rho, S, Pinv, gamma); });
}
int gref_stage_pcg(int B, const float* S, const float* Pinv, const float* gamma, float* lambda, const float* eps, int max_iters, const int* kkt_conv, int* iters)
{
        return dispatch((uint32_t)B, AllBatches{}, [&](auto bc) { return stage_pcg<decltype(bc)::value>(S, Pinv, gamma, lambda, eps, max_iters, kkt_conv, iters); });
}
int gref_stage_dz(int B, const float* lambda, const float* Qinv, const float* Rinv, float* q, float* r, const float* A, const float* Bm, float* dz)
{
        return dispatch((uint32_t)B, AllBatches{}, [&](auto bc) { return stage_dz<decltype(bc)::value>(lambda, Qinv, Rinv, q, r, A, Bm, dz); });
}
int gref_stage_merit(int B, const float* xu, const float* dz, const float* xs, const float* ref, const float* mu, const float* fext, float dt, const float* cost7, int num_alphas, float* merit)
{
        return dispatch((uint32_t)B, AllBatches{}, [&](auto bc) { return stage_merit<decltype(bc)::value>(xu, dz, xs, ref, mu, fext, dt, cost7, num_alphas, merit); });
}
int gref_stage_linesearch(int B, float* xu, const float* dz, const float* merit8, float* merit_init, float* step, float* rho, float* drho, int adapt)
{
        return dispatch((uint32_t)B, AllBatches{}, [&](auto bc) { return stage_linesearch<decltype(bc)::value>(xu, dz, merit8, merit_init, step, rho, drho, adapt); });
}

// n samples of (x[nx], u[nu], fext[6]) -> qdd[nq], dqdd[nq*3nq], ee[6], dee[6*nq]
int gref_dyn_dump(int n, const float* x, const float* u, const float* fext, float* qdd, float* dqdd, float* ee, float* dee)
{
        constexpr int nq = kNX / 2;
        DevBuf        dx(size_t(n) * kNX), du(size_t(n) * kNU), dfe(size_t(n) * 6), dqdd_(size_t(n) * nq), ddq(size_t(n) * nq * 3 * nq), dee_(size_t(n) * 6), ddee(size_t(n) * 6 * nq);
        if (dx.up(x) || du.up(u) || dfe.up(fext)) return -1;
        size_t smem = sizeof(float) * (kNX + kNU + nq + nq * 3 * nq + 6 + 6 * nq + 4096);
        CK(cudaFuncSetAttribute(dynDumpKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dynDumpKernel<<<n, 128, smem>>>(dx.p, du.p, dfe.p, grid_mem(), dqdd_.p, ddq.p, dee_.p, ddee.p);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        if (dqdd_.down(qdd) || ddq.down(dqdd) || dee_.down(ee) || ddee.down(dee)) return -1;
        return 0;
}

}  // extern "C"
