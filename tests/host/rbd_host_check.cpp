// TEST INFRASTRUCTURE: compiles the PRODUCT's per-item math (gato_b200/csrc/rbd.cuh, items.cuh) for the HOST so
// that it can be compared bit-for-bit with the CPU oracle on a machine without a GPU.  sin/cos/log are
// routed to the oracle's libdevice-equivalent implementations (on the GPU the product calls CUDA's own).
#define GATO_HOST_TEST 1
#include <cstring>
#include "../../gato_b200/csrc/items.cuh"
#include "../../oracle/bsqp_oracle.h"

extern "C" float gato_host_sinf(float x) { return gato_oracle_sinf(x); }
extern "C" float gato_host_cosf(float x) { return gato_oracle_cosf(x); }
extern "C" float gato_host_logf(float x) { return gato_oracle_logf(x); }

using namespace gato;

template<class P>
static void dyn_dump(int n, const float* x, const float* u, const float* fext, float* qdd, float* dqdd, float* ee, float* dee)
{
        constexpr int NQ = P::NQ;
        using R = Rbd<P>;
        for (int s = 0; s < n; s++) {
                float q_[NQ], d_[3 * NQ * NQ], e3[3], J[NQ][3];
                R::fd_and_grad(x + s * 2 * NQ, x + s * 2 * NQ + NQ, u + s * NQ, fext + 6 * s, q_, d_);
                memcpy(qdd + s * NQ, q_, sizeof(q_));
                memcpy(dqdd + s * 3 * NQ * NQ, d_, sizeof(d_));
                R::ee_pos_grad(x + s * 2 * NQ, e3, J);
                for (int r = 0; r < 6; r++) ee[6 * s + r] = r < 3 ? e3[r] : 0.0f;
                for (int dj = 0; dj < NQ; dj++)
                        for (int r = 0; r < 6; r++) dee[s * 6 * NQ + 6 * dj + r] = r < 3 ? J[dj][r] : 0.0f;
        }
}

template<class P>
static void stage_kkt(int N, int B, const float* xu, const float* xs, const float* ref, const float* fext, float dt, const float* c7, float* Q, float* R_, float* q, float* r, float* A, float* Bm,
                      float* c)
{
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ;
        const int     traj = (NX + NU) * N - NU;
        const Costs   cs{c7[0], c7[1], c7[2], c7[3], c7[4], c7[5], c7[6]};
        for (int b = 0; b < B; b++)
                for (int k = 0; k < N - 1; k++) {
                        const float* xux = xu + (size_t)b * traj + k * (NX + NU);
                        float*       Ak = A + ((size_t)b * N + k) * NX * NX;
                        float*       Bk = Bm + ((size_t)b * N + k) * NX * NU;
                        float*       ck = c + ((size_t)b * N + k + 1) * NX;
                        Items<P>::linearize(xux, fext + 6 * b, dt, [&](int e, float v) { Ak[e] = v; }, [&](int e, float v) { Bk[e] = v; }, [&](int e, float v) { ck[e] = v; });
                        float* Qk = Q + ((size_t)b * N + k) * NX * NX;
                        float* qk = q + ((size_t)b * N + k) * NX;
                        float* Rk = R_ + ((size_t)b * N + k) * NU * NU;
                        float* rk = r + ((size_t)b * N + k) * NU;
                        Items<P>::template cost_grad_hess<true>(xux, ref + (size_t)b * 6 * N + 6 * k, cs, [&](int e, float v) { Qk[e] = v; }, [&](int e, float v) { qk[e] = v; },
                                                                [&](int e, float v) { Rk[e] = v; }, [&](int e, float v) { rk[e] = v; }, false, Items<P>::pos_form_b_for(N));
                        if (k == N - 2) {
                                float* Qn = Qk + NX * NX;
                                float* qn = qk + NX;
                                Items<P>::template cost_grad_hess<false>(xux, ref + (size_t)b * 6 * N + 6 * (k + 1), cs, [&](int e, float v) { Qn[e] = v; }, [&](int e, float v) { qn[e] = v; },
                                                                         [&](int, float) {}, [&](int, float) {});
                                for (int i = 0; i < NX; i++) c[(size_t)b * N * NX + i] = xu[(size_t)b * traj + i] - xs[(size_t)b * NX + i];
                        }
                }
}

// the same stage through the code paths the KERNELS run: k_kkt's rolled halves (linearize_half_rolled<0/1>) when variant == 1, k_kkt_fine's
// per-column pieces (dyn_prologue + linearize_base + linearize_column) when variant == 2
template<class P>
static void stage_kkt_kernel_paths(int variant, int N, int B, const float* xu, const float* fext, float dt, float* A, float* Bm, float* c)
{
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ;
        const int     traj = (NX + NU) * N - NU;
        for (int b = 0; b < B; b++)
                for (int k = 0; k < N - 1; k++) {
                        const float* xux = xu + (size_t)b * traj + k * (NX + NU);
                        float*       Ak = A + ((size_t)b * N + k) * NX * NX;
                        float*       Bk = Bm + ((size_t)b * N + k) * NX * NU;
                        float*       ck = c + ((size_t)b * N + k + 1) * NX;
                        if (variant == 1) {
                                Items<P>::template linearize_half_rolled<0>(xux, fext + 6 * b, dt, [&](int e, float v) { Ak[e] = v; }, [&](int, float) {}, [&](int e, float v) { ck[e] = v; });
                                Items<P>::template linearize_half_rolled<1>(xux, fext + 6 * b, dt, [&](int e, float v) { Ak[e] = v; }, [&](int e, float v) { Bk[e] = v; }, [&](int, float) {});
                        } else {
                                typename Rbd<P>::DynState st;
                                Rbd<P>::dyn_prologue(xux, xux + NQ, xux + NX, fext + 6 * b, st);
                                Items<P>::linearize_base(st, xux, dt, [&](int e, float v) { Bk[e] = v; }, [&](int e, float v) { ck[e] = v; });
                                sfor<0, NX>([&](auto cc) {
                                        constexpr int cidx = cc;
                                        Items<P>::template linearize_column<cidx / NQ, cidx % NQ>(st, xux + NQ, dt, [&](int e, float v) { Ak[e] = v; });
                                });
                        }
                }
}

// the split merit kernel's halves: merit = fmaf(mu, merit_mid_cons, tracking_cost)
template<class P>
static void stage_merit_split(int N, int B, const float* xu, const float* dz, const float* xs, const float* ref, const float* mu, const float* fext, float dt, const float* c7, int na, float* merit)
{
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ;
        const int     traj = (NX + NU) * N - NU;
        const Costs   cs{c7[0], c7[1], c7[2], c7[3], c7[4], c7[5], c7[6]};
        for (int b = 0; b < B; b++)
                for (int a = 0; a < na; a++) {
                        const float alpha = (float)(1.0 / (double)(1 << a));
                        float       m = 0.0f;
                        for (int k = 0; k < N; k++) {
                                float        xux[2 * NX + NU];
                                const int    cnt = (k == N - 1) ? NX : 2 * NX + NU;
                                const float *xk = xu + (size_t)b * traj + k * (NX + NU), *dk = dz + (size_t)b * traj + k * (NX + NU);
                                for (int i = 0; i < cnt; i++) xux[i] = fmaf(alpha, dk[i], xk[i]);
                                float mk;
                                if (k < N - 1)
                                        mk = fmaf(mu[b], Items<P>::merit_mid_cons(xux, fext + 6 * b, dt), Items<P>::template tracking_cost<false>(xux, ref + (size_t)b * 6 * N + 6 * k, cs));
                                else {
                                        float e0[NX];
                                        for (int i = 0; i < NX; i++) e0[i] = fabsf(fmaf(alpha, dz[(size_t)b * traj + i], xu[(size_t)b * traj + i]) - xs[(size_t)b * NX + i]);
                                        mk = fmaf(mu[b], Items<P>::merit_last_cons(e0), Items<P>::template tracking_cost<true>(xux, ref + (size_t)b * 6 * N + 6 * k, cs));
                                }
                                m = m + mk;
                        }
                        merit[b * na + a] = m;
                }
}

template<class P>
static void stage_merit(int N, int B, const float* xu, const float* dz, const float* xs, const float* ref, const float* mu, const float* fext, float dt, const float* c7, int na, float* merit)
{
        constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ;
        const int     traj = (NX + NU) * N - NU;
        const Costs   cs{c7[0], c7[1], c7[2], c7[3], c7[4], c7[5], c7[6]};
        for (int b = 0; b < B; b++)
                for (int a = 0; a < na; a++) {
                        const float alpha = (float)(1.0 / (double)(1 << a));
                        float       m = 0.0f;
                        for (int k = 0; k < N; k++) {
                                float        xux[2 * NX + NU];
                                const int    cnt = (k == N - 1) ? NX : 2 * NX + NU;
                                const float *xk = xu + (size_t)b * traj + k * (NX + NU), *dk = dz + (size_t)b * traj + k * (NX + NU);
                                for (int i = 0; i < cnt; i++) xux[i] = fmaf(alpha, dk[i], xk[i]);
                                float mk;
                                if (k < N - 1)
                                        mk = Items<P>::merit_mid(xux, ref + (size_t)b * 6 * N + 6 * k, mu[b], fext + 6 * b, dt, cs);
                                else {
                                        float e0[NX];
                                        for (int i = 0; i < NX; i++) e0[i] = fabsf(fmaf(alpha, dz[(size_t)b * traj + i], xu[(size_t)b * traj + i]) - xs[(size_t)b * NX + i]);
                                        mk = Items<P>::merit_last(xux, ref + (size_t)b * 6 * N + 6 * k, mu[b], e0, cs);
                                }
                                m = m + mk;
                        }
                        merit[b * na + a] = m;
                }
}

extern "C" {
int hostchk_dyn_dump(int plant, int n, const float* x, const float* u, const float* fext, float* qdd, float* dqdd, float* ee, float* dee)
{
        if (plant == 1)
                dyn_dump<Iiwa14>(n, x, u, fext, qdd, dqdd, ee, dee);
        else
                dyn_dump<Indy7>(n, x, u, fext, qdd, dqdd, ee, dee);
        return 0;
}
int hostchk_stage_kkt(int plant, int N, int B, const float* xu, const float* xs, const float* ref, const float* fext, float dt, const float* cost7, float* Q, float* R, float* q, float* r, float* A,
                      float* Bm, float* c)
{
        if (plant == 1)
                stage_kkt<Iiwa14>(N, B, xu, xs, ref, fext, dt, cost7, Q, R, q, r, A, Bm, c);
        else
                stage_kkt<Indy7>(N, B, xu, xs, ref, fext, dt, cost7, Q, R, q, r, A, Bm, c);
        return 0;
}
int hostchk_stage_kkt_kernel_paths(int plant, int variant, int N, int B, const float* xu, const float* fext, float dt, float* A, float* Bm, float* c)
{
        if (plant == 1)
                stage_kkt_kernel_paths<Iiwa14>(variant, N, B, xu, fext, dt, A, Bm, c);
        else
                stage_kkt_kernel_paths<Indy7>(variant, N, B, xu, fext, dt, A, Bm, c);
        return 0;
}
int hostchk_stage_merit_split(int plant, int N, int B, const float* xu, const float* dz, const float* xs, const float* ref, const float* mu, const float* fext, float dt, const float* cost7, int na,
                              float* merit)
{
        if (plant == 1)
                stage_merit_split<Iiwa14>(N, B, xu, dz, xs, ref, mu, fext, dt, cost7, na, merit);
        else
                stage_merit_split<Indy7>(N, B, xu, dz, xs, ref, mu, fext, dt, cost7, na, merit);
        return 0;
}
int hostchk_stage_merit(int plant, int N, int B, const float* xu, const float* dz, const float* xs, const float* ref, const float* mu, const float* fext, float dt, const float* cost7, int na,
                        float* merit)
{
        if (plant == 1)
                stage_merit<Iiwa14>(N, B, xu, dz, xs, ref, mu, fext, dt, cost7, na, merit);
        else
                stage_merit<Indy7>(N, B, xu, dz, xs, ref, mu, fext, dt, cost7, na, merit);
        return 0;
}
}
