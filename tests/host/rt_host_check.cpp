// TEST INFRASTRUCTURE: compiles the PRODUCT's table-driven per-item math (gato_b200/csrc/rbd_rt.cuh, items_rt.cuh -- what the kernels of the
// run-time robot models execute) for the HOST, so that it can be compared bit-for-bit with the CPU oracle on a machine without a GPU.
#define GATO_HOST_TEST 1
#include <cstdio>
#include <cstring>
#include "../../gato_b200/csrc/items_rt.cuh"
#include "../../oracle/bsqp_oracle.h"

extern "C" float gato_host_sinf(float x) { return gato_oracle_sinf(x); }
extern "C" float gato_host_cosf(float x) { return gato_oracle_cosf(x); }
extern "C" float gato_host_logf(float x) { return gato_oracle_logf(x); }

using namespace gato;

static RtModel g_model;

template<int NQ>
static void stage_kkt(int variant, int N, int B, const float* xu, const float* xs, const float* ref, const float* fext, float dt, const float* c7, float* Q, float* R_, float* q, float* r,
                      float* A, float* Bm, float* c)
{
        using It = Items<RtPlant<NQ>>;
        constexpr int NX = 2 * NQ, NU = NQ;
        const int     traj = (NX + NU) * N - NU;
        const Costs   cs{c7[0], c7[1], c7[2], c7[3], c7[4], c7[5], c7[6]};
        const It      it(g_model);
        for (int b = 0; b < B; b++)
                for (int k = 0; k < N - 1; k++) {
                        const float* xux = xu + (size_t)b * traj + k * (NX + NU);
                        float*       Ak = A + ((size_t)b * N + k) * NX * NX;
                        float*       Bk = Bm + ((size_t)b * N + k) * NX * NU;
                        float*       ck = c + ((size_t)b * N + k + 1) * NX;
                        if (variant == 1) {  // k_kkt: the two halves
                                it.template linearize_half_rolled<0>(xux, fext + 6 * b, dt, [&](int e, float v) { Ak[e] = v; }, [&](int, float) {}, [&](int e, float v) { ck[e] = v; });
                                it.template linearize_half_rolled<1>(xux, fext + 6 * b, dt, [&](int e, float v) { Ak[e] = v; }, [&](int e, float v) { Bk[e] = v; }, [&](int, float) {});
                        } else {  // k_kkt_fine: prologue + base + one column at a time
                                typename It::DynState st;
                                it.prologue(xux, fext + 6 * b, st);
                                it.linearize_base(st, xux, dt, [&](int e, float v) { Bk[e] = v; }, [&](int e, float v) { ck[e] = v; });
                                for (int col = 0; col < NX; col++) it.linearize_column_any(col, st, xux + NQ, dt, [&](int e, float v) { Ak[e] = v; });
                        }
                        float* Qk = Q + ((size_t)b * N + k) * NX * NX;
                        float* qk = q + ((size_t)b * N + k) * NX;
                        float* Rk = R_ + ((size_t)b * N + k) * NU * NU;
                        float* rk = r + ((size_t)b * N + k) * NU;
                        it.template cost_grad_hess<true>(xux, ref + (size_t)b * 6 * N + 6 * k, cs, [&](int e, float v) { Qk[e] = v; }, [&](int e, float v) { qk[e] = v; },
                                                         [&](int e, float v) { Rk[e] = v; }, [&](int e, float v) { rk[e] = v; }, false, It::pos_form_b_for(N));
                        if (k == N - 2) {
                                float* Qn = Qk + NX * NX;
                                float* qn = qk + NX;
                                it.template cost_grad_hess<true>(xux, ref + (size_t)b * 6 * N + 6 * (k + 1), cs, [&](int e, float v) { Qn[e] = v; }, [&](int e, float v) { qn[e] = v; },
                                                                 [&](int, float) {}, [&](int, float) {}, true, true);  // the terminal item exactly as KktWarp::cost_item calls it
                                for (int i = 0; i < NX; i++) c[(size_t)b * N * NX + i] = xu[(size_t)b * traj + i] - xs[(size_t)b * NX + i];
                        }
                }
}

template<int NQ>
static void stage_merit(int split, int N, int B, const float* xu, const float* dz, const float* xs, const float* ref, const float* mu, const float* fext, float dt, const float* c7, int na,
                        float* merit)
{
        using It = Items<RtPlant<NQ>>;
        constexpr int NX = 2 * NQ, NU = NQ;
        const int     traj = (NX + NU) * N - NU;
        const Costs   cs{c7[0], c7[1], c7[2], c7[3], c7[4], c7[5], c7[6]};
        const It      it(g_model);
        for (int b = 0; b < B; b++)
                for (int a = 0; a < na; a++) {
                        const float alpha = (float)(1.0 / (double)(1 << a));
                        float       m = 0.0f;
                        for (int k = 0; k < N; k++) {
                                float        xux[2 * NX + NU];
                                const int    cnt = (k == N - 1) ? NX : 2 * NX + NU;
                                const float *xk = xu + (size_t)b * traj + k * (NX + NU), *dk = dz + (size_t)b * traj + k * (NX + NU);
                                const float* rf = ref + (size_t)b * 6 * N + 6 * k;
                                for (int i = 0; i < cnt; i++) xux[i] = fmaf(alpha, dk[i], xk[i]);
                                float mk;
                                if (k < N - 1) {
                                        mk = split ? fmaf(mu[b], it.merit_mid_cons(xux, fext + 6 * b, dt), it.template tracking_cost<false>(xux, rf, cs)) : it.merit_mid(xux, rf, mu[b], fext + 6 * b, dt, cs);
                                } else {
                                        float e0[NX];
                                        for (int i = 0; i < NX; i++) e0[i] = fabsf(fmaf(alpha, dz[(size_t)b * traj + i], xu[(size_t)b * traj + i]) - xs[(size_t)b * NX + i]);
                                        mk = split ? fmaf(mu[b], it.merit_last_cons(e0), it.template tracking_cost<true>(xux, rf, cs)) : it.merit_last(xux, rf, mu[b], e0, cs);
                                }
                                m = m + mk;
                        }
                        merit[b * na + a] = m;
                }
}

// forward dynamics + end-effector position per sample: scalar lanes, and the SAME samples two at a time as the lanes of packed pairs
template<int NQ>
static void dyn(int packed, int n, const float* x, const float* u, const float* fext, float* qdd, float* ee)
{
        const Items<RtPlant<NQ>> it(g_model);
        if (!packed) {
                for (int s = 0; s < n; s++) {
                        float q_[NQ], e3[3];
                        RbdRt<NQ, float>::forward_dynamics(g_model, x + s * 2 * NQ, x + s * 2 * NQ + NQ, u + s * NQ, fext + 6 * s, q_);
                        it.ee_pos(x + s * 2 * NQ, e3);
                        memcpy(qdd + s * NQ, q_, sizeof(q_));
                        memcpy(ee + 3 * s, e3, sizeof(e3));
                }
                return;
        }
        for (int s = 0; s < n; s += 2) {
                const int s1 = s + 1 < n ? s + 1 : s;
                f2        q2[NQ], qd2[NQ], u2[NQ], fe2[6], out[NQ], e3[3];
                for (int i = 0; i < NQ; i++) {
                        q2[i] = mk2(x[s * 2 * NQ + i], x[s1 * 2 * NQ + i]);
                        qd2[i] = mk2(x[s * 2 * NQ + NQ + i], x[s1 * 2 * NQ + NQ + i]);
                        u2[i] = mk2(u[s * NQ + i], u[s1 * NQ + i]);
                }
                for (int i = 0; i < 6; i++) fe2[i] = mk2(fext[6 * s + i], fext[6 * s1 + i]);
                RbdRt<NQ, f2>::forward_dynamics(g_model, q2, qd2, u2, fe2, out);
                RbdRt<NQ, f2>::ee_pos(g_model, q2, e3);
                for (int i = 0; i < NQ; i++) qdd[s * NQ + i] = out[i].x, qdd[s1 * NQ + i] = out[i].y;
                for (int i = 0; i < 3; i++) ee[3 * s + i] = e3[i].x, ee[3 * s1 + i] = e3[i].y;
        }
}

extern "C" {
// returns 0, or -1 with the reason in why[256]
int hostchk_rt_set_model(const gato_model* m, char* why)
{
        const char* e = rt_model_from_desc(*m, g_model);
        if (e && why) snprintf(why, 256, "%s", e);
        return e ? -1 : 0;
}
int hostchk_rt_stage_kkt(int variant, int N, int B, const float* xu, const float* xs, const float* ref, const float* fext, float dt, const float* cost7, float* Q, float* R, float* q, float* r,
                         float* A, float* Bm, float* c)
{
        if (g_model.nq == 8)
                stage_kkt<8>(variant, N, B, xu, xs, ref, fext, dt, cost7, Q, R, q, r, A, Bm, c);
        else if (g_model.nq == 7)
                stage_kkt<7>(variant, N, B, xu, xs, ref, fext, dt, cost7, Q, R, q, r, A, Bm, c);
        else
                stage_kkt<6>(variant, N, B, xu, xs, ref, fext, dt, cost7, Q, R, q, r, A, Bm, c);
        return 0;
}
int hostchk_rt_stage_merit(int split, int N, int B, const float* xu, const float* dz, const float* xs, const float* ref, const float* mu, const float* fext, float dt, const float* cost7, int na,
                           float* merit)
{
        if (g_model.nq == 8)
                stage_merit<8>(split, N, B, xu, dz, xs, ref, mu, fext, dt, cost7, na, merit);
        else if (g_model.nq == 7)
                stage_merit<7>(split, N, B, xu, dz, xs, ref, mu, fext, dt, cost7, na, merit);
        else
                stage_merit<6>(split, N, B, xu, dz, xs, ref, mu, fext, dt, cost7, na, merit);
        return 0;
}
int hostchk_rt_dyn(int packed, int n, const float* x, const float* u, const float* fext, float* qdd, float* ee)
{
        if (g_model.nq == 8)
                dyn<8>(packed, n, x, u, fext, qdd, ee);
        else if (g_model.nq == 7)
                dyn<7>(packed, n, x, u, fext, qdd, ee);
        else
                dyn<6>(packed, n, x, u, fext, qdd, ee);
        return 0;
}
}
