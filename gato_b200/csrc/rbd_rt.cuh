// Table-driven rigid-body dynamics: the same building blocks as rbd.cuh (RNEA, direct M^-1, forward dynamics, analytical RNEA gradient,
// end-effector position + Jacobian, integrator, barriers) for a robot whose constants arrive at RUN TIME (rt_model.h), with the loops over
// joints ROLLED -- one copy of the per-joint code, the joint index a run-time value, per-joint state addressed dynamically (local memory),
// constants read from the model table (constant memory on the device).
//
// Replaces the same GRiD-generated device functions as rbd.cuh (iiwa14_grid.cuh:2212, 4742, 5341, 5549, 2596, 2855; iiwa14_fext.cuh:29, 271)
// for ANY fixed-base chain of z-axis revolute joints, not only the two robots the reference ships generated code for.
//
// Arithmetic contract: fp32, explicit fmaf, every dot product in the reference's term order (ascending index from +0).  The only terms left
// out are those of the structurally zero top-right block of a Pluecker transform; a chain that starts from +0 is never -0, so dropping a
// +-0 product changes nothing -- the results are bit-identical to the CPU oracle's dense loops and, for the two compiled robots, to
// rbd.cuh (tests/test_host_math.py, tests/test_gpu_models.py).
//
// Local arrays: every dynamically indexed array of a function lives in ONE work struct (FdWork, DynState with its MinvWork / ColWork union,
// EeWork) -- a single stack object -- and callees get references into it instead of declaring arrays of their own.  This is deliberate: nvcc
// 12.9's optimiser merges the stack slots of separately declared local arrays by lifetime, and with minv's scratch (F, U, Dinv), its output
// Minv and update_X's short-lived sin / cos table as individual locals it gave two SIMULTANEOUSLY LIVE arrays -- F and Minv -- the same address
// (seen in the SASS of k_sim_forward: both addressed off one base register; wrong M^-1 at every optimisation level except -G, while
// k_merit_ls, which inlines the same functions, was correct).  Members of one struct cannot be merged; tests/test_gpu_models.py checks every
// table-driven kernel against the oracle bit for bit.
//
// T is the lane type: float (one work item per thread) or f2 (TWO work items per thread as the lanes of Blackwell's packed FFMA2 / FADD2:
// all items run the same instruction stream, so the packed form halves the instructions per item; each lane is rounded like the scalar
// operation).
#pragma once
#include "rbd.cuh"
#include "rt_model.h"

namespace gato {

// plant tags of the run-time models (P::ID >= 2 selects the table-driven code in Items / the kernels)
template<int NQ_>
struct RtPlant {
        static constexpr int ID = 2, NQ = NQ_;
};
template<class P>
constexpr bool is_rt_plant = (P::ID >= 2);

template<class T>
struct Lane;
template<>
struct Lane<float> {
        static constexpr int W = 1;
        static GATO_HD float set(float s) { return s; }
        static GATO_HD float fma(float a, float b, float c) { return fmaf(a, b, c); }
        static GATO_HD float fmas(float s, float b, float c) { return fmaf(s, b, c); }
        static GATO_HD float add(float a, float b) { return a + b; }
        static GATO_HD float sub(float a, float b) { return a - b; }
        static GATO_HD float mul(float a, float b) { return a * b; }
        static GATO_HD float muls(float s, float b) { return s * b; }
        static GATO_HD float neg(float a) { return -a; }
        template<class F>
        static GATO_HD float map(float a, F&& f)
        {
                return f(a);
        }
        template<class F>
        static GATO_HD float map2(float a, float b, F&& f)
        {
                return f(a, b);
        }
        static GATO_HD float get(float a, int) { return a; }
};
template<>
struct Lane<f2> {
        static constexpr int W = 2;
        static GATO_HD f2 set(float s) { return mk2(s, s); }
        static GATO_HD f2 fma(f2 a, f2 b, f2 c) { return fma2(a, b, c); }
        static GATO_HD f2 fmas(float s, f2 b, f2 c) { return fma2s(s, b, c); }
        static GATO_HD f2 add(f2 a, f2 b) { return add2(a, b); }
        static GATO_HD f2 sub(f2 a, f2 b) { return add2(a, neg2(b)); }  // a - b == a + (-b) bit for bit
        static GATO_HD f2 mul(f2 a, f2 b) { return mul2(a, b); }
        static GATO_HD f2 muls(float s, f2 b) { return mul2s(s, b); }
        static GATO_HD f2 neg(f2 a) { return neg2(a); }
        template<class F>
        static GATO_HD f2 map(f2 a, F&& f)
        {
                return mk2(f(a.x), f(a.y));
        }
        template<class F>
        static GATO_HD f2 map2(f2 a, f2 b, F&& f)
        {
                return mk2(f(a.x, b.x), f(a.y, b.y));
        }
        static GATO_HD float get(f2 a, int l) { return l ? a.y : a.x; }
};

#if defined(__CUDA_ARCH__)
#define GATO_ROLLED _Pragma("unroll 1")
#else
#define GATO_ROLLED
#endif

template<int NQ, class T>
struct RbdRt {
        using L = Lane<T>;
        static constexpr int NX = 2 * NQ, NU = NQ;
        using Xmat = T[NQ][18];  // E | B per joint (rt_model.h: x_compact)
        using V6 = T[NQ][6];

        // ---- X_j(q_j)  (load_update_XImats_helpers) ------------------------------------------------------------
        // t[k] = sin(q_k), t[NQ + k] = cos(q_k)
        static GATO_HD void sincos(const T* q, T (&t)[2 * NQ])
        {
                GATO_ROLLED
                for (int k = 0; k < NQ; k++) {
                        t[k] = L::map(q[k], [](float a) { return g_sin(a); });
                        t[NQ + k] = L::map(q[k], [](float a) { return g_cos(a); });
                }
        }
        static GATO_HD T trig_entry(const RtTrig& e, const T (&t)[2 * NQ])
        {
                const double coef = e.coef;
                return L::map(t[e.k + (e.use_cos ? NQ : 0)], [&](float tv) { return (float)(coef * (double)tv); });
        }
        static GATO_HD void update_X(const RtModel& m, const T* q, T (&t)[2 * NQ], Xmat& X)
        {
                sincos(q, t);
                GATO_ROLLED
                for (int j = 0; j < NQ; j++) {
                        sfor<0, 18>([&](auto ec) { X[j][ec] = L::set(m.X[j][ec]); });
                        const int n = m.nxt[j];
                        GATO_ROLLED
                        for (int i = 0; i < n; i++) X[j][m.xt[j][i].loc] = trig_entry(m.xt[j][i], t);
                }
        }
        // entry (R, C) of X_j; (R < 3, C >= 3) is the zero block and must not be asked for
        template<int R, int C>
        static GATO_HD T xe(const T (&Xj)[18])
        {
                static_assert(!(R < 3 && C >= 3), "structurally zero block");
                if constexpr (C >= 3)
                        return Xj[x_compact(R - 3, C - 3)];  // bottom-right = top-left
                else
                        return Xj[x_compact(R, C)];
        }
        // (X_j v)[R]
        template<int R>
        static GATO_HD T xrow(const T (&Xj)[18], const T (&v)[6])
        {
                T r = L::set(0.0f);
                sfor<0, (R < 3 ? 3 : 6)>([&](auto ic) { r = L::fma(xe<R, ic>(Xj), v[ic], r); });
                return r;
        }
        // (X_j^T f)[C]
        template<int C>
        static GATO_HD T xcol(const T (&Xj)[18], const T (&f)[6])
        {
                T r = L::set(0.0f);
                sfor<(C < 3 ? 0 : 3), 6>([&](auto ic) { r = L::fma(xe<ic, C>(Xj), f[ic], r); });
                return r;
        }
        // (I_j v)[R], dense
        template<int R>
        static GATO_HD T irow(const float (&Ij)[36], const T (&v)[6])
        {
                T r = L::set(0.0f);
                sfor<0, 6>([&](auto ic) { r = L::fmas(Ij[6 * ic + R], v[ic], r); });
                return r;
        }
        // fx(f) * t  (iiwa14_grid.cuh:896-905) with the fma placement nvcc emits for it
        static GATO_HD void fx_times_v(T (&r)[6], const T (&f)[6], const T (&t)[6])
        {
                T s;
                s = L::fma(f[1], t[2], L::neg(L::mul(f[2], t[1])));
                s = L::fma(L::neg(f[5]), t[4], s);
                r[0] = L::fma(f[4], t[5], s);
                s = L::fma(f[2], t[0], L::neg(L::mul(f[0], t[2])));
                s = L::fma(f[5], t[3], s);
                r[1] = L::fma(L::neg(f[3]), t[5], s);
                s = L::fma(f[0], t[1], L::neg(L::mul(f[1], t[0])));
                s = L::fma(L::neg(f[4]), t[3], s);
                r[2] = L::fma(f[3], t[4], s);
                r[3] = L::fma(f[1], t[5], L::neg(L::mul(f[2], t[4])));
                r[4] = L::fma(f[2], t[3], L::neg(L::mul(f[0], t[5])));
                r[5] = L::fma(f[0], t[4], L::neg(L::mul(f[1], t[3])));
        }
        // a_0 = X_0[:, 5] * g: rows 0..2 are the zero block
        template<int R>
        static GATO_HD T gravity_row(const T (&X0)[18])
        {
                if constexpr (R < 3)
                        return L::set(0.0f);
                else
                        return L::muls(kGravity, xe<R, 5>(X0));
        }

        // ---- RNEA  (inverse_dynamics_inner / _vaf with external wrench; oracle: rnea) -----------------------------
        // One forward sweep carries v_{j-1}, a_{j-1} in registers and finishes f_j right away (each value is computed by the same operations as in
        // the reference's three sweeps); v and a are only stored when the caller needs them (KEEP_VA: the gradient does, forward dynamics does not).
        template<bool WITH_QDD, bool KEEP_VA>
        static GATO_HD void rnea(const RtModel& m, const Xmat& X, const T* qd, const T* qdd, const T* fext, V6* v, V6* a, V6& f)
        {
                T vj[6], aj[6];
                sfor<0, 6>([&](auto rc) {
                        constexpr int row = rc;
                        vj[row] = L::set(0.0f);
                        aj[row] = gravity_row<row>(X[0]);
                });
                vj[2] = L::add(vj[2], qd[0]);
                if constexpr (WITH_QDD) aj[2] = L::add(aj[2], qdd[0]);
                GATO_ROLLED
                for (int j = 0; j < NQ; j++) {
                        if (j > 0) {
                                T nv[6], na[6];
                                sfor<0, 6>([&](auto rc) {
                                        constexpr int row = rc;
                                        T             vv = xrow<row>(X[j], vj);
                                        T             aa = xrow<row>(X[j], aj);
                                        if constexpr (row == 2) {
                                                vv = L::add(vv, qd[j]);
                                                if constexpr (WITH_QDD) aa = L::add(aa, qdd[j]);
                                        }
                                        nv[row] = vv, na[row] = aa;
                                });
                                na[0] = L::fma(nv[1], qd[j], na[0]);
                                na[1] = L::fma(L::neg(nv[0]), qd[j], na[1]);
                                na[3] = L::fma(nv[4], qd[j], na[3]);
                                na[4] = L::fma(L::neg(nv[3]), qd[j], na[4]);
                                sfor<0, 6>([&](auto rc) { vj[rc] = nv[rc], aj[rc] = na[rc]; });
                        }
                        if constexpr (KEEP_VA) sfor<0, 6>([&](auto rc) { (*v)[j][rc] = vj[rc], (*a)[j][rc] = aj[rc]; });
                        T fj[6], Iv[6], t[6];
                        sfor<0, 6>([&](auto rc) {
                                constexpr int row = rc;
                                fj[row] = irow<row>(m.I[j], aj);
                                Iv[row] = irow<row>(m.I[j], vj);
                        });
                        fx_times_v(t, vj, Iv);
                        sfor<0, 6>([&](auto rc) {
                                fj[rc] = L::add(fj[rc], t[rc]);
                                if (j == NQ - 1) fj[rc] = L::sub(fj[rc], fext[rc]);
                                f[j][rc] = fj[rc];
                        });
                }
                GATO_ROLLED
                for (int j = NQ - 1; j >= 1; j--) {
                        T val[6], fj[6];
                        sfor<0, 6>([&](auto rc) { fj[rc] = f[j][rc]; });
                        sfor<0, 6>([&](auto rc) { val[rc] = xcol<rc>(X[j], fj); });
                        sfor<0, 6>([&](auto rc) { f[j - 1][rc] = L::add(f[j - 1][rc], val[rc]); });
                }
        }

        // ---- direct M^-1  (direct_minv_inner; oracle: minv); Minv[col*NQ+row], upper triangle valid ------------------
        // scratch of minv: F[j] = column j of the current level's F
        struct MinvWork {
                T F[NQ][6], U[NQ][6], Dinv[NQ];
        };
        static GATO_HD void minv(const RtModel& m, const Xmat& X, MinvWork& w, T (&Minv)[NQ * NQ])
        {
                T IA[36];
                T(&F)[NQ][6] = w.F;
                T(&U)[NQ][6] = w.U;
                T(&Dinv)[NQ] = w.Dinv;
                sfor<0, NQ * NQ>([&](auto ic) { Minv[ic] = L::set(0.0f); });
                GATO_ROLLED
                for (int j = 0; j < NQ; j++) sfor<0, 6>([&](auto rc) { F[j][rc] = L::set(0.0f); });
                sfor<0, 36>([&](auto ec) { IA[ec] = L::set(m.I[NQ - 1][ec]); });
                GATO_ROLLED
                for (int i = NQ - 1; i >= 0; i--) {
                        T Ui[6];
                        sfor<0, 6>([&](auto rc) { Ui[rc] = IA[12 + rc], U[i][rc] = Ui[rc]; });
                        const T Di = L::map(Ui[2], [](float x) { return 1.0f / x; });
                        Dinv[i] = Di;
                        Minv[i * NQ + i] = Di;
                        GATO_ROLLED
                        for (int j = i; j < NQ; j++) {
                                const T mji = L::fma(L::neg(Di), F[j][2], Minv[j * NQ + i]);
                                Minv[j * NQ + i] = mji;
                                if (i > 0) sfor<0, 6>([&](auto rc) { F[j][rc] = L::fma(Ui[rc], mji, F[j][rc]); });
                        }
                        if (i > 0) {
                                T Ia[36], IaT[36];
                                sfor<0, 36>([&](auto ec) {
                                        constexpr int row = ec % 6, col = ec / 6;
                                        Ia[ec] = L::fma(L::neg(L::mul(Ui[row], Di)), Ui[col], IA[ec]);
                                });
                                T Xi[18];
                                sfor<0, 18>([&](auto ec) { Xi[ec] = X[i][ec]; });
                                // F[i-1][:, j] = X_i^T F[i][:, j]  (columns j >= i; the parent's own column i-1 starts at zero)
                                GATO_ROLLED
                                for (int j = i; j < NQ; j++) {
                                        T tmp[6], Fj[6];
                                        sfor<0, 6>([&](auto rc) { Fj[rc] = F[j][rc]; });
                                        sfor<0, 6>([&](auto rc) { tmp[rc] = xcol<rc>(Xi, Fj); });
                                        sfor<0, 6>([&](auto rc) { F[j][rc] = tmp[rc]; });
                                }
                                sfor<0, 6>([&](auto rc) { F[i - 1][rc] = L::set(0.0f); });
                                sfor<0, 6>([&](auto cc) {
                                        constexpr int c = cc;
                                        T             col[6];
                                        sfor<0, 6>([&](auto tc) { col[tc] = Ia[6 * c + tc]; });
                                        sfor<0, 6>([&](auto rc) { IaT[6 * c + rc] = xcol<rc>(Xi, col); });
                                });
                                // IA[i-1] = I[i-1] + IaT * X_i
                                sfor<0, 36>([&](auto ec) {
                                        constexpr int row = ec % 6, col = ec / 6;
                                        T             val = L::set(0.0f);
                                        sfor<(col < 3 ? 0 : 3), 6>([&](auto tc) { val = L::fma(IaT[row + 6 * tc], xe<tc, col>(Xi), val); });
                                        IA[ec] = L::add(L::set(m.I[i - 1][ec]), val);
                                });
                        }
                }
                // forward pass: F[j] now holds column j of the current level's F
                GATO_ROLLED
                for (int j = 0; j < NQ; j++) sfor<0, 6>([&](auto rc) { F[j][rc] = L::muls((rc == 2 ? 1.0f : 0.0f), Minv[j * NQ]); });
                GATO_ROLLED
                for (int i = 1; i < NQ; i++) {
                        T Xi[18], Ui[6];
                        sfor<0, 18>([&](auto ec) { Xi[ec] = X[i][ec]; });
                        sfor<0, 6>([&](auto rc) { Ui[rc] = U[i][rc]; });
                        const T nDi = L::neg(Dinv[i]);
                        GATO_ROLLED
                        for (int j = i; j < NQ; j++) {
                                T tmp[6], Fj[6];
                                sfor<0, 6>([&](auto rc) { Fj[rc] = F[j][rc]; });
                                sfor<0, 6>([&](auto rc) { tmp[rc] = xrow<rc>(Xi, Fj); });
                                T d = L::set(0.0f);
                                sfor<0, 6>([&](auto tc) { d = L::fma(tmp[tc], Ui[tc], d); });
                                const T mji = L::fma(nDi, d, Minv[j * NQ + i]);
                                Minv[j * NQ + i] = mji;
                                if (i < NQ - 1) tmp[2] = L::add(tmp[2], mji);
                                sfor<0, 6>([&](auto rc) { F[j][rc] = tmp[rc]; });
                        }
                }
        }
        static GATO_HD T minv_sym(const T (&Minv)[NQ * NQ], int row, int col) { return (row <= col) ? Minv[col * NQ + row] : Minv[row * NQ + col]; }
        static GATO_HD void fd_finish(const T (&Minv)[NQ * NQ], const T* u, const V6& f, T (&qdd)[NQ])
        {
                T tau[NQ];  // compile-time indices only: registers
                sfor<0, NQ>([&](auto cc) { tau[cc] = L::sub(u[cc], f[cc][2]); });
                GATO_ROLLED
                for (int row = 0; row < NQ; row++) {
                        T val = L::set(0.0f);
                        sfor<0, NQ>([&](auto cc) { val = L::fma(minv_sym(Minv, row, cc), tau[cc], val); });
                        qdd[row] = val;
                }
        }
        struct FdWork {
                T        t[2 * NQ];
                Xmat     X;
                T        Minv[NQ * NQ];
                MinvWork mw;
                V6       f;
                T        qdd[NQ];
        };
        // forwardDynamics with wrench (iiwa14_plant.cuh:171-180)
        static GATO_HD void forward_dynamics(const RtModel& m, const T* q, const T* qd, const T* u, const T* fext, T (&qdd)[NQ])
        {
                FdWork w;
                update_X(m, q, w.t, w.X);
                minv(m, w.X, w.mw, w.Minv);
                rnea<false, false>(m, w.X, qd, nullptr, fext, nullptr, nullptr, w.f);
                fd_finish(w.Minv, u, w.f, w.qdd);  // (fd_finish indexes its output dynamically: it stays inside the work struct)
                sfor<0, NQ>([&](auto ic) { qdd[ic] = w.qdd[ic]; });
        }

        // forwardDynamicsAndGradient with wrench (iiwa14_plant.cuh:229-268): everything the gradient columns share
        // scratch of one gradient column (rnea_grad_col and the M^-1 product that follows it)
        struct ColWork {
                T df[NQ][6], dc[NQ], d[NQ];
        };
        struct DynState {
                Xmat X;
                T    Minv[NQ * NQ];
                V6   v, a, f, Iv;
                T    FxvI[NQ][36];
                T    qdd[NQ];
                T    t[2 * NQ];
                union {  // the prologue's M^-1 scratch is dead when the columns start
                        MinvWork mw;
                        ColWork  cw;
                };
        };
        static GATO_HD void dyn_prologue(const RtModel& m, const T* q, const T* qd, const T* u, const T* fext, DynState& st)
        {
                update_X(m, q, st.t, st.X);
                minv(m, st.X, st.mw, st.Minv);
                rnea<false, false>(m, st.X, qd, nullptr, fext, nullptr, nullptr, st.f);
                fd_finish(st.Minv, u, st.f, st.qdd);
                rnea<true, true>(m, st.X, qd, st.qdd, fext, &st.v, &st.a, st.f);
                GATO_ROLLED
                for (int j = 0; j < NQ; j++) {
                        T vj[6];
                        sfor<0, 6>([&](auto rc) { vj[rc] = st.v[j][rc]; });
                        sfor<0, 6>([&](auto rc) { st.Iv[j][rc] = irow<rc>(m.I[j], vj); });
                        sfor<0, 6>([&](auto cc) {
                                constexpr int c = cc;
                                T             col[6], out[6];
                                sfor<0, 6>([&](auto tc) { col[tc] = L::set(m.I[j][6 * c + tc]); });
                                fx_times_v(out, vj, col);
                                sfor<0, 6>([&](auto rc) { st.FxvI[j][6 * c + rc] = out[rc]; });
                        });
                }
        }
        // ---- RNEA gradient (inverse_dynamics_gradient_inner), column k of d c / d{q | qd} (W = 0 | 1) ---------------------
        // result in st.cw.dc
        template<int W>
        static GATO_HD void rnea_grad_col(const RtModel& m, int k, DynState& st, const T* qd)
        {
                T(&df)[NQ][6] = st.cw.df;
                T(&dc)[NQ] = st.cw.dc;
                T dv[6], da[6];
                GATO_ROLLED
                for (int j = 0; j < NQ; j++) sfor<0, 6>([&](auto rc) { df[j][rc] = L::set(0.0f); });
                sfor<0, 6>([&](auto rc) { dv[rc] = L::set(0.0f), da[rc] = L::set(0.0f); });
                const T z = L::set(0.0f);
                GATO_ROLLED
                for (int j = k; j < NQ; j++) {
                        T Xj[18], ndv[6], nda[6];
                        sfor<0, 18>([&](auto ec) { Xj[ec] = st.X[j][ec]; });
                        const T qdj = qd[j];
                        if (j == k) {
                                // own column: dv = mx2(X v_parent) | S ; da = mx2_scaled(dv, qd) + { mx2(X a_parent) | mx2(v) }
                                T src[6];
                                if constexpr (W == 0) {
                                        T Xv[6], Xa[6];
                                        if (j == 0) {
                                                sfor<0, 6>([&](auto rc) { Xv[rc] = z, Xa[rc] = gravity_row<rc>(Xj); });
                                        } else {
                                                T vp[6], ap[6];
                                                sfor<0, 6>([&](auto rc) { vp[rc] = st.v[j - 1][rc], ap[rc] = st.a[j - 1][rc]; });
                                                sfor<0, 6>([&](auto rc) { Xv[rc] = xrow<rc>(Xj, vp), Xa[rc] = xrow<rc>(Xj, ap); });
                                        }
                                        ndv[0] = Xv[1], ndv[1] = L::neg(Xv[0]), ndv[2] = z, ndv[3] = Xv[4], ndv[4] = L::neg(Xv[3]), ndv[5] = z;
                                        src[0] = Xa[1], src[1] = L::neg(Xa[0]), src[2] = z, src[3] = Xa[4], src[4] = L::neg(Xa[3]), src[5] = z;
                                        if (j == 0) sfor<0, 6>([&](auto rc) { ndv[rc] = z; });
                                } else {
                                        sfor<0, 6>([&](auto rc) { ndv[rc] = L::set(rc == 2 ? 1.0f : 0.0f); });
                                        src[0] = st.v[j][1], src[1] = L::neg(st.v[j][0]), src[2] = z, src[3] = st.v[j][4], src[4] = L::neg(st.v[j][3]), src[5] = z;
                                }
                                nda[0] = L::mul(ndv[1], qdj), nda[1] = L::mul(L::neg(ndv[0]), qdj), nda[2] = z;
                                nda[3] = L::mul(ndv[4], qdj), nda[4] = L::mul(L::neg(ndv[3]), qdj), nda[5] = z;
                                sfor<0, 6>([&](auto rc) { nda[rc] = L::add(nda[rc], src[rc]); });
                        } else {
                                sfor<0, 6>([&](auto rc) { ndv[rc] = xrow<rc>(Xj, dv); });
                                nda[0] = L::mul(ndv[1], qdj), nda[1] = L::mul(L::neg(ndv[0]), qdj), nda[2] = z;
                                nda[3] = L::mul(ndv[4], qdj), nda[4] = L::mul(L::neg(ndv[3]), qdj), nda[5] = z;
                                sfor<0, 6>([&](auto rc) { nda[rc] = L::add(nda[rc], xrow<rc>(Xj, da)); });
                        }
                        sfor<0, 6>([&](auto rc) { dv[rc] = ndv[rc], da[rc] = nda[rc]; });
                        // df_j = fx(dv) I v  +  ( I da + (fx(v) I) dv )
                        T t0[6], Ivj[6];
                        sfor<0, 6>([&](auto rc) { Ivj[rc] = st.Iv[j][rc]; });
                        fx_times_v(t0, dv, Ivj);
                        sfor<0, 6>([&](auto rc) {
                                constexpr int row = rc;
                                const T       d1 = irow<row>(m.I[j], da);
                                T             d2 = z;
                                sfor<0, 6>([&](auto tc) { d2 = L::fma(st.FxvI[j][row + 6 * tc], dv[tc], d2); });
                                df[j][row] = L::add(t0[row], L::add(d1, d2));
                        });
                }
                // backward: df_{j-1} += X_j^T df_j (+ -X_j^T mx2(f_j) on the own dq column)
                GATO_ROLLED
                for (int j = NQ - 1; j >= 1; j--) {
                        T Xj[18], upd[6], dfj[6];
                        sfor<0, 18>([&](auto ec) { Xj[ec] = st.X[j][ec]; });
                        sfor<0, 6>([&](auto rc) { dfj[rc] = df[j][rc]; });
                        sfor<0, 6>([&](auto rc) { upd[rc] = xcol<rc>(Xj, dfj); });
                        if constexpr (W == 0) {
                                if (k == j) {
                                        T mxf[6] = {st.f[j][1], L::neg(st.f[j][0]), z, st.f[j][4], L::neg(st.f[j][3]), z};
                                        sfor<0, 6>([&](auto rc) { upd[rc] = L::add(upd[rc], L::neg(xcol<rc>(Xj, mxf))); });
                                }
                        }
                        sfor<0, 6>([&](auto rc) { df[j - 1][rc] = L::add(df[j - 1][rc], upd[rc]); });
                }
                GATO_ROLLED
                for (int j = 0; j < NQ; j++) dc[j] = df[j][2];
        }

        // ---- end-effector position and positional Jacobian (end_effector_pose(_gradient)_inner; oracle: ee_pos_grad) ----------
        // Xh[j] / dXh[j]: the 4x4 homogeneous transform of joint j and its derivative at q_j
        struct EeWork {
                T t[2 * NQ], Xh[NQ][16], dXh[NQ][16], J[NQ][3];
        };
        static GATO_HD void update_Xhom(const RtModel& m, const T* q, T (&t)[2 * NQ], T (&Xh)[NQ][16], T (*dXh)[16])
        {
                sincos(q, t);
                GATO_ROLLED
                for (int j = 0; j < NQ; j++) {
                        sfor<0, 16>([&](auto ec) { Xh[j][ec] = L::set(m.Xh[j][ec]); });
                        const int n = m.nxht[j];
                        GATO_ROLLED
                        for (int i = 0; i < n; i++) Xh[j][m.xht[j][i].loc] = trig_entry(m.xht[j][i], t);
                        if (dXh) {
                                sfor<0, 16>([&](auto ec) { dXh[j][ec] = L::set(m.dXh[j][ec]); });
                                const int nd = m.ndxht[j];
                                GATO_ROLLED
                                for (int i = 0; i < nd; i++) dXh[j][m.dxht[j][i].loc] = trig_entry(m.dxht[j][i], t);
                        }
                }
        }
        static GATO_HD void hom_apply(const T (&M)[16], T (&p)[4])
        {
                T o[4];
                sfor<0, 4>([&](auto rc) {
                        constexpr int row = rc;
                        T             r = L::set(0.0f);
                        sfor<0, 4>([&](auto ic) { r = L::fma(M[4 * ic + row], p[ic], r); });
                        o[row] = r;
                });
                sfor<0, 4>([&](auto rc) { p[rc] = o[rc]; });
        }
        // d < 0: position; d >= 0: derivative w.r.t. joint d
        static GATO_HD void ee_chain(const T (&Xh)[NQ][16], const T (*dXh)[16], int d, T (&out)[3])
        {
                T p[4];
                sfor<0, 4>([&](auto rc) { p[rc] = (d == NQ - 1) ? dXh[NQ - 1][12 + rc] : Xh[NQ - 1][12 + rc]; });
                GATO_ROLLED
                for (int j = NQ - 2; j >= 0; j--) {
                        T M[16];
                        if (j == d)
                                sfor<0, 16>([&](auto ec) { M[ec] = dXh[j][ec]; });
                        else
                                sfor<0, 16>([&](auto ec) { M[ec] = Xh[j][ec]; });
                        hom_apply(M, p);
                }
                out[0] = p[0], out[1] = p[1], out[2] = p[2];
        }
        struct EePosWork {
                T t[2 * NQ], Xh[NQ][16];
        };
        static GATO_HD void ee_pos(const RtModel& m, const T* q, T (&ee)[3])
        {
                EePosWork w;
                update_Xhom(m, q, w.t, w.Xh, nullptr);
                ee_chain(w.Xh, nullptr, -1, ee);
        }
        // position and Jacobian (w.J[d] = d ee / d q_d)
        static GATO_HD void ee_pos_grad(const RtModel& m, const T* q, T (&ee)[3], EeWork& w)
        {
                update_Xhom(m, q, w.t, w.Xh, w.dXh);
                ee_chain(w.Xh, w.dXh, -1, ee);
                GATO_ROLLED
                for (int d = 0; d < NQ; d++) ee_chain(w.Xh, w.dXh, d, w.J[d]);
        }

        // ---- trapezoidal integrator  (integrator.cuh:34-37, 143-184) ---------------------------------------
        static GATO_HD void integrate(const T* q, const T* qd, const T (&qdd)[NQ], float dt, T (&qn)[NQ], T (&qdn)[NQ])
        {
                sfor<0, NQ>([&](auto ic) {
                        constexpr int i = ic;
                        qdn[i] = L::fmas(dt, qdd[i], qd[i]);
                        const T lin = L::fmas(dt, qd[i], q[i]);
                        qn[i] = L::map2(qdd[i], lin, [&](float a, float l) {
                                const double acc = ((double)a * 0.5) * (double)dt;
                                return (float)fma(acc, (double)dt, (double)l);
                        });
                });
        }
};

// ---- barriers with run-time limits and style (iiwa14_plant.cuh:103-155, indy7_plant.cuh:133-147) ------------------------------
GATO_HD float rt_joint_barrier(float q, float lo, float hi)
{
        float dmin = q - lo, dmax = hi - q;
        dmin = ((double)dmin <= 1e-10) ? (float)1e-10 : dmin;
        dmax = ((double)dmax <= 1e-10) ? (float)1e-10 : dmax;
        return (-g_log(dmin)) - g_log(dmax);
}
GATO_HD float rt_joint_barrier_grad(int style, float q, float lo, float hi)
{
        float dmin = q - lo, dmax = hi - q;
        if (style == 1) {
                const float eps = 1e-6f;
                if (dmin >= 0.0f) {
                        if (dmin < eps) dmin = eps;
                } else {
                        if (dmin > -eps) dmin = -eps;
                }
                if (dmax >= 0.0f) {
                        if (dmax < eps) dmax = eps;
                } else {
                        if (dmax > -eps) dmax = -eps;
                }
        } else {
                dmin = ((double)dmin <= 1e-6) ? (float)1e-6 : dmin;
                dmax = ((double)dmax <= 1e-6) ? (float)1e-6 : dmax;
        }
        return (-1.0f / dmin) + (1.0f / dmax);
}
GATO_HD float rt_joint_barrier_hess(float q, float lo, float hi)
{
        float       dmin = q - lo, dmax = hi - q;
        const float eps = 1e-6f;
        float       amin = dmin >= 0.0f ? dmin : -dmin, amax = dmax >= 0.0f ? dmax : -dmax;
        if (amin < eps) amin = eps;
        if (amax < eps) amax = eps;
        return 1.0f / (amin * amin) + 1.0f / (amax * amax);
}

}  // namespace gato
