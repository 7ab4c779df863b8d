// Rigid-body dynamics for fixed-base serial chains of z-axis revolute joints, specialised at compile
// time per robot from data tables (robot_model_data.h) — the building blocks of the BSQP hot path:
//   RNEA, direct M^-1, forward dynamics, analytical RNEA gradient, end-effector position + Jacobian,
//   trapezoidal integrator and its Jacobians, joint-limit barriers, tracking cost.
//
// Replaces (same results, different program): the GRiD-generated device functions the reference calls —
//   load_update_XImats_helpers iiwa14_grid.cuh:2212, direct_minv_inner :4742, inverse_dynamics_inner(_vaf)
//   iiwa14_fext.cuh:29/271, forward_dynamics_finish grid:5341, inverse_dynamics_gradient_inner :5549,
//   end_effector_pose(_gradient)_inner :2596/:2855 (indy7: indy7_grid.cuh:1597,2918,2281,2497,3322,3373,1834,1933),
//   the plant layer iiwa14_plant.cuh:103-450 / indy7_plant.cuh and integrator.cuh:20-257.
//
// Design (B200-first): ONE THREAD PER WORK ITEM (a knot, or a knot x line-search step).  Everything is
// fully unrolled over joints with the robot's constants folded into immediates and its structural zeros
// skipped at compile time, so there are no block barriers, no shuffles and no shared-memory staging in
// the dynamics (the reference spends 113 __syncthreads per linearisation with 6-98 of 128-352 threads
// active).  Arithmetic contract: fp32, every fused multiply-add is an explicit fmaf() (translation units
// including this header are compiled with -fmad=false), term order identical to the CPU oracle, so that
// results are bit-identical to the oracle and integer outcomes (PCG counts, line-search indices) match.
#pragma once

#include <cmath>
#include <cstdint>
#include <type_traits>

#include "costs.h"
#include "sfor.h"
#include "robot_model_data.h"

namespace gato {

// ---- scalar math ---------------------------------------------------------------------------------
// Device: CUDA's precise sinf/cosf/logf.  GATO_HOST_TEST (tests only) lets a host build inject
// bit-equivalent stand-ins so this header can be unit-tested on a CPU-only machine.
#if defined(GATO_HOST_TEST) && !defined(__CUDA_ARCH__)
extern "C" float gato_host_sinf(float);
extern "C" float gato_host_cosf(float);
extern "C" float gato_host_logf(float);
GATO_HD float g_sin(float x) { return gato_host_sinf(x); }
GATO_HD float g_cos(float x) { return gato_host_cosf(x); }
GATO_HD float g_log(float x) { return gato_host_logf(x); }
#else
GATO_HD float g_sin(float x) { return sinf(x); }
GATO_HD float g_cos(float x) { return cosf(x); }
GATO_HD float g_log(float x) { return logf(x); }
#endif

// ---- packed pairs ----------------------------------------------------------------------------------
// Two fp32 lanes through ONE instruction: Blackwell's FFMA2 / FADD2 / FMUL2 (fma.rn.f32x2 ...) round each lane exactly like the scalar
// operation, so "two vectors through the same matrix" costs half the instructions with bit-identical results -- what matters for kernels
// that are bound by instruction fetch and issue.  The host build (tests) evaluates the two lanes one after the other.
#if defined(__CUDACC__)
using f2 = float2;
#else
struct f2 {
        float x, y;
};
#endif
GATO_HD f2 mk2(float x, float y)
{
        f2 r;
        r.x = x, r.y = y;
        return r;
}
#ifndef GATO_F2_EMULATE
#define GATO_F2_EMULATE 0  // debugging aid: bit 0 / 1 / 2 evaluate fma2 / add2 / mul2 lane by lane with scalar instructions on the device too
#endif
GATO_HD f2 fma2(f2 a, f2 b, f2 c)
{
#if defined(__CUDA_ARCH__) && !(GATO_F2_EMULATE & 1)
        return __ffma2_rn(a, b, c);
#else
        return mk2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
GATO_HD f2 fma2s(float s, f2 b, f2 c) { return fma2(mk2(s, s), b, c); }  // scalar coefficient, broadcast
GATO_HD f2 add2(f2 a, f2 b)
{
#if defined(__CUDA_ARCH__) && !(GATO_F2_EMULATE & 2)
        return __fadd2_rn(a, b);
#else
        return mk2(a.x + b.x, a.y + b.y);
#endif
}
// The product is evaluated as two SCALAR multiplies on purpose.  ptxas 12.9 fuses a packed multiply with a following packed add into one
// FFMA2 -- even when the PTX says mul.rn.f32x2 / add.rn.f32x2 with explicit rounding modifiers, which forbid contraction, and even under
// -fmad=false (seen in SASS; it also folds fma.rn.f32x2(a, b, -0.0) back into a multiply first).  One rounding instead of two changes last
// bits against the scalar path and the oracle (found as 1-ulp differences in A between k_kkt and k_kkt_fine).  Scalar FMUL + FADD2 is left alone.
GATO_HD f2 mul2(f2 a, f2 b) { return mk2(a.x * b.x, a.y * b.y); }
GATO_HD f2 mul2s(float s, f2 b) { return mul2(mk2(s, s), b); }
GATO_HD f2 neg2(f2 a) { return mk2(-a.x, -a.y); }

constexpr float kGravity = 9.81f;  // iiwa14_plant.cuh:25-28

// ---- robot description -----------------------------------------------------------------------------
struct Iiwa14 {
        static constexpr int                    ID = 1, NQ = 7;
        static constexpr const double*          TABLE = iiwa14_TABLE;
        static constexpr const gato_trig_entry *XT = iiwa14_X_TRIG, *XHT = iiwa14_XH_TRIG, *DXHT = iiwa14_DXH_TRIG;
        static constexpr int                    NXT = iiwa14_X_TRIG_LEN, NXHT = iiwa14_XH_TRIG_LEN, NDXHT = iiwa14_DXH_TRIG_LEN;
        // iiwa14_plant.cuh:36-70
        static constexpr double JL[7] = {2.96706, 2.09440, 2.96706, 2.09440, 2.96706, 2.09440, 3.05433};
        static constexpr double VL[7] = {1.48353, 1.48353, 1.74533, 1.30900, 2.26893, 2.35619, 2.35619};
        static constexpr double CL[7] = {320.0, 320.0, 176.0, 176.0, 110.0, 40.0, 40.0};
};
struct Indy7 {
        static constexpr int                    ID = 0, NQ = 6;
        static constexpr const double*          TABLE = indy7_TABLE;
        static constexpr const gato_trig_entry *XT = indy7_X_TRIG, *XHT = indy7_XH_TRIG, *DXHT = indy7_DXH_TRIG;
        static constexpr int                    NXT = indy7_X_TRIG_LEN, NXHT = indy7_XH_TRIG_LEN, NDXHT = indy7_DXH_TRIG_LEN;
        // indy7_plant.cuh:66-96
        static constexpr double JL[6] = {3.0543, 3.0543, 3.0543, 3.0543, 3.0543, 3.7520};
        static constexpr double VL[6] = {2.61, 2.61, 2.61, 3.14, 3.14, 3.14};
        static constexpr double CL[6] = {431.97, 431.97, 197.23, 79.79, 79.79, 79.79};
};

constexpr double kMargin = (double)(float)(-0.1);  // JOINT_LIMIT_MARGIN<float>()
template<class P, int WHICH, int J, int SIDE>  // WHICH 0 joint, 1 velocity, 2 control; SIDE 0 lower, 1 upper
constexpr float limit()
{
        const double L = WHICH == 0 ? P::JL[J] : (WHICH == 1 ? P::VL[J] : P::CL[J]);
        return SIDE == 0 ? (float)(-L - kMargin) : (float)(L + kMargin);
}

constexpr int trig_find(const gato_trig_entry* t, int n, int idx)
{
        for (int i = 0; i < n; i++)
                if (t[i].idx == idx) return i;
        return -1;
}

// Compile-time description of entry (R,C) of the 6x6 Plücker transform X_J(q_J) (col-major table layout,
// bottom-right 3x3 = copy of top-left, iiwa14_grid.cuh:2287-2291).
template<class P, int J, int R, int C>
struct XE {
        static constexpr bool   br = (R >= 3 && C >= 3);
        static constexpr int    r0 = br ? R - 3 : R, c0 = br ? C - 3 : C;
        static constexpr int    loc = 6 * c0 + r0;
        static constexpr int    ti = trig_find(P::XT, P::NXT, 36 * J + loc);
        static constexpr float  cst = (float)P::TABLE[36 * J + loc];
        static constexpr bool   nz = (ti >= 0) || (cst != 0.0f);
        static constexpr double coef = ti >= 0 ? P::XT[ti >= 0 ? ti : 0].coef : 0.0;
        static constexpr int    tk = ti >= 0 ? P::XT[ti >= 0 ? ti : 0].k : 0;
};
template<class P, int J, int R, int C>
constexpr float inertia()
{
        return (float)P::TABLE[36 * P::NQ + 36 * J + 6 * C + R];
}
// 4x4 homogeneous transforms (col-major, idx = 4*C + R)
template<class P, int J, int R, int C, bool DERIV>
struct HE {
        static constexpr int    base = 72 * P::NQ + (DERIV ? 16 * P::NQ : 0);
        static constexpr int    loc = 4 * C + R;
        static constexpr int    ti = DERIV ? trig_find(P::DXHT, P::NDXHT, 16 * J + loc) : trig_find(P::XHT, P::NXHT, 16 * J + loc);
        static constexpr float  cst = (float)P::TABLE[base + 16 * J + loc];
        static constexpr bool   nz = (ti >= 0) || (cst != 0.0f);
        static constexpr double coef = ti >= 0 ? (DERIV ? P::DXHT : P::XHT)[ti >= 0 ? ti : 0].coef : 0.0;
        static constexpr int    tk = ti >= 0 ? (DERIV ? P::DXHT : P::XHT)[ti >= 0 ? ti : 0].k : 0;
};

// entry = (float)(coef * (double)t) with the +-1 cases exact without the fp64 round trip
template<class E>
GATO_HD float trig_entry(const float* t)
{
        if constexpr (E::coef == 1.0)
                return t[E::tk];
        else if constexpr (E::coef == -1.0)
                return -t[E::tk];
        else
                return (float)(E::coef * (double)t[E::tk]);
}

template<class P>
struct Rbd {
        static constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ;
        using Xmat = float[NQ][36];  // only structurally non-zero, non-aliased entries are ever touched
        using V6 = float[NQ][6];

        // ---- sin/cos + X update  (load_update_XImats_helpers) ------------------------------------
        static GATO_HD void sincos(const float* q, float (&t)[2 * NQ])
        {
                sfor<0, NQ>([&](auto kc) {
                        constexpr int k = kc;
                        t[k] = g_sin(q[k]);
                        t[k + NQ] = g_cos(q[k]);
                });
        }
        static GATO_HD void update_X(const float (&t)[2 * NQ], Xmat& X)
        {
                sfor<0, NQ>([&](auto jc) {
                        constexpr int j = jc;
                        sfor<0, 36>([&](auto ec) {
                                constexpr int e = ec;
                                using E = XE<P, j, e % 6, e / 6>;
                                if constexpr (!E::br && E::ti >= 0) X[j][e] = trig_entry<E>(t);
                        });
                });
        }
        template<int J, int R, int C>
        static GATO_HD float xv(const Xmat& X)
        {
                using E = XE<P, J, R, C>;
                if constexpr (E::ti >= 0)
                        return X[J][E::loc];
                else
                        return E::cst;
        }
        // (X_J v)[R]: ascending-index fma chain from +0, structural zeros skipped
        template<int J, int R>
        static GATO_HD float xrow(const Xmat& X, const float (&v)[6])
        {
                float r = 0.0f;
                sfor<0, 6>([&](auto ic) {
                        constexpr int i = ic;
                        if constexpr (XE<P, J, R, i>::nz) r = fmaf(xv<J, R, i>(X), v[i], r);
                });
                return r;
        }
        // (X_J^T f)[C]
        template<int J, int C>
        static GATO_HD float xcol(const Xmat& X, const float (&f)[6])
        {
                float r = 0.0f;
                sfor<0, 6>([&](auto ic) {
                        constexpr int i = ic;
                        if constexpr (XE<P, J, i, C>::nz) r = fmaf(xv<J, i, C>(X), f[i], r);
                });
                return r;
        }
        // (I_J v)[R]
        template<int J, int R>
        static GATO_HD float irow(const float (&v)[6])
        {
                float r = 0.0f;
                sfor<0, 6>([&](auto ic) {
                        constexpr int i = ic;
                        if constexpr (inertia<P, J, R, i>() != 0.0f) r = fmaf(inertia<P, J, R, i>(), v[i], r);
                });
                return r;
        }
        // the same three products for a PAIR of vectors (lane x and lane y of every entry): identical chains, one packed instruction per term
        template<int J, int R>
        static GATO_HD f2 xrow2(const Xmat& X, const f2 (&v)[6])
        {
                f2 r = mk2(0.0f, 0.0f);
                sfor<0, 6>([&](auto ic) {
                        constexpr int i = ic;
                        if constexpr (XE<P, J, R, i>::nz) r = fma2s(xv<J, R, i>(X), v[i], r);
                });
                return r;
        }
        template<int J, int C>
        static GATO_HD f2 xcol2(const Xmat& X, const f2 (&f)[6])
        {
                f2 r = mk2(0.0f, 0.0f);
                sfor<0, 6>([&](auto ic) {
                        constexpr int i = ic;
                        if constexpr (XE<P, J, i, C>::nz) r = fma2s(xv<J, i, C>(X), f[i], r);
                });
                return r;
        }
        template<int J, int R>
        static GATO_HD f2 irow2(const f2 (&v)[6])
        {
                f2 r = mk2(0.0f, 0.0f);
                sfor<0, 6>([&](auto ic) {
                        constexpr int i = ic;
                        if constexpr (inertia<P, J, R, i>() != 0.0f) r = fma2s(inertia<P, J, R, i>(), v[i], r);
                });
                return r;
        }
        // fx(f) * t for a pair of f and one t (the lanes of f are two gradient columns, t = I v of the joint)
        static GATO_HD void fx2_times_v(f2 (&r)[6], const f2 (&f)[6], const float (&t)[6])
        {
                f2 s;
                s = fma2s(t[2], f[1], neg2(mul2s(t[1], f[2])));
                s = fma2s(t[4], neg2(f[5]), s);
                r[0] = fma2s(t[5], f[4], s);
                s = fma2s(t[0], f[2], neg2(mul2s(t[2], f[0])));
                s = fma2s(t[3], f[5], s);
                r[1] = fma2s(t[5], neg2(f[3]), s);
                s = fma2s(t[1], f[0], neg2(mul2s(t[0], f[1])));
                s = fma2s(t[3], neg2(f[4]), s);
                r[2] = fma2s(t[4], f[3], s);
                r[3] = fma2s(t[5], f[1], neg2(mul2s(t[4], f[2])));
                r[4] = fma2s(t[3], f[2], neg2(mul2s(t[5], f[0])));
                r[5] = fma2s(t[4], f[0], neg2(mul2s(t[3], f[1])));
        }

        // fx(f) * t   (iiwa14_grid.cuh:896-905) with the fma placement nvcc emits for it
        static GATO_HD void fx_times_v(float (&r)[6], const float (&f)[6], const float (&t)[6])
        {
                float s;
                s = fmaf(f[1], t[2], -(f[2] * t[1]));
                s = fmaf(-f[5], t[4], s);
                r[0] = fmaf(f[4], t[5], s);
                s = fmaf(f[2], t[0], -(f[0] * t[2]));
                s = fmaf(f[5], t[3], s);
                r[1] = fmaf(-f[3], t[5], s);
                s = fmaf(f[0], t[1], -(f[1] * t[0]));
                s = fmaf(-f[4], t[3], s);
                r[2] = fmaf(f[3], t[4], s);
                r[3] = fmaf(f[1], t[5], -(f[2] * t[4]));
                r[4] = fmaf(f[2], t[3], -(f[0] * t[5]));
                r[5] = fmaf(f[0], t[4], -(f[1] * t[3]));
        }

        // ---- RNEA  (inverse_dynamics_inner / _vaf with external wrench) ---------------------------
        template<bool WITH_QDD>
        static GATO_HD void rnea(const Xmat& X, const float* qd, const float* qdd, const float* fext, V6& v, V6& a, V6& f)
        {
                sfor<0, 6>([&](auto rc) {
                        constexpr int row = rc;
                        v[0][row] = 0.0f;
                        a[0][row] = XE<P, 0, row, 5>::nz ? xv<0, row, 5>(X) * kGravity : 0.0f;
                });
                v[0][2] = v[0][2] + qd[0];
                if constexpr (WITH_QDD) a[0][2] = a[0][2] + qdd[0];
                sfor<1, NQ>([&](auto jc) {
                        constexpr int j = jc;
                        sfor<0, 6>([&](auto rc) {
                                constexpr int row = rc;
                                float         vv = xrow<j, row>(X, v[j - 1]);
                                float         aa = xrow<j, row>(X, a[j - 1]);
                                if constexpr (row == 2) {
                                        vv = vv + qd[j];
                                        if constexpr (WITH_QDD) aa = aa + qdd[j];
                                }
                                v[j][row] = vv;
                                a[j][row] = aa;
                        });
                        a[j][0] = fmaf(v[j][1], qd[j], a[j][0]);
                        a[j][1] = fmaf(-v[j][0], qd[j], a[j][1]);
                        a[j][3] = fmaf(v[j][4], qd[j], a[j][3]);
                        a[j][4] = fmaf(-v[j][3], qd[j], a[j][4]);
                });
                sfor<0, NQ>([&](auto jc) {
                        constexpr int j = jc;
                        float         Iv[6], t[6];
                        sfor<0, 6>([&](auto rc) {
                                constexpr int row = rc;
                                f[j][row] = irow<j, row>(a[j]);
                                Iv[row] = irow<j, row>(v[j]);
                        });
                        fx_times_v(t, v[j], Iv);
                        sfor<0, 6>([&](auto rc) {
                                constexpr int row = rc;
                                f[j][row] = f[j][row] + t[row];
                                if constexpr (j == NQ - 1) f[j][row] = f[j][row] - fext[row];
                        });
                });
                sfor_down<1, NQ>([&](auto jc) {
                        constexpr int j = jc;
                        float         val[6];
                        sfor<0, 6>([&](auto rc) { val[rc] = xcol<j, rc>(X, f[j]); });
                        sfor<0, 6>([&](auto rc) { f[j - 1][rc] = f[j - 1][rc] + val[rc]; });
                });
        }

        // ---- direct M^-1  (direct_minv_inner); Minv[col*NQ+row], upper triangle valid ---------------
        static GATO_HD void minv(const Xmat& X, float (&Minv)[NQ * NQ])
        {
                float IA[36], U[NQ][6], Dinv[NQ];
                float F[NQ][6];  // F[i][:, j] of the current level i, indexed by column j
                sfor<0, NQ * NQ>([&](auto ic) { Minv[ic] = 0.0f; });
                sfor<0, NQ>([&](auto jc) { sfor<0, 6>([&](auto rc) { F[jc][rc] = 0.0f; }); });
                sfor<0, 36>([&](auto ec) { IA[ec] = inertia<P, NQ - 1, ec % 6, ec / 6>(); });
                sfor_down<0, NQ>([&](auto ic_) {
                        constexpr int i = ic_;
                        sfor<0, 6>([&](auto rc) { U[i][rc] = IA[12 + rc]; });
                        Dinv[i] = 1.0f / U[i][2];
                        Minv[i * NQ + i] = Dinv[i];
                        sfor<i, NQ>([&](auto jc) {
                                constexpr int j = jc;
                                Minv[j * NQ + i] = fmaf(-Dinv[i], F[j][2], Minv[j * NQ + i]);
                                if constexpr (i > 0) sfor<0, 6>([&](auto rc) { F[j][rc] = fmaf(U[i][rc], Minv[j * NQ + i], F[j][rc]); });
                        });
                        if constexpr (i > 0) {
                                float Ia[36], IaT[36];
                                sfor<0, 36>([&](auto ec) {
                                        constexpr int row = ec % 6, col = ec / 6;
                                        Ia[ec] = fmaf(-(U[i][row] * Dinv[i]), U[i][col], IA[ec]);
                                });
                                // F[i-1][:, j] = X_i^T F[i][:, j]  (columns j >= i; the parent's own column i-1 starts at zero)
                                sfor<i, NQ>([&](auto jc) {
                                        constexpr int j = jc;
                                        float         tmp[6];
                                        sfor<0, 6>([&](auto rc) { tmp[rc] = xcol<i, rc>(X, F[j]); });
                                        sfor<0, 6>([&](auto rc) { F[j][rc] = tmp[rc]; });
                                });
                                sfor<0, 6>([&](auto rc) { F[i - 1][rc] = 0.0f; });
                                sfor<0, 6>([&](auto cc) {
                                        constexpr int c = cc;
                                        float         col[6];
                                        sfor<0, 6>([&](auto tc) { col[tc] = Ia[6 * c + tc]; });
                                        sfor<0, 6>([&](auto rc) { IaT[6 * c + rc] = xcol<i, rc>(X, col); });
                                });
                                // IA[i-1] = I[i-1] + IaT * X_i
                                sfor<0, 36>([&](auto ec) {
                                        constexpr int row = ec % 6, col = ec / 6;
                                        float         val = 0.0f;
                                        sfor<0, 6>([&](auto tc) {
                                                constexpr int t = tc;
                                                if constexpr (XE<P, i, t, col>::nz) val = fmaf(IaT[row + 6 * t], xv<i, t, col>(X), val);
                                        });
                                        IA[ec] = inertia<P, i - 1, row, col>() + val;
                                });
                        }
                });
                // forward pass: F[j] now holds column j of the current level's F
                sfor<0, NQ>([&](auto jc) {
                        constexpr int j = jc;
                        sfor<0, 6>([&](auto rc) { F[j][rc] = (rc == 2 ? 1.0f : 0.0f) * Minv[j * NQ]; });
                });
                sfor<1, NQ>([&](auto ic_) {
                        constexpr int i = ic_;
                        sfor<i, NQ>([&](auto jc) {
                                constexpr int j = jc;
                                float         tmp[6];
                                sfor<0, 6>([&](auto rc) { tmp[rc] = xrow<i, rc>(X, F[j]); });
                                sfor<0, 6>([&](auto rc) { F[j][rc] = tmp[rc]; });
                                float d = 0.0f;
                                sfor<0, 6>([&](auto tc) { d = fmaf(F[j][tc], U[i][tc], d); });
                                Minv[j * NQ + i] = fmaf(-Dinv[i], d, Minv[j * NQ + i]);
                                if constexpr (i < NQ - 1) F[j][2] = F[j][2] + Minv[j * NQ + i];
                        });
                });
        }
        template<int ROW, int COL>
        static GATO_HD float minv_sym(const float (&Minv)[NQ * NQ])
        {
                return (ROW <= COL) ? Minv[COL * NQ + ROW] : Minv[ROW * NQ + COL];
        }
        static GATO_HD void fd_finish(const float (&Minv)[NQ * NQ], const float* u, const V6& f, float (&qdd)[NQ])
        {
                sfor<0, NQ>([&](auto rc) {
                        constexpr int row = rc;
                        float         val = 0.0f;
                        sfor<0, NQ>([&](auto cc) {
                                constexpr int col = cc;
                                val = fmaf(minv_sym<row, col>(Minv), (u[col] - f[col][2]), val);
                        });
                        qdd[row] = val;
                });
        }
        // forwardDynamics with wrench (iiwa14_plant.cuh:171-180)
        static GATO_HD void forward_dynamics(const float* q, const float* qd, const float* u, const float* fext, float (&qdd)[NQ])
        {
                float t[2 * NQ];
                Xmat  X;
                sincos(q, t);
                update_X(t, X);
                float Minv[NQ * NQ];
                minv(X, Minv);
                V6 v, a, f;
                rnea<false>(X, qd, nullptr, fext, v, a, f);
                fd_finish(Minv, u, f, qdd);
        }

        // ---- RNEA gradient (inverse_dynamics_gradient_inner), one (dq | dqd, column K) at a time ------------
        // dc[j] = d c_j / d{q|qd}_K for j = 0..NQ-1
        template<int W, int K>
        static GATO_HD void rnea_grad_col(const Xmat& X, const float* qd, const V6& v, const V6& a, const V6& f, const V6& Iv, const float (&FxvI)[NQ][36], float (&dc)[NQ])
        {
                float df[NQ][6];
                float dv[6], da[6];
                sfor<0, NQ>([&](auto jc) { sfor<0, 6>([&](auto rc) { df[jc][rc] = 0.0f; }); });
                sfor<K, NQ>([&](auto jc) {
                        constexpr int j = jc;
                        float         ndv[6], nda[6];
                        if constexpr (j == K) {
                                // own column: dv = mx2(X v_parent) | S ; da = mx2_scaled(dv, qd) + { mx2(X a_parent) | mx2(v) }
                                float src[6];
                                if constexpr (W == 0) {
                                        float Xv[6], Xa[6];
                                        sfor<0, 6>([&](auto rc) {
                                                constexpr int row = rc;
                                                if constexpr (j == 0) {
                                                        Xv[row] = 0.0f;
                                                        Xa[row] = XE<P, 0, row, 5>::nz ? xv<0, row, 5>(X) * kGravity : 0.0f;
                                                } else {
                                                        Xv[row] = xrow<j, row>(X, v[j > 0 ? j - 1 : 0]);
                                                        Xa[row] = xrow<j, row>(X, a[j > 0 ? j - 1 : 0]);
                                                }
                                        });
                                        ndv[0] = Xv[1], ndv[1] = -Xv[0], ndv[2] = 0.0f, ndv[3] = Xv[4], ndv[4] = -Xv[3], ndv[5] = 0.0f;
                                        src[0] = Xa[1], src[1] = -Xa[0], src[2] = 0.0f, src[3] = Xa[4], src[4] = -Xa[3], src[5] = 0.0f;
                                        if constexpr (j == 0) sfor<0, 6>([&](auto rc) { ndv[rc] = 0.0f; });
                                } else {
                                        sfor<0, 6>([&](auto rc) { ndv[rc] = (rc == 2) ? 1.0f : 0.0f; });
                                        src[0] = v[j][1], src[1] = -v[j][0], src[2] = 0.0f, src[3] = v[j][4], src[4] = -v[j][3], src[5] = 0.0f;
                                }
                                nda[0] = ndv[1] * qd[j], nda[1] = (-ndv[0]) * qd[j], nda[2] = 0.0f, nda[3] = ndv[4] * qd[j], nda[4] = (-ndv[3]) * qd[j], nda[5] = 0.0f;
                                sfor<0, 6>([&](auto rc) { nda[rc] = nda[rc] + src[rc]; });
                        } else {
                                sfor<0, 6>([&](auto rc) { ndv[rc] = xrow<j, rc>(X, dv); });
                                nda[0] = ndv[1] * qd[j], nda[1] = (-ndv[0]) * qd[j], nda[2] = 0.0f, nda[3] = ndv[4] * qd[j], nda[4] = (-ndv[3]) * qd[j], nda[5] = 0.0f;
                                sfor<0, 6>([&](auto rc) { nda[rc] = nda[rc] + xrow<j, rc>(X, da); });
                        }
                        sfor<0, 6>([&](auto rc) {
                                dv[rc] = ndv[rc];
                                da[rc] = nda[rc];
                        });
                        // df_j = fx(dv) I v  +  ( I da + (fx(v) I) dv )
                        float t0[6];
                        fx_times_v(t0, dv, Iv[j]);
                        sfor<0, 6>([&](auto rc) {
                                constexpr int row = rc;
                                float         d1 = irow<j, row>(da);
                                float         d2 = 0.0f;
                                sfor<0, 6>([&](auto tc) { d2 = fmaf(FxvI[j][row + 6 * tc], dv[tc], d2); });
                                df[j][row] = t0[row] + (d1 + d2);
                        });
                });
                // backward: df_{j-1} += X_j^T df_j (+ -X_j^T mx2(f_j) on the own dq column)
                sfor_down<1, NQ>([&](auto jc) {
                        constexpr int j = jc;
                        float         upd[6];
                        sfor<0, 6>([&](auto rc) { upd[rc] = xcol<j, rc>(X, df[j]); });
                        if constexpr (W == 0 && K == j) {
                                float mxf[6] = {f[j][1], -f[j][0], 0.0f, f[j][4], -f[j][3], 0.0f};
                                sfor<0, 6>([&](auto rc) { upd[rc] = upd[rc] + (-xcol<j, rc>(X, mxf)); });
                        }
                        sfor<0, 6>([&](auto rc) { df[j - 1][rc] = df[j - 1][rc] + upd[rc]; });
                });
                sfor<0, NQ>([&](auto jc) { dc[jc] = df[jc][2]; });
        }

        // The same column with K as a RUN-TIME value (uniform across the warp): one copy of the per-joint code serves all columns, so a
        // loop over k re-executes ~2k instructions that stay in the instruction caches instead of streaming 7 specialised copies (the
        // fully unrolled gradient is instruction-fetch bound).  Joints below k are skipped by uniform branches; every operation that is
        // executed is the one rnea_grad_col<W, k> executes, on the same values.
        template<int W>
        static GATO_HD void rnea_grad_col_rt(int k, const Xmat& X, const float* qd, const V6& v, const V6& a, const V6& f, const V6& Iv, const float (&FxvI)[NQ][36], float (&dc)[NQ])
        {
                float df[NQ][6];
                float dv[6], da[6];
                sfor<0, NQ>([&](auto jc) { sfor<0, 6>([&](auto rc) { df[jc][rc] = 0.0f; }); });
                sfor<0, 6>([&](auto rc) { dv[rc] = 0.0f, da[rc] = 0.0f; });
                sfor<0, NQ>([&](auto jc) {
                        constexpr int j = jc;
                        if (j >= k) {
                                float ndv[6], nda[6];
                                if (j == k) {
                                        float src[6];
                                        if constexpr (W == 0) {
                                                float Xv[6], Xa[6];
                                                sfor<0, 6>([&](auto rc) {
                                                        constexpr int row = rc;
                                                        if constexpr (j == 0) {
                                                                Xv[row] = 0.0f;
                                                                Xa[row] = XE<P, 0, row, 5>::nz ? xv<0, row, 5>(X) * kGravity : 0.0f;
                                                        } else {
                                                                Xv[row] = xrow<j, row>(X, v[j > 0 ? j - 1 : 0]);
                                                                Xa[row] = xrow<j, row>(X, a[j > 0 ? j - 1 : 0]);
                                                        }
                                                });
                                                ndv[0] = Xv[1], ndv[1] = -Xv[0], ndv[2] = 0.0f, ndv[3] = Xv[4], ndv[4] = -Xv[3], ndv[5] = 0.0f;
                                                src[0] = Xa[1], src[1] = -Xa[0], src[2] = 0.0f, src[3] = Xa[4], src[4] = -Xa[3], src[5] = 0.0f;
                                                if constexpr (j == 0) sfor<0, 6>([&](auto rc) { ndv[rc] = 0.0f; });
                                        } else {
                                                sfor<0, 6>([&](auto rc) { ndv[rc] = (rc == 2) ? 1.0f : 0.0f; });
                                                src[0] = v[j][1], src[1] = -v[j][0], src[2] = 0.0f, src[3] = v[j][4], src[4] = -v[j][3], src[5] = 0.0f;
                                        }
                                        nda[0] = ndv[1] * qd[j], nda[1] = (-ndv[0]) * qd[j], nda[2] = 0.0f, nda[3] = ndv[4] * qd[j], nda[4] = (-ndv[3]) * qd[j], nda[5] = 0.0f;
                                        sfor<0, 6>([&](auto rc) { nda[rc] = nda[rc] + src[rc]; });
                                } else {
                                        if constexpr (j > 0) {
                                                sfor<0, 6>([&](auto rc) { ndv[rc] = xrow<j, rc>(X, dv); });
                                                nda[0] = ndv[1] * qd[j], nda[1] = (-ndv[0]) * qd[j], nda[2] = 0.0f, nda[3] = ndv[4] * qd[j], nda[4] = (-ndv[3]) * qd[j], nda[5] = 0.0f;
                                                sfor<0, 6>([&](auto rc) { nda[rc] = nda[rc] + xrow<j, rc>(X, da); });
                                        } else {
                                                sfor<0, 6>([&](auto rc) { ndv[rc] = 0.0f, nda[rc] = 0.0f; });  // unreachable: j > k >= 0
                                        }
                                }
                                sfor<0, 6>([&](auto rc) {
                                        dv[rc] = ndv[rc];
                                        da[rc] = nda[rc];
                                });
                                float t0[6];
                                fx_times_v(t0, dv, Iv[j]);
                                sfor<0, 6>([&](auto rc) {
                                        constexpr int row = rc;
                                        float         d1 = irow<j, row>(da);
                                        float         d2 = 0.0f;
                                        sfor<0, 6>([&](auto tc) { d2 = fmaf(FxvI[j][row + 6 * tc], dv[tc], d2); });
                                        df[j][row] = t0[row] + (d1 + d2);
                                });
                        }
                });
                sfor_down<1, NQ>([&](auto jc) {
                        constexpr int j = jc;
                        float         upd[6];
                        sfor<0, 6>([&](auto rc) { upd[rc] = xcol<j, rc>(X, df[j]); });
                        if constexpr (W == 0) {
                                if (k == j) {
                                        float mxf[6] = {f[j][1], -f[j][0], 0.0f, f[j][4], -f[j][3], 0.0f};
                                        sfor<0, 6>([&](auto rc) { upd[rc] = upd[rc] + (-xcol<j, rc>(X, mxf)); });
                                }
                        }
                        sfor<0, 6>([&](auto rc) { df[j - 1][rc] = df[j - 1][rc] + upd[rc]; });
                });
                sfor<0, NQ>([&](auto jc) { dc[jc] = df[jc][2]; });
        }

        // TWO columns k0 and k0 + 1 at a time, as the two lanes of packed pairs (dc[j].x = d c_j / d{q|qd}_k0, dc[j].y = ... k0+1).  Every operation
        // of rnea_grad_col_rt is "a coefficient times a column's vector": the coefficient is shared and the two columns ride in one packed
        // instruction.  Column k0 + 1 starts one joint later; until then its lane carries exact zeros, which change nothing downstream (every
        // chain starts from +0).  With k0 + 1 == NQ the second lane stays zero and is ignored by the caller.
        template<int W>
        static GATO_HD void rnea_grad_col2_rt(int k0, const Xmat& X, const float* qd, const V6& v, const V6& a, const V6& f, const V6& Iv, const float (&FxvI)[NQ][36], f2 (&dc)[NQ])
        {
                f2 df[NQ][6];
                f2 dv[6], da[6];
                const f2 z2 = mk2(0.0f, 0.0f);
                sfor<0, NQ>([&](auto jc) { sfor<0, 6>([&](auto rc) { df[jc][rc] = z2; }); });
                sfor<0, 6>([&](auto rc) { dv[rc] = z2, da[rc] = z2; });
                sfor<0, NQ>([&](auto jc) {
                        constexpr int j = jc;
                        if (j >= k0) {
                                f2 ndv[6], nda[6];
                                sfor<0, 6>([&](auto rc) { ndv[rc] = z2, nda[rc] = z2; });
                                if constexpr (j > 0) {
                                        if (j > k0) {
                                                // propagate the columns that have started (lane y still carries zeros at j == k0 + 1: they stay zeros)
                                                sfor<0, 6>([&](auto rc) { ndv[rc] = xrow2<j, rc>(X, dv); });
                                                nda[0] = mul2s(qd[j], ndv[1]), nda[1] = mul2s(qd[j], neg2(ndv[0])), nda[2] = z2;
                                                nda[3] = mul2s(qd[j], ndv[4]), nda[4] = mul2s(qd[j], neg2(ndv[3])), nda[5] = z2;
                                                sfor<0, 6>([&](auto rc) { nda[rc] = add2(nda[rc], xrow2<j, rc>(X, da)); });
                                        }
                                }
                                if (j == k0 || j == k0 + 1) {
                                        // the column that starts at this joint (scalar): dv = mx2(X v_parent) | S ; da = mx2_scaled(dv, qd) + { mx2(X a_parent) | mx2(v) }
                                        float sdv[6], sda[6], src[6];
                                        if constexpr (W == 0) {
                                                float Xv[6], Xa[6];
                                                sfor<0, 6>([&](auto rc) {
                                                        constexpr int row = rc;
                                                        if constexpr (j == 0) {
                                                                Xv[row] = 0.0f;
                                                                Xa[row] = XE<P, 0, row, 5>::nz ? xv<0, row, 5>(X) * kGravity : 0.0f;
                                                        } else {
                                                                Xv[row] = xrow<j, row>(X, v[j > 0 ? j - 1 : 0]);
                                                                Xa[row] = xrow<j, row>(X, a[j > 0 ? j - 1 : 0]);
                                                        }
                                                });
                                                sdv[0] = Xv[1], sdv[1] = -Xv[0], sdv[2] = 0.0f, sdv[3] = Xv[4], sdv[4] = -Xv[3], sdv[5] = 0.0f;
                                                src[0] = Xa[1], src[1] = -Xa[0], src[2] = 0.0f, src[3] = Xa[4], src[4] = -Xa[3], src[5] = 0.0f;
                                                if constexpr (j == 0) sfor<0, 6>([&](auto rc) { sdv[rc] = 0.0f; });
                                        } else {
                                                sfor<0, 6>([&](auto rc) { sdv[rc] = (rc == 2) ? 1.0f : 0.0f; });
                                                src[0] = v[j][1], src[1] = -v[j][0], src[2] = 0.0f, src[3] = v[j][4], src[4] = -v[j][3], src[5] = 0.0f;
                                        }
                                        sda[0] = sdv[1] * qd[j], sda[1] = (-sdv[0]) * qd[j], sda[2] = 0.0f, sda[3] = sdv[4] * qd[j], sda[4] = (-sdv[3]) * qd[j], sda[5] = 0.0f;
                                        sfor<0, 6>([&](auto rc) { sda[rc] = sda[rc] + src[rc]; });
                                        if (j == k0)
                                                sfor<0, 6>([&](auto rc) { ndv[rc].x = sdv[rc], nda[rc].x = sda[rc]; });
                                        else
                                                sfor<0, 6>([&](auto rc) { ndv[rc].y = sdv[rc], nda[rc].y = sda[rc]; });
                                }
                                sfor<0, 6>([&](auto rc) {
                                        dv[rc] = ndv[rc];
                                        da[rc] = nda[rc];
                                });
                                // df_j = fx(dv) I v  +  ( I da + (fx(v) I) dv )
                                f2 t0[6];
                                fx2_times_v(t0, dv, Iv[j]);
                                sfor<0, 6>([&](auto rc) {
                                        constexpr int row = rc;
                                        const f2      d1 = irow2<j, row>(da);
                                        f2            d2 = z2;
                                        sfor<0, 6>([&](auto tc) { d2 = fma2s(FxvI[j][row + 6 * tc], dv[tc], d2); });
                                        df[j][row] = add2(t0[row], add2(d1, d2));
                                });
                        }
                });
                sfor_down<1, NQ>([&](auto jc) {
                        constexpr int j = jc;
                        f2            upd[6];
                        sfor<0, 6>([&](auto rc) { upd[rc] = xcol2<j, rc>(X, df[j]); });
                        if constexpr (W == 0) {
                                if (j == k0 || j == k0 + 1) {
                                        float mxf[6] = {f[j][1], -f[j][0], 0.0f, f[j][4], -f[j][3], 0.0f};
                                        if (j == k0)
                                                sfor<0, 6>([&](auto rc) { upd[rc].x = upd[rc].x + (-xcol<j, rc>(X, mxf)); });
                                        else
                                                sfor<0, 6>([&](auto rc) { upd[rc].y = upd[rc].y + (-xcol<j, rc>(X, mxf)); });
                                }
                        }
                        sfor<0, 6>([&](auto rc) { df[j - 1][rc] = add2(df[j - 1][rc], upd[rc]); });
                });
                sfor<0, NQ>([&](auto jc) { dc[jc] = df[jc][2]; });
        }

        // forwardDynamicsAndGradient with wrench (iiwa14_plant.cuh:229-268), split so that the column blocks of the gradient can be
        // computed by different threads: dyn_prologue (X, M^-1, RNEA, qdd, RNEA with qdd, I v, fx(v) I) + grad_block<W>.
        struct DynState {
                Xmat  X;
                float Minv[NQ * NQ];
                V6    v, a, f, Iv;
                float FxvI[NQ][36];
                float qdd[NQ];
        };
        static GATO_HD void dyn_prologue(const float* q, const float* qd, const float* u, const float* fext, DynState& st)
        {
                float t[2 * NQ];
                sincos(q, t);
                update_X(t, st.X);
                minv(st.X, st.Minv);
                rnea<false>(st.X, qd, nullptr, fext, st.v, st.a, st.f);
                fd_finish(st.Minv, u, st.f, st.qdd);
                rnea<true>(st.X, qd, st.qdd, fext, st.v, st.a, st.f);
                sfor<0, NQ>([&](auto jc) {
                        constexpr int j = jc;
                        sfor<0, 6>([&](auto rc) { st.Iv[j][rc] = irow<j, rc>(st.v[j]); });
                        sfor<0, 6>([&](auto cc) {
                                constexpr int c = cc;
                                float         col[6], out[6];
                                sfor<0, 6>([&](auto tc) { col[tc] = inertia<P, j, tc, c>(); });
                                fx_times_v(out, st.v[j], col);
                                sfor<0, 6>([&](auto rc) { st.FxvI[j][6 * c + rc] = out[rc]; });
                        });
                });
        }
        // W = 0: d qdd / dq, W = 1: d qdd / dqd;  out[k*NQ + row] = -(Minv * dc_k)[row]  (symmetric-upper lookup, plant:249-259)
        template<int W>
        static GATO_HD void grad_block(const DynState& st, const float* qd, float (&out)[NQ * NQ])
        {
                sfor<0, NQ>([&](auto kc) {
                        constexpr int k = kc;
                        float         dc[NQ];
                        rnea_grad_col<W, k>(st.X, qd, st.v, st.a, st.f, st.Iv, st.FxvI, dc);
                        sfor<0, NQ>([&](auto rc) {
                                constexpr int row = rc;
                                float         val = 0.0f;
                                sfor<0, NQ>([&](auto cc) { val = fmaf(minv_sym<row, cc>(st.Minv), dc[cc], val); });
                                out[k * NQ + row] = -val;
                        });
                });
        }
        // dqdd col-major NQ x 3NQ = [ dqdd/dq | dqdd/dqd | Minv ]
        static GATO_HD void fd_and_grad(const float* q, const float* qd, const float* u, const float* fext, float (&qdd)[NQ], float (&dqdd)[3 * NQ * NQ])
        {
                DynState st;
                dyn_prologue(q, qd, u, fext, st);
                sfor<0, NQ>([&](auto ic) { qdd[ic] = st.qdd[ic]; });
                float blk[NQ * NQ];
                grad_block<0>(st, qd, blk);
                sfor<0, NQ * NQ>([&](auto ec) { dqdd[ec] = blk[ec]; });
                grad_block<1>(st, qd, blk);
                sfor<0, NQ * NQ>([&](auto ec) { dqdd[NQ * NQ + ec] = blk[ec]; });
                sfor<0, NQ * NQ>([&](auto ec) { dqdd[2 * NQ * NQ + ec] = minv_sym<ec % NQ, ec / NQ>(st.Minv); });
        }

        // ---- end-effector position and positional Jacobian ---------------------------------------------
        template<int J, int R, int C, bool D>
        static GATO_HD float hv(const float* t)
        {
                using E = HE<P, J, R, C, D>;
                if constexpr (E::ti >= 0)
                        return trig_entry<E>(t);
                else
                        return E::cst;
        }
        template<int J, bool D>
        static GATO_HD void hom_apply(const float* t, float (&p)[4])
        {
                float o[4];
                sfor<0, 4>([&](auto rc) {
                        constexpr int row = rc;
                        float         r = 0.0f;
                        sfor<0, 4>([&](auto ic) {
                                constexpr int i = ic;
                                if constexpr (HE<P, J, row, i, D>::nz) r = fmaf(hv<J, row, i, D>(t), p[i], r);
                        });
                        o[row] = r;
                });
                sfor<0, 4>([&](auto rc) { p[rc] = o[rc]; });
        }
        template<int D>  // D = -1: position; D >= 0: derivative w.r.t. joint D
        static GATO_HD void ee_chain(const float* t, float (&out)[3])
        {
                float p[4];
                sfor<0, 4>([&](auto rc) {
                        constexpr int row = rc;
                        if constexpr (D == NQ - 1)
                                p[row] = HE<P, NQ - 1, row, 3, true>::nz ? hv<NQ - 1, row, 3, true>(t) : 0.0f;
                        else
                                p[row] = HE<P, NQ - 1, row, 3, false>::nz ? hv<NQ - 1, row, 3, false>(t) : 0.0f;
                });
                sfor_down<0, NQ - 1>([&](auto jc) {
                        constexpr int j = jc;
                        if constexpr (j == D)
                                hom_apply<j, true>(t, p);
                        else
                                hom_apply<j, false>(t, p);
                });
                out[0] = p[0], out[1] = p[1], out[2] = p[2];
        }
        static GATO_HD void ee_pos(const float* q, float (&ee)[3])
        {
                float t[2 * NQ];
                sincos(q, t);
                ee_chain<-1>(t, ee);
        }
        static GATO_HD void ee_pos_grad(const float* q, float (&ee)[3], float (&J)[NQ][3])
        {
                float t[2 * NQ];
                sincos(q, t);
                ee_chain<-1>(t, ee);
                sfor<0, NQ>([&](auto dc) { ee_chain<dc>(t, J[dc]); });
        }

        // ---- trapezoidal integrator  (integrator.cuh:34-37, 143-184) ---------------------------------------
        static GATO_HD void integrate(const float* q, const float* qd, const float (&qdd)[NQ], float dt, float (&qn)[NQ], float (&qdn)[NQ])
        {
                sfor<0, NQ>([&](auto ic) {
                        constexpr int i = ic;
                        qdn[i] = fmaf(dt, qdd[i], qd[i]);
                        float  lin = fmaf(dt, qd[i], q[i]);
                        double acc = ((double)qdd[i] * 0.5) * (double)dt;
                        qn[i] = (float)fma(acc, (double)dt, (double)lin);
                });
        }

        // ---- barriers  (iiwa14_plant.cuh:103-155, indy7_plant.cuh:133-147) --------------------------------
        static GATO_HD float joint_barrier(float q, float lo, float hi)
        {
                float dmin = q - lo, dmax = hi - q;
                dmin = ((double)dmin <= 1e-10) ? (float)1e-10 : dmin;
                dmax = ((double)dmax <= 1e-10) ? (float)1e-10 : dmax;
                return (-g_log(dmin)) - g_log(dmax);
        }
        static GATO_HD float joint_barrier_grad(float q, float lo, float hi)
        {
                float dmin = q - lo, dmax = hi - q;
                if constexpr (P::ID == 1) {
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
        static GATO_HD float joint_barrier_hess(float q, float lo, float hi)
        {
                float       dmin = q - lo, dmax = hi - q;
                const float eps = 1e-6f;
                float       amin = dmin >= 0.0f ? dmin : -dmin, amax = dmax >= 0.0f ? dmax : -dmax;
                if (amin < eps) amin = eps;
                if (amax < eps) amax = eps;
                return 1.0f / (amin * amin) + 1.0f / (amax * amax);
        }
};

// block::reduce's halving tree with odd carry (linalg.cuh:329-353), evaluated by one thread on registers
template<int N>
GATO_HD float tree_reduce(float (&x)[N])
{
        if constexpr (N > 3) {
                constexpr int odd = N % 2, half = (N - odd) / 2;
                float         y[half];
                sfor<0, half>([&](auto ic) { y[ic] = x[ic] + x[ic + half]; });
                if constexpr (odd) y[0] = y[0] + x[2 * half];
                return tree_reduce<half>(y);
        } else {
                float r = x[0];
                sfor<1, N>([&](auto ic) { r = r + x[ic]; });
                return r;
        }
}

}  // namespace gato
