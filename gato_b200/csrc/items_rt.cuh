// Items<P> for the run-time models (P = RtPlant<NQ>): the same per-work-item stages as items.cuh -- tracking-cost gradient / Hessian,
// linearised dynamics, merit contribution -- on the table-driven dynamics of rbd_rt.cuh, with limits and barrier style taken from the model.
// Same arithmetic as items.cuh statement by statement (both restate the oracle's cost_grad_hess / kkt_one / merit_one); the object holds the
// model reference, the member functions have the signatures of the static ones of the compiled plants so the kernels are written once.
#pragma once
#include "items.cuh"
#include "rbd_rt.cuh"

namespace gato {

template<int NQ_>
struct Items<RtPlant<NQ_>> {
        static constexpr int NQ = NQ_, NX = 2 * NQ, NU = NQ;
        using R = RbdRt<NQ, float>;
        using DynState = typename R::DynState;
        const RtModel& m;
        GATO_HD explicit Items(const RtModel& model) : m(model) {}

        static constexpr int kPosFormBMinKnots = 4, kPosFormBMaxKnots = 9;
        static GATO_HD bool pos_form_b_for(int knot_points) { return knot_points >= kPosFormBMinKnots && knot_points <= kPosFormBMaxKnots; }

        template<bool WITH_R, class FQ, class Fq, class FR, class Fr>
        GATO_HD void cost_grad_hess(const float* xu, const float* ref3, const Costs& cs, FQ&& putQ, Fq&& putq, FR&& putR, Fr&& putr, bool terminal = !WITH_R,
                                    bool pos_form_b = !WITH_R) const
        {
                // one stack object for every dynamically indexed array of this function (see the note on local arrays in rbd_rt.cuh)
                struct Work {
                        typename R::EeWork ew;
                        float              h[NQ], bq[NQ], bv[NQ], bu[NQ], dq[NQ], dv[NQ], du[NQ];
                } wk;
                float ee[3], e[3];
                R::ee_pos_grad(m, xu, ee, wk.ew);
                float(&J)[NQ][3] = wk.ew.J;
                float(&h)[NQ] = wk.h;
                sfor<0, 3>([&](auto rc) { e[rc] = ee[rc] - ref3[rc]; });
                const float w = cs.q_cost;
                sfor<0, NQ>([&](auto ic) {
                        constexpr int i = ic;
                        float         s = J[i][1] * e[1];
                        s = fmaf(J[i][0], e[0], s);
                        h[i] = fmaf(J[i][2], e[2], s);
                });
                const int style = m.style;
                float(&bq)[NQ] = wk.bq;  // barrier gradients
                float(&bv)[NQ] = wk.bv;
                float(&bu)[NQ] = wk.bu;
                GATO_ROLLED
                for (int i = 0; i < NQ; i++) {
                        bq[i] = rt_joint_barrier_grad(style, xu[i], m.jl[i][0], m.jl[i][1]);
                        bv[i] = rt_joint_barrier_grad(style, xu[NQ + i], m.vl[i][0], m.vl[i][1]);
                        putq(i, pos_form_b ? fmaf(cs.q_lim_cost, bq[i], h[i] * w) : fmaf(h[i], w, cs.q_lim_cost * bq[i]));
                        putq(NQ + i, terminal ? fmaf(cs.vel_lim_cost, bv[i], cs.qd_cost * xu[NQ + i]) : fmaf(cs.qd_cost, xu[NQ + i], cs.vel_lim_cost * bv[i]));
                        if constexpr (WITH_R) {
                                bu[i] = rt_joint_barrier_grad(style, xu[NX + i], m.cl[i][0], m.cl[i][1]);
                                putr(i, fmaf(cs.u_cost, xu[NX + i], cs.ctrl_lim_cost * bu[i]));
                        }
                }
                // diagonal barrier terms first (one evaluation per joint), then the blocks
                float(&dq)[NQ] = wk.dq;
                float(&dv)[NQ] = wk.dv;
                float(&du)[NQ] = wk.du;
                GATO_ROLLED
                for (int i = 0; i < NQ; i++) {
                        if (style == 1) {
                                dq[i] = rt_joint_barrier_hess(xu[i], m.jl[i][0], m.jl[i][1]);
                                dv[i] = fmaf(cs.vel_lim_cost, rt_joint_barrier_hess(xu[NQ + i], m.vl[i][0], m.vl[i][1]), cs.qd_cost);
                                if constexpr (WITH_R) du[i] = fmaf(cs.ctrl_lim_cost, rt_joint_barrier_hess(xu[NX + i], m.cl[i][0], m.cl[i][1]), cs.u_cost);
                        } else {
                                dq[i] = 0.0f;
                                dv[i] = fmaf(cs.vel_lim_cost * bv[i], bv[i], cs.qd_cost);
                                if constexpr (WITH_R) du[i] = fmaf(cs.ctrl_lim_cost * bu[i], bu[i], cs.u_cost);
                        }
                }
                sfor<0, NX>([&](auto ic) {
                        constexpr int i = ic;
                        sfor<0, NX>([&](auto jc) {
                                constexpr int j = jc;
                                float         val;
                                if constexpr (j < NQ && i < NQ) {
                                        val = (h[i] * h[j]) * w;
                                        if (style == 1) {
                                                if constexpr (i == j) val = fmaf(cs.q_lim_cost, dq[i], val);
                                        } else {
                                                val = fmaf(cs.q_lim_cost * bq[i], bq[j], val);
                                        }
                                } else if constexpr (i == j) {
                                        val = dv[i - NQ];
                                } else {
                                        val = 0.0f;
                                }
                                putQ(i * NX + j, val);
                        });
                });
                if constexpr (WITH_R) {
                        sfor<0, NU>([&](auto oc) {
                                constexpr int o = oc;
                                sfor<0, NU>([&](auto jc) {
                                        constexpr int j = jc;
                                        putR(o * NU + j, o == j ? du[o] : 0.0f);
                                });
                        });
                }
        }

        // ---- linearised dynamics ----------------------------------------------------------------------------------
        GATO_HD void prologue(const float* xux, const float* fext, DynState& st) const { R::dyn_prologue(m, xux, xux + NQ, xux + NX, fext, st); }
        // column c = k + W*NQ of A
        template<int W, class FA>
        GATO_HD void column(int k, DynState& st, const float* qd, float dt, FA&& putA) const
        {
                R::template rnea_grad_col<W>(m, k, st, qd);
                float(&dc)[NQ] = st.cw.dc;
                float(&d)[NQ] = st.cw.d;
                GATO_ROLLED
                for (int row = 0; row < NQ; row++) {
                        float val = 0.0f;
                        sfor<0, NQ>([&](auto cc) { val = fmaf(R::minv_sym(st.Minv, row, cc), dc[cc], val); });
                        d[row] = -val;
                }
                const float dt_sq_half = (float)((0.5 * (double)dt) * (double)dt);
                const int   c = k + W * NQ;
                GATO_ROLLED
                for (int r = 0; r < NX; r++) {
                        const int rd = r < NQ ? r : r - NQ;
                        float     val = (r == c) ? 1.0f : 0.0f;
                        if (r < NQ) {
                                if (W == 1 && r == k) val = val + dt;
                                val = fmaf(dt_sq_half, d[rd], val);
                        } else {
                                val = fmaf(dt, d[rd], val);
                        }
                        putA(c * NX + r, val);
                }
        }
        // the defect c_{k+1} and one column of B, for callers that stage one column at a time
        template<class Fc>
        GATO_HD void defect(const DynState& st, const float* xux, float dt, Fc&& putc) const
        {
                float qn[NQ], qdn[NQ];
                R::integrate(xux, xux + NQ, st.qdd, dt, qn, qdn);
                sfor<0, NQ>([&](auto ic) {
                        constexpr int i = ic;
                        putc(i, xux[NX + NU + i] - qn[i]);
                        putc(i + NQ, xux[NX + NU + NQ + i] - qdn[i]);
                });
        }
        template<class FB>
        GATO_HD void b_column(const DynState& st, int c, float dt, FB&& putB) const
        {
                const float dt_sq_half = (float)((0.5 * (double)dt) * (double)dt);
                GATO_ROLLED
                for (int r = 0; r < NX; r++) {
                        const float d = R::minv_sym(st.Minv, r < NQ ? r : r - NQ, c);
                        putB(c * NX + r, (r < NQ) ? (dt_sq_half * d) : (dt * d));
                }
        }
        template<class FA>
        GATO_HD void linearize_column_any(int col, DynState& st, const float* qd, float dt, FA&& putA) const
        {
                if (col < NQ)
                        column<0>(col, st, qd, dt, putA);
                else
                        column<1>(col - NQ, st, qd, dt, putA);
        }
        template<class FB, class Fc>
        GATO_HD void linearize_base(const DynState& st, const float* xux, float dt, FB&& putB, Fc&& putc) const
        {
                const float dt_sq_half = (float)((0.5 * (double)dt) * (double)dt);
                float       qn[NQ], qdn[NQ];
                R::integrate(xux, xux + NQ, st.qdd, dt, qn, qdn);
                sfor<0, NQ>([&](auto ic) {
                        constexpr int i = ic;
                        putc(i, xux[NX + NU + i] - qn[i]);
                        putc(i + NQ, xux[NX + NU + NQ + i] - qdn[i]);
                });
                GATO_ROLLED
                for (int c = 0; c < NU; c++) {
                        GATO_ROLLED
                        for (int r = 0; r < NX; r++) {
                                const float d = R::minv_sym(st.Minv, r < NQ ? r : r - NQ, c);
                                putB(c * NX + r, (r < NQ) ? (dt_sq_half * d) : (dt * d));
                        }
                }
        }
        //   HALF 0: columns 0..NQ-1 of A (d/dq) and the defect c          HALF 1: columns NQ..NX-1 of A (d/dqd) and B
        template<int HALF, class FA, class FB, class Fc>
        GATO_HD void linearize_half_rolled(const float* xux, const float* fext, float dt, FA&& putA, FB&& putB, Fc&& putc) const
        {
                DynState st;
                prologue(xux, fext, st);
                GATO_ROLLED
                for (int k = 0; k < NQ; k++) column<HALF>(k, st, xux + NQ, dt, putA);
                if constexpr (HALF == 0) {
                        float qn[NQ], qdn[NQ];
                        R::integrate(xux, xux + NQ, st.qdd, dt, qn, qdn);
                        sfor<0, NQ>([&](auto ic) {
                                constexpr int i = ic;
                                putc(i, xux[NX + NU + i] - qn[i]);
                                putc(i + NQ, xux[NX + NU + NQ + i] - qdn[i]);
                        });
                } else {
                        const float dt_sq_half = (float)((0.5 * (double)dt) * (double)dt);
                        GATO_ROLLED
                        for (int c = 0; c < NU; c++) {
                                GATO_ROLLED
                                for (int r = 0; r < NX; r++) {
                                        const float d = R::minv_sym(st.Minv, r < NQ ? r : r - NQ, c);
                                        putB(c * NX + r, (r < NQ) ? (dt_sq_half * d) : (dt * d));
                                }
                        }
                }
        }

        // ---- merit contribution of knot k ------------------------------------------------------------------
        template<bool LAST>
        GATO_HD float tracking_cost(const float* xu, const float* ref3, const Costs& cs) const
        {
                constexpr int TN = NQ + (LAST ? 0 : NU);
                float         cv[TN + 3], ee[3];  // cv: the only dynamically indexed array of this function
                R::ee_pos(m, xu, ee);
                GATO_ROLLED
                for (int i = 0; i < NQ; i++) {
                        const float err = xu[i + NQ];
                        float       c = ((0.5f * cs.qd_cost) * err) * err;
                        // zero-weight barriers are not evaluated (see items.cuh: tracking_cost)
                        if (cs.q_lim_cost != 0.0f) c = fmaf(cs.q_lim_cost, rt_joint_barrier(xu[i], m.jl[i][0], m.jl[i][1]), c);
                        if (cs.vel_lim_cost != 0.0f) c = fmaf(cs.vel_lim_cost, rt_joint_barrier(xu[i + NQ], m.vl[i][0], m.vl[i][1]), c);
                        cv[i] = c;
                }
                if constexpr (!LAST) {
                        GATO_ROLLED
                        for (int j = 0; j < NU; j++) {
                                const float err = xu[NX + j];
                                float       c = ((0.5f * cs.u_cost) * err) * err;
                                if (cs.ctrl_lim_cost != 0.0f) c = fmaf(cs.ctrl_lim_cost, rt_joint_barrier(xu[NX + j], m.cl[j][0], m.cl[j][1]), c);
                                cv[NQ + j] = c;
                        }
                }
                const float w = LAST ? cs.N_cost : cs.q_cost;
                sfor<0, 3>([&](auto ic) {
                        constexpr int i = ic;
                        const float   err = ee[i] - ref3[i];
                        cv[TN + i] = (float)((((double)w * 0.5) * (double)err) * (double)err);
                });
                return tree_reduce<TN + 3>(cv);
        }
        GATO_HD float merit_mid_cons(const float* xux, const float* fext, float dt) const
        {
                float qdd[NQ], qn[NQ], qdn[NQ], err[NX];
                R::forward_dynamics(m, xux, xux + NQ, xux + NX, fext, qdd);
                R::integrate(xux, xux + NQ, qdd, dt, qn, qdn);
                sfor<0, NQ>([&](auto ic) {
                        constexpr int i = ic;
                        err[i] = fabsf(xux[NX + NU + i] - qn[i]);
                        err[i + NQ] = fabsf(xux[NX + NU + NQ + i] - qdn[i]);
                });
                return tree_reduce<NX>(err);
        }
        GATO_HD float merit_mid(const float* xux, const float* ref3, float mu, const float* fext, float dt, const Costs& cs) const
        {
                const float cost = tracking_cost<false>(xux, ref3, cs);
                return fmaf(mu, merit_mid_cons(xux, fext, dt), cost);
        }
        static GATO_HD float merit_last_cons(const float* x0err)
        {
                float err[NX];
                sfor<0, NX>([&](auto ic) { err[ic] = x0err[ic]; });
                return tree_reduce<NX>(err);
        }
        GATO_HD float merit_last(const float* x, const float* ref3, float mu, const float* x0err, const Costs& cs) const
        {
                const float cost = tracking_cost<true>(x, ref3, cs);
                return fmaf(mu, merit_last_cons(x0err), cost);
        }
        // ---- what the thread-per-solve kernels use ---------------------------------------------------------------
        GATO_HD void sim_step(const float* x, const float* u, const float* fext, float dt, float (&qn)[NQ], float (&qdn)[NQ]) const
        {
                float qdd[NQ];
                R::forward_dynamics(m, x, x + NQ, u, fext, qdd);
                R::integrate(x, x + NQ, qdd, dt, qn, qdn);
        }
        GATO_HD void ee_pos(const float* q, float (&ee)[3]) const { R::ee_pos(m, q, ee); }
};

}  // namespace gato
