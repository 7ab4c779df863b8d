// Per-work-item stages of the BSQP path built on rbd.cuh: what ONE thread computes for one knot.
//   kkt_item   : linearised dynamics (A_k, B_k, defect c_{k+1}) + quadraticised tracking cost (Q,q,R,r)
//                replaces setupKKTSystemBatchedKernel's per-block body (setup_kkt.cuh:52-107),
//                compute_linearized_dynamics (integrator.cuh:235-257) and
//                trackingCostGradientAndHessian[_lastblock] (iiwa14_plant.cuh:338-450, indy7_plant.cuh:325-440)
//   merit_item : cost_k + mu * |defect_k|_1 at z + alpha dz
//                replaces computeMeritBatchedKernel's per-block body (merit.cuh:36-86), trackingcost
//                (iiwa14_plant.cuh:275-327) and compute_integrator_error (integrator.cuh:211-233)
// Outputs are handed to caller-supplied sinks `put(index, value)` so the same code feeds the CUDA kernels
// (shared-memory transpose staging) and the host unit test.
#pragma once
#include "rbd.cuh"

namespace gato {

template<class P>
struct Items {
        using R = Rbd<P>;
        static constexpr int NQ = P::NQ, NX = 2 * NQ, NU = NQ;
        using DynState = typename R::DynState;
        // The kernels are written once for the compiled plants and the run-time models (items_rt.cuh), whose Items object carries the model
        // reference: they construct an Items<P> and call members.  The compiled plants carry their constants in the code: nothing to hold.
        GATO_HD Items() {}
        static GATO_HD void prologue(const float* xux, const float* fext, DynState& st) { R::dyn_prologue(xux, xux + NQ, xux + NX, fext, st); }
        static GATO_HD void sim_step(const float* x, const float* u, const float* fext, float dt, float (&qn)[NQ], float (&qdn)[NQ])
        {
                float qdd[NQ];
                R::forward_dynamics(x, x + NQ, u, fext, qdd);
                R::integrate(x, x + NQ, qdd, dt, qn, qdn);
        }
        static GATO_HD void ee_pos(const float* q, float (&ee)[3]) { R::ee_pos(q, ee); }

        // ---- tracking cost gradient / Hessian at (x,u) against ref xyz, weight q_cost (see oracle note) ----
        // terminal: the block is Q_{N-1}, q_{N-1} (the reference's computeR = false instantiation).  pos_form_b: the position entries of the
        // gradient use fma(lim, barrier', round(a * b)) instead of fma(a, b, round(lim * barrier')) -- what nvcc emits for the terminal block
        // and, when the reference is compiled for a horizon of 4 ... 9 knots, for every block (see the oracle's cost_grad_hess).
        static constexpr int kPosFormBMinKnots = 4, kPosFormBMaxKnots = 9;
        static GATO_HD bool pos_form_b_for(int knot_points) { return knot_points >= kPosFormBMinKnots && knot_points <= kPosFormBMaxKnots; }
        template<bool WITH_R, class FQ, class Fq, class FR, class Fr>
        static GATO_HD void cost_grad_hess(const float* xu, const float* ref3, const Costs& cs, FQ&& putQ, Fq&& putq, FR&& putR, Fr&& putr, bool terminal = !WITH_R,
                                           bool pos_form_b = !WITH_R)
        {
                float ee[3], J[NQ][3], e[3], h[NQ];
                R::ee_pos_grad(xu, ee, J);
                sfor<0, 3>([&](auto rc) { e[rc] = ee[rc] - ref3[rc]; });
                const float w = cs.q_cost;
                sfor<0, NQ>([&](auto ic) {
                        constexpr int i = ic;
                        float         s = J[i][1] * e[1];
                        s = fmaf(J[i][0], e[0], s);
                        h[i] = fmaf(J[i][2], e[2], s);
                });
                float bq[NQ], bv[NQ], bu[NQ];  // barrier gradients
                sfor<0, NQ>([&](auto ic) {
                        constexpr int i = ic;
                        bq[i] = R::joint_barrier_grad(xu[i], limit<P, 0, i, 0>(), limit<P, 0, i, 1>());
                        bv[i] = R::joint_barrier_grad(xu[NQ + i], limit<P, 1, i, 0>(), limit<P, 1, i, 1>());
                        putq(i, pos_form_b ? fmaf(cs.q_lim_cost, bq[i], h[i] * w) : fmaf(h[i], w, cs.q_lim_cost * bq[i]));
                        putq(NQ + i, terminal ? fmaf(cs.vel_lim_cost, bv[i], cs.qd_cost * xu[NQ + i]) : fmaf(cs.qd_cost, xu[NQ + i], cs.vel_lim_cost * bv[i]));
                        if constexpr (WITH_R) {
                                bu[i] = R::joint_barrier_grad(xu[NX + i], limit<P, 2, i, 0>(), limit<P, 2, i, 1>());
                                putr(i, fmaf(cs.u_cost, xu[NX + i], cs.ctrl_lim_cost * bu[i]));
                        }
                });
                sfor<0, NX>([&](auto ic) {
                        constexpr int i = ic;
                        sfor<0, NX>([&](auto jc) {
                                constexpr int j = jc;
                                float         val;
                                if constexpr (j < NQ && i < NQ) {
                                        val = (h[i] * h[j]) * w;
                                        if constexpr (P::ID == 1) {
                                                if constexpr (i == j) val = fmaf(cs.q_lim_cost, R::joint_barrier_hess(xu[i], limit<P, 0, i, 0>(), limit<P, 0, i, 1>()), val);
                                        } else {
                                                val = fmaf(cs.q_lim_cost * bq[i], bq[j], val);
                                        }
                                } else if constexpr (i == j) {
                                        if constexpr (P::ID == 1)
                                                val = fmaf(cs.vel_lim_cost, R::joint_barrier_hess(xu[i], limit<P, 1, i - NQ, 0>(), limit<P, 1, i - NQ, 1>()), cs.qd_cost);
                                        else
                                                val = fmaf(cs.vel_lim_cost * bv[i - NQ], bv[i - NQ], cs.qd_cost);
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
                                        float         val = 0.0f;
                                        if constexpr (o == j) {
                                                if constexpr (P::ID == 1)
                                                        val = fmaf(cs.ctrl_lim_cost, R::joint_barrier_hess(xu[NX + o], limit<P, 2, o, 0>(), limit<P, 2, o, 1>()), cs.u_cost);
                                                else
                                                        val = fmaf(cs.ctrl_lim_cost * bu[o], bu[o], cs.u_cost);
                                        }
                                        putR(o * NU + j, val);
                                });
                        });
                }
        }

        // ---- linearised dynamics: A (NX x NX col-major), B (NX x NU col-major), defect c (NX) -----------------
        template<class FA, class FB, class Fc>
        static GATO_HD void linearize(const float* xux, const float* fext, float dt, FA&& putA, FB&& putB, Fc&& putc)
        {
                float qdd[NQ], dqdd[3 * NQ * NQ], qn[NQ], qdn[NQ];
                R::fd_and_grad(xux, xux + NQ, xux + NX, fext, qdd, dqdd);
                R::integrate(xux, xux + NQ, qdd, dt, qn, qdn);
                sfor<0, NQ>([&](auto ic) {
                        constexpr int i = ic;
                        putc(i, xux[NX + NU + i] - qn[i]);
                        putc(i + NQ, xux[NX + NU + NQ + i] - qdn[i]);
                });
                const float dt_sq_half = (float)((0.5 * (double)dt) * (double)dt);
                sfor<0, NX * NX>([&](auto ec) {
                        constexpr int i = ec, c = i / NX, r = i % NX, rd = r % NQ;
                        const float   d = dqdd[c * NQ + rd];
                        float         val = (r == c) ? 1.0f : 0.0f;
                        if constexpr (r < NQ) {
                                if constexpr (c >= NQ && r == c - NQ) val = val + dt;
                                val = fmaf(dt_sq_half, d, val);
                        } else {
                                val = fmaf(dt, d, val);
                        }
                        putA(i, val);
                });
                sfor<0, NX * NU>([&](auto ec) {
                        constexpr int i = ec, c = i / NX, r = i % NX, rd = r % NQ;
                        const float   d = dqdd[NX * NQ + c * NQ + rd];
                        putB(i, (r < NQ) ? (dt_sq_half * d) : (dt * d));
                });
        }

        // ---- the same linearisation split in two halves that different threads compute (each repeats the prologue) ----------
        //   HALF 0: columns 0..NQ-1 of A (d/dq) and the defect c          HALF 1: columns NQ..NX-1 of A (d/dqd) and B
        // putA receives indices of the full col-major A; results are bit-identical to linearize().
        template<int HALF, class FA, class FB, class Fc>
        static GATO_HD void linearize_half(const float* xux, const float* fext, float dt, FA&& putA, FB&& putB, Fc&& putc)
        {
                typename R::DynState st;
                R::dyn_prologue(xux, xux + NQ, xux + NX, fext, st);
                const float dt_sq_half = (float)((0.5 * (double)dt) * (double)dt);
                float       blk[NQ * NQ];
                R::template grad_block<HALF>(st, xux + NQ, blk);
                sfor<0, NX * NQ>([&](auto ec) {
                        constexpr int e = ec, cl = e / NX, r = e % NX, rd = r % NQ, c = cl + HALF * NQ;
                        const float   d = blk[cl * NQ + rd];
                        float         val = (r == c) ? 1.0f : 0.0f;
                        if constexpr (r < NQ) {
                                if constexpr (c >= NQ && r == c - NQ) val = val + dt;
                                val = fmaf(dt_sq_half, d, val);
                        } else {
                                val = fmaf(dt, d, val);
                        }
                        putA(c * NX + r, val);
                });
                if constexpr (HALF == 0) {
                        float qn[NQ], qdn[NQ];
                        R::integrate(xux, xux + NQ, st.qdd, dt, qn, qdn);
                        sfor<0, NQ>([&](auto ic) {
                                constexpr int i = ic;
                                putc(i, xux[NX + NU + i] - qn[i]);
                                putc(i + NQ, xux[NX + NU + NQ + i] - qdn[i]);
                        });
                } else {
                        sfor<0, NX * NU>([&](auto ec) {
                                constexpr int i = ec, c = i / NX, r = i % NX, rd = r % NQ;
                                const float   d = R::template minv_sym<rd, c>(st.Minv);
                                putB(i, (r < NQ) ? (dt_sq_half * d) : (dt * d));
                        });
                }
        }

        // ---- linearize_half with the loop over the columns rolled (rnea_grad_col_rt): what k_kkt runs -----------------------------------
        template<int HALF, class FA, class FB, class Fc>
        static GATO_HD void linearize_half_rolled(const float* xux, const float* fext, float dt, FA&& putA, FB&& putB, Fc&& putc)
        {
                typename R::DynState st;
                R::dyn_prologue(xux, xux + NQ, xux + NX, fext, st);
                const float dt_sq_half = (float)((0.5 * (double)dt) * (double)dt);
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
                for (int k0 = 0; k0 < NQ; k0 += 2) {
                        // two columns per pass as the lanes of packed pairs (rnea_grad_col2_rt); with NQ odd the last pass has one
                        f2 dc[NQ], d[NQ];
                        R::template rnea_grad_col2_rt<HALF>(k0, st.X, xux + NQ, st.v, st.a, st.f, st.Iv, st.FxvI, dc);
                        sfor<0, NQ>([&](auto rc) {
                                constexpr int row = rc;
                                f2            val = mk2(0.0f, 0.0f);
                                sfor<0, NQ>([&](auto cc) { val = fma2s(R::template minv_sym<row, cc>(st.Minv), dc[cc], val); });
                                d[row] = neg2(val);
                        });
                        for (int l = 0; l < 2; l++) {
                                const int k = k0 + l;
                                if (k >= NQ) break;
                                const int c = k + HALF * NQ;
                                sfor<0, NX>([&](auto rc) {
                                        constexpr int r = rc, rd = r % NQ;
                                        const float   dd = l == 0 ? d[rd].x : d[rd].y;
                                        float         val = (r == c) ? 1.0f : 0.0f;
                                        if constexpr (r < NQ) {
                                                if (HALF == 1 && r == k) val = val + dt;  // c >= NQ && r == c - NQ
                                                val = fmaf(dt_sq_half, dd, val);
                                        } else {
                                                val = fmaf(dt, dd, val);
                                        }
                                        putA(c * NX + r, val);
                                });
                        }
                }
                if constexpr (HALF == 0) {
                        float qn[NQ], qdn[NQ];
                        R::integrate(xux, xux + NQ, st.qdd, dt, qn, qdn);
                        sfor<0, NQ>([&](auto ic) {
                                constexpr int i = ic;
                                putc(i, xux[NX + NU + i] - qn[i]);
                                putc(i + NQ, xux[NX + NU + NQ + i] - qdn[i]);
                        });
                } else {
                        sfor<0, NX * NU>([&](auto ec) {
                                constexpr int i = ec, c = i / NX, r = i % NX, rd = r % NQ;
                                const float   d = R::template minv_sym<rd, c>(st.Minv);
                                putB(i, (r < NQ) ? (dt_sq_half * d) : (dt * d));
                        });
                }
        }

        // ---- the same linearisation one column per thread (small batches: k_kkt_fine) -------------------------------------------------
        // column c = K + W*NQ of A (W = 0: d/dq_K, W = 1: d/dqd_K); bit-identical to linearize() / linearize_half()
        template<int W, int K, class FA>
        static GATO_HD void linearize_column(const typename R::DynState& st, const float* qd, float dt, FA&& putA)
        {
                float dc[NQ], d[NQ];
                R::template rnea_grad_col<W, K>(st.X, qd, st.v, st.a, st.f, st.Iv, st.FxvI, dc);
                sfor<0, NQ>([&](auto rc) {
                        constexpr int row = rc;
                        float         val = 0.0f;
                        sfor<0, NQ>([&](auto cc) { val = fmaf(R::template minv_sym<row, cc>(st.Minv), dc[cc], val); });
                        d[row] = -val;
                });
                const float   dt_sq_half = (float)((0.5 * (double)dt) * (double)dt);
                constexpr int c = K + W * NQ;
                sfor<0, NX>([&](auto rc) {
                        constexpr int r = rc, rd = r % NQ;
                        float         val = (r == c) ? 1.0f : 0.0f;
                        if constexpr (r < NQ) {
                                if constexpr (c >= NQ && r == c - NQ) val = val + dt;
                                val = fmaf(dt_sq_half, d[rd], val);
                        } else {
                                val = fmaf(dt, d[rd], val);
                        }
                        putA(c * NX + r, val);
                });
        }
        // column `col` (a run-time value, uniform across the warp) of A: dispatch to the specialised column
        template<class FA>
        static GATO_HD void linearize_column_any(int col, const DynState& st, const float* qd, float dt, FA&& putA)
        {
                sfor<0, NX>([&](auto cc) {
                        constexpr int cidx = cc;
                        if (col == cidx) linearize_column<cidx / NQ, cidx % NQ>(st, qd, dt, putA);
                });
        }
        // B (from M^-1) and the defect c_{k+1}
        template<class FB, class Fc>
        static GATO_HD void linearize_base(const typename R::DynState& st, const float* xux, float dt, FB&& putB, Fc&& putc)
        {
                const float dt_sq_half = (float)((0.5 * (double)dt) * (double)dt);
                float       qn[NQ], qdn[NQ];
                R::integrate(xux, xux + NQ, st.qdd, dt, qn, qdn);
                sfor<0, NQ>([&](auto ic) {
                        constexpr int i = ic;
                        putc(i, xux[NX + NU + i] - qn[i]);
                        putc(i + NQ, xux[NX + NU + NQ + i] - qdn[i]);
                });
                sfor<0, NX * NU>([&](auto ec) {
                        constexpr int i = ec, c = i / NX, r = i % NX, rd = r % NQ;
                        const float   d = R::template minv_sym<rd, c>(st.Minv);
                        putB(i, (r < NQ) ? (dt_sq_half * d) : (dt * d));
                });
        }

        // ---- merit contribution of knot k ------------------------------------------------------------------
        // xux = z + alpha dz at knot k: x_k,u_k,x_{k+1} (k < N-1) or x_{N-1} only; x0err = |x_0 + alpha dz_0 - x_s| entries (used at k = N-1)
        template<bool LAST>
        static GATO_HD float tracking_cost(const float* xu, const float* ref3, const Costs& cs)
        {
                constexpr int TN = NQ + (LAST ? 0 : NU);
                float         cv[TN + 3], ee[3];
                R::ee_pos(xu, ee);
                sfor<0, NQ>([&](auto ic) {
                        constexpr int i = ic;
                        const float   err = xu[i + NQ];
                        float         c = ((0.5f * cs.qd_cost) * err) * err;
                        // A barrier whose weight is exactly zero is not evaluated (two logarithms each; the default parameters have vel_lim_cost =
                        // ctrl_lim_cost = 0): fmaf(0, barrier, c) == c bit for bit for every finite barrier value -- the clamped logarithms are finite for
                        // every finite state, and c >= +0 here -- and a NaN state still reaches the sum through the quadratic terms / the end-effector
                        // position.  (Only an infinite state entry under a zero weight would give NaN instead of the skipped term.)
                        if (cs.q_lim_cost != 0.0f) c = fmaf(cs.q_lim_cost, R::joint_barrier(xu[i], limit<P, 0, i, 0>(), limit<P, 0, i, 1>()), c);
                        if (cs.vel_lim_cost != 0.0f) c = fmaf(cs.vel_lim_cost, R::joint_barrier(xu[i + NQ], limit<P, 1, i, 0>(), limit<P, 1, i, 1>()), c);
                        cv[i] = c;
                });
                if constexpr (!LAST) {
                        sfor<0, NU>([&](auto jc) {
                                constexpr int j = jc;
                                const float   err = xu[NX + j];
                                float         c = ((0.5f * cs.u_cost) * err) * err;
                                if (cs.ctrl_lim_cost != 0.0f) c = fmaf(cs.ctrl_lim_cost, R::joint_barrier(xu[NX + j], limit<P, 2, j, 0>(), limit<P, 2, j, 1>()), c);
                                cv[NQ + j] = c;
                        });
                }
                const float w = LAST ? cs.N_cost : cs.q_cost;
                sfor<0, 3>([&](auto ic) {
                        constexpr int i = ic;
                        const float   err = ee[i] - ref3[i];
                        cv[TN + i] = (float)((((double)w * 0.5) * (double)err) * (double)err);
                });
                return tree_reduce<TN + 3>(cv);
        }
        static GATO_HD float merit_mid(const float* xux, const float* ref3, float mu, const float* fext, float dt, const Costs& cs)
        {
                const float cost = tracking_cost<false>(xux, ref3, cs);
                float       qdd[NQ], qn[NQ], qdn[NQ], err[NX];
                R::forward_dynamics(xux, xux + NQ, xux + NX, fext, qdd);
                R::integrate(xux, xux + NQ, qdd, dt, qn, qdn);
                sfor<0, NQ>([&](auto ic) {
                        constexpr int i = ic;
                        err[i] = fabsf(xux[NX + NU + i] - qn[i]);
                        err[i + NQ] = fabsf(xux[NX + NU + NQ + i] - qdn[i]);
                });
                const float cons = tree_reduce<NX>(err);
                return fmaf(mu, cons, cost);
        }
        // the two halves of merit_mid for the split kernel (small batches): merit = fmaf(mu, merit_mid_cons, tracking_cost<false>)
        static GATO_HD float merit_mid_cons(const float* xux, const float* fext, float dt)
        {
                float qdd[NQ], qn[NQ], qdn[NQ], err[NX];
                R::forward_dynamics(xux, xux + NQ, xux + NX, fext, qdd);
                R::integrate(xux, xux + NQ, qdd, dt, qn, qdn);
                sfor<0, NQ>([&](auto ic) {
                        constexpr int i = ic;
                        err[i] = fabsf(xux[NX + NU + i] - qn[i]);
                        err[i + NQ] = fabsf(xux[NX + NU + NQ + i] - qdn[i]);
                });
                return tree_reduce<NX>(err);
        }
        static GATO_HD float merit_last_cons(const float* x0err)
        {
                float err[NX];
                sfor<0, NX>([&](auto ic) { err[ic] = x0err[ic]; });
                return tree_reduce<NX>(err);
        }
        static GATO_HD float merit_last(const float* x, const float* ref3, float mu, const float* x0err, const Costs& cs)
        {
                const float cost = tracking_cost<true>(x, ref3, cs);
                float       err[NX];
                sfor<0, NX>([&](auto ic) { err[ic] = x0err[ic]; });
                const float cons = tree_reduce<NX>(err);
                return fmaf(mu, cons, cost);
        }
};

}  // namespace gato
