// Run-time robot model: the constants the compiled plants (rbd.cuh: Iiwa14, Indy7) fold into immediates, as a table that the
// table-driven dynamics (rbd_rt.cuh) reads from constant memory -- what lets a new robot in from a data file without code generation
// (the reference bakes these numbers into generated source: iiwa14_grid.cuh:1211-2087, 2212-2293, 2365-2448).
//
// Robots covered: fixed-base serial chains of z-axis revolute joints (what GRiD generates for the reference), nq = 6, 7 or 8 (the
// kernels are instantiated for nx = 12, 14 and 16).  A Pluecker transform X_j(q_j) = [E 0; B E] is held as its 18 independent entries: E = X[0:3,0:3]
// (the bottom-right block is a copy, iiwa14_grid.cuh:2287-2291) and B = X[3:6,0:3]; the top-right block is structurally zero for every joint.
#pragma once
#include "../../include/gato_b200.h"

namespace gato {

constexpr int kRtMaxQ = GATO_MODEL_MAX_NQ;
constexpr int kRtMaxXTrig = GATO_MODEL_MAX_TRIG;  // sin/cos-dependent entries of one joint's X (8 for both reference robots)
constexpr int kRtMaxHTrig = 8;                    // ... of one joint's 4x4 homogeneous transform (4)
constexpr int kRtSlots = 8;                       // models resident in constant memory at a time (plant ids 2 .. 2 + kRtSlots - 1); 8 x 7 KB of the 64 KB

// entry `loc` of a joint's matrix = (float)(coef * (double)(use_cos ? cos(q_k) : sin(q_k)))
struct RtTrig {
        int    loc;  // X: compact index 0..17 (see x_compact); Xhom / dXhom: 4 * col + row
        short  k, use_cos;
        double coef;
};
static_assert(sizeof(RtTrig) == 16, "two trig entries per 32 bytes of constant memory");

struct RtModel {
        int   nq, style;  // style 1: iiwa14-type limit barriers in the cost Hessian (iiwa14_plant.cuh:103-155), 0: indy7-type (indy7_plant.cuh:133-147)
        float X[kRtMaxQ][18];  // constant parts of E | B, x_compact order
        float I[kRtMaxQ][36];  // spatial inertias, column-major
        float Xh[kRtMaxQ][16], dXh[kRtMaxQ][16];  // constant parts of the homogeneous transforms and of their derivatives, column-major
        int   nxt[kRtMaxQ], nxht[kRtMaxQ], ndxht[kRtMaxQ];
        RtTrig xt[kRtMaxQ][kRtMaxXTrig], xht[kRtMaxQ][kRtMaxHTrig], dxht[kRtMaxQ][kRtMaxHTrig];
        float  jl[kRtMaxQ][2], vl[kRtMaxQ][2], cl[kRtMaxQ][2];  // joint / velocity / control limits with the reference's margin folded in
};

// (row, col) of the 6x6 transform, col < 3 -> index into the 18 stored entries: E column-major at 0..8, B column-major at 9..17
constexpr int x_compact(int row, int col) { return row < 3 ? 3 * col + row : 9 + 3 * col + (row - 3); }

// gato_model (file content, doubles) -> RtModel; returns 0 or a message describing what the model violates
inline const char* rt_model_from_desc(const gato_model& d, RtModel& m)
{
        if (d.nq < 6 || d.nq > 8) return "nq must be 6, 7 or 8 (the kernels are instantiated for nx = 12, 14 and 16)";
        if (d.style != 0 && d.style != 1) return "style must be 0 (indy7-type barriers) or 1 (iiwa14-type)";
        m = RtModel{};
        m.nq = d.nq, m.style = d.style;
        const int nq = d.nq;
        for (int j = 0; j < nq; j++) {
                for (int c = 0; c < 6; c++)
                        for (int r = 0; r < 6; r++) {
                                const double v = d.X[36 * j + 6 * c + r];
                                if (c < 3)
                                        m.X[j][x_compact(r, c)] = (float)v;
                                else if (r < 3 && v != 0.0)
                                        return "the top-right 3x3 block of every X must be zero (Pluecker transform of a revolute joint)";
                        }
                for (int e = 0; e < 36; e++) m.I[j][e] = (float)d.I[36 * j + e];
                for (int e = 0; e < 16; e++) m.Xh[j][e] = (float)d.Xhom[16 * j + e], m.dXh[j][e] = (float)d.dXhom[16 * j + e];
                // limits: "double literal -/+ float margin", folded in double, stored as float (iiwa14_plant.cuh:31-70, indy7_plant.cuh:61-96)
                const double margin = (double)(float)(-0.1);
                m.jl[j][0] = (float)(-d.joint_limit[j] - margin), m.jl[j][1] = (float)(d.joint_limit[j] + margin);
                m.vl[j][0] = (float)(-d.vel_limit[j] - margin), m.vl[j][1] = (float)(d.vel_limit[j] + margin);
                m.cl[j][0] = (float)(-d.ctrl_limit[j] - margin), m.cl[j][1] = (float)(d.ctrl_limit[j] + margin);
        }
        auto add = [&](const gato_trig* src, int n, int per, int cap, int* cnt, RtTrig* dst, bool is_x) -> const char* {
                if (n < 0 || n > kRtMaxQ * (is_x ? GATO_MODEL_MAX_TRIG : 8)) return "trig entry count out of range";
                for (int i = 0; i < n; i++) {
                        const int idx = src[i].idx, k = src[i].k;
                        if (idx < 0 || idx >= per * nq || k < 0 || k >= 2 * nq) return "trig entry index out of range";
                        const int j = idx / per, loc = idx % per;
                        if (cnt[j] >= cap) return "too many trig entries for one joint";
                        RtTrig& t = dst[j * cap + cnt[j]++];
                        if (is_x) {
                                const int c = loc / 6, r = loc % 6;
                                if (c >= 3) return "trig entries of X must lie in its first three columns (the bottom-right block is a copy, the top-right zero)";
                                t.loc = x_compact(r, c);
                        } else {
                                t.loc = loc;
                        }
                        t.k = (short)(k % nq), t.use_cos = (short)(k >= nq ? 1 : 0), t.coef = src[i].coef;
                }
                return nullptr;
        };
        if (const char* e = add(d.x_trig, d.n_x_trig, 36, kRtMaxXTrig, m.nxt, &m.xt[0][0], true)) return e;
        if (const char* e = add(d.xh_trig, d.n_xh_trig, 16, kRtMaxHTrig, m.nxht, &m.xht[0][0], false)) return e;
        if (const char* e = add(d.dxh_trig, d.n_dxh_trig, 16, kRtMaxHTrig, m.ndxht, &m.dxht[0][0], false)) return e;
        return nullptr;
}

}  // namespace gato
