// Geometric calibration reductions (utils/model_calibration.py).
//
// K5 rk_calib_normal_eq_f64: H = sum J^T J, g = sum J^T r over (measurement, tool) -- a CTA
//    alternates phase 1 (two warp roles walk the reduced chain backwards once per measurement and
//    emit 6 QR-compressed rows [J | r] in the end-effector frame into a shared-memory slab) and
//    phase 2 (Gram accumulation of the slab on the FP64 tensor cores); see calib_normal_eq6_kernel.
// K6 rk_calib_tool_base_eq_f64: the blocks of estimate_tool_base's A^T A / A^T dp are sums of 45
//    per-measurement numbers, accumulated in registers (calib_tool_base_stats_kernel);
//    rk_calib_eval_f64: residual sums and maxima.
// Every reduction ends the same way: CTA partials go to the workspace and the last CTA to finish
// (atomic ticket) adds them in index order, so results are deterministic.
#include <cstdlib>
#include <type_traits>

#include "rk_common.cuh"

namespace rk {

constexpr int kMeas = 128;     // threads per CTA
constexpr int kW = 20;         // logical row width: <= 19 columns + residual
constexpr int kMaxGrid = 1024;
// 6-row kernel (tensor-core Gram): rows of kW doubles at an odd stride (phase-1 stores and the
// DMMA fragment loads are both bank-conflict free, see gram_dmma); 8x8 tiles of a 24x24 Gram
constexpr int kRS6 = 21;
constexpr int kTiles6 = 6;
constexpr int kPart6 = kTiles6 * 64;                  // doubles per CTA partial
constexpr int kPartStride = kPart6 + 8;               // + r^T r; also covers the other reductions

struct CalibP {
  int nj, np, nt;
  int jaxis[RK_CALIB_MAX_JOINTS], jkind[RK_CALIB_MAX_JOINTS];
  double jdir[RK_CALIB_MAX_JOINTS], joff[RK_CALIB_MAX_JOINTS];
  double seg[RK_CALIB_MAX_JOINTS][12];
  int pstart[RK_CALIB_MAX_JOINTS + 1];  // parameters of joint j: [pstart[j], pstart[j+1])
  int pkind[RK_CALIB_MAX_PARAMS];
  double pax[RK_CALIB_MAX_PARAMS][3], porg[RK_CALIB_MAX_PARAMS][3];
  double Tb[12];
  double tool[RK_CALIB_MAX_TOOLS][3];
  int seg_trans[RK_CALIB_MAX_JOINTS];  // seg[j] is a pure translation (true at the nominal model)
};

static int make_calib(const RkCalibModel* m, CalibP& P, bool identity_base) {
  if (m->njoints < 1 || m->njoints > RK_CALIB_MAX_JOINTS) return RK_ERR_UNSUPPORTED;
  if (m->nparams < 0 || m->nparams > RK_CALIB_MAX_PARAMS) return RK_ERR_UNSUPPORTED;
  if (m->ntools < 1 || m->ntools > RK_CALIB_MAX_TOOLS) return RK_ERR_UNSUPPORTED;
  P.nj = m->njoints, P.np = m->nparams, P.nt = m->ntools;
  for (int j = 0; j < RK_CALIB_MAX_JOINTS; ++j) {
    const bool on = j < m->njoints;
    P.jaxis[j] = on ? m->joint_axis[j] : 0;
    P.jkind[j] = on ? m->joint_kind[j] : 1;
    P.jdir[j] = on ? m->joint_dir[j] : 0.0;
    P.joff[j] = on ? m->joint_off[j] : 0.0;
    for (int k = 0; k < 12; ++k) P.seg[j][k] = on ? m->seg[j][k] : ((k % 5 == 0) ? 1.0 : 0.0);
    if (on && (P.jaxis[j] < 0 || P.jaxis[j] > 2)) return RK_ERR_INVALID_ARG;
  }
  int p = 0;
  for (int j = 0; j < m->njoints; ++j) {  // parameters must be grouped by joint, ascending
    P.pstart[j] = p;
    while (p < m->nparams && m->param_joint[p] == j) ++p;
  }
  if (p != m->nparams) return RK_ERR_INVALID_ARG;
  for (int j = m->njoints; j <= RK_CALIB_MAX_JOINTS; ++j) P.pstart[j] = p;
  for (int q = 0; q < RK_CALIB_MAX_PARAMS; ++q) {
    const bool on = q < m->nparams;
    P.pkind[q] = on ? m->param_kind[q] : 0;
    for (int k = 0; k < 3; ++k) {
      P.pax[q][k] = on ? m->param_axis[q][k] : 0.0;
      P.porg[q][k] = on ? m->param_origin[q][k] : 0.0;
    }
  }
  for (int j = 0; j < RK_CALIB_MAX_JOINTS; ++j) {
    bool ident = true;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) ident = ident && P.seg[j][r * 4 + c] == ((r == c) ? 1.0 : 0.0);
    P.seg_trans[j] = ident ? 1 : 0;
  }
  for (int k = 0; k < 12; ++k)
    P.Tb[k] = identity_base ? ((k % 5 == 0) ? 1.0 : 0.0) : m->T_base[k];
  for (int t = 0; t < RK_CALIB_MAX_TOOLS; ++t)
    for (int k = 0; k < 3; ++k)
      P.tool[t][k] = (identity_base || t >= m->ntools) ? 0.0 : m->tool_p[t][k];
  return RK_OK;
}

struct Fr {
  double R[9], p[3];
};

__device__ __forceinline__ void fr_init(Fr& F, const double* T12) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) F.R[i * 3 + j] = T12[i * 4 + j];
    F.p[i] = T12[i * 4 + 3];
  }
}

template <int A>
__device__ __forceinline__ void fr_rot(Fr& F, double s, double c) {
  constexpr int I = (A + 1) % 3, Jc = (A + 2) % 3;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double a = F.R[r * 3 + I], b = F.R[r * 3 + Jc];
    F.R[r * 3 + I] = c * a + s * b;
    F.R[r * 3 + Jc] = c * b - s * a;
  }
}

// apply joint j (value = sin/cos for revolute, displacement in `s` for prismatic)
__device__ __forceinline__ void fr_joint(Fr& F, int axis, int kind, double s, double c) {
  if (kind == 1) {
    switch (axis) {
      case 0: fr_rot<0>(F, s, c); break;
      case 1: fr_rot<1>(F, s, c); break;
      default: fr_rot<2>(F, s, c); break;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i) F.p[i] += s * F.R[i * 3 + axis];
  }
}

__device__ __forceinline__ void fr_mul(Fr& F, const double* X) {
  double Rn[9], pn[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Rn[i * 3 + j] = F.R[i * 3] * X[j] + F.R[i * 3 + 1] * X[4 + j] + F.R[i * 3 + 2] * X[8 + j];
    pn[i] = F.R[i * 3] * X[3] + F.R[i * 3 + 1] * X[7] + F.R[i * 3 + 2] * X[11] + F.p[i];
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) F.R[i] = Rn[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) F.p[i] = pn[i];
}

// joint values of one measurement: (sin, cos) for revolute joints, displacement otherwise
__device__ __forceinline__ void load_joints(const CalibP& P, const double* __restrict__ q, int64_t i,
                                            bool live, double (&s)[RK_CALIB_MAX_JOINTS],
                                            double (&c)[RK_CALIB_MAX_JOINTS]) {
  // all joint values are fetched before the first sincos (independent loads in flight together;
  // inside the loop below every uniform branch would fence the next load behind it)
  double raw[RK_CALIB_MAX_JOINTS];
#pragma unroll
  for (int j = 0; j < RK_CALIB_MAX_JOINTS; ++j) raw[j] = (live && j < P.nj) ? __ldcs(q + i * P.nj + j) : 0.0;
#pragma unroll
  for (int j = 0; j < RK_CALIB_MAX_JOINTS; ++j) {
    s[j] = 0.0, c[j] = 1.0;
    if (j < P.nj) {
      const double v = __dadd_rn(__dmul_rn(raw[j], P.jdir[j]), P.joff[j]);  // model_calibration.py:557
      if (P.jkind[j] == 1)
        fast_sincos(v, &s[j], &c[j]);
      else
        s[j] = v;
    }
  }
}

// whole chain (no tool): get_pose with T_tool = I (model_calibration.py:541-548)
__device__ __forceinline__ void chain_end(const CalibP& P, const double (&s)[RK_CALIB_MAX_JOINTS],
                                          const double (&c)[RK_CALIB_MAX_JOINTS], Fr& F) {
  fr_init(F, P.Tb);
#pragma unroll
  for (int j = 0; j < RK_CALIB_MAX_JOINTS; ++j) {
    if (j < P.nj) {
      fr_joint(F, P.jaxis[j], P.jkind[j], s[j], c[j]);
      if (P.seg_trans[j]) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
          F.p[i] += F.R[i * 3] * P.seg[j][3] + F.R[i * 3 + 1] * P.seg[j][7] + F.R[i * 3 + 2] * P.seg[j][11];
      } else {
        fr_mul(F, P.seg[j]);
      }
    }
  }
}

// -------------------------------------------------------------------------------------------
// K5 fast path (normal equations): 6 rows per measurement instead of 3 per (measurement, tool).
//
// (1) H and g are invariant under a rotation of each measurement's rows, so the Jacobian is
//     written in the END-EFFECTOR frame.  With S_j = X_j Rot_{j+1} X_{j+1} ... = (Q_j, s_j) the
//     constant-plus-joint suffix after joint j, a parameter with axis a and origin o in joint j's
//     frame has the column  a' x (b' + tool_t),  a' = Q_j^T a,  b' = Q_j^T (s_j - o)  (rotation
//     parameter) or a' (translation parameter).  Only suffixes appear, so ONE backward walk
//     over the chain serves every parameter and every tool (the forward formulation walked the
//     chain once per tool), and the walk ends with the whole chain, which the residual needs.
// (2) The 3*nt rows of all tools are  M [c; a'],  c = a' x b',  M = [I | -[tool_t]x] stacked: a
//     CONSTANT (3 nt) x 6 matrix.  With M = Q R (Householder, host) the Gram matrix of the
//     3 nt rows equals that of the 6 rows  y = R [c; a'],  and  J^T r = Y^T (Q^T rho)  with rho
//     the stacked end-frame residuals.  Phase 2 therefore accumulates 6 rows per measurement
//     (2/3 of the DFMAs for three tools); r^T r is summed separately.
// A CTA pass covers kMeas/2 measurements; warps 0,1 and warps 2,3 split the parameters of a
// measurement by joint (`split`, chosen by the host to balance the two roles): role B walks the
// whole chain and owns joints < split, the residual and the padding; role A walks joints >= split.
// -------------------------------------------------------------------------------------------
struct CalibN {
  double tbar[3];                         // centroid of the tools
  double R[6][6];                         // block diagonal: R11 diagonal, R22 upper triangular
  double Qt[6][3 * RK_CALIB_MAX_TOOLS];   // first 6 rows of Q^T
  int split;
};

// Householder QR of the (3 nt) x 6 matrix M = [I | -[tool_t]x]; rows of R / Q^T beyond
// min(3 nt, 6) are zero.
static void make_calibn(const CalibP& P, CalibN& N) {
  const int mrows = 3 * P.nt;
  double A[3 * RK_CALIB_MAX_TOOLS][6], Qm[3 * RK_CALIB_MAX_TOOLS][3 * RK_CALIB_MAX_TOOLS];
  // tools are taken relative to their centroid (the kernel adds the centroid to b'): then the
  // skew blocks sum to zero and R = diag(R11 (diagonal 3x3), R22 (upper triangular 3x3))
  for (int k = 0; k < 3; ++k) {
    N.tbar[k] = 0.0;
    for (int t = 0; t < P.nt; ++t) N.tbar[k] += P.tool[t][k] / P.nt;
  }
  for (int t = 0; t < P.nt; ++t) {
    const double w[3] = {P.tool[t][0] - N.tbar[0], P.tool[t][1] - N.tbar[1], P.tool[t][2] - N.tbar[2]};
    const double K[3][3] = {{0.0, w[2], -w[1]}, {-w[2], 0.0, w[0]}, {w[1], -w[0], 0.0}};  // -[w]x
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        A[3 * t + r][c] = (r == c) ? 1.0 : 0.0;
        A[3 * t + r][3 + c] = K[r][c];
      }
  }
  for (int i = 0; i < mrows; ++i)
    for (int j = 0; j < mrows; ++j) Qm[i][j] = (i == j) ? 1.0 : 0.0;  // accumulates Q^T
  const int steps = mrows < 6 ? mrows : 6;
  for (int k = 0; k < steps; ++k) {
    double norm = 0.0;
    for (int i = k; i < mrows; ++i) norm += A[i][k] * A[i][k];
    norm = sqrt(norm);
    if (norm == 0.0) continue;
    double v[3 * RK_CALIB_MAX_TOOLS];
    const double alpha = A[k][k] > 0.0 ? -norm : norm;
    double vn = 0.0;
    for (int i = k; i < mrows; ++i) {
      v[i] = A[i][k] - (i == k ? alpha : 0.0);
      vn += v[i] * v[i];
    }
    if (vn == 0.0) continue;
    for (int c = 0; c < 6; ++c) {
      double d = 0.0;
      for (int i = k; i < mrows; ++i) d += v[i] * A[i][c];
      d *= 2.0 / vn;
      for (int i = k; i < mrows; ++i) A[i][c] -= d * v[i];
    }
    for (int c = 0; c < mrows; ++c) {
      double d = 0.0;
      for (int i = k; i < mrows; ++i) d += v[i] * Qm[i][c];
      d *= 2.0 / vn;
      for (int i = k; i < mrows; ++i) Qm[i][c] -= d * v[i];
    }
  }
  for (int i = 0; i < 6; ++i) {
    // the off-diagonal block and the off-diagonal of R11 vanish analytically (rounding noise only)
    for (int c = 0; c < 6; ++c)
      N.R[i][c] = (i < mrows && c >= i && (i < 3 ? c == i : true)) ? A[i][c] : 0.0;
    for (int c = 0; c < 3 * RK_CALIB_MAX_TOOLS; ++c) N.Qt[i][c] = (i < mrows && c < mrows) ? Qm[i][c] : 0.0;
  }
  // role balance (FP64 operation estimates: sincos 45, walk step 55, rotation / translation
  // parameter 48 / 15, residual 36 + 18 per tool)
  int best = 0;
  double best_cost = 1e300;
  for (int sp = 0; sp <= P.nj; ++sp) {
    double ca = 0.0, cb = 45.0 * P.nj + 55.0 * P.nj + 54.0 * P.nt;
    for (int j = 0; j < P.nj; ++j) {
      double pc = 0.0;
      for (int p = P.pstart[j]; p < P.pstart[j + 1]; ++p) pc += P.pkind[p] == 1 ? 48.0 : 15.0;
      if (j >= sp) ca += pc + (j > sp ? 100.0 : 0.0);
      else cb += pc;
    }
    const double cost = ca > cb ? ca : cb;
    if (cost < best_cost) best_cost = cost, best = sp;
  }
  N.split = best;
}

// Suffix transform S = (Q, s), 3x4 row-major in registers.
struct Sfx {
  double Q[9], s[3];
};

// S <- Rot_axis(theta) * S  (revolute) or Trans_axis(d) * S (prismatic: d passed in `sn`)
template <int A>
__device__ __forceinline__ void sfx_rot(Sfx& S, double sn, double cs) {
  constexpr int I = (A + 1) % 3, Jc = (A + 2) % 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double a = S.Q[I * 3 + c], b = S.Q[Jc * 3 + c];
    S.Q[I * 3 + c] = cs * a - sn * b;
    S.Q[Jc * 3 + c] = sn * a + cs * b;
  }
  const double a = S.s[I], b = S.s[Jc];
  S.s[I] = cs * a - sn * b;
  S.s[Jc] = sn * a + cs * b;
}
__device__ __forceinline__ void sfx_joint(Sfx& S, int axis, int kind, double sn, double cs) {
  if (kind == 1) {
    switch (axis) {
      case 0: sfx_rot<0>(S, sn, cs); break;
      case 1: sfx_rot<1>(S, sn, cs); break;
      default: sfx_rot<2>(S, sn, cs); break;
    }
  } else {  // (static indices keep S in registers)
    S.s[0] += axis == 0 ? sn : 0.0;
    S.s[1] += axis == 1 ? sn : 0.0;
    S.s[2] += axis == 2 ? sn : 0.0;
  }
}
// S <- X * S for a constant 3x4 X
__device__ __forceinline__ void sfx_premul(Sfx& S, const double* X) {
  double Qn[9], sn[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      Qn[i * 3 + c] = X[i * 4] * S.Q[c] + X[i * 4 + 1] * S.Q[3 + c] + X[i * 4 + 2] * S.Q[6 + c];
    sn[i] = X[i * 4] * S.s[0] + X[i * 4 + 1] * S.s[1] + X[i * 4 + 2] * S.s[2] + X[i * 4 + 3];
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) S.Q[i] = Qn[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) S.s[i] = sn[i];
}

// Phase 1 of the 6-row kernel for one (measurement, role): writes this role's columns of the six
// slab rows (rows + r * 64 * kRS6, r < 6); returns this measurement's r^T r (role B) or 0.
__device__ __forceinline__ double fill_rows6(const CalibP& P, const CalibN& N, bool role_a,
                                             const double (&qv)[RK_CALIB_MAX_JOINTS],
                                             const double (&pv)[3 * RK_CALIB_MAX_TOOLS], bool live,
                                             int ncols, double* __restrict__ rows) {
  constexpr int kRowStep = (kMeas / 2) * kRS6;  // row r of measurement slot ms is slab row r*64 + ms
  const int wpad = kW;  // the tensor-core Gram reads all kW columns of a row
  const int split = N.split;
  // columns owned by this role: A [pstart[split], np), B [0, pstart[split]) and [np, wpad)
  if (!live) {
    const int c0 = role_a ? P.pstart[split] : 0, c1 = role_a ? P.np : P.pstart[split];
    for (int r = 0; r < 6; ++r) {
      for (int k = c0; k < c1; ++k) rows[r * kRowStep + k] = 0.0;
      if (!role_a)
        for (int k = P.np; k < wpad; ++k) rows[r * kRowStep + k] = 0.0;
    }
    return 0.0;
  }
  const int jstop = role_a ? split : 0;  // last joint whose suffix this role needs
  Sfx S;
#pragma unroll
  for (int k = 0; k < 9; ++k) S.Q[k] = (k % 4 == 0) ? 1.0 : 0.0;
  S.s[0] = S.s[1] = S.s[2] = 0.0;
  // rolled on purpose: unrolled over 8 joints the walk is ~38 KB of code (> the 32 KB L1.5
  // instruction cache; ncu: no_instruction 0.8 per issue)
#pragma unroll 1
  for (int j = P.nj - 1; j >= jstop; --j) {
    {
      if (P.seg_trans[j]) {      // S_j = X_j Rot_{j+1} S_{j+1}
        S.s[0] += P.seg[j][3], S.s[1] += P.seg[j][7], S.s[2] += P.seg[j][11];
      } else {
        sfx_premul(S, P.seg[j]);
      }
      if (role_a || j < split) {
        for (int p = P.pstart[j]; p < P.pstart[j + 1]; ++p) {
          const double ax = P.pax[p][0], ay = P.pax[p][1], az = P.pax[p][2];
          // a' = Q^T a
          const double a0 = S.Q[0] * ax + S.Q[3] * ay + S.Q[6] * az;
          const double a1 = S.Q[1] * ax + S.Q[4] * ay + S.Q[7] * az;
          const double a2 = S.Q[2] * ax + S.Q[5] * ay + S.Q[8] * az;
          double y[6];
          if (P.pkind[p] == 1) {
            // b'' = Q^T (s - o) + centroid of the tools
            const double dx = S.s[0] - P.porg[p][0], dy = S.s[1] - P.porg[p][1], dz = S.s[2] - P.porg[p][2];
            const double b0 = S.Q[0] * dx + S.Q[3] * dy + S.Q[6] * dz + N.tbar[0];
            const double b1 = S.Q[1] * dx + S.Q[4] * dy + S.Q[7] * dz + N.tbar[1];
            const double b2 = S.Q[2] * dx + S.Q[5] * dy + S.Q[8] * dz + N.tbar[2];
            y[0] = N.R[0][0] * (a1 * b2 - a2 * b1);
            y[1] = N.R[1][1] * (a2 * b0 - a0 * b2);
            y[2] = N.R[2][2] * (a0 * b1 - a1 * b0);
            y[3] = N.R[3][3] * a0 + N.R[3][4] * a1 + N.R[3][5] * a2;
            y[4] = N.R[4][4] * a1 + N.R[4][5] * a2;
            y[5] = N.R[5][5] * a2;
          } else {
            y[0] = N.R[0][0] * a0, y[1] = N.R[1][1] * a1, y[2] = N.R[2][2] * a2;
            y[3] = y[4] = y[5] = 0.0;
          }
#pragma unroll
          for (int r = 0; r < 6; ++r) rows[r * kRowStep + p] = y[r];
        }
      }
      if (j > jstop || !role_a) {  // S <- Rot_j(q_j) S_j  (role A never needs Rot_split)
        const double v = __dadd_rn(__dmul_rn(qv[j], P.jdir[j]), P.joff[j]);  // model_calibration.py:557
        double sn = v, cs = 1.0;
        if (P.jkind[j] == 1) fast_sincos(v, &sn, &cs);
        sfx_joint(S, P.jaxis[j], P.jkind[j], sn, cs);
      }
    }
  }
  if (role_a) return 0.0;
  // role B: S = Rot_0 S_0; T_end = T_base S.  Residual of tool t in the end frame:
  // rho_t = R_end^T (p_meas - p_end) - tool_t   (r = p_meas - T[:3,3], model_calibration.py:420)
  sfx_premul(S, P.Tb);
  double rho[3 * RK_CALIB_MAX_TOOLS];
  double rr = 0.0;
#pragma unroll
  for (int t = 0; t < RK_CALIB_MAX_TOOLS; ++t) {
    if (t < P.nt) {
      const double dx = pv[3 * t] - S.s[0], dy = pv[3 * t + 1] - S.s[1], dz = pv[3 * t + 2] - S.s[2];
      rho[3 * t + 0] = S.Q[0] * dx + S.Q[3] * dy + S.Q[6] * dz - P.tool[t][0];
      rho[3 * t + 1] = S.Q[1] * dx + S.Q[4] * dy + S.Q[7] * dz - P.tool[t][1];
      rho[3 * t + 2] = S.Q[2] * dx + S.Q[5] * dy + S.Q[8] * dz - P.tool[t][2];
      rr += rho[3 * t] * rho[3 * t] + rho[3 * t + 1] * rho[3 * t + 1] + rho[3 * t + 2] * rho[3 * t + 2];
    } else {
      rho[3 * t] = rho[3 * t + 1] = rho[3 * t + 2] = 0.0;
    }
  }
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 3 * RK_CALIB_MAX_TOOLS; ++k) acc = fma(N.Qt[r][k], rho[k], acc);
    rows[r * kRowStep + P.np] = acc;
  }
  for (int k = ncols; k < wpad; ++k) {
#pragma unroll
    for (int r = 0; r < 6; ++r) rows[r * kRowStep + k] = 0.0;
  }
  return rr;
}

// Phase 2 of the 6-row kernel on the FP64 tensor cores: G += Y^T Y with mma.sync.m8n8k4.f64.
// The Gram matrix is cut into 8x8 tiles (I, J), I <= J < 3 (columns 20..23 are zero padding).  For
// a group of 4 slab rows k the fragment "row k, columns 8c..8c+7" -- lane l holds
// Y[k = l%4][8c + l/4] -- is BOTH the A operand (Y^T block c, row-major) of tiles (c, .) and the B
// operand (Y block c, column-major) of tiles (., c), so a group costs 3 LDS.64 and 6 DMMA
// (= 1536 FMAs); the CUDA-core version needed 4 LDS.128 per 16 DFMA and ran into the
// shared-memory bandwidth (ncu: L1 data pipe 68 %, FP64 pipe 42 %).  Measured DMMA rate on B200:
// 1.86e13 FMA/s, the same as DFMA (rk_probe_dmma), so this is not about peak but about feeding it.
// Rows of a group are {r, r+8, r+16, r+24}: with the odd row stride kRS6 the four rows start 16
// banks apart, which makes the fragment load two conflict-free wavefronts.
struct GramDmma {
  double acc[kTiles6][2];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int t = 0; t < kTiles6; ++t) acc[t][0] = acc[t][1] = 0.0;
  }
  static __device__ __forceinline__ void mma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
  }
  // all kMeas*3 slab rows; warp w takes the row groups g = w (mod 4)
  __device__ __forceinline__ void accumulate(const double* __restrict__ slab, int warp, int lane) {
    const int k = lane & 3, c = lane >> 2;
    constexpr int kGroups = kMeas * 3 / 4;
#pragma unroll 2
    for (int g = warp; g < kGroups; g += kMeas / kWarp) {
      const int row = (g >> 3) * 32 + (g & 7) + 8 * k;
      const double* rp = slab + row * kRS6 + c;
      const double v0 = rp[0], v1 = rp[8];
      const double v2 = c < 4 ? rp[16] : 0.0;
      mma(acc[0], v0, v0), mma(acc[1], v0, v1), mma(acc[2], v0, v2);
      mma(acc[3], v1, v1), mma(acc[4], v1, v2), mma(acc[5], v2, v2);
    }
  }
};

__global__ void __launch_bounds__(kMeas, 3)
calib_normal_eq6_kernel(const __grid_constant__ CalibP P, const __grid_constant__ CalibN N,
                        const double* __restrict__ q, const double* __restrict__ pts, int64_t m,
                        int ncols, double* __restrict__ partial, unsigned int* __restrict__ counter,
                        double* __restrict__ out) {
  extern __shared__ __align__(16) double slab[];  // [(kMeas/2)*6 rows][kRS6]
  __shared__ int s_last;
  __shared__ double s_rr[kMeas / kWarp];
  constexpr int kHalf = kMeas / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool role_a = tid >= kHalf;  // warp-uniform
  const int ms = tid & (kHalf - 1);
  GramDmma G;
  G.init();
  double rr = 0.0;
  // software pipeline: the joint vector (and, for role B, the measured points) of the NEXT pass
  // are loaded before this pass's phase 2 and consumed after it
  double qn[RK_CALIB_MAX_JOINTS], pn[3 * RK_CALIB_MAX_TOOLS];
  auto prefetch = [&](int64_t i) {
    const bool live = i < m;
#pragma unroll
    for (int j = 0; j < RK_CALIB_MAX_JOINTS; ++j) qn[j] = (live && j < P.nj) ? __ldcs(q + i * P.nj + j) : 0.0;
#pragma unroll
    for (int k = 0; k < 3 * RK_CALIB_MAX_TOOLS; ++k)
      pn[k] = (live && !role_a && k < 3 * P.nt) ? __ldcs(pts + i * 3 * P.nt + k) : 0.0;
  };
  const int64_t stride = (int64_t)gridDim.x * kHalf;
  int64_t base = (int64_t)blockIdx.x * kHalf;
  prefetch(base + ms);
  for (; base < m; base += stride) {
    double qv[RK_CALIB_MAX_JOINTS], pv[3 * RK_CALIB_MAX_TOOLS];
#pragma unroll
    for (int j = 0; j < RK_CALIB_MAX_JOINTS; ++j) qv[j] = qn[j];
#pragma unroll
    for (int k = 0; k < 3 * RK_CALIB_MAX_TOOLS; ++k) pv[k] = pn[k];
    __syncthreads();  // previous phase 2 is done with the slab
    rr += fill_rows6(P, N, role_a, qv, pv, base + ms < m, ncols, slab + ms * kRS6);
    prefetch(base + stride + ms);
    __syncthreads();
    G.accumulate(slab, warp, lane);
  }

  // warps -> CTA partial (reuse the slab): lane l holds C[l/4][2(l%4) + e] of every tile
  __syncthreads();
#pragma unroll
  for (int t = 0; t < kTiles6; ++t)
#pragma unroll
    for (int e = 0; e < 2; ++e)
      slab[(warp * kTiles6 + t) * 64 + (lane >> 2) * 8 + 2 * (lane & 3) + e] = G.acc[t][e];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rr += __shfl_xor_sync(0xffffffffu, rr, o);
  if (lane == 0) s_rr[warp] = rr;
  __syncthreads();
  double* mine = partial + (int64_t)blockIdx.x * kPartStride;
  for (int e = tid; e < kPart6; e += kMeas) {
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < kMeas / kWarp; ++w) sum += slab[w * kPart6 + e];
    mine[e] = sum;
  }
  if (tid == 0) mine[kPart6] = (s_rr[0] + s_rr[1]) + (s_rr[2] + s_rr[3]);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA: deterministic sum over CTAs (index order), then scatter into the packed format
  const int np = ncols - 1;
  for (int e = tid; e <= kPart6; e += kMeas) {
    double sum = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) sum += __ldcg(partial + (int64_t)b * kPartStride + e);
    if (e == kPart6) {
      out[np * (np + 1) / 2 + np] = sum;  // r^T r
      continue;
    }
    const int t = e >> 6, r = (e >> 3) & 7, c = e & 7;
    const int I = t < 3 ? 0 : (t < 5 ? 1 : 2), J = t < 3 ? t : (t < 5 ? t - 2 : 2);
    const int gi = 8 * I + r, gj = 8 * J + c;
    if (gi > gj || gj >= ncols || gi >= np) continue;
    if (gj < np)
      out[gi * np - gi * (gi - 1) / 2 + (gj - gi)] = sum;  // packed upper H
    else
      out[np * (np + 1) / 2 + gi] = sum;                   // g
  }
  if (tid == 0) *counter = 0;  // workspace is reusable without a memset
}

// -------------------------------------------------------------------------------------------
// K5 with the model STRUCTURE fixed at compile time (values stay runtime).
//
// The generic kernel above interprets the model: the joint walk is a rolled loop over runtime
// tables (parameter lists per joint, parameter kinds, joint axes), so every constant it touches
// costs an LDC, every parameter a kind branch, every joint an axis switch -- three quarters of its
// instructions were phase 1 and only a third of those FP64 (ncu, round 1).  Here a Shape type
// carries the structure as constant expressions: after full unrolling every table index is
// static, the constants become constant-bank operands of the DFMAs themselves, and a joint is ONE
// straight-line variant instead of a switch over three.  The Shape below is the Fanuc165F reduced
// model of fanuc_calibration.py (the reference's one real calibration model: six revolute joints
// z y y x y x, 18 parameters); any other model takes the generic kernel.
//
// Columns are also reordered in the slab: the 8 translation parameters first, then the 10 rotation
// parameters, the residual and one zero column.  The three "angular" rows of a measurement
// (y[3..5] = R22 a') are identically zero in the translation columns, so their Gram update needs
// only the 3 tiles of the rotation blocks, not 6: 6.75 DMMA per measurement instead of 9.
// -------------------------------------------------------------------------------------------
struct FanucReducedShape {
  static constexpr int NJ = 6, NP = 18;
  static constexpr int SPLIT = 1;  // role A: joints >= SPLIT (make_calibn's cost model picks 1 for this model)
  static __host__ __device__ constexpr int jaxis(int j) {
    constexpr int a[NJ] = {2, 1, 1, 0, 1, 0};
    return a[j];
  }
  static __host__ __device__ constexpr int pstart(int j) {
    constexpr int a[NJ + 1] = {0, 3, 7, 11, 15, 18, 18};
    return a[j];
  }
  static __host__ __device__ constexpr int pkind(int p) {  // 1 rotation, 0 translation
    constexpr int a[NP] = {0, 0, 1, 1, 0, 1, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0, 1};
    return a[p];
  }
  // slab column of parameter p: translations 0..7 in parameter order, then rotations 8..17
  // (checked against pkind by shape_matches)
  static __host__ __device__ constexpr int pcol(int p) {
    constexpr int a[NP] = {0, 1, 8, 9, 2, 10, 11, 12, 3, 4, 13, 14, 5, 6, 15, 16, 7, 17};
    return a[p];
  }
  static constexpr int NTRANS = 8;  // == number of translation parameters (one 8-column block)
};

// Slab of the specialised kernel: the three linear rows of a measurement are 20 wide (stride 21), the
// three angular rows only hold the rotation / residual / zero columns 8..19 (12 wide, stride 13):
// 52 KB per CTA instead of 64.5 KB, so FOUR CTAs are resident per SM instead of three.  (ncu: the
// kernel is latency-bound -- issue 38 %, FP64 pipe 28 %, 12 warps per SM -- so resident warps count.)
constexpr int kRSB = 13;                                  // row stride of the angular rows
constexpr int kSlabA = (kMeas / 2) * 3 * kRS6;            // doubles in the linear region
constexpr int kSlabB = (kMeas / 2) * 3 * kRSB;            // doubles in the angular region
constexpr int kSlabStatic = kSlabA + kSlabB;

template <class Shape>
static bool shape_matches(const CalibP& P) {
  if (P.nj != Shape::NJ || P.np != Shape::NP) return false;
  int ntr = 0;
  for (int j = 0; j < Shape::NJ; ++j)
    if (P.jkind[j] != 1 || P.jaxis[j] != Shape::jaxis(j) || P.pstart[j] != Shape::pstart(j)) return false;
  if (P.pstart[Shape::NJ] != Shape::pstart(Shape::NJ)) return false;
  int nrot = 0;
  for (int p = 0; p < Shape::NP; ++p) {
    if (P.pkind[p] != Shape::pkind(p)) return false;
    if (Shape::pcol(p) != (P.pkind[p] ? Shape::NTRANS + nrot : ntr)) return false;  // the table above
    (P.pkind[p] ? nrot : ntr) += 1;
  }
  return ntr == Shape::NTRANS;
}

// parameter index of every slab column (the residual column maps to NP); host side
struct ColMap {
  int param_of_col[kW];
};
template <class Shape>
static ColMap make_colmap() {
  ColMap m;
  for (int c = 0; c < kW; ++c) m.param_of_col[c] = Shape::NP;
  for (int p = 0; p < Shape::NP; ++p) m.param_of_col[Shape::pcol(p)] = p;
  return m;
}

// Phase 1 for one (measurement, role), every index static: joints and parameters are expanded by
// template recursion (a `#pragma unroll` over bounds that come out of constexpr functions stayed a
// rolled loop with table look-ups in local memory).  Column layout: see Shape::pcol; the residual is
// column NP, column NP + 1 is zero.
template <class Shape, int PI>
__device__ __forceinline__ void static_param(const CalibP& P, const CalibN& N, const Sfx& S,
                                             double* __restrict__ rows, double* __restrict__ rowsB) {
  constexpr int kRowStep = (kMeas / 2) * kRS6;
  constexpr int kRowStepB = (kMeas / 2) * kRSB;
  const double ax = P.pax[PI][0], ay = P.pax[PI][1], az = P.pax[PI][2];
  const double a0 = S.Q[0] * ax + S.Q[3] * ay + S.Q[6] * az;  // a' = Q^T a
  const double a1 = S.Q[1] * ax + S.Q[4] * ay + S.Q[7] * az;
  const double a2 = S.Q[2] * ax + S.Q[5] * ay + S.Q[8] * az;
  double* col = rows + Shape::pcol(PI);
  if constexpr (Shape::pkind(PI) == 1) {
    double* colB = rowsB + (Shape::pcol(PI) - Shape::NTRANS);
    const double dx = S.s[0] - P.porg[PI][0], dy = S.s[1] - P.porg[PI][1], dz = S.s[2] - P.porg[PI][2];
    const double b0 = S.Q[0] * dx + S.Q[3] * dy + S.Q[6] * dz + N.tbar[0];  // b'' = Q^T (s - o) + tbar
    const double b1 = S.Q[1] * dx + S.Q[4] * dy + S.Q[7] * dz + N.tbar[1];
    const double b2 = S.Q[2] * dx + S.Q[5] * dy + S.Q[8] * dz + N.tbar[2];
    col[0 * kRowStep] = N.R[0][0] * (a1 * b2 - a2 * b1);
    col[1 * kRowStep] = N.R[1][1] * (a2 * b0 - a0 * b2);
    col[2 * kRowStep] = N.R[2][2] * (a0 * b1 - a1 * b0);
    colB[0 * kRowStepB] = N.R[3][3] * a0 + N.R[3][4] * a1 + N.R[3][5] * a2;
    colB[1 * kRowStepB] = N.R[4][4] * a1 + N.R[4][5] * a2;
    colB[2 * kRowStepB] = N.R[5][5] * a2;
  } else {
    // translation column: the angular rows are zero and, by the column order, not even stored
    col[0 * kRowStep] = N.R[0][0] * a0;
    col[1 * kRowStep] = N.R[1][1] * a1;
    col[2 * kRowStep] = N.R[2][2] * a2;
  }
}
template <class Shape, int P0, int P1>
__device__ __forceinline__ void static_params(const CalibP& P, const CalibN& N, const Sfx& S,
                                              double* __restrict__ rows, double* __restrict__ rowsB) {
  if constexpr (P0 < P1) {
    static_param<Shape, P0>(P, N, S, rows, rowsB);
    static_params<Shape, P0 + 1, P1>(P, N, S, rows, rowsB);
  }
}
// joints J, J - 1, ..., 0 of the backward walk
template <class Shape, bool ROLE_A, int J>
__device__ __forceinline__ void static_joints(const CalibP& P, const CalibN& N, Sfx& S,
                                              const double (&qv)[RK_CALIB_MAX_JOINTS],
                                              double* __restrict__ rows, double* __restrict__ rowsB) {
  constexpr int SPLIT = Shape::SPLIT;
  if constexpr (!(ROLE_A && J < SPLIT)) {
    if (P.seg_trans[J]) {  // S_j = X_j Rot_{j+1} S_{j+1}; a pure translation at the nominal model
      S.s[0] += P.seg[J][3], S.s[1] += P.seg[J][7], S.s[2] += P.seg[J][11];
    } else {
      sfx_premul(S, P.seg[J]);
    }
    if constexpr (ROLE_A ? (J >= SPLIT) : (J < SPLIT))
      static_params<Shape, Shape::pstart(J), Shape::pstart(J + 1)>(P, N, S, rows, rowsB);
    if constexpr (!(ROLE_A && J == SPLIT)) {  // S <- Rot_j(q_j) S_j  (role A never needs Rot_SPLIT)
      const double v = __dadd_rn(__dmul_rn(qv[J], P.jdir[J]), P.joff[J]);  // model_calibration.py:557
      double sn, cs;
      fast_sincos(v, &sn, &cs);
      sfx_rot<Shape::jaxis(J)>(S, sn, cs);
    }
    if constexpr (J > 0) static_joints<Shape, ROLE_A, J - 1>(P, N, S, qv, rows, rowsB);
  }
}
template <class Shape, bool ROLE_A, int PI>
__device__ __forceinline__ void static_zero_cols(double* __restrict__ rows, double* __restrict__ rowsB) {
  constexpr int kRowStep = (kMeas / 2) * kRS6;
  constexpr int kRowStepB = (kMeas / 2) * kRSB;
  if constexpr (PI < Shape::NP) {
    if constexpr (ROLE_A ? (PI >= Shape::pstart(Shape::SPLIT)) : (PI < Shape::pstart(Shape::SPLIT))) {
#pragma unroll
      for (int r = 0; r < 3; ++r) rows[r * kRowStep + Shape::pcol(PI)] = 0.0;
      if constexpr (Shape::pkind(PI) == 1) {
#pragma unroll
        for (int r = 0; r < 3; ++r) rowsB[r * kRowStepB + Shape::pcol(PI) - Shape::NTRANS] = 0.0;
      }
    }
    static_zero_cols<Shape, ROLE_A, PI + 1>(rows, rowsB);
  }
}

template <class Shape, bool ROLE_A>
__device__ __forceinline__ double fill_rows6_static(const CalibP& P, const CalibN& N,
                                                    const double (&qv)[RK_CALIB_MAX_JOINTS],
                                                    const double (&pv)[3 * RK_CALIB_MAX_TOOLS], bool live,
                                                    double* __restrict__ rows, double* __restrict__ rowsB) {
  constexpr int kRowStep = (kMeas / 2) * kRS6;
  constexpr int kRowStepB = (kMeas / 2) * kRSB;
  constexpr int NJ = Shape::NJ, NP = Shape::NP, NB0 = Shape::NTRANS;
  if (!live) {  // past the end of the data: this role's columns are zero
    static_zero_cols<Shape, ROLE_A, 0>(rows, rowsB);
    if (!ROLE_A) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        rows[r * kRowStep + NP] = 0.0, rows[r * kRowStep + NP + 1] = 0.0;
        rowsB[r * kRowStepB + NP - NB0] = 0.0, rowsB[r * kRowStepB + NP + 1 - NB0] = 0.0;
      }
    }
    return 0.0;
  }
  Sfx S;
#pragma unroll
  for (int k = 0; k < 9; ++k) S.Q[k] = (k % 4 == 0) ? 1.0 : 0.0;
  S.s[0] = S.s[1] = S.s[2] = 0.0;
  static_joints<Shape, ROLE_A, NJ - 1>(P, N, S, qv, rows, rowsB);
  if (ROLE_A) return 0.0;
  // role B: S = Rot_0 S_0; T_end = T_base S.  Residual of tool t in the end frame:
  // rho_t = R_end^T (p_meas - p_end) - tool_t   (r = p_meas - T[:3,3], model_calibration.py:420)
  sfx_premul(S, P.Tb);
  double rho[3 * RK_CALIB_MAX_TOOLS];
  double rr = 0.0;
#pragma unroll
  for (int t = 0; t < RK_CALIB_MAX_TOOLS; ++t) {
    if (t < P.nt) {
      const double dx = pv[3 * t] - S.s[0], dy = pv[3 * t + 1] - S.s[1], dz = pv[3 * t + 2] - S.s[2];
      rho[3 * t + 0] = S.Q[0] * dx + S.Q[3] * dy + S.Q[6] * dz - P.tool[t][0];
      rho[3 * t + 1] = S.Q[1] * dx + S.Q[4] * dy + S.Q[7] * dz - P.tool[t][1];
      rho[3 * t + 2] = S.Q[2] * dx + S.Q[5] * dy + S.Q[8] * dz - P.tool[t][2];
      rr += rho[3 * t] * rho[3 * t] + rho[3 * t + 1] * rho[3 * t + 1] + rho[3 * t + 2] * rho[3 * t + 2];
    } else {
      rho[3 * t] = rho[3 * t + 1] = rho[3 * t + 2] = 0.0;
    }
  }
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 3 * RK_CALIB_MAX_TOOLS; ++k) acc = fma(N.Qt[r][k], rho[k], acc);
    if (r < 3) {
      rows[r * kRowStep + NP] = acc;
      rows[r * kRowStep + NP + 1] = 0.0;
    } else {
      rowsB[(r - 3) * kRowStepB + NP - NB0] = acc;
      rowsB[(r - 3) * kRowStepB + NP + 1 - NB0] = 0.0;
    }
  }
  return rr;
}

// Phase 2 for the reordered columns: block 0 = translations (zero in the angular rows), block 1 =
// rotations 0..7, block 2 = rotations 8, 9, residual, zero (+ 4 columns that exist only as the
// zero fill of the fragment).  Row groups [0, kGroups / 2) hold the linear rows y[0..2] of the
// pass's measurements (6 tiles), the other half the angular rows y[3..5] (3 tiles).
struct GramDmmaSplit {
  double acc[kTiles6][2];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int t = 0; t < kTiles6; ++t) acc[t][0] = acc[t][1] = 0.0;
  }
  __device__ __forceinline__ void accumulate(const double* __restrict__ slab, int warp, int lane) {
    const int k = lane & 3, c = lane >> 2;
    constexpr int kGroups = kMeas * 3 / 4;
#pragma unroll 2
    for (int g = warp; g < kGroups / 2; g += kMeas / kWarp) {
      const int row = (g >> 3) * 32 + (g & 7) + 8 * k;
      const double* rp = slab + row * kRS6 + c;
      const double v0 = rp[0], v1 = rp[8];
      const double v2 = c < 4 ? rp[16] : 0.0;
      GramDmma::mma(acc[0], v0, v0), GramDmma::mma(acc[1], v0, v1), GramDmma::mma(acc[2], v0, v2);
      GramDmma::mma(acc[3], v1, v1), GramDmma::mma(acc[4], v1, v2), GramDmma::mma(acc[5], v2, v2);
    }
    const double* slabB = slab + kSlabA;  // angular rows: columns 8..19 only, stride kRSB
#pragma unroll 2
    for (int g = warp; g < kGroups / 2; g += kMeas / kWarp) {
      const int row = (g >> 3) * 32 + (g & 7) + 8 * k;
      const double* rp = slabB + row * kRSB + c;
      const double v1 = rp[0];
      const double v2 = c < 4 ? rp[8] : 0.0;
      GramDmma::mma(acc[3], v1, v1), GramDmma::mma(acc[4], v1, v2), GramDmma::mma(acc[5], v2, v2);
    }
  }
};

template <class Shape>
__global__ void __launch_bounds__(kMeas, 4)
calib_normal_eq6_static_kernel(const __grid_constant__ CalibP P, const __grid_constant__ CalibN N,
                               const __grid_constant__ ColMap CM, const double* __restrict__ q,
                               const double* __restrict__ pts, int64_t m,
                               double* __restrict__ partial, unsigned int* __restrict__ counter,
                               double* __restrict__ out) {
  extern __shared__ __align__(16) double slab[];  // linear rows [(kMeas/2)*3][kRS6], angular rows [..][kRSB]
  __shared__ int s_last;
  __shared__ double s_rr[kMeas / kWarp];
  constexpr int kHalf = kMeas / 2;
  constexpr int NP = Shape::NP;
  static_assert(kSlabStatic >= (kMeas / kWarp) * kPart6, "the slab is reused for the warp partials");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool role_a = tid >= kHalf;  // warp-uniform
  const int ms = tid & (kHalf - 1);
  GramDmmaSplit G;
  G.init();
  double rr = 0.0;
  double qn[RK_CALIB_MAX_JOINTS], pn[3 * RK_CALIB_MAX_TOOLS];
  auto prefetch = [&](int64_t i) {
    const bool live = i < m;
#pragma unroll
    for (int j = 0; j < RK_CALIB_MAX_JOINTS; ++j) qn[j] = (live && j < Shape::NJ) ? __ldcs(q + i * Shape::NJ + j) : 0.0;
#pragma unroll
    for (int k = 0; k < 3 * RK_CALIB_MAX_TOOLS; ++k)
      pn[k] = (live && !role_a && k < 3 * P.nt) ? __ldcs(pts + i * 3 * P.nt + k) : 0.0;
  };
  const int64_t stride = (int64_t)gridDim.x * kHalf;
  int64_t base = (int64_t)blockIdx.x * kHalf;
  prefetch(base + ms);
  for (; base < m; base += stride) {
    double qv[RK_CALIB_MAX_JOINTS], pv[3 * RK_CALIB_MAX_TOOLS];
#pragma unroll
    for (int j = 0; j < RK_CALIB_MAX_JOINTS; ++j) qv[j] = qn[j];
#pragma unroll
    for (int k = 0; k < 3 * RK_CALIB_MAX_TOOLS; ++k) pv[k] = pn[k];
    __syncthreads();  // previous phase 2 is done with the slab
    if (role_a)
      fill_rows6_static<Shape, true>(P, N, qv, pv, base + ms < m, slab + ms * kRS6, slab + kSlabA + ms * kRSB);
    else
      rr += fill_rows6_static<Shape, false>(P, N, qv, pv, base + ms < m, slab + ms * kRS6,
                                            slab + kSlabA + ms * kRSB);
    prefetch(base + stride + ms);
    __syncthreads();
    G.accumulate(slab, warp, lane);
  }

  // warps -> CTA partial (reuse the slab): lane l holds C[l/4][2(l%4) + e] of every tile
  __syncthreads();
#pragma unroll
  for (int t = 0; t < kTiles6; ++t)
#pragma unroll
    for (int e = 0; e < 2; ++e)
      slab[(warp * kTiles6 + t) * 64 + (lane >> 2) * 8 + 2 * (lane & 3) + e] = G.acc[t][e];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rr += __shfl_xor_sync(0xffffffffu, rr, o);
  if (lane == 0) s_rr[warp] = rr;
  __syncthreads();
  double* mine = partial + (int64_t)blockIdx.x * kPartStride;
  for (int e = tid; e < kPart6; e += kMeas) {
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < kMeas / kWarp; ++w) sum += slab[w * kPart6 + e];
    mine[e] = sum;
  }
  if (tid == 0) mine[kPart6] = (s_rr[0] + s_rr[1]) + (s_rr[2] + s_rr[3]);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA: deterministic sum over CTAs (index order), then scatter into the packed format,
  // undoing the column order (slab column -> parameter)
  for (int e = tid; e <= kPart6; e += kMeas) {
    double sum = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) sum += __ldcg(partial + (int64_t)b * kPartStride + e);
    if (e == kPart6) {
      out[NP * (NP + 1) / 2 + NP] = sum;  // r^T r
      continue;
    }
    const int t = e >> 6, r = (e >> 3) & 7, c = e & 7;
    const int I = t < 3 ? 0 : (t < 5 ? 1 : 2), J = t < 3 ? t : (t < 5 ? t - 2 : 2);
    const int gi = 8 * I + r, gj = 8 * J + c;  // slab columns
    if (gi > gj || gj > NP || gi >= NP) continue;
    const int pi = CM.param_of_col[gi], pj = CM.param_of_col[gj];  // NP = the residual column
    if (gj == NP) {
      out[NP * (NP + 1) / 2 + pi] = sum;  // g
    } else {
      const int a = pi < pj ? pi : pj, b2 = pi < pj ? pj : pi;
      out[a * NP - a * (a - 1) / 2 + (b2 - a)] = sum;  // packed upper H
    }
  }
  if (tid == 0) *counter = 0;
}

// -------------------------------------------------------------------------------------------
// estimate_tool_base (model_calibration.py:296-332) without forming the rows.  The reference sums
// A^T A and A^T dp over (measurement i, tool t) with A = [I | skew(p_i)^T | R_i in block t],
// dp = p_meas(i,t) - p_i, where (R_i, p_i) is the flange pose.  Every block of those sums is a
// plain sum of a few per-measurement quantities:
//   I^T I = n nt I;  I^T skew^T = nt skew(sum p)^T;  I^T R = sum R;  skew skew^T = nt sum(|p|^2 I - p p^T);
//   skew R = sum [p x R[:,b]];  R^T R = n I (orthonormal);  and for the right-hand side
//   sum_t dp = d,  skew dp -> p x d,  R^T dp per tool.
// So a thread accumulates 33 + 3 nt numbers per measurement (~150 FMAs after the chain) instead
// of feeding 9 rows of width 16 to a Gram accumulation (1224 FMAs); the last CTA assembles the
// caller's dim x dim + dim layout.
// -------------------------------------------------------------------------------------------
// Forward chain get_pose(q) with the model's structure static (see FanucReducedShape): joint J's
// rotation is one straight-line variant and the tables are constant-bank operands.  SHAPE = void
// selects the generic walk (chain_end) for any other model.
template <class Shape, int J>
__device__ __forceinline__ void static_forward(const CalibP& P, const double (&raw)[RK_CALIB_MAX_JOINTS], Fr& F) {
  if constexpr (J < Shape::NJ) {
    const double v = __dadd_rn(__dmul_rn(raw[J], P.jdir[J]), P.joff[J]);  // model_calibration.py:557
    double sn, cs;
    fast_sincos(v, &sn, &cs);
    fr_rot<Shape::jaxis(J)>(F, sn, cs);
    if (P.seg_trans[J]) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
        F.p[i] += F.R[i * 3] * P.seg[J][3] + F.R[i * 3 + 1] * P.seg[J][7] + F.R[i * 3 + 2] * P.seg[J][11];
    } else {
      fr_mul(F, P.seg[J]);
    }
    static_forward<Shape, J + 1>(P, raw, F);
  }
}

// flange pose of measurement i: the generic walk, or the static one when SHAPE is a model shape
template <class SHAPE>
__device__ __forceinline__ void measurement_pose(const CalibP& P, const double* __restrict__ q, int64_t i, Fr& F) {
  if constexpr (std::is_void<SHAPE>::value) {
    double sj[RK_CALIB_MAX_JOINTS], cj[RK_CALIB_MAX_JOINTS];
    load_joints(P, q, i, true, sj, cj);
    chain_end(P, sj, cj, F);
  } else {
    double raw[RK_CALIB_MAX_JOINTS];
#pragma unroll
    for (int j = 0; j < RK_CALIB_MAX_JOINTS; ++j) raw[j] = j < SHAPE::NJ ? __ldcs(q + i * SHAPE::NJ + j) : 0.0;
    fr_init(F, P.Tb);
    static_forward<SHAPE, 0>(P, raw, F);
  }
}

constexpr int kTbAcc = 33 + 3 * RK_CALIB_MAX_TOOLS;

template <class SHAPE>
__global__ void __launch_bounds__(kMeas, 3)
calib_tool_base_stats_kernel(const __grid_constant__ CalibP P, const double* __restrict__ q,
                             const double* __restrict__ pts, int64_t m, double* __restrict__ partial,
                             unsigned int* __restrict__ counter, double* __restrict__ out) {
  __shared__ double red[kMeas / kWarp][kTbAcc];
  __shared__ double tot[kTbAcc];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double acc[kTbAcc];
#pragma unroll
  for (int k = 0; k < kTbAcc; ++k) acc[k] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * kMeas + tid; i < m; i += (int64_t)gridDim.x * kMeas) {
    double pv[3 * RK_CALIB_MAX_TOOLS];
#pragma unroll
    for (int k = 0; k < 3 * RK_CALIB_MAX_TOOLS; ++k) pv[k] = k < 3 * P.nt ? __ldcs(pts + i * 3 * P.nt + k) : 0.0;
    Fr F;
    measurement_pose<SHAPE>(P, q, i, F);
    const double px = F.p[0], py = F.p[1], pz = F.p[2];
    acc[0] += px, acc[1] += py, acc[2] += pz;
    acc[3] += px * px, acc[4] += px * py, acc[5] += px * pz;
    acc[6] += py * py, acc[7] += py * pz, acc[8] += pz * pz;
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[9 + k] += F.R[k];
#pragma unroll
    for (int b = 0; b < 3; ++b) {  // p x R[:, b]
      const double rx = F.R[b], ry = F.R[3 + b], rz = F.R[6 + b];
      acc[18 + 0 * 3 + b] += py * rz - pz * ry;
      acc[18 + 1 * 3 + b] += pz * rx - px * rz;
      acc[18 + 2 * 3 + b] += px * ry - py * rx;
    }
    double dx = 0.0, dy = 0.0, dz = 0.0;
#pragma unroll
    for (int t = 0; t < RK_CALIB_MAX_TOOLS; ++t) {
      if (t < P.nt) {
        const double ex = pv[3 * t] - px, ey = pv[3 * t + 1] - py, ez = pv[3 * t + 2] - pz;  // dp (:328)
        dx += ex, dy += ey, dz += ez;
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[33 + 3 * t + k] += F.R[k] * ex + F.R[3 + k] * ey + F.R[6 + k] * ez;
      }
    }
    acc[27] += dx, acc[28] += dy, acc[29] += dz;
    acc[30] += py * dz - pz * dy, acc[31] += pz * dx - px * dz, acc[32] += px * dy - py * dx;
  }
#pragma unroll
  for (int k = 0; k < kTbAcc; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (tid < kTbAcc) {
    double v = 0.0;
    for (int w = 0; w < kMeas / kWarp; ++w) v += red[w][tid];
    partial[(int64_t)blockIdx.x * kPartStride + tid] = v;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < kTbAcc) {
    double v = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(partial + (int64_t)b * kPartStride + tid);
    tot[tid] = v;
  }
  __syncthreads();
  // assemble A^T A (dim x dim, row-major) and A^T dp (dim)
  const int nt = P.nt, dim = 6 + 3 * nt;
  const double n = (double)m, fnt = (double)nt;
  for (int e = tid; e < dim * dim + dim; e += kMeas) {
    double v = 0.0;
    if (e >= dim * dim) {
      const int r = e - dim * dim;
      v = r < 3 ? tot[27 + r] : (r < 6 ? tot[30 + (r - 3)] : tot[33 + (r - 6)]);
    } else {
      int r = e / dim, c = e - r * dim;
      if (r > c) {  // symmetric
        const int t = r;
        r = c, c = t;
      }
      if (c < 3) {
        v = (r == c) ? n * fnt : 0.0;
      } else if (c < 6) {
        const int k = c - 3;
        if (r < 3) {  // nt * skew(sum p)^T = [[0, Pz, -Py], [-Pz, 0, Px], [Py, -Px, 0]]
          const double St[9] = {0.0, tot[2], -tot[1], -tot[2], 0.0, tot[0], tot[1], -tot[0], 0.0};
          v = fnt * St[r * 3 + k];
        } else {      // nt * sum(|p|^2 I - p p^T)
          const int a = r - 3;
          const double pp[9] = {tot[3], tot[4], tot[5], tot[4], tot[6], tot[7], tot[5], tot[7], tot[8]};
          v = fnt * (((a == k) ? (tot[3] + tot[6] + tot[8]) : 0.0) - pp[a * 3 + k]);
        }
      } else {
        const int k = (c - 6) % 3;
        if (r < 3) {
          v = tot[9 + r * 3 + k];          // sum R
        } else if (r < 6) {
          v = tot[18 + (r - 3) * 3 + k];   // sum skew(p) R
        } else {
          v = (r == c) ? n : 0.0;          // R^T R = I; different tools do not couple
        }
      }
    }
    out[e] = v;
  }
  if (tid == 0) *counter = 0;
}

// evaluate (model_calibration.py:480-505): sums of squares and maxima of the residuals
template <class SHAPE>
__global__ void __launch_bounds__(kMeas, 4)
calib_eval_kernel(const __grid_constant__ CalibP P, const double* __restrict__ q,
                  const double* __restrict__ pts, int64_t m, double* __restrict__ partial,
                  unsigned int* __restrict__ counter, double* __restrict__ out) {
  __shared__ double red[kMeas / kWarp][8];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double acc[8] = {0.0, 0.0, 0.0, 0.0, -1.0, -1.0, -1.0, -1.0};  // maxima start at -1 (:481)
  for (int64_t i = (int64_t)blockIdx.x * kMeas + tid; i < m; i += (int64_t)gridDim.x * kMeas) {
    double pv[3 * RK_CALIB_MAX_TOOLS];
#pragma unroll
    for (int k = 0; k < 3 * RK_CALIB_MAX_TOOLS; ++k) pv[k] = k < 3 * P.nt ? __ldcs(pts + i * 3 * P.nt + k) : 0.0;
    Fr F;
    measurement_pose<SHAPE>(P, q, i, F);
#pragma unroll
    for (int t = 0; t < RK_CALIB_MAX_TOOLS; ++t) {
      if (t >= P.nt) break;
      const double* pm = pv + 3 * t;
      double d[3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
        d[r] = fabs(pm[r] - (F.p[r] + F.R[r * 3] * P.tool[t][0] + F.R[r * 3 + 1] * P.tool[t][1] +
                             F.R[r * 3 + 2] * P.tool[t][2]));
      const double d2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      const double dist = sqrt(d2);
      acc[0] += dist * dist;  // :490 squares the rounded norm
      acc[1] += d[0] * d[0], acc[2] += d[1] * d[1], acc[3] += d[2] * d[2];
      acc[4] = fmax(acc[4], dist);
      acc[5] = fmax(acc[5], d[0]), acc[6] = fmax(acc[6], d[1]), acc[7] = fmax(acc[7], d[2]);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_xor_sync(0xffffffffu, acc[k], o);
      acc[k] = (k < 4) ? acc[k] + other : fmax(acc[k], other);
    }
    if (lane == 0) red[warp][k] = acc[k];
  }
  __syncthreads();
  if (tid < 8) {
    double v = red[0][tid];
    for (int w = 1; w < kMeas / kWarp; ++w) v = (tid < 4) ? v + red[w][tid] : fmax(v, red[w][tid]);
    partial[(int64_t)blockIdx.x * 8 + tid] = v;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < 8) {
    double v = __ldcg(partial + tid);
    for (unsigned b = 1; b < gridDim.x; ++b) {
      const double o = __ldcg(partial + (int64_t)b * 8 + tid);
      v = (tid < 4) ? v + o : fmax(v, o);
    }
    out[tid] = v;
  }
  if (tid == 0) *counter = 0;
}

constexpr size_t kWsCounterOffset = (size_t)kMaxGrid * kPartStride * sizeof(double);
constexpr size_t kWsBytes = kWsCounterOffset + 256;
// RK_CALIB_GENERIC=1 in the environment sends every model through the generic (interpreting)
// kernel: lets the tests and tools/bench compare the two on the same model
static const bool g_calib_force_generic = [] {
  const char* e = getenv("RK_CALIB_GENERIC");
  return e && e[0] == '1';
}();

}  // namespace rk

extern "C" {

int rk_calib_out_len(int p) { return p * (p + 1) / 2 + p + 1; }
size_t rk_calib_workspace_bytes(void) { return rk::kWsBytes; }

int rk_calib_normal_eq_f64(const RkCalibModel* model, const double* q, const double* pts, int64_t m,
                           double* out, void* workspace, size_t ws_bytes, void* stream) {
  using namespace rk;
  if (!model || !out || m < 0) return RK_ERR_INVALID_ARG;
  if (!workspace || ws_bytes < kWsBytes) return RK_ERR_WORKSPACE;
  if (m > 0 && (!q || !pts)) return RK_ERR_INVALID_ARG;
  CalibP P;
  int rc = make_calib(model, P, false);
  if (rc != RK_OK) return rc;
  const int ncols = P.np + 1;
  if (ncols > kW) return RK_ERR_UNSUPPORTED;
  CalibN N;
  make_calibn(P, N);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)kMeas * 3 * kRS6 * sizeof(double);
  unsigned int* counter = reinterpret_cast<unsigned int*>((char*)workspace + kWsCounterOffset);
  if (cudaMemsetAsync(counter, 0, sizeof(unsigned int), st) != cudaSuccess) return check_launch();
  if (shape_matches<FanucReducedShape>(P) && !g_calib_force_generic) {
    auto kern = calib_normal_eq6_static_kernel<FanucReducedShape>;
    const size_t smem_s = (size_t)kSlabStatic * sizeof(double);
    if (!allow_dynamic_smem(kern, smem_s)) return RK_ERR_CUDA;
    int grid = persistent_grid(kern, kMeas, smem_s, m > 0 ? 2 * m : 1);
    if (grid > kMaxGrid) grid = kMaxGrid;
    kern<<<grid, kMeas, smem_s, st>>>(P, N, make_colmap<FanucReducedShape>(), q, pts, m, (double*)workspace,
                                    counter, out);
    return check_launch();
  }
  if (!allow_dynamic_smem(calib_normal_eq6_kernel, smem)) return RK_ERR_CUDA;
  int grid = persistent_grid(calib_normal_eq6_kernel, kMeas, smem, m > 0 ? 2 * m : 1);
  if (grid > kMaxGrid) grid = kMaxGrid;
  calib_normal_eq6_kernel<<<grid, kMeas, smem, st>>>(P, N, q, pts, m, ncols, (double*)workspace, counter, out);
  return check_launch();
}

int rk_calib_tool_base_eq_f64(const RkCalibModel* model, const double* q, const double* pts, int64_t m,
                              double* out, void* workspace, size_t ws_bytes, void* stream) {
  using namespace rk;
  if (!model || !out || m < 0) return RK_ERR_INVALID_ARG;
  if (!workspace || ws_bytes < kWsBytes) return RK_ERR_WORKSPACE;
  if (m > 0 && (!q || !pts)) return RK_ERR_INVALID_ARG;
  CalibP P;
  int rc = make_calib(model, P, true);
  if (rc != RK_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* counter = reinterpret_cast<unsigned int*>((char*)workspace + kWsCounterOffset);
  if (cudaMemsetAsync(counter, 0, sizeof(unsigned int), st) != cudaSuccess) return check_launch();
  auto kern = (shape_matches<FanucReducedShape>(P) && !g_calib_force_generic)
                  ? calib_tool_base_stats_kernel<FanucReducedShape>
                  : calib_tool_base_stats_kernel<void>;
  int grid = persistent_grid(kern, kMeas, 0, m > 0 ? m : 1);
  if (grid > kMaxGrid) grid = kMaxGrid;
  kern<<<grid, kMeas, 0, st>>>(P, q, pts, m, (double*)workspace, counter, out);
  return check_launch();
}

int rk_calib_eval_f64(const RkCalibModel* model, const double* q, const double* pts, int64_t m,
                      double* out, void* workspace, size_t ws_bytes, void* stream) {
  if (!model || !out || m < 0) return RK_ERR_INVALID_ARG;
  if (!workspace || ws_bytes < rk::kWsBytes) return RK_ERR_WORKSPACE;
  if (m > 0 && (!q || !pts)) return RK_ERR_INVALID_ARG;
  rk::CalibP P;
  int rc = rk::make_calib(model, P, false);
  if (rc != RK_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* counter = reinterpret_cast<unsigned int*>((char*)workspace + rk::kWsCounterOffset);
  if (cudaMemsetAsync(counter, 0, sizeof(unsigned int), st) != cudaSuccess) return rk::check_launch();
  auto kern = (rk::shape_matches<rk::FanucReducedShape>(P) && !rk::g_calib_force_generic)
                  ? rk::calib_eval_kernel<rk::FanucReducedShape>
                  : rk::calib_eval_kernel<void>;
  int grid = rk::persistent_grid(kern, rk::kMeas, 0, m > 0 ? m : 1);
  if (grid > rk::kMaxGrid) grid = rk::kMaxGrid;
  kern<<<grid, rk::kMeas, 0, st>>>(P, q, pts, m, (double*)workspace, counter, out);
  return rk::check_launch();
}

}  // extern "C"
