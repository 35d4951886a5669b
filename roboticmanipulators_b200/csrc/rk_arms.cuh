// Device functions for the 3R spatial arm (RRRRobot) and the 2R planar arm (RRRobot),
// shared by the pose kernels (rk_arms.cu) and the trajectory kernels (rk_traj.cu).
#pragma once
#include "rk_common.cuh"

namespace rk {

template <typename S>
struct ArmPT {
  S l[3], lo[3], hi[3], eps;
  S Tb[12], Tt[12], Tbi[12], Tti[12];
  int has_base, has_tool;
};
using ArmP = ArmPT<double>;

template <typename S>
inline ArmPT<S> make_arm_params_t(const RkArm* r) {
  ArmPT<S> p;
  for (int i = 0; i < 3; ++i) {
    p.l[i] = (S)r->l[i];
    p.lo[i] = (S)r->lim_lo[i];
    p.hi[i] = (S)r->lim_hi[i];
  }
  p.eps = (S)r->eps;
  for (int i = 0; i < 12; ++i) {
    p.Tb[i] = (S)r->T_base[i];
    p.Tt[i] = (S)r->T_tool[i];
    p.Tbi[i] = (S)r->T_base_inv[i];
    p.Tti[i] = (S)r->T_tool_inv[i];
  }
  p.has_base = r->has_base;
  p.has_tool = r->has_tool;
  return p;
}
inline ArmP make_arm_params(const RkArm* r) { return make_arm_params_t<double>(r); }

// Apply T_tool on the right and T_base on the left of a local pose (R columns-major
// 3x3 row-major array, origin On) and of NC Jacobian columns; write T (16) and J (6 x NC).
template <typename S, int NC>
__device__ __forceinline__ void arm_finish(const ArmPT<S>& P, S (&R)[9], V3<S> On, V3<S> (&jv)[NC],
                                           V3<S> (&jw)[NC], const V3<S> (&Ok)[NC], S (&T)[16],
                                           S (&J)[6 * NC]) {
  if (P.has_tool) {
    V3<S> pt = {P.Tt[3], P.Tt[7], P.Tt[11]};
    V3<S> dv = rot(R, pt);
    On = {On.x + dv.x, On.y + dv.y, On.z + dv.z};
    S Rn[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        Rn[i * 3 + j] = R[i * 3] * P.Tt[j] + R[i * 3 + 1] * P.Tt[4 + j] + R[i * 3 + 2] * P.Tt[8 + j];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = Rn[i];
  }
#pragma unroll
  for (int k = 0; k < NC; ++k) jv[k] = cross(jw[k], sub(On, Ok[k]));  // u x (O_n - O_k)
  if (P.has_base) {
    const S Rb[9] = {P.Tb[0], P.Tb[1], P.Tb[2], P.Tb[4], P.Tb[5], P.Tb[6], P.Tb[8], P.Tb[9], P.Tb[10]};
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      jv[k] = rot(Rb, jv[k]);
      jw[k] = rot(Rb, jw[k]);
    }
    V3<S> p = rot(Rb, On);
    On = {p.x + P.Tb[3], p.y + P.Tb[7], p.z + P.Tb[11]};
    S Rn[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        Rn[i * 3 + j] = Rb[i * 3] * R[j] + Rb[i * 3 + 1] * R[3 + j] + Rb[i * 3 + 2] * R[6 + j];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = Rn[i];
  }
  T[0] = R[0], T[1] = R[1], T[2] = R[2], T[3] = On.x;
  T[4] = R[3], T[5] = R[4], T[6] = R[5], T[7] = On.y;
  T[8] = R[6], T[9] = R[7], T[10] = R[8], T[11] = On.z;
  T[12] = S(0), T[13] = S(0), T[14] = S(0), T[15] = S(1);
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    J[0 * NC + k] = jv[k].x, J[1 * NC + k] = jv[k].y, J[2 * NC + k] = jv[k].z;
    J[3 * NC + k] = jw[k].x, J[4 * NC + k] = jw[k].y, J[5 * NC + k] = jw[k].z;
  }
}

// RRRRobot.forward_kinematics (rrr_robot.py:96-120): Rz(q0) Tz(l0) Ry(q1) Tx(l1) Ry(q2) Tx(l2)
template <typename S>
__device__ __forceinline__ void rrr_fk_jac_pose(const ArmPT<S>& P, const S (&q)[3], S (&T)[16], S (&J)[18]) {
  using M = Sc<S>;
  S s0, c0, s1, c1, s2, c2;
  M::sincos(q[0], &s0, &c0);
  M::sincos(q[1], &s1, &c1);
  M::sincos(q[2], &s2, &c2);
  const S c12 = c1 * c2 - s1 * s2, s12 = s1 * c2 + c1 * s2;
  const S l0 = P.l[0], l1 = P.l[1], l2 = P.l[2];
  const V3<S> y1 = {-s0, c0, S(0)};
  const V3<S> x3 = {c12 * c0, c12 * s0, -s12};
  const V3<S> z3 = {s12 * c0, s12 * s0, c12};
  const S rho2 = l1 * c1, h2 = l0 - l1 * s1;
  const S rho3 = rho2 + l2 * c12, h3 = h2 - l2 * s12;
  V3<S> On = {rho3 * c0, rho3 * s0, h3};
  S R[9] = {x3.x, y1.x, z3.x, x3.y, y1.y, z3.y, x3.z, y1.z, z3.z};
  V3<S> jv[3], jw[3] = {{S(0), S(0), S(1)}, y1, y1};
  const V3<S> Ok[3] = {{S(0), S(0), S(0)}, {S(0), S(0), l0}, {rho2 * c0, rho2 * s0, h2}};
  arm_finish<S, 3>(P, R, On, jv, jw, Ok, T, J);
}

// RRRobot.forward_kinematics (rr_robot.py:102-126): Rz(q0) Tx(l0) Rz(q1) Tx(l1)
template <typename S>
__device__ __forceinline__ void rr_fk_jac_pose(const ArmPT<S>& P, const S (&q)[2], S (&T)[16], S (&J)[12]) {
  using M = Sc<S>;
  S s0, c0, s1, c1;
  M::sincos(q[0], &s0, &c0);
  M::sincos(q[1], &s1, &c1);
  const S c01 = c0 * c1 - s0 * s1, s01 = s0 * c1 + c0 * s1;
  const S l0 = P.l[0], l1 = P.l[1];
  V3<S> On = {l0 * c0 + l1 * c01, l0 * s0 + l1 * s01, S(0)};
  S R[9] = {c01, -s01, S(0), s01, c01, S(0), S(0), S(0), S(1)};
  V3<S> jv[2], jw[2] = {{S(0), S(0), S(1)}, {S(0), S(0), S(1)}};
  const V3<S> Ok[2] = {{S(0), S(0), S(0)}, {l0 * c0, l0 * s0, S(0)}};
  arm_finish<S, 2>(P, R, On, jv, jw, Ok, T, J);
}

// translation of T_0 = T_base^-1 T T_tool^-1 (rrr_robot.py:159, rr_robot.py:162)
template <typename S>
__device__ __forceinline__ void arm_ik_p0(const ArmPT<S>& P, const S (&Tin)[12], S (&p0)[3]) {
  S A[12], B[12];
  if (P.has_base) {
    xf_mul(P.Tbi, Tin, A);
  } else {
#pragma unroll
    for (int i = 0; i < 12; ++i) A[i] = Tin[i];
  }
  if (P.has_tool) {
    xf_mul(A, P.Tti, B);
  } else {
#pragma unroll
    for (int i = 0; i < 12; ++i) B[i] = A[i];
  }
  p0[0] = B[3], p0[1] = B[7], p0[2] = B[11];
}

// RRRRobot.inverse_kinematics (rrr_robot.py:137-201) for the elbow sign m, both shoulder-independent
// pieces shared by the caller: everything the reference takes a sine or cosine of is an angle it
// has just produced with acos/atan2, so those are formed from the operands (as in the Fanuc IK):
// sin(q2) = m sqrt(1 - cc^2); A = l1 cos q1 + l2 cos(q1 + q2) (:181) is the radial coordinate of the
// target, i.e. x' itself up to rounding.  acos(cc) = atan2(sqrt((1-cc)(1+cc)), cc).
template <typename S>
struct RrrK {          // what depends on the target and on k only
  S xp, zp, cc, sabs, acosc, beta, gamma;
  bool reach;
};

template <typename S>
__device__ __forceinline__ RrrK<S> rrr_ik_k(const ArmPT<S>& P, const S (&p0)[3], S rxy, int k) {
  using M = Sc<S>;
  const S l0 = P.l[0], l1 = P.l[1], l2 = P.l[2];
  RrrK<S> r;
  r.xp = k > 0 ? rxy : -rxy;                                                     // :162
  r.zp = M::sub(l0, p0[2]);                                                      // :163
  const S num = M::sub(M::sub(M::add(M::mul(r.xp, r.xp), M::mul(r.zp, r.zp)), M::mul(l1, l1)),
                       M::mul(l2, l2));                                          // :165-166
  r.cc = M::div(num, M::mul(M::mul(S(2), l1), l2));                              // :167-168
  r.reach = !(M::abs(r.cc) > S(1));                                              // :171
  const S c = r.reach ? r.cc : S(0.5);   // unreachable: computed on harmless operands, discarded
  r.sabs = M::sqrt_fast((S(1) - c) * (S(1) + c));
  r.acosc = M::atan2_fast(r.sabs, c);                                            // :175 (|q2|)
  r.beta = M::atan2_fast(l2 * r.sabs, l1 + l2 * c);                              // :177-178, m = +1
  r.gamma = M::atan2_fast(r.zp, r.xp);                                           // :179
  return r;
}

// finishes branch (m, k): limits (:191-196) and the zeros-on-failure convention
template <typename S>
__device__ __forceinline__ bool rrr_ik_finish(const ArmPT<S>& P, const RrrK<S>& r, S q0atan, int m,
                                              S (&qo)[3]) {
  using M = Sc<S>;
  const S q2 = m > 0 ? r.acosc : -r.acosc;                                       // :175
  const S q1 = m > 0 ? M::sub(r.gamma, r.beta) : M::add(r.gamma, r.beta);        // :179
  const S q0 = (M::abs(r.xp) < P.eps) ? S(0) : q0atan;                           // :181-187
  const S qs[3] = {q0, q1, q2};
  bool out = !r.reach;
#pragma unroll
  for (int j = 0; j < 3; ++j) out = out || (qs[j] < P.lo[j]) || (qs[j] > P.hi[j]);
#pragma unroll
  for (int j = 0; j < 3; ++j) qo[j] = out ? S(0) : qs[j];
  return !out;
}

template <typename S>
__device__ __forceinline__ bool rrr_ik_branch(const ArmPT<S>& P, const S (&p0)[3], int m, int k, S (&qo)[3]) {
  using M = Sc<S>;
  const S x = p0[0], y = p0[1];
  const S rxy = M::sqrt(M::add(M::mul(x, x), M::mul(y, y)));                     // :162
  const RrrK<S> r = rrr_ik_k(P, p0, rxy, k);
  return rrr_ik_finish(P, r, M::atan2_fast(y, x), m, qo);
}

// Branch (m, k) = (1, 1) together with the sines / cosines of the solution it returns, formed from
// the operands of its own atan2 / acos calls (unit vectors + angle addition, as everywhere in the
// IK): trig = {c0, s0, c1, s1, c12, s12} with c12 = cos(q1 + q2).  A caller that needs the Jacobian
// AT the returned solution (get_dq_from_dx) then spends no sincos on it.  A failed branch returns
// the zero vector, hence the trig of zero.
template <typename S>
__device__ __forceinline__ bool rrr_ik11_trig(const ArmPT<S>& P, const S (&p0)[3], S (&qo)[3], S (&trig)[6]) {
  using M = Sc<S>;
  const S x = p0[0], y = p0[1];
  const S rxy = M::sqrt(M::add(M::mul(x, x), M::mul(y, y)));                     // :162
  const RrrK<S> r = rrr_ik_k(P, p0, rxy, 1);
  const bool ok = rrr_ik_finish(P, r, M::atan2_fast(y, x), 1, qo);
  const S l1 = P.l[1], l2 = P.l[2];
  const S c2 = r.reach ? r.cc : S(0.5), s2 = r.sabs;                             // q2 = +acos(cc)
  const S rg2 = r.xp * r.xp + r.zp * r.zp;
  const bool rg_ok = rg2 > S(1e-30);
  const S inv = M::rsqrt_fast(rg_ok ? rg2 : S(1));
  const S cg = rg_ok ? r.xp * inv : S(1), sg = rg_ok ? r.zp * inv : S(0);
  // |(l2 s2, l1 + l2 c2)|^2 = l1^2 + l2^2 + 2 l1 l2 c2 = x'^2 + z'^2 (cosine rule): the same 1 / |.|
  const S cb = rg_ok ? (l1 + l2 * c2) * inv : S(1), sb = rg_ok ? (l2 * s2) * inv : S(0);
  const S c1 = cg * cb + sg * sb, s1 = sg * cb - cg * sb;                        // q1 = gamma - beta
  const bool axis = M::abs(r.xp) < P.eps;                                        // :181-187 -> q0 = 0
  const bool rxy_ok = rxy > S(1e-30);
  const S irxy = M::rcp_fast(rxy_ok ? rxy : S(1));
  const S c0 = (axis || !rxy_ok) ? S(1) : x * irxy, s0 = (axis || !rxy_ok) ? S(0) : y * irxy;
  trig[0] = ok ? c0 : S(1), trig[1] = ok ? s0 : S(0);
  trig[2] = ok ? c1 : S(1), trig[3] = ok ? s1 : S(0);
  trig[4] = ok ? c1 * c2 - s1 * s2 : S(1), trig[5] = ok ? s1 * c2 + c1 * s2 : S(0);
  return ok;
}

// The canonical 4 branches (m outer, k inner) of one target in one thread: atan2(y, x) once; k only
// flips the sign of x' = k r_xy, which enters cos(q2) through x'^2 -- (-x)^2 == x^2 bit for bit -- so
// acos and beta are shared by both k as well (beta is odd in m) and only gamma = atan2(z', +-r_xy) is
// needed twice, and the second one is the supplement of the first: 4 atan2 per target instead of 16.
template <typename S>
__device__ __forceinline__ void rrr_ik4_pose(const ArmPT<S>& P, const S (&p0)[3], S (&qo)[4][3],
                                             bool (&ok)[4]) {
  using M = Sc<S>;
  const S x = p0[0], y = p0[1];
  const S rxy = M::sqrt(M::add(M::mul(x, x), M::mul(y, y)));
  const S q0 = M::atan2_fast(y, x);
  RrrK<S> r = rrr_ik_k(P, p0, rxy, 1);
#pragma unroll
  for (int ki = 0; ki < 2; ++ki) {
    if (ki) {  // k = -1: the same triangle seen from x' = -r_xy
      r.xp = -rxy;                                                               // :162
      // gamma = atan2(z', -r_xy) (:179) is the supplement of the k = +1 angle: sign(z') pi - gamma+
      // (exact for z' = +-0 and r_xy = 0, within an ulp of pi otherwise) -- the fifth atan2 goes
      r.gamma = M::sub(copysign(M::pi(), r.zp), r.gamma);
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) ok[mi * 2 + ki] = rrr_ik_finish(P, r, q0, mi ? -1 : 1, qo[mi * 2 + ki]);
  }
}

// RRRobot.inverse_kinematics (rr_robot.py:147-195): both elbow signs share acos, beta (odd in m)
// and atan2(y, x).
template <typename S>
__device__ __forceinline__ void rr_ik2_pose(const ArmPT<S>& P, const S (&p0)[3], S (&qo)[2][2],
                                            bool (&ok)[2]) {
  using M = Sc<S>;
  const S l0 = P.l[0], l1 = P.l[1];
  const S x = p0[0], y = p0[1];
  const S num = M::sub(M::sub(M::add(M::mul(x, x), M::mul(y, y)), M::mul(l0, l0)),
                       M::mul(l1, l1));                                          // :167
  const S cc = M::div(num, M::mul(M::mul(S(2), l0), l1));                        // :168-169
  const bool reach = !(M::abs(cc) > S(1));                                       // :172
  const S c = reach ? cc : S(0.5);
  const S sabs = M::sqrt_fast((S(1) - c) * (S(1) + c));
  const S acosc = M::atan2_fast(sabs, c);                                        // :176
  const S beta = M::atan2_fast(l1 * sabs, l0 + l1 * c);                          // :177-178, m = +1
  const S th = M::atan2_fast(y, x);
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    const S q1 = mi ? -acosc : acosc;
    const S q0 = mi ? M::add(th, beta) : M::sub(th, beta);                       // :179
    const S qs[2] = {q0, q1};
    bool out = !reach;
#pragma unroll
    for (int j = 0; j < 2; ++j) out = out || (qs[j] < P.lo[j]) || (qs[j] > P.hi[j]);
    qo[mi][0] = out ? S(0) : q0, qo[mi][1] = out ? S(0) : q1;
    ok[mi] = !out;
  }
}

template <typename S>
__device__ __forceinline__ bool rr_ik_branch(const ArmPT<S>& P, const S (&p0)[3], int m, S (&qo)[2]) {
  S q2[2][2];
  bool ok[2];
  rr_ik2_pose(P, p0, q2, ok);
  const int mi = m > 0 ? 0 : 1;
  qo[0] = q2[mi][0], qo[1] = q2[mi][1];
  return ok[mi];
}

}  // namespace rk
