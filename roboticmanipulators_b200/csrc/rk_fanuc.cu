// Fanuc165F batched kernels: K1 fk+jacobian+det, K2 Pieper IK over all branches.
// One thread owns one pose (K1) or one (pose, elbow/root branch) pair (K2); the screw
// chain is unrolled in registers.  See include/rk_b200.h for the reference citations.
#include <cmath>

#include "rk_common.cuh"

namespace rk {

template <typename S>
struct FanucP {
  S l[7], off[6], dir[6], lo[6], hi[6];
  S d, dq, eps, cdq, sdq;
  S dd, l3l3, den, rden;  // d*d, l3*l3, 2*d*l3 rounded exactly as fanuc.py:233-234 does; 1/den
  S Tb[12], Tt[12], Tbi[12], Tti[12];
  int has_base, has_tool, user_jac;
  unsigned smask[6];  // sign-bit mask of dir[j] (used when every direction is +-1)
  int std_limits;     // see fanuc_std_limits()
};

// The stock Fanuc165F constants (fanuc.py:79-86, :101-111) have a structure the IK limit test can
// use: every direction is +-1, joints 0, 3, 4, 5 have no offset, and every joint but 1 has limits
// symmetric about zero.  Then (q - off) * dir is a sign flip and lo <= u <= hi is one |.| compare,
// bit-identical to the reference's two compares.  Any other constants take the general code.
static bool fanuc_std_limits(const RkFanuc* r) {
  for (int i = 0; i < 6; ++i)
    if (r->directions[i] != 1.0 && r->directions[i] != -1.0) return false;
  for (int i : {0, 3, 4, 5})
    if (r->offsets[i] != 0.0) return false;
  for (int i : {0, 2, 3, 4, 5})
    if (r->lim_lo[i] != -r->lim_hi[i]) return false;
  return true;
}

template <typename S>
static FanucP<S> make_params(const RkFanuc* r, int flags) {
  FanucP<S> p;
  for (int i = 0; i < 7; ++i) p.l[i] = (S)r->l[i];
  for (int i = 0; i < 6; ++i) {
    p.off[i] = (S)r->offsets[i];
    p.dir[i] = (S)r->directions[i];
    p.lo[i] = (S)r->lim_lo[i];
    p.hi[i] = (S)r->lim_hi[i];
  }
  p.d = (S)r->d;
  p.dq = (S)r->dq;
  p.eps = (S)r->eps;
  p.dd = p.d * p.d;
  p.l3l3 = p.l[3] * p.l[3];
  p.den = (S(2) * p.d) * p.l[3];
  p.rden = S(1) / p.den;
  p.cdq = (S)cos(r->dq);
  p.sdq = (S)sin(r->dq);
  for (int i = 0; i < 12; ++i) {
    p.Tb[i] = (S)r->T_base[i];
    p.Tt[i] = (S)r->T_tool[i];
    p.Tbi[i] = (S)r->T_base_inv[i];
    p.Tti[i] = (S)r->T_tool_inv[i];
  }
  p.has_base = r->has_base;
  p.has_tool = r->has_tool;
  p.user_jac = (flags & RK_JAC_USER_ANGLES) ? 1 : 0;
  for (int i = 0; i < 6; ++i) p.smask[i] = r->directions[i] < 0 ? 0x80000000u : 0u;
  p.std_limits = fanuc_std_limits(r) ? 1 : 0;
  return p;
}

// -----------------------------------------------------------------------------------
// K1: forward kinematics + screw Jacobian + det J for one pose.
//
// Chain (robots/fanuc.py:122-125, :173):
//   T_base Tz(l0) Rz(q0) Tz(l1) Tx(l2) Ry(q1) Tx(l3) Ry(q2) Tz(l4) Tx(l5) Rx(q3) Ry(q4) Rx(q5) Tx(l6) T_tool
// Because every translation is along a frame axis, the frame columns follow from
// Givens-style updates: joints 1,2 share the axis y1 and only the sum q1+q2 rotates
// the arm plane; the wrist axes x3, y4, x5 intersect at the wrist centre O3.
// Jacobian column k = [u_k x (O_n - O_k); u_k] (utils/jacobians.py:203-233).
// det J: move the reference point to O3 (unimodular row operation): J becomes block
// lower triangular, det J = det[u_k x (O3 - O_k)]_{k<3} * det[u3 u4 u5]
//        = rho3 * l3 * (l5 s2 - l4 c2) * (-s4).
// -----------------------------------------------------------------------------------
// Results leave through `out` (out.T(k, v), out.J(k, v), out.det(v)) as soon as they exist, so
// the SOA kernel can turn each into a store immediately and keep register pressure low.
template <typename S, typename Out>
__device__ __forceinline__ void fanuc_fk_jac_emit(const FanucP<S>& P, const S (&q)[6], Out& out) {
  using M = Sc<S>;
  S c[6], s[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    S qm = M::add(M::mul(q[k], P.dir[k]), P.off[k]);  // fanuc.py:163
    M::sincos(qm, &s[k], &c[k]);
  }
  const S l0 = P.l[0], l1 = P.l[1], l2 = P.l[2], l3 = P.l[3], l4 = P.l[4], l5 = P.l[5], l6 = P.l[6];
  const S c12 = c[1] * c[2] - s[1] * s[2];
  const S s12 = s[1] * c[2] + c[1] * s[2];
  // arm plane coordinates (rho along x1, h along z)
  const S rho2 = l2 + l3 * c[1];
  const S h1 = l0 + l1;
  const S h2 = h1 - l3 * s[1];
  const S rho3 = rho2 + l4 * s12 + l5 * c12;
  const S h3 = h2 + l4 * c12 - l5 * s12;

  const V3<S> y1 = {-s[0], c[0], S(0)};
  const V3<S> x3 = {c12 * c[0], c12 * s[0], -s12};
  const V3<S> z3 = {s12 * c[0], s12 * s[0], c12};
  const V3<S> O3 = {rho3 * c[0], rho3 * s[0], h3};
  const V3<S> y4 = lin2(c[3], y1, s[3], z3);
  const V3<S> z4 = lin2(-s[3], y1, c[3], z3);
  const V3<S> x5 = lin2(c[4], x3, -s[4], z4);
  const V3<S> z5 = lin2(s[4], x3, c[4], z4);
  const V3<S> y6 = lin2(c[5], y4, s[5], z5);
  const V3<S> z6 = lin2(-s[5], y4, c[5], z5);
  V3<S> On = axpy(l6, x5, O3);
  // R6 = [x5 y6 z6] (columns)
  S R[9] = {x5.x, y6.x, z6.x, x5.y, y6.y, z6.y, x5.z, y6.z, z6.z};
  if (P.has_tool) {
    const S* Tt = P.Tt;
    V3<S> pt = {Tt[3], Tt[7], Tt[11]};
    V3<S> dv = rot(R, pt);
    On = {On.x + dv.x, On.y + dv.y, On.z + dv.z};
    S Rn[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        Rn[i * 3 + j] = R[i * 3 + 0] * Tt[0 * 4 + j] + R[i * 3 + 1] * Tt[1 * 4 + j] + R[i * 3 + 2] * Tt[2 * 4 + j];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = Rn[i];
  }

  S dt = -s[4] * rho3 * l3 * (l5 * s[2] - l4 * c[2]);
  if (P.user_jac) {
#pragma unroll
    for (int k = 0; k < 6; ++k) dt *= P.dir[k];
  }
  out.det(dt);

  // Jacobian column k = [u x (O_n - O_k); u], rotated into the world and emitted at once so
  // that only the frame vectors stay live
  const bool hb = P.has_base != 0;
  const S* Tb = P.Tb;
  const S Rb[9] = {Tb[0], Tb[1], Tb[2], Tb[4], Tb[5], Tb[6], Tb[8], Tb[9], Tb[10]};
  auto emit_col = [&](int k, const V3<S>& u, const V3<S>& Ok) {
    V3<S> lin = cross(u, sub(On, Ok));
    V3<S> ang = u;
    if (hb) {
      lin = rot(Rb, lin);
      ang = rot(Rb, ang);
    }
    if (P.user_jac) {
      const S dk = P.dir[k];
      lin = {lin.x * dk, lin.y * dk, lin.z * dk};
      ang = {ang.x * dk, ang.y * dk, ang.z * dk};
    }
    out.J(0 * 6 + k, lin.x), out.J(1 * 6 + k, lin.y), out.J(2 * 6 + k, lin.z);
    out.J(3 * 6 + k, ang.x), out.J(4 * 6 + k, ang.y), out.J(5 * 6 + k, ang.z);
  };
  const V3<S> zax = {S(0), S(0), S(1)};
  const V3<S> O0 = {S(0), S(0), l0};
  const V3<S> O1 = {l2 * c[0], l2 * s[0], h1};
  const V3<S> O2 = {rho2 * c[0], rho2 * s[0], h2};
  emit_col(0, zax, O0);
  emit_col(1, y1, O1);
  emit_col(2, y1, O2);
  emit_col(3, x3, O3);
  emit_col(4, y4, O3);
  emit_col(5, x5, O3);

  if (hb) {
    V3<S> p = rot(Rb, On);
    On = {p.x + Tb[3], p.y + Tb[7], p.z + Tb[11]};
    S Rn[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        Rn[i * 3 + j] = Rb[i * 3 + 0] * R[0 * 3 + j] + Rb[i * 3 + 1] * R[1 * 3 + j] + Rb[i * 3 + 2] * R[2 * 3 + j];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = Rn[i];
  }
  out.T(0, R[0]), out.T(1, R[1]), out.T(2, R[2]), out.T(3, On.x);
  out.T(4, R[3]), out.T(5, R[4]), out.T(6, R[5]), out.T(7, On.y);
  out.T(8, R[6]), out.T(9, R[7]), out.T(10, R[8]), out.T(11, On.z);
  out.T(12, S(0)), out.T(13, S(0)), out.T(14, S(0)), out.T(15, S(1));
}

template <typename S>
struct ArraySink {
  S (&Tv)[16];
  S (&Jv)[36];
  S& dv;
  __device__ __forceinline__ void T(int k, S v) { Tv[k] = v; }
  __device__ __forceinline__ void J(int k, S v) { Jv[k] = v; }
  __device__ __forceinline__ void det(S v) { dv = v; }
};

template <typename S>
__device__ __forceinline__ void fanuc_fk_jac_pose(const FanucP<S>& P, const S (&q)[6], S (&T)[16],
                                                  S (&J)[36], S& det) {
  ArraySink<S> sink{T, J, det};
  fanuc_fk_jac_emit(P, q, sink);
}

// Streams every component straight to its SOA plane.
template <typename S>
struct PlaneSink {
  S* Tp;
  S* Jp;
  S* dp;
  int64_t ld;
  __device__ __forceinline__ void T(int k, S v) {
    if (Tp) __stcs(Tp + k * ld, v);
  }
  __device__ __forceinline__ void J(int k, S v) {
    if (Jp) __stcs(Jp + k * ld, v);
  }
  __device__ __forceinline__ void det(S v) {
    if (dp) __stcs(dp, v);
  }
};

// SOA: component planes, every warp access is one contiguous 256-byte (f64) segment.
template <typename S>
__global__ void __launch_bounds__(kBlock, 4)
fanuc_fk_jac_soa_kernel(const __grid_constant__ FanucP<S> P, const S* __restrict__ q, int64_t n,
                        int64_t ld, S* __restrict__ T, S* __restrict__ J, S* __restrict__ det) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // software pipeline: the next trip's joint vector is in flight while this trip computes
  // (ncu: 55 % of warp time was long-scoreboard on these six loads)
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  S qn[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) qn[k] = (i < n) ? __ldcs(q + k * ld + i) : S(0);
  for (; i < n; i += stride) {
    S qv[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) qv[k] = qn[k];
    const int64_t inext = i + stride;
    if (inext < n) {
#pragma unroll
      for (int k = 0; k < 6; ++k) qn[k] = __ldcs(q + k * ld + inext);
    }
    PlaneSink<S> sink{T ? T + i : nullptr, J ? J + i : nullptr, det ? det + i : nullptr, ld};
    fanuc_fk_jac_emit(P, qv, sink);
  }
}

// SOA fast path of K1: a CTA walks 128-pose tiles.  The tile's six joint planes arrive by one
// 2-D TMA load (double buffered, one tile ahead); every result is written to a dense
// shared-memory tile the moment it exists and the tile leaves as three TMA stores
// (T [16][128], J [36][128], det [1][128]).  The store path then runs at the memory system's pace
// instead of through 53 LSU store instructions per thread (ncu: lg_throttle + long_scoreboard).
template <typename S>
struct FkTile {
  // 1 KB rows in both precisions: FP32 threads own two poses of a tile (half the barriers and TMA
  // instructions per pose, two independent chains per thread)
  static constexpr int kPerThread = sizeof(S) == 4 ? 2 : 1;
  static constexpr int kPoses = kBlock * kPerThread;
  S q[2][6][kPoses];
  S T[16][kPoses];
  S J[36][kPoses];
  S det[kPoses];
  uint64_t bar[2];
};

// Emits into the shared tile; the first emission of a trip (det) first makes sure the previous
// trip's TMA stores have finished reading the tile.
template <typename S>
struct TileSink {
  FkTile<S>& sm;
  int tid;     // thread index (elects the waiter)
  int col;     // pose within the tile
  bool first;  // first pose of the trip: wait for the previous trip's stores
  __device__ __forceinline__ void det(S v) {
    if (first) {
      if (tid == 0) bulk_wait_read0();
      __syncthreads();
    }
    sm.det[col] = v;
  }
  __device__ __forceinline__ void T(int k, S v) { sm.T[k][col] = v; }
  __device__ __forceinline__ void J(int k, S v) { sm.J[k][col] = v; }
};

template <typename S>
__global__ void __launch_bounds__(kBlock, 3)
fanuc_fk_jac_tma_kernel(const __grid_constant__ FanucP<S> P, const __grid_constant__ CUtensorMap tmq,
                        const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmJ,
                        const __grid_constant__ CUtensorMap tmD, int64_t ntiles, int wantT, int wantJ,
                        int wantD, int chunk) {
  using Tile = FkTile<S>;
  constexpr int TP = Tile::kPoses;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Tile& sm = *reinterpret_cast<Tile*>(smem_raw);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue_load = [&](int64_t tile, int buf) {
    mbar_expect_tx(&sm.bar[buf], 6 * TP * sizeof(S));
    tma_load_2d(&sm.q[buf][0][0], &tmq, (int)(tile * TP), 0, &sm.bar[buf]);
  };
  // A CTA owns `chunk` consecutive tiles and the grid has one CTA per chunk: the hardware block
  // scheduler hands chunks to SMs as they free up, so SMs that get more memory bandwidth process
  // more poses.  (A persistent grid with a static tile stride left the average SM idle 14 % of the
  // kernel: ncu sm__cycles_active min / avg / max = 6.7 / 9.4 / 10.9 M cycles; 5.6 -> 5.0 ms.)
  int64_t tile = (int64_t)blockIdx.x * chunk;
  const int64_t tend = (tile + chunk < ntiles) ? tile + chunk : ntiles;
  if (tid == 0 && tile < tend) issue_load(tile, 0);
  int it = 0;
  for (; tile < tend; ++tile, ++it) {
    const int buf = it & 1;
    const int64_t next = tile + 1;
    if (tid == 0 && next < tend) issue_load(next, buf ^ 1);
    mbar_wait(&sm.bar[buf], (it >> 1) & 1);
#pragma unroll
    for (int h = 0; h < Tile::kPerThread; ++h) {
      const int col = tid + h * kBlock;
      S qv[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) qv[k] = sm.q[buf][k][col];
      TileSink<S> sink{sm, tid, col, h == 0};
      fanuc_fk_jac_emit(P, qv, sink);
    }
    fence_async_smem();
    __syncthreads();  // tile complete; also everyone is done with sm.q[buf]
    if (tid == 0) {
      const int x = (int)(tile * TP);
      if (wantT) tma_store_2d(&tmT, x, 0, &sm.T[0][0]);
      if (wantJ) tma_store_2d(&tmJ, x, 0, &sm.J[0][0]);
      if (wantD) tma_store_2d(&tmD, x, 0, &sm.det[0]);
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait_all();
}

// AOS: the reference's shapes ([n][6] -> [n][4][4], [n][6][6]); each warp stages its 32
// poses through a private shared-memory tile so that global traffic is whole lines.
template <typename S>
__global__ void __launch_bounds__(kBlock)
fanuc_fk_jac_aos_kernel(const __grid_constant__ FanucP<S> P, const S* __restrict__ q, int64_t n,
                        S* __restrict__ T, S* __restrict__ J, S* __restrict__ det) {
  __shared__ S stage_all[(kBlock / kWarp) * kWarp * Stride<36>::value];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  S* stage = stage_all + warp * kWarp * Stride<36>::value;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + warp * kWarp; base < n; base += stride) {
    const int64_t rem = n - base;
    const int cnt = rem < kWarp ? (int)rem : kWarp;
    S qv[6];
    warp_load_aos<S, 6>(stage, qv, q + base * 6, cnt, lane);
    if (lane >= cnt) {
#pragma unroll
      for (int k = 0; k < 6; ++k) qv[k] = S(0);
    }
    S Tv[16], Jv[36], dv;
    fanuc_fk_jac_pose(P, qv, Tv, Jv, dv);
    if (T) warp_store_aos<S, 16>(stage, Tv, T + base * 16, cnt, lane);
    if (J) warp_store_aos<S, 36>(stage, Jv, J + base * 36, cnt, lane);
    if (det && lane < cnt) det[base + lane] = dv;
  }
}

// -----------------------------------------------------------------------------------
// K2: Pieper inverse kinematics (robots/fanuc.py:196-302).
// -----------------------------------------------------------------------------------

// T_0 = Tz(l0)^-1 T_base^-1 T T_tool^-1 Tx(l6)^-1  (fanuc.py:222-225)
template <typename S>
__device__ __forceinline__ void fanuc_ik_t0(const FanucP<S>& P, const S (&Tin)[12], S (&R0)[9],
                                            S (&p0)[3]) {
  using M = Sc<S>;
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
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) R0[i * 3 + j] = B[i * 4 + j];
  const S ml6 = -P.l[6];
  p0[0] = M::add(M::mul(B[0], ml6), B[3]);
  p0[1] = M::add(M::mul(B[4], ml6), B[7]);
  p0[2] = M::sub(M::add(M::mul(B[8], ml6), B[11]), P.l[0]);
}

// One (m, k) pair: position solve once, wrist solve once, then one solution per wrist
// flag w in wl[0..NW).  Writes user-space angles or the all-zero failure sentinel.
//
// Transcendental budget.  The reference evaluates acos, 7 atan2 and 8 sin/cos per branch.
// Every sine/cosine it takes is of an angle that was itself produced by acos/atan2 of
// known operands (or a sum of such angles), so here they are formed algebraically from
// those operands (unit vectors + angle-addition products): cos(psi) = psi_cos,
// sin(psi) = -m sqrt(1 - psi_cos^2), cos(gamma) = x'/|(x',z')|, ...; acos(c) is
// atan2(sqrt((1-c)(1+c)), c); and r13 / cos(q5) is +-hypot(r12, r13), which makes
// q4(w=+-1) = -q4(w=0).  That leaves 4 atan2 that depend only on (pose, k) -- acos(c),
// beta, gamma, q0 -- and 3 per (m,k) -- the wrist (fanuc_ik8_pose shares them across the
// branches of a pose; this routine, the general one with every special case, does not).
// Differences to the reference's literal evaluation are rounding-level (<= 1e-15), except
// within 1e-11 of the measure-zero corners where the reference itself divides by
// cos(a5 - pi) ~ 1e-5.
template <typename S, int NW>
__device__ __forceinline__ void fanuc_ik_mk(const FanucP<S>& P, const S (&R0)[9], const S (&p0)[3],
                                            int m, int k, const int (&wl)[NW], S (&qout)[NW][6],
                                            bool (&valid)[NW]) {
  using M = Sc<S>;
  const S fm = (S)m, fk = (S)k;
  const S l1 = P.l[1], l2 = P.l[2], l3 = P.l[3], d = P.d;
  const S x = p0[0], y = p0[1], z = p0[2];
  // fanuc.py:228-234, evaluated operation by operation (no contraction): these decide
  // reachability
  const S rxy = M::sqrt(M::add(M::mul(x, x), M::mul(y, y)));
  const S xp = M::sub(M::mul(fk, rxy), l2);
  const S zp = M::sub(l1, z);
  const S rg2 = M::add(M::mul(xp, xp), M::mul(zp, zp));
  const S num = M::sub(M::sub(rg2, M::mul(d, d)), M::mul(l3, l3));
  const S den = M::mul(M::mul(S(2), d), l3);
  const S psi_cos = M::div(num, den);
  const bool reach = !(M::abs(psi_cos) > S(1));  // fanuc.py:237-239 (NaN goes on, as there)
  const S c = reach ? psi_cos : S(0.5);          // unreachable lanes keep computing (discarded);
                                                 // 0.5 keeps their atan2 operands off the libm path

  const S sabs = M::sqrt_fast((S(1) - c) * (S(1) + c));  // sin(acos c) >= 0
  const S sb = -d * sabs, cb = l3 + d * c;          // beta = atan2(d sin(m psi), l3 + d cos psi)
  const S acosc = M::atan2_fast(sabs, c);   // :241  acos(psi_cos)
  const S beta = M::atan2_fast(sb, cb);     // :242-243
  const S gamma = M::atan2_fast(zp, xp);    // :244
  S q0 = M::atan2_fast(y, x);               // :246
  const S psi = -fm * acosc;                      // :241
  const S q1 = M::sub(gamma, M::mul(fm, beta));   // :247
  const S q2 = M::add(psi, P.dq);                 // :248

  // unit vectors of gamma, beta, q0
  // |(sb, cb)|^2 = d^2 + l3^2 + 2 d l3 c = x'^2 + z'^2 (cosine rule), so one rsqrt serves both
  const bool rg_ok = rg2 > S(1e-200);
  const S irg = M::rsqrt_fast(rg_ok ? rg2 : S(1));
  const S cg = rg_ok ? xp * irg : S(1), sg = rg_ok ? zp * irg : S(0);
  const S cbn = rg_ok ? cb * irg : S(1), sbn = rg_ok ? sb * irg : S(0);
  const S c1 = cg * cbn + fm * (sg * sbn);        // cos(q1), q1 = gamma - m beta
  const S s1 = sg * cbn - fm * (cg * sbn);
  const S sps = -fm * sabs;                        // sin(psi)
  const S c1p = c1 * c - s1 * sps;                 // cos(q1 + psi)
  const S s1p = s1 * c + c1 * sps;
  const S A = d * c1p + l2 + l3 * c1;              // :250-251
  const bool up = M::abs(A) < P.eps;               // :253-255
  if (up) q0 = S(0);
  const bool rxy_ok = rxy > S(1e-100);
  const S irxy = M::rsqrt_fast(rxy_ok ? rxy * rxy : S(1));
  const bool deg0 = up || !rxy_ok;
  const S c0 = deg0 ? S(1) : x * irxy, s0 = deg0 ? S(0) : y * irxy;
  const S ct = c1p * P.cdq - s1p * P.sdq;          // cos(q1 + q2), q2 = psi + dq
  const S st = s1p * P.cdq + c1p * P.sdq;

  // T_rot = Ry(-dq) T_012^-1 T_0 (:262-264).  Its rotation is Ry(-(q1+q2)) Rz(-q0) R0.
  const S a0 = c0 * R0[0] + s0 * R0[3], a1 = c0 * R0[1] + s0 * R0[4], a2 = c0 * R0[2] + s0 * R0[5];
  const S r11 = ct * a0 - st * R0[6];
  const S r12 = ct * a1 - st * R0[7];
  const S r13 = ct * a2 - st * R0[8];
  const S r21 = c0 * R0[3] - s0 * R0[0];
  const S r31 = st * a0 + ct * R0[6];

  const bool regular = M::abs(r11 - S(1)) > P.eps;  // :272
  const bool use13 = M::abs(r13) > P.eps;           // :279
  const S a3 = M::atan2_fast(r21, -r31);                 // :273
  const S a5 = M::atan2_fast(r12, r13);                  // :274
  const S q4p = M::atan2_fast(M::sqrt_fast(r12 * r12 + r13 * r13), r11);  // :280 with w = 0
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    S q3 = S(0), q4 = S(0), q5 = S(0);
    if (regular) {
      const S wpi = (S)wl[i] * M::pi();
      q3 = M::sub(a3, wpi);  // :276
      q5 = M::sub(a5, wpi);  // :277
      if (use13)
        q4 = (wl[i] == 0) ? q4p : -q4p;          // :280: r13 / cos(q5) = +-hypot(r12, r13)
      else
        q4 = M::atan2(r12 / M::sin(q5), r11);    // :282 (rare; NaN quirk preserved)
    }
    const S qs[6] = {q0, q1, q2, q3, q4, q5};
    bool out = false;
    S qu[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      qu[j] = M::mul(M::sub(qs[j], P.off[j]), P.dir[j]);   // :290
      out = out || (qu[j] < P.lo[j]) || (qu[j] > P.hi[j]);  // :292-297 (NaN passes)
    }
    valid[i] = reach && !out;
#pragma unroll
    for (int j = 0; j < 6; ++j) qout[i][j] = valid[i] ? qu[j] : S(0);
  }
}

// All 8 canonical branches of ONE pose in ONE thread (the SOA fast path of K2).  What depends
// only on the pose (q0, the Rz(-q0) R0 rows) is computed once, what depends only on k (acos,
// beta, gamma and their unit vectors) twice, the wrist four times.  m and k act as sign-bit flips.  The code has no
// data-dependent branch: the wrist-singular case (:272, 0.15 % of uniformly sampled poses) is a
// warp-voted select, and the measure-zero special cases of fanuc_ik_mk (arm on the axis :253-255,
// |r13| <= eps :279-282, degenerate lengths, operands outside fast_atan2_core's domain) only
// raise the returned `rare` flag; the caller then recomputes that pose with fanuc_ik_mk.
// Must be called by all 32 lanes of a warp.
// Branch index b = 4*[m==-1] + 2*[k==-1] + w; emitted through out.branch(b, ok, u[6]).
template <typename S, bool STD, typename Out>
__device__ __forceinline__ bool fanuc_ik8_pose(const FanucP<S>& P, const S (&R0)[9], const S (&p0)[3],
                                               Out& out) {
  using M = Sc<S>;
  bool rare = false;
  const S l1 = P.l[1], l2 = P.l[2], l3 = P.l[3], d = P.d;
  const S x = p0[0], y = p0[1], z = p0[2];
  // fanuc.py:228-230 (operation by operation: these decide reachability)
  const S rxy = M::sqrt(M::add(M::mul(x, x), M::mul(y, y)));
  const S zp = M::sub(l1, z);
  const S zp2 = M::mul(zp, zp);
  const S q0 = M::atan2_core(y, x, rare);  // :246
  rare = rare || !(rxy > S(1e-100));
  const S irxy = M::rcp_fast(rxy);
  const S c0 = x * irxy, s0 = y * irxy;
  // rows of Rz(-q0) R0
  const S a0 = c0 * R0[0] + s0 * R0[3], a1 = c0 * R0[1] + s0 * R0[4], a2 = c0 * R0[2] + s0 * R0[5];
  const S r21 = c0 * R0[3] - s0 * R0[0];
  // user angle u = (q - off) * dir (:290) and its limit test (:292-297; NaN passes).  STD: the
  // constants satisfy fanuc_std_limits(), so u is a sign flip and symmetric limits need one compare.
  S u0;
  bool out0;
  if (STD) {
    u0 = M::flipm(q0, P.smask[0]);
    out0 = M::abs(q0) > P.hi[0];
  } else {
    u0 = M::mul(M::sub(q0, P.off[0]), P.dir[0]);
    out0 = (u0 < P.lo[0]) || (u0 > P.hi[0]);
  }
  // The k and m loops stay rolled (ki/mi = 1 means k/m = -1, applied as sign-bit flips): fully
  // unrolled the pose is 32 KB of straight-line code, the size of the L1.5 instruction cache, and
  // measured 7 % slower; each iteration still holds three independent atan2 chains.
#pragma unroll 1
  for (int ki = 0; ki < 2; ++ki) {
    const S xp = M::sub(M::flip(rxy, ki), l2);                // k * rxy - l2
    const S rg2 = M::add(M::mul(xp, xp), zp2);
    // psi_cos = num / den (:231-234).  |RN(num/den)| > 1 <=> |num| > |den| for IEEE doubles, so
    // the reachability test (:237-239) needs no division; the quotient itself is num * (1/den)
    // with one residual correction (<= 1 ulp).
    const S num = M::sub(M::sub(rg2, P.dd), P.l3l3);
    const bool reach = !(M::abs(num) > M::abs(P.den));
    S pc = num * P.rden;
    pc = fma(fma(-pc, P.den, num), P.rden, pc);
    const S c = reach ? pc : S(0.5);                          // unreachable: computed, discarded
    const S sabs = M::sqrt_fast((S(1) - c) * (S(1) + c));     // sin(acos c) >= 0
    const S sb = -d * sabs, cb = l3 + d * c;                  // beta operands for m = +1
    const S acosc = M::atan2_core(sabs, c, rare);             // :241
    const S beta = M::atan2_core(sb, cb, rare);               // :242-243 (m = +1; odd in m)
    const S gamma = M::atan2_core(zp, xp, rare);              // :244
    rare = rare || !(rg2 > S(1e-200));
    // |(sb, cb)|^2 = d^2 + l3^2 + 2 d l3 c = x'^2 + z'^2 (cosine rule): one rsqrt serves both
    const S irg = M::rsqrt_fast(rg2);
    const S cg = xp * irg, sg = zp * irg, cbn = cb * irg, sbn = sb * irg;
    const S cc = cg * cbn, ss = sg * sbn, sc = sg * cbn, cs = cg * sbn;
#pragma unroll 1
    for (int mi = 0; mi < 2; ++mi) {
      // m = +1: psi = -acos, q1 = gamma - beta;  m = -1: psi = +acos, q1 = gamma + beta
      const S q1 = M::sub(gamma, M::flip(beta, mi));                 // :247
      const S q2 = M::sub(P.dq, M::flip(acosc, mi));                 // :248 (psi + dq)
      const S c1 = cc + M::flip(ss, mi);                             // cos(q1)
      const S s1 = sc - M::flip(cs, mi);
      const S sps = M::flip(sabs, 1 - mi);                           // sin(psi)
      const S c1p = c1 * c - s1 * sps;                               // cos(q1 + psi)
      const S s1p = s1 * c + c1 * sps;
      const S A = d * c1p + l2 + l3 * c1;                            // :250-251
      rare = rare || (M::abs(A) < P.eps);                            // :253-255
      const S ct = c1p * P.cdq - s1p * P.sdq;                        // cos(q1 + q2)
      const S st = s1p * P.cdq + c1p * P.sdq;
      // T_rot = Ry(-dq) T_012^-1 T_0 (:262-264): rotation Ry(-(q1+q2)) Rz(-q0) R0
      const S r11 = ct * a0 - st * R0[6];
      const S r12 = ct * a1 - st * R0[7];
      const S r13 = ct * a2 - st * R0[8];
      const S r31 = st * a0 + ct * R0[6];
      const bool regular = M::abs(r11 - S(1)) > P.eps;               // :272
      rare = rare || (regular && !(M::abs(r13) > P.eps));            // :279 -> :282
      S a3 = M::atan2_core(r21, -r31, rare);                         // :273
      S a5 = M::atan2_core(r12, r13, rare);                          // :274
      S q4p = M::atan2_core(M::sqrt_fast(r12 * r12 + r13 * r13), r11, rare);  // :280, w = 0
      // wrist singular (|q4| < ~4.5e-3: 0.15 % of uniformly sampled poses): q3 = q4 = q5 = 0 for
      // every w (:266-268); pi is added back so that the w = 1 subtraction below returns 0
      S wpi = M::pi();
      if (__any_sync(0xffffffffu, !regular)) {
        if (!regular) a3 = S(0), a5 = S(0), q4p = S(0), wpi = S(0);
      }
      S u1, u2;
      bool bad = !reach || out0;
      if (STD) {
        const S t2 = M::sub(q2, P.off[2]);
        u1 = M::flipm(M::sub(q1, P.off[1]), P.smask[1]);
        u2 = M::flipm(t2, P.smask[2]);
        bad = bad || (u1 < P.lo[1]) || (u1 > P.hi[1]) || (M::abs(t2) > P.hi[2]) ||
              (M::abs(q4p) > P.hi[4]);  // |q4| is the same for w = 0 and 1
      } else {
        u1 = M::mul(M::sub(q1, P.off[1]), P.dir[1]);
        u2 = M::mul(M::sub(q2, P.off[2]), P.dir[2]);
        bad = bad || (u1 < P.lo[1]) || (u1 > P.hi[1]) || (u2 < P.lo[2]) || (u2 > P.hi[2]);
      }
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const S q3 = w ? M::sub(a3, wpi) : a3;                       // :276
        const S q5 = w ? M::sub(a5, wpi) : a5;                       // :277
        const S q4 = w ? -q4p : q4p;      // :280: r13 / cos(q5) = +-hypot(r12, r13)
        S u3, u4, u5;
        bool ok;
        if (STD) {
          u3 = M::flipm(q3, P.smask[3]), u4 = M::flipm(q4, P.smask[4]), u5 = M::flipm(q5, P.smask[5]);
          ok = !(bad || (M::abs(q3) > P.hi[3]) || (M::abs(q5) > P.hi[5]));
        } else {
          u3 = M::mul(M::sub(q3, P.off[3]), P.dir[3]);
          u4 = M::mul(M::sub(q4, P.off[4]), P.dir[4]);
          u5 = M::mul(M::sub(q5, P.off[5]), P.dir[5]);
          ok = !(bad || (u3 < P.lo[3]) || (u3 > P.hi[3]) || (u4 < P.lo[4]) || (u4 > P.hi[4]) ||
                 (u5 < P.lo[5]) || (u5 > P.hi[5]));
        }
        const S u[6] = {u0, u1, u2, u3, u4, u5};
        out.branch(mi * 4 + ki * 2 + w, ok, u);  // solution or the all-zero sentinel (:239, :300)
      }
      out.pair_done(mi * 4 + ki * 2);
    }
  }
  return rare;
}

template <typename S, bool SOA>
__device__ __forceinline__ void load_T12(const S* __restrict__ T, int64_t i, int64_t ld, S (&Tin)[12]) {
  const S* p = SOA ? T + i : T + i * 16;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    Tin[k] = __ldcs(p);
    p += SOA ? ld : 1;
  }
}

// SOA fast path, thread-per-pose: one thread computes all 8 branches of its pose with
// fanuc_ik8_pose (about half the instructions per pose of the earlier 4-lanes-per-pose mapping, which
// replicated the pose- and k-level work and spent ~200 instructions per lane on uniform-register
// spills and lane pairing).  Every warp runs its own pipeline over 32-pose tiles: the tile's 12
// input planes arrive by one 2-D TMA load ([12][32] box; the next tile's load is issued as soon as
// the lanes hold the current inputs in registers; per-warp mbarrier -- no CTA barrier anywhere);
// each result is stored to its SOA plane the moment it exists (a warp store = one contiguous
// 256-byte run; `st.global.cs`), without a guard: lanes past the end of a ragged tile redo the
// last pose.
template <typename S>
struct IkPlaneSink {
  S* qp;        // q_out + pose index
  uint8_t* vp;  // valid + pose index
  int64_t ld;
  __device__ __forceinline__ void q(int b, int j, S v) { __stcs(qp + (int64_t)(b * 6 + j) * ld, v); }
  __device__ __forceinline__ void branch(int b, bool okf, const S (&u)[6]) {
    S* p = qp + (int64_t)(b * 6) * ld;  // one 64-bit multiply per branch, then a running pointer
#pragma unroll
    for (int j = 0; j < 6; ++j, p += ld) __stcs(p, okf ? u[j] : S(0));
    ok(b, okf);
  }
  __device__ __forceinline__ void ok(int b, bool f) { vp[(int64_t)b * ld] = f ? 1 : 0; }
  __device__ __forceinline__ void pair_done(int) {}
};

template <typename S>
struct alignas(128) IkPoseTile {
  static constexpr int kPoses = kWarp;
  // Two buffers, the next tile requested one iteration ahead.  (Round 1 had ONE buffer and issued
  // the next TMA load right after the lanes had "read" the current tile.  But a shared-memory load
  // that has been issued has not necessarily been performed: under a long MIO queue the TMA write
  // of the next tile landed first, and a handful of poses per million -- always the first rows of a
  // tile -- were computed from the NEXT tile's input.  Found by the conditioned FP32 parity gate.
  // Now a buffer is overwritten a whole iteration after it was read: every value read from it has
  // been consumed by arithmetic whose results were already stored.)
  S in[2][12][kPoses];
  uint64_t bar[2];
};

template <typename S, bool STD>
__global__ void __launch_bounds__(kBlock, sizeof(S) == 4 ? 6 : 4)
fanuc_ik8_pose_kernel(const __grid_constant__ FanucP<S> P, const __grid_constant__ CUtensorMap tmT,
                      S* __restrict__ q_out, uint8_t* __restrict__ valid, int64_t n, int64_t ld,
                      int64_t ntiles, int chunk) {
  using Tile = IkPoseTile<S>;
  constexpr int TP = Tile::kPoses;
  constexpr int WPB = kBlock / kWarp;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = uniform_warp_idx();
  Tile& sm = reinterpret_cast<Tile*>(smem_raw)[warp];
  if (lane == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue_load = [&](int64_t tile, int buf) {  // one lane
    mbar_expect_tx(&sm.bar[buf], 12 * TP * sizeof(S));
    tma_load_2d(&sm.in[buf][0][0], &tmT, (int)(tile * TP), 0, &sm.bar[buf]);
  };
  // A CTA owns `chunk` consecutive groups of WPB tiles, so the CTAs resident at any moment write one
  // compact window of every output plane.  (With a grid-strided tile loop a warp's successive tiles
  // were gridDim.x * 128 poses apart and the chip was writing 16 separate regions of each of the
  // 56 planes at once: 5.53 -> 5.31 ms per 2^26 poses.)
  int64_t tile = (int64_t)blockIdx.x * WPB * chunk + warp;
  const int64_t tlast = (int64_t)(blockIdx.x + 1) * WPB * chunk;
  const int64_t tend = tlast < ntiles ? tlast : ntiles;
  if (lane == 0 && tile < tend) issue_load(tile, 0);
  int it = 0;
  for (; tile < tend; tile += WPB, ++it) {
    const int buf = it & 1;
    const int64_t next = tile + WPB;
    __syncwarp();  // every lane has left iteration it - 1 (whose reads of buffer buf ^ 1 are consumed)
    if (lane == 0 && next < tend) issue_load(next, buf ^ 1);
    mbar_wait(&sm.bar[buf], (it >> 1) & 1);
    // every lane runs (fanuc_ik8_pose votes across the warp) and stores without a guard: lanes
    // past the end of a ragged last tile redo the last pose and write the same values again
    const int64_t rem = n - tile * TP;
    const int src = lane < rem ? lane : (int)rem - 1;
    S Tin[12];
#pragma unroll
    for (int c = 0; c < 12; ++c) Tin[c] = sm.in[buf][c][src];
    const int64_t i = tile * TP + src;
    S R0[9], p0[3];
    fanuc_ik_t0(P, Tin, R0, p0);
    IkPlaneSink<S> sink{q_out + i, valid + i, ld};
    if (fanuc_ik8_pose<S, STD>(P, R0, p0, sink)) {
      // measure-zero configurations: the branch-by-branch routine with all its special cases
#pragma unroll 1
      for (int mk = 0; mk < 4; ++mk) {
        const int wl[2] = {0, 1};
        S qo[2][6];
        bool ok[2];
        fanuc_ik_mk<S, 2>(P, R0, p0, (mk & 2) ? -1 : 1, (mk & 1) ? -1 : 1, wl, qo, ok);
#pragma unroll
        for (int w = 0; w < 2; ++w) {
#pragma unroll
          for (int j = 0; j < 6; ++j) sink.q(mk * 2 + w, j, qo[w][j]);
          sink.ok(mk * 2 + w, ok[w]);
        }
      }
    }
  }
}

// K1 + K2 on the same poses in one pass (the round trip of tests/test_robots.py::test_fanuc_ifk and
// fanuc_kinematics.py: FK of q, then every IK branch of that pose): the pose never comes back from
// HBM -- its 3x4 part stays in registers between the two halves -- so a pose costs
// 48 B in + 128 + 288 + 8 + 384 + 8 B out = 864 B instead of 472 + 520 = 992 B.  Same per-warp
// pipeline as fanuc_ik8_pose_kernel ([6][32] joint tile by TMA, next tile requested as soon as the
// lanes hold their joints), every result stored to its plane the moment it exists, unguarded.
template <typename S>
struct FkIkSink {
  S* Tp;
  S* Jp;
  S* dp;
  int64_t ld;
  S (&Tin)[12];
  __device__ __forceinline__ void T(int k, S v) {
    __stcs(Tp + k * ld, v);
    if (k < 12) Tin[k] = v;
  }
  __device__ __forceinline__ void J(int k, S v) { __stcs(Jp + k * ld, v); }
  __device__ __forceinline__ void det(S v) { __stcs(dp, v); }
};

template <typename S>
struct alignas(128) FkIkTile {
  static constexpr int kPoses = kWarp;
  S q[2][6][kPoses];  // two buffers, see IkPoseTile
  uint64_t bar[2];
};

template <typename S, bool STD>
__global__ void __launch_bounds__(kBlock, sizeof(S) == 4 ? 6 : 4)
fanuc_fk_jac_ik8_kernel(const __grid_constant__ FanucP<S> P, const __grid_constant__ CUtensorMap tmq,
                        S* __restrict__ T, S* __restrict__ J, S* __restrict__ det,
                        S* __restrict__ q_out, uint8_t* __restrict__ valid, int64_t n, int64_t ld,
                        int64_t ntiles, int chunk) {
  using Tile = FkIkTile<S>;
  constexpr int TP = Tile::kPoses;
  constexpr int WPB = kBlock / kWarp;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = uniform_warp_idx();
  Tile& sm = reinterpret_cast<Tile*>(smem_raw)[warp];
  if (lane == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue_load = [&](int64_t tile, int buf) {  // one lane
    mbar_expect_tx(&sm.bar[buf], 6 * TP * sizeof(S));
    tma_load_2d(&sm.q[buf][0][0], &tmq, (int)(tile * TP), 0, &sm.bar[buf]);
  };
  int64_t tile = (int64_t)blockIdx.x * WPB * chunk + warp;  // see fanuc_ik8_pose_kernel
  const int64_t tlast = (int64_t)(blockIdx.x + 1) * WPB * chunk;
  const int64_t tend = tlast < ntiles ? tlast : ntiles;
  if (lane == 0 && tile < tend) issue_load(tile, 0);
  int it = 0;
  for (; tile < tend; tile += WPB, ++it) {
    const int buf = it & 1;
    const int64_t next = tile + WPB;
    __syncwarp();
    if (lane == 0 && next < tend) issue_load(next, buf ^ 1);
    mbar_wait(&sm.bar[buf], (it >> 1) & 1);
    const int64_t rem = n - tile * TP;  // ragged last tile: the spare lanes redo the last pose
    const int src = lane < rem ? lane : (int)rem - 1;
    S qv[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) qv[c] = sm.q[buf][c][src];
    const int64_t i = tile * TP + src;
    S Tin[12];
    FkIkSink<S> fk{T + i, J + i, det + i, ld, Tin};
    fanuc_fk_jac_emit(P, qv, fk);
    S R0[9], p0[3];
    fanuc_ik_t0(P, Tin, R0, p0);
    IkPlaneSink<S> sink{q_out + i, valid + i, ld};
    if (fanuc_ik8_pose<S, STD>(P, R0, p0, sink)) {
#pragma unroll 1
      for (int mk = 0; mk < 4; ++mk) {
        const int wl[2] = {0, 1};
        S qo[2][6];
        bool ok[2];
        fanuc_ik_mk<S, 2>(P, R0, p0, (mk & 2) ? -1 : 1, (mk & 1) ? -1 : 1, wl, qo, ok);
#pragma unroll
        for (int w = 0; w < 2; ++w) {
#pragma unroll
          for (int j = 0; j < 6; ++j) sink.q(mk * 2 + w, j, qo[w][j]);
          sink.ok(mk * 2 + w, ok[w]);
        }
      }
    }
  }
}

// Pose-major arrays, SMALL batches: K1 then K2 in ONE launch, one thread per pose, every value stored
// where it belongs the moment it exists.  A pose's stores are scattered (no staging through shared
// memory), which wastes sectors the two staged AOS kernels do not -- so only batches small enough to
// be launch-bound (n <= kSmallAos poses, ~3 us of kernel) come here: it halves the launch cost of
// the reference-shaped call (Fanuc165F.fk_jac_ik_batch on [N, 6] -> [N, 4, 4] ...; bench.py
// `latency_us`).  Same per-pose code, hence the same bits as the two calls.
constexpr int64_t kSmallAos = 16384;

template <typename S>
struct AosFkSink {
  S* Tp;
  S* Jp;
  S* dp;
  S (&Tin)[12];
  __device__ __forceinline__ void T(int k, S v) {
    Tp[k] = v;
    if (k < 12) Tin[k] = v;
  }
  __device__ __forceinline__ void J(int k, S v) {
    if (Jp) Jp[k] = v;
  }
  __device__ __forceinline__ void det(S v) {
    if (dp) *dp = v;
  }
};
template <typename S>
struct AosIkDirectSink {
  S* qp;
  uint8_t* vp;
  __device__ __forceinline__ void q(int b, int j, S v) { qp[b * 6 + j] = v; }
  __device__ __forceinline__ void branch(int b, bool okf, const S (&u)[6]) {
#pragma unroll
    for (int j = 0; j < 6; ++j) qp[b * 6 + j] = okf ? u[j] : S(0);
    vp[b] = okf ? 1 : 0;
  }
  __device__ __forceinline__ void ok(int b, bool f) { vp[b] = f ? 1 : 0; }
  __device__ __forceinline__ void pair_done(int) {}
};

template <typename S, bool STD>
__global__ void __launch_bounds__(kBlock)
fanuc_fk_jac_ik8_small_aos_kernel(const __grid_constant__ FanucP<S> P, const S* __restrict__ q, int64_t n,
                                  S* __restrict__ T, S* __restrict__ J, S* __restrict__ det,
                                  S* __restrict__ q_out, uint8_t* __restrict__ valid) {
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 - (threadIdx.x & 31) >= n) return;  // whole warps only leave (fanuc_ik8_pose votes)
  const int64_t i = i0 < n ? i0 : n - 1;     // spare lanes redo the last pose and store the same values
  S qv[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) qv[k] = q[i * 6 + k];
  S Tin[12];
  AosFkSink<S> fk{T + i * 16, J ? J + i * 36 : nullptr, det ? det + i : nullptr, Tin};
  fanuc_fk_jac_emit(P, qv, fk);
  S R0[9], p0[3];
  fanuc_ik_t0(P, Tin, R0, p0);
  AosIkDirectSink<S> sink{q_out + i * 48, valid + i * 8};
  if (fanuc_ik8_pose<S, STD>(P, R0, p0, sink)) {
#pragma unroll 1
    for (int mk = 0; mk < 4; ++mk) {
      const int wl[2] = {0, 1};
      S qo[2][6];
      bool ok[2];
      fanuc_ik_mk<S, 2>(P, R0, p0, (mk & 2) ? -1 : 1, (mk & 1) ? -1 : 1, wl, qo, ok);
#pragma unroll
      for (int w = 0; w < 2; ++w) {
#pragma unroll
        for (int j = 0; j < 6; ++j) sink.q(mk * 2 + w, j, qo[w][j]);
        sink.ok(mk * 2 + w, ok[w]);
      }
    }
  }
}

// AOS twin of fanuc_ik8_pose_kernel (the reference's shapes: T [n][4][4] -> q [n][8][6],
// valid [n][8]).  A warp's 32 poses are one contiguous 4 KB run of T: it arrives by one 2-D TMA
// load of the [32 poses][16] box with the hardware swizzle (128 B rows for FP64, 64 B for FP32), so
// that lane L reads ITS row with 16-byte loads and no bank conflict.  The two solutions of an
// (m, k) pair (12 values per pose) are staged in a warp-private tile with a padded row stride
// (conflict-free 16-byte stores) and copied out cooperatively: the pair is 32 runs of 12 values at
// a 48-value pitch, walked in 16-byte pieces so that every store instruction writes whole runs
// (FP32 stages the whole pose instead, see IkPoseTileAos).
// The 8 validity bytes of a pose are packed into one 64-bit store.
template <typename S>
struct alignas(1024) IkPoseTileAos {
  static constexpr int kPoses = kWarp;
  // FP64: one (m, k) pair is staged at a time (its 96-byte runs are whole 32-byte sectors).  FP32: a
  // pair's run is 48 bytes = 1.5 sectors, and partial-sector writes make L2 fetch the sector first,
  // so the whole pose (8 x 6 values, 192 bytes) is staged and the tile leaves as one contiguous run.
  static constexpr bool kFull = sizeof(S) == 4;
  static constexpr int kRun = kFull ? 48 : 12;
  // staging row stride (elements): rows stay 16-byte aligned and 8 consecutive rows start in 8
  // different 16-byte bank groups (112 B for FP64, 208 B for FP32)
  static constexpr int kStage = kFull ? 52 : 14;
  // two input buffers, the next tile requested one iteration ahead (see IkPoseTile)
  S in[2][kPoses][16];
  S stage[kPoses][kStage];
  uint64_t bar[2];
};

template <typename S>
struct IkAosSink {
  using Tile = IkPoseTileAos<S>;
  Tile& sm;
  S* qwarp;       // q_out of the warp's first pose
  int lane, cnt;  // cnt = live poses of this tile
  unsigned long long okbits;
  static constexpr int kPer16 = 16 / sizeof(S);         // elements per 16-byte piece
  static constexpr int kPieces = Tile::kRun / kPer16;   // pieces per staged run (6 for FP64, 12 for FP32)
  __device__ __forceinline__ void branch(int b, bool okf, const S (&u)[6]) {
    const int col = Tile::kFull ? b * 6 : (b & 1) * 6;
#pragma unroll
    for (int j = 0; j < 6; ++j) sm.stage[lane][col + j] = okf ? u[j] : S(0);
    okbits |= (unsigned long long)(okf ? 1 : 0) << (8 * b);
  }
  // Copy-out: the staged runs are walked in 16-byte pieces, piece p = it * 32 + lane of run
  // p / kPieces, so one store instruction writes 512 contiguous staged bytes = whole runs.  (run,
  // piece-in-run) advance by fixed steps per iteration: the loop carries two byte offsets instead of
  // dividing.
  __device__ __forceinline__ void pair_done(int b0) {
    if (Tile::kFull && b0 != 6) return;  // (m, k) = (-1, -1) is the last pair of a pose
    __syncwarp();
    constexpr int kRowB = Tile::kStage * (int)sizeof(S);   // staging row pitch in bytes
    constexpr int kPoseB = 48 * (int)sizeof(S);            // pose pitch in q_out
    constexpr int kStepRun = kWarp / kPieces, kStepPart = kWarp % kPieces;
    int run = lane / kPieces, part = lane - run * kPieces;
    const unsigned char* sp = reinterpret_cast<const unsigned char*>(&sm.stage[0][0]) + run * kRowB + part * 16;
    unsigned char* gp = reinterpret_cast<unsigned char*>(qwarp + (Tile::kFull ? 0 : b0 * 6)) +
                        (int64_t)run * kPoseB + part * 16;
#pragma unroll
    for (int it = 0; it < kPieces; ++it) {
      // default (write-back) policy: a 128-byte line receives the runs of two different pairs a few
      // microseconds apart and should stay in L2 until it is complete
      if (run < cnt) *reinterpret_cast<uint4*>(gp) = *reinterpret_cast<const uint4*>(sp);
      run += kStepRun, part += kStepPart;
      sp += kStepRun * kRowB + kStepPart * 16, gp += kStepRun * kPoseB + kStepPart * 16;
      if (part >= kPieces) {
        part -= kPieces, run += 1;
        sp += kRowB - kPieces * 16, gp += kPoseB - kPieces * 16;
      }
    }
    __syncwarp();
  }
};

template <typename S, bool STD>
__global__ void __launch_bounds__(kBlock, sizeof(S) == 4 ? 6 : 4)
fanuc_ik8_pose_aos_kernel(const __grid_constant__ FanucP<S> P, const __grid_constant__ CUtensorMap tmT,
                          S* __restrict__ q_out, uint8_t* __restrict__ valid, int64_t n, int64_t ntiles,
                          int chunk) {
  using Tile = IkPoseTileAos<S>;
  constexpr int TP = Tile::kPoses;
  constexpr int WPB = kBlock / kWarp;
  constexpr int kPer16 = 16 / sizeof(S);
  constexpr unsigned kSwz = sizeof(S) == 8 ? 7u : 3u;  // 128-byte / 64-byte swizzle of the row
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = uniform_warp_idx();
  // the swizzle pattern is a function of the absolute shared-memory address: tiles start on 1 KB
  unsigned char* smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Tile& sm = reinterpret_cast<Tile*>(smem_al)[warp];
  if (lane == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue_load = [&](int64_t tile, int buf) {  // one lane
    mbar_expect_tx(&sm.bar[buf], 16 * TP * sizeof(S));
    tma_load_2d(&sm.in[buf][0][0], &tmT, 0, (int)(tile * TP), &sm.bar[buf]);
  };
  int64_t tile = (int64_t)blockIdx.x * WPB * chunk + warp;  // see fanuc_ik8_pose_kernel
  const int64_t tlast = (int64_t)(blockIdx.x + 1) * WPB * chunk;
  const int64_t tend = tlast < ntiles ? tlast : ntiles;
  if (lane == 0 && tile < tend) issue_load(tile, 0);
  int it = 0;
  for (; tile < tend; tile += WPB, ++it) {
    const int buf = it & 1;
    const int64_t next = tile + WPB;
    __syncwarp();
    if (lane == 0 && next < tend) issue_load(next, buf ^ 1);
    mbar_wait(&sm.bar[buf], (it >> 1) & 1);
    // lane L owns row L; 16-byte piece c of the row sits at piece c ^ (swizzle bits of the address)
    S Tin[12];
    const unsigned char* rowp = reinterpret_cast<const unsigned char*>(&sm.in[buf][lane][0]);
    const unsigned rowoff = (unsigned)((buf * TP + lane) * 16 * sizeof(S));
#pragma unroll
    for (int c = 0; c < 12 / kPer16; ++c) {
      const unsigned off = ((unsigned)c ^ ((rowoff >> 7) & kSwz)) << 4;
      const uint4 v = *reinterpret_cast<const uint4*>(rowp + off);
      S tmp[kPer16];
      *reinterpret_cast<uint4*>(tmp) = v;
#pragma unroll
      for (int e = 0; e < kPer16; ++e) Tin[c * kPer16 + e] = tmp[e];
    }
    // every lane runs (fanuc_ik8_pose votes across the warp); rows past the end were zero-filled
    const int64_t base = tile * TP;
    const int64_t i = base + lane;
    const int64_t rem = n - base;
    S R0[9], p0[3];
    fanuc_ik_t0(P, Tin, R0, p0);
    IkAosSink<S> sink{sm, q_out + base * 48, lane, rem < TP ? (int)rem : TP, 0ull};
    const bool rare = fanuc_ik8_pose<S, STD>(P, R0, p0, sink);
    if (i < n) {
      if (rare) {
        // measure-zero configurations: the branch-by-branch routine with all its special cases
        sink.okbits = 0ull;
#pragma unroll 1
        for (int mk = 0; mk < 4; ++mk) {
          const int wl[2] = {0, 1};
          S qo[2][6];
          bool ok[2];
          fanuc_ik_mk<S, 2>(P, R0, p0, (mk & 2) ? -1 : 1, (mk & 1) ? -1 : 1, wl, qo, ok);
#pragma unroll
          for (int w = 0; w < 2; ++w) {
#pragma unroll
            for (int j = 0; j < 6; ++j) q_out[i * 48 + (mk * 2 + w) * 6 + j] = qo[w][j];
            sink.okbits |= (unsigned long long)(ok[w] ? 1 : 0) << (8 * (mk * 2 + w));
          }
        }
      }
      *reinterpret_cast<unsigned long long*>(valid + i * 8) = sink.okbits;
    }
  }
}

// Arbitrary branch list: one thread per pose.  The host groups the requested branches by (m, k);
// a group's position and wrist solve is done once (fanuc_ik_mk) for the wrist flags it needs --
// NW = 1: every group asks for a single w (e.g. one branch per pose); NW = 3: w = 0, 1, -1 are all
// evaluated and each requested branch picks its own.  Every store is a coalesced plane access in
// SOA (the earlier thread-per-(pose, branch) mapping scattered 8-byte pieces over nb planes:
// 12.4 ms per 2^24 poses for the reference's 12-branch list, now ~2 ms).
struct BranchList {
  int nb, ngroups;
  int8_t gm[4], gk[4];   // distinct (m, k) pairs in order of first appearance
  int8_t gw[4];          // NW = 1: the group's single w
  int8_t grp[16];        // group of branch b
  int8_t wi[16];         // NW = 3: index of branch b's w in {0, 1, -1}; NW = 1: 0
};

template <typename S, bool SOA, int NW>
__global__ void __launch_bounds__(kBlock)
fanuc_ik_list_kernel(const __grid_constant__ FanucP<S> P, const __grid_constant__ BranchList BL,
                     const S* __restrict__ T, int64_t n, int64_t ld, S* __restrict__ q_out,
                     uint8_t* __restrict__ valid) {
  const int nb = BL.nb;
  // AOS ([n][nb][6]): a warp's 32 poses are one contiguous run of 32 * nb * 6 values; the lanes stage
  // their rows in shared memory (odd row stride) and the warp copies the run out linearly
  extern __shared__ __align__(16) unsigned char stage_raw[];
  const int lane = threadIdx.x & 31;
  const int row = nb * 6 + 1;
  S* stage = reinterpret_cast<S*>(stage_raw) + (size_t)(threadIdx.x >> 5) * kWarp * row;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t nround = (n + kWarp - 1) / kWarp * kWarp;  // whole warps stay in the loop (AOS copy-out)
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nround; i0 += stride) {
    const int64_t i = i0 < n ? i0 : n - 1;  // lanes past the end redo the last pose
    S Tin[12];
    load_T12<S, SOA>(T, i, ld, Tin);
    S R0[9], p0[3];
    fanuc_ik_t0(P, Tin, R0, p0);
#pragma unroll 1
    for (int g = 0; g < BL.ngroups; ++g) {
      int wl[NW];
      if (NW == 1) {
        wl[0] = BL.gw[g];
      } else {
        wl[0] = 0, wl[NW > 1 ? 1 : 0] = 1, wl[NW > 2 ? 2 : 0] = -1;
      }
      S qo[NW][6];
      bool ok[NW];
      fanuc_ik_mk<S, NW>(P, R0, p0, BL.gm[g], BL.gk[g], wl, qo, ok);
#pragma unroll 1
      for (int b = 0; b < nb; ++b) {
        if (BL.grp[b] != g) continue;
        const int w = BL.wi[b];
        bool okb = ok[0];
        S v[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) v[j] = qo[0][j];
        if (NW > 1) {
          if (w == 1) okb = ok[NW > 1 ? 1 : 0];
          if (w == 2) okb = ok[NW > 2 ? 2 : 0];
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            if (w == 1) v[j] = qo[NW > 1 ? 1 : 0][j];
            if (w == 2) v[j] = qo[NW > 2 ? 2 : 0][j];
          }
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          if (SOA)
            __stcs(q_out + (int64_t)(b * 6 + j) * ld + i, v[j]);
          else
            stage[lane * row + b * 6 + j] = v[j];
        }
        if (SOA)
          valid[(int64_t)b * ld + i] = okb ? 1 : 0;
        else if (i0 < n)
          valid[i * nb + b] = okb ? 1 : 0;
      }
    }
    if (!SOA) {
      __syncwarp();
      const int64_t base = i0 - lane;                       // first pose of this warp pass
      const int64_t rem = n - base;
      const int total = (int)(rem < kWarp ? rem : kWarp) * (row - 1);
      S* g = q_out + base * (row - 1);
      int item = 0, c = lane;                                // element (item, c) = staged row, column
      while (c >= row - 1) c -= row - 1, ++item;
      for (int idx = lane; idx < total; idx += kWarp) {
        g[idx] = stage[item * row + c];
        c += kWarp;
        while (c >= row - 1) c -= row - 1, ++item;
      }
      __syncwarp();
    }
  }
}

// TrajectoryGenerator.get_dq_from_dx (utils/trajectory_generator.py:161-188) for the Fanuc165F:
// dq = J(q)^-1 dx with the 6x6 Jacobian of K1 (the twist dx = [v; omega]); the Jacobian never
// leaves the registers, the system is solved in place with partial pivoting.
template <typename S, bool SOA>
__global__ void __launch_bounds__(kBlock, 5)
fanuc_dq_from_dx_kernel(const __grid_constant__ FanucP<S> P, const S* __restrict__ q, const S* __restrict__ dx,
                        int64_t n, int64_t ld, S* __restrict__ dq, uint8_t* __restrict__ ok) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    S qv[6], b[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      qv[k] = SOA ? __ldcs(q + k * ld + i) : q[i * 6 + k];
      b[k] = SOA ? __ldcs(dx + k * ld + i) : dx[i * 6 + k];
    }
    S Tv[16], Jv[36], dv;
    fanuc_fk_jac_pose(P, qv, Tv, Jv, dv);
    S A[6][6];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) A[r][c] = Jv[r * 6 + c];
    const bool good = gauss_solve<S, 6>(A, b);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      if (SOA) __stcs(dq + k * ld + i, b[k]);
      else dq[i * 6 + k] = b[k];
    }
    if (ok) ok[i] = good ? 1 : 0;
  }
}

template <typename S>
static int launch_dq_from_dx(const RkFanuc* robot, const S* q, const S* dx, int64_t n, int64_t ld, int flags,
                             S* dq, uint8_t* ok, void* stream) {
  if (!robot || n < 0) return RK_ERR_INVALID_ARG;
  if (n == 0) return RK_OK;
  if (!q || !dx || !dq) return RK_ERR_INVALID_ARG;
  const bool soa = (flags & RK_SOA) != 0;
  if (soa && ld < n) return RK_ERR_INVALID_ARG;
  FanucP<S> P = make_params<S>(robot, flags);
  cudaStream_t st = (cudaStream_t)stream;
  if (soa) {
    int grid = stream_grid(fanuc_dq_from_dx_kernel<S, true>, kBlock, 0, n);
    fanuc_dq_from_dx_kernel<S, true><<<grid, kBlock, 0, st>>>(P, q, dx, n, ld, dq, ok);
  } else {
    int grid = stream_grid(fanuc_dq_from_dx_kernel<S, false>, kBlock, 0, n);
    fanuc_dq_from_dx_kernel<S, false><<<grid, kBlock, 0, st>>>(P, q, dx, n, ld, dq, ok);
  }
  return check_launch();
}

// -----------------------------------------------------------------------------------
// host launchers
// -----------------------------------------------------------------------------------
// Grid of the warp-per-tile kernels: a CTA owns `chunk` consecutive groups of (kBlock / kWarp) tiles.
// Large batches: `want` groups per CTA and as many CTAs as that takes (the hardware block scheduler
// balances the SMs); batches below 4 resident waves of such CTAs: one resident wave of equal chunks.
struct TileGrid {
  int grid, chunk;
};
template <typename K>
static TileGrid tile_grid(K kernel, size_t smem, int64_t ntiles, int want) {
  constexpr int WPB = kBlock / kWarp;
  const int64_t groups = (ntiles + WPB - 1) / WPB;
  const int64_t cap = persistent_grid(kernel, kBlock, smem, groups * kBlock);
  int64_t chunk = want;
  if (groups / chunk < 4 * cap) chunk = (groups + cap - 1) / cap;
  const int64_t grid = (groups + chunk - 1) / chunk;
  return {(int)grid, (int)chunk};
}
template <typename S>
static int launch_fk_jac(const RkFanuc* robot, const S* q, int64_t n, int64_t ld, int flags, S* T,
                         S* J, S* det, void* stream) {
  if (!robot || n < 0) return RK_ERR_INVALID_ARG;
  if (n == 0) return RK_OK;
  if (!q) return RK_ERR_INVALID_ARG;
  const bool soa = (flags & RK_SOA) != 0;
  if (soa && ld < n) return RK_ERR_INVALID_ARG;
  FanucP<S> P = make_params<S>(robot, flags);
  cudaStream_t st = (cudaStream_t)stream;
  if (soa) {
    // Whole 128-pose tiles go through the TMA kernel; the ragged tail (< 1 tile) through the LSU
    // kernel.  TMA stores clip a box at 16-byte granularity, so a tensor map whose extent is not a
    // multiple of 16 bytes would let the last store spill one element into the caller's padding
    // (tests/test_more_gpu.py::test_no_out_of_range_writes); full tiles never clip.
    constexpr int TP = FkTile<S>::kPoses;
    int64_t n_main = (n / TP) * TP;
    if (n_main > 0) {
      CUtensorMap tmq, tmT, tmJ, tmD;
      const uint64_t pitch = (uint64_t)ld * sizeof(S);
      // a missing output gets a (never used) alias of q so the kernel signature stays fixed
      const bool tma_ok =
          make_tmap_2d(&tmq, TmapType<S>::value, sizeof(S), q, n_main, 6, pitch, TP, 6) &&
          make_tmap_2d(&tmT, TmapType<S>::value, sizeof(S), T ? T : (S*)q, n_main, T ? 16 : 6, pitch, TP, T ? 16 : 6) &&
          make_tmap_2d(&tmJ, TmapType<S>::value, sizeof(S), J ? J : (S*)q, n_main, J ? 36 : 6, pitch, TP, J ? 36 : 6) &&
          make_tmap_2d(&tmD, TmapType<S>::value, sizeof(S), det ? det : (S*)q, n_main, 1, pitch, TP, 1);
      if (tma_ok) {
        const int64_t ntiles = n_main / TP;
        const size_t smem = sizeof(FkTile<S>);
        if (!allow_dynamic_smem(fanuc_fk_jac_tma_kernel<S>, smem)) return RK_ERR_CUDA;
        // tiles per CTA: 8 (measured 4: 4.86, 8: 4.87, 16: 4.90, 64: 5.03 ms per 2^26) when that
        // gives at least 4 waves, otherwise one resident wave of equal chunks (see stream_grid)
        const int64_t cap = persistent_grid(fanuc_fk_jac_tma_kernel<S>, kBlock, smem, ntiles * kBlock);
        const int chunk = (ntiles / 8 >= 4 * cap) ? 8 : (int)((ntiles + cap - 1) / cap);
        const int64_t grid = (ntiles + chunk - 1) / chunk;
        fanuc_fk_jac_tma_kernel<S><<<(unsigned)grid, kBlock, smem, st>>>(P, tmq, tmT, tmJ, tmD, ntiles, T != nullptr,
                                                                         J != nullptr, det != nullptr, chunk);
      } else {
        n_main = 0;  // unaligned planes: everything through the LSU kernel
      }
    }
    if (n_main < n) {
      const int64_t tail = n - n_main;
      int grid = stream_grid(fanuc_fk_jac_soa_kernel<S>, kBlock, 0, tail);
      fanuc_fk_jac_soa_kernel<S><<<grid, kBlock, 0, st>>>(P, q + n_main, tail, ld, T ? T + n_main : nullptr,
                                                          J ? J + n_main : nullptr, det ? det + n_main : nullptr);
    }
  } else {
    int grid = stream_grid(fanuc_fk_jac_aos_kernel<S>, kBlock, 0, n);
    fanuc_fk_jac_aos_kernel<S><<<grid, kBlock, 0, st>>>(P, q, n, T, J, det);
  }
  return check_launch();
}

template <typename S>
static int launch_ik(const RkFanuc* robot, const S* T, int64_t n, int64_t ld, int flags,
                     const int8_t* branches, int nb, S* q_out, uint8_t* valid, void* stream) {
  if (!robot || n < 0 || nb < 1 || nb > 16) return RK_ERR_INVALID_ARG;
  if (!branches && nb != 8) return RK_ERR_INVALID_ARG;
  if (n == 0) return RK_OK;
  if (!T || !q_out || !valid) return RK_ERR_INVALID_ARG;
  const bool soa = (flags & RK_SOA) != 0;
  if (soa && ld < n) return RK_ERR_INVALID_ARG;
  FanucP<S> P = make_params<S>(robot, flags);
  cudaStream_t st = (cudaStream_t)stream;
  bool canonical = (branches == nullptr);
  if (branches && nb == 8) {
    canonical = true;
    int b = 0;
    for (int m = 1; m >= -1; m -= 2)
      for (int k = 1; k >= -1; k -= 2)
        for (int w = 0; w <= 1; ++w, ++b)
          if (branches[b * 3] != m || branches[b * 3 + 1] != k || branches[b * 3 + 2] != w)
            canonical = false;
  }
  // the canonical 8 as an explicit list, for the general (list) kernel when the fast paths below
  // cannot be used
  static const int8_t kCanonical8[24] = {1, 1, 0, 1, 1, 1, 1, -1, 0, 1, -1, 1, -1, 1, 0, -1, 1, 1, -1, -1, 0, -1, -1, 1};
  if (canonical) {
    branches = kCanonical8;
    if (soa) {
      // thread-per-pose kernel when T satisfies the tensor-map alignment rules (16-byte base and pitch)
      constexpr int TP = IkPoseTile<S>::kPoses;
      CUtensorMap tmT;
      if (make_tmap_2d(&tmT, TmapType<S>::value, sizeof(S), T, n, 12, ld * sizeof(S), TP, 12)) {
        const int64_t ntiles = (n + TP - 1) / TP;
        const size_t smem = sizeof(IkPoseTile<S>) * (kBlock / kWarp);
        auto kern = P.std_limits ? fanuc_ik8_pose_kernel<S, true> : fanuc_ik8_pose_kernel<S, false>;
        // groups per CTA, ms per 2^26 poses: 1: 5.72, 2: 5.42, 4: 5.31
        const TileGrid tg = tile_grid(kern, smem, ntiles, 4);
        kern<<<tg.grid, kBlock, smem, st>>>(P, tmT, q_out, valid, n, ld, ntiles, tg.chunk);
        return check_launch();
      }
    } else {
      // thread-per-pose kernel when the arrays allow 16-byte accesses
      constexpr int TP = IkPoseTileAos<S>::kPoses;
      CUtensorMap tmT;
      const bool aligned = (((uintptr_t)q_out & 15) == 0) && (((uintptr_t)valid & 7) == 0);
      if (aligned && make_tmap_2d(&tmT, TmapType<S>::value, sizeof(S), T, 16, n, 16 * sizeof(S), 16, TP,
                                  sizeof(S) == 8 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)) {
        const int64_t ntiles = (n + TP - 1) / TP;
        const size_t smem = sizeof(IkPoseTileAos<S>) * (kBlock / kWarp) + 1024;
        auto kern = P.std_limits ? fanuc_ik8_pose_aos_kernel<S, true> : fanuc_ik8_pose_aos_kernel<S, false>;
        if (!allow_dynamic_smem(kern, smem)) return RK_ERR_CUDA;
        const TileGrid tg = tile_grid(kern, smem, ntiles, 4);  // 1: 6.80, 4: 6.42 ms per 2^26 poses
        kern<<<tg.grid, kBlock, smem, st>>>(P, tmT, q_out, valid, n, ntiles, tg.chunk);
        return check_launch();
      }
    }
  }
  BranchList BL = {};
  BL.nb = nb;
  bool single_w = true;
  int per_group[4] = {0, 0, 0, 0};
  for (int b = 0; b < nb; ++b) {
    int m = branches[b * 3], k = branches[b * 3 + 1], w = branches[b * 3 + 2];
    if (m != 1 && m != -1) m = 1;               // fanuc.py:210-212
    if (k != 1 && k != -1) k = 1;               // :214-216
    if (w != 1 && w != -1 && w != 0) w = 1;     // :218-220
    int g = 0;
    while (g < BL.ngroups && (BL.gm[g] != m || BL.gk[g] != k)) ++g;
    if (g == BL.ngroups) {
      BL.gm[g] = (int8_t)m, BL.gk[g] = (int8_t)k, BL.gw[g] = (int8_t)w;
      ++BL.ngroups;
    }
    BL.grp[b] = (int8_t)g;
    BL.wi[b] = (int8_t)(w == 0 ? 0 : (w == 1 ? 1 : 2));
    if (++per_group[g] > 1 && w != BL.gw[g]) single_w = false;
  }
  const size_t aos_smem = (size_t)kBlock * (nb * 6 + 1) * sizeof(S);  // AOS staging rows
  if (single_w) {
    for (int b = 0; b < nb; ++b) BL.wi[b] = 0;
    if (soa) {
      int grid = stream_grid(fanuc_ik_list_kernel<S, true, 1>, kBlock, 0, n);
      fanuc_ik_list_kernel<S, true, 1><<<grid, kBlock, 0, st>>>(P, BL, T, n, ld, q_out, valid);
    } else {
      if (!allow_dynamic_smem(fanuc_ik_list_kernel<S, false, 1>, aos_smem)) return RK_ERR_CUDA;
      int grid = stream_grid(fanuc_ik_list_kernel<S, false, 1>, kBlock, aos_smem, n);
      fanuc_ik_list_kernel<S, false, 1><<<grid, kBlock, aos_smem, st>>>(P, BL, T, n, ld, q_out, valid);
    }
  } else if (soa) {
    int grid = stream_grid(fanuc_ik_list_kernel<S, true, 3>, kBlock, 0, n);
    fanuc_ik_list_kernel<S, true, 3><<<grid, kBlock, 0, st>>>(P, BL, T, n, ld, q_out, valid);
  } else {
    if (!allow_dynamic_smem(fanuc_ik_list_kernel<S, false, 3>, aos_smem)) return RK_ERR_CUDA;
    int grid = stream_grid(fanuc_ik_list_kernel<S, false, 3>, kBlock, aos_smem, n);
    fanuc_ik_list_kernel<S, false, 3><<<grid, kBlock, aos_smem, st>>>(P, BL, T, n, ld, q_out, valid);
  }
  return check_launch();
}

// K1 then K2 (canonical 8 branches) on the same poses.  One fused launch when the layout is SOA, every
// output is requested and the planes satisfy the tensor-map alignment rules; otherwise the two
// kernels back to back on `stream` (identical results: the fused kernel runs the same per-pose code
// on the pose it has just written).
template <typename S>
static int launch_fk_jac_ik8(const RkFanuc* robot, const S* q, int64_t n, int64_t ld, int flags, S* T,
                             S* J, S* det, S* q_out, uint8_t* valid, void* stream) {
  if (!robot || n < 0) return RK_ERR_INVALID_ARG;
  if (n == 0) return RK_OK;
  if (!q || !T || !q_out || !valid) return RK_ERR_INVALID_ARG;
  const bool soa = (flags & RK_SOA) != 0;
  if (soa && ld < n) return RK_ERR_INVALID_ARG;
  if (soa && J && det) {
    constexpr int TP = FkIkTile<S>::kPoses;
    CUtensorMap tmq;
    if (make_tmap_2d(&tmq, TmapType<S>::value, sizeof(S), q, n, 6, ld * sizeof(S), TP, 6)) {
      FanucP<S> P = make_params<S>(robot, flags);
      const int64_t ntiles = (n + TP - 1) / TP;
      const size_t smem = sizeof(FkIkTile<S>) * (kBlock / kWarp);
      auto kern = P.std_limits ? fanuc_fk_jac_ik8_kernel<S, true> : fanuc_fk_jac_ik8_kernel<S, false>;
      // One group per CTA: this kernel is 94 % writes (109 planes) and runs fastest with short-lived
      // CTAs -- ms per 2^26 poses by groups per CTA: 1: 8.99, 2: 9.36, 4: 9.55, 8: 9.66 (the opposite
      // of the read-heavier IK kernel above); 3 or 5 CTAs per SM instead of 4 and padded plane
      // strides were also measured and are no better.  Round 2, with the two-buffer tile pipeline:
      // groups per CTA 1: 9.19 - 9.23, 2: 9.44, 3: 9.62, 4: 9.64 ms; CTA size 32: 9.57, 64: 9.58,
      // 128: 9.14 - 9.20, 256: 9.36 - 9.39 ms (same box, 1965 MHz).
      const TileGrid tg = tile_grid(kern, smem, ntiles, 1);
      kern<<<tg.grid, kBlock, smem, (cudaStream_t)stream>>>(P, tmq, T, J, det, q_out, valid, n, ld, ntiles,
                                                            tg.chunk);
      return check_launch();
    }
  }
  if (!soa && n <= kSmallAos) {  // launch-bound: one kernel instead of two
    FanucP<S> P = make_params<S>(robot, flags);
    auto kern = P.std_limits ? fanuc_fk_jac_ik8_small_aos_kernel<S, true> : fanuc_fk_jac_ik8_small_aos_kernel<S, false>;
    kern<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(P, q, n, T, J, det, q_out, valid);
    return check_launch();
  }
  int rc = launch_fk_jac<S>(robot, q, n, ld, flags, T, J, det, stream);
  if (rc != RK_OK) return rc;
  return launch_ik<S>(robot, T, n, ld, flags & ~RK_JAC_USER_ANGLES, nullptr, 8, q_out, valid, stream);
}

}  // namespace rk

extern "C" {

int rk_fanuc_fk_jac_f64(const RkFanuc* robot, const double* q, int64_t n, int64_t ld, int flags,
                        double* T, double* J, double* det, void* stream) {
  return rk::launch_fk_jac<double>(robot, q, n, ld, flags, T, J, det, stream);
}
int rk_fanuc_fk_jac_f32(const RkFanuc* robot, const float* q, int64_t n, int64_t ld, int flags,
                        float* T, float* J, float* det, void* stream) {
  return rk::launch_fk_jac<float>(robot, q, n, ld, flags, T, J, det, stream);
}
int rk_fanuc_ik_f64(const RkFanuc* robot, const double* T, int64_t n, int64_t ld, int flags,
                    const int8_t* branches_mkw, int nb, double* q_out, uint8_t* valid,
                    void* stream) {
  return rk::launch_ik<double>(robot, T, n, ld, flags, branches_mkw, nb, q_out, valid, stream);
}
int rk_fanuc_ik_f32(const RkFanuc* robot, const float* T, int64_t n, int64_t ld, int flags,
                    const int8_t* branches_mkw, int nb, float* q_out, uint8_t* valid,
                    void* stream) {
  return rk::launch_ik<float>(robot, T, n, ld, flags, branches_mkw, nb, q_out, valid, stream);
}

int rk_fanuc_dq_from_dx_f64(const RkFanuc* robot, const double* q, const double* dx, int64_t n, int64_t ld,
                            int flags, double* dq, uint8_t* ok, void* stream) {
  return rk::launch_dq_from_dx<double>(robot, q, dx, n, ld, flags, dq, ok, stream);
}
int rk_fanuc_dq_from_dx_f32(const RkFanuc* robot, const float* q, const float* dx, int64_t n, int64_t ld,
                            int flags, float* dq, uint8_t* ok, void* stream) {
  return rk::launch_dq_from_dx<float>(robot, q, dx, n, ld, flags, dq, ok, stream);
}

int rk_fanuc_fk_jac_ik8_f64(const RkFanuc* robot, const double* q, int64_t n, int64_t ld, int flags,
                            double* T, double* J, double* det, double* q_out, uint8_t* valid,
                            void* stream) {
  return rk::launch_fk_jac_ik8<double>(robot, q, n, ld, flags, T, J, det, q_out, valid, stream);
}
int rk_fanuc_fk_jac_ik8_f32(const RkFanuc* robot, const float* q, int64_t n, int64_t ld, int flags,
                            float* T, float* J, float* det, float* q_out, uint8_t* valid,
                            void* stream) {
  return rk::launch_fk_jac_ik8<float>(robot, q, n, ld, flags, T, J, det, q_out, valid, stream);
}

}  // extern "C"
