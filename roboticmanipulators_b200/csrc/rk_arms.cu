// RRRRobot (chain RzTzRyTxRyTx) and RRRobot (chain RzTxRzTx): FK + screw Jacobian, IK.
// Reference: robots/rrr_robot.py:96-201, robots/rr_robot.py:102-195.
#include "rk_arms.cuh"

namespace rk {

// One thread per pose; AOS goes through the warp-private staging tile.
template <typename S, int NQ, bool SOA>
__global__ void __launch_bounds__(kBlock, 4)
arm_fk_jac_kernel(const __grid_constant__ ArmPT<S> P, const S* __restrict__ q, int64_t n,
                  int64_t ld, S* __restrict__ T, S* __restrict__ J) {
  constexpr int NJ = 6 * NQ;
  constexpr int NST = Stride<NJ>::value > Stride<16>::value ? Stride<NJ>::value : Stride<16>::value;
  __shared__ S stage_all[SOA ? 1 : (kBlock / kWarp) * kWarp * NST];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + warp * kWarp; base < n; base += stride) {
    const int64_t i = base + lane;
    const int64_t rem = n - base;
    const int cnt = rem < kWarp ? (int)rem : kWarp;
    S qv[NQ], Tv[16], Jv[NJ];
    S* stage = stage_all + (SOA ? 0 : warp * kWarp * NST);
    if (SOA) {
#pragma unroll
      for (int k = 0; k < NQ; ++k) qv[k] = (lane < cnt) ? __ldcs(q + k * ld + i) : S(0);
    } else {
      warp_load_aos<S, NQ>(stage, qv, q + base * NQ, cnt, lane);
      if (lane >= cnt) {
#pragma unroll
        for (int k = 0; k < NQ; ++k) qv[k] = S(0);
      }
    }
    if constexpr (NQ == 3)
      rrr_fk_jac_pose(P, qv, Tv, Jv);
    else
      rr_fk_jac_pose(P, qv, Tv, Jv);
    if (SOA) {
      if (lane < cnt) {
        if (T) {
#pragma unroll
          for (int k = 0; k < 16; ++k) __stcs(T + k * ld + i, Tv[k]);
        }
        if (J) {
#pragma unroll
          for (int k = 0; k < NJ; ++k) __stcs(J + k * ld + i, Jv[k]);
        }
      }
    } else {
      if (T) warp_store_aos<S, 16>(stage, Tv, T + base * 16, cnt, lane);
      if (J) warp_store_aos<S, NJ>(stage, Jv, J + base * NJ, cnt, lane);
    }
  }
}

// FP32 planes, two neighbouring poses per thread: 8-byte accesses (a 4-byte plane access is 128 B per
// warp request -- the one-pose kernel sat at 0.80 of HBM with `lg_throttle` 3.7 and long_scoreboard
// 16 cycles per issue), next trip's joints in flight while this one computes.  n and ld even,
// pointers 8-byte aligned (the launcher checks; the odd tail goes through arm_fk_jac_kernel).
template <int NQ>
__global__ void __launch_bounds__(kBlock, 4)
arm_fk_jac_soa2_kernel(const __grid_constant__ ArmPT<float> P, const float* __restrict__ q, int64_t n,
                       int64_t ld, float* __restrict__ T, float* __restrict__ J) {
  constexpr int NJ = 6 * NQ;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 2;
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  float2 qn[NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k)
    qn[k] = (i < n) ? __ldcs(reinterpret_cast<const float2*>(q + k * ld + i)) : make_float2(0.f, 0.f);
  for (; i < n; i += stride) {
    float qa[NQ], qb[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k) qa[k] = qn[k].x, qb[k] = qn[k].y;
    const int64_t in = i + stride;
    if (in < n) {
#pragma unroll
      for (int k = 0; k < NQ; ++k) qn[k] = __ldcs(reinterpret_cast<const float2*>(q + k * ld + in));
    }
    float Ta[16], Ja[NJ], Tb[16], Jb[NJ];
    if constexpr (NQ == 3) {
      rrr_fk_jac_pose(P, qa, Ta, Ja);
      rrr_fk_jac_pose(P, qb, Tb, Jb);
    } else {
      rr_fk_jac_pose(P, qa, Ta, Ja);
      rr_fk_jac_pose(P, qb, Tb, Jb);
    }
    if (T) {
#pragma unroll
      for (int k = 0; k < 16; ++k) __stcs(reinterpret_cast<float2*>(T + k * ld + i), make_float2(Ta[k], Tb[k]));
    }
    if (J) {
#pragma unroll
      for (int k = 0; k < NJ; ++k) __stcs(reinterpret_cast<float2*>(J + k * ld + i), make_float2(Ja[k], Jb[k]));
    }
  }
}

struct ArmBranches {
  int8_t mk[2 * 8];
  int nb;
};

// One thread per (pose, branch).
template <typename S, int NQ, bool SOA>
__global__ void __launch_bounds__(kBlock)
arm_ik_kernel(const __grid_constant__ ArmPT<S> P, const __grid_constant__ ArmBranches BL,
              const S* __restrict__ T, int64_t n, int64_t ld, S* __restrict__ q_out,
              uint8_t* __restrict__ valid) {
  const int nb = BL.nb;
  const int64_t total = n * nb;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t tix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tix < total; tix += stride) {
    const int64_t i = tix / nb;
    const int b = (int)(tix - i * nb);
    S Tin[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) Tin[k] = SOA ? T[k * ld + i] : T[i * 16 + k];
    S p0[3];
    arm_ik_p0(P, Tin, p0);
    S qo[NQ];
    bool ok;
    if constexpr (NQ == 3)
      ok = rrr_ik_branch(P, p0, BL.mk[b * 2], BL.mk[b * 2 + 1], qo);
    else
      ok = rr_ik_branch(P, p0, BL.mk[b * 2], qo);
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      if (SOA)
        q_out[(int64_t)(b * NQ + j) * ld + i] = qo[j];
      else
        q_out[(i * nb + b) * NQ + j] = qo[j];
    }
    if (SOA)
      valid[(int64_t)b * ld + i] = ok ? 1 : 0;
    else
      valid[i * nb + b] = ok ? 1 : 0;
  }
}

// Canonical branch set (RRR: 4 = m outer, k inner; RR: 2), one thread per target: the work shared
// by the branches is done once (rrr_ik4_pose / rr_ik2_pose).  The arms' IK is position-only, so
// without a base / tool transform only the translation column of T is read (24 of 128 bytes).
template <typename S, int NQ, bool SOA>
__global__ void __launch_bounds__(kBlock, 8)
arm_ik_pose_kernel(const __grid_constant__ ArmPT<S> P, const S* __restrict__ T, int64_t n, int64_t ld,
                   S* __restrict__ q_out, uint8_t* __restrict__ valid) {
  constexpr int NB = (NQ == 3) ? 4 : 2;
  const bool full = P.has_base || P.has_tool;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // software pipeline (identity base / tool, RR arm): the next trip's translation column is in flight
  // while this trip's atan2 chains run.  Measured per 2^25 targets: RR 0.436 -> 0.415 ms (FP32 0.265
  // -> 0.237); the RRR arm sits at its 64-register cap and got slower (0.923 -> 0.965), so it loads
  // in place.
  constexpr bool PRE = NQ == 2;
  auto load3 = [&](int64_t at, S (&v)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = SOA ? __ldcs(T + (4 * k + 3) * ld + at) : T[at * 16 + 4 * k + 3];
  };
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  S pn[3] = {S(0), S(0), S(0)};
  if (PRE && !full && i < n) load3(i, pn);
  for (; i < n; i += stride) {
    S p0[3];
    if (full) {
      S Tin[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) Tin[k] = SOA ? __ldcs(T + k * ld + i) : T[i * 16 + k];
      arm_ik_p0(P, Tin, p0);
    } else {
      if constexpr (PRE) {
#pragma unroll
        for (int k = 0; k < 3; ++k) p0[k] = pn[k];
        if (i + stride < n) load3(i + stride, pn);
      } else {
        load3(i, p0);
      }
    }
    S qo[NB][NQ];
    bool ok[NB];
    if constexpr (NQ == 3)
      rrr_ik4_pose(P, p0, qo, ok);
    else
      rr_ik2_pose(P, p0, qo, ok);
    if (SOA) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
#pragma unroll
        for (int j = 0; j < NQ; ++j) __stcs(q_out + (int64_t)(b * NQ + j) * ld + i, qo[b][j]);
        valid[(int64_t)b * ld + i] = ok[b] ? 1 : 0;
      }
    } else {
      // [n][NB][NQ]: a thread's NB*NQ values are contiguous (96 / 32 bytes in FP64)
      S* qp = q_out + i * (NB * NQ);
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int j = 0; j < NQ; j += 1) qp[b * NQ + j] = qo[b][j];
#pragma unroll
      for (int b = 0; b < NB; ++b) valid[i * NB + b] = ok[b] ? 1 : 0;
    }
  }
}

template <typename S, int NQ>
static int launch_arm_fk(const RkArm* robot, const S* q, int64_t n, int64_t ld, int flags,
                         S* T, S* J, void* stream) {
  if (!robot || n < 0) return RK_ERR_INVALID_ARG;
  if (n == 0) return RK_OK;
  if (!q) return RK_ERR_INVALID_ARG;
  const bool soa = (flags & RK_SOA) != 0;
  if (soa && ld < n) return RK_ERR_INVALID_ARG;
  ArmPT<S> P = make_arm_params_t<S>(robot);
  cudaStream_t st = (cudaStream_t)stream;
  if (soa) {
    if constexpr (sizeof(S) == 4) {
      const bool pairs = n >= 2 && ld % 2 == 0 && (((uintptr_t)q | (uintptr_t)T | (uintptr_t)J) & 7) == 0;
      if (pairs) {
        const int64_t n2 = n & ~(int64_t)1;
        int grid = stream_grid(arm_fk_jac_soa2_kernel<NQ>, kBlock, 0, n2 / 2);
        arm_fk_jac_soa2_kernel<NQ><<<grid, kBlock, 0, st>>>(P, q, n2, ld, T, J);
        if (n2 < n)  // odd tail: one pose
          arm_fk_jac_kernel<S, NQ, true><<<1, kBlock, 0, st>>>(P, q + n2, 1, ld, T ? T + n2 : nullptr,
                                                               J ? J + n2 : nullptr);
        return check_launch();
      }
    }
    // (a prefetching one-pose SOA kernel measured slower in FP64: 1.596 -> 1.625 ms per 2^25, 124
    // registers; the shared kernel stays)
    int grid = stream_grid(arm_fk_jac_kernel<S, NQ, true>, kBlock, 0, n);
    arm_fk_jac_kernel<S, NQ, true><<<grid, kBlock, 0, st>>>(P, q, n, ld, T, J);
  } else {
    int grid = stream_grid(arm_fk_jac_kernel<S, NQ, false>, kBlock, 0, n);
    arm_fk_jac_kernel<S, NQ, false><<<grid, kBlock, 0, st>>>(P, q, n, ld, T, J);
  }
  return check_launch();
}

template <typename S, int NQ>
static int launch_arm_ik(const RkArm* robot, const S* T, int64_t n, int64_t ld, int flags,
                         const int8_t* branches, int nb, S* q_out, uint8_t* valid,
                         void* stream) {
  constexpr int canon = (NQ == 3) ? 4 : 2;
  if (!robot || n < 0 || nb < 1 || nb > 8) return RK_ERR_INVALID_ARG;
  if (!branches && nb != canon) return RK_ERR_INVALID_ARG;
  if (n == 0) return RK_OK;
  if (!T || !q_out || !valid) return RK_ERR_INVALID_ARG;
  const bool soa = (flags & RK_SOA) != 0;
  if (soa && ld < n) return RK_ERR_INVALID_ARG;
  ArmPT<S> P = make_arm_params_t<S>(robot);
  ArmBranches BL;
  BL.nb = nb;
  for (int b = 0; b < nb; ++b) {
    int m, k = 1;
    if (NQ == 3) {
      m = branches ? branches[b * 2] : ((b & 2) ? -1 : 1);
      k = branches ? branches[b * 2 + 1] : ((b & 1) ? -1 : 1);
    } else {
      m = branches ? branches[b] : (b ? -1 : 1);
    }
    if (m != 1 && m != -1) m = 1;   // rrr_robot.py:150-152
    if (k != 1 && k != -1) k = 1;   // :154-156
    BL.mk[b * 2] = (int8_t)m;
    BL.mk[b * 2 + 1] = (int8_t)k;
  }
  cudaStream_t st = (cudaStream_t)stream;
  bool canonical = nb == canon;
  for (int b = 0; b < nb && canonical; ++b) {
    const int m = (NQ == 3) ? ((b & 2) ? -1 : 1) : (b ? -1 : 1), k = (NQ == 3) ? ((b & 1) ? -1 : 1) : 1;
    canonical = BL.mk[b * 2] == m && BL.mk[b * 2 + 1] == k;
  }
  if (canonical) {
    if (soa) {
      int grid = stream_grid(arm_ik_pose_kernel<S, NQ, true>, kBlock, 0, n);
      arm_ik_pose_kernel<S, NQ, true><<<grid, kBlock, 0, st>>>(P, T, n, ld, q_out, valid);
    } else {
      int grid = stream_grid(arm_ik_pose_kernel<S, NQ, false>, kBlock, 0, n);
      arm_ik_pose_kernel<S, NQ, false><<<grid, kBlock, 0, st>>>(P, T, n, ld, q_out, valid);
    }
    return check_launch();
  }
  if (soa) {
    int grid = stream_grid(arm_ik_kernel<S, NQ, true>, kBlock, 0, n * nb);
    arm_ik_kernel<S, NQ, true><<<grid, kBlock, 0, st>>>(P, BL, T, n, ld, q_out, valid);
  } else {
    int grid = stream_grid(arm_ik_kernel<S, NQ, false>, kBlock, 0, n * nb);
    arm_ik_kernel<S, NQ, false><<<grid, kBlock, 0, st>>>(P, BL, T, n, ld, q_out, valid);
  }
  return check_launch();
}

}  // namespace rk

extern "C" {
#define RK_ARM_ENTRY(SFX, S)                                                                                \
  int rk_rrr_fk_jac_##SFX(const RkArm* robot, const S* q, int64_t n, int64_t ld, int flags, S* T, S* J,     \
                          void* stream) {                                                                   \
    return rk::launch_arm_fk<S, 3>(robot, q, n, ld, flags, T, J, stream);                                   \
  }                                                                                                         \
  int rk_rr_fk_jac_##SFX(const RkArm* robot, const S* q, int64_t n, int64_t ld, int flags, S* T, S* J,      \
                         void* stream) {                                                                    \
    return rk::launch_arm_fk<S, 2>(robot, q, n, ld, flags, T, J, stream);                                   \
  }                                                                                                         \
  int rk_rrr_ik_##SFX(const RkArm* robot, const S* T, int64_t n, int64_t ld, int flags,                     \
                      const int8_t* branches_mk, int nb, S* q_out, uint8_t* valid, void* stream) {          \
    return rk::launch_arm_ik<S, 3>(robot, T, n, ld, flags, branches_mk, nb, q_out, valid, stream);          \
  }                                                                                                         \
  int rk_rr_ik_##SFX(const RkArm* robot, const S* T, int64_t n, int64_t ld, int flags,                      \
                     const int8_t* branches_m, int nb, S* q_out, uint8_t* valid, void* stream) {            \
    return rk::launch_arm_ik<S, 2>(robot, T, n, ld, flags, branches_m, nb, q_out, valid, stream);           \
  }
RK_ARM_ENTRY(f64, double)
RK_ARM_ENTRY(f32, float)
#undef RK_ARM_ENTRY
}
