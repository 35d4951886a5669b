// RRRRobot (chain RzTzRyTxRyTx) and RRRobot (chain RzTxRzTx): FK + screw Jacobian, IK.
// Reference: robots/rrr_robot.py:96-201, robots/rr_robot.py:102-195.
#include "rk_arms.cuh"

namespace rk {

static ArmP make_arm(const RkArm* r) {
  ArmP p;
  for (int i = 0; i < 3; ++i) {
    p.l[i] = r->l[i];
    p.lo[i] = r->lim_lo[i];
    p.hi[i] = r->lim_hi[i];
  }
  p.eps = r->eps;
  for (int i = 0; i < 12; ++i) {
    p.Tb[i] = r->T_base[i];
    p.Tt[i] = r->T_tool[i];
    p.Tbi[i] = r->T_base_inv[i];
    p.Tti[i] = r->T_tool_inv[i];
  }
  p.has_base = r->has_base;
  p.has_tool = r->has_tool;
  return p;
}

ArmP make_arm_params(const RkArm* r) { return make_arm(r); }

// One thread per pose; AOS goes through the warp-private staging tile.
template <int NQ, bool SOA>
__global__ void __launch_bounds__(kBlock, 4)
arm_fk_jac_kernel(const __grid_constant__ ArmP P, const double* __restrict__ q, int64_t n,
                  int64_t ld, double* __restrict__ T, double* __restrict__ J) {
  constexpr int NJ = 6 * NQ;
  constexpr int NST = Stride<NJ>::value > Stride<16>::value ? Stride<NJ>::value : Stride<16>::value;
  __shared__ double stage_all[SOA ? 1 : (kBlock / kWarp) * kWarp * NST];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + warp * kWarp; base < n; base += stride) {
    const int64_t i = base + lane;
    const int64_t rem = n - base;
    const int cnt = rem < kWarp ? (int)rem : kWarp;
    double qv[NQ], Tv[16], Jv[NJ];
    double* stage = stage_all + (SOA ? 0 : warp * kWarp * NST);
    if (SOA) {
#pragma unroll
      for (int k = 0; k < NQ; ++k) qv[k] = (lane < cnt) ? __ldcs(q + k * ld + i) : 0.0;
    } else {
      warp_load_aos<double, NQ>(stage, qv, q + base * NQ, cnt, lane);
      if (lane >= cnt) {
#pragma unroll
        for (int k = 0; k < NQ; ++k) qv[k] = 0.0;
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
      if (T) warp_store_aos<double, 16>(stage, Tv, T + base * 16, cnt, lane);
      if (J) warp_store_aos<double, NJ>(stage, Jv, J + base * NJ, cnt, lane);
    }
  }
}

struct ArmBranches {
  int8_t mk[2 * 8];
  int nb;
};

// One thread per (pose, branch).
template <int NQ, bool SOA>
__global__ void __launch_bounds__(kBlock)
arm_ik_kernel(const __grid_constant__ ArmP P, const __grid_constant__ ArmBranches BL,
              const double* __restrict__ T, int64_t n, int64_t ld, double* __restrict__ q_out,
              uint8_t* __restrict__ valid) {
  const int nb = BL.nb;
  const int64_t total = n * nb;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t tix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tix < total; tix += stride) {
    const int64_t i = tix / nb;
    const int b = (int)(tix - i * nb);
    double Tin[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) Tin[k] = SOA ? T[k * ld + i] : T[i * 16 + k];
    double p0[3];
    arm_ik_p0(P, Tin, p0);
    double qo[NQ];
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
template <int NQ, bool SOA>
__global__ void __launch_bounds__(kBlock)
arm_ik_pose_kernel(const __grid_constant__ ArmP P, const double* __restrict__ T, int64_t n, int64_t ld,
                   double* __restrict__ q_out, uint8_t* __restrict__ valid) {
  constexpr int NB = (NQ == 3) ? 4 : 2;
  const bool full = P.has_base || P.has_tool;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double p0[3];
    if (full) {
      double Tin[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) Tin[k] = SOA ? __ldcs(T + k * ld + i) : T[i * 16 + k];
      arm_ik_p0(P, Tin, p0);
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) p0[k] = SOA ? __ldcs(T + (4 * k + 3) * ld + i) : T[i * 16 + 4 * k + 3];
    }
    double qo[NB][NQ];
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
      // [n][NB][NQ]: a thread's NB*NQ values are contiguous (96 / 32 bytes)
      double* qp = q_out + i * (NB * NQ);
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int j = 0; j < NQ; j += 1) qp[b * NQ + j] = qo[b][j];
#pragma unroll
      for (int b = 0; b < NB; ++b) valid[i * NB + b] = ok[b] ? 1 : 0;
    }
  }
}

template <int NQ>
static int launch_arm_fk(const RkArm* robot, const double* q, int64_t n, int64_t ld, int flags,
                         double* T, double* J, void* stream) {
  if (!robot || n < 0) return RK_ERR_INVALID_ARG;
  if (n == 0) return RK_OK;
  if (!q) return RK_ERR_INVALID_ARG;
  const bool soa = (flags & RK_SOA) != 0;
  if (soa && ld < n) return RK_ERR_INVALID_ARG;
  ArmP P = make_arm(robot);
  cudaStream_t st = (cudaStream_t)stream;
  if (soa) {
    int grid = stream_grid(arm_fk_jac_kernel<NQ, true>, kBlock, 0, n);
    arm_fk_jac_kernel<NQ, true><<<grid, kBlock, 0, st>>>(P, q, n, ld, T, J);
  } else {
    int grid = stream_grid(arm_fk_jac_kernel<NQ, false>, kBlock, 0, n);
    arm_fk_jac_kernel<NQ, false><<<grid, kBlock, 0, st>>>(P, q, n, ld, T, J);
  }
  return check_launch();
}

template <int NQ>
static int launch_arm_ik(const RkArm* robot, const double* T, int64_t n, int64_t ld, int flags,
                         const int8_t* branches, int nb, double* q_out, uint8_t* valid,
                         void* stream) {
  constexpr int canon = (NQ == 3) ? 4 : 2;
  if (!robot || n < 0 || nb < 1 || nb > 8) return RK_ERR_INVALID_ARG;
  if (!branches && nb != canon) return RK_ERR_INVALID_ARG;
  if (n == 0) return RK_OK;
  if (!T || !q_out || !valid) return RK_ERR_INVALID_ARG;
  const bool soa = (flags & RK_SOA) != 0;
  if (soa && ld < n) return RK_ERR_INVALID_ARG;
  ArmP P = make_arm(robot);
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
      int grid = stream_grid(arm_ik_pose_kernel<NQ, true>, kBlock, 0, n);
      arm_ik_pose_kernel<NQ, true><<<grid, kBlock, 0, st>>>(P, T, n, ld, q_out, valid);
    } else {
      int grid = stream_grid(arm_ik_pose_kernel<NQ, false>, kBlock, 0, n);
      arm_ik_pose_kernel<NQ, false><<<grid, kBlock, 0, st>>>(P, T, n, ld, q_out, valid);
    }
    return check_launch();
  }
  if (soa) {
    int grid = stream_grid(arm_ik_kernel<NQ, true>, kBlock, 0, n * nb);
    arm_ik_kernel<NQ, true><<<grid, kBlock, 0, st>>>(P, BL, T, n, ld, q_out, valid);
  } else {
    int grid = stream_grid(arm_ik_kernel<NQ, false>, kBlock, 0, n * nb);
    arm_ik_kernel<NQ, false><<<grid, kBlock, 0, st>>>(P, BL, T, n, ld, q_out, valid);
  }
  return check_launch();
}

}  // namespace rk

extern "C" {
int rk_rrr_fk_jac_f64(const RkArm* robot, const double* q, int64_t n, int64_t ld, int flags,
                      double* T, double* J, void* stream) {
  return rk::launch_arm_fk<3>(robot, q, n, ld, flags, T, J, stream);
}
int rk_rr_fk_jac_f64(const RkArm* robot, const double* q, int64_t n, int64_t ld, int flags,
                     double* T, double* J, void* stream) {
  return rk::launch_arm_fk<2>(robot, q, n, ld, flags, T, J, stream);
}
int rk_rrr_ik_f64(const RkArm* robot, const double* T, int64_t n, int64_t ld, int flags,
                  const int8_t* branches_mk, int nb, double* q_out, uint8_t* valid, void* stream) {
  return rk::launch_arm_ik<3>(robot, T, n, ld, flags, branches_mk, nb, q_out, valid, stream);
}
int rk_rr_ik_f64(const RkArm* robot, const double* T, int64_t n, int64_t ld, int flags,
                 const int8_t* branches_m, int nb, double* q_out, uint8_t* valid, void* stream) {
  return rk::launch_arm_ik<2>(robot, T, n, ld, flags, branches_m, nb, q_out, valid, stream);
}
}
