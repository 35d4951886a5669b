// Geometric calibration reductions (utils/model_calibration.py).
//
// K5 rk_calib_normal_eq_f64: H = sum J^T J, g = sum J^T r over (measurement, tool).
// Each CTA alternates two phases on a tile of kMeas measurements, one tool at a time:
//   phase 1  thread = measurement: joint sincos, walk the reduced chain (constant token
//            runs pre-multiplied by the host), emit the 3 rows [J | r] of width <= 20 into
//            a shared-memory slab (row stride 20 doubles);
//   phase 2  thread = (4x4 tile of the upper triangle of [J|r]^T [J|r], k-slice): two
//            LDS.128 pairs feed 16 register FMAs per row.
// Accumulators stay in registers for the whole grid-stride loop; then k-slices are summed
// through shared memory, CTA partials go to the workspace, and the last CTA to finish adds
// them in index order (deterministic) and writes the packed result.
#include "rk_common.cuh"

namespace rk {

constexpr int kMeas = 128;     // measurements per CTA iteration (= threads per CTA)
constexpr int kW = 20;         // logical row width: <= 19 columns + residual, 5 blocks of 4
constexpr int kRS = 22;        // slab row stride (doubles): 16-byte aligned, 2-way banked at worst
constexpr int kMaxTiles = 15;  // upper-triangular 4x4 tiles of a 5x5 block grid
constexpr int kTileElems = kMaxTiles * 16;
constexpr int kMaxGrid = 1024;

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
#pragma unroll
  for (int j = 0; j < RK_CALIB_MAX_JOINTS; ++j) {
    s[j] = 0.0, c[j] = 1.0;
    if (j < P.nj) {
      const double raw = live ? q[i * P.nj + j] : 0.0;
      const double v = __dadd_rn(__dmul_rn(raw, P.jdir[j]), P.joff[j]);  // model_calibration.py:557
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
      fr_mul(F, P.seg[j]);
    }
  }
}

enum { MODE_NORMAL_EQ = 0, MODE_TOOL_BASE = 1 };

// Phase 1: the three slab rows of (measurement, tool t).
template <int MODE>
__device__ __forceinline__ void fill_rows(const CalibP& P, const double (&s)[RK_CALIB_MAX_JOINTS],
                                          const double (&c)[RK_CALIB_MAX_JOINTS], const Fr& Fend,
                                          const double* __restrict__ pts, int64_t i, bool live, int t,
                                          int ncols, double* __restrict__ rows /* [3][kRS] */) {
  const int wpad = ((ncols + 3) / 4) * 4;
  if (!live || MODE == MODE_TOOL_BASE) {
    for (int r = 0; r < 3; ++r)
      for (int k = 0; k < wpad; ++k) rows[r * kRS + k] = 0.0;
    if (!live) return;
  } else {
    for (int r = 0; r < 3; ++r)
      for (int k = ncols; k < wpad; ++k) rows[r * kRS + k] = 0.0;
  }
  const double* pm = pts + (i * P.nt + t) * 3;
  if (MODE == MODE_NORMAL_EQ) {
    const double On[3] = {
        Fend.p[0] + Fend.R[0] * P.tool[t][0] + Fend.R[1] * P.tool[t][1] + Fend.R[2] * P.tool[t][2],
        Fend.p[1] + Fend.R[3] * P.tool[t][0] + Fend.R[4] * P.tool[t][1] + Fend.R[5] * P.tool[t][2],
        Fend.p[2] + Fend.R[6] * P.tool[t][0] + Fend.R[7] * P.tool[t][1] + Fend.R[8] * P.tool[t][2]};
    Fr F;
    fr_init(F, P.Tb);
#pragma unroll
    for (int j = 0; j < RK_CALIB_MAX_JOINTS; ++j) {
      if (j < P.nj) {
        fr_joint(F, P.jaxis[j], P.jkind[j], s[j], c[j]);
        for (int p = P.pstart[j]; p < P.pstart[j + 1]; ++p) {
          const V3<double> ax = {P.pax[p][0], P.pax[p][1], P.pax[p][2]};
          const V3<double> u = rot(F.R, ax);
          V3<double> col = u;                         // translation parameter: u
          if (P.pkind[p] == 1) {                      // rotation parameter: u x (O_n - O_p)
            const V3<double> og = {P.porg[p][0], P.porg[p][1], P.porg[p][2]};
            const V3<double> ro = rot(F.R, og);
            const V3<double> d = {On[0] - (F.p[0] + ro.x), On[1] - (F.p[1] + ro.y),
                                  On[2] - (F.p[2] + ro.z)};
            col = cross(u, d);
          }
          rows[0 * kRS + p] = col.x;
          rows[1 * kRS + p] = col.y;
          rows[2 * kRS + p] = col.z;
        }
        fr_mul(F, P.seg[j]);
      }
    }
    // residual column: r = p_meas - T[:3,3] (model_calibration.py:420)
#pragma unroll
    for (int r = 0; r < 3; ++r) rows[r * kRS + P.np] = pm[r] - On[r];
  } else {
    // estimate_tool_base rows [I | skew(p)^T | R at tool block | dp] (model_calibration.py:306-329)
    const double px = Fend.p[0], py = Fend.p[1], pz = Fend.p[2];
    const double St[9] = {0.0, pz, -py, -pz, 0.0, px, py, -px, 0.0};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      rows[r * kRS + r] = 1.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        rows[r * kRS + 3 + k] = St[r * 3 + k];
        rows[r * kRS + 6 + 3 * t + k] = Fend.R[r * 3 + k];
      }
      rows[r * kRS + ncols - 1] = pm[r] - Fend.p[r];
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(kMeas, 3)
calib_accumulate_kernel(const __grid_constant__ CalibP P, const double* __restrict__ q,
                        const double* __restrict__ pts, int64_t m, int ncols,
                        double* __restrict__ partial, unsigned int* __restrict__ counter,
                        double* __restrict__ out) {
  extern __shared__ __align__(16) double slab[];  // [kMeas*3][kRS]
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int nblk = (ncols + 3) / 4;
  const int ntiles = nblk * (nblk + 1) / 2;
  const int ksplit = kMeas / ntiles;
  const bool worker = tid < ntiles * ksplit;
  const int tile = tid % ntiles, ks = tid / ntiles;
  int bi = 0, bj = 0;
  {
    int rem = tile;
    while (rem >= nblk - bi) {
      rem -= nblk - bi;
      ++bi;
    }
    bj = bi + rem;
  }
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

  for (int64_t base = (int64_t)blockIdx.x * kMeas; base < m; base += (int64_t)gridDim.x * kMeas) {
    const int64_t i = base + tid;
    const bool live = i < m;
    double sj[RK_CALIB_MAX_JOINTS], cj[RK_CALIB_MAX_JOINTS];
    load_joints(P, q, i, live, sj, cj);
    Fr Fend;
    chain_end(P, sj, cj, Fend);
    for (int t = 0; t < P.nt; ++t) {
      __syncthreads();  // previous phase 2 is done with the slab
      fill_rows<MODE>(P, sj, cj, Fend, pts, i, live, t, ncols, slab + tid * 3 * kRS);
      __syncthreads();
      if (worker) {
        const double* ra = slab + bi * 4;
        const double* rb = slab + bj * 4;
#pragma unroll 4
        for (int sl = ks; sl < kMeas * 3; sl += ksplit) {
          const double2 a01 = *reinterpret_cast<const double2*>(ra + sl * kRS);
          const double2 a23 = *reinterpret_cast<const double2*>(ra + sl * kRS + 2);
          const double2 b01 = *reinterpret_cast<const double2*>(rb + sl * kRS);
          const double2 b23 = *reinterpret_cast<const double2*>(rb + sl * kRS + 2);
          const double av[4] = {a01.x, a01.y, a23.x, a23.y};
          const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
      }
    }
  }

  // k-slices -> CTA partial (reuse the slab)
  __syncthreads();
  if (worker) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) slab[(ks * ntiles + tile) * 16 + a * 4 + b] = acc[a][b];
  }
  __syncthreads();
  for (int e = tid; e < ntiles * 16; e += kMeas) {
    double sum = 0.0;
    for (int k = 0; k < ksplit; ++k) sum += slab[k * ntiles * 16 + e];
    partial[(int64_t)blockIdx.x * kTileElems + e] = sum;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA: deterministic sum over CTAs, then scatter into the caller's format
  for (int e = tid; e < ntiles * 16; e += kMeas) {
    double sum = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) sum += __ldcg(partial + (int64_t)b * kTileElems + e);
    slab[e] = sum;
  }
  __syncthreads();
  for (int e = tid; e < ntiles * 16; e += kMeas) {
    const int tl = e / 16, a = (e % 16) / 4, b = e % 4;
    int ti = 0, rem = tl;
    while (rem >= nblk - ti) {
      rem -= nblk - ti;
      ++ti;
    }
    const int tj = ti + rem;
    const int gi = ti * 4 + a, gj = tj * 4 + b;
    if (gi > gj || gj >= ncols) continue;
    const double v = slab[e];
    if (MODE == MODE_NORMAL_EQ) {
      const int np = ncols - 1;
      if (gj < np) {
        out[gi * np - gi * (gi - 1) / 2 + (gj - gi)] = v;             // packed upper H
      } else if (gi < np) {
        out[np * (np + 1) / 2 + gi] = v;                               // g
      } else {
        out[np * (np + 1) / 2 + np] = v;                               // r^T r
      }
    } else {
      const int dim = ncols - 1;
      if (gj < dim) {
        out[gi * dim + gj] = v;
        out[gj * dim + gi] = v;
      } else if (gi < dim) {
        out[dim * dim + gi] = v;
      }
    }
  }
  if (tid == 0) *counter = 0;  // workspace is reusable without a memset
}

// evaluate (model_calibration.py:480-505): sums of squares and maxima of the residuals
__global__ void __launch_bounds__(kMeas)
calib_eval_kernel(const __grid_constant__ CalibP P, const double* __restrict__ q,
                  const double* __restrict__ pts, int64_t m, double* __restrict__ partial,
                  unsigned int* __restrict__ counter, double* __restrict__ out) {
  __shared__ double red[kMeas / kWarp][8];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double acc[8] = {0.0, 0.0, 0.0, 0.0, -1.0, -1.0, -1.0, -1.0};  // maxima start at -1 (:481)
  for (int64_t i = (int64_t)blockIdx.x * kMeas + tid; i < m; i += (int64_t)gridDim.x * kMeas) {
    double sj[RK_CALIB_MAX_JOINTS], cj[RK_CALIB_MAX_JOINTS];
    load_joints(P, q, i, true, sj, cj);
    Fr F;
    chain_end(P, sj, cj, F);
    for (int t = 0; t < P.nt; ++t) {
      const double* pm = pts + (i * P.nt + t) * 3;
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

constexpr size_t kWsCounterOffset = (size_t)kMaxGrid * kTileElems * sizeof(double);
constexpr size_t kWsBytes = kWsCounterOffset + 256;

template <int MODE>
static int launch_accumulate(const RkCalibModel* model, const double* q, const double* pts, int64_t m,
                             double* out, void* workspace, size_t ws_bytes, void* stream) {
  if (!model || !out || m < 0) return RK_ERR_INVALID_ARG;
  if (!workspace || ws_bytes < kWsBytes) return RK_ERR_WORKSPACE;
  if (m > 0 && (!q || !pts)) return RK_ERR_INVALID_ARG;
  CalibP P;
  int rc = make_calib(model, P, MODE == MODE_TOOL_BASE);
  if (rc != RK_OK) return rc;
  const int ncols = (MODE == MODE_NORMAL_EQ) ? P.np + 1 : 6 + 3 * P.nt + 1;
  if (ncols > kW) return RK_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)kMeas * 3 * kRS * sizeof(double);
  static thread_local bool attr_set[2] = {false, false};
  if (!attr_set[MODE]) {
    if (cudaFuncSetAttribute(calib_accumulate_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
      return check_launch();
    attr_set[MODE] = true;
  }
  int grid = persistent_grid(calib_accumulate_kernel<MODE>, kMeas, smem, m > 0 ? m : 1);
  if (grid > kMaxGrid) grid = kMaxGrid;
  unsigned int* counter = reinterpret_cast<unsigned int*>((char*)workspace + kWsCounterOffset);
  if (cudaMemsetAsync(counter, 0, sizeof(unsigned int), st) != cudaSuccess) return check_launch();
  calib_accumulate_kernel<MODE><<<grid, kMeas, smem, st>>>(P, q, pts, m, ncols, (double*)workspace,
                                                           counter, out);
  return check_launch();
}

}  // namespace rk

extern "C" {

int rk_calib_out_len(int p) { return p * (p + 1) / 2 + p + 1; }
size_t rk_calib_workspace_bytes(void) { return rk::kWsBytes; }

int rk_calib_normal_eq_f64(const RkCalibModel* model, const double* q, const double* pts, int64_t m,
                           double* out, void* workspace, size_t ws_bytes, void* stream) {
  return rk::launch_accumulate<rk::MODE_NORMAL_EQ>(model, q, pts, m, out, workspace, ws_bytes, stream);
}

int rk_calib_tool_base_eq_f64(const RkCalibModel* model, const double* q, const double* pts, int64_t m,
                              double* out, void* workspace, size_t ws_bytes, void* stream) {
  return rk::launch_accumulate<rk::MODE_TOOL_BASE>(model, q, pts, m, out, workspace, ws_bytes, stream);
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
  int grid = rk::persistent_grid(rk::calib_eval_kernel, rk::kMeas, 0, m > 0 ? m : 1);
  if (grid > rk::kMaxGrid) grid = rk::kMaxGrid;
  unsigned int* counter = reinterpret_cast<unsigned int*>((char*)workspace + rk::kWsCounterOffset);
  if (cudaMemsetAsync(counter, 0, sizeof(unsigned int), st) != cudaSuccess) return rk::check_launch();
  rk::calib_eval_kernel<<<grid, rk::kMeas, 0, st>>>(P, q, pts, m, (double*)workspace, counter, out);
  return rk::check_launch();
}

}  // extern "C"
