// Generic token-chain interpreter: batched Transformation(sequence, values) and
// JacobianCalculator.calculate_numerically (utils/robo_math.py:364-397,
// utils/jacobians.py:166-240).  One thread per sample; the opcode table lives in the
// kernel parameter space so every thread of the grid walks the same uniform switch.
#include "rk_common.cuh"

namespace rk {

struct Frame {
  double R[9];  // row-major
  double p[3];
};

__device__ __forceinline__ void frame_identity(Frame& F) {
#pragma unroll
  for (int i = 0; i < 9; ++i) F.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  F.p[0] = F.p[1] = F.p[2] = 0.0;
}

__device__ __forceinline__ void frame_from12(Frame& F, const double* T) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) F.R[i * 3 + j] = T[i * 4 + j];
    F.p[i] = T[i * 4 + 3];
  }
}

// F <- F * Trans(axis A, v)
template <int A>
__device__ __forceinline__ void frame_translate(Frame& F, double v) {
#pragma unroll
  for (int i = 0; i < 3; ++i) F.p[i] = fma(v, F.R[i * 3 + A], F.p[i]);
}

// F <- F * Rot(axis A, angle with sin s, cos c): Givens update of the two other columns
template <int A>
__device__ __forceinline__ void frame_rotate(Frame& F, double s, double c) {
  constexpr int I = (A + 1) % 3, Jc = (A + 2) % 3;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double a = F.R[r * 3 + I], b = F.R[r * 3 + Jc];
    F.R[r * 3 + I] = c * a + s * b;
    F.R[r * 3 + Jc] = c * b - s * a;
  }
}

// op = kind*3 + axis; kind 0 T, 1 R, 2 T inverse, 3 R inverse
__device__ __forceinline__ void frame_apply(Frame& F, int op, double v) {
  const int kind = op / 3, axis = op - kind * 3;
  if (kind == 0 || kind == 2) {
    const double t = (kind == 2) ? -v : v;
    switch (axis) {
      case 0: frame_translate<0>(F, t); break;
      case 1: frame_translate<1>(F, t); break;
      default: frame_translate<2>(F, t); break;
    }
  } else {
    double s, c;
    fast_sincos(v, &s, &c);
    if (kind == 3) s = -s;  // transpose
    switch (axis) {
      case 0: frame_rotate<0>(F, s, c); break;
      case 1: frame_rotate<1>(F, s, c); break;
      default: frame_rotate<2>(F, s, c); break;
    }
  }
}

__device__ __forceinline__ V3<double> frame_axis(const Frame& F, int axis) {
  switch (axis) {
    case 0: return {F.R[0], F.R[3], F.R[6]};
    case 1: return {F.R[1], F.R[4], F.R[7]};
    default: return {F.R[2], F.R[5], F.R[8]};
  }
}

// F <- F * X (X a 3x4 row-major constant)
__device__ __forceinline__ void frame_mul12(Frame& F, const double* X) {
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

template <bool SOA>
__global__ void __launch_bounds__(kBlock)
chain_fk_jac_kernel(const __grid_constant__ RkChain Ch, const double* __restrict__ values,
                    int nvals, int n_active, int64_t n, int64_t ld, double* __restrict__ T,
                    double* __restrict__ J) {
  extern __shared__ __align__(16) double smem_d[];  // token values [ntok][kBlock]
  double* tokv = smem_d + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    // All token values are fetched before the walk, eight independent loads at a time: fetched
    // inside the walk, every token waited for its own DRAM round trip (ncu: long_scoreboard 7.7
    // per issue, the kernel ran at one third of its memory roofline).
    for (int t0 = 0; t0 < Ch.ntok; t0 += 8) {
      double v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int t = t0 + k;
        const int sc = t < Ch.ntok ? Ch.src[t] : -1;
        v[k] = sc < 0 ? (t < Ch.ntok ? Ch.cst[t] : 0.0)
                      : (SOA ? __ldcs(values + (int64_t)sc * ld + i) : values[i * nvals + sc]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (t0 + k < Ch.ntok) tokv[(t0 + k) * kBlock] = v[k];
    }
    auto val = [&](int t) -> double { return tokv[t * kBlock]; };
    Frame F;
    if (Ch.has_base)
      frame_from12(F, Ch.T_base);
    else
      frame_identity(F);
    for (int t = 0; t < Ch.ntok; ++t) frame_apply(F, Ch.op[t], val(t));
    if (Ch.has_tool) frame_mul12(F, Ch.T_tool);
    if (T) {
      const double Tv[16] = {F.R[0], F.R[1], F.R[2], F.p[0], F.R[3], F.R[4], F.R[5], F.p[1],
                             F.R[6], F.R[7], F.R[8], F.p[2], 0.0,    0.0,    0.0,    1.0};
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (SOA)
          __stcs(T + (int64_t)k * ld + i, Tv[k]);
        else
          T[i * 16 + k] = Tv[k];
      }
    }
    if (J) {
      const V3<double> On = {F.p[0], F.p[1], F.p[2]};
      // second walk: the frame BEFORE each active token gives its screw axis and origin
      if (Ch.has_base)
        frame_from12(F, Ch.T_base);
      else
        frame_identity(F);
      int col = 0;
      for (int t = 0; t < Ch.ntok; ++t) {
        const int op = Ch.op[t];
        if (Ch.active[t]) {
          const int kind = op / 3, axis = op - kind * 3;
          const V3<double> u = frame_axis(F, axis);
          V3<double> lin, ang;
          if (kind == 1) {  // revolute: [u x (O_n - O_k); u]
            const V3<double> O = {F.p[0], F.p[1], F.p[2]};
            lin = cross(u, sub(On, O));
            ang = u;
          } else {          // prismatic: [u; 0]
            lin = u;
            ang = {0.0, 0.0, 0.0};
          }
          const double cv[6] = {lin.x, lin.y, lin.z, ang.x, ang.y, ang.z};
#pragma unroll
          for (int r = 0; r < 6; ++r) {
            if (SOA)
              __stcs(J + (int64_t)(r * n_active + col) * ld + i, cv[r]);
            else
              J[(i * 6 + r) * n_active + col] = cv[r];
          }
          ++col;
        }
        frame_apply(F, op, val(t));
      }
    }
  }
}

}  // namespace rk

extern "C" int rk_chain_fk_jac_f64(const RkChain* chain, const double* values, int nvals,
                                   int64_t n, int64_t ld, int flags, double* T, double* J,
                                   void* stream) {
  if (!chain || n < 0 || nvals < 0) return RK_ERR_INVALID_ARG;
  if (chain->ntok < 0 || chain->ntok > RK_MAX_TOKENS) return RK_ERR_UNSUPPORTED;
  int n_active = 0;
  for (int t = 0; t < chain->ntok; ++t) {
    const int op = chain->op[t];
    if (op < 0 || op >= 12) return RK_ERR_INVALID_ARG;
    if (chain->src[t] >= nvals) return RK_ERR_INVALID_ARG;
    if (chain->active[t]) {
      if (op >= 6) return RK_ERR_UNSUPPORTED;  // the reference has no 'Rxid' token either
      ++n_active;
    }
  }
  if (n == 0) return RK_OK;
  if (nvals > 0 && !values) return RK_ERR_INVALID_ARG;
  const bool soa = (flags & RK_SOA) != 0;
  if (soa && ld < n) return RK_ERR_INVALID_ARG;
  if (J && n_active == 0) J = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)chain->ntok * rk::kBlock * sizeof(double);
  if (soa) {
    int grid = rk::stream_grid(rk::chain_fk_jac_kernel<true>, rk::kBlock, smem, n);
    rk::chain_fk_jac_kernel<true><<<grid, rk::kBlock, smem, st>>>(*chain, values, nvals, n_active, n, ld, T, J);
  } else {
    int grid = rk::stream_grid(rk::chain_fk_jac_kernel<false>, rk::kBlock, smem, n);
    rk::chain_fk_jac_kernel<false><<<grid, rk::kBlock, smem, st>>>(*chain, values, nvals, n_active, n, ld, T, J);
  }
  return rk::check_launch();
}
