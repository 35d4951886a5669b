// Generic token-chain interpreter: batched Transformation(sequence, values) and
// JacobianCalculator.calculate_numerically (utils/robo_math.py:364-397, :556-608,
// utils/jacobians.py:166-240), and TrajectoryGenerator.get_dq_from_dx for any such chain
// (utils/trajectory_generator.py:161-188).  One thread per sample.
//
// The host turns the token list into a short program (ChainProg, in the kernel parameter space, so
// every thread of the grid walks the same uniform switch):
//   * runs of constant tokens are multiplied out once into a 3x4 constant (or a plain translation
//     vector) -- a constant rotation costs the batch nothing, not a sincos per sample;
//   * inverse tokens become a sign on the value, derivative tokens (Txd .. Rzd) their own step kind;
//   * the values of the variable tokens are fetched together before the walk (eight independent
//     loads in flight) and parked in shared memory.
// The Jacobian needs O_n, so it takes a second walk (column [u x (O_n - O_k); u] / [u; 0] from the
// frame before each active token) -- a cheap one: the first walk parks every sine / cosine in shared
// memory.  The frame itself lives in shared memory too (see chain_kernel).
#include <cmath>
#include <cstring>

#include "rk_chain.cuh"

namespace rk {

// ---- host: compile RkChain -> ChainProg -------------------------------------------------
static void h_ident(double* X) {
  for (int i = 0; i < 12; ++i) X[i] = (i % 5 == 0) ? 1.0 : 0.0;
}
static void h_mul(double* A, const double* B) {  // A <- A * B (3x4 homogeneous)
  double C[12];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j)
      C[i * 4 + j] = A[i * 4] * B[j] + A[i * 4 + 1] * B[4 + j] + A[i * 4 + 2] * B[8 + j] + (j == 3 ? A[i * 4 + 3] : 0.0);
  memcpy(A, C, sizeof(C));
}
static void h_elem(int kind4, int axis, double v, double* X) {  // kind4: 0 T, 1 R, 2 Ti, 3 Ri
  h_ident(X);
  if (kind4 == 0 || kind4 == 2) {
    X[axis * 4 + 3] = kind4 == 2 ? -v : v;
    return;
  }
  const double s = kind4 == 3 ? -sin(v) : sin(v), c = cos(v);
  const int i = (axis + 1) % 3, j = (axis + 2) % 3;
  X[i * 4 + i] = c, X[i * 4 + j] = -s, X[j * 4 + i] = s, X[j * 4 + j] = c;
}

template <typename S>
static int compile_chain(const RkChain* ch, int nvals, ChainProg<S>& P) {
  memset(&P, 0, sizeof(P));
  if (ch->ntok < 0 || ch->ntok > RK_MAX_TOKENS) return RK_ERR_UNSUPPORTED;
  double pend[12];
  bool have_pend = false, pend_rot = false;
  int npool = 0;
  auto flush = [&]() -> int {
    if (!have_pend) return RK_OK;
    if (npool >= kMaxConst) return RK_ERR_UNSUPPORTED;
    const int st = P.nstep++;
    P.kind[st] = pend_rot ? ST_CONST : ST_CONST_T;
    P.slot[st] = (int8_t)npool;
    for (int i = 0; i < 12; ++i) P.X[npool][i] = (S)pend[i];
    ++npool;
    have_pend = pend_rot = false;
    return RK_OK;
  };
  for (int t = 0; t < ch->ntok; ++t) {
    const int op = ch->op[t];
    if (op < 0 || op >= 18) return RK_ERR_INVALID_ARG;
    const int kind6 = op / 3, axis = op % 3;  // 0 T, 1 R, 2 Ti, 3 Ri, 4 Td, 5 Rd
    const int src = ch->src[t];
    if (src >= nvals) return RK_ERR_INVALID_ARG;
    if (ch->active[t] && kind6 >= 2) return RK_ERR_UNSUPPORTED;  // the reference has no 'Rxid' / 'Rxdd' token
    if (kind6 >= 4) P.has_deriv = 1;
    const bool foldable = src < 0 && !ch->active[t] && kind6 < 4;
    if (foldable) {
      double E[12];
      h_elem(kind6, axis, ch->cst[t], E);
      if (!have_pend) h_ident(pend), have_pend = true;
      h_mul(pend, E);
      pend_rot = pend_rot || kind6 == 1 || kind6 == 3;
      continue;
    }
    int rc = flush();
    if (rc != RK_OK) return rc;
    const int st = P.nstep++;
    P.kind[st] = (int8_t)(kind6 == 0 || kind6 == 2 ? ST_TRANS : (kind6 == 1 || kind6 == 3 ? ST_ROT : (kind6 == 4 ? ST_DTRANS : ST_DROT)));
    P.axis[st] = (int8_t)axis;
    P.neg[st] = (int8_t)(kind6 == 2 || kind6 == 3);
    P.active[st] = (int8_t)(ch->active[t] ? 1 : 0);
    if (ch->active[t]) ++P.na;
    if (src >= 0) {
      P.slot[st] = (int8_t)P.nvar;
      P.plane[P.nvar++] = (int16_t)src;
    } else {  // a constant that could not be folded (active or derivative): its value sits in the pool
      if (npool >= kMaxConst) return RK_ERR_UNSUPPORTED;
      P.slot[st] = (int8_t)(-1 - npool);
      P.X[npool][0] = (S)ch->cst[t];
      ++npool;
    }
  }
  int rc = flush();
  if (rc != RK_OK) return rc;
  P.nslot = P.nvar;
  for (int st = 0; st < P.nstep; ++st) {
    const bool cst = P.kind[st] == ST_CONST || P.kind[st] == ST_CONST_T;
    P.code[st] = (int8_t)(P.kind[st] * 3 + (cst ? 0 : P.axis[st]));
    if (P.kind[st] == ST_ROT && P.slot[st] >= 0) P.cslot[st] = (int8_t)P.nslot++;
  }
  for (int i = 0; i < 12; ++i) P.Tb[i] = (S)ch->T_base[i], P.Tt[i] = (S)ch->T_tool[i];
  P.has_base = ch->has_base, P.has_tool = ch->has_tool;
  return RK_OK;
}

// ---- device ------------------------------------------------------------------------------
// The walk is an interpreter: which column a step touches is data (the axis of the token), uniform
// over the grid.  With the frame in REGISTERS every axis needs its own code path and the paths must
// agree on a register assignment where they meet: ncu showed 26 % of the instructions of the
// register version to be IMAD.MOV shuffles and 9 % branches around a switch of 18 cases (~200
// instructions per step where the arithmetic is 12).  Here the rotation part of the frame lives in
// SHARED memory, column-major (fr[(3 c + r) * kBlock + tid]), where a column is selected by address
// arithmetic in the uniform datapath; only the origin p -- which no step permutes -- stays in
// registers.  A rotation is 6 LDS + 12 FP64 + 6 STS, a translation 3 LDS + 3 DFMA, whatever the axis.
//
// Shared memory per thread: 9 frame entries + nslot staged values.  Slot v < nvar holds the value of
// variable v; a variable ROTATION's sine replaces its value during the first walk and its cosine
// goes to slot cslot[st], so the second (Jacobian) walk repeats no sincos.
// MODE 0: pose and / or Jacobian; MODE 1: dq = J[rows]^-1 dx (J is na x na after the row selection)
struct Pass0 {
  static constexpr int value = 0;
};
struct Pass1 {
  static constexpr int value = 1;
};

template <typename S, bool SOA, int MODE, int NA>
__global__ void __launch_bounds__(kBlock)
chain_kernel(const __grid_constant__ ChainProg<S> Pg, const S* __restrict__ values, int nvals, int64_t n,
             int64_t ld, S* __restrict__ T, S* __restrict__ J, const S* __restrict__ dx, S* __restrict__ dq,
             uint8_t* __restrict__ ok) {
  using M = Sc<S>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S* fr = reinterpret_cast<S*>(smem_raw) + threadIdx.x;  // frame entry k at fr[k * kBlock]
  S* tokv = fr + 9 * kBlock;                             // staged value of slot v at tokv[v * kBlock]
  const int na = Pg.na;
  const bool want_jac = (MODE == 1) || (J != nullptr && na > 0);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    // all values first, eight independent loads at a time: fetched inside the walk every token
    // waited for its own DRAM round trip (ncu, round 1: long_scoreboard 7.7 per issue)
    for (int v0 = 0; v0 < Pg.nvar; v0 += 8) {
      S v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int sl = v0 + k;
        const int pl = sl < Pg.nvar ? Pg.plane[sl] : 0;
        v[k] = sl < Pg.nvar ? (SOA ? __ldcs(values + (int64_t)pl * ld + i) : values[i * nvals + pl]) : S(0);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (v0 + k < Pg.nvar) tokv[(v0 + k) * kBlock] = v[k];
    }
    S p[3];
    // one walk over the program; PASS 0 computes (and parks) the sincos, PASS 1 reuses them and
    // hands the frame BEFORE every active token to `at_active`
    auto walk = [&](auto pass_tag, auto&& at_active) {
      constexpr int PASS = decltype(pass_tag)::value;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) fr[(3 * c + r) * kBlock] = Pg.has_base ? Pg.Tb[r * 4 + c] : S(r == c ? 1 : 0);
#pragma unroll
      for (int r = 0; r < 3; ++r) p[r] = Pg.has_base ? Pg.Tb[r * 4 + 3] : S(0);
      int col = 0;
      for (int st = 0; st < Pg.nstep; ++st) {
        const int kind = Pg.kind[st], axis = Pg.axis[st], sl = Pg.slot[st];  // uniform
        if (PASS == 1 && Pg.active[st]) {
          at_active(st, col, axis, kind);
          ++col;
        }
        S* cA = fr + 3 * axis * kBlock;                                   // column `axis`
        S* cI = fr + 3 * (axis == 2 ? 0 : axis + 1) * kBlock;             // the two columns a rotation mixes
        S* cJ = fr + 3 * (axis == 0 ? 2 : axis - 1) * kBlock;
        if (kind == ST_TRANS) {
          S v = sl >= 0 ? tokv[sl * kBlock] : Pg.X[-1 - sl][0];
          if (Pg.neg[st]) v = -v;
#pragma unroll
          for (int r = 0; r < 3; ++r) p[r] = fma(v, cA[r * kBlock], p[r]);
        } else if (kind == ST_ROT || kind == ST_DROT) {
          S sn, cs;
          if (PASS == 0 || sl < 0 || kind == ST_DROT) {
            S v = sl >= 0 ? tokv[sl * kBlock] : Pg.X[-1 - sl][0];
            if (Pg.neg[st]) v = -v;
            M::sincos(v, &sn, &cs);
            if (PASS == 0 && want_jac && sl >= 0 && kind == ST_ROT) {
              tokv[sl * kBlock] = sn;
              tokv[Pg.cslot[st] * kBlock] = cs;
            }
          } else {
            sn = tokv[sl * kBlock], cs = tokv[Pg.cslot[st] * kBlock];
          }
          if (kind == ST_ROT) {  // F <- F Rot_axis: Givens update of columns I, J
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              const S a = cI[r * kBlock], b = cJ[r * kBlock];
              cI[r * kBlock] = cs * a + sn * b;
              cJ[r * kBlock] = cs * b - sn * a;
            }
          } else {               // derivative (robo_math.py:583-608): d/dv Rot, zero corner
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              const S a = cI[r * kBlock], b = cJ[r * kBlock];
              cI[r * kBlock] = cs * b - sn * a;
              cJ[r * kBlock] = -cs * a - sn * b;
              cA[r * kBlock] = S(0);
              p[r] = S(0);
            }
          }
        } else if (kind == ST_CONST_T || kind == ST_CONST) {  // F <- F X, X a folded constant
          const S* X = Pg.X[sl];
          S R[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) R[k] = fr[k * kBlock];   // R[3 c + r]
#pragma unroll
          for (int r = 0; r < 3; ++r) p[r] += R[r] * X[3] + R[3 + r] * X[7] + R[6 + r] * X[11];
          if (kind == ST_CONST) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
              for (int r = 0; r < 3; ++r)
                fr[(3 * c + r) * kBlock] = R[r] * X[c] + R[3 + r] * X[4 + c] + R[6 + r] * X[8 + c];
          }
        } else {  // ST_DTRANS (robo_math.py:556-581): p <- column `axis`, R <- 0
#pragma unroll
          for (int r = 0; r < 3; ++r) p[r] = cA[r * kBlock];
#pragma unroll
          for (int k = 0; k < 9; ++k) fr[k * kBlock] = S(0);
        }
      }
      if (Pg.has_tool) {
        S R[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = fr[k * kBlock];
#pragma unroll
        for (int r = 0; r < 3; ++r) p[r] += R[r] * Pg.Tt[3] + R[3 + r] * Pg.Tt[7] + R[6 + r] * Pg.Tt[11];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int r = 0; r < 3; ++r)
            fr[(3 * c + r) * kBlock] = R[r] * Pg.Tt[c] + R[3 + r] * Pg.Tt[4 + c] + R[6 + r] * Pg.Tt[8 + c];
      }
    };
    walk(Pass0{}, [](int, int, int, int) {});
    if (MODE == 0 && T) {
      const S w = Pg.has_deriv ? S(0) : S(1);  // a derivative factor zeroes the corner of the product
      S* tp = SOA ? T + i : T + i * 16;
      const int64_t step = SOA ? ld : 1;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c, tp += step) {
          const S v = r == 3 ? (c == 3 ? w : S(0)) : (c == 3 ? p[r] : fr[(3 * c + r) * kBlock]);
          if (SOA) __stcs(tp, v);
          else *tp = v;
        }
    }
    if (!want_jac) continue;
    const V3<S> On = {p[0], p[1], p[2]};
    // column a of the Jacobian: revolute [u x (O_n - O_k); u], prismatic [u; 0]
    // (utils/jacobians.py:203-233; equal to calculate_numerically's matrix-derivative form)
    auto column = [&](int axis, int kind, S (&cv)[6]) {
      const S* cA = fr + 3 * axis * kBlock;
      const V3<S> u = {cA[0], cA[kBlock], cA[2 * kBlock]};
      if (kind == ST_ROT) {
        const V3<S> O = {p[0], p[1], p[2]};
        const V3<S> lin = cross(u, sub(On, O));
        cv[0] = lin.x, cv[1] = lin.y, cv[2] = lin.z, cv[3] = u.x, cv[4] = u.y, cv[5] = u.z;
      } else {
        cv[0] = u.x, cv[1] = u.y, cv[2] = u.z, cv[3] = cv[4] = cv[5] = S(0);
      }
    };
    if constexpr (MODE == 0) {
      walk(Pass1{}, [&](int, int a, int axis, int kind) {
        S cv[6];
        column(axis, kind, cv);
        S* jp = SOA ? J + (int64_t)a * ld + i : J + i * 6 * na + a;
        const int64_t step = SOA ? (int64_t)na * ld : na;
#pragma unroll
        for (int r = 0; r < 6; ++r, jp += step) {
          if (SOA) __stcs(jp, cv[r]);
          else *jp = cv[r];
        }
      });
    } else {
      // get_dq_from_dx (utils/trajectory_generator.py:161-188): the reference evaluates a symbolic
      // inverse of the square block J[rows, :]; here that block is solved with partial pivoting
      S A[NA][NA], b[NA];
      walk(Pass1{}, [&](int, int a, int axis, int kind) {
        S cv[6];
        column(axis, kind, cv);
#pragma unroll
        for (int c2 = 0; c2 < NA; ++c2) {
          if (c2 != a) continue;
#pragma unroll
          for (int r = 0; r < NA; ++r) {
            const int row = Pg.rows[r];
            S pick = cv[0];
#pragma unroll
            for (int k = 1; k < 6; ++k) pick = (row == k) ? cv[k] : pick;
            A[r][c2] = pick;
          }
        }
      });
#pragma unroll
      for (int r = 0; r < NA; ++r) b[r] = SOA ? dx[(int64_t)r * ld + i] : dx[i * NA + r];
      const bool good = gauss_solve<S, NA>(A, b);
#pragma unroll
      for (int r = 0; r < NA; ++r) {
        if (SOA) dq[(int64_t)r * ld + i] = b[r];
        else dq[i * NA + r] = b[r];
      }
      if (ok) ok[i] = good ? 1 : 0;
    }
  }
}

template <typename S>
static int launch_chain(const RkChain* chain, const S* values, int nvals, int64_t n, int64_t ld, int flags, S* T,
                        S* J, void* stream) {
  if (!chain || n < 0 || nvals < 0) return RK_ERR_INVALID_ARG;
  ChainProg<S> Pg;
  int rc = compile_chain<S>(chain, nvals, Pg);
  if (rc != RK_OK) return rc;
  if (Pg.has_deriv && J && Pg.na > 0) return RK_ERR_UNSUPPORTED;  // no Jacobian of a derivative chain
  if (n == 0) return RK_OK;
  if (Pg.nvar > 0 && !values) return RK_ERR_INVALID_ARG;
  const bool soa = (flags & RK_SOA) != 0;
  if (soa && ld < n) return RK_ERR_INVALID_ARG;
  if (J && Pg.na == 0) J = nullptr;
  if (!T && !J) return RK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // straight-line code specialised for this chain (compiled once, rk_chain_jit.cu) when NVRTC is
  // there; the interpreting kernel below otherwise
  rc = chain_jit_launch<S>(Pg, soa, values, nvals, n, ld, T, J, st);
  if (rc != RK_ERR_UNSUPPORTED) return rc;
  const size_t smem = (size_t)(9 + (J ? Pg.nslot : Pg.nvar)) * kBlock * sizeof(S);
  if (soa) {
    auto kern = chain_kernel<S, true, 0, 1>;
    if (!allow_dynamic_smem(kern, smem)) return RK_ERR_CUDA;
    int grid = stream_grid(kern, kBlock, smem, n);
    kern<<<grid, kBlock, smem, st>>>(Pg, values, nvals, n, ld, T, J, nullptr, nullptr, nullptr);
  } else {
    auto kern = chain_kernel<S, false, 0, 1>;
    if (!allow_dynamic_smem(kern, smem)) return RK_ERR_CUDA;
    int grid = stream_grid(kern, kBlock, smem, n);
    kern<<<grid, kBlock, smem, st>>>(Pg, values, nvals, n, ld, T, J, nullptr, nullptr, nullptr);
  }
  return check_launch();
}

template <typename S, bool SOA, int NA>
static int launch_solve_na(const ChainProg<S>& Pg, const S* values, int nvals, int64_t n, int64_t ld, const S* dx,
                           S* dq, uint8_t* ok, cudaStream_t st) {
  auto kern = chain_kernel<S, SOA, 1, NA>;
  const size_t smem = (size_t)(9 + Pg.nslot) * kBlock * sizeof(S);
  if (!allow_dynamic_smem(kern, smem)) return RK_ERR_CUDA;
  int grid = stream_grid(kern, kBlock, smem, n);
  kern<<<grid, kBlock, smem, st>>>(Pg, values, nvals, n, ld, nullptr, nullptr, dx, dq, ok);
  return check_launch();
}

static int launch_chain_solve(const RkChain* chain, const double* values, int nvals, const int32_t* rows,
                              const double* dx, int64_t n, int64_t ld, int flags, double* dq, uint8_t* ok,
                              void* stream) {
  if (!chain || n < 0 || nvals < 0) return RK_ERR_INVALID_ARG;
  ChainProg<double> Pg;
  int rc = compile_chain<double>(chain, nvals, Pg);
  if (rc != RK_OK) return rc;
  if (Pg.has_deriv || Pg.na < 1 || Pg.na > 6) return RK_ERR_UNSUPPORTED;
  for (int r = 0; r < Pg.na; ++r) {
    const int row = rows ? rows[r] : r;  // default: the first na rows (position rows for na <= 3)
    if (row < 0 || row > 5) return RK_ERR_INVALID_ARG;
    for (int k = 0; k < r; ++k)
      if (Pg.rows[k] == row) return RK_ERR_INVALID_ARG;
    Pg.rows[r] = (int8_t)row;
  }
  if (n == 0) return RK_OK;
  if ((Pg.nvar > 0 && !values) || !dx || !dq) return RK_ERR_INVALID_ARG;
  const bool soa = (flags & RK_SOA) != 0;
  if (soa && ld < n) return RK_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  // straight-line specialisation first (rk_chain_jit.cu), the interpreting kernel otherwise
  rc = chain_jit_solve_launch<double>(Pg, soa, values, nvals, n, ld, dx, dq, ok, st);
  if (rc != RK_ERR_UNSUPPORTED) return rc;
#define RK_NA(NA)                                                                                         \
  case NA:                                                                                                \
    return soa ? launch_solve_na<double, true, NA>(Pg, values, nvals, n, ld, dx, dq, ok, st)              \
               : launch_solve_na<double, false, NA>(Pg, values, nvals, n, ld, dx, dq, ok, st);
  switch (Pg.na) { RK_NA(1) RK_NA(2) RK_NA(3) RK_NA(4) RK_NA(5) RK_NA(6) }
#undef RK_NA
  return RK_ERR_UNSUPPORTED;
}

template <typename S>
static int64_t jit_source(const RkChain* chain, int nvals, int flags, int want, char* buf, int64_t cap) {
  ChainProg<S> Pg;
  if (compile_chain<S>(chain, nvals, Pg) != RK_OK) return -1;
  if (want & 8)
    for (int r = 0; r < Pg.na && r < 6; ++r) Pg.rows[r] = (int8_t)r;
  const std::string src = chain_jit_source<S>(Pg, flags & RK_SOA, want & 1, want & 2, want & 4, want & 8);
  if (src.empty()) return -1;
  if (buf && cap > 0) {
    const size_t m = src.size() < (size_t)cap - 1 ? src.size() : (size_t)cap - 1;
    memcpy(buf, src.data(), m);
    buf[m] = 0;
  }
  return (int64_t)src.size() + 1;
}

}  // namespace rk

extern "C" {
int64_t rk_chain_jit_source(const RkChain* chain, int nvals, int flags, int want, int f32, char* buf, int64_t cap) {
  if (!chain) return -1;
  return f32 ? rk::jit_source<float>(chain, nvals, flags, want, buf, cap)
             : rk::jit_source<double>(chain, nvals, flags, want, buf, cap);
}
int rk_chain_fk_jac_f64(const RkChain* chain, const double* values, int nvals, int64_t n, int64_t ld, int flags,
                        double* T, double* J, void* stream) {
  return rk::launch_chain<double>(chain, values, nvals, n, ld, flags, T, J, stream);
}
int rk_chain_fk_jac_f32(const RkChain* chain, const float* values, int nvals, int64_t n, int64_t ld, int flags,
                        float* T, float* J, void* stream) {
  return rk::launch_chain<float>(chain, values, nvals, n, ld, flags, T, J, stream);
}
int rk_chain_dq_from_dx_f64(const RkChain* chain, const double* values, int nvals, const int32_t* rows,
                            const double* dx, int64_t n, int64_t ld, int flags, double* dq, uint8_t* ok,
                            void* stream) {
  return rk::launch_chain_solve(chain, values, nvals, rows, dx, n, ld, flags, dq, ok, stream);
}
}
