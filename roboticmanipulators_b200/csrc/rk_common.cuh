// Shared device helpers for librk_b200 (sm_100a).  No torch, no third-party headers.
#pragma once
#include <type_traits>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "rk_b200.h"

namespace rk {

constexpr int kBlock = 128;  // threads per CTA for the pose-per-thread kernels
constexpr int kWarp = 32;

// Set by launch helpers when a CUDA runtime call fails (reported by rk_last_cuda_error()).
extern thread_local cudaError_t g_last_cuda;

inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_last_cuda = e;
    return RK_ERR_CUDA;
  }
  return RK_OK;
}

// Opt a kernel in to more than 48 KB of dynamic shared memory.  Called before every such launch:
// the attribute is per device and per context, the call costs about a microsecond, and a cached
// "already done" flag would be wrong as soon as one host thread drives two GPUs.
template <typename K>
inline bool allow_dynamic_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return true;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess)
    return true;
  g_last_cuda = cudaGetLastError();
  return false;
}

// Grid for a kernel that must stay resident for the whole problem (per-CTA accumulators that are
// reduced at the end): enough CTAs to fill every SM at the kernel's occupancy, never more than
// the work needs ("a multiple of 148 x resident CTAs per SM").
// Occupancy of (kernel, block, smem) on the current device, looked up once: the runtime query
// costs ~2 us and used to be made on every launch (at 1e4 poses the kernel itself runs 1.5 us).
// Returns resident CTAs on the whole device (SM count x CTAs per SM).  Defined in rk_api.cu.
int64_t cached_resident_ctas(const void* kernel, int block, size_t smem);
int sm_count();  // SMs of the current device (rk_api.cu)

template <typename K>
inline int persistent_grid(K kernel, int block, size_t smem, int64_t work_items) {
  int64_t need = (work_items + block - 1) / block;
  const int64_t cap = cached_resident_ctas(reinterpret_cast<const void*>(kernel), block, smem);
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// Grid for a streaming grid-stride kernel: NOT persistent on purpose.  Every thread makes about
// `trips` loop trips, so a large problem gets many times more CTAs than fit on the machine and
// the hardware block scheduler hands them to SMs as they free up.  SMs do not get equal shares
// of the HBM bandwidth; with a persistent grid and a static stride the average SM sat idle for
// 14 % of K1 (ncu sm__cycles_active min / avg / max = 6.7 / 9.4 / 10.9 M cycles) waiting for the
// slowest one.  Problems that would give fewer than 4 waves keep the resident grid: a partly
// filled last wave costs more than the balancing gains (measured on 32 K trajectories: +25 %).
// trips = 8 suits the thread-per-pose kernels (their software prefetch wants a few trips); the
// write-only warp-per-trajectory samplers run 8-15 % faster with trips = 2 (rk_traj.cu).
template <typename K>
inline int stream_grid(K kernel, int block, size_t smem, int64_t work_items, int trips = 8) {
  const int64_t cap = persistent_grid(kernel, block, smem, work_items);
  const int64_t need = (work_items + (int64_t)block * trips - 1) / ((int64_t)block * trips);
  const int64_t g = need >= 4 * cap ? need : cap;
  return (int)(g < 0x7fffffff ? g : 0x7fffffff);
}

// ---------------------------------------------------------------------------------
// Scalar traits: correctly rounded, never-contracted primitives (the reference's
// threshold expressions are plain IEEE mul/add sequences) and libm-grade transcendentals.
// ---------------------------------------------------------------------------------
template <typename S>
struct Sc;

template <>
struct Sc<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
  static __device__ __forceinline__ void sincos(double x, double* s, double* c);
  static __device__ __forceinline__ double sin(double x) { return ::sin(x); }
  static __device__ __forceinline__ double cos(double x) { return ::cos(x); }
  static __device__ __forceinline__ double acos(double x) { return ::acos(x); }
  static __device__ __forceinline__ double atan2(double y, double x) { return ::atan2(y, x); }
  static __device__ __forceinline__ double atan2_fast(double y, double x);
  static __device__ __forceinline__ double atan2_core(double y, double x, bool& rare);
  static __device__ __forceinline__ double rcp_fast(double x);
  static __device__ __forceinline__ double rsqrt_fast(double x);
  static __device__ __forceinline__ double sqrt_fast(double x);
  static __device__ __forceinline__ double abs(double x) { return ::fabs(x); }
  // neg ? -x : x by an integer XOR on the sign bit (keeps the FP64 pipe free; neg is 0 or 1)
  static __device__ __forceinline__ double flip(double x, int neg) {
    return __hiloint2double(__double2hiint(x) ^ (neg << 31), __double2loint(x));
  }
  // x with its sign bit XORed by mask (0 or 0x80000000): multiplication by +-1 without the FP64 pipe
  static __device__ __forceinline__ double flipm(double x, unsigned mask) {
    return __hiloint2double(__double2hiint(x) ^ (int)mask, __double2loint(x));
  }
  static __device__ __forceinline__ double pi() { return 3.141592653589793; }
};

template <>
struct Sc<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
  static __device__ __forceinline__ void sincos(float x, float* s, float* c) { ::sincosf(x, s, c); }
  static __device__ __forceinline__ float sin(float x) { return ::sinf(x); }
  static __device__ __forceinline__ float cos(float x) { return ::cosf(x); }
  static __device__ __forceinline__ float acos(float x) { return ::acosf(x); }
  static __device__ __forceinline__ float atan2(float y, float x) { return ::atan2f(y, x); }
  static __device__ __forceinline__ float atan2_fast(float y, float x);
  static __device__ __forceinline__ float atan2_core(float y, float x, bool& rare);
  static __device__ __forceinline__ float rcp_fast(float x) {  // normal finite x, <= 1 ulp
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
  }
  static __device__ __forceinline__ float rsqrt_fast(float x) { return ::rsqrtf(x); }
  static __device__ __forceinline__ float sqrt_fast(float x) {  // 0 for x <= 0; no slow path
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(x, 0.0f)));
    return r;
  }
  static __device__ __forceinline__ float abs(float x) { return ::fabsf(x); }
  static __device__ __forceinline__ float flip(float x, int neg) {
    return __int_as_float(__float_as_int(x) ^ (neg << 31));
  }
  static __device__ __forceinline__ float flipm(float x, unsigned mask) {
    return __int_as_float(__float_as_int(x) ^ (int)mask);
  }
  static __device__ __forceinline__ float pi() { return 3.14159265358979323846f; }
};

// ---------------------------------------------------------------------------------
// Lean FP64 primitives for the FP64-pipe-bound kernels.  CUDA's libm atan2/acos cost
// ~170 issue slots per call on sm_100a, two thirds of them UMOV/IMAD moves that
// materialise 64-bit polynomial constants and special-case selects.  Here the polynomial
// lives in constant memory (FFMA/DFMA take a constant-bank operand directly), the
// reciprocal is MUFU.RCP64H + two Newton steps, and everything outside the finite,
// normal range falls back to libm.
// ---------------------------------------------------------------------------------
__constant__ double kAtanPoly[11] = {
    -0x1.5555555555555p-2, 0x1.999999999934ap-3, -0x1.24924924360cbp-3, 0x1.c71c7185314cbp-4,
    -0x1.745d0b26b83e7p-4, 0x1.3b1262d95579ep-4, -0x1.10fa75382537fp-4, 0x1.dfe61e80903d2p-5,
    -0x1.a098bb6ba4941p-5, 0x1.41603647c7a7cp-5, -0x1.3a2b7a07caea9p-6};

// 1/x for normal finite x: hardware seed (2^-23) + two Newton iterations (<= 1 ulp).
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// 1/sqrt(x) for normal finite x > 0 (<= 1 ulp).
__device__ __forceinline__ double fast_rsqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * r, r, 0.5);
  r = fma(r, e, r);
  e = fma(-hx * r, r, 0.5);
  return fma(r, e, r);
}

// atan2(y, x), <= 2 ulp whenever max(|x|,|y|) is normal and moderate (exact zeros of the smaller
// operand included); `rare` is raised for everything else (both zero, denormal/huge magnitudes,
// inf, NaN) and the caller then takes libm.  Octant reduction to |u| <= tan(pi/8) with ONE division:
//   t = min/max;  t > tan(pi/8):  atan t = pi/4 + atan((min - max)/(min + max)).
// atan(u) = u + u^3 g(u^2), g = degree-10 interpolant at Chebyshev nodes (error 8e-18).
// Straight-line code: no branch, so independent calls interleave in the FP64 pipe.
__device__ __forceinline__ double fast_atan2_core(double y, double x, bool& rare) {
  // |x|, |y| by integer masks and the octant flag from the sign of mn - tan(pi/8) mx: the FP64
  // pipe sees one compare and one FMA here
  const double ax = __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x));
  const double ay = __hiloint2double(__double2hiint(y) & 0x7fffffff, __double2loint(y));
  const bool swap = ay > ax;
  const double mx = swap ? ay : ax, mn = swap ? ax : ay;
  const double t8 = fma(-0x1.a827999fcef32p-2, mx, mn);
  // bigf is 1.0 (upper octant) or 0.0: the octant shift is then plain FMAs
  const double bigf = __hiloint2double(~(__double2hiint(t8) >> 31) & 0x3ff00000, 0);
  const double num = fma(-bigf, mx, mn);  // big ? mn - mx : mn
  const double den = fma(bigf, mn, mx);   // big ? mn + mx : mx
  // u = num / den: hardware seed (2^-23), one Newton step (2^-46), then one residual
  // correction of the quotient itself (<= 0.5 ulp + 2^-92)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
  r = fma(r, fma(-den, r, 1.0), r);
  double u = num * r;
  u = fma(fma(-den, u, num), r, u);
  const double s = u * u, s2 = s * s;
  // even / odd Horner chains in s^2 for instruction-level parallelism
  double pe = kAtanPoly[10];
  pe = fma(pe, s2, kAtanPoly[8]);
  pe = fma(pe, s2, kAtanPoly[6]);
  pe = fma(pe, s2, kAtanPoly[4]);
  pe = fma(pe, s2, kAtanPoly[2]);
  pe = fma(pe, s2, kAtanPoly[0]);
  double po = kAtanPoly[9];
  po = fma(po, s2, kAtanPoly[7]);
  po = fma(po, s2, kAtanPoly[5]);
  po = fma(po, s2, kAtanPoly[3]);
  po = fma(po, s2, kAtanPoly[1]);
  const double p = fma(po, s, pe);
  double a = fma(u * s, p, u);
  a = fma(bigf, 0x1.921fb54442d18p-1, a);       // + pi/4 in the upper octant
  if (swap) a = 0x1.921fb54442d18p+0 - a;       // pi/2
  if (__double2hiint(x) < 0) a = 0x1.921fb54442d18p+1 - a;  // pi
  a = copysign(a, y);
  // valid when den = max (+ min) is in [2^-900, 2^+900]: exponent window on its high word
  // (den >= 0).  Zero, denormal, huge, inf and NaN fail it.
  rare = rare || ((unsigned)__double2hiint(den) - 0x07b00000u > 0x70800000u);
  return a;
}

__device__ __forceinline__ double fast_atan2(double y, double x) {
  bool rare = false;
  double a = fast_atan2_core(y, x, rare);
  if (rare) a = ::atan2(y, x);
  return a;
}

// sin and cos together, <= ~1.5 ulp for |x| <= 1e5 (libm beyond that and for NaN/inf).
// Cody-Waite reduction by pi/2 in three FMA steps, degree-11/12 polynomials in r^2 whose
// coefficients sit in constant memory, quadrant fix-up with selects.  No branch until the
// final range test, so the six joint sincos of a pose interleave instead of serialising
// behind libm's Payne-Hanek branch (K1 went from 873 to ~480 issue slots per pose).
__constant__ double kSinPoly[6] = {-0x1.5555555555555p-3, 0x1.1111111110b23p-7,  -0x1.a01a019e623afp-13,
                                   0x1.71de376b9c575p-19, -0x1.ae5fdeb852566p-26, 0x1.5dfb4dfaf15abp-33};
__constant__ double kCosPoly[6] = {0x1.5555555555555p-5,  -0x1.6c16c16c1691fp-10, 0x1.a01a019f3df42p-16,
                                   -0x1.27e4fa025e03dp-22, 0x1.1eeb52bea9072p-29, -0x1.906dd38cb66a1p-37};

// libm's sincos behind a real call: inlined, its Payne-Hanek path (a local-memory table walk, ~400
// instructions) was replicated at every sincos site of every unrolled chain -- 10 copies in the
// structure-specialised calibration kernel, 6 in K1 -- and blew the instruction cache for a path
// that joint angles never take.
static __device__ __noinline__ double2 slow_sincos(double x) {  // by value: results stay in registers
  double2 r;
  ::sincos(x, &r.x, &r.y);
  return r;
}

__device__ __forceinline__ void fast_sincos(double x, double* sp, double* cp) {
  const double kf = rint(x * 0x1.45f306dc9c883p-1);  // x * 2/pi
  const int q = (int)kf;
  double r = fma(-kf, 0x1.921fb54442d18p+0, x);
  r = fma(-kf, 0x1.1a62633145c07p-54, r);
  r = fma(-kf, -0x1.f1976b7ed8fbcp-110, r);
  const double z = r * r;
  double ps = kSinPoly[5];
  ps = fma(ps, z, kSinPoly[4]);
  ps = fma(ps, z, kSinPoly[3]);
  ps = fma(ps, z, kSinPoly[2]);
  ps = fma(ps, z, kSinPoly[1]);
  ps = fma(ps, z, kSinPoly[0]);
  double pc = kCosPoly[5];
  pc = fma(pc, z, kCosPoly[4]);
  pc = fma(pc, z, kCosPoly[3]);
  pc = fma(pc, z, kCosPoly[2]);
  pc = fma(pc, z, kCosPoly[1]);
  pc = fma(pc, z, kCosPoly[0]);
  const double sn = fma(r * z, ps, r);
  const double cs = fma(z * z, pc, fma(-0.5, z, 1.0));
  double s = (q & 1) ? cs : sn;
  double c = (q & 1) ? sn : cs;
  if (q & 2) s = -s;
  if ((q + 1) & 2) c = -c;
  if (!(fabs(x) <= 1.0e5)) {
    const double2 r2 = slow_sincos(x);
    s = r2.x, c = r2.y;
  }
  *sp = s;
  *cp = c;
}

// sqrt(t) for t >= 0 without the IEEE slow-path branch (<= 1 ulp); 0 for t <= 0 or tiny.
// Hardware seed (2^-23), one Newton step on 1/sqrt (2^-45), then the residual correction of the
// root itself, which squares the error again.
__device__ __forceinline__ double fast_sqrt(double t) {
  const bool ok = t > 1e-290;
  const double x = ok ? t : 1.0;
  double rs;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rs) : "d"(x));
  rs = fma(rs, fma(-0.5 * x * rs, rs, 0.5), rs);
  double r = x * rs;
  r = fma(fma(-r, r, x), 0.5 * rs, r);
  return ok ? r : 0.0;
}

// FP32 twin of fast_atan2_core (same octant reduction; degree-4 interpolant, error 1e-9 before
// rounding; about 1.5 ulp overall).  `rare` outside a moderate exponent window -> caller uses libm.
__device__ __forceinline__ float fast_atan2f_core(float y, float x, bool& rare) {
  const float ax = fabsf(x), ay = fabsf(y);
  const bool swap = ay > ax;
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float bigf = fmaf(-0.41421356237f, mx, mn) >= 0.0f ? 1.0f : 0.0f;
  const float num = fmaf(-bigf, mx, mn), den = fmaf(bigf, mn, mx);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
  float u = num * r;
  u = fmaf(fmaf(-den, u, num), r, u);
  const float s = u * u;
  float p = -0.06451927870512009f;
  p = fmaf(p, s, 0.10743731260299683f);
  p = fmaf(p, s, -0.14263956248760223f);
  p = fmaf(p, s, 0.19999539852142334f);
  p = fmaf(p, s, -0.3333333134651184f);
  float a = fmaf(u * s, p, u);
  a = fmaf(bigf, 0.78539816339744831f, a);
  if (swap) a = 1.57079632679489662f - a;
  if (__float_as_int(x) < 0) a = 3.14159265358979324f - a;
  a = copysignf(a, y);
  // den in [2^-100, 2^+100]: zero, denormal, huge, inf and NaN fail
  rare = rare || ((unsigned)__float_as_int(den) - 0x0d800000u > 0x64000000u);
  return a;
}
__device__ __forceinline__ float Sc<float>::atan2_core(float y, float x, bool& rare) {
  return fast_atan2f_core(y, x, rare);
}
__device__ __forceinline__ float Sc<float>::atan2_fast(float y, float x) {
  bool rare = false;
  float a = fast_atan2f_core(y, x, rare);
  if (rare) a = ::atan2f(y, x);
  return a;
}

__device__ __forceinline__ double Sc<double>::atan2_fast(double y, double x) { return fast_atan2(y, x); }
__device__ __forceinline__ double Sc<double>::atan2_core(double y, double x, bool& rare) {
  return fast_atan2_core(y, x, rare);
}
__device__ __forceinline__ double Sc<double>::rcp_fast(double x) { return fast_rcp(x); }
__device__ __forceinline__ double Sc<double>::rsqrt_fast(double x) { return fast_rsqrt(x); }
__device__ __forceinline__ void Sc<double>::sincos(double x, double* s, double* c) { fast_sincos(x, s, c); }
__device__ __forceinline__ double Sc<double>::sqrt_fast(double x) { return fast_sqrt(x); }

// ---------------------------------------------------------------------------------
// TMA bulk copies (1-D cp.async.bulk) + mbarrier: the SOA planes of a tile move between HBM
// and shared memory without occupying the LSU or registers, asynchronously to the math.
// Addresses and sizes must be multiples of 16 bytes.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RK_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra RK_DONE_%=;\n"
      "bra RK_WAIT_%=;\n"
      "RK_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 2-D tiled tensor maps: one TMA instruction moves a whole [rows][cols] box between the SOA
// planes in HBM and a dense shared-memory tile; out-of-range columns are zero-filled on load, so
// ragged LOAD tails need no special code.  Stores are clipped too, but at 16-byte granularity
// (measured: an 8-byte extent that ends on an odd element spills one element), so callers store
// whole tiles only.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(smem_u32(smem_dst)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int x, int y, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map),
               "r"(x), "r"(y), "r"(smem_u32(smem_src))
               : "memory");
}

// Warp index as a value the compiler knows to be warp-uniform (broadcast from lane 0): tile
// indices, shared-memory tile addresses and TMA coordinates derived from it then live in uniform
// registers instead of being re-uniformised by a loop before every TMA instruction.
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// Host: encode a [rows][cols] tensor with row pitch `pitch_bytes` and box [box_rows][box_cols].
// Returns false when the driver entry point is unavailable or the layout violates TMA's
// alignment rules (16-byte base and pitch) -- callers then use the LSU kernels.
inline bool make_tmap_2d(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* base,
                         uint64_t cols, uint64_t rows, uint64_t pitch_bytes, uint32_t box_cols,
                         uint32_t box_rows, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_NONE) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)p;
    cudaGetLastError();
  });
  if (!fn) return false;
  if (((uintptr_t)base & 15) || (pitch_bytes & 15) || ((uint64_t)box_cols * elem_bytes & 15)) return false;
  if (cols == 0 || cols > 0x7fffffffull || rows == 0) return false;
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {pitch_bytes};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swizzle, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename S>
struct TmapType;
template <>
struct TmapType<double> {
  static constexpr CUtensorMapDataType value = CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
};
template <>
struct TmapType<float> {
  static constexpr CUtensorMapDataType value = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
};

// 3-vectors as plain scalars keep everything in registers.
template <typename S>
struct V3 {
  S x, y, z;
};

template <typename S>
__device__ __forceinline__ V3<S> cross(const V3<S>& a, const V3<S>& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename S>
__device__ __forceinline__ V3<S> sub(const V3<S>& a, const V3<S>& b) {
  return {a.x - b.x, a.y - b.y, a.z - b.z};
}
template <typename S>
__device__ __forceinline__ V3<S> axpy(S a, const V3<S>& x, const V3<S>& y) {  // a*x + y
  return {a * x.x + y.x, a * x.y + y.y, a * x.z + y.z};
}
template <typename S>
__device__ __forceinline__ V3<S> lin2(S a, const V3<S>& x, S b, const V3<S>& y) {  // a*x + b*y
  return {a * x.x + b * y.x, a * x.y + b * y.y, a * x.z + b * y.z};
}
// R (row-major 3x3) * v
template <typename S>
__device__ __forceinline__ V3<S> rot(const S* R, const V3<S>& v) {
  return {R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z,
          R[6] * v.x + R[7] * v.y + R[8] * v.z};
}

// C(3x4) = A(3x4) * B(3x4) as homogeneous transforms, row-major, last row implied.
template <typename S>
__device__ __forceinline__ void xf_mul(const S* A, const S* B, S* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      S acc = A[i * 4 + 0] * B[0 * 4 + j] + A[i * 4 + 1] * B[1 * 4 + j] + A[i * 4 + 2] * B[2 * 4 + j];
      if (j == 3) acc += A[i * 4 + 3];
      C[i * 4 + j] = acc;
    }
  }
}

// Solve A x = b for a small dense system held in registers (N <= 6): Gaussian elimination with
// partial pivoting.  Every index is static -- the pivot row is brought up by conditional swaps, not
// by an indexed access -- so A and b never leave the register file.  x overwrites b; returns false
// when a pivot is exactly zero (the result then holds inf / nan, like a failed np.linalg.inv).
// (The loops are unrolled by template recursion: with `#pragma unroll` on the triangular loops nvcc
// 12.9 kept the matrix on the stack and indexed it dynamically -- 384 bytes of local memory and ~230
// LDL/STL per solve, 3.6 ms per 2^24 Fanuc poses.)
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}
template <typename S, int N>
__device__ __forceinline__ bool gauss_solve(S (&A)[N][N], S (&b)[N]) {
  bool ok = true;
  static_for<0, N>([&](auto kc) {
    constexpr int k = decltype(kc)::value;
    static_for<k + 1, N>([&](auto rc) {
      constexpr int r = decltype(rc)::value;
      const bool sw = Sc<S>::abs(A[r][k]) > Sc<S>::abs(A[k][k]);
      static_for<k, N>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        const S t = A[k][c];
        A[k][c] = sw ? A[r][c] : t;
        A[r][c] = sw ? t : A[r][c];
      });
      const S t = b[k];
      b[k] = sw ? b[r] : t;
      b[r] = sw ? t : b[r];
    });
    ok = ok && (A[k][k] != S(0));
    const S inv = S(1) / A[k][k];
    static_for<k + 1, N>([&](auto rc) {
      constexpr int r = decltype(rc)::value;
      const S f = A[r][k] * inv;
      static_for<k + 1, N>([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        A[r][c] -= f * A[k][c];
      });
      b[r] -= f * b[k];
    });
  });
  static_for<0, N>([&](auto kc) {
    constexpr int k = N - 1 - decltype(kc)::value;
    S acc = b[k];
    static_for<k + 1, N>([&](auto cc) {
      constexpr int c = decltype(cc)::value;
      acc -= A[k][c] * b[c];
    });
    b[k] = acc / A[k][k];
  });
  return ok;
}

// ---------------------------------------------------------------------------------
// Warp-private staging between the per-thread register layout and pose-major (AOS)
// global arrays: lane t owns NC consecutive components of item (warp_base + t).  The
// staging row stride is odd (in elements) so the per-lane accesses are bank-conflict
// free, and the global side is walked as one contiguous run -> full 128-byte lines.
// ---------------------------------------------------------------------------------
template <int NC>
struct Stride {
  static constexpr int value = (NC % 2 == 0) ? NC + 1 : NC;
};

template <typename S, int NC>
__device__ __forceinline__ void warp_store_aos(S* stage, const S (&v)[NC], S* g_warp_base,
                                               int n_items, int lane) {
  constexpr int ST = Stride<NC>::value;
  __syncwarp();
#pragma unroll
  for (int c = 0; c < NC; ++c) stage[lane * ST + c] = v[c];
  __syncwarp();
  const int total = n_items * NC;
  for (int idx = lane; idx < total; idx += kWarp) {
    int item = idx / NC, c = idx - item * NC;
    g_warp_base[idx] = stage[item * ST + c];
  }
}

// Same, with eight shared-memory loads in flight before their eight stores.  Helps the
// trajectory kernels (small NC, few live registers: Cartesian line + IK 1.56 -> 1.42 ms); it costs
// K1's AOS kernel, which is already at its register limit, 19 %.
template <typename S, int NC>
__device__ __forceinline__ void warp_store_aos_ilp(S* stage, const S (&v)[NC], S* g_warp_base,
                                                   int n_items, int lane) {
  constexpr int ST = Stride<NC>::value;
  __syncwarp();
#pragma unroll
  for (int c = 0; c < NC; ++c) stage[lane * ST + c] = v[c];
  __syncwarp();
  const int total = n_items * NC;
  if constexpr (sizeof(S) == 4 && NC % 2 == 0) {
    // FP32 runs: a 4-byte element per lane is 128 B per store instruction; with a full tile and an
    // 8-byte aligned run every lane writes element PAIRS (rrr_p2p_fk_f32: LSU data pipe 54 %)
    if (n_items == kWarp && (reinterpret_cast<uintptr_t>(g_warp_base) & 7) == 0) {
      float2 tmp2[NC / 2];
#pragma unroll
      for (int k = 0; k < NC / 2; ++k) {
        const int idx = 2 * (lane + k * kWarp);
        const int item = idx / NC, c = idx - item * NC;  // NC even: the pair stays inside one item
        tmp2[k] = make_float2(stage[item * ST + c], stage[item * ST + c + 1]);
      }
#pragma unroll
      for (int k = 0; k < NC / 2; ++k)
        reinterpret_cast<float2*>(g_warp_base)[lane + k * kWarp] = tmp2[k];
      return;
    }
  }
  constexpr int CH = NC < 8 ? NC : 8;
#pragma unroll
  for (int k0 = 0; k0 < NC; k0 += CH) {
    S tmp[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int idx = lane + (k0 + k) * kWarp;
      const int item = idx / NC, c = idx - item * NC;
      tmp[k] = (k0 + k < NC && idx < total) ? stage[item * ST + c] : S(0);
    }
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int idx = lane + (k0 + k) * kWarp;
      if (k0 + k < NC && idx < total) g_warp_base[idx] = tmp[k];
    }
  }
}

template <typename S, int NC>
__device__ __forceinline__ void warp_load_aos(S* stage, S (&v)[NC], const S* g_warp_base,
                                              int n_items, int lane) {
  constexpr int ST = Stride<NC>::value;
  __syncwarp();
  const int total = n_items * NC;
  // the warp's run is NC * 32 elements: NC independent loads in flight, then the stores (a rolled
  // loop would make every store wait for its own load before the next load is issued)
  S tmp[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const int idx = lane + k * kWarp;
    tmp[k] = idx < total ? g_warp_base[idx] : S(0);
  }
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const int idx = lane + k * kWarp;
    const int item = idx / NC, c = idx - item * NC;
    if (idx < total) stage[item * ST + c] = tmp[k];
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < NC; ++c) v[c] = stage[lane * ST + c];
}

}  // namespace rk
