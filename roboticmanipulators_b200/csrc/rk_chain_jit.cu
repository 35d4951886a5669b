// Run-time specialisation of the generic token chain (Transformation / JacobianCalculator,
// utils/robo_math.py:364-397, utils/jacobians.py:166-240).
//
// The interpreting kernel (rk_chain.cu) pays ~2000 instructions per sample on the 13-token Fanuc
// sequence where the arithmetic is ~650: which column a step touches is DATA there, so the frame
// has to live in shared memory (or be shuffled between register assignments), every step decodes
// its tables, and the Jacobian needs a second walk.  For a given chain all of that is known before
// the first sample: this file writes the chain out as straight-line CUDA -- the frame in named
// registers, one walk, the screw axis / origin of every active token kept in registers until O_n is
// known, constants (folded 3x4 segments, base, tool) in the kernel parameter block so that their
// VALUES do not take part in the cache key -- compiles it once per (structure, dtype, layout,
// outputs) with NVRTC for sm_100a, and launches it through the driver API.  libnvrtc is looked up at
// run time; without it (or for very wide chains) callers keep the interpreting kernel.
#include <dlfcn.h>

#include <map>
#include <mutex>
#include <string>

#include "rk_chain.cuh"

namespace rk {
namespace {

std::mutex g_mu;
std::string g_status = "not used yet";

// ---- NVRTC, resolved with dlopen --------------------------------------------------------
struct Nvrtc {
  typedef struct _nvrtcProgram* Program;
  int (*createProgram)(Program*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*compileProgram)(Program, int, const char* const*) = nullptr;
  int (*getCUBINSize)(Program, size_t*) = nullptr;
  int (*getCUBIN)(Program, char*) = nullptr;
  int (*getProgramLogSize)(Program, size_t*) = nullptr;
  int (*getProgramLog)(Program, char*) = nullptr;
  int (*destroyProgram)(Program*) = nullptr;
  bool ok = false;
};

Nvrtc& nvrtc() {
  static Nvrtc N;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
      h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (h) break;
    }
    if (!h) return;
    auto sym = [&](const char* s) { return dlsym(h, s); };
    N.createProgram = (decltype(N.createProgram))sym("nvrtcCreateProgram");
    N.compileProgram = (decltype(N.compileProgram))sym("nvrtcCompileProgram");
    N.getCUBINSize = (decltype(N.getCUBINSize))sym("nvrtcGetCUBINSize");
    N.getCUBIN = (decltype(N.getCUBIN))sym("nvrtcGetCUBIN");
    N.getProgramLogSize = (decltype(N.getProgramLogSize))sym("nvrtcGetProgramLogSize");
    N.getProgramLog = (decltype(N.getProgramLog))sym("nvrtcGetProgramLog");
    N.destroyProgram = (decltype(N.destroyProgram))sym("nvrtcDestroyProgram");
    N.ok = N.createProgram && N.compileProgram && N.getCUBINSize && N.getCUBIN && N.getProgramLogSize &&
           N.getProgramLog && N.destroyProgram;
  });
  return N;
}

// ---- driver API entry points (the library links the runtime statically) ------------------
struct Driver {
  CUresult (*moduleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*moduleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                           CUstream, void**, void**) = nullptr;
  bool ok = false;
};

Driver& driver() {
  static Driver D;
  static std::once_flag once;
  std::call_once(once, [] {
    auto get = [](const char* name) -> void* {
      void* p = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        p = nullptr;
      cudaGetLastError();
      return p;
    };
    D.moduleLoadData = (decltype(D.moduleLoadData))get("cuModuleLoadData");
    D.moduleGetFunction = (decltype(D.moduleGetFunction))get("cuModuleGetFunction");
    D.launchKernel = (decltype(D.launchKernel))get("cuLaunchKernel");
    D.ok = D.moduleLoadData && D.moduleGetFunction && D.launchKernel;
  });
  return D;
}

// ---- source generation ---------------------------------------------------------------------
const char* kPreludeF64 = R"SRC(
typedef double S;
#define LD(p) ld_cs(p)
#define ST(p, v) st_cs(p, v)
typedef double2 S2;
#define LD2(p) ld_cs2(p)
#define ST2(p, a, b) st_cs2(p, a, b)
__device__ __forceinline__ double2 ld_cs2(const double* p) { double2 v; asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p)); return v; }
__device__ __forceinline__ void st_cs2(double* p, double a, double b) { asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" :: "l"(p), "d"(a), "d"(b) : "memory"); }
__device__ __forceinline__ double ld_cs(const double* p) { double v; asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
__device__ __forceinline__ void st_cs(double* p, double v) { asm volatile("st.global.cs.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory"); }
// sin and cos together (Cody-Waite by pi/2, degree-11/12 polynomials; <= ~1.5 ulp for |x| <= 1e5):
// the same routine as fast_sincos of rk_common.cuh
// (results by value and by reference: a pointer to a local would put every sin/cos pair through the
// stack -- STL + LDL on the main path, measured as the long_scoreboard stall of the first version)
__device__ __noinline__ double2 slow_sc(double x) { double s, c; sincos(x, &s, &c); return make_double2(s, c); }
__device__ __forceinline__ void rk_sincos(double x, double& so, double& co) {
  const double kf = rint(x * 0x1.45f306dc9c883p-1);
  const int q = (int)kf;
  double r = fma(-kf, 0x1.921fb54442d18p+0, x);
  r = fma(-kf, 0x1.1a62633145c07p-54, r);
  r = fma(-kf, -0x1.f1976b7ed8fbcp-110, r);
  const double z = r * r;
  double ps = 0x1.5dfb4dfaf15abp-33;
  ps = fma(ps, z, -0x1.ae5fdeb852566p-26);
  ps = fma(ps, z, 0x1.71de376b9c575p-19);
  ps = fma(ps, z, -0x1.a01a019e623afp-13);
  ps = fma(ps, z, 0x1.1111111110b23p-7);
  ps = fma(ps, z, -0x1.5555555555555p-3);
  double pc = -0x1.906dd38cb66a1p-37;
  pc = fma(pc, z, 0x1.1eeb52bea9072p-29);
  pc = fma(pc, z, -0x1.27e4fa025e03dp-22);
  pc = fma(pc, z, 0x1.a01a019f3df42p-16);
  pc = fma(pc, z, -0x1.6c16c16c1691fp-10);
  pc = fma(pc, z, 0x1.5555555555555p-5);
  const double sn = fma(r * z, ps, r);
  const double cs = fma(z * z, pc, fma(-0.5, z, 1.0));
  double s = (q & 1) ? cs : sn;
  double c = (q & 1) ? sn : cs;
  if (q & 2) s = -s;
  if ((q + 1) & 2) c = -c;
  if (!(fabs(x) <= 1.0e5)) { const double2 sc = slow_sc(x); s = sc.x; c = sc.y; }
  so = s;
  co = c;
}
)SRC";

const char* kPreludeF32 = R"SRC(
typedef float S;
#define LD(p) ld_cs(p)
#define ST(p, v) st_cs(p, v)
typedef float2 S2;
#define LD2(p) ld_cs2(p)
#define ST2(p, a, b) st_cs2(p, a, b)
__device__ __forceinline__ float2 ld_cs2(const float* p) { float2 v; asm volatile("ld.global.cs.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p)); return v; }
__device__ __forceinline__ void st_cs2(float* p, float a, float b) { asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(a), "f"(b) : "memory"); }
__device__ __forceinline__ float ld_cs(const float* p) { float v; asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
__device__ __forceinline__ void st_cs(float* p, float v) { asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void rk_sincos(float x, float& so, float& co) { float s, c; sincosf(x, &s, &c); so = s; co = c; }
)SRC";

// The per-sample work is generated once as text in which every per-sample name ends in '@'; the
// kernel body instantiates it once (scalar) or twice ('a', 'b': two neighbouring samples per thread,
// 16-byte loads and stores on the planes -- half the LSU instructions of a plane-bound T-only walk).
std::string subst(const std::string& text, const char* suffix) {
  std::string out;
  out.reserve(text.size() + 64);
  for (char ch : text) {
    if (ch == '@') out += suffix;
    else out.push_back(ch);
  }
  return out;
}

// resident CTAs per SM the register allocation is held to (RK_CHAIN_MINB overrides).  Measured on the
// Fanuc sequence, 2^24 samples: the walks are fastest uncapped (T + J 1.41 / 1.42 / 1.45 / 1.48 ms at
// 1 / 3 / 4 / 5), the dq solve at 4 (0.885 uncapped, 0.881 / 0.819 / 0.951 ms at 3 / 4 / 5).
int min_blocks(bool solve) {
  static const int env = [] {
    const char* e = getenv("RK_CHAIN_MINB");
    const int m = e ? atoi(e) : 0;
    return m < 0 ? 0 : (m > 16 ? 16 : m);
  }();
  return env ? env : (solve ? 4 : 1);
}

template <typename S>
std::string generate(const ChainProg<S>& Pg, bool soa, bool wantT, bool wantJ, bool vec2, bool solve = false) {
  if (solve) wantT = false, wantJ = true, vec2 = false;  // J[rows, :] dq = dx, solved in registers
  std::string src = sizeof(S) == 8 ? kPreludeF64 : kPreludeF32;
  std::string body;
  auto num = [](long long v) { return std::to_string(v); };
  auto R = [&](int i, int j) { return "r" + num(i) + num(j) + "@"; };
  auto P = [&](int i) { return "p" + num(i) + "@"; };
  const std::string per = vec2 ? "256" : "128";   // samples per CTA and trip
  src += "struct Consts { S X[" + num(kMaxConst) + "][12]; S Tb[12]; S Tt[12]; };\n";
  src +=
      "extern \"C\" __global__ void __launch_bounds__(128, " + num(min_blocks(solve)) + ") rk_chain_jit(const S* __restrict__ values, int nvals,\n"
      "    long long n, long long ld, S* __restrict__ T, S* __restrict__ J, const Consts C,\n"
      "    const S* __restrict__ dx, S* __restrict__ dq, unsigned char* __restrict__ ok) {\n"
      "  const long long stride = (long long)gridDim.x * " + per + ";\n"
      "  long long i = ((long long)blockIdx.x * 128 + threadIdx.x)" + (vec2 ? " * 2" : "") + ";\n";
  // software pipeline: the next trip's values are in flight while this trip computes (the K1 LSU
  // kernel spent 55 % of its warp time waiting for its six loads before it did the same)
  auto load_vals = [&](const std::string& idx, bool declare) {
    for (int v = 0; v < Pg.nvar; ++v) {
      const std::string pl = num(Pg.plane[v]);
      src += std::string("    ") + (declare ? (vec2 ? "S2 " : "S ") : "") + "nv" + num(v) + " = " +
             (vec2 ? "LD2(values + " + pl + "ll * ld + " + idx + ");\n"
                   : soa ? "LD(values + " + pl + "ll * ld + " + idx + ");\n"
                         : "values[" + idx + " * nvals + " + pl + "];\n");
    }
  };
  src += "  if (i >= n) return;\n";
  load_vals("i", true);
  src += "  for (; i < n; i += stride) {\n";
  for (int v = 0; v < Pg.nvar; ++v) {
    if (vec2)
      src += "    const S v" + num(v) + "a = nv" + num(v) + ".x, v" + num(v) + "b = nv" + num(v) + ".y;\n";
    else
      src += "    const S v" + num(v) + " = nv" + num(v) + ";\n";
  }
  src += "    { const long long in = i + stride; if (in < n) {\n";
  load_vals("in", false);
  src += "    } }\n";
  if (solve)  // the right-hand side is requested before the walk and needed after it
    for (int r = 0; r < Pg.na; ++r)
      src += "    S b" + num(r) + " = " +
             (soa ? "LD(dx + " + num(r) + "ll * ld + i);\n" : "dx[i * " + num(Pg.na) + " + " + num(r) + "];\n");
  // frame
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      body += "    S " + R(i, j) + " = " +
              (Pg.has_base ? "C.Tb[" + num(i * 4 + j) + "]" : std::string(i == j ? "S(1)" : "S(0)")) + ";\n";
    body += "    S " + P(i) + " = " + (Pg.has_base ? "C.Tb[" + num(i * 4 + 3) + "]" : std::string("S(0)")) + ";\n";
  }
  auto mul12 = [&](const std::string& X) {  // F <- F * X (3x4 constant in the parameter block)
    body += "    {\n";
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j)
        body += "      const S t" + num(i) + num(j) + " = " + R(i, 0) + " * " + X + "[" + num(j) + "] + " + R(i, 1) +
                " * " + X + "[" + num(4 + j) + "] + " + R(i, 2) + " * " + X + "[" + num(8 + j) + "];\n";
      body += "      " + P(i) + " += " + R(i, 0) + " * " + X + "[3] + " + R(i, 1) + " * " + X + "[7] + " + R(i, 2) +
              " * " + X + "[11];\n";
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) body += "      " + R(i, j) + " = t" + num(i) + num(j) + ";\n";
    body += "    }\n";
  };
  int col = 0;
  for (int st = 0; st < Pg.nstep; ++st) {
    const int kind = Pg.kind[st], A = Pg.axis[st], I = (A + 1) % 3, Jc = (A + 2) % 3, sl = Pg.slot[st];
    const std::string val = std::string(Pg.neg[st] ? "-" : "") +
                            (sl >= 0 ? "v" + num(sl) + "@" : "C.X[" + num(-1 - sl) + "][0]");
    if (Pg.active[st] && wantJ) {  // the frame BEFORE an active token: its screw axis and origin
      const std::string k = num(col);
      body += "    const S u" + k + "x@ = " + R(0, A) + ", u" + k + "y@ = " + R(1, A) + ", u" + k + "z@ = " + R(2, A) +
              ";\n";
      if (kind == ST_ROT)
        body += "    const S o" + k + "x@ = p0@, o" + k + "y@ = p1@, o" + k + "z@ = p2@;\n";
    }
    if (Pg.active[st]) ++col;
    switch (kind) {
      case ST_TRANS:
        body += "    { const S v = " + val + ";";
        for (int r = 0; r < 3; ++r) body += " " + P(r) + " = fma(v, " + R(r, A) + ", " + P(r) + ");";
        body += " }\n";
        break;
      case ST_ROT:
      case ST_DROT:
        body += "    { S s, c; rk_sincos(" + val + ", s, c);\n";
        for (int r = 0; r < 3; ++r) {
          body += "      { const S a = " + R(r, I) + ", b = " + R(r, Jc) + "; ";
          if (kind == ST_ROT)
            body += R(r, I) + " = c * a + s * b; " + R(r, Jc) + " = c * b - s * a; }\n";
          else
            body += R(r, I) + " = c * b - s * a; " + R(r, Jc) + " = -c * a - s * b; " + R(r, A) + " = S(0); " + P(r) +
                    " = S(0); }\n";
        }
        body += "    }\n";
        break;
      case ST_CONST: mul12("C.X[" + num(sl) + "]"); break;
      case ST_CONST_T:
        for (int r = 0; r < 3; ++r)
          body += "    " + P(r) + " += " + R(r, 0) + " * C.X[" + num(sl) + "][3] + " + R(r, 1) + " * C.X[" + num(sl) +
                  "][7] + " + R(r, 2) + " * C.X[" + num(sl) + "][11];\n";
        break;
      default:  // ST_DTRANS
        for (int r = 0; r < 3; ++r) body += "    " + P(r) + " = " + R(r, A) + ";\n";
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) body += "    " + R(i, j) + " = S(0);\n";
    }
  }
  if (Pg.has_tool) mul12("C.Tt");
  if (wantJ) {  // lever arms of the revolute columns
    int a = 0;
    for (int st = 0; st < Pg.nstep; ++st) {
      if (!Pg.active[st]) continue;
      const std::string k = num(a++);
      if (Pg.kind[st] == ST_ROT)
        body += "    const S dx" + k + "@ = p0@ - o" + k + "x@, dy" + k + "@ = p1@ - o" + k + "y@, dz" + k +
                "@ = p2@ - o" + k + "z@;\n";
    }
  }
  if (vec2)
    src += subst(body, "a") + subst(body, "b");
  else
    src += subst(body, "");
  auto store = [&](const std::string& base, long long comp, long long ncomp, const std::string& v) {
    if (vec2)
      src += "    ST2(" + base + " + " + num(comp) + "ll * ld + i, " + subst(v, "a") + ", " + subst(v, "b") + ");\n";
    else if (soa)
      src += "    ST(" + base + " + " + num(comp) + "ll * ld + i, " + subst(v, "") + ");\n";
    else
      src += "    " + base + "[i * " + num(ncomp) + " + " + num(comp) + "] = " + subst(v, "") + ";\n";
  };
  if (wantT) {
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) store("T", i * 4 + j, 16, R(i, j));
      store("T", i * 4 + 3, 16, P(i));
    }
    store("T", 12, 16, "S(0)"), store("T", 13, 16, "S(0)"), store("T", 14, 16, "S(0)");
    store("T", 15, 16, Pg.has_deriv ? "S(0)" : "S(1)");
  }
  if (solve) {
    // entry (row, column a) of the Jacobian as an expression of the walk's registers
    const int N = Pg.na;
    auto entry = [&](int row, int a, int kind) -> std::string {
      const std::string k = num(a);
      const char* ax[3] = {"x", "y", "z"};
      if (kind != ST_ROT) return row < 3 ? "u" + k + ax[row] : std::string("S(0)");
      if (row >= 3) return "u" + k + ax[row - 3];
      const int i1 = (row + 1) % 3, i2 = (row + 2) % 3;
      return "u" + k + ax[i1] + " * d" + ax[i2] + k + " - u" + k + ax[i2] + " * d" + ax[i1] + k;
    };
    auto A = [&](int r, int c) { return "a" + num(r) + "_" + num(c); };
    int a = 0;
    for (int st = 0; st < Pg.nstep; ++st) {
      if (!Pg.active[st]) continue;
      for (int r = 0; r < N; ++r) src += "    S " + A(r, a) + " = " + entry(Pg.rows[r], a, Pg.kind[st]) + ";\n";
      ++a;
    }
    // Gaussian elimination with partial pivoting, the same operation order as gauss_solve<S, N>
    // of rk_common.cuh (row exchanges are selects, nothing is indexed)
    src += "    bool good = true;\n";
    for (int k = 0; k < N; ++k) {
      for (int r = k + 1; r < N; ++r) {
        src += "    { const bool sw = fabs(" + A(r, k) + ") > fabs(" + A(k, k) + ");\n";
        for (int c = k; c < N; ++c)
          src += "      { const S t = " + A(k, c) + "; " + A(k, c) + " = sw ? " + A(r, c) + " : t; " + A(r, c) +
                 " = sw ? t : " + A(r, c) + "; }\n";
        src += "      { const S t = b" + num(k) + "; b" + num(k) + " = sw ? b" + num(r) + " : t; b" + num(r) +
               " = sw ? t : b" + num(r) + "; } }\n";
      }
      src += "    good = good && (" + A(k, k) + " != S(0));\n";
      src += "    { const S inv = S(1) / " + A(k, k) + ";\n";
      for (int r = k + 1; r < N; ++r) {
        src += "      { const S f = " + A(r, k) + " * inv;";
        for (int c = k + 1; c < N; ++c) src += " " + A(r, c) + " -= f * " + A(k, c) + ";";
        src += " b" + num(r) + " -= f * b" + num(k) + "; }\n";
      }
      src += "    }\n";
    }
    for (int k = N - 1; k >= 0; --k) {
      src += "    { S acc = b" + num(k) + ";";
      for (int c = k + 1; c < N; ++c) src += " acc -= " + A(k, c) + " * b" + num(c) + ";";
      src += " b" + num(k) + " = acc / " + A(k, k) + "; }\n";
    }
    for (int r = 0; r < N; ++r)
      src += soa ? "    ST(dq + " + num(r) + "ll * ld + i, b" + num(r) + ");\n"
                 : "    dq[i * " + num(N) + " + " + num(r) + "] = b" + num(r) + ";\n";
    src += "    if (ok) ok[i] = good ? 1 : 0;\n";
  } else if (wantJ) {
    int a = 0;
    for (int st = 0; st < Pg.nstep; ++st) {
      if (!Pg.active[st]) continue;
      const std::string k = num(a);
      const char* ax[3] = {"x@", "y@", "z@"};
      if (Pg.kind[st] == ST_ROT) {  // [u x (O_n - O_k); u]
        store("J", 0 * Pg.na + a, 6 * Pg.na, "u" + k + "y@ * dz" + k + "@ - u" + k + "z@ * dy" + k + "@");
        store("J", 1 * Pg.na + a, 6 * Pg.na, "u" + k + "z@ * dx" + k + "@ - u" + k + "x@ * dz" + k + "@");
        store("J", 2 * Pg.na + a, 6 * Pg.na, "u" + k + "x@ * dy" + k + "@ - u" + k + "y@ * dx" + k + "@");
        for (int r = 0; r < 3; ++r) store("J", (3 + r) * Pg.na + a, 6 * Pg.na, "u" + k + ax[r]);
      } else {                      // prismatic: [u; 0]
        for (int r = 0; r < 3; ++r) store("J", r * Pg.na + a, 6 * Pg.na, "u" + k + ax[r]);
        for (int r = 3; r < 6; ++r) store("J", r * Pg.na + a, 6 * Pg.na, "S(0)");
      }
      ++a;
    }
  }
  src += "  }\n}\n";
  return src;
}

struct Entry {
  CUfunction fn = nullptr;
  bool failed = false;
};
std::map<std::pair<int, std::string>, Entry> g_cache;

CUfunction build(const std::string& src) {
  Nvrtc& N = nvrtc();
  Driver& D = driver();
  if (!N.ok) {
    g_status = "libnvrtc not found";
    return nullptr;
  }
  if (!D.ok) {
    g_status = "driver entry points not available";
    return nullptr;
  }
  Nvrtc::Program prog = nullptr;
  if (N.createProgram(&prog, src.c_str(), "rk_chain_jit.cu", 0, nullptr, nullptr) != 0) {
    g_status = "nvrtcCreateProgram failed";
    return nullptr;
  }
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
  const int rc = N.compileProgram(prog, 3, opts);
  if (rc != 0) {
    size_t ls = 0;
    N.getProgramLogSize(prog, &ls);
    std::string log(ls, '\0');
    if (ls) N.getProgramLog(prog, &log[0]);
    g_status = "nvrtcCompileProgram failed: " + log.substr(0, 2000);
    N.destroyProgram(&prog);
    return nullptr;
  }
  size_t sz = 0;
  N.getCUBINSize(prog, &sz);
  std::string cubin(sz, '\0');
  N.getCUBIN(prog, &cubin[0]);
  N.destroyProgram(&prog);
  cudaFree(nullptr);  // make sure the primary context is current
  CUmodule mod = nullptr;
  CUfunction fn = nullptr;
  if (D.moduleLoadData(&mod, cubin.data()) != CUDA_SUCCESS ||
      D.moduleGetFunction(&fn, mod, "rk_chain_jit") != CUDA_SUCCESS) {
    g_status = "cuModuleLoadData / cuModuleGetFunction failed";
    return nullptr;
  }
  g_status = "ok";
  return fn;  // the module stays loaded for the life of the process
}

}  // namespace

const char* chain_jit_status() {
  std::lock_guard<std::mutex> lock(g_mu);
  static std::string copy;
  copy = g_status;
  return copy.c_str();
}

template <typename S>
static int jit_run(const ChainProg<S>& Pg, bool soa, const S* values, int nvals, int64_t n, int64_t ld, S* T, S* J,
                   const S* dx, S* dq, uint8_t* ok, bool solve, cudaStream_t st) {
  static const bool disabled = [] {
    const char* e = getenv("RK_CHAIN_JIT");
    return e && e[0] == '0';
  }();
  if (disabled) return RK_ERR_UNSUPPORTED;
  // the axes and origins of the active tokens stay in registers until the end of the walk
  if (Pg.na > 12 || Pg.nstep > RK_MAX_TOKENS) return RK_ERR_UNSUPPORTED;
  const bool wantT = !solve && T != nullptr, wantJ = solve || (J != nullptr && Pg.na > 0);
  // two samples per thread (8-byte plane accesses for FP32, which are otherwise 128 B per warp and
  // LSU-instruction bound: 0.84 -> 0.72 ms T+J, 0.33 -> 0.31 ms T on 2^24 samples; FP64 planes are
  // already 256 B per warp access and measured no faster): RK_CHAIN_VEC2 = 0 never, 1 FP32 walks
  // (default), 2 every SOA walk
  static const int vec2_mode = [] {
    const char* e = getenv("RK_CHAIN_VEC2");
    return e ? atoi(e) : 1;
  }();
  const bool vec2 = !solve && soa && vec2_mode > 0 && (vec2_mode > 1 || sizeof(S) == 4) && n % 2 == 0 && ld % 2 == 0 &&
                    (((uintptr_t)values | (uintptr_t)T | (uintptr_t)J) & 15) == 0;
  int dev = 0;
  cudaGetDevice(&dev);
  CUfunction fn = nullptr;
  {
    // the cache key is the STRUCTURE of the program (not its constants): a few dozen bytes
    std::string key;
    key.reserve(16 + 8 * Pg.nstep);
    key.push_back((char)sizeof(S)), key.push_back((char)soa), key.push_back((char)wantT), key.push_back((char)wantJ);
    key.push_back((char)Pg.has_base), key.push_back((char)Pg.has_tool), key.push_back((char)Pg.has_deriv);
    key.push_back((char)Pg.nstep), key.push_back((char)Pg.nvar), key.push_back((char)vec2);
    key.push_back((char)solve);
    if (solve)
      for (int r = 0; r < Pg.na; ++r) key.push_back((char)Pg.rows[r]);
    for (int st2 = 0; st2 < Pg.nstep; ++st2) {
      key.push_back((char)Pg.kind[st2]), key.push_back((char)Pg.axis[st2]), key.push_back((char)Pg.neg[st2]);
      key.push_back((char)Pg.active[st2]), key.push_back((char)Pg.slot[st2]);
    }
    for (int v = 0; v < Pg.nvar; ++v) key.push_back((char)(Pg.plane[v] & 0xff)), key.push_back((char)(Pg.plane[v] >> 8));
    std::lock_guard<std::mutex> lock(g_mu);
    Entry& e = g_cache[{dev, key}];
    if (!e.fn && !e.failed) {
      e.fn = build(generate(Pg, soa, wantT, wantJ, vec2, solve));
      e.failed = e.fn == nullptr;
    }
    fn = e.fn;
  }
  if (!fn) return RK_ERR_UNSUPPORTED;
  struct Consts {
    S X[kMaxConst][12];
    S Tb[12];
    S Tt[12];
  } C;
  memcpy(C.X, Pg.X, sizeof(C.X));
  memcpy(C.Tb, Pg.Tb, sizeof(C.Tb));
  memcpy(C.Tt, Pg.Tt, sizeof(C.Tt));
  long long nn = n, lld = ld;
  void* args[] = {(void*)&values, (void*)&nvals, (void*)&nn,  (void*)&lld, (void*)&T,
                  (void*)&J,      (void*)&C,     (void*)&dx,  (void*)&dq,  (void*)&ok};
  // same grid policy as stream_grid: ~8 trips per thread for large batches
  const int64_t cap = (int64_t)sm_count() * 8, per = vec2 ? 256 : 128;
  const int64_t need = (n + per * 8 - 1) / (per * 8);
  int64_t grid = need >= 4 * cap ? need : ((n + per - 1) / per < cap ? (n + per - 1) / per : cap);
  if (grid < 1) grid = 1;
  if (grid > 0x7fffffff) grid = 0x7fffffff;
  if (driver().launchKernel(fn, (unsigned)grid, 1, 1, 128, 1, 1, 0, (CUstream)st, args, nullptr) != CUDA_SUCCESS) {
    std::lock_guard<std::mutex> lock(g_mu);
    g_status = "cuLaunchKernel failed";
    return RK_ERR_CUDA;
  }
  return RK_OK;
}

template <typename S>
int chain_jit_launch(const ChainProg<S>& Pg, bool soa, const S* values, int nvals, int64_t n, int64_t ld, S* T,
                     S* J, cudaStream_t st) {
  return jit_run<S>(Pg, soa, values, nvals, n, ld, T, J, nullptr, nullptr, nullptr, false, st);
}
template <typename S>
int chain_jit_solve_launch(const ChainProg<S>& Pg, bool soa, const S* values, int nvals, int64_t n, int64_t ld,
                           const S* dx, S* dq, uint8_t* ok, cudaStream_t st) {
  if (Pg.na < 1 || Pg.na > 6) return RK_ERR_UNSUPPORTED;
  return jit_run<S>(Pg, soa, values, nvals, n, ld, nullptr, nullptr, dx, dq, ok, true, st);
}
template int chain_jit_solve_launch<double>(const ChainProg<double>&, bool, const double*, int, int64_t, int64_t,
                                            const double*, double*, uint8_t*, cudaStream_t);

template <typename S>
std::string chain_jit_source(const ChainProg<S>& Pg, bool soa, bool wantT, bool wantJ, bool vec2, bool solve) {
  if (solve && (Pg.na < 1 || Pg.na > 6)) return std::string();
  return generate(Pg, soa, wantT, wantJ && Pg.na > 0, vec2 && soa, solve);
}
template std::string chain_jit_source<double>(const ChainProg<double>&, bool, bool, bool, bool, bool);
template std::string chain_jit_source<float>(const ChainProg<float>&, bool, bool, bool, bool, bool);

template int chain_jit_launch<double>(const ChainProg<double>&, bool, const double*, int, int64_t, int64_t, double*,
                                      double*, cudaStream_t);
template int chain_jit_launch<float>(const ChainProg<float>&, bool, const float*, int, int64_t, int64_t, float*,
                                     float*, cudaStream_t);

}  // namespace rk

extern "C" const char* rk_chain_jit_status(void) { return rk::chain_jit_status(); }
