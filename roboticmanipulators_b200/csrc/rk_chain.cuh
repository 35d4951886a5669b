// The compiled form of a token chain (see rk_chain.cu), shared by the interpreting kernel and the
// run-time specialiser (rk_chain_jit.cu).
#pragma once
#include <string>
#include "rk_common.cuh"

namespace rk {

constexpr int kMaxConst = 16;   // folded constant segments per chain

enum : int { ST_TRANS = 0, ST_ROT = 1, ST_CONST = 2, ST_CONST_T = 3, ST_DTRANS = 4, ST_DROT = 5 };

template <typename S>
struct ChainProg {
  int nstep, nvar, na, has_deriv, has_base, has_tool;
  int nslot;                     // staged values per sample: nvar + one cosine slot per variable rotation
  int8_t kind[RK_MAX_TOKENS], axis[RK_MAX_TOKENS], neg[RK_MAX_TOKENS], active[RK_MAX_TOKENS];
  int8_t code[RK_MAX_TOKENS];    // kind * 3 + axis (axis 0 for the constant kinds)
  int8_t slot[RK_MAX_TOKENS];    // variable steps: index into the staged values; constant steps: pool index
  int8_t cslot[RK_MAX_TOKENS];   // variable rotation steps: slot of the cosine parked by the first walk
  int16_t plane[RK_MAX_TOKENS];  // values column of variable slot v
  int8_t rows[6];                // solve mode: Jacobian rows that form the square system
  S X[kMaxConst][12];
  S Tb[12], Tt[12];
};


// Run-time specialisation (rk_chain_jit.cu): straight-line code for THIS chain, compiled once with
// NVRTC and cached.  Returns RK_OK when the kernel was launched, RK_ERR_UNSUPPORTED when there is no
// specialised kernel (NVRTC not present, compilation failed, chain too wide) -- the caller then runs
// the interpreting kernel.
template <typename S>
int chain_jit_launch(const ChainProg<S>& Pg, bool soa, const S* values, int nvals, int64_t n, int64_t ld, S* T,
                     S* J, cudaStream_t st);
// the CUDA source the specialiser would compile (diagnostics and the CPU-side build test)
template <typename S>
std::string chain_jit_source(const ChainProg<S>& Pg, bool soa, bool wantT, bool wantJ, bool vec2, bool solve);
// dq = J[Pg.rows, :]^-1 dx per sample, elimination unrolled over the chain's own columns
template <typename S>
int chain_jit_solve_launch(const ChainProg<S>& Pg, bool soa, const S* values, int nvals, int64_t n, int64_t ld,
                           const S* dx, S* dq, uint8_t* ok, cudaStream_t st);
const char* chain_jit_status();  // last NVRTC / driver message (diagnostics)

}  // namespace rk
