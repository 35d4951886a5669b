// Trapezoidal trajectory sampling: one warp per trajectory, lanes stride over samples.
// Reference: utils/trapezoidal_generator.py:36-227, utils/trajectory_generator.py:71-188,
// robots/rrr_robot.py:203-280.  Every expression that decides a sample count or a sample
// value is evaluated with explicit round-to-nearest mul/add/div in the reference's
// operation order (no FMA contraction), so counts and samples are bit-identical to NumPy.
#include "rk_arms.cuh"

namespace rk {

using M = Sc<double>;

// generate_coefficients (trapezoidal_generator.py:36-79)
__device__ __forceinline__ void trap_coefficients(double q0, double qf, double vmax, double amax,
                                                  double freq, double& t1, double& tau) {
  const double dq = fabs(M::sub(qf, q0));
  const double vp = M::sqrt(M::mul(dq, amax));
  if (vp < vmax) {  // triangular (:58-61)
    t1 = M::div(vp, amax);
    tau = t1;
  } else {          // trapezoidal (:63-66)
    t1 = M::div(vmax, amax);
    tau = M::div(dq, vmax);
  }
  if (freq != 0.0) {  // :71-77
    t1 = M::div(ceil(M::mul(t1, freq)), freq);
    tau = M::div(ceil(M::mul(tau, freq)), freq);
  }
}

// _get_times (trapezoidal_generator.py:208-227) + np.linspace semantics
struct TrapTimes {
  double t1, tau, tend;              // tend = tau + t1
  double st_acc, st_flat, st_dec;    // linspace steps
  double tl;                         // t_acc[-1]
  double tfl;                        // t_no_acc[-1] (if Lf > 0)
  int La, Lf, L;
};

__device__ __forceinline__ TrapTimes trap_times(double t1, double tau, int n) {
  TrapTimes tt;
  tt.t1 = t1, tt.tau = tau;
  tt.tend = M::add(tau, t1);
  const double step = M::div((double)n, tt.tend);           // :222
  const double fa = M::mul(step, t1);
  const double ff = M::mul(step, M::sub(tau, t1));
  // int() truncates toward zero; NaN/inf (zero-length motion) -> the reference raises
  const bool bad = !(fa == fa) || !(ff == ff) || fabs(fa) > 2e9 || fabs(ff) > 2e9;
  tt.La = bad ? 0 : (int)fa;                                 // :224, :226
  tt.Lf = bad ? 0 : (int)ff;                                 // :225
  if (tt.La < 0) tt.La = 0;
  if (tt.Lf < 0) tt.Lf = 0;
  tt.L = (tt.La == 0) ? 0 : 2 * tt.La + tt.Lf;               // empty t_acc -> IndexError (:142)
  tt.st_acc = (tt.La > 1) ? M::div(M::sub(t1, 0.0), (double)(tt.La - 1)) : 0.0;
  tt.st_flat = (tt.Lf > 1) ? M::div(M::sub(tau, t1), (double)(tt.Lf - 1)) : 0.0;
  tt.st_dec = (tt.La > 1) ? M::div(M::sub(tt.tend, tau), (double)(tt.La - 1)) : 0.0;
  tt.tl = (tt.La > 1) ? t1 : 0.0;
  tt.tfl = (tt.Lf > 1) ? tau : t1;
  return tt;
}

// np.linspace(a, b, L)[k]: k*step + a, last element forced to b, [a] when L == 1
__device__ __forceinline__ double linspace_at(double a, double b, double step, int L, int k) {
  if (L > 1 && k == L - 1) return b;
  return M::add(M::mul((double)k, step), a);
}

// sample i of the profile: segment, the local time used by the formulas, the output time
__device__ __forceinline__ void trap_sample_time(const TrapTimes& tt, int i, int& seg, double& tk,
                                                 double& tout) {
  if (i < tt.La) {
    seg = 0;
    tk = linspace_at(0.0, tt.t1, tt.st_acc, tt.La, i);
    tout = tk;
  } else if (i < tt.La + tt.Lf) {
    seg = 1;
    tout = linspace_at(tt.t1, tt.tau, tt.st_flat, tt.Lf, i - tt.La);
    tk = M::sub(tout, tt.tl);                                 // :142
  } else {
    seg = 2;
    const int k = i - tt.La - tt.Lf;
    tk = linspace_at(0.0, tt.t1, tt.st_acc, tt.La, k);        // get_q/get_dq reuse t_acc
    tout = linspace_at(tt.tau, tt.tend, tt.st_dec, tt.La, k);
  }
}

// get_q / get_dq / get_ddq (trapezoidal_generator.py:125-206) for one joint
__device__ __forceinline__ void trap_sample_joint(const TrapTimes& tt, int seg, double tk, double q0,
                                                  double v, double a, double& q, double& dq,
                                                  double& ddq) {
  const double ha = M::mul(0.5, a);
  const double q_acc_last = M::add(q0, M::mul(ha, M::mul(tt.tl, tt.tl)));   // q_acc[-1]
  if (seg == 0) {
    q = M::add(q0, M::mul(ha, M::mul(tk, tk)));                             // :147
    dq = M::mul(a, tk);                                                     // :182
    ddq = a;                                                                // :202
  } else if (seg == 1) {
    q = M::add(q_acc_last, M::mul(v, tk));                                  // :148
    dq = v;                                                                 // :183
    ddq = 0.0;                                                              // :203
  } else {
    const double start_q = (tt.Lf != 0)
                               ? M::add(q_acc_last, M::mul(v, M::sub(tt.tfl, tt.tl)))
                               : q_acc_last;                                // :151-154
    q = M::sub(M::add(start_q, M::mul(v, tk)), M::mul(ha, M::mul(tk, tk)));  // :156-157
    dq = M::sub(M::mul(a, tt.tl), M::mul(a, tk));                           // :184
    ddq = -a;                                                               // :204
  }
}

// Equalised profile of one trajectory (trajectory_generator.py:118-153)
template <int NJ>
struct Equalized {
  TrapTimes tt;
  double q0[NJ], v[NJ], a[NJ];
};

template <int NJ>
__device__ __forceinline__ Equalized<NJ> equalize(const double* __restrict__ q0,
                                                  const double* __restrict__ qf, double vmax,
                                                  double amax, double freq, int n) {
  Equalized<NJ> e;
  double t1m = 0.0, taum = 0.0, qfv[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    e.q0[j] = q0[j];
    qfv[j] = qf[j];
    double t1, tau;
    trap_coefficients(e.q0[j], qfv[j], vmax, amax, freq, t1, tau);
    t1m = (j == 0 || t1 > t1m) ? t1 : t1m;      // max() (:144-145)
    taum = (j == 0 || tau > taum) ? tau : taum;
  }
  e.tt = trap_times(t1m, taum, n);
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const double dq = M::sub(qfv[j], e.q0[j]);
    e.v[j] = M::div(dq, taum);                                 // :150
    e.a[j] = M::div(dq, M::mul(taum, t1m));                    // :151
  }
  return e;
}

// mode 0: joint-space sampler; 1: + RRR FK/Jacobian per sample; 2: Cartesian + RRR IK + dq.
// SO = type of the OUTPUTS and of the per-sample FK / Jacobian (mode 1).  The profile itself --
// coefficients, sample counts, sample times, q / dq / ddq -- is always evaluated in FP64 in the
// reference's operation order (counts must be bit-exact), and rounded to SO on the way out.
template <int NJ, int MODE, typename SO>
__global__ void __launch_bounds__(kBlock, MODE == 2 ? 5 : (MODE == 1 ? (sizeof(SO) == 4 ? 4 : 2) : 3))
traj_kernel(const __grid_constant__ ArmPT<SO> P, const double* __restrict__ q0, const double* __restrict__ qf,
            int64_t ntraj, int nsteps, double vmax, double amax, double freq, SO* __restrict__ q,
            SO* __restrict__ dq, SO* __restrict__ ddq, SO* __restrict__ t,
            int32_t* __restrict__ len, SO* __restrict__ o1, SO* __restrict__ o2,
            uint8_t* __restrict__ valid) {
  constexpr int NST = (MODE == 1) ? Stride<18>::value : Stride<NJ>::value;
  __shared__ SO stage_all[(kBlock / kWarp) * kWarp * NST];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  SO* stage = stage_all + warp * kWarp * NST;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  ArmPT<SO> bare = P;  // get_symbolic_jacobian_inverse uses identity base/tool (rrr_robot_jacobians.py:19-20)
  bare.has_base = bare.has_tool = 0;
  for (int64_t tr = (((int64_t)blockIdx.x * blockDim.x) >> 5) + warp; tr < ntraj; tr += warps) {
    const Equalized<NJ> e = equalize<NJ>(q0 + tr * NJ, qf + tr * NJ, vmax, amax, freq, nsteps);
    if (lane == 0) len[tr] = e.tt.L;
    for (int base = 0; base < nsteps; base += kWarp) {
      const int i = base + lane;
      const int cnt = (nsteps - base) < kWarp ? (nsteps - base) : kWarp;
      double qv[NJ], dqv[NJ], ddqv[NJ], tout = 0.0;
#pragma unroll
      for (int j = 0; j < NJ; ++j) qv[j] = dqv[j] = ddqv[j] = 0.0;
      const bool live = i < e.tt.L;
      if (live) {
        int seg;
        double tk;
        trap_sample_time(e.tt, i, seg, tk, tout);
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          trap_sample_joint(e.tt, seg, tk, e.q0[j], e.v[j], e.a[j], qv[j], dqv[j], ddqv[j]);
      }
      const int64_t s0 = tr * nsteps + base;  // first sample of this warp pass
      SO qo[NJ], dqo[NJ], ddqo[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) qo[j] = (SO)qv[j], dqo[j] = (SO)dqv[j], ddqo[j] = (SO)ddqv[j];
      if constexpr (MODE == 2) {
        // 3-wide rows: every lane writes its own 24 bytes (three strided 8-byte stores that the L2
        // merges into full sectors); the staged transpose costs ~30 instructions per array here and
        // this kernel is issue-bound: 1.184 -> 1.142 ms per 2^15 x 1000 samples
        if (lane < cnt) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            if (q) q[(s0 + lane) * NJ + j] = qo[j];
            if (dq) dq[(s0 + lane) * NJ + j] = dqo[j];
          }
        }
      } else {
        if (q) warp_store_aos_ilp<SO, NJ>(stage, qo, q + s0 * NJ, cnt, lane);
        if (dq) warp_store_aos_ilp<SO, NJ>(stage, dqo, dq + s0 * NJ, cnt, lane);
        if (ddq) warp_store_aos_ilp<SO, NJ>(stage, ddqo, ddq + s0 * NJ, cnt, lane);
      }
      if (t && lane < cnt) t[s0 + lane] = (SO)tout;
      if constexpr (MODE == 1) {
        SO Tv[16], Jv[18];
        rrr_fk_jac_pose(P, reinterpret_cast<const SO(&)[3]>(qo), Tv, Jv);
        if (!live) {
#pragma unroll
          for (int k = 0; k < 16; ++k) Tv[k] = SO(0);
#pragma unroll
          for (int k = 0; k < 18; ++k) Jv[k] = SO(0);
        }
        if (o1) warp_store_aos_ilp<SO, 16>(stage, Tv, o1 + s0 * 16, cnt, lane);
        if (o2) warp_store_aos_ilp<SO, 18>(stage, Jv, o2 + s0 * 18, cnt, lane);
      }
      if constexpr (MODE == 2) {
        // move_via_points: T_IK = [I | p] -> inverse_kinematics(m=1, k=1) (rrr_robot.py:263-270)
        double jq[3] = {0.0, 0.0, 0.0}, jdq[3] = {0.0, 0.0, 0.0};
        bool ok = false;
        if (live) {
          const double Tin[12] = {1.0, 0.0, 0.0, qv[0], 0.0, 1.0, 0.0, qv[1], 0.0, 0.0, 1.0, qv[2]};
          double p0[3];
          arm_ik_p0(P, Tin, p0);
          double tg[6];
          ok = rrr_ik11_trig(P, p0, jq, tg);
          // get_dq_from_dx: dq = J_v(q)^-1 dx (trajectory_generator.py:161-188) with the position
          // Jacobian of the bare arm (identity base / tool, rrr_robot_jacobians.py:19-20) AT the IK's
          // answer -- written out from the sines / cosines the IK already holds:
          //   column 0 = z x O_n, column 1 = y1 x (O_n - O_1), column 2 = y1 x (O_n - O_2)
          const double c0 = tg[0], sn0 = tg[1], c1 = tg[2], s1 = tg[3], c12 = tg[4], s12 = tg[5];
          const double l1 = bare.l[1], l2 = bare.l[2];
          const double rho3 = l1 * c1 + l2 * c12, vz = -l1 * s1 - l2 * s12;
          const double a = -rho3 * sn0, b = c0 * vz, c = -l2 * s12 * c0;
          const double d = rho3 * c0, f = sn0 * vz, g = -l2 * s12 * sn0;
          const double h = 0.0, k = -rho3, l = -l2 * c12;
          const double A0 = f * l - g * k, A1 = c * k - b * l, A2 = b * g - c * f;
          const double B0 = g * h - d * l, B1 = a * l - c * h, B2 = c * d - a * g;
          const double C0 = d * k - f * h, C1 = b * h - a * k, C2 = a * f - b * d;
          const double det = a * A0 + b * B0 + c * C0;
          const double inv = 1.0 / det;  // singular J -> inf/nan, like a failed np.linalg.inv
          jdq[0] = (A0 * dqv[0] + A1 * dqv[1] + A2 * dqv[2]) * inv;
          jdq[1] = (B0 * dqv[0] + B1 * dqv[1] + B2 * dqv[2]) * inv;
          jdq[2] = (C0 * dqv[0] + C1 * dqv[1] + C2 * dqv[2]) * inv;
        }
        if (lane < cnt) {
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (o1) o1[(s0 + lane) * 3 + j] = jq[j];
            if (o2) o2[(s0 + lane) * 3 + j] = jdq[j];
          }
        }
        if (valid && lane < cnt) valid[s0 + lane] = ok ? 1 : 0;
      }
    }
  }
}


// generate_coefficients for a batch of scalar moves: one thread per PAIR of moves with 16-byte
// accesses when the arrays allow it (2 loads + 2 stores of 8 bytes per move left this kernel
// latency-bound at 0.79 of HBM: ncu long_scoreboard 10.7 per issue), one move per thread otherwise.
template <bool VEC2>
__global__ void __launch_bounds__(kBlock)
trap_coefficients_kernel(const double* __restrict__ q0, const double* __restrict__ qf, int64_t n,
                         double vmax, double amax, double freq, double* __restrict__ t1,
                         double* __restrict__ tau) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (VEC2) {
    const int64_t pairs = n >> 1;
    // software pipeline: the next trip's operands are in flight while this trip's square roots and
    // divisions run (ncu: long_scoreboard was the top stall, 6.8 cycles per issue)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double2 an = make_double2(0.0, 0.0), bn = an;
    if (i < pairs) {
      an = __ldcs(reinterpret_cast<const double2*>(q0) + i);
      bn = __ldcs(reinterpret_cast<const double2*>(qf) + i);
    }
    for (; i < pairs; i += stride) {
      const double2 a = an, b = bn;
      const int64_t in = i + stride;
      if (in < pairs) {
        an = __ldcs(reinterpret_cast<const double2*>(q0) + in);
        bn = __ldcs(reinterpret_cast<const double2*>(qf) + in);
      }
      double2 o1, o2;
      trap_coefficients(a.x, b.x, vmax, amax, freq, o1.x, o2.x);
      trap_coefficients(a.y, b.y, vmax, amax, freq, o1.y, o2.y);
      __stcs(reinterpret_cast<double2*>(t1) + i, o1);
      __stcs(reinterpret_cast<double2*>(tau) + i, o2);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
      double a, b;
      trap_coefficients(q0[n - 1], qf[n - 1], vmax, amax, freq, a, b);
      t1[n - 1] = a, tau[n - 1] = b;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      double a, b;
      trap_coefficients(q0[i], qf[i], vmax, amax, freq, a, b);
      t1[i] = a, tau[i] = b;
    }
  }
}

// get_q / get_dq / get_ddq / get_t for explicit per-profile state (q_init, current
// velocity / acceleration limits, t_1, tau): one warp per profile.
__global__ void __launch_bounds__(kBlock)
trap_sample_kernel(const double* __restrict__ q0, const double* __restrict__ v,
                   const double* __restrict__ a, const double* __restrict__ t1,
                   const double* __restrict__ tau, int64_t nprof, int nsteps, double* __restrict__ q,
                   double* __restrict__ dq, double* __restrict__ ddq, double* __restrict__ t,
                   int32_t* __restrict__ len) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t pr = (((int64_t)blockIdx.x * blockDim.x) >> 5) + warp; pr < nprof; pr += warps) {
    const TrapTimes tt = trap_times(t1[pr], tau[pr], nsteps);
    const double q0v = q0[pr], vv = v[pr], av = a[pr];
    if (lane == 0) len[pr] = tt.L;
    for (int i = lane; i < nsteps; i += kWarp) {
      double qo = 0.0, dqo = 0.0, ddqo = 0.0, tout = 0.0;
      if (i < tt.L) {
        int seg;
        double tk;
        trap_sample_time(tt, i, seg, tk, tout);
        trap_sample_joint(tt, seg, tk, q0v, vv, av, qo, dqo, ddqo);
      }
      const int64_t o = pr * nsteps + i;
      if (q) q[o] = qo;
      if (dq) dq[o] = dqo;
      if (ddq) ddq[o] = ddqo;
      if (t) t[o] = tout;
    }
  }
}

// Quintic boundary-value profiles (utils/polynomial_generator.py:16-96): one warp per profile.
// The reference inverts the 6x6 constraint matrix; the closed form in the shifted variable
// tau = t - t0 is the same polynomial: a0 = q0, a1 = v0, a2 = acc0 / 2 and a3..a5 from the end
// conditions; coefficients in powers of t follow by binomial expansion.
__global__ void __launch_bounds__(kBlock)
poly5_kernel(const double* __restrict__ b0, const double* __restrict__ bf, const double* __restrict__ t0a,
             const double* __restrict__ tfa, int64_t nprof, int nsteps, double* __restrict__ coef,
             double* __restrict__ q, double* __restrict__ dq, double* __restrict__ ddq) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t pr = (((int64_t)blockIdx.x * blockDim.x) >> 5) + warp; pr < nprof; pr += warps) {
    const double q0 = b0[pr * 3], v0 = b0[pr * 3 + 1], a0 = b0[pr * 3 + 2];
    const double qf = bf[pr * 3], vf = bf[pr * 3 + 1], af = bf[pr * 3 + 2];
    const double t0 = t0a[pr], tf = tfa[pr];
    const double T = tf - t0, T2 = T * T, T3 = T2 * T, T4 = T3 * T, T5 = T4 * T;
    const double h = qf - q0;
    double a[6];
    a[0] = q0, a[1] = v0, a[2] = 0.5 * a0;
    a[3] = (20.0 * h - (8.0 * vf + 12.0 * v0) * T - (3.0 * a0 - af) * T2) / (2.0 * T3);
    a[4] = (-30.0 * h + (14.0 * vf + 16.0 * v0) * T + (3.0 * a0 - 2.0 * af) * T2) / (2.0 * T4);
    a[5] = (12.0 * h - 6.0 * (vf + v0) * T + (af - a0) * T2) / (2.0 * T5);
    // c_k = sum_{i>=k} a_i C(i,k) (-t0)^(i-k)
    const double binom[6][6] = {{1, 0, 0, 0, 0, 0}, {1, 1, 0, 0, 0, 0},  {1, 2, 1, 0, 0, 0},
                                {1, 3, 3, 1, 0, 0}, {1, 4, 6, 4, 1, 0}, {1, 5, 10, 10, 5, 1}};
    double c[6];
    const double mt = -t0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double acc = 0.0, pw = 1.0;
#pragma unroll
      for (int i = k; i < 6; ++i) {
        acc += a[i] * binom[i][k] * pw;
        pw *= mt;
      }
      c[k] = acc;
    }
    if (coef && lane < 6) {
      double v = c[0];
#pragma unroll
      for (int k = 1; k < 6; ++k) v = (lane == k) ? c[k] : v;
      coef[pr * 6 + lane] = v;
    }
    const double step = (nsteps > 1) ? M::div(M::sub(tf, t0), (double)(nsteps - 1)) : 0.0;
    for (int i = lane; i < nsteps; i += kWarp) {
      const double t = linspace_at(t0, tf, step, nsteps, i);
      const double t2 = t * t, t3 = t2 * t, t4 = t3 * t, t5 = t4 * t;
      const int64_t o = pr * nsteps + i;
      if (q) q[o] = c[0] + c[1] * t + c[2] * t2 + c[3] * t3 + c[4] * t4 + c[5] * t5;
      if (dq) dq[o] = c[1] + 2.0 * c[2] * t + 3.0 * c[3] * t2 + 4.0 * c[4] * t3 + 5.0 * c[5] * t4;
      if (ddq) ddq[o] = 2.0 * c[2] + 6.0 * c[3] * t + 12.0 * c[4] * t2 + 20.0 * c[5] * t3;
    }
  }
}

template <int NJ, int MODE, typename SO = double>
static int launch_traj(const ArmPT<SO>& P, const double* q0, const double* qf, int64_t ntraj, int nsteps,
                       double vmax, double amax, double freq, SO* q, SO* dq, SO* ddq, SO* t, int32_t* len,
                       SO* o1, SO* o2, uint8_t* valid, cudaStream_t st) {
  // 2 trajectories per warp (ms per 2^15 x 1000 samples, 6 joints: 8 per warp 0.85, 2: 0.79, 1: 0.79)
  int grid = stream_grid(traj_kernel<NJ, MODE, SO>, kBlock, 0, ntraj * kWarp, 2);
  traj_kernel<NJ, MODE, SO><<<grid, kBlock, 0, st>>>(P, q0, qf, ntraj, nsteps, vmax, amax, freq, q, dq,
                                                     ddq, t, len, o1, o2, valid);
  return check_launch();
}

}  // namespace rk

extern "C" {

int rk_trap_p2p_f64(const double* q0, const double* qf, int64_t ntraj, int nj, int nsteps,
                    double v_max, double a_max, double control_freq, double* q, double* dq,
                    double* ddq, double* t, int32_t* len, void* stream) {
  if (ntraj < 0 || nsteps < 1 || !len) return RK_ERR_INVALID_ARG;
  if (nj < 1 || nj > RK_MAX_TRAJ_JOINTS) return RK_ERR_UNSUPPORTED;
  if (ntraj == 0) return RK_OK;
  if (!q0 || !qf) return RK_ERR_INVALID_ARG;
  rk::ArmP P = {};
  cudaStream_t st = (cudaStream_t)stream;
#define RK_CASE(N)                                                                            \
  case N:                                                                                     \
    return rk::launch_traj<N, 0, double>(P, q0, qf, ntraj, nsteps, v_max, a_max, control_freq, q, dq, \
                                 ddq, t, len, nullptr, nullptr, nullptr, st);
  switch (nj) {
    RK_CASE(1) RK_CASE(2) RK_CASE(3) RK_CASE(4) RK_CASE(5) RK_CASE(6) RK_CASE(7) RK_CASE(8)
  }
#undef RK_CASE
  return RK_ERR_UNSUPPORTED;
}

int rk_trap_coefficients_f64(const double* q0, const double* qf, int64_t n, double v_max,
                             double a_max, double control_freq, double* t1, double* tau,
                             void* stream) {
  if (n < 0) return RK_ERR_INVALID_ARG;
  if (n == 0) return RK_OK;
  if (!q0 || !qf || !t1 || !tau) return RK_ERR_INVALID_ARG;
  const bool vec = n >= 2 && (((uintptr_t)q0 | (uintptr_t)qf | (uintptr_t)t1 | (uintptr_t)tau) & 15) == 0;
  if (vec) {
    int grid = rk::stream_grid(rk::trap_coefficients_kernel<true>, rk::kBlock, 0, n / 2);
    rk::trap_coefficients_kernel<true><<<grid, rk::kBlock, 0, (cudaStream_t)stream>>>(
        q0, qf, n, v_max, a_max, control_freq, t1, tau);
  } else {
    int grid = rk::stream_grid(rk::trap_coefficients_kernel<false>, rk::kBlock, 0, n);
    rk::trap_coefficients_kernel<false><<<grid, rk::kBlock, 0, (cudaStream_t)stream>>>(
        q0, qf, n, v_max, a_max, control_freq, t1, tau);
  }
  return rk::check_launch();
}

int rk_trap_sample_f64(const double* q_init, const double* v_cur, const double* a_cur,
                       const double* t1, const double* tau, int64_t nprof, int nsteps, double* q,
                       double* dq, double* ddq, double* t, int32_t* len, void* stream) {
  if (nprof < 0 || nsteps < 1) return RK_ERR_INVALID_ARG;
  if (nprof == 0) return RK_OK;
  if (!q_init || !v_cur || !a_cur || !t1 || !tau || !len) return RK_ERR_INVALID_ARG;
  int grid = rk::stream_grid(rk::trap_sample_kernel, rk::kBlock, 0, nprof * rk::kWarp, 2);  // 8: 0.40, 2: 0.35 ms
  rk::trap_sample_kernel<<<grid, rk::kBlock, 0, (cudaStream_t)stream>>>(
      q_init, v_cur, a_cur, t1, tau, nprof, nsteps, q, dq, ddq, t, len);
  return rk::check_launch();
}

int rk_poly5_f64(const double* b0, const double* bf, const double* t0, const double* tf, int64_t nprof,
                 int nsteps, double* coef, double* q, double* dq, double* ddq, void* stream) {
  if (nprof < 0 || nsteps < 1) return RK_ERR_INVALID_ARG;
  if (nprof == 0) return RK_OK;
  if (!b0 || !bf || !t0 || !tf) return RK_ERR_INVALID_ARG;
  int grid = rk::stream_grid(rk::poly5_kernel, rk::kBlock, 0, nprof * rk::kWarp, 2);  // 8: 0.32, 2: 0.27 ms
  rk::poly5_kernel<<<grid, rk::kBlock, 0, (cudaStream_t)stream>>>(b0, bf, t0, tf, nprof, nsteps, coef, q,
                                                                    dq, ddq);
  return rk::check_launch();
}

int rk_rrr_p2p_fk_f64(const RkArm* robot, const double* q0, const double* qf, int64_t ntraj,
                      int nsteps, double v_max, double a_max, double control_freq, double* q,
                      double* dq, double* ddq, double* t, int32_t* len, double* T, double* J,
                      void* stream) {
  if (!robot || ntraj < 0 || nsteps < 1 || !len) return RK_ERR_INVALID_ARG;
  if (ntraj == 0) return RK_OK;
  if (!q0 || !qf) return RK_ERR_INVALID_ARG;
  rk::ArmP P = rk::make_arm_params(robot);
  return rk::launch_traj<3, 1, double>(P, q0, qf, ntraj, nsteps, v_max, a_max, control_freq, q, dq, ddq, t,
                               len, T, J, nullptr, (cudaStream_t)stream);
}

int rk_rrr_p2p_fk_f32(const RkArm* robot, const double* q0, const double* qf, int64_t ntraj, int nsteps,
                      double v_max, double a_max, double control_freq, float* q, float* dq, float* ddq,
                      float* t, int32_t* len, float* T, float* J, void* stream) {
  if (!robot || ntraj < 0 || nsteps < 1 || !len) return RK_ERR_INVALID_ARG;
  if (ntraj == 0) return RK_OK;
  if (!q0 || !qf) return RK_ERR_INVALID_ARG;
  rk::ArmPT<float> P = rk::make_arm_params_t<float>(robot);
  return rk::launch_traj<3, 1, float>(P, q0, qf, ntraj, nsteps, v_max, a_max, control_freq, q, dq, ddq, t,
                                      len, T, J, nullptr, (cudaStream_t)stream);
}

int rk_rrr_lin_ik_f64(const RkArm* robot, const double* p0, const double* pf, int64_t ntraj,
                      int nsteps, double dx_max, double ddx_max, double control_freq, double* x,
                      double* dx, double* t, double* qj, double* dqj, uint8_t* valid, int32_t* len,
                      void* stream) {
  if (!robot || ntraj < 0 || nsteps < 1 || !len) return RK_ERR_INVALID_ARG;
  if (ntraj == 0) return RK_OK;
  if (!p0 || !pf) return RK_ERR_INVALID_ARG;
  rk::ArmP P = rk::make_arm_params(robot);
  return rk::launch_traj<3, 2, double>(P, p0, pf, ntraj, nsteps, dx_max, ddx_max, control_freq, x, dx,
                               nullptr, t, len, qj, dqj, valid, (cudaStream_t)stream);
}

}  // extern "C"
