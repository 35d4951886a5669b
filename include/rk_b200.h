/*
 * rk_b200.h -- C ABI of librk_b200.so: batched manipulator kinematics on B200 (sm_100a).
 *
 * This is the drop-in boundary for the data-parallel hot path of
 * Sarrasor/RoboticManipulators.  The reference is pure Python and has no FFI of its
 * own; each entry point below cites the reference interface (file:line, relative to
 * the reference repo) whose per-pose work it replaces with one batched launch.
 *
 * Conventions
 *  - plain C types only; every array pointer is a DEVICE pointer owned by the caller
 *    (for example torch tensors' data_ptr()); parameter structs are HOST pointers and
 *    are copied into the launch (kernel parameter space) -- no global state.
 *  - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it.
 *  - every function returns RK_OK (0) or a negative RkStatus; nothing throws.
 *  - output pointers documented "may be NULL" are skipped (no traffic) when NULL.
 *  - `flags` selects the memory layout of the batched arrays:
 *      RK_AOS (0): pose-major, exactly the reference's shapes stacked on a leading
 *                  batch axis: q[n][6], T[n][4][4], J[n][6][6], ...
 *      RK_SOA (1): component-major planes with leading dimension ld >= n:
 *                  q[6][ld], T[16][ld], J[36][ld] (row-major component index), ...
 *                  every warp access is then a contiguous 256-byte segment.
 *  - 4x4 transforms inside parameter structs are 3x4 row-major (the constant last row
 *    0 0 0 1 is implied); batched T outputs carry all 16 entries like the reference.
 */
#ifndef RK_B200_H
#define RK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RK_VERSION 100

typedef enum RkStatus {
  RK_OK = 0,
  RK_ERR_INVALID_ARG = -1, /* NULL/negative/inconsistent argument                 */
  RK_ERR_CUDA = -2,        /* a CUDA call failed; see rk_last_cuda_error()        */
  RK_ERR_UNSUPPORTED = -3, /* e.g. more tokens/parameters than the compiled limit */
  RK_ERR_WORKSPACE = -4    /* workspace too small; see rk_calib_workspace_bytes() */
} RkStatus;

enum { RK_AOS = 0, RK_SOA = 1, RK_JAC_USER_ANGLES = 2 };

int rk_version(void);
const char* rk_strerror(int status);
const char* rk_last_cuda_error(void);

/* ------------------------------------------------------------------------------------
 * Fanuc165F  (reference: robots/fanuc.py, robots/robot.py)
 * ---------------------------------------------------------------------------------- */

/* Per-robot constants: Fanuc165F.__init__ / set_lengths / set_joint_offsets /
 * set_joint_directions / _generate_value_pairs / _calculate_limits_radians
 * (robots/fanuc.py:43-111), Robot.set_transforms and Robot.epsilon
 * (robots/robot.py:17,39-57).  Filled by the host exactly as the reference computes
 * them (d, dq, radian limits are NOT recomputed on the device). */
typedef struct RkFanuc {
  double l[7];          /* link lengths l_0..l_6 (fanuc.py:73)                      */
  double offsets[6];    /* qs_offsets (fanuc.py:79-80)                              */
  double directions[6]; /* qs_directions (fanuc.py:86)                              */
  double lim_lo[6];     /* qs_lim_rad[i][0] (fanuc.py:110-111)                      */
  double lim_hi[6];     /* qs_lim_rad[i][1]                                         */
  double d;             /* sqrt(l4^2 + l5^2)   (fanuc.py:98)                        */
  double dq;            /* atan2(l4, l5)       (fanuc.py:99)                        */
  double eps;           /* Robot.epsilon = 1e-5 (robot.py:17)                       */
  double T_base[12];    /* world -> base, 3x4 row-major (robot.py:49-52)            */
  double T_tool[12];    /* flange -> tool (robot.py:54-57)                          */
  double T_base_inv[12];/* inverse of T_base, used by IK (fanuc.py:225)             */
  double T_tool_inv[12];/* inverse of T_tool                                        */
  int32_t has_base;     /* 0: T_base is the identity (multiplications are skipped)  */
  int32_t has_tool;     /* 0: T_tool is the identity                                */
} RkFanuc;

/* Host-only convenience for callers without the Python layer: fills `out` the way
 * Fanuc165F.__init__ does (robots/fanuc.py:43-111, robots/robot.py:39-57).  lengths[7],
 * offsets[6], directions[6]: NULL = the reference's defaults; T_base / T_tool: row-major 4x4
 * (16 doubles, last row ignored), NULL = identity.  d, dq and the radian limits are computed as
 * the reference computes them; the inverse transforms are those of the affine part. */
int rk_fanuc_init(RkFanuc* out, const double* lengths, const double* offsets,
                  const double* directions, const double* T_base, const double* T_tool);

/* Batched Fanuc165F.forward_kinematics (robots/fanuc.py:151-179) fused with the 6x6
 * geometric Jacobian of JacobianCalculator.calculate_numerically / calculate_scew on
 * the sequence TzRzTzTxRyTxRyTzTxRxRyRxTx (utils/jacobians.py:166-240, driver
 * fanuc_jacobians.py:17-57) and the singularity measure det J
 * (fanuc_jacobians.py:56-57; manipulability = |det J|).
 *   q   in : n joint vectors (USER angles; q' = q*dir + off is applied, fanuc.py:163)
 *   T   out: n 4x4 poses T_base * chain * T_tool (fanuc.py:173); may be NULL
 *   J   out: n 6x6 Jacobians, columns w.r.t. the MODEL angles q' (the convention of
 *            fanuc_jacobians.py) unless RK_JAC_USER_ANGLES is set, in which case
 *            column i is multiplied by directions[i]; may be NULL
 *   det out: n determinants of J; may be NULL
 * Algorithmic traffic (FP64): 48 B in + 128 + 288 + 8 B out = 472 B per pose. */
int rk_fanuc_fk_jac_f64(const RkFanuc* robot, const double* q, int64_t n, int64_t ld,
                        int flags, double* T, double* J, double* det, void* stream);
int rk_fanuc_fk_jac_f32(const RkFanuc* robot, const float* q, int64_t n, int64_t ld,
                        int flags, float* T, float* J, float* det, void* stream);

/* Batched Fanuc165F.inverse_kinematics (robots/fanuc.py:196-302) over a list of
 * branches (m, k, w).
 *   T            in : n target poses (4x4)
 *   branches_mkw in : HOST pointer to nb triples (m,k,w), m,k in {1,-1}, w in {0,1,-1};
 *                     NULL selects the canonical 8 in the order
 *                     for m in (1,-1): for k in (1,-1): for w in (0,1)
 *                     (then nb must be 8).  Invalid values default to 1 like
 *                     fanuc.py:210-220.
 *   q_out        out: AOS [n][nb][6]; SOA [nb][6][ld].  A failed branch is the
 *                     reference's all-zero sentinel (fanuc.py:239, :300).
 *   valid        out: AOS [n][nb]; SOA [nb][ld]; 1 = solution returned, 0 = failure
 *                     (unreachable, fanuc.py:237, or out of joint limits, :292-300).
 * Algorithmic traffic (FP64, nb = 8): 128 B in + 384 + 8 B out = 520 B per pose. */
int rk_fanuc_ik_f64(const RkFanuc* robot, const double* T, int64_t n, int64_t ld, int flags,
                    const int8_t* branches_mkw, int nb, double* q_out, uint8_t* valid,
                    void* stream);
int rk_fanuc_ik_f32(const RkFanuc* robot, const float* T, int64_t n, int64_t ld, int flags,
                    const int8_t* branches_mkw, int nb, float* q_out, uint8_t* valid,
                    void* stream);

/* TrajectoryGenerator.get_dq_from_dx (utils/trajectory_generator.py:161-188) for the Fanuc165F:
 * dq = J(q)^-1 dx with the 6x6 Jacobian of rk_fanuc_fk_jac_* (dx = [v; omega] per pose; columns
 * w.r.t. the model angles unless RK_JAC_USER_ANGLES), solved in registers with partial pivoting.
 * q, dx, dq: AOS [n][6] / SOA [6][ld]; ok [n]: 0 where J is exactly singular; may be NULL. */
int rk_fanuc_dq_from_dx_f64(const RkFanuc* robot, const double* q, const double* dx, int64_t n,
                            int64_t ld, int flags, double* dq, uint8_t* ok, void* stream);
int rk_fanuc_dq_from_dx_f32(const RkFanuc* robot, const float* q, const float* dx, int64_t n,
                            int64_t ld, int flags, float* dq, uint8_t* ok, void* stream);

/* forward_kinematics + Jacobian + det followed by inverse_kinematics over the canonical 8
 * branches of the pose just computed -- the FK -> IK round trip of fanuc_kinematics.py and
 * tests/test_robots.py:33-52 (test_fanuc_ifk) -- as ONE call.  Arguments and results are
 * exactly those of rk_fanuc_fk_jac_* followed by rk_fanuc_ik_*(T, branches = NULL, nb = 8);
 * T, q_out and valid are required, J and det may be NULL.  With RK_SOA and all five
 * outputs this is a single kernel in which the pose never returns from memory:
 * algorithmic traffic (FP64) 48 B in + 128 + 288 + 8 + 384 + 8 B out = 864 B per pose
 * (992 B for the two calls). */
int rk_fanuc_fk_jac_ik8_f64(const RkFanuc* robot, const double* q, int64_t n, int64_t ld,
                            int flags, double* T, double* J, double* det, double* q_out,
                            uint8_t* valid, void* stream);
int rk_fanuc_fk_jac_ik8_f32(const RkFanuc* robot, const float* q, int64_t n, int64_t ld,
                            int flags, float* T, float* J, float* det, float* q_out,
                            uint8_t* valid, void* stream);

/* The same round trip on HOST arrays (the reference's NumPy arrays: fanuc_kinematics.py:26-54 keeps q,
 * T and the IK answers in host memory).  The batch is streamed through the GPU in chunks over three
 * streams (upload / kernels / download) and a ring of device staging slots owned by an RkHostPipe,
 * one cudaMemcpyAsync per array and chunk, so that both directions of the host link stay busy; the
 * link is the roofline of this call (48 B up, up to 816 B down per pose in FP64).
 *   q, T, J, det, q_out, valid: HOST pointers with the layouts of rk_fanuc_fk_jac_ik8_* (RK_AOS:
 *   the reference's shapes; RK_SOA: planes of pitch ld); page-locked memory gives full link speed,
 *   pageable memory works at the driver's staging speed.  A NULL output is neither computed nor
 *   shipped (q_out and valid come together).  The call returns when the host arrays are complete.
 * A pipe belongs to the device that was current at creation and to one host thread at a time. */
typedef struct RkHostPipe RkHostPipe;
int rk_host_pipe_create(RkHostPipe** pipe, int64_t chunk_poses, int nslots /* 2..8 */);
int rk_host_pipe_destroy(RkHostPipe* pipe);
float rk_host_pipe_last_ms(const RkHostPipe* pipe); /* device time of the last call (first upload -> last download) */
int rk_fanuc_fk_jac_ik8_host_f64(RkHostPipe* pipe, const RkFanuc* robot, const double* q, int64_t n,
                                 int64_t ld, int flags, double* T, double* J, double* det,
                                 double* q_out, uint8_t* valid);
int rk_fanuc_fk_jac_ik8_host_f32(RkHostPipe* pipe, const RkFanuc* robot, const float* q, int64_t n,
                                 int64_t ld, int flags, float* T, float* J, float* det, float* q_out,
                                 uint8_t* valid);
/* Inverse kinematics only on HOST arrays (BASELINE config 2; reference: one
 * Fanuc165F.inverse_kinematics call per target and branch, robots/fanuc.py:196-302): targets
 * T [n][4][4] (RK_SOA: 16 planes of pitch ld) up, the 8 canonical solutions q_out [n][8][6] and
 * valid [n][8] down, through the same pipe -- 128 B up and 392 B down per target in FP64. */
int rk_fanuc_ik8_host_f64(RkHostPipe* pipe, const RkFanuc* robot, const double* T, int64_t n, int64_t ld,
                          int flags, double* q_out, uint8_t* valid);
int rk_fanuc_ik8_host_f32(RkHostPipe* pipe, const RkFanuc* robot, const float* T, int64_t n, int64_t ld,
                          int flags, float* q_out, uint8_t* valid);

/* ------------------------------------------------------------------------------------
 * RRRRobot / RRRobot  (reference: robots/rrr_robot.py, robots/rr_robot.py)
 * ---------------------------------------------------------------------------------- */
typedef struct RkArm {
  double l[3];          /* RRRRobot.ls (rrr_robot.py:36) / RRRobot._ls (rr_robot.py:61-65) */
  double lim_lo[3];     /* qs_lim_rad (rrr_robot.py:37-39, rr_robot.py:36)                 */
  double lim_hi[3];
  double eps;           /* Robot.epsilon                                                    */
  double T_base[12], T_tool[12], T_base_inv[12], T_tool_inv[12];
  int32_t has_base, has_tool;
} RkArm;

/* Host-only: RkArm as RRRRobot() (njoints = 3) / RRRobot() (njoints = 2) build it -- link lengths
 * (1, 1, 1) / (0.8, 0.8) unless `lengths[njoints]` is given, +-360 degree limits, epsilon 1e-5; T_base / T_tool
 * row-major 4x4 or NULL = identity. */
int rk_arm_init(RkArm* out, int njoints, const double* lengths, const double* T_base,
                const double* T_tool);

/* RRRRobot.forward_kinematics (rrr_robot.py:96-120), chain RzTzRyTxRyTx, + the 6x3
 * Jacobian of rrr_robot_jacobians.py:49-64.  q[n][3] -> T[n][4][4], J[n][6][3]. */
int rk_rrr_fk_jac_f64(const RkArm* robot, const double* q, int64_t n, int64_t ld, int flags,
                      double* T, double* J, void* stream);
/* RRRRobot.inverse_kinematics (rrr_robot.py:137-201); branches (m,k), NULL = canonical
 * 4: for m in (1,-1): for k in (1,-1).  q_out AOS [n][nb][3], valid [n][nb]. */
int rk_rrr_ik_f64(const RkArm* robot, const double* T, int64_t n, int64_t ld, int flags,
                  const int8_t* branches_mk, int nb, double* q_out, uint8_t* valid,
                  void* stream);
/* RRRobot.forward_kinematics (rr_robot.py:102-126), chain RzTxRzTx, + 6x2 Jacobian. */
int rk_rr_fk_jac_f64(const RkArm* robot, const double* q, int64_t n, int64_t ld, int flags,
                     double* T, double* J, void* stream);
/* RRRobot.inverse_kinematics (rr_robot.py:147-195); branches m, NULL = (1,-1). */
int rk_rr_ik_f64(const RkArm* robot, const double* T, int64_t n, int64_t ld, int flags,
                 const int8_t* branches_m, int nb, double* q_out, uint8_t* valid,
                 void* stream);
/* FP32 variants of the four entry points above (same arguments; the reachability and limit
 * tests of the IK are evaluated in FP32 on the FP32 inputs). */
int rk_rrr_fk_jac_f32(const RkArm* robot, const float* q, int64_t n, int64_t ld, int flags,
                      float* T, float* J, void* stream);
int rk_rrr_ik_f32(const RkArm* robot, const float* T, int64_t n, int64_t ld, int flags,
                  const int8_t* branches_mk, int nb, float* q_out, uint8_t* valid, void* stream);
int rk_rr_fk_jac_f32(const RkArm* robot, const float* q, int64_t n, int64_t ld, int flags,
                     float* T, float* J, void* stream);
int rk_rr_ik_f32(const RkArm* robot, const float* T, int64_t n, int64_t ld, int flags,
                 const int8_t* branches_m, int nb, float* q_out, uint8_t* valid, void* stream);

/* ------------------------------------------------------------------------------------
 * Generic token chains  (reference: utils/robo_math.py Transformation,
 * utils/jacobians.py JacobianCalculator.calculate_numerically)
 * ---------------------------------------------------------------------------------- */
#define RK_MAX_TOKENS 32

/* Opcode = kind*3 + axis, kind: 0 T, 1 R, 2 T-inverse, 3 R-inverse, 4 T-derivative, 5 R-derivative;
 * axis 0 x, 1 y, 2 z -- the 18 tokens of utils/robo_math.py:368-387 (Tx .. Rz, Txi .. Rzi,
 * Txd .. Rzd).  A derivative token is the matrix dM/dv with a ZERO corner (robo_math.py:556-608), so
 * the [3][3] entry of a product that contains one is 0. */
typedef struct RkChain {
  int32_t ntok;
  int32_t op[RK_MAX_TOKENS];
  int32_t active[RK_MAX_TOKENS]; /* joint_indices mask (jacobians.py:23): Jacobian column  */
  int32_t src[RK_MAX_TOKENS];    /* >=0: column of `values` feeding this token; <0: cst[]  */
  double cst[RK_MAX_TOKENS];     /* constant token value when src < 0                      */
  double T_base[12], T_tool[12];
  int32_t has_base, has_tool;
} RkChain;

/* Batched Transformation(sequence, values).transformation (utils/robo_math.py:364-397)
 * and JacobianCalculator.calculate_numerically(values) (utils/jacobians.py:166-184).
 *   values in : AOS [n][nvals]; SOA [nvals][ld]
 *   T      out: T_base * prod(tokens) * T_tool; may be NULL
 *   J      out: 6 x n_active per sample (AOS [n][6][na]; SOA [6*na][ld]); may be NULL */
int rk_chain_fk_jac_f64(const RkChain* chain, const double* values, int nvals, int64_t n,
                        int64_t ld, int flags, double* T, double* J, void* stream);
int rk_chain_fk_jac_f32(const RkChain* chain, const float* values, int nvals, int64_t n,
                        int64_t ld, int flags, float* T, float* J, void* stream);
/* rk_chain_fk_jac_* and rk_chain_dq_from_dx_f64 run straight-line code specialised for the chain's
 * STRUCTURE (compiled once per structure with NVRTC for sm_100a and cached; constants stay run-time
 * parameters) when libnvrtc can be loaded, and an interpreting kernel otherwise (or with
 * RK_CHAIN_JIT=0 in the environment).
 * Returns "ok", or why the last specialisation was not used. */
const char* rk_chain_jit_status(void);
/* The CUDA source of that specialisation (host only, no GPU needed).  want = 1 T | 2 J | 4 two
 * samples per thread | 8 the dq = J^-1 dx solve of rk_chain_dq_from_dx_f64 on rows 0..na-1.  Two
 * samples per thread (SOA planes, two-element vector accesses) is chosen automatically for FP32
 * walks with an even n and ld; RK_CHAIN_VEC2=0/1/2 in the environment = never / FP32 walks / every
 * SOA walk.  Writes at most cap bytes (NUL-terminated) into buf and returns the size needed, or -1
 * for a bad chain. */
int64_t rk_chain_jit_source(const RkChain* chain, int nvals, int flags, int want, int f32, char* buf,
                            int64_t cap);

/* TrajectoryGenerator.get_dq_from_dx (utils/trajectory_generator.py:161-188) for any chain: the
 * reference evaluates a SYMBOLIC inverse of a square block of the Jacobian per sample
 * (rrr_robot_jacobians.py:14-46 builds it as calculate_numeric()[:3, :3].inv()); here the na x na
 * block J[rows, :] of the chain's Jacobian (na = number of active tokens <= 6) is formed in
 * registers and  J[rows, :] dq = dx  is solved in place with partial pivoting.
 *   rows in : HOST pointer to na distinct row indices in 0..5 (0-2 linear, 3-5 angular);
 *             NULL = rows 0 .. na-1
 *   dx   in : AOS [n][na]; SOA [na][ld];   dq out: same layout
 *   ok   out: [n], 0 where a pivot was exactly zero (singular block); may be NULL */
int rk_chain_dq_from_dx_f64(const RkChain* chain, const double* values, int nvals,
                            const int32_t* rows, const double* dx, int64_t n, int64_t ld, int flags,
                            double* dq, uint8_t* ok, void* stream);

/* ------------------------------------------------------------------------------------
 * Trapezoidal trajectories (reference: utils/trapezoidal_generator.py,
 * utils/trajectory_generator.py, robots/rrr_robot.py:203-280)
 * ---------------------------------------------------------------------------------- */
#define RK_MAX_TRAJ_JOINTS 8

/* TrapezoidalGenerator.generate_coefficients (trapezoidal_generator.py:36-79) for n scalar
 * moves: t1[i], tau[i] from (q0[i], qf[i]) under (v_max, a_max, control_freq). */
int rk_trap_coefficients_f64(const double* q0, const double* qf, int64_t n, double v_max,
                             double a_max, double control_freq, double* t1, double* tau,
                             void* stream);

/* TrapezoidalGenerator.get_q / get_dq / get_ddq / get_t (trapezoidal_generator.py:110-227)
 * for nprof profiles with explicit state: q_init, the CURRENT velocity / acceleration
 * limits (set_current_limits, :81-93), t_1 and tau.  Outputs [nprof][nsteps] padded with
 * zeros; len[i] = the reference's sample count (0 when it raises).  Any of q, dq, ddq, t
 * may be NULL. */
int rk_trap_sample_f64(const double* q_init, const double* v_cur, const double* a_cur,
                       const double* t1, const double* tau, int64_t nprof, int nsteps, double* q,
                       double* dq, double* ddq, double* t, int32_t* len, void* stream);

/* PolynomialGenerator.generate_coefficients + polynomial_from_coefs / dpolynomial_from_coefs
 * (utils/polynomial_generator.py:16-96; TrajectoryGenerator.generate_joint_poly_trajectory,
 * utils/trajectory_generator.py:27-69) for nprof quintic profiles.
 *   b0, bf in : [nprof][3] = (position, velocity, acceleration) at t0[i] and tf[i]
 *   coef  out : [nprof][6] coefficients in powers of t; may be NULL
 *   q,dq,ddq out: [nprof][nsteps] samples at np.linspace(t0, tf, nsteps); any may be NULL */
int rk_poly5_f64(const double* b0, const double* bf, const double* t0, const double* tf, int64_t nprof,
                 int nsteps, double* coef, double* q, double* dq, double* ddq, void* stream);

/* TrajectoryGenerator.generate_p2p_trajectory / _generate_equalized_profiles
 * (trajectory_generator.py:71-92,118-159) over ntraj trajectories of nj joints:
 * per-joint generate_coefficients (trapezoidal_generator.py:36-79), equalisation
 * (:144-153) and sampling get_q/get_dq/get_ddq/get_t (:110-227).  The reference returns
 * FEWER than nsteps samples (three np.linspace segments of truncated lengths); outputs
 * are padded with zeros to nsteps and len[i] holds the reference's sample count (0 when
 * the reference raises because the acceleration segment is empty).  Sample counts and
 * sample values are bit-identical to the reference (no FMA contraction).
 *   q0,qf  in : [ntraj][nj]
 *   q,dq,ddq out: [ntraj][nsteps][nj] (the reference's np.array(qs).T per trajectory);
 *                 any may be NULL
 *   t      out: [ntraj][nsteps]; may be NULL;  len out: [ntraj] */
int rk_trap_p2p_f64(const double* q0, const double* qf, int64_t ntraj, int nj, int nsteps,
                    double v_max, double a_max, double control_freq, double* q, double* dq,
                    double* ddq, double* t, int32_t* len, void* stream);

/* The same sampler fused with RRRRobot.forward_kinematics + Jacobian per sample
 * (RRRRobot.move_joints, rrr_robot.py:203-224): additionally writes
 * T [ntraj][nsteps][4][4] and J [ntraj][nsteps][6][3] (either may be NULL).  nj == 3.
 * Algorithmic traffic: 72 + 8 + 128 + 144 = 352 B per sample. */
int rk_rrr_p2p_fk_f64(const RkArm* robot, const double* q0, const double* qf, int64_t ntraj,
                      int nsteps, double v_max, double a_max, double control_freq, double* q,
                      double* dq, double* ddq, double* t, int32_t* len, double* T, double* J,
                      void* stream);

/* FP32 variant: the profile (coefficients, sample COUNTS, times, q / dq / ddq) is evaluated in
 * FP64 exactly as above -- len[] is bit-identical to the FP64 call -- and rounded to FP32 on the
 * way out; the per-sample forward kinematics and Jacobian are computed in FP32.  176 B per sample. */
int rk_rrr_p2p_fk_f32(const RkArm* robot, const double* q0, const double* qf, int64_t ntraj,
                      int nsteps, double v_max, double a_max, double control_freq, float* q,
                      float* dq, float* ddq, float* t, int32_t* len, float* T, float* J,
                      void* stream);

/* TrajectoryGenerator.generate_lin_trajectory (trajectory_generator.py:94-116) followed
 * per sample by RRRRobot.move_via_points' inverse_kinematics(T_IK=[I|p], m=k=1)
 * (rrr_robot.py:263-270) and get_dq_from_dx, dq = J_v(q)^-1 dx
 * (trajectory_generator.py:161-188, rrr_robot_jacobians.py:14-46).
 *   p0,pf in : [ntraj][3];  x,dx out: [ntraj][nsteps][3]; t: [ntraj][nsteps]
 *   qj,dqj out: [ntraj][nsteps][3]; valid: [ntraj][nsteps]; len: [ntraj] */
int rk_rrr_lin_ik_f64(const RkArm* robot, const double* p0, const double* pf, int64_t ntraj,
                      int nsteps, double dx_max, double ddx_max, double control_freq,
                      double* x, double* dx, double* t, double* qj, double* dqj,
                      uint8_t* valid, int32_t* len, void* stream);

/* ------------------------------------------------------------------------------------
 * Geometric calibration (reference: utils/model_calibration.py)
 * ---------------------------------------------------------------------------------- */
#define RK_CALIB_MAX_JOINTS 8
#define RK_CALIB_MAX_PARAMS 19 /* P + 1 residual column <= 20 = 5 blocks of 4 */
#define RK_CALIB_MAX_TOOLS 4

/* One linearisation point of the reduced model (ModelCalibrator.get_reduced_model,
 * model_calibration.py:122-266), compiled by the host once per iteration: the constant
 * token runs between consecutive joint tokens are pre-multiplied at (nominal +
 * parameters) (model_calibration.py:555-583).  For parameter token p living in the run
 * after joint jp: axis[p] / origin[p] are the token's axis and origin expressed in the
 * frame right after joint jp's rotation, so that in the world
 *   u_p = R_jp * axis[p],  O_p = O_jp + R_jp * origin[p]
 * and the position Jacobian column (utils/jacobians.py:203-233, position rows only,
 * model_calibration.py:550-553) is u_p x (O_n - O_p) for a rotation parameter and u_p
 * for a translation parameter. */
typedef struct RkCalibModel {
  int32_t njoints, nparams, ntools;
  int32_t joint_axis[RK_CALIB_MAX_JOINTS];      /* 0 x, 1 y, 2 z                          */
  int32_t joint_kind[RK_CALIB_MAX_JOINTS];      /* 1 revolute, 0 prismatic                */
  double joint_dir[RK_CALIB_MAX_JOINTS];        /* qs_directions (model_calibration.py:557) */
  double joint_off[RK_CALIB_MAX_JOINTS];        /* qs_offsets                              */
  double seg[RK_CALIB_MAX_JOINTS][12];          /* constant 3x4 after joint j              */
  int32_t param_joint[RK_CALIB_MAX_PARAMS];     /* jp                                      */
  int32_t param_kind[RK_CALIB_MAX_PARAMS];      /* 1 rotation, 0 translation               */
  double param_axis[RK_CALIB_MAX_PARAMS][3];
  double param_origin[RK_CALIB_MAX_PARAMS][3];
  double T_base[12];                            /* includes any tokens before joint 0      */
  double tool_p[RK_CALIB_MAX_TOOLS][3];         /* translation column of each T_tool       */
} RkCalibModel;

/* Number of doubles rk_calib_normal_eq_f64 writes: P(P+1)/2 packed upper triangle of
 * H = sum J^T J (row-major: H00 H01 .. H0P-1 H11 ..), then g = sum J^T r (P), then
 * r^T r (1). */
int rk_calib_out_len(int nparams);
size_t rk_calib_workspace_bytes(void);

/* The accumulation loop of ModelCalibrator.estimate_parameters
 * (utils/model_calibration.py:413-422): for every measurement i and tool j,
 * T = T_base * chain(q_i) * T_tool_j (get_pose, :541-548), J = position rows of the
 * parameter Jacobian (get_jacobian, :550-553), r = p_ij - T[:3,3]; H += J^T J,
 * g += J^T r.  Thread-tile accumulation in registers, block partials in `workspace`,
 * last-block deterministic final sum into `out` (ready to be all-reduced).
 *   q   in : [m][njoints] joint angles (AOS) ; pts in: [m][3*ntools] measured points
 * Algorithmic traffic: (njoints + 3*ntools)*8 B in per measurement (120 B for Fanuc). */
int rk_calib_normal_eq_f64(const RkCalibModel* model, const double* q, const double* pts,
                           int64_t m, double* out, void* workspace, size_t ws_bytes,
                           void* stream);

/* ModelCalibrator.evaluate (utils/model_calibration.py:456-539): out[0..3] = sums of
 * dist^2, dx^2, dy^2, dz^2 over all (measurement, tool); out[4..7] = max dist, |dx|,
 * |dy|, |dz|. */
int rk_calib_eval_f64(const RkCalibModel* model, const double* q, const double* pts,
                      int64_t m, double* out, void* workspace, size_t ws_bytes,
                      void* stream);

/* The A^T A / A^T dp accumulation of ModelCalibrator.estimate_tool_base
 * (utils/model_calibration.py:296-332), dim = 6 + 3*ntools.  out = dim*dim (full,
 * row-major) followed by dim entries of A^T dp.  model->T_base and tool_p are ignored
 * (the reference evaluates get_pose(q) with identity base/tool, :303). */
int rk_calib_tool_base_eq_f64(const RkCalibModel* model, const double* q, const double* pts,
                              int64_t m, double* out, void* workspace, size_t ws_bytes,
                              void* stream);

/* ------------------------------------------------------------------------------------
 * Measurement helpers
 * ---------------------------------------------------------------------------------- */
/* Dependent-free DFMA loop on every SM: runs `iters` rounds of 8 independent FMA chains
 * per thread; *total_fma (host) receives the number of FP64 FMA instructions (per
 * thread-lane) executed so the caller can divide by the event time. */
int rk_probe_dfma(int64_t iters, double* sink, int64_t* total_fma, void* stream);
/* Same for the FP64 tensor-core path: 4 independent mma.sync.m8n8k4.f64 chains per warp;
 * *total_fma receives the number of FP64 FMAs (256 per warp-level MMA). */
int rk_probe_dmma(int64_t iters, double* sink, int64_t* total_fma, void* stream);

/* out[i] = the library's internal FP64 atan2(y[i], x[i]) (octant reduction + degree-10
 * polynomial, used by the IK kernels instead of libm); exposed so the accuracy tests can
 * compare it against libm over the whole plane. */
int rk_probe_atan2(const double* y, const double* x, int64_t n, double* out, void* stream);
/* Access-pattern ceiling: each thread reads `nin` SOA planes and writes `nout` planes at its
 * index with 8-byte (vec = 1) or 16-byte (vec = 2) accesses, streaming (cs = 1) or default
 * stores; used by tools/stream_probe.py to find what K1's 6-in / 53-out pattern can reach. */
int rk_probe_stream(const double* in, int nin, double* out, int nout, int64_t n, int64_t ld, int vec,
                    int cs, int blocks_per_sm, void* stream);
/* s[i], c[i] = the library's internal FP64 sincos(x[i]) (Cody-Waite + polynomials). */
int rk_probe_sincos(const double* x, int64_t n, double* s, double* c, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RK_B200_H */
