// TEMPORARY: entry points not implemented yet return RK_ERR_UNSUPPORTED.
#include "rk_common.cuh"
extern "C" {
int rk_rrr_fk_jac_f64(const RkArm*, const double*, int64_t, int64_t, int, double*, double*, void*) { return RK_ERR_UNSUPPORTED; }
int rk_rrr_ik_f64(const RkArm*, const double*, int64_t, int64_t, int, const int8_t*, int, double*, uint8_t*, void*) { return RK_ERR_UNSUPPORTED; }
int rk_rr_fk_jac_f64(const RkArm*, const double*, int64_t, int64_t, int, double*, double*, void*) { return RK_ERR_UNSUPPORTED; }
int rk_rr_ik_f64(const RkArm*, const double*, int64_t, int64_t, int, const int8_t*, int, double*, uint8_t*, void*) { return RK_ERR_UNSUPPORTED; }
int rk_chain_fk_jac_f64(const RkChain*, const double*, int, int64_t, int64_t, int, double*, double*, void*) { return RK_ERR_UNSUPPORTED; }
int rk_trap_p2p_f64(const double*, const double*, int64_t, int, int, double, double, double, double*, double*, double*, double*, int32_t*, void*) { return RK_ERR_UNSUPPORTED; }
int rk_rrr_p2p_fk_f64(const RkArm*, const double*, const double*, int64_t, int, double, double, double, double*, double*, double*, double*, int32_t*, double*, double*, void*) { return RK_ERR_UNSUPPORTED; }
int rk_rrr_lin_ik_f64(const RkArm*, const double*, const double*, int64_t, int, double, double, double, double*, double*, double*, double*, double*, uint8_t*, int32_t*, void*) { return RK_ERR_UNSUPPORTED; }
int rk_calib_out_len(int p) { return p * (p + 1) / 2 + p + 1; }
size_t rk_calib_workspace_bytes(void) { return 0; }
int rk_calib_normal_eq_f64(const RkCalibModel*, const double*, const double*, int64_t, double*, void*, size_t, void*) { return RK_ERR_UNSUPPORTED; }
int rk_calib_eval_f64(const RkCalibModel*, const double*, const double*, int64_t, double*, void*, size_t, void*) { return RK_ERR_UNSUPPORTED; }
int rk_calib_tool_base_eq_f64(const RkCalibModel*, const double*, const double*, int64_t, double*, void*, size_t, void*) { return RK_ERR_UNSUPPORTED; }
}
