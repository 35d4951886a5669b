/* Plain C against include/rk_b200.h: nothing in the ABI depends on Python or torch.
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/c_abi_demo.c \
 *       -L roboticmanipulators_b200 -lrk_b200 -L /usr/local/cuda/lib64 -lcudart -lm -o c_abi_demo
 *   LD_LIBRARY_PATH=roboticmanipulators_b200 ./c_abi_demo [n]
 *
 * n joint vectors of the stock Fanuc165F go through the fused call on SOA planes (pose, 6x6
 * Jacobian, det J, the 8 canonical IK solutions with validity), then the same batch through the
 * host-array streaming call; the program prints how many poses found their own joint vector among
 * the valid IK branches and checks that both calls agree. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rk_b200.h"

#define CHECK(x)                                                                         \
  do {                                                                                   \
    int rc_ = (x);                                                                       \
    if (rc_ != 0) {                                                                      \
      fprintf(stderr, "%s failed: %d (%s / %s)\n", #x, rc_, rk_strerror(rc_), rk_last_cuda_error()); \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 100000;
  RkFanuc robot;
  CHECK(rk_fanuc_init(&robot, NULL, NULL, NULL, NULL, NULL));

  /* joint vectors inside the limits, component-major planes [6][n] */
  double* hq = (double*)malloc(sizeof(double) * 6 * n);
  unsigned s = 12345u;
  for (int j = 0; j < 6; ++j)
    for (int64_t i = 0; i < n; ++i) {
      s = s * 1664525u + 1013904223u;
      const double u = (double)(s >> 8) / 16777216.0;
      hq[j * n + i] = robot.lim_lo[j] + u * (robot.lim_hi[j] - robot.lim_lo[j]);
    }
  double *q, *T, *J, *det, *qik;
  uint8_t* valid;
  CHECK(cudaMalloc((void**)&q, sizeof(double) * 6 * n));
  CHECK(cudaMalloc((void**)&T, sizeof(double) * 16 * n));
  CHECK(cudaMalloc((void**)&J, sizeof(double) * 36 * n));
  CHECK(cudaMalloc((void**)&det, sizeof(double) * n));
  CHECK(cudaMalloc((void**)&qik, sizeof(double) * 48 * n));
  CHECK(cudaMalloc((void**)&valid, 8 * n));
  CHECK(cudaMemcpy(q, hq, sizeof(double) * 6 * n, cudaMemcpyHostToDevice));
  CHECK(rk_fanuc_fk_jac_ik8_f64(&robot, q, n, n, RK_SOA, T, J, det, qik, valid, NULL));
  CHECK(cudaDeviceSynchronize());
  double* hqik = (double*)malloc(sizeof(double) * 48 * n);
  uint8_t* hvalid = (uint8_t*)malloc(8 * n);
  CHECK(cudaMemcpy(hqik, qik, sizeof(double) * 48 * n, cudaMemcpyDeviceToHost));
  CHECK(cudaMemcpy(hvalid, valid, 8 * n, cudaMemcpyDeviceToHost));

  int64_t found = 0, nvalid = 0;
  for (int64_t i = 0; i < n; ++i)
    for (int b = 0; b < 8; ++b) {
      if (!hvalid[b * n + i]) continue;
      ++nvalid;
      double worst = 0.0;
      for (int j = 0; j < 6; ++j) {
        double d = fabs(hqik[(b * 6 + j) * n + i] - hq[j * n + i]);
        d = fmin(d, fabs(d - 6.283185307179586));
        worst = fmax(worst, d);
      }
      if (worst < 1e-6) {
        ++found;
        break;
      }
    }
  printf("poses %lld  valid solutions per pose %.3f  own joint vector recovered %.3f\n", (long long)n,
         (double)nvalid / (double)n, (double)found / (double)n);

  /* the same batch on host arrays (pageable here; page-locked memory gives link speed) */
  RkHostPipe* pipe;
  CHECK(rk_host_pipe_create(&pipe, 65536, 3));
  double* sqik = (double*)malloc(sizeof(double) * 48 * n);
  uint8_t* svalid = (uint8_t*)malloc(8 * n);
  CHECK(rk_fanuc_fk_jac_ik8_host_f64(pipe, &robot, hq, n, n, RK_SOA, NULL, NULL, NULL, sqik, svalid));
  CHECK(rk_host_pipe_destroy(pipe));
  int64_t mask_diff = 0;
  double worst = 0.0;
  for (int64_t k = 0; k < 8 * n; ++k) mask_diff += svalid[k] != hvalid[k];
  for (int64_t k = 0; k < 48 * n; ++k) worst = fmax(worst, fabs(sqik[k] - hqik[k]));
  printf("streaming call: %lld differing validity flags, max |dq| %.3g\n", (long long)mask_diff, worst);
  return (mask_diff == 0 && worst < 1e-6 && found > n / 2) ? 0 : 2;
}
