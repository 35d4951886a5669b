// Host-buffer entry points: the batched Fanuc165F round trip on arrays that live in HOST memory
// (the reference's own NumPy arrays), streamed through the GPU in chunks.
//
// A pose costs 48 B up and up to 816 B down while the kernels need ~0.14 ns of GPU time for it, so
// the link is the roofline here (PCIe 5 x16: ~55 GB/s per direction, measured by
// tools/probe_host_link.py) and the only job of this file is to keep BOTH directions of the link
// busy all the time: three streams (upload, kernels, download) over a ring of device staging slots,
// one cudaMemcpyAsync (cudaMemcpy2DAsync for component-major host planes) per array and chunk, and
// nothing shipped that the caller did not ask for (a NULL output is neither computed nor copied).
#include <new>

#include "rk_common.cuh"

struct RkHostPipe {
  int device = 0;
  int64_t chunk = 0;
  int nslots = 0;
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  // per slot: inputs landed / kernels done / downloads done
  cudaEvent_t ev_in[8] = {}, ev_k[8] = {}, ev_out[8] = {};
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  unsigned char* slab = nullptr;  // nslots device staging slots
  size_t slot_bytes = 0;
  float last_ms = 0.0f;
};

namespace rk {

// component counts of one pose: q, T, J, det, q_ik (elements) and valid (bytes)
constexpr int kNq = 6, kNT = 16, kNJ = 36, kND = 1, kNQ = 48, kNV = 8;

static size_t slot_bytes_for(int64_t chunk) {
  const size_t elems = (size_t)chunk * (kNq + kNT + kNJ + kND + kNQ);
  return ((elems * sizeof(double) + (size_t)chunk * kNV + 255) / 256) * 256 + 1024;
}

template <typename S>
struct SlotView {
  S *q, *T, *J, *det, *qik;
  uint8_t* valid;
  SlotView(unsigned char* base, int64_t chunk) {
    // every array starts on a 128-byte boundary (chunk is a multiple of 32 poses)
    S* p = reinterpret_cast<S*>(base);
    q = p, p += chunk * kNq;
    T = p, p += chunk * kNT;
    J = p, p += chunk * kNJ;
    det = p, p += chunk * kND;
    qik = p, p += chunk * kNQ;
    valid = reinterpret_cast<uint8_t*>(p);
  }
};

// one array of one chunk between the host (pitch ld_host elements per plane in SOA) and the device
// staging slot (SOA planes at the fixed pitch ld_dev = chunk size, so tensor maps stay aligned for a
// ragged last chunk)
template <typename E>
static cudaError_t copy_chunk(E* dst, const E* src, int ncomp, int64_t cnt, int64_t lo, int64_t ld_host,
                              int64_t ld_dev, bool soa, bool to_device, cudaStream_t st) {
  const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  if (!soa) {
    // pose-major: the chunk is one contiguous run of the host array
    E* d = to_device ? dst : dst + lo * ncomp;
    const E* s = to_device ? src + lo * ncomp : src;
    return cudaMemcpyAsync(d, s, (size_t)cnt * ncomp * sizeof(E), kind, st);
  }
  E* d = to_device ? dst : dst + lo;
  const E* s = to_device ? src + lo : src;
  const size_t dpitch = (to_device ? (size_t)ld_dev : (size_t)ld_host) * sizeof(E);
  const size_t spitch = (to_device ? (size_t)ld_host : (size_t)ld_dev) * sizeof(E);
  return cudaMemcpy2DAsync(d, dpitch, s, spitch, (size_t)cnt * sizeof(E), ncomp, kind, st);
}

// A failing call must not leave copies in flight on the caller's host arrays: the three streams are
// drained before the error is returned.
static void drain(RkHostPipe* h) {
  if (!h) return;
  if (h->s_out) cudaStreamSynchronize(h->s_out);
  if (h->s_k) cudaStreamSynchronize(h->s_k);
  if (h->s_in) cudaStreamSynchronize(h->s_in);
}

#define RK_TRY(expr)                 \
  do {                               \
    cudaError_t e_ = (expr);         \
    if (e_ != cudaSuccess) {         \
      g_last_cuda = e_;              \
      drain(h);                      \
      cudaGetLastError();            \
      return RK_ERR_CUDA;            \
    }                                \
  } while (0)

template <typename S, typename FkFn, typename FkIkFn>
static int run_host(RkHostPipe* h, const RkFanuc* robot, const S* q, int64_t n, int64_t ld, int flags, S* T,
                    S* J, S* det, S* q_out, uint8_t* valid, FkFn fk, FkIkFn fkik) {
  if (!h || !robot || n < 0) return RK_ERR_INVALID_ARG;
  if (n == 0) return RK_OK;
  if (!q) return RK_ERR_INVALID_ARG;
  if ((q_out == nullptr) != (valid == nullptr)) return RK_ERR_INVALID_ARG;  // the IK returns both
  if (!T && !J && !det && !q_out) return RK_ERR_INVALID_ARG;
  const bool soa = (flags & RK_SOA) != 0;
  if (soa && ld < n) return RK_ERR_INVALID_ARG;
  int cur = 0;
  RK_TRY(cudaGetDevice(&cur));
  if (cur != h->device) return RK_ERR_INVALID_ARG;
  const bool ik = q_out != nullptr;
  const int64_t C = h->chunk;
  RK_TRY(cudaEventRecord(h->ev_t0, h->s_in));
  int64_t lo = 0;
  for (int64_t i = 0; lo < n; ++i, lo += C) {
    const int slot = (int)(i % h->nslots);
    const int64_t cnt = (n - lo) < C ? (n - lo) : C;
    SlotView<S> v(h->slab + (size_t)slot * h->slot_bytes, C);
    // upload: the slot's previous kernels must have consumed its q
    if (i >= h->nslots) RK_TRY(cudaStreamWaitEvent(h->s_in, h->ev_k[slot], 0));
    RK_TRY(copy_chunk<S>(v.q, q, kNq, cnt, lo, ld, C, soa, true, h->s_in));
    RK_TRY(cudaEventRecord(h->ev_in[slot], h->s_in));
    // kernels: inputs landed, and the slot's previous results have left
    RK_TRY(cudaStreamWaitEvent(h->s_k, h->ev_in[slot], 0));
    if (i >= h->nslots) RK_TRY(cudaStreamWaitEvent(h->s_k, h->ev_out[slot], 0));
    int rc;
    if (ik)
      rc = fkik(robot, v.q, cnt, C, flags, v.T, J ? v.J : nullptr, det ? v.det : nullptr, v.qik, v.valid,
                (void*)h->s_k);
    else
      rc = fk(robot, v.q, cnt, C, flags, T ? v.T : nullptr, J ? v.J : nullptr, det ? v.det : nullptr,
              (void*)h->s_k);
    if (rc != RK_OK) {
      drain(h);
      return rc;
    }
    RK_TRY(cudaEventRecord(h->ev_k[slot], h->s_k));
    // download what was asked for
    RK_TRY(cudaStreamWaitEvent(h->s_out, h->ev_k[slot], 0));
    if (T) RK_TRY(copy_chunk<S>(T, v.T, kNT, cnt, lo, ld, C, soa, false, h->s_out));
    if (J) RK_TRY(copy_chunk<S>(J, v.J, kNJ, cnt, lo, ld, C, soa, false, h->s_out));
    if (det) RK_TRY(copy_chunk<S>(det, v.det, kND, cnt, lo, ld, C, soa, false, h->s_out));
    if (ik) {
      RK_TRY(copy_chunk<S>(q_out, v.qik, kNQ, cnt, lo, ld, C, soa, false, h->s_out));
      RK_TRY(copy_chunk<uint8_t>(valid, v.valid, kNV, cnt, lo, ld, C, soa, false, h->s_out));
    }
    RK_TRY(cudaEventRecord(h->ev_out[slot], h->s_out));
  }
  RK_TRY(cudaEventRecord(h->ev_t1, h->s_out));
  RK_TRY(cudaStreamSynchronize(h->s_out));  // the caller reads the host arrays next
  RK_TRY(cudaStreamSynchronize(h->s_k));
  RK_TRY(cudaStreamSynchronize(h->s_in));
  cudaEventElapsedTime(&h->last_ms, h->ev_t0, h->ev_t1);
  return RK_OK;
}

// Inverse kinematics only (BASELINE config 2: targets in, the 8 canonical solutions out): 128 B up,
// 392 B down per pose in FP64 -- the same three-stream ring, the slot's T / q_ik / valid areas.
template <typename S, typename IkFn>
static int run_host_ik(RkHostPipe* h, const RkFanuc* robot, const S* T, int64_t n, int64_t ld, int flags,
                       S* q_out, uint8_t* valid, IkFn ikfn) {
  if (!h || !robot || n < 0) return RK_ERR_INVALID_ARG;
  if (n == 0) return RK_OK;
  if (!T || !q_out || !valid) return RK_ERR_INVALID_ARG;
  const bool soa = (flags & RK_SOA) != 0;
  if (soa && ld < n) return RK_ERR_INVALID_ARG;
  int cur = 0;
  RK_TRY(cudaGetDevice(&cur));
  if (cur != h->device) return RK_ERR_INVALID_ARG;
  const int64_t C = h->chunk;
  RK_TRY(cudaEventRecord(h->ev_t0, h->s_in));
  int64_t lo = 0;
  for (int64_t i = 0; lo < n; ++i, lo += C) {
    const int slot = (int)(i % h->nslots);
    const int64_t cnt = (n - lo) < C ? (n - lo) : C;
    SlotView<S> v(h->slab + (size_t)slot * h->slot_bytes, C);
    if (i >= h->nslots) RK_TRY(cudaStreamWaitEvent(h->s_in, h->ev_k[slot], 0));
    RK_TRY(copy_chunk<S>(v.T, T, kNT, cnt, lo, ld, C, soa, true, h->s_in));
    RK_TRY(cudaEventRecord(h->ev_in[slot], h->s_in));
    RK_TRY(cudaStreamWaitEvent(h->s_k, h->ev_in[slot], 0));
    if (i >= h->nslots) RK_TRY(cudaStreamWaitEvent(h->s_k, h->ev_out[slot], 0));
    const int rc = ikfn(robot, v.T, cnt, C, flags, nullptr, 8, v.qik, v.valid, (void*)h->s_k);
    if (rc != RK_OK) {
      drain(h);
      return rc;
    }
    RK_TRY(cudaEventRecord(h->ev_k[slot], h->s_k));
    RK_TRY(cudaStreamWaitEvent(h->s_out, h->ev_k[slot], 0));
    RK_TRY(copy_chunk<S>(q_out, v.qik, kNQ, cnt, lo, ld, C, soa, false, h->s_out));
    RK_TRY(copy_chunk<uint8_t>(valid, v.valid, kNV, cnt, lo, ld, C, soa, false, h->s_out));
    RK_TRY(cudaEventRecord(h->ev_out[slot], h->s_out));
  }
  RK_TRY(cudaEventRecord(h->ev_t1, h->s_out));
  RK_TRY(cudaStreamSynchronize(h->s_out));
  RK_TRY(cudaStreamSynchronize(h->s_k));
  RK_TRY(cudaStreamSynchronize(h->s_in));
  cudaEventElapsedTime(&h->last_ms, h->ev_t0, h->ev_t1);
  return RK_OK;
}

}  // namespace rk

extern "C" {

int rk_host_pipe_create(RkHostPipe** pipe, int64_t chunk_poses, int nslots) {
  using namespace rk;
  if (!pipe || chunk_poses < 32 || nslots < 2 || nslots > 8) return RK_ERR_INVALID_ARG;
  RkHostPipe* h = new (std::nothrow) RkHostPipe();
  if (!h) return RK_ERR_INVALID_ARG;
  h->chunk = (chunk_poses + 31) / 32 * 32;
  h->nslots = nslots;
  h->slot_bytes = slot_bytes_for(h->chunk);
  bool ok = cudaGetDevice(&h->device) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&h->s_k, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking) == cudaSuccess;
  for (int s = 0; ok && s < nslots; ++s) {
    ok = ok && cudaEventCreateWithFlags(&h->ev_in[s], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_k[s], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_out[s], cudaEventDisableTiming) == cudaSuccess;
  }
  ok = ok && cudaEventCreate(&h->ev_t0) == cudaSuccess && cudaEventCreate(&h->ev_t1) == cudaSuccess;
  ok = ok && cudaMalloc(&h->slab, h->slot_bytes * nslots) == cudaSuccess;
  if (!ok) {
    g_last_cuda = cudaGetLastError();
    rk_host_pipe_destroy(h);
    return RK_ERR_CUDA;
  }
  *pipe = h;
  return RK_OK;
}

int rk_host_pipe_destroy(RkHostPipe* h) {
  if (!h) return RK_OK;
  if (h->slab) cudaFree(h->slab);
  for (int s = 0; s < 8; ++s) {
    if (h->ev_in[s]) cudaEventDestroy(h->ev_in[s]);
    if (h->ev_k[s]) cudaEventDestroy(h->ev_k[s]);
    if (h->ev_out[s]) cudaEventDestroy(h->ev_out[s]);
  }
  if (h->ev_t0) cudaEventDestroy(h->ev_t0);
  if (h->ev_t1) cudaEventDestroy(h->ev_t1);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_k) cudaStreamDestroy(h->s_k);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  delete h;
  return RK_OK;
}

float rk_host_pipe_last_ms(const RkHostPipe* h) { return h ? h->last_ms : -1.0f; }

int rk_fanuc_fk_jac_ik8_host_f64(RkHostPipe* pipe, const RkFanuc* robot, const double* q, int64_t n, int64_t ld,
                                 int flags, double* T, double* J, double* det, double* q_out, uint8_t* valid) {
  return rk::run_host<double>(pipe, robot, q, n, ld, flags, T, J, det, q_out, valid, rk_fanuc_fk_jac_f64,
                              rk_fanuc_fk_jac_ik8_f64);
}
int rk_fanuc_fk_jac_ik8_host_f32(RkHostPipe* pipe, const RkFanuc* robot, const float* q, int64_t n, int64_t ld,
                                 int flags, float* T, float* J, float* det, float* q_out, uint8_t* valid) {
  return rk::run_host<float>(pipe, robot, q, n, ld, flags, T, J, det, q_out, valid, rk_fanuc_fk_jac_f32,
                             rk_fanuc_fk_jac_ik8_f32);
}

int rk_fanuc_ik8_host_f64(RkHostPipe* pipe, const RkFanuc* robot, const double* T, int64_t n, int64_t ld, int flags,
                          double* q_out, uint8_t* valid) {
  return rk::run_host_ik<double>(pipe, robot, T, n, ld, flags, q_out, valid, rk_fanuc_ik_f64);
}
int rk_fanuc_ik8_host_f32(RkHostPipe* pipe, const RkFanuc* robot, const float* T, int64_t n, int64_t ld, int flags,
                          float* q_out, uint8_t* valid) {
  return rk::run_host_ik<float>(pipe, robot, T, n, ld, flags, q_out, valid, rk_fanuc_ik_f32);
}

}  // extern "C"
