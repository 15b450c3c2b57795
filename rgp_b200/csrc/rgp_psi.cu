// rgp_psi.cu - C ABI of librgp_psi.so (see include/rgp_psi.h for the contract and the
// reference interfaces each entry point replaces).
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <new>
#include <thread>
#include <vector>

#include "context.cuh"
#include "common.cuh"
#include "ref_kernels.cuh"
#include "fast_path.cuh"
#include "fp64_peak.cuh"
#include "lag_kernels.cuh"
#include "mlp_kernels.cuh"

namespace rgp {

thread_local char g_last_error[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

static int check_common(rgp_psi_ctx* h, int64_t N, int M, int Q, const void* mu, const void* S,
                        const void* Z, const void* ell, double variance) {
  if (!h) return set_error(RGP_PSI_ERR_INVALID, "null handle");
  if (N <= 0 || M <= 0 || Q <= 0)
    return set_error(RGP_PSI_ERR_INVALID, "N, M, Q must be positive (got %lld, %d, %d)",
                     (long long)N, M, Q);
  if (!mu || !S || !Z || !ell) return set_error(RGP_PSI_ERR_INVALID, "null input pointer");
  if (!(variance > 0.0)) return set_error(RGP_PSI_ERR_INVALID, "variance must be positive");
  if ((int64_t)M * M > (int64_t)1 << 30)
    return set_error(RGP_PSI_ERR_INVALID, "M = %d too large", M);
  if (Q > RGP_PSI_MAX_Q)
    return set_error(RGP_PSI_ERR_INVALID, "Q = %d is larger than the supported maximum of %d", Q, RGP_PSI_MAX_Q);
  return 0;
}

static bool use_fast(const rgp_psi_ctx* h, int M, int Q) {
  if (h->impl == RGP_PSI_IMPL_REFERENCE) return false;
  return fast::supported(M, Q);
}

// ------------------------------------------------------------------ reference path
namespace refdrv {

static int64_t pick_chunk(const rgp_psi_ctx* h, int64_t N, int M, int Q) {
  if (h->row_chunk > 0) return std::min<int64_t>(N, h->row_chunk);
  // keep the per-chunk workspace (L1: chunk*M, dinv: chunk*Q) under ~1 GiB
  int64_t per_row = 8ll * (M + Q + 2);
  int64_t c = ((int64_t)1 << 30) / per_row;
  return std::max<int64_t>(1, std::min<int64_t>(N, c));
}

static int psi2_splits(const rgp_psi_ctx* h, int64_t rows, int M) {
  int tiles = ceil_div(M, 16) * ceil_div(M, 16);
  int64_t want = std::max<int64_t>(1, (int64_t)h->sm_count * 16 / tiles);
  return (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(rows, want), 65535));
}

static int forward(rgp_psi_ctx* h, cudaStream_t st, int64_t N, int M, int Q, const double* mu,
                   const double* S, const double* Z, const double* ell, double variance,
                   double* psi0, double* psi1, double* psi2) {
  const int64_t rc = pick_chunk(h, N, M, Q);
  size_t need = bump_size(rc, 8) * 2 + bump_size(rc * Q, 8) + bump_size((size_t)M * M, 8);
  RGP_TRY(arena_reserve(&h->ws, &h->ws_bytes, need));
  Bump b(h->ws, h->ws_bytes);
  double* c1 = b.take<double>(rc);
  double* c2 = b.take<double>(rc);
  double* dinv = b.take<double>(rc * Q);
  double* zz = b.take<double>((size_t)M * M);
  RGP_LAUNCH(h, st, "ref_pair_terms", ref::pair_terms, ceil_div((int64_t)M * M, 256), 256, 0, M, Q,
             Z, ell, (const double*)nullptr, zz, (double*)nullptr);
  if (!h->accumulate) RGP_CUDA(cudaMemsetAsync(psi2, 0, sizeof(double) * M * M, st));
  if (psi0) RGP_LAUNCH(h, st, "ref_fill", ref::fill, ceil_div(N, 256), 256, 0, N, variance, psi0);
  for (int64_t s = 0; s < N; s += rc) {
    int64_t r = std::min(rc, N - s);
    const double* mu_c = mu + s * Q;
    const double* S_c = S + s * Q;
    RGP_LAUNCH(h, st, "ref_row_terms", ref::row_terms, ceil_div(r, 128), 128, 0, r, Q, S_c, ell, c1,
               c2, dinv);
    if (psi1)
      RGP_LAUNCH(h, st, "ref_psi1", ref::psi1, ceil_div(r * M, 256), 256, 0, r, M, Q, mu_c, S_c, Z,
                 ell, c1, variance, (const double*)nullptr, psi1 + s * M);
    dim3 grid(ceil_div(M, 16), ceil_div(M, 16), psi2_splits(h, r, M));
    RGP_LAUNCH(h, st, "ref_psi2", ref::psi2, grid, dim3(16, 16), 0, r, M, Q, mu_c, Z, c2, dinv, zz,
               variance, psi2);
  }
  return 0;
}

static int backward(rgp_psi_ctx* h, cudaStream_t st, int64_t N, int M, int Q, const double* mu,
                    const double* S, const double* Z, const double* ell, double variance,
                    const double* dL0, double dL0c, const double* dL1, const double* dL2,
                    double* dmu, double* dS, double* dZ, double* dell, double* dvar) {
  const int64_t rc = pick_chunk(h, N, M, Q);
  size_t need = bump_size(rc, 8) * 2 + bump_size(rc * Q, 8) + bump_size((size_t)M * M, 8) * 3 +
                bump_size(rc * M, 8);
  RGP_TRY(arena_reserve(&h->ws, &h->ws_bytes, need));
  Bump b(h->ws, h->ws_bytes);
  double* c1 = b.take<double>(rc);
  double* c2 = b.take<double>(rc);
  double* dinv = b.take<double>(rc * Q);
  double* zz = b.take<double>((size_t)M * M);
  double* dLs = b.take<double>((size_t)M * M);
  double* p2 = b.take<double>((size_t)M * M);
  double* L1 = b.take<double>(rc * M);

  RGP_CUDA(cudaMemsetAsync(dmu, 0, sizeof(double) * N * Q, st));
  RGP_CUDA(cudaMemsetAsync(dS, 0, sizeof(double) * N * Q, st));
  if (!h->accumulate) {
    RGP_CUDA(cudaMemsetAsync(dZ, 0, sizeof(double) * M * Q, st));
    RGP_CUDA(cudaMemsetAsync(dell, 0, sizeof(double) * Q, st));
    RGP_CUDA(cudaMemsetAsync(dvar, 0, sizeof(double), st));
  }
  RGP_CUDA(cudaMemsetAsync(p2, 0, sizeof(double) * M * M, st));
  if (dL0) {
    RGP_LAUNCH(h, st, "ref_sum_dL0", ref::sum_to, std::min(1024, ceil_div(N, 256)), 256, 0, N, dL0,
               dvar);
  } else {
    RGP_LAUNCH(h, st, "ref_add", ref::add_scalar, 1, 32, 0, dL0c * (double)N, dvar);
  }
  RGP_LAUNCH(h, st, "ref_pair_terms", ref::pair_terms, ceil_div((int64_t)M * M, 256), 256, 0, M, Q,
             Z, ell, dL2, zz, dLs);
  for (int64_t s = 0; s < N; s += rc) {
    int64_t r = std::min(rc, N - s);
    const double* mu_c = mu + s * Q;
    const double* S_c = S + s * Q;
    RGP_LAUNCH(h, st, "ref_row_terms", ref::row_terms, ceil_div(r, 128), 128, 0, r, Q, S_c, ell, c1,
               c2, dinv);
    if (dL1) {
      RGP_LAUNCH(h, st, "ref_psi1", ref::psi1, ceil_div(r * M, 256), 256, 0, r, M, Q, mu_c, S_c, Z,
                 ell, c1, variance, dL1 + s * M, L1);
      RGP_LAUNCH(h, st, "ref_psi1_bwd_rows", ref::psi1_bwd_rows, ceil_div(r * Q, 128), 128, 0, r, M,
                 Q, mu_c, S_c, Z, ell, L1, variance, dmu + s * Q, dS + s * Q, dell, dvar);
      dim3 gz(ceil_div((int64_t)M * Q, 128), (unsigned)std::max<int64_t>(1, std::min<int64_t>(r / 64, 64)));
      RGP_LAUNCH(h, st, "ref_psi1_bwd_Z", ref::psi1_bwd_Z, gz, 128, 0, r, M, Q, mu_c, S_c, Z, ell, L1,
                 dZ);
    }
    dim3 grid(ceil_div(M, 16), ceil_div(M, 16), psi2_splits(h, r, M));
    RGP_LAUNCH(h, st, "ref_psi2", ref::psi2, grid, dim3(16, 16), 0, r, M, Q, mu_c, Z, c2, dinv, zz,
               variance, p2);
    int rows_grid = (int)std::min<int64_t>(r, (int64_t)h->sm_count * 8);
    size_t smem = sizeof(double) * (1 + 3 * (size_t)Q);
    RGP_LAUNCH(h, st, "ref_psi2_bwd_rows", (ref::psi2_bwd_rows<128>), rows_grid, 128, smem, r, M, Q,
               mu_c, S_c, Z, ell, c2, dinv, zz, dLs, variance, dmu + s * Q, dS + s * Q, dZ, dell,
               dvar);
  }
  RGP_LAUNCH(h, st, "ref_psi2_bwd_tails", ref::psi2_bwd_tails, ceil_div((int64_t)M * Q, 128), 128, 0,
             M, Q, Z, ell, dLs, p2, dZ, dell);
  return 0;
}

}  // namespace refdrv
}  // namespace rgp

using namespace rgp;

extern "C" {

int rgp_psi_abi_version(void) { return RGP_PSI_ABI_VERSION; }
const char* rgp_psi_last_error(void) { return g_last_error; }

int rgp_psi_create(int device, rgp_psi_handle_t* out) {
  if (!out) return set_error(RGP_PSI_ERR_INVALID, "null out pointer");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return set_error(RGP_PSI_ERR_NODEVICE,
                     "no CUDA device available (%s); librgp_psi has no CPU fallback",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= count)
    return set_error(RGP_PSI_ERR_INVALID, "device %d out of range [0,%d)", device, count);
  cudaDeviceProp prop;
  RGP_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return set_error(RGP_PSI_ERR_NODEVICE,
                     "device %d is sm_%d%d; this library carries sm_100a code only", device,
                     prop.major, prop.minor);
  RGP_CUDA(cudaSetDevice(device));
  rgp_psi_ctx* h = new (std::nothrow) rgp_psi_ctx();
  if (!h) return set_error(RGP_PSI_ERR_NOMEM, "host allocation failed");
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  int s = fast::init(h);
  if (s != 0) {
    delete h;
    return s;
  }
  *out = h;
  return 0;
}

int rgp_psi_destroy(rgp_psi_handle_t h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (auto& p : h->pending) {
    cudaEventDestroy(p.start);
    cudaEventDestroy(p.stop);
  }
  for (auto e : h->event_pool) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
    if (h->ev_cmp[i]) cudaEventDestroy(h->ev_cmp[i]);
    if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
  }
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_cmp) cudaStreamDestroy(h->s_cmp);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  if (h->ws) cudaFree(h->ws);
  if (h->io) cudaFree(h->io);
  if (h->pin) cudaFreeHost(h->pin);
  delete h;
  return 0;
}

int rgp_psi_set_option(rgp_psi_handle_t h, const char* key, int64_t value) {
  if (!h || !key) return set_error(RGP_PSI_ERR_INVALID, "null handle or key");
  if (!strcmp(key, "impl")) {
    if (value < 0 || value > 2) return set_error(RGP_PSI_ERR_INVALID, "impl must be 0, 1 or 2");
    h->impl = (int)value;
  } else if (!strcmp(key, "row_chunk")) {
    if (value < 0) return set_error(RGP_PSI_ERR_INVALID, "row_chunk must be >= 0");
    h->row_chunk = value;
  } else if (!strcmp(key, "host_chunk")) {
    if (value < 0) return set_error(RGP_PSI_ERR_INVALID, "host_chunk must be >= 0");
    h->host_chunk = value;
  } else if (!strcmp(key, "host_threads")) {
    if (value < 0 || value > 256) return set_error(RGP_PSI_ERR_INVALID, "host_threads must be in [0, 256]");
    h->host_threads = (int)value;
  } else if (!strcmp(key, "profile")) {
    h->profile = value != 0;
  } else if (!strcmp(key, "bwd_pipe")) {
#ifdef RGP_DEBUG
    if (value < 0 || value > 3) return set_error(RGP_PSI_ERR_INVALID, "bwd_pipe must be 0 ... 3");
#else
    if (value < 0 || value > 2) return set_error(RGP_PSI_ERR_INVALID, "bwd_pipe must be 0, 1 or 2");
#endif
    h->bwd_pipe = (int)value;
  } else if (!strcmp(key, "small_m")) {
    if (value < 0 || value > 2) return set_error(RGP_PSI_ERR_INVALID, "small_m must be 0, 1 or 2");
    h->small_m = (int)value;
  } else if (!strcmp(key, "small_ks")) {
    if (value != 0 && value != 1 && value != 2 && value != 4)
      return set_error(RGP_PSI_ERR_INVALID, "small_ks must be 0, 1, 2 or 4");
    h->small_ks = (int)value;
  } else if (!strcmp(key, "small_warps")) {
    if (value != 0 && value != 8 && value != 16) return set_error(RGP_PSI_ERR_INVALID, "small_warps must be 0, 8 or 16");
    h->small_warps = (int)value;
#ifdef RGP_DEBUG
  // experiment knobs: they make kernels skip work (wrong results) or change occupancy, so the
  // production library does not know them
  } else if (!strcmp(key, "debug_skip")) {
    h->debug_skip = (int)value;
  } else if (!strcmp(key, "fwd_smem_pad")) {
    if (value < 0 || value > 100000) return set_error(RGP_PSI_ERR_INVALID, "fwd_smem_pad must be in [0, 100000]");
    h->fwd_smem_pad = (int)value;
#endif
  } else {
    return set_error(RGP_PSI_ERR_INVALID, "unknown option '%s'", key);
  }
  return 0;
}

int rgp_psi_forward_dev(rgp_psi_handle_t h, void* stream, int64_t N, int M, int Q, const double* mu,
                        const double* S, const double* Z, const double* ell, double variance,
                        double* psi0_out, double* psi1_out, double* psi2_out) {
  RGP_TRY(check_common(h, N, M, Q, mu, S, Z, ell, variance));
  if (!psi2_out) return set_error(RGP_PSI_ERR_INVALID, "psi2_out must not be null");
  RGP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (h->impl == RGP_PSI_IMPL_FAST && !fast::supported(M, Q))
    return set_error(RGP_PSI_ERR_INVALID, "fast path does not support M=%d Q=%d", M, Q);
  if (use_fast(h, M, Q))
    return fast::forward(h, st, N, M, Q, mu, S, Z, ell, variance, psi0_out, psi1_out, psi2_out);
  return refdrv::forward(h, st, N, M, Q, mu, S, Z, ell, variance, psi0_out, psi1_out, psi2_out);
}

int rgp_psi_backward_dev(rgp_psi_handle_t h, void* stream, int64_t N, int M, int Q,
                         const double* mu, const double* S, const double* Z, const double* ell,
                         double variance, const double* dL_dpsi0, double dL_dpsi0_const,
                         const double* dL_dpsi1, const double* dL_dpsi2, double* dmu_out,
                         double* dS_out, double* dZ_out, double* dell_out, double* dvar_out) {
  RGP_TRY(check_common(h, N, M, Q, mu, S, Z, ell, variance));
  if (!dL_dpsi2 || !dmu_out || !dS_out || !dZ_out || !dell_out || !dvar_out)
    return set_error(RGP_PSI_ERR_INVALID, "null dL_dpsi2 or output pointer");
  RGP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (h->impl == RGP_PSI_IMPL_FAST && !fast::supported(M, Q))
    return set_error(RGP_PSI_ERR_INVALID, "fast path does not support M=%d Q=%d", M, Q);
  if (use_fast(h, M, Q))
    return fast::backward(h, st, N, M, Q, mu, S, Z, ell, variance, dL_dpsi0, dL_dpsi0_const,
                          dL_dpsi1, dL_dpsi2, dmu_out, dS_out, dZ_out, dell_out, dvar_out);
  return refdrv::backward(h, st, N, M, Q, mu, S, Z, ell, variance, dL_dpsi0, dL_dpsi0_const,
                          dL_dpsi1, dL_dpsi2, dmu_out, dS_out, dZ_out, dell_out, dvar_out);
}

int rgp_psi_fused_dev(rgp_psi_handle_t h, void* stream, int64_t N, int M, int Q, const double* mu,
                      const double* S, const double* Z, const double* ell, double variance,
                      const double* dL_dpsi0, double dL_dpsi0_const, const double* dL_dpsi1,
                      const double* dL_dpsi2, double* psi1_out, double* psi2_out, double* dmu_out,
                      double* dS_out, double* dZ_out, double* dell_out, double* dvar_out) {
  RGP_TRY(check_common(h, N, M, Q, mu, S, Z, ell, variance));
  if (!dL_dpsi2 || !psi2_out || !dmu_out || !dS_out || !dZ_out || !dell_out || !dvar_out)
    return set_error(RGP_PSI_ERR_INVALID, "null dL_dpsi2 or output pointer");
  RGP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (h->impl == RGP_PSI_IMPL_FAST && !fast::supported(M, Q))
    return set_error(RGP_PSI_ERR_INVALID, "fast path does not support M=%d Q=%d", M, Q);
  if (use_fast(h, M, Q)) {
    if (fast::fused_supported(Q))
      return fast::backward(h, st, N, M, Q, mu, S, Z, ell, variance, dL_dpsi0, dL_dpsi0_const, dL_dpsi1, dL_dpsi2,
                            dmu_out, dS_out, dZ_out, dell_out, dvar_out, psi1_out, psi2_out);
    RGP_TRY(fast::forward(h, st, N, M, Q, mu, S, Z, ell, variance, nullptr, psi1_out, psi2_out));   // 64 < Q <= 128
    return fast::backward(h, st, N, M, Q, mu, S, Z, ell, variance, dL_dpsi0, dL_dpsi0_const, dL_dpsi1, dL_dpsi2,
                          dmu_out, dS_out, dZ_out, dell_out, dvar_out);
  }
  // reference kernels: two passes
  RGP_TRY(refdrv::forward(h, st, N, M, Q, mu, S, Z, ell, variance, nullptr, psi1_out, psi2_out));
  return refdrv::backward(h, st, N, M, Q, mu, S, Z, ell, variance, dL_dpsi0, dL_dpsi0_const, dL_dpsi1, dL_dpsi2,
                          dmu_out, dS_out, dZ_out, dell_out, dvar_out);
}

// ------------------------------------------------------------- host-buffer wrappers
// Rows are streamed in chunks through double-buffered device mirrors on three streams (copy-in,
// compute, copy-out), so host<->device traffic overlaps the kernels and device memory is bounded by
// the chunk size, not by N.  The copies are only asynchronous from page-locked memory, and GPy hands
// over ordinary (pageable) numpy arrays: every caller buffer that is not page-locked is therefore
// bounced through a pinned staging ring owned by the handle - the calling thread copies chunk c+1
// into the ring (several threads, memory-bandwidth bound) while the GPU works on chunk c, and drains
// the results of chunk c-1 out of it.  Page-locked caller buffers are used in place.
} // extern "C"

namespace rgp {
static int host_pipeline_init(rgp_psi_ctx* h) {
  if (h->s_in) return 0;
  RGP_CUDA(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
  RGP_CUDA(cudaStreamCreateWithFlags(&h->s_cmp, cudaStreamNonBlocking));
  RGP_CUDA(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    RGP_CUDA(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
    RGP_CUDA(cudaEventCreateWithFlags(&h->ev_cmp[i], cudaEventDisableTiming));
    RGP_CUDA(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
  }
  return 0;
}
// Host threads this process may use for its copies / digests: the cores divided by the ranks sharing the node
// (torchrun exports LOCAL_WORLD_SIZE), so eight ranks do not start 8 x 32 threads on 32 cores.
static int host_parallelism() {
  unsigned cores = std::max(1u, std::thread::hardware_concurrency());
  const char* lw = getenv("LOCAL_WORLD_SIZE");
  int ranks = lw ? atoi(lw) : 1;
  if (ranks < 1) ranks = 1;
  return (int)std::max(1u, cores / (unsigned)ranks);
}
static bool is_pinned(const void* p) {
  if (!p) return true;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
static int64_t host_chunk_rows(const rgp_psi_ctx* h, int64_t N, bool staged) {
  // staged chunks are half as long: the pinned ring holds two chunks of every pageable array
  int64_t c = h->host_chunk > 0 ? h->host_chunk : (staged ? 131072 : 262144);
  return std::min<int64_t>(N, c);
}
static int pin_reserve(rgp_psi_ctx* h, size_t need) {
  if (need <= h->pin_bytes) return 0;
  if (h->pin) {
    RGP_CUDA(cudaDeviceSynchronize());
    RGP_CUDA(cudaFreeHost(h->pin));
    h->pin = nullptr;
    h->pin_bytes = 0;
  }
  cudaError_t e = cudaHostAlloc((void**)&h->pin, need, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    h->pin = nullptr;
    return set_error(RGP_PSI_ERR_NOMEM, "cudaHostAlloc of %zu staging bytes failed: %s", need, cudaGetErrorString(e));
  }
  h->pin_bytes = need;
  return 0;
}
// memcpy split over a few threads (a single thread moves ~10 GB/s; the host side of a 17 GB Psi1 must
// keep up with the GPU).  Small copies stay on the calling thread.
static void par_memcpy(const rgp_psi_ctx* h, void* dst, const void* src, size_t bytes) {
  if (!bytes) return;
  int nt = h->host_threads > 0 ? h->host_threads : std::min(8, std::max(1, host_parallelism() / 2));
  if (bytes < (size_t)8 << 20 || nt <= 1) {
    memcpy(dst, src, bytes);
    return;
  }
  const size_t per = ((bytes + nt - 1) / nt + 4095) & ~size_t(4095);
  std::vector<std::thread> th;
  for (int i = 1; i < nt; ++i) {
    const size_t off = per * i;
    if (off >= bytes) break;
    th.emplace_back([=] { memcpy((char*)dst + off, (const char*)src + off, std::min(per, bytes - off)); });
  }
  memcpy(dst, src, std::min(per, bytes));
  for (auto& t : th) t.join();
}
struct AccumulateGuard {       // restores the handle's accumulate flag on every exit path
  rgp_psi_ctx* h;
  int saved;
  explicit AccumulateGuard(rgp_psi_ctx* c) : h(c), saved(c->accumulate) {}
  ~AccumulateGuard() { h->accumulate = saved; }
};
// One row-indexed caller array ([N][width] doubles) of a *_host call: where its chunks live on the host
// side of the DMA (the caller's own memory if page-locked, else a slot of the staging ring).
struct HostArray {
  const double* in = nullptr;   // caller input (or null)
  double* out = nullptr;        // caller output (or null)
  int64_t width = 0;
  bool staged = false;
  double* slot[2] = {nullptr, nullptr};
  const double* src(int k, int64_t r0) const { return staged ? slot[k] : in + r0 * width; }
  double* dst(int k, int64_t r0) const { return staged ? slot[k] : out + r0 * width; }
};
}  // namespace rgp

extern "C" {

int rgp_psi_forward_host(rgp_psi_handle_t h, int64_t N, int M, int Q, const double* mu,
                         const double* S, const double* Z, const double* ell, double variance,
                         double* psi0_out, double* psi1_out, double* psi2_out) {
  RGP_TRY(check_common(h, N, M, Q, mu, S, Z, ell, variance));
  if (!psi2_out) return set_error(RGP_PSI_ERR_INVALID, "psi2_out must not be null");
  RGP_CUDA(cudaSetDevice(h->device));
  RGP_TRY(host_pipeline_init(h));
  HostArray a_mu, a_S, a_p1;
  a_mu.in = mu; a_mu.width = Q; a_mu.staged = !is_pinned(mu);
  a_S.in = S; a_S.width = Q; a_S.staged = !is_pinned(S);
  a_p1.out = psi1_out; a_p1.width = M; a_p1.staged = psi1_out && !is_pinned(psi1_out);
  const bool any_staged = a_mu.staged || a_S.staged || a_p1.staged;
  const int64_t R = host_chunk_rows(h, N, any_staged);
  size_t rq = (size_t)R * Q, rm = (size_t)R * M, mq = (size_t)M * Q, mm = (size_t)M * M;
  size_t need = bump_size(mq, 8) + bump_size(Q, 8) + bump_size(mm, 8) +
                2 * (bump_size(rq, 8) * 2 + (psi1_out ? bump_size(rm, 8) : 0));
  RGP_TRY(arena_reserve(&h->io, &h->io_bytes, need));
  RGP_TRY(pin_reserve(h, 2 * ((a_mu.staged ? bump_size(rq, 8) : 0) + (a_S.staged ? bump_size(rq, 8) : 0) +
                              (a_p1.staged ? bump_size(rm, 8) : 0))));
  Bump b(h->io, h->io_bytes), pb(h->pin, h->pin_bytes);
  double* d_Z = b.take<double>(mq);
  double* d_ell = b.take<double>(Q);
  double* d_p2 = b.take<double>(mm);
  double *d_mu[2], *d_S[2], *d_p1[2];
  for (int i = 0; i < 2; ++i) {
    d_mu[i] = b.take<double>(rq);
    d_S[i] = b.take<double>(rq);
    d_p1[i] = psi1_out ? b.take<double>(rm) : nullptr;
    if (a_mu.staged) a_mu.slot[i] = pb.take<double>(rq);
    if (a_S.staged) a_S.slot[i] = pb.take<double>(rq);
    if (a_p1.staged) a_p1.slot[i] = pb.take<double>(rm);
  }
  AccumulateGuard guard(h);
  RGP_CUDA(cudaMemcpyAsync(d_Z, Z, mq * 8, cudaMemcpyHostToDevice, h->s_cmp));
  RGP_CUDA(cudaMemcpyAsync(d_ell, ell, (size_t)Q * 8, cudaMemcpyHostToDevice, h->s_cmp));
  // results of chunk (c, rows) staged in slot k go back to the caller once their D2H copy has landed
  auto drain = [&](int k, int64_t r0, int64_t rows) -> int {
    if (!a_p1.staged) return 0;
    RGP_CUDA(cudaEventSynchronize(h->ev_out[k]));
    par_memcpy(h, psi1_out + r0 * M, a_p1.slot[k], (size_t)rows * M * 8);
    return 0;
  };
  int c = 0;
  int64_t prev_r0 = 0, prev_rows = 0;
  for (int64_t r0 = 0; r0 < N; r0 += R, ++c) {
    const int64_t rows = std::min(R, N - r0);
    const int k = c & 1;
    if (c >= 2) RGP_CUDA(cudaStreamWaitEvent(h->s_in, h->ev_cmp[k], 0));     // device inputs of chunk c-2 consumed
    if (c >= 2 && (a_mu.staged || a_S.staged)) RGP_CUDA(cudaEventSynchronize(h->ev_in[k]));   // ring slot k copied out
    if (a_mu.staged) par_memcpy(h, a_mu.slot[k], mu + r0 * Q, (size_t)rows * Q * 8);
    if (a_S.staged) par_memcpy(h, a_S.slot[k], S + r0 * Q, (size_t)rows * Q * 8);
    RGP_CUDA(cudaMemcpyAsync(d_mu[k], a_mu.src(k, r0), (size_t)rows * Q * 8, cudaMemcpyHostToDevice, h->s_in));
    RGP_CUDA(cudaMemcpyAsync(d_S[k], a_S.src(k, r0), (size_t)rows * Q * 8, cudaMemcpyHostToDevice, h->s_in));
    RGP_CUDA(cudaEventRecord(h->ev_in[k], h->s_in));
    RGP_CUDA(cudaStreamWaitEvent(h->s_cmp, h->ev_in[k], 0));
    if (c >= 2) RGP_CUDA(cudaStreamWaitEvent(h->s_cmp, h->ev_out[k], 0));    // psi1 of chunk c-2 drained
    h->accumulate = c > 0;
    RGP_TRY(rgp_psi_forward_dev(h, h->s_cmp, rows, M, Q, d_mu[k], d_S[k], d_Z, d_ell, variance, nullptr,
                                d_p1[k], d_p2));
    RGP_CUDA(cudaEventRecord(h->ev_cmp[k], h->s_cmp));
    if (psi1_out) {
      RGP_CUDA(cudaStreamWaitEvent(h->s_out, h->ev_cmp[k], 0));
      RGP_CUDA(cudaMemcpyAsync(a_p1.dst(k, r0), d_p1[k], (size_t)rows * M * 8, cudaMemcpyDeviceToHost, h->s_out));
      RGP_CUDA(cudaEventRecord(h->ev_out[k], h->s_out));
    }
    if (c >= 1) RGP_TRY(drain(k ^ 1, prev_r0, prev_rows));   // the GPU is busy with chunk c meanwhile
    prev_r0 = r0;
    prev_rows = rows;
  }
  RGP_CUDA(cudaMemcpyAsync(psi2_out, d_p2, mm * 8, cudaMemcpyDeviceToHost, h->s_cmp));
  if (psi0_out)
    for (int64_t i = 0; i < N; ++i) psi0_out[i] = variance;                 // Psi0[n] = variance
  RGP_TRY(drain((c - 1) & 1, prev_r0, prev_rows));
  RGP_CUDA(cudaStreamSynchronize(h->s_in));
  RGP_CUDA(cudaStreamSynchronize(h->s_cmp));
  RGP_CUDA(cudaStreamSynchronize(h->s_out));
  return 0;
}

int rgp_psi_backward_host(rgp_psi_handle_t h, int64_t N, int M, int Q, const double* mu,
                          const double* S, const double* Z, const double* ell, double variance,
                          const double* dL_dpsi0, double dL_dpsi0_const, const double* dL_dpsi1,
                          const double* dL_dpsi2, double* dmu_out, double* dS_out, double* dZ_out,
                          double* dell_out, double* dvar_out) {
  RGP_TRY(check_common(h, N, M, Q, mu, S, Z, ell, variance));
  if (!dL_dpsi2 || !dmu_out || !dS_out || !dZ_out || !dell_out || !dvar_out)
    return set_error(RGP_PSI_ERR_INVALID, "null dL_dpsi2 or output pointer");
  RGP_CUDA(cudaSetDevice(h->device));
  RGP_TRY(host_pipeline_init(h));
  HostArray a_mu, a_S, a_d0, a_d1, a_gm, a_gs;
  a_mu.in = mu; a_mu.width = Q; a_mu.staged = !is_pinned(mu);
  a_S.in = S; a_S.width = Q; a_S.staged = !is_pinned(S);
  a_d0.in = dL_dpsi0; a_d0.width = 1; a_d0.staged = dL_dpsi0 && !is_pinned(dL_dpsi0);
  a_d1.in = dL_dpsi1; a_d1.width = M; a_d1.staged = dL_dpsi1 && !is_pinned(dL_dpsi1);
  a_gm.out = dmu_out; a_gm.width = Q; a_gm.staged = !is_pinned(dmu_out);
  a_gs.out = dS_out; a_gs.width = Q; a_gs.staged = !is_pinned(dS_out);
  HostArray* ins[4] = {&a_mu, &a_S, &a_d0, &a_d1};
  HostArray* outs[2] = {&a_gm, &a_gs};
  bool in_staged = false, out_staged = false;
  for (auto* a : ins) in_staged |= a->staged;
  for (auto* a : outs) out_staged |= a->staged;
  const int64_t R = host_chunk_rows(h, N, in_staged || out_staged);
  size_t rq = (size_t)R * Q, rm = (size_t)R * M, mq = (size_t)M * Q, mm = (size_t)M * M;
  size_t need = bump_size(mq, 8) * 2 + bump_size(Q, 8) * 2 + bump_size(mm, 8) + bump_size(1, 8) +
                2 * (bump_size(rq, 8) * 4 + (dL_dpsi0 ? bump_size(R, 8) : 0) + (dL_dpsi1 ? bump_size(rm, 8) : 0));
  RGP_TRY(arena_reserve(&h->io, &h->io_bytes, need));
  size_t pneed = 0;
  for (auto* a : ins) if (a->staged) pneed += 2 * bump_size((size_t)R * a->width, 8);
  for (auto* a : outs) if (a->staged) pneed += 2 * bump_size((size_t)R * a->width, 8);
  RGP_TRY(pin_reserve(h, pneed));
  Bump b(h->io, h->io_bytes), pb(h->pin, h->pin_bytes);
  double* d_Z = b.take<double>(mq);
  double* d_dZ = b.take<double>(mq);
  double* d_ell = b.take<double>(Q);
  double* d_dell = b.take<double>(Q);
  double* d_dL2 = b.take<double>(mm);
  double* d_dvar = b.take<double>(1);
  double *d_mu[2], *d_S[2], *d_dmu[2], *d_dS[2], *d_dL0[2], *d_dL1[2];
  for (int i = 0; i < 2; ++i) {
    d_mu[i] = b.take<double>(rq);
    d_S[i] = b.take<double>(rq);
    d_dmu[i] = b.take<double>(rq);
    d_dS[i] = b.take<double>(rq);
    d_dL0[i] = dL_dpsi0 ? b.take<double>(R) : nullptr;
    d_dL1[i] = dL_dpsi1 ? b.take<double>(rm) : nullptr;
    for (auto* a : ins) if (a->staged) a->slot[i] = pb.take<double>((size_t)R * a->width);
    for (auto* a : outs) if (a->staged) a->slot[i] = pb.take<double>((size_t)R * a->width);
  }
  AccumulateGuard guard(h);
  RGP_CUDA(cudaMemcpyAsync(d_Z, Z, mq * 8, cudaMemcpyHostToDevice, h->s_cmp));
  RGP_CUDA(cudaMemcpyAsync(d_ell, ell, (size_t)Q * 8, cudaMemcpyHostToDevice, h->s_cmp));
  RGP_CUDA(cudaMemcpyAsync(d_dL2, dL_dpsi2, mm * 8, cudaMemcpyHostToDevice, h->s_cmp));
  auto drain = [&](int k, int64_t r0, int64_t rows) -> int {
    if (!out_staged) return 0;
    RGP_CUDA(cudaEventSynchronize(h->ev_out[k]));
    for (auto* a : outs)
      if (a->staged) par_memcpy(h, a->out + r0 * a->width, a->slot[k], (size_t)rows * a->width * 8);
    return 0;
  };
  int c = 0;
  int64_t prev_r0 = 0, prev_rows = 0;
  for (int64_t r0 = 0; r0 < N; r0 += R, ++c) {
    const int64_t rows = std::min(R, N - r0);
    const int k = c & 1;
    if (c >= 2) RGP_CUDA(cudaStreamWaitEvent(h->s_in, h->ev_cmp[k], 0));
    if (c >= 2 && in_staged) RGP_CUDA(cudaEventSynchronize(h->ev_in[k]));    // ring slot k copied out
    for (auto* a : ins)
      if (a->staged) par_memcpy(h, a->slot[k], a->in + r0 * a->width, (size_t)rows * a->width * 8);
    RGP_CUDA(cudaMemcpyAsync(d_mu[k], a_mu.src(k, r0), (size_t)rows * Q * 8, cudaMemcpyHostToDevice, h->s_in));
    RGP_CUDA(cudaMemcpyAsync(d_S[k], a_S.src(k, r0), (size_t)rows * Q * 8, cudaMemcpyHostToDevice, h->s_in));
    if (dL_dpsi0)
      RGP_CUDA(cudaMemcpyAsync(d_dL0[k], a_d0.src(k, r0), (size_t)rows * 8, cudaMemcpyHostToDevice, h->s_in));
    if (dL_dpsi1)
      RGP_CUDA(cudaMemcpyAsync(d_dL1[k], a_d1.src(k, r0), (size_t)rows * M * 8, cudaMemcpyHostToDevice, h->s_in));
    RGP_CUDA(cudaEventRecord(h->ev_in[k], h->s_in));
    RGP_CUDA(cudaStreamWaitEvent(h->s_cmp, h->ev_in[k], 0));
    if (c >= 2) RGP_CUDA(cudaStreamWaitEvent(h->s_cmp, h->ev_out[k], 0));    // dmu/dS of chunk c-2 drained
    h->accumulate = c > 0;
    RGP_TRY(rgp_psi_backward_dev(h, h->s_cmp, rows, M, Q, d_mu[k], d_S[k], d_Z, d_ell, variance, d_dL0[k],
                                 dL_dpsi0_const, d_dL1[k], d_dL2, d_dmu[k], d_dS[k], d_dZ, d_dell, d_dvar));
    RGP_CUDA(cudaEventRecord(h->ev_cmp[k], h->s_cmp));
    RGP_CUDA(cudaStreamWaitEvent(h->s_out, h->ev_cmp[k], 0));
    RGP_CUDA(cudaMemcpyAsync(a_gm.dst(k, r0), d_dmu[k], (size_t)rows * Q * 8, cudaMemcpyDeviceToHost, h->s_out));
    RGP_CUDA(cudaMemcpyAsync(a_gs.dst(k, r0), d_dS[k], (size_t)rows * Q * 8, cudaMemcpyDeviceToHost, h->s_out));
    RGP_CUDA(cudaEventRecord(h->ev_out[k], h->s_out));
    if (c >= 1) RGP_TRY(drain(k ^ 1, prev_r0, prev_rows));
    prev_r0 = r0;
    prev_rows = rows;
  }
  RGP_CUDA(cudaMemcpyAsync(dZ_out, d_dZ, mq * 8, cudaMemcpyDeviceToHost, h->s_cmp));
  RGP_CUDA(cudaMemcpyAsync(dell_out, d_dell, (size_t)Q * 8, cudaMemcpyDeviceToHost, h->s_cmp));
  RGP_CUDA(cudaMemcpyAsync(dvar_out, d_dvar, 8, cudaMemcpyDeviceToHost, h->s_cmp));
  RGP_TRY(drain((c - 1) & 1, prev_r0, prev_rows));
  RGP_CUDA(cudaStreamSynchronize(h->s_in));
  RGP_CUDA(cudaStreamSynchronize(h->s_cmp));
  RGP_CUDA(cudaStreamSynchronize(h->s_out));
  return 0;
}

// ------------------------------------------------------------------ lag window
int rgp_lag_gather_dev(rgp_psi_handle_t h, void* stream, int nseq, const int64_t* seq_desc, int64_t N,
                       int Xwin, int Dx, int Uwin, int Du, const double* lat, const double* ctl,
                       double* X_out) {
  if (!h) return set_error(RGP_PSI_ERR_INVALID, "null handle");
  if (nseq <= 0 || N <= 0 || !seq_desc || !X_out || Xwin < 0 || Uwin < 0 || Dx < 0 || Du < 0)
    return set_error(RGP_PSI_ERR_INVALID, "bad lag-window arguments");
  if ((Xwin > 0 && (!lat || Dx <= 0)) || (Uwin > 0 && (!ctl || Du <= 0)) || Xwin * Dx + Uwin * Du <= 0)
    return set_error(RGP_PSI_ERR_INVALID, "window / source mismatch");
  RGP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)std::min<int64_t>(ceil_div(N, 8), (int64_t)h->sm_count * 32);
  RGP_LAUNCH(h, st, "lag_gather", lag::k_gather, blocks, 256, 0, nseq, seq_desc, N, Xwin, Dx, Uwin, Du, lat,
             ctl, X_out);
  return 0;
}

int rgp_lag_scatter_dev(rgp_psi_handle_t h, void* stream, int nseq, const int64_t* seq_desc, int64_t N,
                        int Xwin, int Dx, int Uwin, int Du, const double* dX, int64_t lat_total,
                        double* lat_grad, int64_t ctl_total, double* ctl_grad) {
  if (!h) return set_error(RGP_PSI_ERR_INVALID, "null handle");
  if (nseq <= 0 || N <= 0 || !seq_desc || !dX || Xwin < 0 || Uwin < 0)
    return set_error(RGP_PSI_ERR_INVALID, "bad lag-window arguments");
  RGP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int Q = Xwin * Dx + Uwin * Du;
  if (Xwin > 0 && lat_grad && lat_total > 0)
    RGP_LAUNCH(h, st, "lag_scatter", lag::k_scatter, ceil_div(lat_total * Dx, 256), 256, 0, nseq, seq_desc,
               Xwin, Dx, 0, Q, 2, dX, lat_total, lat_grad);
  if (Uwin > 0 && ctl_grad && ctl_total > 0)
    RGP_LAUNCH(h, st, "lag_scatter", lag::k_scatter, ceil_div(ctl_total * Du, 256), 256, 0, nseq, seq_desc,
               Uwin, Du, Xwin * Dx, Q, 4, dX, ctl_total, ctl_grad);
  return 0;
}

int rgp_latent_terms_dev(rgp_psi_handle_t h, void* stream, int nseq, const int64_t* seq_desc, int Xwin, int D,
                         const double* lat_mean, const double* lat_var, int64_t lat_total,
                         const double* dL_dYmean, const double* dL_dYvar, int dyvar_cols, double* lat_gmean,
                         double* lat_gvar, double* value_out) {
  if (!h) return set_error(RGP_PSI_ERR_INVALID, "null handle");
  if (nseq <= 0 || !seq_desc || Xwin < 0 || D <= 0 || lat_total <= 0 || !lat_mean || !lat_var || !dL_dYmean ||
      !dL_dYvar || !lat_gmean || !lat_gvar || !value_out)
    return set_error(RGP_PSI_ERR_INVALID, "bad latent-terms arguments");
  if (dyvar_cols != 1 && dyvar_cols != D)
    return set_error(RGP_PSI_ERR_INVALID, "dyvar_cols must be 1 or D (got %d, D = %d)", dyvar_cols, D);
  RGP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)std::min<int64_t>(ceil_div(lat_total * D, 256), (int64_t)h->sm_count * 8);
  RGP_TRY(arena_reserve(&h->ws, &h->ws_bytes, bump_size(blocks, sizeof(double))));
  double* partial = (double*)h->ws;
  RGP_LAUNCH(h, st, "latent_terms", lag::k_latent_terms, blocks, 256, 0, nseq, seq_desc, Xwin, D, lat_mean,
             lat_var, dL_dYmean, dL_dYvar, dyvar_cols, lat_total, lat_gmean, lat_gvar, partial);
  RGP_LAUNCH(h, st, "latent_terms_sum", lag::k_sum_partials, 1, 256, 0, blocks, partial, value_out);
  return 0;
}

static int mlp_common(rgp_psi_handle_t h, int nseq, const int64_t* seq_desc, int Xwin, int Dx, int Uwin, int Du,
                      int nlayers, const int* units, mlp::Shape* sh) {
  if (!h) return set_error(RGP_PSI_ERR_INVALID, "null handle");
  if (nseq <= 0 || !seq_desc || Xwin <= 0 || Dx <= 0 || Uwin < 0 || Du < 0 || !units)
    return set_error(RGP_PSI_ERR_INVALID, "bad MLP free-run arguments");
  if (!mlp::make_shape(nlayers, units, sh))
    return set_error(RGP_PSI_ERR_INVALID, "MLP needs 1..%d layers with positive widths", mlp::MAXL);
  if (sh->u[0] != Xwin * Dx + Uwin * Du || sh->u[nlayers] != Dx)
    return set_error(RGP_PSI_ERR_INVALID, "MLP widths [%d ... %d] do not match the window (%d inputs, %d outputs)",
                     sh->u[0], sh->u[nlayers], Xwin * Dx + Uwin * Du, Dx);
  if (Xwin * Dx + Uwin * Du > mlp::THREADS)
    return set_error(RGP_PSI_ERR_INVALID, "MLP input of %d values is wider than the block (%d)", Xwin * Dx + Uwin * Du, mlp::THREADS);
  if (mlp::bwd_smem(*sh, Xwin, Dx) > 200 * 1024)
    return set_error(RGP_PSI_ERR_INVALID, "MLP with %d parameters does not fit in shared memory", sh->nparams);
  return 0;
}

int rgp_mlp_freerun_dev(rgp_psi_handle_t h, void* stream, int nseq, const int64_t* seq_desc, int Xwin, int Dx,
                        int Uwin, int Du, int nlayers, const int* units, const double* params, double* lat_mean,
                        const double* ctl_mean, double* hidden_acts) {
  mlp::Shape sh;
  RGP_TRY(mlp_common(h, nseq, seq_desc, Xwin, Dx, Uwin, Du, nlayers, units, &sh));
  if (!params || !lat_mean || (Uwin > 0 && !ctl_mean) || (sh.nhid > 0 && !hidden_acts))
    return set_error(RGP_PSI_ERR_INVALID, "null MLP free-run pointer");
  RGP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  RGP_CUDA(cudaFuncSetAttribute(mlp::k_freerun, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mlp::fwd_smem(sh)));
  RGP_LAUNCH(h, st, "mlp_freerun", mlp::k_freerun, nseq, mlp::THREADS, mlp::fwd_smem(sh), sh, seq_desc, Xwin, Dx, Uwin,
             Du, params, lat_mean, ctl_mean, hidden_acts, h->debug_skip);
  return 0;
}

int rgp_mlp_freerun_bwd_dev(rgp_psi_handle_t h, void* stream, int nseq, const int64_t* seq_desc, int Xwin, int Dx,
                            int Uwin, int Du, int nlayers, const int* units, const double* params,
                            const double* lat_mean, const double* ctl_mean, const double* hidden_acts,
                            double* lat_gmean, double* ctl_gmean, double* param_grads) {
  mlp::Shape sh;
  RGP_TRY(mlp_common(h, nseq, seq_desc, Xwin, Dx, Uwin, Du, nlayers, units, &sh));
  if (!params || !lat_mean || (Uwin > 0 && !ctl_mean) || (sh.nhid > 0 && !hidden_acts) || !lat_gmean || !param_grads)
    return set_error(RGP_PSI_ERR_INVALID, "null MLP back-propagation pointer");
  RGP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  RGP_CUDA(cudaFuncSetAttribute(mlp::k_freerun_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)mlp::bwd_smem(sh, Xwin, Dx)));
  RGP_LAUNCH(h, st, "mlp_freerun_bwd", mlp::k_freerun_bwd, nseq, mlp::THREADS, mlp::bwd_smem(sh, Xwin, Dx), sh, seq_desc, Xwin,
             Dx, Uwin, Du, params, lat_mean, ctl_mean, hidden_acts, lat_gmean, ctl_gmean, param_grads);
  return 0;
}

// ------------------------------------------------------------------ host-side content digest
// 128-bit order-sensitive digest of a host buffer (two xxh64-style lanes sets with different seeds over
// 8 MiB slices hashed on several threads; slice digests are folded in slice order).  The plugin keys its
// result cache on it: the layer rewrites q(X) IN PLACE every evaluation (autoreg/layers.py:528-550), so
// only the content identifies an input, and at the headline shape the key covers 21 GB per call.
}  // extern "C"
namespace rgp {
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static const uint64_t HP1 = 0x9E3779B185EBCA87ull, HP2 = 0xC2B2AE3D27D4EB4Full, HP3 = 0x165667B19E3779F9ull;
static inline uint64_t hround(uint64_t acc, uint64_t w) { return rotl64(acc + w * HP2, 31) * HP1; }
static inline uint64_t havalanche(uint64_t h) {
  h ^= h >> 33; h *= HP2; h ^= h >> 29; h *= HP3; h ^= h >> 32;
  return h;
}
// Two independent 64-bit digests per slice.  Lanes a: XXH3-style accumulate - one 32x32->64 multiply per
// 8 input bytes plus the neighbouring word added in, scrambled every 1 KiB - strong and memory-bound on a
// few threads.  Lanes b: rotate / xor / add only, with different lane pairing.  Word position matters in
// both (lane index, rotation count), so permuted input gives another digest.
static void digest_slice(const unsigned char* p, size_t n, uint64_t seed, uint64_t out[2]) {
  uint64_t a[4] = {seed + HP1 + HP2, seed + HP2, seed ^ HP3, seed - HP1};
  uint64_t b[4] = {~seed + HP3, seed ^ HP1, seed * HP2 + 1, seed + HP3 * 3};
  const uint64_t k[4] = {0xBE4BA423396CFEB8ull, 0x1CAD21F72C81017Cull, 0xDB979083E96DD4DEull, 0x1F67B3B7A4A44072ull};
  size_t i = 0;
  while (i + 32 <= n) {
    const size_t stop = std::min(n - 31, i + 1024);
    for (; i < stop; i += 32) {
      uint64_t w[4];
      memcpy(w, p + i, 32);
      for (int j = 0; j < 4; ++j) {
        const uint64_t dk = w[j] ^ k[j];
        a[j] += (dk & 0xFFFFFFFFull) * (dk >> 32) + w[j ^ 1];
        b[j] = (rotl64(b[j], 29) ^ w[(j + 1) & 3]) + w[(j + 2) & 3];
      }
    }
    for (int j = 0; j < 4; ++j) {                       // scramble: keep the accumulators from staying linear
      a[j] = (a[j] ^ (a[j] >> 47) ^ k[(j + 1) & 3]) * HP1;
      b[j] = (b[j] ^ (b[j] >> 31)) * HP2;
    }
  }
  uint64_t tail[4] = {0, 0, 0, 0};
  memcpy(tail, p + i, n - i);
  for (int j = 0; j < 4; ++j) {
    a[j] = hround(a[j], tail[j] ^ (uint64_t)(n - i));
    b[j] = hround(b[j], rotl64(tail[(j + 1) & 3], 17) + (uint64_t)n);
  }
  uint64_t ha = rotl64(a[0], 1) + rotl64(a[1], 7) + rotl64(a[2], 12) + rotl64(a[3], 18) + (uint64_t)n;
  uint64_t hb = rotl64(b[0], 3) ^ rotl64(b[1], 11) ^ rotl64(b[2], 27) ^ rotl64(b[3], 41) ^ ((uint64_t)n * HP1);
  out[0] = havalanche(ha);
  out[1] = havalanche(hb);
}
}  // namespace rgp
extern "C" {

int rgp_host_digest(const void* data, int64_t nbytes, int threads, uint64_t out[2]) {
  if (!out || nbytes < 0 || (!data && nbytes > 0)) return set_error(RGP_PSI_ERR_INVALID, "bad digest arguments");
  const size_t SL = (size_t)8 << 20;
  const size_t n = (size_t)nbytes, nsl = n ? (n + SL - 1) / SL : 1;
  std::vector<uint64_t> part(2 * nsl);
  int nt = threads > 0 ? threads : std::min(32, rgp::host_parallelism());
  nt = (int)std::min<size_t>(nt, nsl);
  auto work = [&](int tix) {
    for (size_t sl = tix; sl < nsl; sl += nt) {
      const size_t off = sl * SL;
      digest_slice((const unsigned char*)data + off, std::min(SL, n - off), 0x5DEECE66Dull + sl, &part[2 * sl]);
    }
  };
  if (nt <= 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
  }
  if (nsl == 1) {
    out[0] = part[0];
    out[1] = part[1];
  } else {
    digest_slice((const unsigned char*)part.data(), part.size() * 8, 0x2545F4914F6CDD1Dull ^ (uint64_t)n, out);
  }
  return 0;
}

// ------------------------------------------------------------------ measurement
int64_t rgp_psi_launch_count(rgp_psi_handle_t h) { return h ? h->launches : -1; }

static void drain_pending(rgp_psi_ctx* h) {
  for (auto& p : h->pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.start, p.stop) == cudaSuccess) {
      bool found = false;
      for (auto& s : h->stats)
        if (s.name == p.name || !strcmp(s.name, p.name)) {
          s.total_ms += ms;
          s.launches++;
          found = true;
          break;
        }
      if (!found) h->stats.push_back({p.name, (double)ms, 1});
    } else {
      cudaGetLastError();
    }
    h->event_pool.push_back(p.start);
    h->event_pool.push_back(p.stop);
  }
  h->pending.clear();
}

int rgp_psi_reset_counters(rgp_psi_handle_t h) {
  if (!h) return set_error(RGP_PSI_ERR_INVALID, "null handle");
  RGP_CUDA(cudaSetDevice(h->device));
  RGP_CUDA(cudaDeviceSynchronize());
  drain_pending(h);
  h->stats.clear();
  h->launches = 0;
  return 0;
}

int rgp_psi_kernel_times(rgp_psi_handle_t h, int cap, const char** names, double* total_ms,
                         int64_t* launches) {
  if (!h) return set_error(RGP_PSI_ERR_INVALID, "null handle");
  RGP_CUDA(cudaSetDevice(h->device));
  RGP_CUDA(cudaDeviceSynchronize());
  drain_pending(h);
  int n = (int)h->stats.size();
  for (int i = 0; i < n && i < cap; ++i) {
    if (names) names[i] = h->stats[i].name;
    if (total_ms) total_ms[i] = h->stats[i].total_ms;
    if (launches) launches[i] = h->stats[i].launches;
  }
  return n;
}

int rgp_psi_small_schedule(int M, int Q, int ks, int backward, signed char* out, int out_bytes) {
  static_assert(sizeof(fast::SmallSched) == 258, "layout documented in rgp_psi.h");
  if (!out || out_bytes < (int)sizeof(fast::SmallSched) + 5)
    return set_error(RGP_PSI_ERR_INVALID, "small_schedule: out must hold %d bytes", (int)sizeof(fast::SmallSched) + 5);
  if (M < 1 || Q < 1 || Q > RGP_PSI_MAX_Q || !(ks == 0 || ks == 1 || ks == 2 || ks == 4) || backward < 0 || backward > 2)
    return set_error(RGP_PSI_ERR_INVALID, "small_schedule: bad shape, k split or pass");
  rgp_psi_ctx tmp;
  tmp.small_m = 1;
  tmp.small_ks = ks;
  const fast::Shape s = fast::make_shape(&tmp, 1, M, Q);
  const fast::SmallPlan p = fast::small_plan(&tmp, s);
  if (!p.ok) return set_error(RGP_PSI_ERR_INVALID, "small_schedule: M=%d Q=%d is served by the block kernels", M, Q);
  const fast::SmallVariant& v = backward == 0 ? p.fwd : (backward == 1 ? p.bwd : p.fused);
  memcpy(out, &v.sched, sizeof(fast::SmallSched));
  out[258] = (signed char)v.warps;
  out[259] = (signed char)v.s1;
  out[260] = (signed char)v.nbuf;
  out[261] = (signed char)v.JMAX;
  tmp.small_m = 2;                                  // would the default rule pick the small kernels for this shape?
  out[262] = (signed char)(fast::small_plan(&tmp, s).ok ? 1 : 0);
  return (int)sizeof(fast::SmallSched) + 5;
}

int rgp_psi_fp64_peak(rgp_psi_handle_t h, void* stream, int reps, double* tflops_out) {
  if (!h || !tflops_out) return set_error(RGP_PSI_ERR_INVALID, "null handle or output");
  RGP_CUDA(cudaSetDevice(h->device));
  return peak::measure(h, (cudaStream_t)stream, reps, tflops_out);
}

int64_t rgp_psi_workspace_bytes(rgp_psi_handle_t h) {
  return h ? (int64_t)(h->ws_bytes + h->io_bytes) : -1;
}

}  // extern "C"
