// context.cuh - handle state, error reporting, workspace arena, launch bookkeeping.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/rgp_psi.h"

namespace rgp {

extern thread_local char g_last_error[512];

int set_error(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));

struct KernelStat {
  const char* name;
  double total_ms;
  int64_t launches;
};

struct PendingEvent {
  const char* name;
  cudaEvent_t start, stop;
};

}  // namespace rgp

struct rgp_psi_ctx {
  int device = 0;
  int sm_count = 148;
  int impl = RGP_PSI_IMPL_AUTO;
  int64_t row_chunk = 0;
  int profile = 0;
  int accumulate = 0;      // internal: Psi2 / dZ / dell / dvar outputs are added to, not overwritten
  int64_t host_chunk = 0;  // rows per pipelined chunk of the *_host entry points (0 = 262144)
  cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;   // *_host pipeline streams
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_cmp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  int bwd_pipe = 2;       // Psi2 backward kernel: 0 row-at-a-time, 1 software-pipelined (psi2_bwdp.cuh), 2 (default) = the faster of the two per pass type
  int small_m = 2;        // small-inducing-set kernels (psi2_small.cuh): 0 never, 1 whenever the shape fits, 2 (default) when they also save work
  int small_ks = 0;       // their stage-2 k split: 0 = default, or 1 / 2 / 4
  int small_warps = 0;    // their CTA size: 0 / 16 = 16 warps, one CTA per SM; 8 = 8 warps, two CTAs per SM, where that fits (M <= 64; the default there)
  int debug_skip = 0;     // timing experiments only, settable in RGP_DEBUG builds; always 0 in production
  int fwd_smem_pad = 0;   // tuning knob (RGP_DEBUG builds): extra dynamic smem for k_psi2_fwd (forces 1 CTA/SM)
  // device workspace arena (grow-only) and a bump pointer valid for one call
  char* ws = nullptr;
  size_t ws_bytes = 0;
  size_t ws_off = 0;
  // device mirror buffers for the *_host entry points
  char* io = nullptr;
  size_t io_bytes = 0;
  size_t io_off = 0;
  // pinned host staging ring of the *_host entry points (used for caller buffers that are pageable)
  char* pin = nullptr;
  size_t pin_bytes = 0;
  int host_threads = 0;    // threads of the pageable <-> pinned copies (0 = min(8, cores / 2))
  int64_t launches = 0;
  std::vector<rgp::KernelStat> stats;
  std::vector<rgp::PendingEvent> pending;
  std::vector<cudaEvent_t> event_pool;
};

namespace rgp {

#define RGP_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess)                                                              \
      return rgp::set_error(RGP_PSI_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, \
                            __LINE__, cudaGetErrorString(e__));                          \
  } while (0)

#define RGP_TRY(expr)          \
  do {                         \
    int s__ = (expr);          \
    if (s__ != 0) return s__;  \
  } while (0)

// Grow-only arena.  reserve() may free + reallocate (synchronising), so it is called
// once per API call, before any kernel that uses pointers from take().
inline int arena_reserve(char** base, size_t* cap, size_t need) {
  if (need <= *cap) return 0;
  if (*base) {
    RGP_CUDA(cudaDeviceSynchronize());
    RGP_CUDA(cudaFree(*base));
    *base = nullptr;
    *cap = 0;
  }
  size_t want = need + (need >> 3) + (1u << 20);
  cudaError_t e = cudaMalloc((void**)base, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    e = cudaMalloc((void**)base, need);
    want = need;
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    *base = nullptr;
    return set_error(RGP_PSI_ERR_NOMEM, "cudaMalloc of %zu workspace bytes failed: %s", need,
                     cudaGetErrorString(e));
  }
  *cap = want;
  return 0;
}

struct Bump {
  char* base;
  size_t cap;
  size_t off = 0;
  Bump(char* b, size_t c) : base(b), cap(c) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    T* p = (T*)(base + off);
    off += bytes;
    return p;
  }
  bool ok() const { return off <= cap; }
};
inline size_t bump_size(size_t count, size_t elem) { return (count * elem + 255) & ~size_t(255); }

// Launch bookkeeping: counts launches and, in profile mode, brackets each launch with
// an event pair recorded on the launch stream.
struct LaunchScope {
  rgp_psi_ctx* ctx;
  cudaStream_t stream;
  const char* name;
  cudaEvent_t start = nullptr, stop = nullptr;
  LaunchScope(rgp_psi_ctx* c, cudaStream_t s, const char* n) : ctx(c), stream(s), name(n) {
    ctx->launches++;
    if (ctx->profile) {
      start = get_event();
      stop = get_event();
      cudaEventRecord(start, stream);
    }
  }
  ~LaunchScope() {
    if (ctx->profile) {
      cudaEventRecord(stop, stream);
      ctx->pending.push_back({name, start, stop});
    }
  }
  cudaEvent_t get_event() {
    if (!ctx->event_pool.empty()) {
      cudaEvent_t e = ctx->event_pool.back();
      ctx->event_pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
};

#define RGP_LAUNCH(ctx, stream, name, kernel, grid, block, smem, ...)                  \
  do {                                                                                 \
    {                                                                                  \
      rgp::LaunchScope scope__(ctx, stream, name);                                     \
      kernel<<<grid, block, smem, stream>>>(__VA_ARGS__);                              \
    }                                                                                  \
    cudaError_t le__ = cudaGetLastError();                                             \
    if (le__ != cudaSuccess)                                                           \
      return rgp::set_error(RGP_PSI_ERR_CUDA, "launch of %s failed: %s", name,         \
                            cudaGetErrorString(le__));                                 \
  } while (0)

}  // namespace rgp
