// fp64_peak.cuh - DFMA-chain microbenchmark: the measured FP64 (CUDA-core) roofline
// denominator.  MEASURED_PEAKS.json carries HBM and bf16 only, so the
// fp64 peak is measured on the same GPU, same clocks, inside the bench run.
#pragma once
#include "context.cuh"

namespace rgp {
namespace peak {

constexpr int kChains = 16;
constexpr int kInner = 512;

__global__ void __launch_bounds__(256) dfma_chain(int outer, double seed, double* __restrict__ out) {
  double a[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) a[i] = seed + 1e-3 * (threadIdx.x + i);
  const double m = 0.999999, c = 1e-7;
  for (int o = 0; o < outer; ++o) {
#pragma unroll 4
    for (int k = 0; k < kInner; ++k) {
#pragma unroll
      for (int i = 0; i < kChains; ++i) a[i] = fma(a[i], m, c);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += a[i];
  if (s == 123.456) out[0] = s;   // never true; keeps the chains alive
}

static int measure(rgp_psi_ctx* h, cudaStream_t st, int reps, double* tflops) {
  RGP_TRY(arena_reserve(&h->ws, &h->ws_bytes, 256));
  const int threads = 256, blocks = h->sm_count * 8, outer = 64;
  cudaEvent_t e0, e1;
  RGP_CUDA(cudaEventCreate(&e0));
  RGP_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  if (reps < 1) reps = 1;
  for (int r = 0; r < reps + 1; ++r) {   // first launch is the warm-up
    RGP_CUDA(cudaEventRecord(e0, st));
    h->launches++;
    dfma_chain<<<blocks, threads, 0, st>>>(outer, 0.5, (double*)h->ws);
    RGP_CUDA(cudaEventRecord(e1, st));
    RGP_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    RGP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * (double)blocks * threads * outer * kInner * kChains;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  RGP_CUDA(cudaGetLastError());
  *tflops = best;
  return 0;
}

}  // namespace peak
}  // namespace rgp
