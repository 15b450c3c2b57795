// microbench.cu - standalone hardware probes that shape the kernel design (run on the
// GPU box; results are summarised in profiles/).  Not part of librgp_psi.so.
//   1. DFMA-chain peak (the fp64 roofline denominator) at several occupancies
//   2. DMMA (mma.sync m8n8k4 f64) peak, alone and interleaved with DFMA
//   3. LDS.128 cost versus number of distinct addresses per warp (broadcast merging)
//   4. exp throughput: library exp() versus rgp::exp_neg()
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e = (x);                                                          \
    if (e != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                    \
    }                                                                             \
  } while (0)

template <int CH>
__global__ void __launch_bounds__(256) k_dfma(int outer, double* out) {
  double a[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) a[i] = 0.5 + 1e-3 * (threadIdx.x + i);
  const double m = 0.999999, c = 1e-7;
  for (int o = 0; o < outer; ++o)
#pragma unroll 4
    for (int k = 0; k < 256; ++k)
#pragma unroll
      for (int i = 0; i < CH; ++i) a[i] = fma(a[i], m, c);
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += a[i];
  if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// MODE 0: DMMA only; 1: DMMA + DFMA interleaved (per 1 DMMA, NF DFMAs)
template <int MODE, int NF>
__global__ void __launch_bounds__(256) k_dmma(int outer, double* out) {
  double c[8][2];
  double f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    c[i][0] = 0.0;
    c[i][1] = 0.0;
    f[i] = 0.5 + i;
  }
  double a = 1e-3 * threadIdx.x, b = 1.0 + 1e-4 * threadIdx.x;
  const double m = 0.999999, cc = 1e-7;
  for (int o = 0; o < outer; ++o) {
#pragma unroll 2
    for (int k = 0; k < 64; ++k) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dmma884(c[i][0], c[i][1], a, b);
        if (MODE == 1) {
#pragma unroll
          for (int j = 0; j < NF; ++j) f[(i + j) & 7] = fma(f[(i + j) & 7], m, cc);
        }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
  if (s == 123.456) out[0] = s;
}

// LDS.128 probe: every lane loads double2 at smem[(pattern(lane)) * stride]; NLD loads per
// iteration into independent registers, summed afterwards.
__global__ void __launch_bounds__(1024) k_lds(int iters, int distinct, int stride_d2, double* out,
                                              long long* cycles) {
  extern __shared__ double2 sm2[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm2[i] = make_double2(i, -i);
  __syncthreads();
  int lane = threadIdx.x & 31;
  int idx;
  if (distinct >= 32) idx = lane;                 // all distinct, contiguous
  else if (distinct == 8) idx = lane & 7;         // 8 distinct, each quarter-warp sees all 8
  else if (distinct == 4) idx = lane >> 3;        // one address per quarter-warp
  else if (distinct == 2) idx = lane >> 4;
  else idx = 0;                                   // full broadcast
  const double2* p = sm2 + idx * stride_d2;
  double2 acc = make_double2(0, 0);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      double2 v;
      unsigned addr = (unsigned)__cvta_generic_to_shared(p + j * 64);
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
      acc.x += v.x;
      acc.y += v.y;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc.x == 123.456) out[0] = acc.y;
}

template <int WHICH>
__global__ void __launch_bounds__(256) k_exp(int outer, double* out) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -1e-3 * (threadIdx.x + 1) - i;
  double s = 0;
  for (int o = 0; o < outer; ++o) {
#pragma unroll 4
    for (int k = 0; k < 64; ++k) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double e = WHICH == 0 ? exp(x[i]) : rgp::exp_neg(x[i]);
        s += e;
        x[i] = x[i] * 0.999 - 1e-4;
      }
    }
  }
  if (s == 123.456) out[0] = s;
}

// DMMA throughput versus independent accumulators in flight: NACC accumulators per warp,
// blockDim/32 warps per CTA, one CTA per SM (so warps per scheduler = blockDim/128).
template <int NACC>
__global__ void k_dmma_inflight(int iters, double* out) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1e-3 * threadIdx.x, b = 1.0 + 1e-4 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

template <int NACC>
static void probe_inflight(int sms, double* out) {
  for (int warps_per_sched : {1, 2, 4, 8}) {
    int threads = 128 * warps_per_sched;
    int iters = 4096 / NACC;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k_dmma_inflight<NACC><<<sms, threads>>>(iters, out);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_dmma_inflight<NACC><<<sms, threads>>>(iters, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double dmma_per_sched = (double)iters * 8 * NACC * warps_per_sched;
    double cycles = ms * 1e-3 * 1.965e9;
    printf("{\"probe\": \"dmma_inflight\", \"acc_per_warp\": %d, \"warps_per_scheduler\": %d, \"in_flight\": %d, "
           "\"cycles_per_dmma_per_scheduler\": %.2f, \"frac_of_peak\": %.3f}\n",
           NACC, warps_per_sched, NACC * warps_per_sched, cycles / dmma_per_sched, 16.0 / (cycles / dmma_per_sched));
  }
}

// DMMA throughput while fragments stream from shared memory: per iteration NA + NB LDS.64
// fragment loads (conflict-free [64][68] tile pattern) feed NA x NB DMMAs.
template <int NA, int NB>
__global__ void k_dmma_lds(int iters, double* out) {
  extern __shared__ double smd[];
  for (int i = threadIdx.x; i < 2 * 64 * 68; i += blockDim.x) smd[i] = 1e-3 * (i % 97);
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const double* pa = smd + ((wid * 8) % 32 + g) * 68 + t;
  const double* pb = smd + 64 * 68 + ((wid * 8) % 32 + g) * 68 + t;
  double c[NA][NB][2];
#pragma unroll
  for (int i = 0; i < NA; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 4
    for (int k0 = 0; k0 < 64; k0 += 4) {
      double a[NA], b[NB];
#pragma unroll
      for (int i = 0; i < NA; ++i) a[i] = pa[(i * 8 % 32) * 68 + k0];
#pragma unroll
      for (int j = 0; j < NB; ++j) b[j] = pb[(j * 8 % 32) * 68 + k0];
#pragma unroll
      for (int i = 0; i < NA; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NA; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) s += c[i][j][0] + c[i][j][1];
  if (s == 123.456) out[0] = s;
}

template <int NA, int NB>
static void probe_dmma_lds(int sms, double* out) {
  CK(cudaFuncSetAttribute(k_dmma_lds<NA, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 64 * 68 * 8));
  for (int warps : {8, 16}) {
    int iters = 4000 / (NA * NB);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k_dmma_lds<NA, NB><<<sms, warps * 32, 2 * 64 * 68 * 8>>>(iters, out);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_dmma_lds<NA, NB><<<sms, warps * 32, 2 * 64 * 68 * 8>>>(iters, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double dmma_per_sched = (double)iters * 16 * NA * NB * warps / 4.0;
    double cycles = ms * 1e-3 * 1.965e9;
    printf("{\"probe\": \"dmma_lds\", \"tile\": \"%dx%d\", \"lds_per_dmma\": %.3f, \"warps_per_sm\": %d, "
           "\"frac_of_dmma_peak\": %.3f}\n", NA, NB, (double)(NA + NB) / (NA * NB), warps,
           16.0 / (cycles / dmma_per_sched));
  }
}

// Row-loop mimic of the backward kernel: per "row" LOOPS short MMA loops of KSTEPS k-steps on a
// 2x2 tile, accumulators reset per loop and folded into persistent registers with a few DFMAs;
// optional __syncthreads per row.  Isolates the cost of short loops / phase boundaries.
template <int LOOPS, int KSTEPS, bool SYNC>
__global__ void k_rowloop(int rows, double* out) {
  extern __shared__ double smd[];
  for (int i = threadIdx.x; i < 3 * 64 * 68; i += blockDim.x) smd[i] = 1e-3 * (i % 97);
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int wr = wid >> 2, wc = wid & 3;
  double keep[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) keep[i][j][0] = keep[i][j][1] = 0.0;
  for (int n = 0; n < rows; ++n) {
#pragma unroll
    for (int l = 0; l < LOOPS; ++l) {
      const double* pa = smd + (l % 3) * 64 * 68 + (16 * wr + g) * 68 + t;
      const double* pb = smd + ((l + 1) % 3) * 64 * 68 + (16 * wc + g) * 68 + t;
      double c[2][2][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) c[i][j][0] = c[i][j][1] = 0.0;
#pragma unroll 4
      for (int k0 = 0; k0 < 4 * KSTEPS; k0 += 4) {
        double a[2], b[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) a[i] = pa[i * 8 * 68 + k0];
#pragma unroll
        for (int j = 0; j < 2; ++j) b[j] = pb[j * 8 * 68 + k0];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          keep[i][j][0] = fma(c[i][j][0], 0.5, keep[i][j][0]);
          keep[i][j][1] = fma(c[i][j][1], 0.5, keep[i][j][1]);
        }
    }
    if (SYNC) __syncthreads();
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) s += keep[i][j][0] + keep[i][j][1];
  if (s == 123.456) out[0] = s;
}

// Same row loop plus a scalar FP64 "epilogue" per row: NEXP table-assisted exps per lane on the
// accumulators of the first loop (8 FP64 ops + 1 LDS each, like the real kernel), a multiply and
// two adds per element.  Measures how much DMMA throughput a realistic DMMA/DFMA mix can reach.
template <int NEXP, bool SKEW>
__global__ void k_rowloop_mix(int rows, double* out) {
  extern __shared__ double smd[];
  double* tab = smd + 3 * 64 * 68;
  for (int i = threadIdx.x; i < 3 * 64 * 68; i += blockDim.x) smd[i] = 1e-3 * (i % 97);
  rgp::exp_table_init(tab, threadIdx.x);
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int wr = wid >> 2, wc = wid & 3;
  const bool grpB = SKEW && ((wid >> 2) & 1);
  double keep[2][2][2], sums = 0.0;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) keep[i][j][0] = keep[i][j][1] = 0.0;
  auto mma_loop = [&](int l, double (&c)[2][2][2]) {
    const double* pa = smd + (l % 3) * 64 * 68 + (16 * wr + g) * 68 + t;
    const double* pb = smd + ((l + 1) % 3) * 64 * 68 + (16 * wc + g) * 68 + t;
#pragma unroll 4
    for (int k0 = 0; k0 < 64; k0 += 4) {
      double a[2], b[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = pa[i * 8 * 68 + k0];
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = pb[j * 8 * 68 + k0];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
  };
  auto epilogue = [&](double (&c)[2][2][2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (i * 4 + j * 2 + e < NEXP) {
            double l = 1.0001 * rgp::exp_tab(-1e-3 * c[i][j][e] - 0.5, tab);
            sums += l;
            keep[i][j][e] += l;
          }
  };
  for (int n = 0; n < rows; ++n) {
    double c0[2][2][2], c1[2][2][2], c2[2][2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) c0[i][j][e] = c1[i][j][e] = c2[i][j][e] = 0.0;
    if (!grpB) {
      mma_loop(0, c0); epilogue(c0); mma_loop(1, c1); mma_loop(2, c2);
    } else {
      mma_loop(1, c1); mma_loop(0, c0); epilogue(c0); mma_loop(2, c2);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        keep[i][j][0] = fma(c1[i][j][0], 0.5, keep[i][j][0]) + c2[i][j][0];
        keep[i][j][1] = fma(c1[i][j][1], 0.5, keep[i][j][1]) + c2[i][j][1];
      }
    __syncthreads();
  }
  double s = sums;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) s += keep[i][j][0] + keep[i][j][1];
  if (s == 123.456) out[0] = s;
}

template <int NEXP, bool SKEW>
static void probe_rowloop_mix(int sms, double* out) {
  const int smem = (3 * 64 * 68 + 256) * 8;
  CK(cudaFuncSetAttribute(k_rowloop_mix<NEXP, SKEW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int rows = 6000;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  k_rowloop_mix<NEXP, SKEW><<<sms, 512, smem>>>(rows, out);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_rowloop_mix<NEXP, SKEW><<<sms, 512, smem>>>(rows, out);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  double dmma_per_sched = (double)rows * 3 * 16 * 4 * 4;
  double cycles = ms * 1e-3 * 1.965e9;
  double scalar_cycles = (double)rows * 4 * (NEXP * 11 + 16) * 2;     // per scheduler: 4 warps x FP64 scalar instrs x 2 cycles
  printf("{\"probe\": \"rowloop_mix\", \"exps_per_lane\": %d, \"skew\": %d, \"dmma_frac_of_peak\": %.3f, "
         "\"dmma_plus_scalar_frac\": %.3f}\n", NEXP, (int)SKEW, 16.0 / (cycles / dmma_per_sched),
         (dmma_per_sched * 16.0 + scalar_cycles) / cycles);
}

template <int LOOPS, int KSTEPS, bool SYNC>
static void probe_rowloop(int sms, double* out) {
  const int smem = 3 * 64 * 68 * 8;
  CK(cudaFuncSetAttribute(k_rowloop<LOOPS, KSTEPS, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int rows = 6000 * 3 * 16 / (LOOPS * KSTEPS);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  k_rowloop<LOOPS, KSTEPS, SYNC><<<sms, 512, smem>>>(rows, out);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_rowloop<LOOPS, KSTEPS, SYNC><<<sms, 512, smem>>>(rows, out);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  double dmma_per_sched = (double)rows * LOOPS * KSTEPS * 4 * 4;   // 4 warps/scheduler, 4 DMMA per k-step
  double cycles = ms * 1e-3 * 1.965e9;
  printf("{\"probe\": \"rowloop\", \"loops_per_row\": %d, \"ksteps\": %d, \"sync\": %d, \"frac_of_dmma_peak\": %.3f}\n",
         LOOPS, KSTEPS, (int)SYNC, 16.0 / (cycles / dmma_per_sched));
}

template <typename F>
static float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
  double* out;
  CK(cudaMalloc(&out, 1024));
  long long* cyc;
  CK(cudaMalloc(&cyc, sizeof(long long) * 1024));

  // 1. DFMA peak vs chains / occupancy
  {
    const int outer = 64;
    for (int bps : {1, 2, 4, 8}) {
      int blocks = sms * bps;
      float ms8 = time_ms([&] { k_dfma<8><<<blocks, 256>>>(outer, out); });
      float ms16 = time_ms([&] { k_dfma<16><<<blocks, 256>>>(outer, out); });
      double f8 = 2.0 * blocks * 256.0 * outer * 256 * 8, f16 = 2.0 * blocks * 256.0 * outer * 256 * 16;
      printf("{\"probe\": \"dfma\", \"blocks_per_sm\": %d, \"tflops_ch8\": %.3f, \"tflops_ch16\": %.3f}\n",
             bps, f8 / ms8 / 1e9, f16 / ms16 / 1e9);
    }
    // sustained: ~2 s of back-to-back launches
    int blocks = sms * 8;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n = 200;
    for (int i = 0; i < n; ++i) k_dfma<16><<<blocks, 256>>>(64, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("{\"probe\": \"dfma_sustained\", \"seconds\": %.3f, \"tflops\": %.3f}\n", ms / 1e3,
           2.0 * blocks * 256.0 * 64 * 256 * 16 * n / ms / 1e9);
  }
  // 2. DMMA
  {
    const int outer = 32;
    int blocks = sms * 8;
    double mma_flops = 2.0 * 256.0 * (double)blocks * (256 / 32) * outer * 64 * 8;  // 256 FMA / warp-instr
    float ms = time_ms([&] { k_dmma<0, 0><<<blocks, 256>>>(outer, out); });
    printf("{\"probe\": \"dmma884\", \"tflops\": %.3f}\n", mma_flops / ms / 1e9);
    float ms4 = time_ms([&] { k_dmma<1, 4><<<blocks, 256>>>(outer, out); });
    double fma4 = 2.0 * blocks * 256.0 * outer * 64 * 8 * 4;
    printf("{\"probe\": \"dmma884+4dfma\", \"tflops_total\": %.3f, \"tflops_mma\": %.3f, \"tflops_fma\": %.3f}\n",
           (mma_flops + fma4) / ms4 / 1e9, mma_flops / ms4 / 1e9, fma4 / ms4 / 1e9);
    float ms8 = time_ms([&] { k_dmma<1, 8><<<blocks, 256>>>(outer, out); });
    double fma8 = 2.0 * blocks * 256.0 * outer * 64 * 8 * 8;
    printf("{\"probe\": \"dmma884+8dfma\", \"tflops_total\": %.3f, \"tflops_mma\": %.3f, \"tflops_fma\": %.3f}\n",
           (mma_flops + fma8) / ms8 / 1e9, mma_flops / ms8 / 1e9, fma8 / ms8 / 1e9);
  }
  // 3. LDS.128 vs distinct addresses
  {
    CK(cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int warps : {1, 4, 16, 32}) {
      for (int distinct : {1, 2, 4, 8, 32}) {
        int iters = 2000;
        k_lds<<<1, warps * 32, 65536>>>(iters, distinct, 1, out, cyc);
        CK(cudaDeviceSynchronize());
        long long c;
        CK(cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost));
        printf("{\"probe\": \"lds128\", \"warps\": %d, \"distinct\": %d, \"cycles_per_lds_per_sm\": %.3f}\n",
               warps, distinct, (double)c / (iters * 16.0 * warps));
      }
    }
  }
  // 8. realistic DMMA + scalar-epilogue mix
  probe_rowloop_mix<0, false>(sms, out);
  probe_rowloop_mix<4, false>(sms, out);
  probe_rowloop_mix<8, false>(sms, out);
  probe_rowloop_mix<8, true>(sms, out);
  // 7. short-loop / phase-boundary cost
  probe_rowloop<3, 16, false>(sms, out);
  probe_rowloop<3, 16, true>(sms, out);
  probe_rowloop<1, 48, false>(sms, out);
  probe_rowloop<1, 48, true>(sms, out);
  probe_rowloop<6, 8, false>(sms, out);
  // 6. DMMA fed from shared memory
  probe_dmma_lds<1, 1>(sms, out);
  probe_dmma_lds<2, 2>(sms, out);
  probe_dmma_lds<2, 4>(sms, out);
  probe_dmma_lds<4, 4>(sms, out);
  // 5. DMMA latency / accumulators in flight
  probe_inflight<1>(sms, out);
  probe_inflight<2>(sms, out);
  probe_inflight<4>(sms, out);
  probe_inflight<8>(sms, out);
  // 4. exp
  {
    int blocks = sms * 8, outer = 16;
    double n = (double)blocks * 256 * outer * 64 * 8;
    float m0 = time_ms([&] { k_exp<0><<<blocks, 256>>>(outer, out); });
    float m1 = time_ms([&] { k_exp<1><<<blocks, 256>>>(outer, out); });
    printf("{\"probe\": \"exp\", \"lib_Gexp_s\": %.2f, \"exp_neg_Gexp_s\": %.2f}\n", n / m0 / 1e6, n / m1 / 1e6);
  }
  return 0;
}
