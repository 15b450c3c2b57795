// mixbench.cu - how DMMA (mma.sync.m8n8k4.f64) and scalar FP64 instructions share the FP64 pipe.
// Each warp runs, per iteration, DM DMMAs (8 independent accumulators, register operands) followed by
// SC scalar DFMAs (8 independent chains); W warps per CTA, one CTA per SM; optional CTA barrier per
// iteration (keeps the phases of all warps aligned).  Reported: measured cycles per iteration per
// scheduler against the ideal W/4 * (16 DM + 2 SC), i.e. the pipe time the instructions need.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mixbench mixbench.cu
#include <cuda_runtime.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dfma(double& f, double m, double c) {
  asm volatile("fma.rn.f64 %0, %0, %1, %2;\n" : "+d"(f) : "d"(m), "d"(c));
}

// SPLIT: 0 = every warp runs [DM DMMA][SC DFMA]; 1 = warp-specialised: even warps only DMMA (2 DM per
// iteration), odd warps only DFMA (2 SC per iteration) - same totals per scheduler pair
template <int DM, int SC, bool SYNC, int SPLIT>
__global__ void __launch_bounds__(1024) k_mix(int iters, double* out) {
  double c[8][2], f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = c[i][1] = 0.0; f[i] = 0.5 + i; }
  const double a = 1e-3 * threadIdx.x, b = 1.0 + 1e-4 * threadIdx.x, m = 0.999999, cc = 1e-7;
  const int wid = threadIdx.x >> 5;
  const bool do_m = SPLIT == 0 || ((wid >> 2) & 1) == 0;   // warps 0-3, 8-11: DMMA; 4-7, 12-15: DFMA (same schedulers)
  const bool do_s = SPLIT == 0 || ((wid >> 2) & 1) == 1;
  constexpr int MULT = SPLIT ? 2 : 1;
  // runs of >= 8 are loops (not unrolled) over bodies of 8, so ptxas cannot mix the two runs; shorter
  // runs are straight-line code and ptxas orders them as it likes (dump the SASS to see)
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (do_m) {
      if constexpr (DM * MULT >= 8) {
#pragma unroll 1
        for (int r = 0; r < DM * MULT / 8; ++r) {
#pragma unroll
          for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
        }
      } else {
#pragma unroll
        for (int i = 0; i < DM * MULT; ++i) dmma(c[i & 7][0], c[i & 7][1], a, b);
      }
    }
    if (do_s) {
      if constexpr (SC * MULT >= 8) {
#pragma unroll 1
        for (int r = 0; r < SC * MULT / 8; ++r) {
#pragma unroll
          for (int j = 0; j < 8; ++j) dfma(f[j], m, cc);
        }
#pragma unroll
        for (int j = 0; j < (SC * MULT) % 8; ++j) dfma(f[j], m, cc);
      } else {
#pragma unroll
        for (int j = 0; j < SC * MULT; ++j) dfma(f[j & 7], m, cc);
      }
    }
    if (SYNC) __syncthreads();
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
  if (s == 123.456) out[0] = s;
}

template <int DM, int SC, bool SYNC, int SPLIT>
static int run(int sms, int warps, double* out) {
  const long long work = 1 << 22;                      // ~DMMA-equivalents per warp
  int iters = (int)(work / (DM * 8 + SC + 1));
  if (iters < 16) iters = 16;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  k_mix<DM, SC, SYNC, SPLIT><<<sms, 32 * warps>>>(iters / 8, out);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    CK(cudaEventRecord(e0));
    k_mix<DM, SC, SYNC, SPLIT><<<sms, 32 * warps>>>(iters, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double cyc = best * 1e-3 * 1.965e9 / iters;    // cycles per iteration (per scheduler: warps/4 warps share it)
  const double ideal = warps / 4.0 * (16.0 * DM + 2.0 * SC);
  printf("{\"probe\": \"mix\", \"dmma_run\": %d, \"dfma_run\": %d, \"warps\": %d, \"sync\": %d, \"split\": %d, "
         "\"cycles_per_iter\": %.1f, \"ideal\": %.1f, \"pipe_eff\": %.3f, \"extra_cycles_per_dfma\": %.2f}\n",
         DM, SC, warps, (int)SYNC, SPLIT, cyc, ideal, ideal / cyc,
         SC ? (cyc - warps / 4.0 * 16.0 * DM * 1.0) / (warps / 4.0 * SC) - 2.0 : 0.0);
  return 0;
}

#define RUN(DM, SC, SYNC, SPLIT, W) if (run<DM, SC, SYNC, SPLIT>(sms, W, out)) return 1

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double* out;
  CK(cudaMalloc(&out, 1024));
  for (int w : {4, 8, 16}) {
    if (run<8, 0, false, 0>(sms, w, out)) return 1;
    if (run<0, 8, false, 0>(sms, w, out)) return 1;
    if (run<1, 1, false, 0>(sms, w, out)) return 1;
    if (run<2, 2, false, 0>(sms, w, out)) return 1;
    if (run<4, 4, false, 0>(sms, w, out)) return 1;
    if (run<8, 8, false, 0>(sms, w, out)) return 1;
    if (run<8, 1, false, 0>(sms, w, out)) return 1;
    if (run<8, 2, false, 0>(sms, w, out)) return 1;
    if (run<8, 4, false, 0>(sms, w, out)) return 1;
    if (run<32, 32, false, 0>(sms, w, out)) return 1;
    if (run<128, 128, false, 0>(sms, w, out)) return 1;
    if (run<384, 330, false, 0>(sms, w, out)) return 1;
    if (run<128, 128, true, 0>(sms, w, out)) return 1;
    if (run<384, 330, true, 0>(sms, w, out)) return 1;
    if (run<32, 32, true, 0>(sms, w, out)) return 1;
  }
  for (int w : {8, 16}) {
    if (run<8, 8, false, 1>(sms, w, out)) return 1;
    if (run<64, 64, false, 1>(sms, w, out)) return 1;
    if (run<64, 16, false, 1>(sms, w, out)) return 1;
    if (run<384, 330, false, 1>(sms, w, out)) return 1;
  }
  return 0;
}
