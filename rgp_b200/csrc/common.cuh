// common.cuh - shared device helpers for librgp_psi (sm_100a, fp64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rgp {

#define RGP_DEVINL __device__ __forceinline__

// Pad (doubles) of the shared-memory Z' tiles: row stride = QC + pad.
//   4: stride == 4 (mod 16): every 8-byte fragment load hits each bank exactly twice (the minimum for 256 B).
//   8: stride == 8 (mod 16): the stage-2 access patterns keep that property, and a 16-byte load of two
//      consecutive k columns is conflict free per quarter warp, so stage 1 fetches the fragments of TWO k-steps
//      with one LDS.128 (thread t holds k = k0 + 2t and k0 + 2t + 1; the first MMA contracts over the even columns
//      of the 8-wide step, the second over the odd ones - A and B use the same assignment, so the sum over k is
//      unchanged; an odd number of k-steps ends with one ordinary step).
// Measured (profiles/SUMMARY_r02.md): pad 8 gives the forward kernel +2.3 % and the backward kernel +1.4 % at
// M = 512, Q = 64 and +1.7 % at (1024, 128); at QC <= 32 (few k-steps per row) it costs 1-2 %, so those keep pad 4.
__host__ __device__ constexpr int tile_pad(int QC) { return QC >= 64 ? 8 : 4; }

// ---------------------------------------------------------------------------------
// exp(x) for the psi exponents (logs of quantities <= 1, so x <= ~0; positive x up to
// ~700 also works).  Range reduction x = k ln2 + r, |r| <= ln2/2 with a single FMA
// against the full-precision ln2 (error <= |k| * 4e-17, i.e. < 1e-15 for |x| < 40 and
// 2e-14 at |x| ~ 700), degree-10 near-minimax polynomial (Chebyshev interpolation
// fitted in 60-digit arithmetic; max relative error 3.9e-16 on the reduced range),
// exponent assembled with integer arithmetic.  13 FP64-pipe ops + one compare,
// versus ~25 for the CUDA library exp(); results below ~3e-308 flush to 0 (x < -708),
// which is what a sum over rows wants, and NaN/Inf garbage from the integer path at
// absurdly negative x (padded inducing points use -1e300) is discarded by the select.
// ---------------------------------------------------------------------------------
RGP_DEVINL double exp_neg(double x) {
  const double LOG2E = 1.4426950408889634074;
  const double LN2 = 6.93147180559945286227e-01;
  const double MAGIC = 6755399441055744.0;             // 2^52 + 2^51
  double kd = fma(x, LOG2E, MAGIC);
  int k = __double2loint(kd);
  double kf = kd - MAGIC;
  double r = fma(kf, -LN2, x);
  double p = 2.76263572414472227e-07;
  p = fma(p, r, 2.76401807962098502e-06);
  p = fma(p, r, 2.48015043469976862e-05);
  p = fma(p, r, 1.98411702704400671e-04);
  p = fma(p, r, 1.38888889324885988e-03);
  p = fma(p, r, 8.33333338566778249e-03);
  p = fma(p, r, 4.16666666665731419e-02);
  p = fma(p, r, 1.66666666665544055e-01);
  p = fma(p, r, 5.00000000000000555e-01);
  p = fma(p, r, 1.00000000000000666e+00);
  p = fma(p, r, 1.0);
  int hi = __double2hiint(p) + (k << 20);
  double res = __hiloint2double(hi, __double2loint(p));
  return (x < -708.0) ? 0.0 : res;
}

// Table-assisted variant for the two O(N M^2 Q) kernels: x = (256 k + j) ln2/256 + r with
// |r| <= ln2/512, exp(x) = 2^k * T[j] * e^r, T[j] = 2^(j/256) from a 2 KB shared-memory table
// (filled once per CTA with exp2()), e^r by a degree-4 Taylor polynomial (remainder
// r^5/120 < 4e-17).  8 FP64-pipe ops and one LDS; the underflow test is an integer compare
// on the high word (x < ~-708 -> 0), so it costs the FP64 pipe nothing.
RGP_DEVINL void exp_table_init(double* tab, int tid) {
  if (tid < 256) tab[tid] = exp2((double)tid * (1.0 / 256.0));
}

RGP_DEVINL double exp_tab(double x, const double* __restrict__ tab) {
  const double INV = 369.32993046757463;                // 256 / ln2
  const double STEP = 2.7076061740622863e-03;           // ln2 / 256
  const double MAGIC = 6755399441055744.0;              // 2^52 + 2^51
  double kd = fma(x, INV, MAGIC);
  int n = __double2loint(kd);
  double nf = kd - MAGIC;
  double r = fma(nf, -STEP, x);
  double p = fma(r, 4.16666666666666666667e-02, 1.66666666666666666667e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  double res = tab[n & 255] * p;
  int hi = __double2hiint(res) + ((n >> 8) << 20);
  res = __hiloint2double(hi, __double2loint(res));
  return ((unsigned)__double2hiint(x) > 0xC0862000u) ? 0.0 : res;
}

// Sum `v[0..7]` over the 8 lanes that differ in lane bits 2..4 (recursive halving: 7 shuffles
// instead of 24).  Returns, in every lane, the total of element c = 4*b4 + 2*b3 + b2 where
// b4,b3,b2 are that lane's bits 4,3,2.
RGP_DEVINL double reduce8_over_g(const double (&v)[8], int lane) {
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
  double u[4], w[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double send = h4 ? v[i] : v[i + 4];
    const double keep = h4 ? v[i + 4] : v[i];
    u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double send = h3 ? u[i] : u[i + 2];
    const double keep = h3 ? u[i + 2] : u[i];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const double send = h2 ? w[0] : w[1];
  const double keep = h2 ? w[1] : w[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 4);
}

// fp64 tensor-style MMA (DMMA.8x8x4 in SASS): D[8x8] += A[8x4] B[4x8]; lane (g = lane / 4, t = lane % 4) holds
// A[g][t], B[t][g] and D[g][2t], D[g][2t+1].  Runs on the FP64 pipe (same peak as DFMA, far fewer operand bytes).
RGP_DEVINL void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

RGP_DEVINL double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (all threads must call; result on every thread).
RGP_DEVINL double block_sum(double v, double* scratch /* >= 33 doubles */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = (lane < nw) ? scratch[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

}  // namespace rgp
