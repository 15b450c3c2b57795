// common.cuh - shared device helpers for librgp_psi (sm_100a, fp64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rgp {

#define RGP_DEVINL __device__ __forceinline__

// ---------------------------------------------------------------------------------
// exp for x <= ~0 (the psi exponents are logs of quantities <= 1; positive x up to
// ~700 also works).  Range reduction x = k ln2 + r, |r| <= ln2/2, degree-11 Taylor/
// Horner in fp64 (relative error < 1e-15 from the polynomial; the single-constant
// ln2 reduction adds <= |k| * 2^-53 * ... ~ 1e-14 relative at |x| ~ 700), exponent
// assembled with integer arithmetic.  13-14 FP64-pipe ops, versus ~25 for the CUDA
// library exp(), and it never touches the slow denormal path: results below
// ~2.2e-308 flush to 0, which is what the sum over rows wants anyway.
// ---------------------------------------------------------------------------------
RGP_DEVINL double exp_neg(double x) {
  const double LOG2E = 1.4426950408889634074;
  const double LN2_HI = 6.93147180369123816490e-01;   // high part of ln2 (fdlibm)
  const double LN2_LO = 1.90821492927058770002e-10;   // low part
  const double MAGIC = 6755399441055744.0;            // 2^52 + 2^51
  double xc = fmax(x, -708.0);
  double kd = fma(xc, LOG2E, MAGIC);
  int k = __double2loint(kd);
  double kf = kd - MAGIC;
  double r = fma(kf, -LN2_HI, xc);
  r = fma(kf, -LN2_LO, r);
  double p = 2.50521083854417187751e-08;               // 1/11!
  p = fma(p, r, 2.75573192239858906526e-07);           // 1/10!
  p = fma(p, r, 2.75573192239858906526e-06);           // 1/9!
  p = fma(p, r, 2.48015873015873015873e-05);           // 1/8!
  p = fma(p, r, 1.98412698412698412698e-04);           // 1/7!
  p = fma(p, r, 1.38888888888888888889e-03);           // 1/6!
  p = fma(p, r, 8.33333333333333333333e-03);           // 1/5!
  p = fma(p, r, 4.16666666666666666667e-02);           // 1/4!
  p = fma(p, r, 1.66666666666666666667e-01);           // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  int hi = __double2hiint(p) + (k << 20);
  double res = __hiloint2double(hi, __double2loint(p));
  return (x < -708.0) ? 0.0 : res;
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
