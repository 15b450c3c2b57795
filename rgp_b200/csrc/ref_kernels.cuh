// ref_kernels.cuh - simple "one thread per output" kernels written straight from the
// definitions in SURVEY.md section 8 (a1-a5): direct (mu - Zbar)^2 exponent, separate
// Z-Z' term, tails through the summed Psi2.  They share NO algebra with the fast path
// (psi2_fwd.cuh / psi2_bwd.cuh use the factorised exponent), so they double as an
// on-device cross-check at sizes the CPU oracle cannot reach, and they serve shapes
// the fast path does not cover (Q > 128).  Still CUDA: there is no CPU fallback.
#pragma once
#include "common.cuh"

namespace rgp {
namespace ref {

// c1[n] = -1/2 sum log(S/l2+1), c2[n] = -1/2 sum log(2S/l2+1), dinv[n,q] = 1/(2S+l2)
__global__ void row_terms(int64_t N, int Q, const double* __restrict__ S,
                          const double* __restrict__ ell, double* __restrict__ c1,
                          double* __restrict__ c2, double* __restrict__ dinv) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double a = 0.0, b = 0.0;
  for (int q = 0; q < Q; ++q) {
    double l2 = ell[q] * ell[q];
    double s = S[n * Q + q];
    a += log1p(s / l2);
    b += log1p(2.0 * s / l2);
    dinv[n * Q + q] = 1.0 / (2.0 * s + l2);
  }
  c1[n] = -0.5 * a;
  c2[n] = -0.5 * b;
}

// out[n,m] = scale[n,m] * variance * exp(c1[n] - 1/2 sum_q (mu-Z)^2/(S+l2)); scale may be null
__global__ void psi1(int64_t N, int M, int Q, const double* __restrict__ mu,
                     const double* __restrict__ S, const double* __restrict__ Z,
                     const double* __restrict__ ell, const double* __restrict__ c1,
                     double variance, const double* __restrict__ scale,
                     double* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * (int64_t)M) return;
  int64_t n = idx / M;
  int m = (int)(idx - n * M);
  double quad = 0.0;
  for (int q = 0; q < Q; ++q) {
    double l2 = ell[q] * ell[q];
    double d = mu[n * Q + q] - Z[m * Q + q];
    quad += d * d / (S[n * Q + q] + l2);
  }
  double v = variance * exp(c1[n] - 0.5 * quad);
  out[idx] = scale ? scale[idx] * v : v;
}

// zz[m,m'] = -sum_q (Z_m - Z_m')^2/(4 l2);  dLs = (dL + dL^T)/2 (if dL given)
__global__ void pair_terms(int M, int Q, const double* __restrict__ Z,
                           const double* __restrict__ ell, const double* __restrict__ dL,
                           double* __restrict__ zz, double* __restrict__ dLs) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * M) return;
  int m = idx / M, mp = idx - m * M;
  double acc = 0.0;
  for (int q = 0; q < Q; ++q) {
    double d = Z[m * Q + q] - Z[mp * Q + q];
    acc += d * d / (ell[q] * ell[q]);
  }
  zz[idx] = -0.25 * acc;
  if (dL) dLs[idx] = 0.5 * (dL[m * M + mp] + dL[mp * M + m]);
}

// psi2[m,m'] += sum_{n in split} variance^2 exp(c2[n] + zz - sum_q dinv (mu - zbar)^2)
// block (16,16) over (m',m); gridDim.z row splits; output must be zeroed first.
__global__ void psi2(int64_t N, int M, int Q, const double* __restrict__ mu,
                     const double* __restrict__ Z, const double* __restrict__ c2,
                     const double* __restrict__ dinv, const double* __restrict__ zz,
                     double variance, double* __restrict__ psi2_out) {
  int mp = blockIdx.x * 16 + threadIdx.x;
  int m = blockIdx.y * 16 + threadIdx.y;
  if (m >= M || mp >= M) return;
  int64_t per = (N + gridDim.z - 1) / gridDim.z;
  int64_t n0 = per * blockIdx.z, n1 = n0 + per < N ? n0 + per : N;
  double e1 = zz[m * M + mp];
  double acc = 0.0;
  for (int64_t n = n0; n < n1; ++n) {
    double s = 0.0;
    for (int q = 0; q < Q; ++q) {
      double zb = 0.5 * (Z[m * Q + q] + Z[mp * Q + q]);
      double d = mu[n * Q + q] - zb;
      s = fma(d * d, dinv[n * Q + q], s);
    }
    acc += exp(c2[n] + e1 - s);
  }
  atomicAdd(&psi2_out[m * M + mp], variance * variance * acc);
}

// Psi1 gradients, row-local part.  L1 = dL_dpsi1 * Psi1 precomputed.  thread per (n,q).
__global__ void psi1_bwd_rows(int64_t N, int M, int Q, const double* __restrict__ mu,
                              const double* __restrict__ S, const double* __restrict__ Z,
                              const double* __restrict__ ell, const double* __restrict__ L1,
                              double variance, double* __restrict__ dmu,
                              double* __restrict__ dS, double* __restrict__ dell,
                              double* __restrict__ dvar) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * (int64_t)Q) return;
  int64_t n = idx / Q;
  int q = (int)(idx - n * Q);
  double l = ell[q], l2 = l * l;
  double s = S[idx], m_ = mu[idx];
  double e = 1.0 / (s + l2);
  double A = 0.0, B = 0.0, Lam = 0.0;
  for (int m = 0; m < M; ++m) {
    double w = L1[n * M + m];
    double a = m_ - Z[m * Q + q];
    A = fma(w, a, A);
    B = fma(w * a, a, B);
    Lam += w;
  }
  dmu[idx] += -e * A;
  dS[idx] += 0.5 * e * (e * B - Lam);
  atomicAdd(&dell[q], l * e * (e * B + (s / l2) * Lam));
  if (q == 0) atomicAdd(dvar, Lam / variance);
}

// dZ[m,q] += sum_n L1[n,m] (mu-Z)/(S+l2).  thread per (m,q), gridDim.y row splits.
__global__ void psi1_bwd_Z(int64_t N, int M, int Q, const double* __restrict__ mu,
                           const double* __restrict__ S, const double* __restrict__ Z,
                           const double* __restrict__ ell, const double* __restrict__ L1,
                           double* __restrict__ dZ) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * Q) return;
  int m = idx / Q, q = idx - m * Q;
  int64_t per = (N + gridDim.y - 1) / gridDim.y;
  int64_t n0 = per * blockIdx.y, n1 = n0 + per < N ? n0 + per : N;
  double l2 = ell[q] * ell[q], z = Z[idx], acc = 0.0;
  for (int64_t n = n0; n < n1; ++n)
    acc = fma(L1[n * M + m], (mu[n * Q + q] - z) / (S[n * Q + q] + l2), acc);
  atomicAdd(&dZ[idx], acc);
}

// Psi2 gradients, one CTA per row (grid-stride).  Thread m walks m' and keeps
// lam_m = sum_m' L[m,m'] and tmp[q] = sum_m' L[m,m'] Z[m',q] (GPy's `tmp`).
template <int QMAX>
__global__ void psi2_bwd_rows(int64_t N, int M, int Q, const double* __restrict__ mu,
                              const double* __restrict__ S, const double* __restrict__ Z,
                              const double* __restrict__ ell, const double* __restrict__ c2,
                              const double* __restrict__ dinv, const double* __restrict__ zz,
                              const double* __restrict__ dLs, double variance,
                              double* __restrict__ dmu, double* __restrict__ dS,
                              double* __restrict__ dZ, double* __restrict__ dell,
                              double* __restrict__ dvar) {
  extern __shared__ double sm[];            // Lam, U[Q], V[Q], W[Q]
  double* sU = sm + 1;
  double* sV = sU + Q;
  double* sW = sV + Q;
  const double v2 = variance * variance;
  for (int64_t n = blockIdx.x; n < N; n += gridDim.x) {
    for (int i = threadIdx.x; i < 1 + 3 * Q; i += blockDim.x) sm[i] = 0.0;
    __syncthreads();
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
      double tmp[QMAX];
      for (int q = 0; q < Q; ++q) tmp[q] = 0.0;
      double lam = 0.0;
      for (int mp = 0; mp < M; ++mp) {
        double s = 0.0;
        for (int q = 0; q < Q; ++q) {
          double zb = 0.5 * (Z[m * Q + q] + Z[mp * Q + q]);
          double d = mu[n * Q + q] - zb;
          s = fma(d * d, dinv[n * Q + q], s);
        }
        double L = dLs[m * M + mp] * v2 * exp(c2[n] + zz[m * M + mp] - s);
        lam += L;
        for (int q = 0; q < Q; ++q) tmp[q] = fma(L, Z[mp * Q + q], tmp[q]);
      }
      atomicAdd(&sm[0], lam);
      for (int q = 0; q < Q; ++q) {
        double z = Z[m * Q + q], d = dinv[n * Q + q], mq = mu[n * Q + q];
        atomicAdd(&dZ[m * Q + q], d * (2.0 * mq * lam - lam * z - tmp[q]));
        atomicAdd(&sU[q], lam * z);
        atomicAdd(&sV[q], lam * z * z);
        atomicAdd(&sW[q], z * tmp[q]);
      }
    }
    __syncthreads();
    double Lam = sm[0];
    for (int q = threadIdx.x; q < Q; q += blockDim.x) {
      double d = dinv[n * Q + q], mq = mu[n * Q + q], s = S[n * Q + q], l = ell[q];
      double quad = 2.0 * mq * mq * Lam - 4.0 * mq * sU[q] + sV[q] + sW[q];
      dmu[n * Q + q] += -2.0 * d * (mq * Lam - sU[q]);
      dS[n * Q + q] += -d * Lam + d * d * quad;
      atomicAdd(&dell[q], Lam * 2.0 * s / (l * (2.0 * s + l * l)) + l * d * d * quad);
    }
    if (threadIdx.x == 0) atomicAdd(dvar, 2.0 * Lam / variance);
    __syncthreads();
  }
}

// Row-independent tails through the Z-Z' term: LN = dLs * Psi2.  thread per (m,q).
__global__ void psi2_bwd_tails(int M, int Q, const double* __restrict__ Z,
                               const double* __restrict__ ell, const double* __restrict__ dLs,
                               const double* __restrict__ psi2, double* __restrict__ dZ,
                               double* __restrict__ dell) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * Q) return;
  int m = idx / Q, q = idx - m * Q;
  double rs = 0.0, lnz = 0.0;
  for (int mp = 0; mp < M; ++mp) {
    double ln = dLs[m * M + mp] * psi2[m * M + mp];
    rs += ln;
    lnz = fma(ln, Z[mp * Q + q], lnz);
  }
  double l = ell[q], z = Z[idx];
  dZ[idx] += -(rs * z - lnz) / (l * l);
  atomicAdd(&dell[q], (rs * z * z - z * lnz) / (l * l * l));
}

__global__ void add_scalar(double v, double* __restrict__ out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] += v;
}

__global__ void fill(int64_t n, double v, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}

// dvar += sum_n dL_dpsi0[n]  (or N * const)
__global__ void sum_to(int64_t n, const double* __restrict__ x, double* __restrict__ out) {
  __shared__ double scratch[33];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    acc += x[i];
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

}  // namespace ref
}  // namespace rgp
