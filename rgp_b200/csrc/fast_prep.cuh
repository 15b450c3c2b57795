// fast_prep.cuh - everything of the fast path that is O(N*M*Q) or smaller: centring,
// padded/tiled operand builds, per-row vectors, the three small GEMMs and the per-row /
// final combiners.  The O(N*M^2*Q) work is in psi2_kernels.cuh.
//
// Algebra (DESIGN.md "Factorised exponent").  With o_q = mean_m Z_mq, Z' = Z - o,
// mu' = mu - o (the statistics only depend on differences), d = 1/(2S+l^2),
// e = 1/(S+l^2):
//   log Psi1[n,m]/s2 = b1_n + sum_q [ (e mu')_q Z'_mq - 1/2 e_q Z'_mq^2 ]
//        b1_n = -1/2 sum_q log1p(S/l^2) - 1/2 sum_q e mu'^2
//   log P_n[m,m']/s2^2 = H_nm + H_nm' + sum_q ws_nq Z'_mq Z'_m'q
//        ws_nq = S/(l^2 (2S+l^2)) >= 0       ( = -w,  w = d/2 - 1/(2 l^2) )
//        H_nm = b2_n + sum_q [ (d mu')_q Z'_mq - 1/4 (d_q + 1/l_q^2) Z'_mq^2 ]
//        b2_n = -1/4 sum_q log1p(2S/l^2) - 1/2 sum_q d mu'^2
// so the Z-Z' term of SURVEY.md 8(a3) is absorbed into w and H and needs no separate
// "tail" in the backward pass.
#pragma once
#include "common.cuh"

namespace rgp {
namespace fast {

constexpr int BM = 64;            // inducing-point tile
constexpr double NEG_BIG = -1.0e300;

// o[q] = mean_m Z[m,q]   (one block per q)
__global__ void k_center(int M, int Q, const double* __restrict__ Z, double* __restrict__ o) {
  __shared__ double scratch[33];
  int q = blockIdx.x;
  double acc = 0.0;
  for (int m = threadIdx.x; m < M; m += blockDim.x) acc += Z[m * Q + q];
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) o[q] = acc / (double)M;
}

// Zt[tile][64][RS]: centred, zero padded, ready for a straight copy into shared memory.
// ZB[Mp][2QC] = [Z' | Z'^2].
__global__ void k_build_Z(int M, int Mp, int Q, int QC, int RS, const double* __restrict__ Z,
                          const double* __restrict__ o, double* __restrict__ Zt,
                          double* __restrict__ ZB) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Mp * RS) return;
  int m = idx / RS, c = idx - m * RS;
  double v = (m < M && c < Q) ? Z[m * Q + c] - o[c] : 0.0;
  Zt[idx] = v;                                  // [m/64][m%64][c] == [m][c] with stride RS
  if (c < QC) {
    ZB[m * 2 * QC + c] = v;
    ZB[m * 2 * QC + QC + c] = v * v;
  }
}

// Ct[b][64][64] = s2^2 * (dL + dL^T)/2 on block b = (I,J), I <= J, zero padded.
__global__ void k_build_C(int M, int nt, const double* __restrict__ dL, double v2,
                          double* __restrict__ Ct) {
  int b = blockIdx.x;
  int I = 0, rem = b;
  while (rem >= nt - I) { rem -= nt - I; ++I; }
  int J = I + rem;
  for (int idx = threadIdx.x; idx < 4096; idx += blockDim.x) {
    int r = idx >> 6, c = idx & 63;
    int m = I * 64 + r, mp = J * 64 + c;
    double v = 0.0;
    if (m < M && mp < M) v = v2 * 0.5 * (dL[m * M + mp] + dL[mp * M + m]);
    Ct[(size_t)b * 4096 + idx] = v;
  }
}

// One warp per row.  Writes ws[n][QC] (= -w), A2[n][2QC] = [d mu' | -1/4 (d+1/l2)], b2[n],
// and (if A1) A1[n][2QC] = [e mu' | -1/2 e], b1[n].
__global__ void k_rowprep(int64_t N, int Q, int QC, const double* __restrict__ mu,
                          const double* __restrict__ S, const double* __restrict__ ell,
                          const double* __restrict__ o, double* __restrict__ w,
                          double* __restrict__ A2, double* __restrict__ b2,
                          double* __restrict__ A1, double* __restrict__ b1) {
  int lane = threadIdx.x & 31;
  int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  double lg2 = 0.0, al2 = 0.0, lg1 = 0.0, al1 = 0.0;
  for (int q = lane; q < QC; q += 32) {
    double wv = 0.0, a2a = 0.0, a2b = 0.0, a1a = 0.0, a1b = 0.0;
    if (q < Q) {
      double l = ell[q], l2 = l * l;
      double s = S[n * Q + q], m = mu[n * Q + q] - o[q];
      double den = 2.0 * s + l2, d = 1.0 / den;
      wv = s / (l2 * den);                      // ws = -w >= 0
      a2a = d * m;
      a2b = -0.25 * (d + 1.0 / l2);
      lg2 += log1p(2.0 * s / l2);
      al2 = fma(a2a, m, al2);
      double e = 1.0 / (s + l2);
      a1a = e * m;
      a1b = -0.5 * e;
      lg1 += log1p(s / l2);
      al1 = fma(a1a, m, al1);
    }
    w[n * QC + q] = wv;
    A2[n * 2 * QC + q] = a2a;
    A2[n * 2 * QC + QC + q] = a2b;
    if (A1) {
      A1[n * 2 * QC + q] = a1a;
      A1[n * 2 * QC + QC + q] = a1b;
    }
  }
  lg2 = warp_sum(lg2);
  al2 = warp_sum(al2);
  lg1 = warp_sum(lg1);
  al1 = warp_sum(al1);
  if (lane == 0) {
    b2[n] = -0.25 * lg2 - 0.5 * al2;
    if (b1) b1[n] = -0.5 * lg1 - 0.5 * al1;
  }
}

// ---------------------------------------------------------------------------------
// Small tiled DGEMM on DMMA: C(i,j) = sum_k A(i,k) B(j,k), 64x64 tile per CTA, 8 warps, each a 16 x 32
// sub-tile of 2 x 4 m8n8k4 accumulators (the tiling of the Psi2 kernels), K staged 16 at a time through
// shared memory with the next chunk prefetched into registers.  Operand element (r,k) lives at
// ptr[r*sr + k*sk]; *_KC says whether k (true) or r (false) is the unit-stride index, which decides how
// the staging loads are coalesced AND the shared layout: k-major operands are kept [r][k] (stride 20),
// r-major ones [k][r] (stride 68) - both strides == 4 (mod 16), so every fragment load hits each bank
// exactly twice and the staging stores are contiguous.  blockIdx.z splits K.  (Round 1 used 4x4 DFMA register
// tiles here: 42-48 % of the FP64 peak; these five GEMMs are 2.5 % of the headline step, 10 % at M = 100.)
// ---------------------------------------------------------------------------------
struct GemmOperand {
  const double* p;
  int64_t sr, sk;
  int rows;       // valid rows (others read as 0)
};

enum { EPI_HP = 0, EPI_PSI1 = 1, EPI_L1 = 2, EPI_PLAIN = 3 };

struct GemmEpi {
  int mode;
  const double* bias;      // [rows_i]
  double variance;
  const double* scale;     // dL_dpsi1 (ld = scale_ld) for EPI_L1
  int64_t scale_ld;
  int M;                   // valid columns
  double* out;
  int64_t out_ld;          // EPI_PSI1/L1/PLAIN: row stride;  EPI_HP: rows per tile (rc)
  int64_t split_stride;    // EPI_PLAIN with split-K: elements between partial outputs
};

constexpr int GEMM_SK = 20;    // [r][k] layout: 16 k + 4
constexpr int GEMM_SR = 68;    // [k][r] layout: 64 r + 4
constexpr int GEMM_TILE = 64 * GEMM_SK;   // >= 16 * GEMM_SR

// global -> registers: the 4 elements of a 64 x 16 operand chunk this thread stages
template <bool KC>
__device__ __forceinline__ void gemm_fetch(const GemmOperand& op, int64_t r0, int64_t k0, int64_t kend, int tid,
                                           double (&v)[4]) {
  if (KC) {
    const int kk = tid & 15, rr = tid >> 4;       // 16 k x 16 rows per pass
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int64_t gr = r0 + rr + 16 * p, gk = k0 + kk;
      v[p] = (gr < op.rows && gk < kend) ? op.p[gr * op.sr + gk * op.sk] : 0.0;
    }
  } else {
    const int rr = tid & 63, kk = tid >> 6;       // 64 rows x 4 k per pass
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int64_t gr = r0 + rr, gk = k0 + kk + 4 * p;
      v[p] = (gr < op.rows && gk < kend) ? op.p[gr * op.sr + gk * op.sk] : 0.0;
    }
  }
}

// registers -> shared, element (r, k) at sm[r * GEMM_SK + k] (KC) or sm[k * GEMM_SR + r]
template <bool KC>
__device__ __forceinline__ void gemm_store(double* sm, int tid, const double (&v)[4]) {
  if (KC) {
    const int kk = tid & 15, rr = tid >> 4;
#pragma unroll
    for (int p = 0; p < 4; ++p) sm[(rr + 16 * p) * GEMM_SK + kk] = v[p];
  } else {
    const int rr = tid & 63, kk = tid >> 6;
#pragma unroll
    for (int p = 0; p < 4; ++p) sm[(kk + 4 * p) * GEMM_SR + rr] = v[p];
  }
}

template <bool KC>
__device__ __forceinline__ double gemm_frag(const double* sm, int r, int k) {
  return KC ? sm[r * GEMM_SK + k] : sm[k * GEMM_SR + r];
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256) k_gemm(GemmOperand A, GemmOperand B, int64_t K, GemmEpi epi) {
  __shared__ __align__(16) double As[GEMM_TILE];
  __shared__ __align__(16) double Bs[GEMM_TILE];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int wr = wid >> 1, wc = wid & 1, g = lane >> 2, t = lane & 3;
  const int64_t i0 = (int64_t)blockIdx.x * 64, j0 = (int64_t)blockIdx.y * 64;
  int64_t kper = (K + gridDim.z - 1) / gridDim.z;
  kper = (kper + 15) / 16 * 16;
  const int64_t kbeg = kper * blockIdx.z, kend = (kbeg + kper < K) ? kbeg + kper : K;
  double acc[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double va[4], vb[4];
  if (kbeg < kend) {
    gemm_fetch<A_KC>(A, i0, kbeg, kend, tid, va);
    gemm_fetch<B_KC>(B, j0, kbeg, kend, tid, vb);
  }
  for (int64_t k0 = kbeg; k0 < kend; k0 += 16) {
    __syncthreads();                              // previous chunk's fragments all read
    gemm_store<A_KC>(As, tid, va);
    gemm_store<B_KC>(Bs, tid, vb);
    __syncthreads();
    if (k0 + 16 < kend) {                         // next chunk in flight while this one is multiplied
      gemm_fetch<A_KC>(A, i0, k0 + 16, kend, tid, va);
      gemm_fetch<B_KC>(B, j0, k0 + 16, kend, tid, vb);
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      double a[2], b[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = gemm_frag<A_KC>(As, 16 * wr + 8 * i + g, 4 * ks + t);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = gemm_frag<B_KC>(Bs, 32 * wc + 8 * j + g, 4 * ks + t);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int64_t row = i0 + 16 * wr + 8 * i + g;
    if (row >= A.rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cl = 32 * wc + 8 * j + 2 * t + e;   // column inside the 64-wide tile
        const int64_t col = j0 + cl;
        const double val = acc[i][j][e];
        if (epi.mode == EPI_HP) {
          // HP[tile][row][64]; padded inducing points get a hugely negative exponent
          const double h = (col < epi.M) ? epi.bias[row] + val : NEG_BIG;
          epi.out[((int64_t)blockIdx.y * epi.out_ld + row) * 64 + cl] = h;
        } else if (epi.mode == EPI_PSI1) {
          if (col < epi.M) epi.out[row * epi.out_ld + col] = epi.variance * exp_neg(epi.bias[row] + val);
        } else if (epi.mode == EPI_L1) {
          // L1[row][Mp], zero in the padding
          double l = 0.0;
          if (col < epi.M)
            l = epi.scale[row * epi.scale_ld + col] * epi.variance * exp_neg(epi.bias[row] + val);
          epi.out[row * epi.out_ld + col] = l;
        } else {
          if (col < B.rows) epi.out[(int64_t)blockIdx.z * epi.split_stride + row * epi.out_ld + col] = val;
        }
      }
  }
}

// ---------------------------------------------------------------------------------
// Per-row combiner of the backward pass: one warp per row.
//   lam  [G][rc][Mp]  row sums of L (psi2)      Wq [G][rc][QC]
//   R2   [rc][2QC] = [lam Z' | lam Z'^2]         L1 [rc][Mp], R1 [rc][2QC] (psi1; may be null)
// writes dmu,dS [rc][Q]; per-CTA partial sums of dell[q] and dvar into part[cta][QC+1].
// Also collapses lam over the G block groups into lam[0] (used by the TN GEMM next).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_rows_finalize(
    int64_t rc, int M, int Mp, int Q, int QC, int G, const double* __restrict__ mu,
    const double* __restrict__ S, const double* __restrict__ ell, const double* __restrict__ o,
    double variance, double* __restrict__ lam, const double* __restrict__ Wq,
    const double* __restrict__ R2, const double* __restrict__ L1, const double* __restrict__ R1,
    const double* __restrict__ dL0, double dL0c, double* __restrict__ dmu, double* __restrict__ dS,
    double* __restrict__ part) {
  extern __shared__ double sacc[];              // [4 warps][QC+1]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double* my = sacc + wid * (QC + 1);
  for (int i = lane; i < QC + 1; i += 32) my[i] = 0.0;
  __syncwarp();
  const int64_t nwarps = (int64_t)gridDim.x * 4;
  for (int64_t n = (int64_t)blockIdx.x * 4 + wid; n < rc; n += nwarps) {
    // Lambda = sum_m lam (summing the block groups), Lambda1 = sum_m L1
    double Lam = 0.0, Lam1 = 0.0;
    for (int m = lane; m < Mp; m += 32) {
      double v = lam[n * Mp + m];
      for (int g = 1; g < G; ++g) v += lam[((int64_t)g * rc + n) * Mp + m];
      if (G > 1) lam[n * Mp + m] = v;
      Lam += v;
      if (L1) Lam1 += L1[n * Mp + m];
    }
    Lam = warp_sum(Lam);
    Lam1 = warp_sum(Lam1);
    for (int q = lane; q < Q; q += 32) {
      double l = ell[q], l2 = l * l;
      double s = S[n * Q + q], m = mu[n * Q + q] - o[q];
      double den = 2.0 * s + l2, d = 1.0 / den;
      double U = R2[n * 2 * QC + q], V = R2[n * 2 * QC + QC + q];
      double W = 0.0;
      for (int g = 0; g < G; ++g) W += Wq[((int64_t)g * rc + n) * QC + q];
      double quad = 2.0 * m * m * Lam - 4.0 * m * U + V + W;
      double gmu = -2.0 * d * (m * Lam - U);
      double gS = -d * Lam + d * d * quad;
      double gl = Lam * 2.0 * s / (l * den) + l * d * d * quad + (V - W) / (l2 * l);
      if (L1) {
        double e = 1.0 / (s + l2);
        double LZ = R1[n * 2 * QC + q], LZ2 = R1[n * 2 * QC + QC + q];
        double A = m * Lam1 - LZ;
        double B = m * m * Lam1 - 2.0 * m * LZ + LZ2;
        gmu += -e * A;
        gS += 0.5 * e * (e * B - Lam1);
        gl += l * e * (e * B + (s / l2) * Lam1);
      }
      dmu[n * Q + q] = gmu;
      dS[n * Q + q] = gS;
      my[q] += gl;
    }
    if (lane == 0) my[QC] += (2.0 * Lam + Lam1) / variance + (dL0 ? dL0[n] : dL0c);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < QC + 1; i += blockDim.x)
    part[(int64_t)blockIdx.x * (QC + 1) + i] =
        sacc[i] + sacc[(QC + 1) + i] + sacc[2 * (QC + 1) + i] + sacc[3 * (QC + 1) + i];
}

// Final small combine for one row chunk (outputs accumulate across chunks):
//   dZ[m,q]  += 2 Gl[m,q] + 4 Z'[m,q] Gl[m,QC+q] + 2 ACC[m,q]  +  GL[m,q] + 2 Z' GL[m,QC+q]
//   (ACC = sum_n ws (L_n Z') = -sum_n w (L_n Z'), hence the plus sign)
//   dell[q]  += sum_cta part[cta][q];   dvar += sum_cta part[cta][QC]
// Gl / GL are split-K partials [splits][Mp][2QC]; ACC partials [ncta][Mp][QC].
__global__ void k_final_small(int M, int Mp, int Q, int QC, const double* __restrict__ ZB,
                              const double* __restrict__ Gl, int splits_l,
                              const double* __restrict__ GL, int splits_L,
                              const double* __restrict__ ACCp, int ncta,
                              const double* __restrict__ part, int nparts,
                              double* __restrict__ dZ, double* __restrict__ dell,
                              double* __restrict__ dvar) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < M * Q) {
    int m = idx / Q, q = idx - m * Q;
    double z = ZB[m * 2 * QC + q];
    double ga = 0.0, gb = 0.0, Ga = 0.0, Gb = 0.0, acc = 0.0;
    for (int s = 0; s < splits_l; ++s) {
      ga += Gl[((int64_t)s * Mp + m) * 2 * QC + q];
      gb += Gl[((int64_t)s * Mp + m) * 2 * QC + QC + q];
    }
    if (GL)
      for (int s = 0; s < splits_L; ++s) {
        Ga += GL[((int64_t)s * Mp + m) * 2 * QC + q];
        Gb += GL[((int64_t)s * Mp + m) * 2 * QC + QC + q];
      }
    {                                            // CTA partials in fixed order, eight loads in flight
      const double* p = ACCp + (int64_t)m * QC + q;
      const int64_t stride = (int64_t)Mp * QC;
      int c = 0;
      for (; c + 8 <= ncta; c += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = p[(c + u) * stride];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
      }
      for (; c < ncta; ++c) acc += p[c * stride];
    }
    dZ[idx] += 2.0 * ga + 4.0 * z * gb + 2.0 * acc + Ga + 2.0 * z * Gb;
  }
  if (idx < Q + 1) {
    int col = (idx < Q) ? idx : QC;
    double s = 0.0;
    for (int c = 0; c < nparts; ++c) s += part[(int64_t)c * (QC + 1) + col];
    if (idx < Q) dell[idx] += s; else dvar[0] += s;
  }
}

// Psi2[m,m'] (+)= s2^2 * sum_r P2p[b][r][64][64], mirrored.  grid = (blocks b, 16 slices of the 64 x 64 tile):
// one tile element per thread, the R partials summed in fixed order (deterministic) with eight loads in flight
// (a single dependent chain over R = 148 ... 296 partials made this the longest kernel of a small evaluation).
__global__ void __launch_bounds__(256) k_psi2_reduce(int M, int nt, int R, double v2,
                                                    const double* __restrict__ P2p,
                                                    int accumulate, double* __restrict__ psi2) {
  int b = blockIdx.x;
  int I = 0, rem = b;
  while (rem >= nt - I) { rem -= nt - I; ++I; }
  int J = I + rem;
  const int idx = blockIdx.y * 256 + threadIdx.x;
  const int r = idx >> 6, c = idx & 63;
  const int m = I * 64 + r, mp = J * 64 + c;
  if (m >= M || mp >= M) return;
  const double* p = P2p + (int64_t)b * R * 4096 + idx;
  double s = 0.0;
  int k = 0;
  for (; k + 8 <= R; k += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = p[(int64_t)(k + u) * 4096];
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; k < R; ++k) s += p[(int64_t)k * 4096];
  s *= v2;
  if (accumulate) {
    psi2[m * M + mp] += s;
    if (I != J) psi2[mp * M + m] += s;
  } else {
    psi2[m * M + mp] = s;
    if (I != J) psi2[mp * M + m] = s;
  }
}

__global__ void k_fill(int64_t n, double v, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}

}  // namespace fast
}  // namespace rgp
