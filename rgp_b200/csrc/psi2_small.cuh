// psi2_small.cuh - Psi2 forward / backward for SMALL inducing sets (M <= 112, Q <= 48): the shapes of the
// reference's own models (M = 50 ... 100, Q = 10 ... 40; autoreg/benchmark/tasks.py:141-177,
// examples/walk_run_2_alex.py:364-393, svi_experiments/rgp_experiments.py).
//
// The 64 x 64 block kernels of psi2_kernels.cuh pay for padding there: M = 100 is computed as 128 (136 8x8 tiles
// of the pair matrix instead of 91), Q = 20 as 32 stage-2 columns, lambda / W go through 3 block passes of
// red.global.add, and 8 warps (2 per scheduler) cannot hide the DMMA issue latency of the short k loops.  Here ONE
// CTA of 16 warps holds the whole problem of a row:
//
//   Z' [Mp16][Qp + 4] and the FULL symmetric L_n [Mp16][Mp16 + 4] in shared memory (Mp16 = M rounded up to 16,
//   Qp = Q rounded up to 8);
//   stage 1   E = H_m + H_m' + sum_q (ws_q Z'_mq) Z'_m'q on the 16 x 16 supertiles of the upper triangle only
//             (Ms (Ms + 1) / 2 supertiles dealt round-robin to the warps), p = exp(E), Psi2 += p (registers),
//             L = C p stored with its mirror image;
//   stage 2   T = L Z' as (16 rows x Qp columns x 1/KS of the k range) jobs dealt round-robin to the warps:
//             acc[m,q] += ws_q T[m,q], W_q += sum_m Z'_mq T[m,q], lambda_m += row sums of the L fragments;
//   per row   lambda_n and W_n are complete inside the CTA: written once with plain stores (no atomics, no
//             block passes), in a fixed order (deterministic).
//
// DMMAs per row at M = 100, Q = 20: 560 + 1092 against 680 + 2048 in the block kernels.  The per-row vectors
// (ws[Qp], H[Mp16]) arrive by TMA bulk copies into a two-slot ring (SmallRowStage), as in the block kernels.
// Two CTA barriers per row (L complete / L consumed): L is single-buffered, 16 warps overlap the phases' tails.
#pragma once
#include "psi2_kernels.cuh"

namespace rgp {
namespace fast {

constexpr int PS_THREADS = 512;
constexpr int PS_WARPS = PS_THREADS / 32;
constexpr int PS_MS_MAX = 7;      // 16-row super rows: M <= 112
constexpr int PS_S1 = 2;          // supertiles per warp: 7 * 8 / 2 = 28 <= 2 * 16
constexpr int PS_VR = 4;          // rows per TMA batch
constexpr int PS_QT_MAX = 6;

// shared-memory size in bytes (host + device agree through this one function)
__host__ __device__ constexpr int small_smem_doubles(int Ms, int QT, bool bwd) {
  const int Mp16 = 16 * Ms, Qp = 8 * QT;
  int d = Mp16 * (Qp + 4) + 2 * PS_VR * (Qp + Mp16) + 258;
  if (bwd) d += 32 * Qp + 4 * Mp16 + Mp16 * (Mp16 + 4);
  return d;
}

// TMA staging of the per-row vectors [ws (Qp) | H (Mp16)]; the logic of RowVecStage with run-time sizes.
struct SmallRowStage {
  double* ring;
  uint64_t* mbar;
  uint32_t phase_bits = 0;
  int64_t r0 = 0, r1 = 0, next = 0, nb = 0;
  const double* wrow = nullptr;
  const double* hp = nullptr;
  int64_t htile = 0;          // doubles between two 64-wide tiles of HP
  int QC = 0, Qp = 0, Mp16 = 0, VB = 0;

  RGP_DEVINL void init_barriers(int tid) {
    if (tid == 0) {
      mbar_init(&mbar[0], 1);
      mbar_init(&mbar[1], 1);
      mbar_fence_init();
    }
  }
  RGP_DEVINL void begin(int64_t r0_, int64_t r1_, int tid) {
    r0 = r0_; r1 = r1_;
    nb = r1 > r0 ? (r1 - r0 + PS_VR - 1) / PS_VR : 0;
    next = 0;
    if (tid == 0 && nb > 0) issue();
  }
  RGP_DEVINL void issue() {
    const int64_t n0 = r0 + next * PS_VR;
    const int rows = (int)((r1 - n0 < PS_VR) ? r1 - n0 : PS_VR);
    const int slot = (int)(next & 1);
    double* dst = ring + slot * PS_VR * VB;
    const int h0 = Mp16 < 64 ? Mp16 : 64;
    mbar_expect_tx(&mbar[slot], (uint32_t)(rows * VB * 8));
    for (int w = 0; w < rows; ++w) {
      bulk_g2s(dst + w * VB, wrow + (n0 + w) * QC, Qp * 8, &mbar[slot]);
      bulk_g2s(dst + w * VB + Qp, hp + (n0 + w) * 64, h0 * 8, &mbar[slot]);
      if (Mp16 > 64) bulk_g2s(dst + w * VB + Qp + 64, hp + htile + (n0 + w) * 64, (Mp16 - 64) * 8, &mbar[slot]);
    }
    ++next;
  }
  // issuing thread, when every thread has finished all rows < r0 + idx
  RGP_DEVINL void refill(int64_t idx) {
    while (next < nb && next <= idx / PS_VR + 1 && (next - 1) * PS_VR <= idx) issue();
  }
  RGP_DEVINL const double* row(int64_t idx) {
    const int slot = (int)((idx / PS_VR) & 1);
    if (idx % PS_VR == 0) {
      mbar_wait(&mbar[slot], (phase_bits >> slot) & 1u);
      phase_bits ^= 1u << slot;
    }
    return ring + slot * PS_VR * VB + (idx % PS_VR) * VB;
  }
};

// MODE 0: forward only (Psi2 partials), 1: backward only, 2: backward + Psi2 partials (fused SVI pass).
// Outputs (strides of the block path, so the small GEMMs and combiners downstream are shared):
//   lam [rc][Mp]   Wq [rc][QC]            complete per row, plain stores
//   ACCp[cta * KS + kr][Mp][QC]           sum_n ws (L_n Z') over this CTA's rows and the k range kr (plain stores)
//   P2s [cta][Mp16][Mp16]                 sum_n p over this CTA's rows (MODE 0 / 2), full symmetric
template <int QT, int MODE, int JMAX>
__global__ void __launch_bounds__(PS_THREADS, 1)
k_psi2_small(int64_t rc, int Mp, int Ms, int nt, int qk, int QC, int KS, int RSz,
             const double* __restrict__ Zt, const double* __restrict__ Ct, const double* __restrict__ wrow,
             const double* __restrict__ HP, double* __restrict__ lam, double* __restrict__ Wq,
             double* __restrict__ ACCp, double* __restrict__ P2s) {
  constexpr int Qp = 8 * QT, RS = Qp + 4;
  constexpr bool BWD = MODE != 0, FWD = MODE != 1;
  const int Mp16 = 16 * Ms, RSL = Mp16 + 4, VB = Qp + Mp16;
  extern __shared__ __align__(16) double smem[];
  double* sZ = smem;                         // [Mp16][RS]
  double* sV = sZ + Mp16 * RS;               // row-vector ring: 2 slots x PS_VR rows x VB
  double* sT = sV + 2 * PS_VR * VB;          // exp table (256) + 2 mbarriers
  double* sW = sT + 258;                     // [jobs <= 32][Qp]   W partials of the row
  double* sLam = sW + 32 * Qp;               // [KS <= 4][Mp16]    lambda partials of the row
  double* sL = sLam + 4 * Mp16;              // [Mp16][RSL]        L_n, full symmetric

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t = lane & 3;
  const int R = gridDim.x;
  const int64_t per = (rc + R - 1) / R;
  const int64_t r0 = per * blockIdx.x, r1 = (r0 + per < rc) ? r0 + per : rc;

  exp_table_init(sT, tid);
  SmallRowStage rv;
  rv.ring = sV;
  rv.mbar = reinterpret_cast<uint64_t*>(sT + 256);
  rv.wrow = wrow;
  rv.hp = HP;
  rv.htile = rc * 64;
  rv.QC = QC; rv.Qp = Qp; rv.Mp16 = Mp16; rv.VB = VB;
  rv.init_barriers(tid);
  for (int idx = tid; idx < Mp16 * Qp; idx += PS_THREADS) {
    const int m = idx / Qp, c = idx - m * Qp;
    sZ[m * RS + c] = Zt[(size_t)m * RSz + c];
  }

  // this warp's supertiles of the upper triangle
  const int NS = Ms * (Ms + 1) / 2;
  int si[PS_S1], sj[PS_S1];
  int ns = 0;
#pragma unroll
  for (int s = 0; s < PS_S1; ++s) {
    const int u = wid + PS_WARPS * s;
    si[s] = sj[s] = 0;
    if (u < NS) {
      int i = 0, rem = u;
      while (rem >= Ms - i) { rem -= Ms - i; ++i; }
      si[s] = i;
      sj[s] = i + rem;
      ns = s + 1;
    }
  }
  double creg[PS_S1][2][2][2], pacc[PS_S1][2][2][2];
#pragma unroll
  for (int s = 0; s < PS_S1; ++s)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        pacc[s][i][j][0] = pacc[s][i][j][1] = 0.0;
        creg[s][i][j][0] = creg[s][i][j][1] = 0.0;
        if (BWD && s < ns) {
          const int m = 16 * si[s] + 8 * i + g, mp = 16 * sj[s] + 8 * j + 2 * t;
          const int I = m >> 6, J = mp >> 6;                       // I <= J (supertiles never straddle a 64-block)
          const int b = I * nt - I * (I - 1) / 2 + (J - I);
          const double2 c2 = *reinterpret_cast<const double2*>(Ct + (size_t)b * 4096 + (m & 63) * 64 + (mp & 63));
          creg[s][i][j][0] = c2.x;
          creg[s][i][j][1] = c2.y;
        }
      }
  // stage-2 jobs: job jb = (k range kr, 16-row strip sp)
  const int njobs = Ms * KS, kper = 4 * Ms / KS;       // k-steps per job
  double accZ[JMAX][2][QT][2];
#pragma unroll
  for (int jj = 0; jj < JMAX; ++jj)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < QT; ++j) accZ[jj][i][j][0] = accZ[jj][i][j][1] = 0.0;

  rv.begin(r0, r1, tid);
  __syncthreads();

  for (int64_t n = r0; n < r1; ++n) {
    if (tid == 0) rv.refill(n - r0);                 // every thread is past the last barrier of row n - 1
    const double* v = rv.row(n - r0);
    const double* H = v + Qp;
    // ------------------------------------------------------------------ stage 1 + exp (+ L)
#pragma unroll
    for (int s = 0; s < PS_S1; ++s) {
      if (s < ns) {
        double acc[2][2][2];
        const int mi = 16 * si[s] + g, mj = 16 * sj[s] + 2 * t;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const double hi = H[mi + 8 * i];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const double2 hj = *reinterpret_cast<const double2*>(H + mj + 8 * j);
            acc[i][j][0] = hi + hj.x;
            acc[i][j][1] = hi + hj.y;
          }
        }
        const double* pa = sZ + (16 * si[s] + g) * RS + t;
        const double* pb = sZ + (16 * sj[s] + g) * RS + t;
#pragma unroll 2
        for (int k0 = 0; k0 < qk; k0 += 4) {
          const double wv = v[k0 + t];
          const double a0 = pa[k0] * wv, a1 = pa[8 * RS + k0] * wv;
          const double b0 = pb[k0], b1 = pb[8 * RS + k0];
          dmma(acc[0][0][0], acc[0][0][1], a0, b0);
          dmma(acc[0][1][0], acc[0][1][1], a0, b1);
          dmma(acc[1][0][0], acc[1][0][1], a1, b0);
          dmma(acc[1][1][0], acc[1][1][1], a1, b1);
        }
        const bool offd = si[s] != sj[s];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const double p0 = exp_tab(acc[i][j][0], sT), p1 = exp_tab(acc[i][j][1], sT);
            if constexpr (FWD) {
              pacc[s][i][j][0] += p0;
              pacc[s][i][j][1] += p1;
            }
            if constexpr (BWD) {
              const double l0 = creg[s][i][j][0] * p0, l1 = creg[s][i][j][1] * p1;
              const int m = mi + 8 * i, mp = mj + 8 * j;
              *reinterpret_cast<double2*>(sL + m * RSL + mp) = make_double2(l0, l1);
              if (offd) {
                sL[mp * RSL + m] = l0;
                sL[(mp + 1) * RSL + m] = l1;
              }
            }
          }
      }
    }
    __syncthreads();                                  // BWD: L_n complete.  forward only: row n's vectors consumed
    if constexpr (BWD) {
      // ---------------------------------------------------------------- stage 2: T = L Z' by jobs
#pragma unroll
      for (int jj = 0; jj < JMAX; ++jj) {
        const int jb = wid + PS_WARPS * jj;
        if (jb < njobs) {
          const int kr = jb / Ms, sp = jb - kr * Ms;
          double T[2][QT][2];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < QT; ++j) T[i][j][0] = T[i][j][1] = 0.0;
          double ls0 = 0.0, ls1 = 0.0;
          const double* pa = sL + (16 * sp + g) * RSL + t;
          const double* pb = sZ + t * RS + g;
          const int kb = 4 * kr * kper, ke = kb + 4 * kper;
#pragma unroll 2
          for (int k0 = kb; k0 < ke; k0 += 4) {
            const double a0 = pa[k0], a1 = pa[8 * RSL + k0];
            double bq[QT];
#pragma unroll
            for (int j = 0; j < QT; ++j) bq[j] = pb[k0 * RS + 8 * j];
            ls0 += a0;
            ls1 += a1;
#pragma unroll
            for (int j = 0; j < QT; ++j) {
              dmma(T[0][j][0], T[0][j][1], a0, bq[j]);
              dmma(T[1][j][0], T[1][j][1], a1, bq[j]);
            }
          }
          // folds: acc += ws T ; W partial = sum over this strip's rows of Z' T
          double wp[2 * QT];
#pragma unroll
          for (int j = 0; j < QT; ++j) {
            const int q = 8 * j + 2 * t;
            const double2 wq = *reinterpret_cast<const double2*>(v + q);
            double w0 = 0.0, w1 = 0.0;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const double2 z = *reinterpret_cast<const double2*>(sZ + (16 * sp + 8 * i + g) * RS + q);
              accZ[jj][i][j][0] = fma(wq.x, T[i][j][0], accZ[jj][i][j][0]);
              accZ[jj][i][j][1] = fma(wq.y, T[i][j][1], accZ[jj][i][j][1]);
              w0 = fma(z.x, T[i][j][0], w0);
              w1 = fma(z.y, T[i][j][1], w1);
            }
            wp[2 * j] = w0;
            wp[2 * j + 1] = w1;
          }
#pragma unroll
          for (int c0 = 0; c0 < 2 * QT; c0 += 8) {
            double v8[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) v8[c] = (c0 + c < 2 * QT) ? wp[(c0 + c < 2 * QT) ? c0 + c : 0] : 0.0;
            const double tot = reduce8_over_g(v8, lane);
            const int cc = c0 + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            if (cc < 2 * QT) sW[jb * Qp + 8 * (cc >> 1) + 2 * t + (cc & 1)] = tot;
          }
          ls0 += __shfl_xor_sync(0xffffffffu, ls0, 1);
          ls0 += __shfl_xor_sync(0xffffffffu, ls0, 2);
          ls1 += __shfl_xor_sync(0xffffffffu, ls1, 1);
          ls1 += __shfl_xor_sync(0xffffffffu, ls1, 2);
          if (t == 0) {
            sLam[kr * Mp16 + 16 * sp + g] = ls0;
            sLam[kr * Mp16 + 16 * sp + 8 + g] = ls1;
          }
        }
      }
      __syncthreads();                                // L_n consumed; partials of row n complete
      if (tid < Qp) {
        double s = 0.0;
        for (int jb = 0; jb < njobs; ++jb) s += sW[jb * Qp + tid];
        Wq[n * QC + tid] = s;
      } else if (tid >= 64 && tid < 64 + Mp16) {
        const int m = tid - 64;
        double s = 0.0;
        for (int kr = 0; kr < KS; ++kr) s += sLam[kr * Mp16 + m];
        lam[n * Mp + m] = s;
      }
    }
  }

  if constexpr (BWD) {
#pragma unroll
    for (int jj = 0; jj < JMAX; ++jj) {
      const int jb = wid + PS_WARPS * jj;
      if (jb < njobs) {
        const int kr = jb / Ms, sp = jb - kr * Ms;
        double* out = ACCp + (size_t)(blockIdx.x * KS + kr) * Mp * QC;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < QT; ++j)
            *reinterpret_cast<double2*>(out + (size_t)(16 * sp + 8 * i + g) * QC + 8 * j + 2 * t) =
                make_double2(accZ[jj][i][j][0], accZ[jj][i][j][1]);
      }
    }
  }
  if constexpr (FWD) {
    double* out = P2s + (size_t)blockIdx.x * Mp16 * Mp16;
#pragma unroll
    for (int s = 0; s < PS_S1; ++s)
      if (s < ns) {
        const bool offd = si[s] != sj[s];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int m = 16 * si[s] + 8 * i + g, mp = 16 * sj[s] + 8 * j + 2 * t;
            *reinterpret_cast<double2*>(out + m * Mp16 + mp) = make_double2(pacc[s][i][j][0], pacc[s][i][j][1]);
            if (offd) {
              out[mp * Mp16 + m] = pacc[s][i][j][0];
              out[(mp + 1) * Mp16 + m] = pacc[s][i][j][1];
            }
          }
      }
  }
}

// Psi2[m,m'] (+)= s2^2 * sum_r P2s[r][m][m']: one element per thread, partials summed in fixed order, eight loads
// in flight (as k_psi2_reduce)
__global__ void __launch_bounds__(256) k_psi2_reduce_small(int M, int Mp16, int R, double v2,
                                                          const double* __restrict__ P2s, int accumulate,
                                                          double* __restrict__ psi2) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= M * M) return;
  const int m = idx / M, mp = idx - m * M;
  const int64_t stride = (int64_t)Mp16 * Mp16;
  const double* p = P2s + m * Mp16 + mp;
  double s = 0.0;
  int k = 0;
  for (; k + 8 <= R; k += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = p[(int64_t)(k + u) * stride];
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; k < R; ++k) s += p[(int64_t)k * stride];
  s *= v2;
  if (accumulate) psi2[idx] += s;
  else psi2[idx] = s;
}

}  // namespace fast
}  // namespace rgp
