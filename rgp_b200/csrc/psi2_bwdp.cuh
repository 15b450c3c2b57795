// psi2_bwdp.cuh - the Psi2 backward kernel, software-pipelined over rows.
//
// Same decomposition, tiles and outputs as the row-at-a-time kernel it replaces (see the header of
// psi2_kernels.cuh): a CTA owns a row range and walks the 64 x 64 blocks of the pair matrix; per row
//   stage 1  E = H_m + H_m' + (ws Z'_I) Z'_J^T        epilogue  L = C * exp(E) -> shared, lambda sums
//   stage 2-I  T = L Z'_J, accI += ws T, W += Z'_I.T   stage 2-J  accJ += L^T (ws Z'_I)
//
// What changed, and why (profiles/SUMMARY_r01.md section 4 measured where the old kernel lost time):
// the scalar FP64 work of a row (4096 table-assisted exps, the C multiply, the lambda / W sums: ~330
// FP64-pipe instructions per thread) ran as separate phases BETWEEN the three DMMA loops.  DMMA and
// scalar FP64 share one pipe; a warp in a scalar phase next to a warp in a DMMA loop is starved, and
// two warps in scalar phases leave the pipe half empty (issue-bound), so ~14 % of the row time was
// scalar work and another ~14 % imperfect overlap.  Here every warp's instruction stream carries both
// at once: iteration n of the row loop runs
//        stage 1 of row n+1                 (DMMA)
//        stage 2-I of row n                 (DMMA)   + W / accI folds
//        stage 2-J of row n                 (DMMA)   interleaved, k-step by k-step, with the exp /
//                                                    L / lambda epilogue of row n+1
// so a DMMA is ready on every scheduler at every moment and the scalar instructions ride in the
// gaps of the same in-order stream (they cannot be starved: the warp's next DMMA is behind them).
// One CTA barrier per row, as before: L(n+1) is written into the other L buffer while L(n) is read.
//
// Row vectors by TMA.  The per-row vectors (ws[QC], H_I[64], H_J[64]) of VR consecutive rows are
// contiguous in HBM; one elected thread fetches them with three cp.async.bulk copies per batch into
// a double-buffered shared slot and an mbarrier carries the byte count (UBLKCP / SYNCS in SASS).  No
// compute thread touches global memory for operands inside the row loop any more.
#pragma once
#include <type_traits>

#include "psi2_kernels.cuh"

namespace rgp {
namespace fast {

// NJ = 8-wide q tiles per warp in stage 2: the stage-2 outputs cover QS = 16 NJ columns per pass (the two
// warps of a column pair take 8 NJ each), so Q = 40 runs NJ = 3 (48 columns) instead of padding to 64
// (measured 17 % faster at M = 200, Q = 40, profiles/kernel_times_small_r02.jsonl).
template <int QC, int NJ_>
struct P2CfgP {
  static constexpr int RS = QC + 4;     // this kernel keeps the pad-4 Z' tiles (the LDS.128 pairing measured 1-4 % slower here)
  static constexpr int NJ = NJ_;
  static constexpr int QS = 16 * NJ_;
  static_assert(QS <= (QC > 64 ? 64 : QC), "stage-2 width exceeds the tile");
  static constexpr int VR = QC > 64 ? 2 : 8;                 // rows per TMA batch
  static constexpr int VBB = VR * (QC + 128);                // doubles per batch slot: ws | H_I | H_J
  static constexpr int SMEM_D = 2 * 64 * RS + 2 * 64 * RSL + 2 * VBB + 2 * 4 * QS + 2 * 2 * 64 + 2 * 4 * 64 + 256 + 2;
  static constexpr int SMEM = SMEM_D * 8;
  static constexpr int FUSED_SMEM = SMEM + 64 * RSL * 8;
};

// stage 1 with the three row vectors given separately (they live in different parts of a batch slot)
template <int QC>
RGP_DEVINL void stage1v(const double* __restrict__ sZI, const double* __restrict__ sZJ,
                        const double* __restrict__ sw, const double* __restrict__ hI,
                        const double* __restrict__ hJ, int qk, int wr, int wc, int lane,
                        double (&acc)[2][4][2]) {
  constexpr int RS = QC + 4;
  const int g = lane >> 2, t = lane & 3;
  const double* pa = sZI + (16 * wr + g) * RS + t;
  const double* pb = sZJ + (32 * wc + g) * RS + t;
  const double* vI = hI + 16 * wr + g;
  const double* vJ = hJ + 32 * wc + 2 * t;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double hi = vI[8 * i];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double2 hj = *reinterpret_cast<const double2*>(vJ + 8 * j);
      acc[i][j][0] = hi + hj.x;
      acc[i][j][1] = hi + hj.y;
    }
  }
#pragma unroll 2
  for (int k0 = 0; k0 < qk; k0 += 4) {
    const double wv = sw[k0 + t];
    double a[2], b[4];
#pragma unroll
    for (int i = 0; i < 2; ++i) a[i] = pa[i * 8 * RS + k0] * wv;
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = pb[j * 8 * RS + k0];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

template <int QC, int CNT>
RGP_DEVINL void stage1v_diag_n(const double* __restrict__ sZ, const double* __restrict__ sw,
                               const double* __restrict__ hI, int qk, const int (&ti)[5], const int (&tj)[5],
                               int lane, double (&acc)[5][2]) {
  constexpr int RS = QC + 4;
  const int g = lane >> 2, t = lane & 3;
  const double* pa[CNT];
  const double* pb[CNT];
#pragma unroll
  for (int s = 0; s < CNT; ++s) {
    pa[s] = sZ + (8 * ti[s] + g) * RS + t;
    pb[s] = sZ + (8 * tj[s] + g) * RS + t;
    const double hi = hI[8 * ti[s] + g];
    const double2 hj = *reinterpret_cast<const double2*>(hI + 8 * tj[s] + 2 * t);
    acc[s][0] = hi + hj.x;
    acc[s][1] = hi + hj.y;
  }
#pragma unroll 2
  for (int k0 = 0; k0 < qk; k0 += 4) {
    const double wv = sw[k0 + t];
    double a[CNT], b[CNT];
#pragma unroll
    for (int s = 0; s < CNT; ++s) {
      a[s] = pa[s][k0] * wv;
      b[s] = pb[s][k0];
    }
#pragma unroll
    for (int s = 0; s < CNT; ++s) dmma(acc[s][0], acc[s][1], a[s], b[s]);
  }
}

template <int QC>
RGP_DEVINL void stage1v_diag(const double* __restrict__ sZ, const double* __restrict__ sw,
                             const double* __restrict__ hI, int qk, const int (&ti)[5], const int (&tj)[5], int cnt,
                             int lane, double (&acc)[5][2]) {
  if (cnt == 5) stage1v_diag_n<QC, 5>(sZ, sw, hI, qk, ti, tj, lane, acc);
  else stage1v_diag_n<QC, 4>(sZ, sw, hI, qk, ti, tj, lane, acc);
}

// =====================================================================================
// grid = (R row ranges, G block groups); outputs as k_psi2_bwd:
//   lam [g][rc][Mp], Wq [g][rc][QC] (red.global.add into rows only this CTA touches),
//   ACCp[cta][Mp][QC] (CTA-private), FUSE: P2p[b][r][64][64] partial Psi2 tiles.
// =====================================================================================
template <int QC, int NJ_, bool FUSE = false>
__global__ void __launch_bounds__(P2_THREADS, 1)
k_psi2_bwdp(int64_t rc, int Mp, int nt, int nblocks, int qk, const double* __restrict__ Zt,
            const double* __restrict__ Ct, const double* __restrict__ wrow, const double* __restrict__ HP,
            double* __restrict__ lam, double* __restrict__ Wq, double* __restrict__ ACCp, int qoff,
            double* __restrict__ P2p = nullptr) {
  using C = P2CfgP<QC, NJ_>;
  constexpr int RS = C::RS, NJ = C::NJ, QS = C::QS, VR = C::VR, VBB = C::VBB;
  extern __shared__ __align__(16) double smem[];
  double* sZI = smem;
  double* sZJ = sZI + 64 * RS;
  double* sL = sZJ + 64 * RS;                     // 2 slots of 64*RSL
  double* sVb = sL + 2 * 64 * RSL;                // 2 batch slots of VBB
  double* sWq = sVb + 2 * VBB;                    // [2][4][QS]
  double* sLr = sWq + 2 * 4 * QS;                 // [2][2][64]  row-sum partials (per wc)
  double* sLc = sLr + 2 * 2 * 64;                 // [2][4][64]  col-sum partials (per wr)
  double* sT = sLc + 2 * 4 * 64;                  // exp table, 256 entries
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sT + 256);   // one per batch slot
  double* sP = sT + 256 + 2;                      // FUSE: Psi2 tile of the block [64][RSL]

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int wr = wid >> 1, wc = wid & 1, g = lane >> 2, t = lane & 3;
  const int R = gridDim.x, G = gridDim.y;
  const int64_t per = (rc + R - 1) / R;
  const int64_t r0 = per * blockIdx.x, r1 = (r0 + per < rc) ? r0 + per : rc;
  const int cta = blockIdx.y * R + blockIdx.x;
  double* lamg = lam + (size_t)blockIdx.y * rc * Mp;
  double* Wqg = Wq + (size_t)blockIdx.y * rc * QC;
  double* accp = ACCp + (size_t)cta * Mp * QC;
  const int qbase = wc * (QS / 2);
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
  }
  exp_table_init(sT, tid);
  uint32_t phase_bits = 0u;                        // bit s: parity of the next completion of slot s's mbarrier
  if (r0 >= r1) {                                  // (uniform) an empty row range adds nothing, but the Psi2 reduction
    if constexpr (FUSE)                            // sums one partial tile per row range: write zeros
      for (int b = blockIdx.y; b < nblocks; b += G) {
        double* out = P2p + ((size_t)b * R + blockIdx.x) * 4096;
        for (int i = tid; i < 4096; i += P2_THREADS) out[i] = 0.0;
      }
    return;
  }

  int curI = -1, curJ = -1;
  for (int b = blockIdx.y; b < nblocks; b += G) {
    int I, J;
    block_ij(b, nt, I, J);
    const bool diag = (I == J);
    __syncthreads();                              // previous block done with every shared buffer
    if (I != curI) copy_tile<64 * RS>(sZI, Zt + (size_t)I * 64 * RS, tid);
    if (J != curJ) copy_tile<64 * RS>(sZJ, Zt + (size_t)J * 64 * RS, tid);
    curI = I;
    curJ = J;
    const double* hI = HP + (size_t)I * rc * 64;
    const double* hJ = HP + (size_t)J * rc * 64;
    const double* cb = Ct + (size_t)b * 4096;
    // batch k of this block = rows [r0 + k VR, ...) -> slot k & 1
    auto issue = [&](int64_t k) {
      const int64_t n0 = r0 + k * VR;
      if (n0 >= r1) return;
      const int rows = (int)((r1 - n0 < VR) ? r1 - n0 : VR);
      const int slot = (int)(k & 1);
      double* dst = sVb + slot * VBB;
      mbar_expect_tx(&mbar[slot], (uint32_t)(rows * (QC + 128) * 8));
      bulk_g2s(dst, wrow + n0 * QC, (uint32_t)(rows * QC * 8), &mbar[slot]);
      bulk_g2s(dst + VR * QC, hI + n0 * 64, (uint32_t)(rows * 512), &mbar[slot]);
      bulk_g2s(dst + VR * QC + VR * 64, hJ + n0 * 64, (uint32_t)(rows * 512), &mbar[slot]);
    };
    auto await = [&](int64_t k) {
      const int slot = (int)(k & 1);
      mbar_wait(&mbar[slot], (phase_bits >> slot) & 1u);
      phase_bits ^= 1u << slot;
    };
    if (tid == 0) {
      issue(0);
      issue(1);
    }
    if constexpr (FUSE)
      for (int i = tid; i < 64 * RSL; i += P2_THREADS) sP[i] = 0.0;
    double accI[2][NJ][2], accJ[2][NJ][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) accI[i][j][0] = accI[i][j][1] = accJ[i][j][0] = accJ[i][j][1] = 0.0;
    __syncthreads();                              // Z' tiles (and the cleared Psi2 tile) visible
    await(0);

    // row n -> its three vectors inside the batch slots
    auto vec = [&](int64_t n, const double*& sw, const double*& vI, const double*& vJ) {
      const int idx = (int)(n - r0);
      const double* base = sVb + ((idx / VR) & 1) * VBB;
      const int w = idx % VR;
      sw = base + w * QC;
      vI = base + VR * QC + w * 64;
      vJ = base + VR * QC + VR * 64 + w * 64;
    };
    // start of iteration n: refill the slot the previous batch has just left, and make sure the batch
    // of row n+1 has landed (all threads poll; it was issued >= VR-1 rows ago)
    auto batches = [&](int64_t n) {
      const int idx = (int)(n - r0);
      if (idx > 0 && idx % VR == 0 && tid == 0) issue(idx / VR + 1);
      if (n + 1 < r1 && (idx + 1) % VR == 0) await((idx + 1) / VR);
    };
    auto flush_wq = [&](int64_t n) {
      if (tid < QS) {
        const double* p = sWq + (int)(n & 1) * 4 * QS + tid;
        const double v = p[0] + p[QS] + p[2 * QS] + p[3 * QS];
        red_add(Wqg + n * QC + qoff + tid, diag ? v : 2.0 * v);
      }
    };

    // stage 2-I of row n (T = L Z'_J) followed by its folds: accI += ws T, W partial -> sWq slot.
    // ED: interleave the epilogue of the NEXT row of a diagonal block (see below) with the k-steps.
    auto stage2I = [&](const double* __restrict__ sw, const double* __restrict__ Lr, int s, auto&& between) {
      double T[2][NJ][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) T[i][j][0] = T[i][j][1] = 0.0;
      const double* pa = Lr + (16 * wr + g) * RSL + t;
      const double* pb = sZJ + t * RS + qoff + qbase + g;
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) {
        const int k0 = 4 * ks;
        double a[2], bq[NJ];
#pragma unroll
        for (int i = 0; i < 2; ++i) a[i] = pa[i * 8 * RSL + k0];
#pragma unroll
        for (int j = 0; j < NJ; ++j) bq[j] = pb[k0 * RS + 8 * j];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) dmma(T[i][j][0], T[i][j][1], a[i], bq[j]);
        between(ks);
      }
      double wp[2 * NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int q = qoff + qbase + 8 * j + 2 * t;
        const double2 wq = *reinterpret_cast<const double2*>(sw + q);
        double w0 = 0.0, w1 = 0.0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const double2 z = *reinterpret_cast<const double2*>(sZI + (16 * wr + 8 * i + g) * RS + q);
          accI[i][j][0] = fma(wq.x, T[i][j][0], accI[i][j][0]);
          accI[i][j][1] = fma(wq.y, T[i][j][1], accI[i][j][1]);
          w0 = fma(z.x, T[i][j][0], w0);
          w1 = fma(z.y, T[i][j][1], w1);
        }
        wp[2 * j] = w0;
        wp[2 * j + 1] = w1;
      }
      if constexpr (NJ == 4) {
        const double tot = reduce8_over_g(wp, lane);
        const int c = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        sWq[s * 4 * QS + wr * QS + qbase + 8 * (c >> 1) + 2 * t + (c & 1)] = tot;
      } else {
#pragma unroll
        for (int c = 0; c < 2 * NJ; ++c) {
          double x = wp[c];
          x += __shfl_xor_sync(0xffffffffu, x, 4);
          x += __shfl_xor_sync(0xffffffffu, x, 8);
          x += __shfl_xor_sync(0xffffffffu, x, 16);
          if (g == 0) sWq[s * 4 * QS + wr * QS + qbase + 8 * (c >> 1) + 2 * t + (c & 1)] = x;
        }
      }
    };
    auto nothing = [](int) {};

    if (!diag) {
      // ------------------------------------------------------------ off-diagonal block
      double creg[2][4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double2 c2 = *reinterpret_cast<const double2*>(cb + (16 * wr + 8 * i + g) * 64 + 32 * wc + 8 * j + 2 * t);
          creg[i][j][0] = c2.x;
          creg[i][j][1] = c2.y;
        }
      double accN[2][4][2];                        // exponents of the next row (stage 1 output)
      double rs[2], cs[8];
      // epilogue of one (i, j) accumulator pair of the row held in accN: exp, Psi2 side sum, L, lambda partials
      auto e_pair = [&](int i, int j, double* __restrict__ Lw) {
        const double p0 = exp_tab(accN[i][j][0], sT), p1 = exp_tab(accN[i][j][1], sT);
        const int off = (16 * wr + 8 * i + g) * RSL + 32 * wc + 8 * j + 2 * t;
        if constexpr (FUSE) {
          double2* pp = reinterpret_cast<double2*>(sP + off);
          double2 o = *pp;
          o.x += p0;
          o.y += p1;
          *pp = o;
        }
        const double l0 = creg[i][j][0] * p0, l1 = creg[i][j][1] * p1;
        *reinterpret_cast<double2*>(Lw + off) = make_double2(l0, l1);
        rs[i] += l0 + l1;
        cs[2 * j] += l0;
        cs[2 * j + 1] += l1;
      };
      auto e_begin = [&]() {
        rs[0] = rs[1] = 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c) cs[c] = 0.0;
      };
      auto e_end = [&](int s) {                    // lambda partials of the row -> slot s
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 1);
          rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 2);
        }
        if (t == 0) {
          sLr[s * 128 + wc * 64 + 16 * wr + g] = rs[0];
          sLr[s * 128 + wc * 64 + 16 * wr + 8 + g] = rs[1];
        }
        const double tot = reduce8_over_g(cs, lane);
        const int c = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        sLc[s * 256 + wr * 64 + 32 * wc + 8 * (c >> 1) + 2 * t + (c & 1)] = tot;
      };
      // stage 2-J of row n: accJ += L^T (ws Z'_I); WITH_E interleaves the epilogue of row n+1
      auto stage2J = [&](const double* __restrict__ sw, const double* __restrict__ Lr, double* __restrict__ Lw,
                         int snext, auto WITH_E) {
        constexpr bool E = decltype(WITH_E)::value;
        const double* pa = Lr + t * RSL + 16 * wr + g;
        const double* pb = sZI + t * RS + qoff + qbase + g;
        double wq[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) wq[j] = sw[qoff + qbase + 8 * j + g];
        if constexpr (E) e_begin();
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          const int k0 = 4 * ks;
          double a[2], bq[NJ];
#pragma unroll
          for (int i = 0; i < 2; ++i) a[i] = pa[k0 * RSL + 8 * i];
#pragma unroll
          for (int j = 0; j < NJ; ++j) bq[j] = pb[k0 * RS + 8 * j] * wq[j];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) dmma(accJ[i][j][0], accJ[i][j][1], a[i], bq[j]);
          if constexpr (E) {
            // 8 accumulator pairs over k-steps 0,1, 3,4, 6,7, 9,10; the lambda reductions follow at 11,
            // so their shuffle chains still have four k-steps of DMMAs behind them
            if (ks % 3 != 2 && ks < 11) {
              const int p = (ks / 3) * 2 + (ks % 3);
              e_pair(p >> 2, p & 3, Lw);
            }
            if (ks == 11) e_end(snext);
          }
        }
      };

      {                                            // prologue: exponents and L of the first row
        const double *sw, *vI, *vJ;
        vec(r0, sw, vI, vJ);
        stage1v<QC>(sZI, sZJ, sw, vI, vJ, qk, wr, wc, lane, accN);
        double* Lw = sL + (int)(r0 & 1) * 64 * RSL;
        e_begin();
#pragma unroll
        for (int p = 0; p < 8; ++p) e_pair(p >> 2, p & 3, Lw);
        e_end((int)(r0 & 1));
      }
      __syncthreads();
      for (int64_t n = r0; n < r1; ++n) {
        const int s = (int)(n & 1);
        const double* Lr = sL + s * 64 * RSL;
        double* Lw = sL + (s ^ 1) * 64 * RSL;
        batches(n);
        if (qoff == 0) {                           // lambda of row n (partials written one iteration ago)
          if (tid >= 64 && tid < 128) {
            const int m = tid - 64;
            const double* p = sLr + s * 128 + m;
            red_add(lamg + n * Mp + I * 64 + m, p[0] + p[64]);
          } else if (tid >= 128 && tid < 192) {
            const int m = tid - 128;
            const double* p = sLc + s * 256 + m;
            red_add(lamg + n * Mp + J * 64 + m, p[0] + p[64] + p[128] + p[192]);
          }
        }
        if (n > r0) flush_wq(n - 1);
        const double *sw, *vI, *vJ;
        vec(n, sw, vI, vJ);
        if (n + 1 < r1) {
          const double *sw1, *vI1, *vJ1;
          vec(n + 1, sw1, vI1, vJ1);
          stage1v<QC>(sZI, sZJ, sw1, vI1, vJ1, qk, wr, wc, lane, accN);
          stage2I(sw, Lr, s, nothing);
          stage2J(sw, Lr, Lw, s ^ 1, std::true_type{});
        } else {
          stage2I(sw, Lr, s, nothing);
          stage2J(sw, Lr, Lw, s ^ 1, std::false_type{});
        }
        __syncthreads();   // L(n+1) + its lambda partials + W(n) complete; every read of L(n) done
      }
    } else {
      // ---------------------------------------------------------------- diagonal block
      int ti[5], tj[5], cnt;
      diag_tiles(wid, ti, tj, cnt);
      double creg[5][2];
#pragma unroll
      for (int s5 = 0; s5 < 5; ++s5) {
        const double2 c2 = *reinterpret_cast<const double2*>(cb + (8 * ti[s5] + g) * 64 + 8 * tj[s5] + 2 * t);
        creg[s5][0] = c2.x;
        creg[s5][1] = c2.y;
      }
      double accN[5][2];
      auto e_tile = [&](int s5, double* __restrict__ Lw) {
        if (s5 < cnt) {
          const double p0 = exp_tab(accN[s5][0], sT), p1 = exp_tab(accN[s5][1], sT);
          const double l0 = creg[s5][0] * p0, l1 = creg[s5][1] * p1;
          const int m = 8 * ti[s5] + g, mp = 8 * tj[s5] + 2 * t;
          if constexpr (FUSE) {
            double2* pp = reinterpret_cast<double2*>(sP + m * RSL + mp);
            double2 o = *pp;
            o.x += p0;
            o.y += p1;
            *pp = o;
          }
          *reinterpret_cast<double2*>(Lw + m * RSL + mp) = make_double2(l0, l1);
          if (ti[s5] != tj[s5]) {
            Lw[mp * RSL + m] = l0;
            Lw[(mp + 1) * RSL + m] = l1;
          }
        }
      };
      {
        const double *sw, *vI, *vJ;
        vec(r0, sw, vI, vJ);
        stage1v_diag<QC>(sZI, sw, vI, qk, ti, tj, cnt, lane, accN);
        double* Lw = sL + (int)(r0 & 1) * 64 * RSL;
#pragma unroll
        for (int s5 = 0; s5 < 5; ++s5) e_tile(s5, Lw);
      }
      __syncthreads();
      for (int64_t n = r0; n < r1; ++n) {
        const int s = (int)(n & 1);
        const double* Lr = sL + s * 64 * RSL;
        double* Lw = sL + (s ^ 1) * 64 * RSL;
        batches(n);
        if (qoff == 0 && tid >= 64 && tid < 128) { // lambda_m = full row sum of the symmetric tile
          const int m = tid - 64;
          const double2* row = reinterpret_cast<const double2*>(Lr + m * RSL);
          double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
          for (int k = 0; k < 32; ++k) {
            const double2 x = row[k];
            s0 += x.x;
            s1 += x.y;
          }
          red_add(lamg + n * Mp + I * 64 + m, s0 + s1);
        }
        if (n > r0) flush_wq(n - 1);
        const double *sw, *vI, *vJ;
        vec(n, sw, vI, vJ);
        if (n + 1 < r1) {
          const double *sw1, *vI1, *vJ1;
          vec(n + 1, sw1, vI1, vJ1);
          stage1v_diag<QC>(sZI, sw1, vI1, qk, ti, tj, cnt, lane, accN);
          stage2I(sw, Lr, s, [&](int ks) {
            if (ks % 3 == 0 && ks < 15) e_tile(ks / 3, Lw);      // 5 tiles over k-steps 0, 3, 6, 9, 12
          });
        } else {
          stage2I(sw, Lr, s, nothing);
        }
        __syncthreads();
      }
    }
    flush_wq(r1 - 1);
    if constexpr (FUSE) {
      double* out = P2p + ((size_t)b * R + blockIdx.x) * 4096;
      if (!diag) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int m = 16 * wr + 8 * i + g, mp = 32 * wc + 8 * j + 2 * t;
            *reinterpret_cast<double2*>(out + m * 64 + mp) = *reinterpret_cast<const double2*>(sP + m * RSL + mp);
          }
      } else {
        int ti[5], tj[5], cnt;
        diag_tiles(wid, ti, tj, cnt);
#pragma unroll
        for (int s5 = 0; s5 < 5; ++s5)
          if (s5 < cnt) {
            const int m = 8 * ti[s5] + g, mp = 8 * tj[s5] + 2 * t;
            const double2 x = *reinterpret_cast<const double2*>(sP + m * RSL + mp);
            *reinterpret_cast<double2*>(out + m * 64 + mp) = x;
            if (ti[s5] != tj[s5]) {
              out[mp * 64 + m] = x.x;
              out[(mp + 1) * 64 + m] = x.y;
            }
          }
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int q = qoff + qbase + 8 * j + 2 * t;
        double2* pI = reinterpret_cast<double2*>(accp + (size_t)(I * 64 + 16 * wr + 8 * i + g) * QC + q);
        double2 o = *pI;
        o.x += accI[i][j][0];
        o.y += accI[i][j][1];
        *pI = o;
        if (!diag) {
          double2* pJ = reinterpret_cast<double2*>(accp + (size_t)(J * 64 + 16 * wr + 8 * i + g) * QC + q);
          double2 u = *pJ;
          u.x += accJ[i][j][0];
          u.y += accJ[i][j][1];
          *pJ = u;
        }
      }
  }
}

}  // namespace fast
}  // namespace rgp
