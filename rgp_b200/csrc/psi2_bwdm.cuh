// psi2_bwdm.cuh - k_psi2_bwd (8 warps, 16 x 32 sub-blocks) with the per-row CTA barrier replaced by
// split-phase synchronisation.
//
// In k_psi2_bwd every warp meets every other warp once per row, right after the exp epilogue, with
// no slack: the arrival spread (1.2-1.8 K cycles of a ~16.7 K-cycle row, profiles/SUMMARY_r01.md
// section 4) is pure idle time for the FP64 pipe.  The data dependencies are weaker than that:
//   stage 2-I of warp (wr, wc) reads rows 16 wr .. 16 wr + 15 of L  -> written by warps (wr, 0), (wr, 1)
//   stage 2-J reads columns 16 wr .. of ALL rows                    -> written by all eight warps
// so here
//   * the two warps of a row pair meet at a 64-thread named barrier (bar.sync 1 + wr) before 2-I,
//   * every warp ARRIVES on mbarrier FULL[slot] after storing its part of L and only WAITS for it
//     after stage 2-I, a third of a row later,
//   * the L slot (double buffered by row parity) is released through mbarrier FREE[slot]: arrive
//     after stage 2-J, wait before the store of row n + 2,
//   * row vectors are staged per warp with cp.async (no cooperative staging, no barrier),
//   * lambda / W_q partials are added to global memory by the warp that owns them (red.global.add)
//     instead of being combined through shared memory behind the barrier.
// Diagonal blocks keep the mirrored-tile scheme and wait for FULL right after the store.
// Arithmetic, tile shapes, stage1 / stage1_diag and the shared tile layout are those of k_psi2_bwd.
#pragma once
#include "psi2_kernels.cuh"
#include "psi2_bwds.cuh"   // mbarrier / cp.async helpers

namespace rgp {
namespace fast {

template <int QC>
struct P2CfgM {
  static constexpr int RS = QC + 4;
  static constexpr int QS = QC > 64 ? 64 : QC;
  static constexpr int NJ = QS / 16;
  static constexpr int VB = QC + 128;                      // ws[QC] | H_I[64] | H_J[64]
  static constexpr int ZI = 0;
  static constexpr int ZJ = 64 * RS;
  static constexpr int LL = 2 * 64 * RS;                   // L tile, 2 slots [2][64][RSL]
  static constexpr int VV = LL + 2 * 64 * RSL;             // row vectors [8 warps][2][VB]
  static constexpr int TT = VV + 8 * 2 * VB;               // exp table [256]
  static constexpr int MB = TT + 256;                      // mbarriers FULL[2], FREE[2]
  static constexpr int SMEM = (MB + 4) * 8;
};

RGP_DEVINL void pair_barrier(int id) { asm volatile("bar.sync %0, 64;\n" ::"r"(id) : "memory"); }

template <int QC>
__global__ void __launch_bounds__(P2_THREADS, 1)
k_psi2_bwdm(int64_t rc, int Mp, int nt, int nblocks, int qk, const double* __restrict__ Zt,
            const double* __restrict__ Ct, const double* __restrict__ wrow,
            const double* __restrict__ HP, double* __restrict__ lam, double* __restrict__ Wq,
            double* __restrict__ ACCp, int qoff, int dbg) {
  // dbg (timing experiments only, results wrong): 1 = no lambda atomics, 2 = no W atomics
  using C = P2CfgM<QC>;
  constexpr int RS = C::RS, VB = C::VB, NJ = C::NJ, QS = C::QS;
  extern __shared__ __align__(16) double smem[];
  double* sZI = smem + C::ZI;
  double* sZJ = smem + C::ZJ;
  double* sL = smem + C::LL;
  double* sT = smem + C::TT;
  const unsigned sb = smem_addr(smem);
  const unsigned mb = sb + 8u * C::MB;

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
  double* myV = smem + C::VV + wid * 2 * VB;
  const unsigned myVs = sb + 8u * (C::VV + wid * 2 * VB);
  exp_table_init(sT, tid);
  if (tid < 4) mbar_init(mb + 8u * tid, 8);
  unsigned it = 0;                                // rows processed by this CTA over all blocks; slot = it & 1
  const int c8 = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);

  int curI = -1, curJ = -1;
  for (int b = blockIdx.y; b < nblocks; b += G) {
    int I, J;
    block_ij(b, nt, I, J);
    const bool diag = (I == J);
    __syncthreads();
    if (I != curI) copy_tile<64 * RS>(sZI, Zt + (size_t)I * 64 * RS, tid);
    if (J != curJ) copy_tile<64 * RS>(sZJ, Zt + (size_t)J * 64 * RS, tid);
    curI = I;
    curJ = J;
    const double* hI = HP + (size_t)I * rc * 64;
    const double* hJ = HP + (size_t)J * rc * 64;
    const double* cb = Ct + (size_t)b * 4096;
    double accI[2][NJ][2], accJ[2][NJ][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) accI[i][j][0] = accI[i][j][1] = accJ[i][j][0] = accJ[i][j][1] = 0.0;
    __syncthreads();

    // this warp's copy of the row vectors of row n -> slot n & 1 (16-byte cp.async chunks)
    auto stage = [&](int64_t n) {
      const unsigned dst = myVs + 8u * ((unsigned)(n & 1) * VB);
      for (int c = lane; c < VB / 2; c += 32) {
        const double* src = c < QC / 2 ? wrow + n * QC + 2 * c
                            : (c < QC / 2 + 32 ? hI + n * 64 + 2 * (c - QC / 2) : hJ + n * 64 + 2 * (c - QC / 2 - 32));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst + 16u * c), "l"(src) : "memory");
      }
      cp_async_commit();
    };
    auto wait_vectors = [&](int64_t n) {
      if (n + 1 < r1) {
        stage(n + 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
    };

    // stage 2-I: T = L ZJ ; accI += ws T ; W partial straight to global
    auto stage2I = [&](const double* v, const double* Lb, double* wq_row) {
      double T[2][NJ][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) T[i][j][0] = T[i][j][1] = 0.0;
      const double* pa = Lb + (16 * wr + g) * RSL + t;
      const double* pb = sZJ + t * RS + qoff + qbase + g;
#pragma unroll 2
      for (int k0 = 0; k0 < 64; k0 += 4) {
        double a[2], bq[NJ];
#pragma unroll
        for (int i = 0; i < 2; ++i) a[i] = pa[i * 8 * RSL + k0];
#pragma unroll
        for (int j = 0; j < NJ; ++j) bq[j] = pb[k0 * RS + 8 * j];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) dmma(T[i][j][0], T[i][j][1], a[i], bq[j]);
      }
      double wp[2 * NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int q = qoff + qbase + 8 * j + 2 * t;
        const double2 wq = *reinterpret_cast<const double2*>(v + q);
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
        if (!(dbg & 2)) red_add(wq_row + qoff + qbase + 8 * (c8 >> 1) + 2 * t + (c8 & 1), diag ? tot : 2.0 * tot);
      } else {
#pragma unroll
        for (int c = 0; c < 2 * NJ; ++c) {
          double x = wp[c];
          x += __shfl_xor_sync(0xffffffffu, x, 4);
          x += __shfl_xor_sync(0xffffffffu, x, 8);
          x += __shfl_xor_sync(0xffffffffu, x, 16);
          if (g == 0) red_add(wq_row + qoff + qbase + 8 * (c >> 1) + 2 * t + (c & 1), diag ? x : 2.0 * x);
        }
      }
    };

    if (r0 < r1) stage(r0);
    if (!diag) {
      // ------------------------------------------------------------ off-diagonal block
      double creg[2][4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double2 c2 = *reinterpret_cast<const double2*>(cb + (16 * wr + 8 * i + g) * 64 + 32 * wc + 8 * j + 2 * t);
          creg[i][j][0] = c2.x;
          creg[i][j][1] = c2.y;
        }
      for (int64_t n = r0; n < r1; ++n, ++it) {
        wait_vectors(n);
        const int s = (int)(it & 1u);
        const unsigned use = it >> 1;
        const double* v = myV + (int)(n & 1) * VB;
        double* Lb = sL + s * 64 * RSL;
        {
          double acc[2][4][2];
          stage1<QC>(sZI, sZJ, v, qk, wr, wc, lane, acc);
          double rs[2] = {0.0, 0.0};
          double cs[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) cs[c] = 0.0;
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[i][j][0] = creg[i][j][0] * exp_tab(acc[i][j][0], sT);
              acc[i][j][1] = creg[i][j][1] * exp_tab(acc[i][j][1], sT);
              rs[i] += acc[i][j][0] + acc[i][j][1];
              cs[2 * j] += acc[i][j][0];
              cs[2 * j + 1] += acc[i][j][1];
            }
          if (use > 0) mbar_wait(mb + 16u + 8u * s, (use - 1u) & 1u);   // FREE: 2-J of row n - 2 has left the slot
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<double2*>(Lb + (16 * wr + 8 * i + g) * RSL + 32 * wc + 8 * j + 2 * t) =
                  make_double2(acc[i][j][0], acc[i][j][1]);
          __syncwarp();
          if (lane == 0) mbar_arrive(mb + 8u * s);                      // FULL: this warp's part of L is stored
          if (qoff == 0 && !(dbg & 1)) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 1);
              rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 2);
              if (t == 0) red_add(lamg + n * Mp + I * 64 + 16 * wr + 8 * i + g, rs[i]);
            }
            const double tot = reduce8_over_g(cs, lane);
            red_add(lamg + n * Mp + J * 64 + 32 * wc + 8 * (c8 >> 1) + 2 * t + (c8 & 1), tot);
          }
        }
        pair_barrier(1 + wr);                     // rows 16 wr .. of L are complete (both column halves)
        stage2I(v, Lb, Wqg + n * QC);
        mbar_wait(mb + 8u * s, use & 1u);         // FULL: all of L is stored
        // stage 2-J: accJ[m',q] += sum_m L[m,m'] (ws_q ZI[m,q]).  A = L^T, B = ws * ZI
        {
          const double* pa = Lb + t * RSL + 16 * wr + g;
          const double* pb = sZI + t * RS + qoff + qbase + g;
          double wq[NJ];
#pragma unroll
          for (int j = 0; j < NJ; ++j) wq[j] = v[qoff + qbase + 8 * j + g];
#pragma unroll 2
          for (int k0 = 0; k0 < 64; k0 += 4) {
            double a[2], bq[NJ];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = pa[k0 * RSL + 8 * i];
#pragma unroll
            for (int j = 0; j < NJ; ++j) bq[j] = pb[k0 * RS + 8 * j] * wq[j];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
              for (int j = 0; j < NJ; ++j) dmma(accJ[i][j][0], accJ[i][j][1], a[i], bq[j]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(mb + 16u + 8u * s);                  // FREE
      }
    } else {
      // ---------------------------------------------------------------- diagonal block
      int ti[5], tj[5], cnt;
      diag_tiles(wid, ti, tj, cnt);
      double creg[5][2];
#pragma unroll
      for (int s5 = 0; s5 < 5; ++s5) {
        double2 c2 = *reinterpret_cast<const double2*>(cb + (8 * ti[s5] + g) * 64 + 8 * tj[s5] + 2 * t);
        creg[s5][0] = c2.x;
        creg[s5][1] = c2.y;
      }
      for (int64_t n = r0; n < r1; ++n, ++it) {
        wait_vectors(n);
        const int s = (int)(it & 1u);
        const unsigned use = it >> 1;
        const double* v = myV + (int)(n & 1) * VB;
        double* Lb = sL + s * 64 * RSL;
        {
          double acc[5][2];
          stage1_diag<QC>(sZI, v, qk, ti, tj, cnt, lane, acc);
#pragma unroll
          for (int s5 = 0; s5 < 5; ++s5)
            if (s5 < cnt) {
              acc[s5][0] = creg[s5][0] * exp_tab(acc[s5][0], sT);
              acc[s5][1] = creg[s5][1] * exp_tab(acc[s5][1], sT);
            }
          if (use > 0) mbar_wait(mb + 16u + 8u * s, (use - 1u) & 1u);   // FREE
#pragma unroll
          for (int s5 = 0; s5 < 5; ++s5)
            if (s5 < cnt) {
              const int m = 8 * ti[s5] + g, mp = 8 * tj[s5] + 2 * t;
              *reinterpret_cast<double2*>(Lb + m * RSL + mp) = make_double2(acc[s5][0], acc[s5][1]);
              if (ti[s5] != tj[s5]) {
                Lb[mp * RSL + m] = acc[s5][0];
                Lb[(mp + 1) * RSL + m] = acc[s5][1];
              }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(mb + 8u * s);
        mbar_wait(mb + 8u * s, use & 1u);         // the mirrored tile is needed whole
        if (qoff == 0) {                          // lambda_m = full row sum; warp w owns rows 8 w .. 8 w + 7
          const double2* row = reinterpret_cast<const double2*>(Lb + (8 * wid + g) * RSL + 16 * t);
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const double2 x = row[k];
            s0 += x.x;
            s1 += x.y;
          }
          s0 += s1;
          s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
          s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
          if (t == 0) red_add(lamg + n * Mp + I * 64 + 8 * wid + g, s0);
        }
        stage2I(v, Lb, Wqg + n * QC);
        __syncwarp();
        if (lane == 0) mbar_arrive(mb + 16u + 8u * s);                  // FREE
      }
    }
    // flush the CTA-private dZ accumulators of this block (each element is owned by one thread)
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
