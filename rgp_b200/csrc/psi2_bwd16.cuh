// psi2_bwd16.cuh - 16-warp variant of the Psi2 backward kernel (512 threads, <= 128
// registers per thread, 1 CTA/SM).  Same algorithm, tiles, shared-memory layout and hazard
// analysis as k_psi2_bwd in psi2_kernels.cuh; every warp owns half as many MMA tiles
// (16 x 16 pairs in stage 1, 16 x QC/4 outputs in stage 2), which doubles the warps per
// scheduler (4 instead of 2) and hides the LDS / shuffle / barrier phases that ncu showed as
// `short_scoreboard` + idle FP64 pipe in the 8-warp kernel (profiles/).  QC in {32, 64}.
#pragma once
#include "psi2_kernels.cuh"

namespace rgp {
namespace fast {

constexpr int P2_THREADS16 = 512;

template <int QC>
struct P2Cfg16 {
  static constexpr int RS = QC + 4;
  static constexpr int NJ = QC / 32;        // 8-wide q tiles per warp in stage 2
  static constexpr int VB = QC + 128;
  static constexpr int BWD_SMEM =
      (4 * 64 * RS + 2 * 64 * RSL + 3 * VB + 2 * 4 * QC + 2 * 4 * 64 + 2 * 4 * 64 + 256) * 8;
};

// Sum v[0..3] over the 8 lanes that differ in lane bits 2..4.  Every lane ends with the total
// of element c = 2*b4 + b3 (lanes that differ only in bit 2 hold the same value).
RGP_DEVINL double reduce4_over_g(const double (&v)[4], int lane) {
  const bool h4 = lane & 16, h3 = lane & 8;
  double u[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double send = h4 ? v[i] : v[i + 2];
    const double keep = h4 ? v[i + 2] : v[i];
    u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  const double send = h3 ? u[0] : u[1];
  const double keep = h3 ? u[1] : u[0];
  double w = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  w += __shfl_xor_sync(0xffffffffu, w, 4);
  return w;
}

template <int COUNT, int NTHR>
RGP_DEVINL void copy_tile_n(double* dst, const double* __restrict__ src, int tid) {
  const double2* s2 = reinterpret_cast<const double2*>(src);
  double2* d2 = reinterpret_cast<double2*>(dst);
  for (int i = tid; i < COUNT / 2; i += NTHR) d2[i] = s2[i];
}

// upper-triangle tiles of a diagonal block over 16 warps: warps 0-3 own 3 tiles, 4-15 own 2
// (every scheduler, i.e. warps w, w+4, w+8, w+12, gets 9 tiles)
RGP_DEVINL void diag_tiles16(int wid, int (&ti)[3], int (&tj)[3], int& cnt) {
  const int first = wid < 4 ? 3 * wid : 12 + 2 * (wid - 4);
  cnt = wid < 4 ? 3 : 2;
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    int idx = first + (s < cnt ? s : cnt - 1);
    int r = 0, off = 0;
    while (idx >= off + 8 - r) { off += 8 - r; ++r; }
    ti[s] = r;
    tj[s] = r + idx - off;
  }
}

template <int QC, int CNT>
RGP_DEVINL void stage1_diag16_n(const double* __restrict__ sZ, const double* __restrict__ sZw,
                                const double* __restrict__ v, int qk,
                                const int (&ti)[3], const int (&tj)[3], int lane, double (&acc)[3][2]) {
  constexpr int RS = P2Cfg16<QC>::RS;
  const int g = lane >> 2, t = lane & 3;
  const double* pa[CNT];
  const double* pb[CNT];
#pragma unroll
  for (int s = 0; s < CNT; ++s) {
    pa[s] = sZw + (8 * ti[s] + g) * RS + t;
    pb[s] = sZ + (8 * tj[s] + g) * RS + t;
    const double hi = v[QC + 8 * ti[s] + g];
    const double2 hj = *reinterpret_cast<const double2*>(v + QC + 8 * tj[s] + 2 * t);
    acc[s][0] = hi + hj.x;
    acc[s][1] = hi + hj.y;
  }
#pragma unroll 2
  for (int k0 = 0; k0 < qk; k0 += 4) {
    double a[CNT], b[CNT];
#pragma unroll
    for (int s = 0; s < CNT; ++s) {
      a[s] = pa[s][k0];
      b[s] = pb[s][k0];
    }
#pragma unroll
    for (int s = 0; s < CNT; ++s) dmma(acc[s][0], acc[s][1], a[s], b[s]);
  }
}

template <int QC>
__global__ void __launch_bounds__(P2_THREADS16, 1)
k_psi2_bwd16(int64_t rc, int Mp, int nt, int nblocks, int qk, const double* __restrict__ Zt,
             const double* __restrict__ Ct, const double* __restrict__ wrow,
             const double* __restrict__ HP, double* __restrict__ lam, double* __restrict__ Wq,
             double* __restrict__ ACCp, int dbg, long long* __restrict__ trace) {
  // dbg: timing-experiment mask (0 in production; results are wrong when non-zero):
  // 1 skip exp, 2 skip lambda sums, 4 skip Wq reduce, 8 skip folds, 16 skip L store, 32 skip flushes,
  // 64 skip pre-weighted tile build, 128 skip the per-row barrier
  using C = P2Cfg16<QC>;
  constexpr int RS = C::RS, VB = C::VB, NJ = C::NJ;
  extern __shared__ __align__(16) double smem[];
  double* sZI = smem;
  double* sZJ = sZI + 64 * RS;
  double* sL = sZJ + 64 * RS;                     // 2 slots of 64*RSL
  double* sV = sL + 2 * 64 * RSL;                 // 3 slots of VB
  double* sWq = sV + 3 * VB;                      // [2][4 wr][QC]
  double* sLr = sWq + 2 * 4 * QC;                 // [2][4 wc][64]
  double* sLc = sLr + 2 * 4 * 64;                 // [2][4 wr][64]
  double* sT = sLc + 2 * 4 * 64;                  // exp table
  double* sZW = sT + 256;                         // 2 slots of 64*RS: ws(n) * Z'_I, built one row ahead

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int wr = wid >> 2, wc = wid & 3, g = lane >> 2, t = lane & 3;
  const int R = gridDim.x, G = gridDim.y;
  const int64_t per = (rc + R - 1) / R;
  const int64_t r0 = per * blockIdx.x, r1 = (r0 + per < rc) ? r0 + per : rc;
  const int cta = blockIdx.y * R + blockIdx.x;
  double* lamg = lam + (size_t)blockIdx.y * rc * Mp;
  double* Wqg = Wq + (size_t)blockIdx.y * rc * QC;
  double* accp = ACCp + (size_t)cta * Mp * QC;
  const int qbase = wc * (QC / 4);                // this warp's q columns in stage 2
  const int cidx = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);   // element held after reduce4_over_g
  exp_table_init(sT, tid);

  int curI = -1, curJ = -1;
  for (int b = blockIdx.y; b < nblocks; b += G) {
    int I, J;
    block_ij(b, nt, I, J);
    const bool diag = (I == J);
    __syncthreads();
    if (I != curI) copy_tile_n<64 * RS, P2_THREADS16>(sZI, Zt + (size_t)I * 64 * RS, tid);
    if (J != curJ) copy_tile_n<64 * RS, P2_THREADS16>(sZJ, Zt + (size_t)J * 64 * RS, tid);
    curI = I;
    curJ = J;
    const double* hI = HP + (size_t)I * rc * 64;
    const double* hJ = HP + (size_t)J * rc * 64;
    const double* cb = Ct + (size_t)b * 4096;
    auto vec_load = [&](int64_t n) -> double {
      if (tid >= VB || n >= r1) return 0.0;
      if (tid < QC) return wrow[n * QC + tid];
      if (tid < QC + 64) return hI[n * 64 + (tid - QC)];
      return hJ[n * 64 + (tid - QC - 64)];
    };
    if (tid < VB) {                                // vectors of the first two rows
      if (r0 < r1) sV[(r0 % 3) * VB + tid] = vec_load(r0);
      if (r0 + 1 < r1) sV[((r0 + 1) % 3) * VB + tid] = vec_load(r0 + 1);
    }
    double accI[2][NJ][2], accJ[2][NJ][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) accI[i][j][0] = accI[i][j][1] = accJ[i][j][0] = accJ[i][j][1] = 0.0;
    __syncthreads();

    // Pre-weighted stage-1 operand: sZW[n & 1][m][q] = ws_nq * Z'_I[m][q].  Keeping the multiply out
    // of the DMMA loops matters: ncu showed warps stalled at in-loop DMULs waiting for an FP64
    // pipe that other warps' 16-cycle DMMAs keep busy (30 % of all stall samples).  Thread t owns
    // column q = t % QC of every row it touches, so it needs a single ws value per row.
    auto ws_of = [&](int64_t n) -> double { return n < r1 ? wrow[n * QC + (tid & (QC - 1))] : 0.0; };
    auto build_zw = [&](int64_t n, double wsq) {
      double* dst = sZW + (n & 1) * 64 * RS;
      constexpr int PER = 64 * QC / P2_THREADS16;
      double tmp[PER];                               // batch the loads: dst may alias sZI for the compiler
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int e = tid + u * P2_THREADS16;
        tmp[u] = sZI[(e / QC) * RS + (e & (QC - 1))];
      }
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int e = tid + u * P2_THREADS16;
        dst[(e / QC) * RS + (e & (QC - 1))] = tmp[u] * wsq;
      }
    };
    if (r0 < r1) build_zw(r0, ws_of(r0));
    if (r0 + 1 < r1) build_zw(r0 + 1, ws_of(r0 + 1));
    __syncthreads();

    auto flush_wq = [&](int64_t n) {
      const int s = (int)(n & 1);
      if (tid < QC) {
        const double* p = sWq + s * 4 * QC + tid;
        double v = p[0] + p[QC] + p[2 * QC] + p[3 * QC];
        red_add(Wqg + n * QC + tid, diag ? v : 2.0 * v);
      }
    };

    // stage 2-I: T = L ZJ ; accI += ws T ; Wq partial
    auto stage2I = [&](const double* v, const double* Lb, int s) {
      double T[2][NJ][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) T[i][j][0] = T[i][j][1] = 0.0;
      const double* pa = Lb + (16 * wr + g) * RSL + t;
      const double* pb = sZJ + t * RS + qbase + g;
      const int kend2i = (dbg & 256) ? 0 : 64;
#pragma unroll 4
      for (int k0 = 0; k0 < kend2i; k0 += 4) {
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
        const int q = qbase + 8 * j + 2 * t;
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
      if (dbg & 4) {
        if (wp[0] == 1.2345e-300) sWq[0] = 1.0;
        return;
      }
      if constexpr (NJ == 2) {
        const double tot = reduce4_over_g(wp, lane);
        if (!(lane & 4))
          sWq[s * 4 * QC + wr * QC + qbase + 8 * (cidx >> 1) + 2 * t + (cidx & 1)] = tot;
      } else {
#pragma unroll
        for (int c = 0; c < 2 * NJ; ++c) {
          double x = wp[c];
          x += __shfl_xor_sync(0xffffffffu, x, 4);
          x += __shfl_xor_sync(0xffffffffu, x, 8);
          x += __shfl_xor_sync(0xffffffffu, x, 16);
          if (g == 0) sWq[s * 4 * QC + wr * QC + qbase + 8 * (c >> 1) + 2 * t + (c & 1)] = x;
        }
      }
    };

    // Row pipeline with skewed warp groups.  Between barrier n and barrier n+1 every warp runs
    //   S2(n)   = stage 2-I + fold + stage 2-J on the L tile of row n        (DMMA heavy)
    //   S1E(n+1)= stage 1 + exp epilogue of row n+1 -> L tile (n+1)&1         (DMMA, then DFMA/SHFL/STS)
    // group A (warps 0-3, 8-11) in that order, group B (4-7, 12-15) in the opposite order, so each
    // scheduler (2 A + 2 B warps) always has warps in both kinds of phase.
    const bool groupB = (wid >> 2) & 1;

    if (!diag) {
      double creg[2][2][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          double2 c2 = *reinterpret_cast<const double2*>(cb + (16 * wr + 8 * i + g) * 64 + 16 * wc + 8 * j + 2 * t);
          creg[i][j][0] = c2.x;
          creg[i][j][1] = c2.y;
        }
      auto s1e = [&](int64_t n) {
        const int s = (int)(n & 1);
        const double* v = sV + (n % 3) * VB;
        double* Lb = sL + s * 64 * RSL;
        double acc[2][2][2];
        const double* pa = sZW + s * 64 * RS + (16 * wr + g) * RS + t;
        const double* pb = sZJ + (16 * wc + g) * RS + t;
        const double* vI = v + QC + 16 * wr + g;
        const double* vJ = v + QC + 64 + 16 * wc + 2 * t;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const double hi = vI[8 * i];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const double2 hj = *reinterpret_cast<const double2*>(vJ + 8 * j);
            acc[i][j][0] = hi + hj.x;
            acc[i][j][1] = hi + hj.y;
          }
        }
        const int kend1 = (dbg & 1024) ? 0 : qk;
#pragma unroll 4
        for (int k0 = 0; k0 < kend1; k0 += 4) {
          double a[2], bb[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) a[i] = pa[i * 8 * RS + k0];
#pragma unroll
          for (int j = 0; j < 2; ++j) bb[j] = pb[j * 8 * RS + k0];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], bb[j]);
        }
        double rs[2] = {0.0, 0.0};
        double cs[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const double l0 = creg[i][j][0] * ((dbg & 1) ? acc[i][j][0] : exp_tab(acc[i][j][0], sT));
            const double l1 = creg[i][j][1] * ((dbg & 1) ? acc[i][j][1] : exp_tab(acc[i][j][1], sT));
            if (!(dbg & 16))
              *reinterpret_cast<double2*>(Lb + (16 * wr + 8 * i + g) * RSL + 16 * wc + 8 * j + 2 * t) =
                  make_double2(l0, l1);
            rs[i] += l0 + l1;
            cs[2 * j] += l0;
            cs[2 * j + 1] += l1;
          }
        if (dbg & 2) {
          if (rs[0] + cs[0] == 1.2345e-300) sLr[0] = 1.0;
          return;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 1);
          rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 2);
        }
        if (t == 0) {
          sLr[s * 256 + wc * 64 + 16 * wr + g] = rs[0];
          sLr[s * 256 + wc * 64 + 16 * wr + 8 + g] = rs[1];
        }
        const double tot = reduce4_over_g(cs, lane);
        if (!(lane & 4)) sLc[s * 256 + wr * 64 + 16 * wc + 8 * (cidx >> 1) + 2 * t + (cidx & 1)] = tot;
      };
      auto s2i = [&](int64_t n) {
        const int s = (int)(n & 1);
        stage2I(sV + (n % 3) * VB, sL + s * 64 * RSL, s);
      };
      auto s2j = [&](int64_t n) {
        const int s = (int)(n & 1);
        const double* v = sV + (n % 3) * VB;
        const double* Lb = sL + s * 64 * RSL;
        // stage 2-J: TJ[m',q] = sum_m L[m,m'] ZI[m,q];  accJ += ws_q TJ   (weight applied after the MMA)
        const double* pa = Lb + t * RSL + 16 * wr + g;
        const double* pb = sZI + t * RS + qbase + g;
        double TJ[2][NJ][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) TJ[i][j][0] = TJ[i][j][1] = 0.0;
        const int kend2j = (dbg & 512) ? 0 : 64;
#pragma unroll 4
        for (int k0 = 0; k0 < kend2j; k0 += 4) {
          double a[2], bq[NJ];
#pragma unroll
          for (int i = 0; i < 2; ++i) a[i] = pa[k0 * RSL + 8 * i];
#pragma unroll
          for (int j = 0; j < NJ; ++j) bq[j] = pb[k0 * RS + 8 * j];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) dmma(TJ[i][j][0], TJ[i][j][1], a[i], bq[j]);
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const double2 wq = *reinterpret_cast<const double2*>(v + qbase + 8 * j + 2 * t);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            accJ[i][j][0] = fma(wq.x, TJ[i][j][0], accJ[i][j][0]);
            accJ[i][j][1] = fma(wq.y, TJ[i][j][1], accJ[i][j][1]);
          }
        }
      };
      if (r0 < r1) s1e(r0);
      __syncthreads();
      for (int64_t n = r0; n < r1; ++n) {
        const int s = (int)(n & 1);
        if (dbg & 32) {
        } else if (tid >= 64 && tid < 128) {
          const int m = tid - 64;
          const double* p = sLr + s * 256 + m;
          red_add(lamg + n * Mp + I * 64 + m, p[0] + p[64] + p[128] + p[192]);
        } else if (tid >= 128 && tid < 192) {
          const int m = tid - 128;
          const double* p = sLc + s * 256 + m;
          red_add(lamg + n * Mp + J * 64 + m, p[0] + p[64] + p[128] + p[192]);
        }
        if (n > r0 && !(dbg & 32)) flush_wq(n - 1);
        // optional timeline trace (profiling builds of the bench only): CTA 0, block 1, rows 8..23
        const bool tr = trace && blockIdx.x == 0 && blockIdx.y == 0 && b == 1 && n >= r0 + 8 && n < r0 + 24 && lane == 0;
        long long* tp = trace + ((n - r0 - 8) * 16 + wid) * 8;
        if (tr) tp[0] = clock64();
        const double nxt = vec_load(n + 2);
        const double wsn = ws_of(n + 2);
        // Both groups start and end the iteration inside an MMA loop; the scalar exp epilogue (E) and
        // the build of the pre-weighted tile of row n+2 (slot n&1, last read by S1E(n) before
        // barrier n) sit in the middle, at different times for the two groups:
        //   A: S1 E build | S2I | S2J          B: S2I | S1 E build | S2J
        if (groupB) {
          s2i(n);
          if (tr) tp[1] = clock64();
          if (n + 1 < r1) s1e(n + 1);
          if (n + 2 < r1 && !(dbg & 64)) build_zw(n + 2, wsn);
          if (tr) tp[2] = clock64();
          s2j(n);
        } else {
          if (n + 1 < r1) s1e(n + 1);
          if (n + 2 < r1 && !(dbg & 64)) build_zw(n + 2, wsn);
          if (tr) tp[1] = clock64();
          s2i(n);
          if (tr) tp[2] = clock64();
          s2j(n);
        }
        if (tid < VB) sV[((n + 2) % 3) * VB + tid] = nxt;
        if (tr) tp[3] = clock64();
        if (!(dbg & 128)) __syncthreads();
        if (tr) tp[4] = clock64();
      }
    } else {
      int ti[3], tj[3], cnt;
      diag_tiles16(wid, ti, tj, cnt);
      double creg[3][2];
#pragma unroll
      for (int s3 = 0; s3 < 3; ++s3) {
        double2 c2 = *reinterpret_cast<const double2*>(cb + (8 * ti[s3] + g) * 64 + 8 * tj[s3] + 2 * t);
        creg[s3][0] = c2.x;
        creg[s3][1] = c2.y;
      }
      auto s1e = [&](int64_t n) {
        const int s = (int)(n & 1);
        const double* v = sV + (n % 3) * VB;
        double* Lb = sL + s * 64 * RSL;
        double acc[3][2];
        if (cnt == 3) stage1_diag16_n<QC, 3>(sZI, sZW + s * 64 * RS, v, qk, ti, tj, lane, acc);
        else stage1_diag16_n<QC, 2>(sZI, sZW + s * 64 * RS, v, qk, ti, tj, lane, acc);
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3)
          if (s3 < cnt) {
            const double l0 = creg[s3][0] * exp_tab(acc[s3][0], sT);
            const double l1 = creg[s3][1] * exp_tab(acc[s3][1], sT);
            const int m = 8 * ti[s3] + g, mp = 8 * tj[s3] + 2 * t;
            *reinterpret_cast<double2*>(Lb + m * RSL + mp) = make_double2(l0, l1);
            if (ti[s3] != tj[s3]) {
              Lb[mp * RSL + m] = l0;
              Lb[(mp + 1) * RSL + m] = l1;
            }
          }
      };
      if (r0 < r1) s1e(r0);
      __syncthreads();
      for (int64_t n = r0; n < r1; ++n) {
        const int s = (int)(n & 1);
        const double* v = sV + (n % 3) * VB;
        const double* Lb = sL + s * 64 * RSL;
        if (tid >= 64 && tid < 128) {             // lambda_m = full row sum of the symmetric tile
          const int m = tid - 64;
          const double2* row = reinterpret_cast<const double2*>(Lb + m * RSL);
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
        const double nxt = vec_load(n + 2);
        const double wsn = ws_of(n + 2);
        if (groupB) {
          if (n + 1 < r1) s1e(n + 1);
          if (n + 2 < r1 && !(dbg & 64)) build_zw(n + 2, wsn);
          stage2I(v, Lb, s);
        } else {
          stage2I(v, Lb, s);
          if (n + 2 < r1 && !(dbg & 64)) build_zw(n + 2, wsn);
          if (n + 1 < r1) s1e(n + 1);
        }
        if (tid < VB) sV[((n + 2) % 3) * VB + tid] = nxt;
        if (!(dbg & 128)) __syncthreads();
      }
    }
    if (r1 > r0) flush_wq(r1 - 1);               // the loop ended with a barrier
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int q = qbase + 8 * j + 2 * t;
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
