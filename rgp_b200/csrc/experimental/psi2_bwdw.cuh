// psi2_bwdw.cuh - Psi2 backward kernel with warp-specialised roles (12 warps per CTA).
//
// Same decomposition, inputs and outputs as k_psi2_bwd / k_psi2_bwdp (psi2_kernels.cuh): a CTA owns a row
// range and walks the 64 x 64 blocks of the pair matrix; per row  stage 1 (exponents), epilogue (exp, L = C p,
// lambda sums), stage 2-I (T = L Z'_J, folds), stage 2-J (L^T Z'_I).
//
// Why roles.  ncu on the 8-warp kernels (profiles/SUMMARY_r02.md section 3): a DMMA holds its warp for ~16
// cycles of fixed issue latency ("wait" is the top stall, even in a kernel stripped to its DMMA loops), the
// warp is in-order, and with two warps per scheduler each warp has only the other warp's 16 DMMA cycles to
// issue everything else.  Every scalar instruction in a DMMA warp's stream - the 16 exps per thread and row,
// the DMULs that scale the stage-1 operand, the lambda sums - therefore delays that warp's next DMMA and the
// FP64 pipe idles (81 % busy, 74 % DMMA).  Here the eight DMMA warps keep only what needs their accumulators:
//
//   warps 0-7  (DMMA)   stage 1 of row n+1 from a PRE-SCALED operand tile (no DMUL), raw exponents -> shared;
//                       stage 2-I / 2-J of row n; the folds of T (accI, W, accJ)
//   warps 8-11 (scalar) one per scheduler: turn the exponent tile of row n+1 into L = C exp(E) IN PLACE,
//                       row / column sums of L (lambda), Psi2 side sum in registers (fused pass), build the
//                       pre-scaled tile ws(n+2) * Z'_I, flush lambda / W of the previous rows, issue the TMA
//                       row-vector batches
//
// Hand-offs: named barrier 1 (DMMA warps arrive after storing the exponents - which also says they are done
// reading the pre-scaled tile; scalar warps wait) and one CTA barrier per row.  Diagonal blocks: the DMMA
// warps compute the 36 upper-triangle tiles (balanced 5/4 per warp) and store them mirrored, the scalar warps
// see a full symmetric tile.
#pragma once
#include "../psi2_bwdp.cuh"

namespace rgp {
namespace fast {

constexpr int PW_THREADS = 384;
constexpr int PW_DMMA = 256;

template <int QC, int NJ_>
struct P2CfgW {
  static constexpr int RS = QC + tile_pad(QC);
  static constexpr int NJ = NJ_;
  static constexpr int QS = 16 * NJ_;
  static constexpr int VR = 8;
  static constexpr int VBB = VR * (QC + 128);
  // Z'_I, Z'_J | 2 exponent / L tiles | pre-scaled tile | 2 vector batches | W partials | lambda rows, cols | exp table | mbar
  static constexpr int SMEM_D = 2 * 64 * RS + 2 * 64 * RSL + 64 * RS + 2 * VBB + 2 * 4 * QS + 2 * 64 + 2 * 4 * 64 + 256 + 2;
  static constexpr int SMEM = SMEM_D * 8;
};

RGP_DEVINL void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
RGP_DEVINL void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <int QC, int NJ_, bool FUSE = false>
__global__ void __launch_bounds__(PW_THREADS, 1)
k_psi2_bwdw(int64_t rc, int Mp, int nt, int nblocks, int qk, const double* __restrict__ Zt,
            const double* __restrict__ Ct, const double* __restrict__ wrow, const double* __restrict__ HP,
            double* __restrict__ lam, double* __restrict__ Wq, double* __restrict__ ACCp,
            double* __restrict__ P2p = nullptr) {
  using C = P2CfgW<QC, NJ_>;
  constexpr int RS = C::RS, NJ = C::NJ, QS = C::QS, VR = C::VR, VBB = C::VBB;
  extern __shared__ __align__(16) double smem[];
  double* sZI = smem;
  double* sZJ = sZI + 64 * RS;
  double* sX = sZJ + 64 * RS;                     // 2 slots [64][RSL]: exponents of a row, then L in place
  double* sWZ = sX + 2 * 64 * RSL;                // [64][RS]  ws(n) * Z'_I, the A operand of stage 1
  double* sVb = sWZ + 64 * RS;                    // 2 batch slots of VBB
  double* sWq = sVb + 2 * VBB;                    // [2][4][QS]
  double* sLr = sWq + 2 * 4 * QS;                 // [2][64]     row sums of L
  double* sLc = sLr + 2 * 64;                     // [2][4][64]  column sums of L per scalar warp
  double* sT = sLc + 2 * 4 * 64;                  // exp table, 256 entries
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sT + 256);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool is_mma = wid < 8;
  const int R = gridDim.x, G = gridDim.y;
  const int64_t per = (rc + R - 1) / R;
  const int64_t r0 = per * blockIdx.x, r1 = (r0 + per < rc) ? r0 + per : rc;
  const int cta = blockIdx.y * R + blockIdx.x;
  double* lamg = lam + (size_t)blockIdx.y * rc * Mp;
  double* Wqg = Wq + (size_t)blockIdx.y * rc * QC;
  double* accp = ACCp + (size_t)cta * Mp * QC;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
  }
  exp_table_init(sT, tid);
  uint32_t phase_bits = 0u;
  if (r0 >= r1) {
    if constexpr (FUSE)
      for (int b = blockIdx.y; b < nblocks; b += G) {
        double* out = P2p + ((size_t)b * R + blockIdx.x) * 4096;
        for (int i = tid; i < 4096; i += PW_THREADS) out[i] = 0.0;
      }
    return;
  }
  // ---- DMMA-warp coordinates (as in k_psi2_bwd)
  const int wr = (wid >> 1) & 3, wc = wid & 1, g = lane >> 2, t = lane & 3;
  const int qbase = wc * (QS / 2);
  // ---- scalar-warp coordinates: warp e owns rows 16 e .. 16 e + 15 of the tile; lane (r4, c8) owns rows
  //      16 e + 4 r4 + i (i < 4) and the column pairs 2 c8 + 16 u (u < 4): a quarter warp reads 128 contiguous bytes
  const int e = wid - 8, r4 = lane >> 3, c8 = lane & 7;

  int curI = -1, curJ = -1;
  for (int b = blockIdx.y; b < nblocks; b += G) {
    int I, J;
    block_ij(b, nt, I, J);
    const bool diag = (I == J);
    __syncthreads();                              // previous block done with every shared buffer
    if (I != curI)
      for (int i = tid; i < 64 * RS / 2; i += PW_THREADS)
        reinterpret_cast<double2*>(sZI)[i] = reinterpret_cast<const double2*>(Zt + (size_t)I * 64 * RS)[i];
    if (J != curJ)
      for (int i = tid; i < 64 * RS / 2; i += PW_THREADS)
        reinterpret_cast<double2*>(sZJ)[i] = reinterpret_cast<const double2*>(Zt + (size_t)J * 64 * RS)[i];
    curI = I;
    curJ = J;
    const double* hI = HP + (size_t)I * rc * 64;
    const double* hJ = HP + (size_t)J * rc * 64;
    const double* cb = Ct + (size_t)b * 4096;
    auto issue = [&](int64_t k) {                  // batch k = rows [r0 + k VR, ...) -> slot k & 1
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
    auto vec = [&](int64_t n, const double*& sw, const double*& vI, const double*& vJ) {
      const int idx = (int)(n - r0);
      const double* base = sVb + ((idx / VR) & 1) * VBB;
      const int w = idx % VR;
      sw = base + w * QC;
      vI = base + VR * QC + w * 64;
      vJ = base + VR * QC + VR * 64 + w * 64;
    };
    if (tid == PW_DMMA) {
      issue(0);
      issue(1);
    }
    __syncthreads();                              // Z' tiles visible
    await(0);
    const int64_t nb = (r1 - r0 + VR - 1) / VR;    // batches of this block
    int64_t awaited = 1;                           // every thread awaits every batch exactly once, in order
    auto need_row = [&](int64_t n) {               // make sure the batch holding row n has landed
      const int64_t k = (n - r0) / VR;
      while (awaited <= k && awaited < nb) { await(awaited); ++awaited; }
    };

    if (is_mma) {
      // =================================================================== DMMA warps
      double accI[2][NJ][2], accJ[2][NJ][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) accI[i][j][0] = accI[i][j][1] = accJ[i][j][0] = accJ[i][j][1] = 0.0;
      int ti[5], tj[5], cnt;
      diag_tiles(wid, ti, tj, cnt);
      // stage 1 of row n from the pre-scaled tile; raw exponents -> X
      auto stage1_store = [&](int64_t n, double* __restrict__ X) {
        const double *sw, *vI, *vJ;
        vec(n, sw, vI, vJ);
        if (!diag) {
          double acc[2][4][2];
          const double* pa = sWZ + (16 * wr + g) * RS + t;
          const double* pb = sZJ + (32 * wc + g) * RS + t;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const double hi = vI[16 * wr + g + 8 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const double2 hj = *reinterpret_cast<const double2*>(vJ + 32 * wc + 2 * t + 8 * j);
              acc[i][j][0] = hi + hj.x;
              acc[i][j][1] = hi + hj.y;
            }
          }
#pragma unroll 2
          for (int k0 = 0; k0 < qk; k0 += 4) {
            double a[2], bq[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = pa[i * 8 * RS + k0];
#pragma unroll
            for (int j = 0; j < 4; ++j) bq[j] = pb[j * 8 * RS + k0];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], bq[j]);
          }
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<double2*>(X + (16 * wr + 8 * i + g) * RSL + 32 * wc + 8 * j + 2 * t) =
                  make_double2(acc[i][j][0], acc[i][j][1]);
        } else {
          double acc[5][2];
          const double* pa[5];
          const double* pb[5];
#pragma unroll
          for (int s = 0; s < 5; ++s) {
            pa[s] = sWZ + (8 * ti[s] + g) * RS + t;
            pb[s] = sZI + (8 * tj[s] + g) * RS + t;
            const double hi = vI[8 * ti[s] + g];
            const double2 hj = *reinterpret_cast<const double2*>(vI + 8 * tj[s] + 2 * t);
            acc[s][0] = hi + hj.x;
            acc[s][1] = hi + hj.y;
          }
          if (cnt == 5) {
#pragma unroll 2
            for (int k0 = 0; k0 < qk; k0 += 4) {
              double a[5], bq[5];
#pragma unroll
              for (int s = 0; s < 5; ++s) { a[s] = pa[s][k0]; bq[s] = pb[s][k0]; }
#pragma unroll
              for (int s = 0; s < 5; ++s) dmma(acc[s][0], acc[s][1], a[s], bq[s]);
            }
          } else {
#pragma unroll 2
            for (int k0 = 0; k0 < qk; k0 += 4) {
              double a[4], bq[4];
#pragma unroll
              for (int s = 0; s < 4; ++s) { a[s] = pa[s][k0]; bq[s] = pb[s][k0]; }
#pragma unroll
              for (int s = 0; s < 4; ++s) dmma(acc[s][0], acc[s][1], a[s], bq[s]);
            }
          }
#pragma unroll
          for (int s = 0; s < 5; ++s)
            if (s < cnt) {
              const int m = 8 * ti[s] + g, mp = 8 * tj[s] + 2 * t;
              *reinterpret_cast<double2*>(X + m * RSL + mp) = make_double2(acc[s][0], acc[s][1]);
              if (ti[s] != tj[s]) {                // mirror: the scalar warps convert a full symmetric tile
                X[mp * RSL + m] = acc[s][0];
                X[(mp + 1) * RSL + m] = acc[s][1];
              }
            }
        }
      };
      // ---- prologue: exponents of the first row (the scalar warps have built its pre-scaled tile)
      __syncthreads();                            // P1: sWZ = ws(r0) Z'_I
      stage1_store(r0, sX + (int)(r0 & 1) * 64 * RSL);
      __syncthreads();                            // P2: exponents of r0 stored
      __syncthreads();                            // P3: L(r0), lambda(r0), sWZ = ws(r0+1) Z'_I ready
      for (int64_t n = r0; n < r1; ++n) {
        const int s = (int)(n & 1);
        const double* Lr = sX + s * 64 * RSL;
        if (n + 1 < r1) {
          need_row(n + 1);
          stage1_store(n + 1, sX + (s ^ 1) * 64 * RSL);
          bar_arrive(1, PW_THREADS);              // exponents of row n+1 stored; done reading the pre-scaled tile
        }
        const double *sw, *vI, *vJ;
        vec(n, sw, vI, vJ);
        // ---- stage 2-I: T = L Z'_J; accI += ws T; W partial
        {
          double T[2][NJ][2];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) T[i][j][0] = T[i][j][1] = 0.0;
          const double* pa = Lr + (16 * wr + g) * RSL + t;
          const double* pb = sZJ + t * RS + qbase + g;
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
            const int q = qbase + 8 * j + 2 * t;
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
        }
        // ---- stage 2-J (off-diagonal blocks): accJ += ws (L^T Z'_I), ws applied after the MMA
        if (!diag) {
          double TJ[2][NJ][2];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) TJ[i][j][0] = TJ[i][j][1] = 0.0;
          const double* pa = Lr + t * RSL + 16 * wr + g;
          const double* pb = sZI + t * RS + qbase + g;
#pragma unroll 2
          for (int k0 = 0; k0 < 64; k0 += 4) {
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
            const double2 wq = *reinterpret_cast<const double2*>(sw + qbase + 8 * j + 2 * t);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              accJ[i][j][0] = fma(wq.x, TJ[i][j][0], accJ[i][j][0]);
              accJ[i][j][1] = fma(wq.y, TJ[i][j][1], accJ[i][j][1]);
            }
          }
        }
        __syncthreads();                          // row barrier: L(n+1), lambda(n+1), W(n), pre-scaled tile of n+2 complete
      }
      // flush the CTA-private dZ accumulators of this block
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
    } else {
      // =================================================================== scalar warps
      const int row0 = 16 * e + 4 * r4;             // first of this lane's 4 rows
      double creg[4][4][2];                         // C = s2^2 sym(dL_dpsi2) at this lane's 32 elements
      double pacc[4][4][2];                         // FUSE: Psi2 partial of the block
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double2 c2 = *reinterpret_cast<const double2*>(cb + (row0 + i) * 64 + 2 * c8 + 16 * u);
          creg[i][u][0] = c2.x;
          creg[i][u][1] = c2.y;
          pacc[i][u][0] = pacc[i][u][1] = 0.0;
        }
      // sWZ <- ws(n) * Z'_I
      auto build_wz = [&](int64_t n) {
        const double *sw, *vI, *vJ;
        vec(n, sw, vI, vJ);
        double2 w2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) w2[u] = *reinterpret_cast<const double2*>(sw + 2 * c8 + 16 * u);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int off = (row0 + i) * RS + 2 * c8 + 16 * u;
            double2 z = *reinterpret_cast<const double2*>(sZI + off);
            z.x *= w2[u].x;
            z.y *= w2[u].y;
            *reinterpret_cast<double2*>(sWZ + off) = z;
          }
        if constexpr (QC > 64) {                    // columns 64 .. QC-1 of wider tiles
          for (int col = 64 + 2 * c8; col < QC; col += 16)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int off = (row0 + i) * RS + col;
              double2 z = *reinterpret_cast<const double2*>(sZI + off);
              const double2 w = *reinterpret_cast<const double2*>(sw + col);
              z.x *= w.x;
              z.y *= w.y;
              *reinterpret_cast<double2*>(sWZ + off) = z;
            }
        }
      };
      // exponents of a row -> L in place, lambda partials -> slot s
      auto convert = [&](double* __restrict__ X, int s) {
        double rs[4] = {0.0, 0.0, 0.0, 0.0};
        double cs[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) cs[c] = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {               // one row (8 exps) at a time keeps the live registers low
          double2 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const double2*>(X + (row0 + i) * RSL + 2 * c8 + 16 * u);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const double p0 = exp_tab(v[u].x, sT), p1 = exp_tab(v[u].y, sT);
            if constexpr (FUSE) {
              pacc[i][u][0] += p0;
              pacc[i][u][1] += p1;
            }
            const double l0 = creg[i][u][0] * p0, l1 = creg[i][u][1] * p1;
            *reinterpret_cast<double2*>(X + (row0 + i) * RSL + 2 * c8 + 16 * u) = make_double2(l0, l1);
            rs[i] += l0 + l1;
            cs[2 * u] += l0;
            cs[2 * u + 1] += l1;
          }
        }
        // row sums: over the 8 lanes of a row group (c8)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 1);
          rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 2);
          rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 4);
        }
        if (c8 < 4) sLr[s * 64 + row0 + c8] = (c8 == 0) ? rs[0] : (c8 == 1) ? rs[1] : (c8 == 2) ? rs[2] : rs[3];
        // column sums over this warp's 16 rows: reduce over r4 (lane bits 3, 4) by recursive halving
        {
          const bool h4 = lane & 16, h3 = lane & 8;
          double u4[4], u2[2];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const double send = h4 ? cs[c] : cs[c + 4];
            const double keep = h4 ? cs[c + 4] : cs[c];
            u4[c] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const double send = h3 ? u4[c] : u4[c + 2];
            const double keep = h3 ? u4[c + 2] : u4[c];
            u2[c] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
          // lane holds cs index c = 4 b4 + 2 b3 + {0, 1}  ->  u = c >> 1, element c & 1
          const int cbase = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int c = cbase + k;
            sLc[s * 256 + e * 64 + 2 * c8 + 16 * (c >> 1) + (c & 1)] = u2[k];
          }
        }
      };
      // lambda of row n (slot n & 1) and W of row n - 1: this role's 128 threads flush them
      auto flush = [&](int64_t n) {
        const int s = (int)(n & 1);
        const int m = tid - PW_DMMA;                // 0 .. 127
        if (m < 64) {
          red_add(lamg + n * Mp + I * 64 + m, sLr[s * 64 + m]);
        } else if (!diag) {
          const double* p = sLc + s * 256 + (m - 64);
          red_add(lamg + n * Mp + J * 64 + (m - 64), p[0] + p[64] + p[128] + p[192]);
        }
      };
      auto flush_wq = [&](int64_t n) {
        const int m = tid - PW_DMMA;
        if (m < QS) {
          const double* p = sWq + (int)(n & 1) * 4 * QS + m;
          const double v = p[0] + p[QS] + p[2 * QS] + p[3 * QS];
          red_add(Wqg + n * QC + m, diag ? v : 2.0 * v);
        }
      };
      // ---- prologue
      build_wz(r0);
      __syncthreads();                            // P1
      __syncthreads();                            // P2: exponents of r0 stored
      if (r0 + 1 < r1) {
        need_row(r0 + 1);
        build_wz(r0 + 1);
      }
      convert(sX + (int)(r0 & 1) * 64 * RSL, (int)(r0 & 1));
      __syncthreads();                            // P3
      for (int64_t n = r0; n < r1; ++n) {
        const int s = (int)(n & 1);
        {                                          // refill the vector slot the previous batch has just left
          const int idx = (int)(n - r0);
          if (idx > 0 && idx % VR == 0 && tid == PW_DMMA) issue(idx / VR + 1);
        }
        flush(n);
        if (n > r0) flush_wq(n - 1);
        if (n + 1 < r1) {
          if (n + 2 < r1) need_row(n + 2);
          bar_sync(1, PW_THREADS);                // exponents of row n+1 stored; pre-scaled tile free
          if (n + 2 < r1) build_wz(n + 2);
          convert(sX + (s ^ 1) * 64 * RSL, s ^ 1);
        }
        __syncthreads();                          // row barrier
      }
      flush_wq(r1 - 1);
      if constexpr (FUSE) {
        double* out = P2p + ((size_t)b * R + blockIdx.x) * 4096;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int u = 0; u < 4; ++u)
            *reinterpret_cast<double2*>(out + (row0 + i) * 64 + 2 * c8 + 16 * u) = make_double2(pacc[i][u][0], pacc[i][u][1]);
      }
    }
  }
}

}  // namespace fast
}  // namespace rgp
