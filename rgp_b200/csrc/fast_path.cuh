// fast_path.cuh - host-side driver of the tiled kernels: workspace plan and launch order.
#pragma once
#include <algorithm>
#include <vector>

#include "context.cuh"
#include "fast_prep.cuh"
#include "psi2_kernels.cuh"
#include "psi2_bwdp.cuh"
#include "psi2_small.cuh"
#ifdef RGP_DEBUG
#include "experimental/psi2_bwdw.cuh"   // warp-specialised variant: a measured negative result, experiment builds only
#endif

namespace rgp {
namespace fast {

static inline bool supported(int M, int Q) { return Q >= 1 && Q <= 128 && M >= 1; }
// the fused pass keeps one more shared tile; at QC = 128 it does not fit next to the Z' tiles
static inline bool fused_supported(int Q) { return Q <= 64; }
static inline int qc_for(int Q) { return Q <= 16 ? 16 : (Q <= 32 ? 32 : (Q <= 64 ? 64 : 128)); }

template <int QC, int NJ>
static int init_bwdp() {
  RGP_CUDA(cudaFuncSetAttribute((k_psi2_bwdp<QC, NJ, false>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                P2CfgP<QC, NJ>::SMEM));
  if constexpr (QC <= 64)
    RGP_CUDA(cudaFuncSetAttribute((k_psi2_bwdp<QC, NJ, true>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  P2CfgP<QC, NJ>::FUSED_SMEM));
  return 0;
}

template <int QC>
static int init_qc() {
  RGP_CUDA(cudaFuncSetAttribute(k_psi2_fwd<QC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                P2Cfg<QC>::FWD_SMEM + (QC == 64 ? 65536 : 0)));
  RGP_CUDA(cudaFuncSetAttribute(k_psi2_bwd<QC>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2Cfg<QC>::BWD_SMEM));
  if constexpr (QC <= 64)
    RGP_CUDA(cudaFuncSetAttribute((k_psi2_bwd<QC, true>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  P2Cfg<QC>::BWD_FUSED_SMEM));
  return 0;
}

template <int QT>
static int init_small() {
  constexpr int fwd = small_smem_doubles(PS_MS_MAX, QT, false) * 8, bwd = small_smem_doubles(PS_MS_MAX, QT, true) * 8;
  RGP_CUDA(cudaFuncSetAttribute((k_psi2_small<QT, 0, 1, 2, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, fwd));
  RGP_CUDA(cudaFuncSetAttribute((k_psi2_small<QT, 1, 1, 2, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, bwd));
  RGP_CUDA(cudaFuncSetAttribute((k_psi2_small<QT, 2, 1, 2, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, bwd));
  if constexpr (QT <= 3) {     // two jobs per warp exist for Q <= 23 only
    RGP_CUDA(cudaFuncSetAttribute((k_psi2_small<QT, 1, 2, 2, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, bwd));
    RGP_CUDA(cudaFuncSetAttribute((k_psi2_small<QT, 2, 2, 2, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, bwd));
  }
#ifdef RGP_DEBUG
  RGP_CUDA(cudaFuncSetAttribute((k_psi2_small<QT, 0, 1, 4, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, fwd));
  if constexpr (QT <= 3)
    RGP_CUDA(cudaFuncSetAttribute((k_psi2_small<QT, 1, 1, 4, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  small_smem_doubles(PS_MS_MAX, QT, true, 1) * 8));
#endif
  return 0;
}

static int init(rgp_psi_ctx*) {
  RGP_TRY(init_small<1>());
  RGP_TRY(init_small<2>());
  RGP_TRY(init_small<3>());
  RGP_TRY(init_small<4>());
  RGP_TRY(init_small<6>());
  RGP_TRY(init_qc<16>());
  RGP_TRY(init_qc<32>());
  RGP_TRY(init_qc<64>());
  RGP_TRY(init_qc<128>());
  RGP_TRY((init_bwdp<16, 1>()));
  RGP_TRY((init_bwdp<32, 2>()));
  RGP_TRY((init_bwdp<64, 3>()));
  RGP_TRY((init_bwdp<64, 4>()));
  RGP_TRY((init_bwdp<128, 4>()));
#ifdef RGP_DEBUG
  RGP_CUDA(cudaFuncSetAttribute((k_psi2_bwdw<64, 4, false>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (P2CfgW<64, 4>::SMEM)));
  RGP_CUDA(cudaFuncSetAttribute((k_psi2_bwdw<64, 4, true>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (P2CfgW<64, 4>::SMEM)));
#endif
  return 0;
}

struct Shape {
  int M, Mp, nt, nblocks, Q, QC, qk, RS;
  int64_t rc;
};

static Shape make_shape(const rgp_psi_ctx* h, int64_t N, int M, int Q) {
  Shape s;
  s.M = M;
  s.Mp = (int)round_up(M, BM);
  s.nt = s.Mp / BM;
  s.nblocks = s.nt * (s.nt + 1) / 2;
  s.Q = Q;
  s.QC = qc_for(Q);
  s.qk = (int)round_up(Q, 4);
  s.RS = s.QC + tile_pad(s.QC);
  int64_t rc = h->row_chunk > 0 ? h->row_chunk : ((int64_t)1 << 20);
  s.rc = std::min<int64_t>(N, rc);
  return s;
}

// row ranges R x block groups G so that R*G ~ target CTAs
static void pick_grid(int64_t rows, int nblocks, int target, int* R, int* G) {
  int64_t rmax = std::max<int64_t>(1, (rows + 3) / 4);
  if (rmax >= target) {
    *R = target;
    *G = 1;
  } else {
    *R = (int)rmax;
    *G = (int)std::min<int64_t>(nblocks, (target + rmax - 1) / rmax);
  }
}

static GemmOperand op(const double* p, int64_t sr, int64_t sk, int64_t rows) {
  GemmOperand o;
  o.p = p;
  o.sr = sr;
  o.sk = sk;
  o.rows = (int)rows;
  return o;
}

static GemmEpi epi_plain(double* out, int64_t ld, int64_t split_stride) {
  GemmEpi e;
  e.mode = EPI_PLAIN;
  e.bias = nullptr;
  e.variance = 0.0;
  e.scale = nullptr;
  e.scale_ld = 0;
  e.M = 0;
  e.out = out;
  e.out_ld = ld;
  e.split_stride = split_stride;
  return e;
}

template <int QC>
static int launch_fwd(rgp_psi_ctx* h, cudaStream_t st, const Shape& s, int64_t rows, int R, int G,
                      const double* Zt, const double* w, const double* HP, double* P2p) {
  RGP_LAUNCH(h, st, "psi2_fwd", (k_psi2_fwd<QC>), dim3(R, G), P2_THREADS, P2Cfg<QC>::FWD_SMEM + h->fwd_smem_pad, rows,
             s.nt, s.nblocks, s.qk, Zt, w, HP, P2p);
  return 0;
}

template <int QC, int NJ>
static int launch_bwdp(rgp_psi_ctx* h, cudaStream_t st, const Shape& s, int64_t rows, int R, int G,
                       const double* Zt, const double* Ct, const double* w, const double* HP,
                       double* lam, double* Wq, double* ACCp, double* P2p) {
  if constexpr (QC <= 64) {
    if (P2p) {     // fused forward + backward: the kernel also accumulates the Psi2 partial tiles
      RGP_LAUNCH(h, st, "psi2_bwd_fused", (k_psi2_bwdp<QC, NJ, true>), dim3(R, G), P2_THREADS,
                 (P2CfgP<QC, NJ>::FUSED_SMEM), rows, s.Mp, s.nt, s.nblocks, s.qk, Zt, Ct, w, HP, lam, Wq, ACCp, 0, P2p);
      return 0;
    }
  }
  for (int qoff = 0; qoff < s.Q; qoff += 16 * NJ)   // two passes for 64 < Q <= 128 (one if Q <= 64)
    RGP_LAUNCH(h, st, "psi2_bwd", (k_psi2_bwdp<QC, NJ, false>), dim3(R, G), P2_THREADS, (P2CfgP<QC, NJ>::SMEM), rows,
               s.Mp, s.nt, s.nblocks, s.qk, Zt, Ct, w, HP, lam, Wq, ACCp, qoff, (double*)nullptr);
  return 0;
}

// Which Psi2 backward kernel serves a pass (profiles/SUMMARY_r02.md section 2): the software-pipelined kernel
// (psi2_bwdp.cuh; pad-4 Z' tiles) for the fused pass and for 32 < Q <= 48 (48 stage-2 columns instead of 64), the
// row-at-a-time kernel (tile_pad(QC) Z' tiles) everywhere else.  bwd_pipe 0 / 1 force one of them (A/B, tests).
static inline bool use_pipelined(const rgp_psi_ctx* h, int QC, int Q, bool fused) {
  const bool narrow = QC == 64 && Q <= 48;
  return h->bwd_pipe == 1 || (h->bwd_pipe == 2 && (fused || narrow));
}

template <int QC>
static int launch_bwd(rgp_psi_ctx* h, cudaStream_t st, const Shape& s, int64_t rows, int R, int G,
                      const double* Zt, const double* Ct, const double* w, const double* HP,
                      double* lam, double* Wq, double* ACCp, double* P2p) {
  // Two Psi2 backward kernels (profiles/SUMMARY_r02.md, "A/B").  The software-pipelined kernel (psi2_bwdp.cuh)
  // sizes its stage-2 width to Q in steps of 16 and is 3 % faster when the pass also accumulates Psi2 (fused);
  // the row-at-a-time kernel is 1.5 - 7 % faster for the plain backward pass at every other shape measured.
  // bwd_pipe: 2 = that choice (default), 0 / 1 = force one of them (A/B measurements, parity tests).
  const bool fused = P2p != nullptr;
#ifdef RGP_DEBUG
  if constexpr (QC == 64) {
    // warp-specialised kernel (experimental/psi2_bwdw.cuh): full stage-2 width, enough rows per CTA to amortise its prologue
    if (h->bwd_pipe == 3 && s.Q > 48 && G == 1 && rows >= (int64_t)64 * R) {
      if (fused)
        RGP_LAUNCH(h, st, "psi2_bwd_fused", (k_psi2_bwdw<64, 4, true>), dim3(R, G), PW_THREADS, (P2CfgW<64, 4>::SMEM),
                   rows, s.Mp, s.nt, s.nblocks, s.qk, Zt, Ct, w, HP, lam, Wq, ACCp, P2p);
      else
        RGP_LAUNCH(h, st, "psi2_bwd", (k_psi2_bwdw<64, 4, false>), dim3(R, G), PW_THREADS, (P2CfgW<64, 4>::SMEM),
                   rows, s.Mp, s.nt, s.nblocks, s.qk, Zt, Ct, w, HP, lam, Wq, ACCp, (double*)nullptr);
      return 0;
    }
  }
#endif
  const bool pipe = use_pipelined(h, QC, s.Q, fused);
  if constexpr (QC == 128) {
    if (fused) return set_error(RGP_PSI_ERR_INVALID, "fused pass is not built for Q > 64");
  }
  if (pipe) {
    if constexpr (QC == 16) return launch_bwdp<16, 1>(h, st, s, rows, R, G, Zt, Ct, w, HP, lam, Wq, ACCp, P2p);
    else if constexpr (QC == 32) return launch_bwdp<32, 2>(h, st, s, rows, R, G, Zt, Ct, w, HP, lam, Wq, ACCp, P2p);
    else if constexpr (QC == 64) {
      if (s.Q <= 48) return launch_bwdp<64, 3>(h, st, s, rows, R, G, Zt, Ct, w, HP, lam, Wq, ACCp, P2p);
      return launch_bwdp<64, 4>(h, st, s, rows, R, G, Zt, Ct, w, HP, lam, Wq, ACCp, P2p);
    } else return launch_bwdp<128, 4>(h, st, s, rows, R, G, Zt, Ct, w, HP, lam, Wq, ACCp, P2p);
  }
  if constexpr (QC <= 64) {
    if (fused) {
      RGP_LAUNCH(h, st, "psi2_bwd_fused", (k_psi2_bwd<QC, true>), dim3(R, G), P2_THREADS, P2Cfg<QC>::BWD_FUSED_SMEM,
                 rows, s.Mp, s.nt, s.nblocks, s.qk, Zt, Ct, w, HP, lam, Wq, ACCp, 0, P2p);
      return 0;
    }
  }
  for (int qoff = 0; qoff < QC; qoff += P2Cfg<QC>::QS) {   // two passes for QC = 128
    if (false) {
    }
#ifdef RGP_DEBUG
#define RGP_ABL(D)                                                                                           \
  else if (h->debug_skip == D) {                                                                             \
    cudaFuncSetAttribute((k_psi2_bwd<QC, false, D>), cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                         P2Cfg<QC>::BWD_SMEM);                                                               \
    RGP_LAUNCH(h, st, "psi2_bwd", (k_psi2_bwd<QC, false, D>), dim3(R, G), P2_THREADS, P2Cfg<QC>::BWD_SMEM,   \
               rows, s.Mp, s.nt, s.nblocks, s.qk, Zt, Ct, w, HP, lam, Wq, ACCp, qoff, (double*)nullptr);     \
  }
    RGP_ABL(1) RGP_ABL(2) RGP_ABL(3) RGP_ABL(4) RGP_ABL(7) RGP_ABL(8) RGP_ABL(14) RGP_ABL(30) RGP_ABL(46) RGP_ABL(78)
    RGP_ABL(142) RGP_ABL(16) RGP_ABL(110) RGP_ABL(174) RGP_ABL(206)
#undef RGP_ABL
#endif
    else {
      RGP_LAUNCH(h, st, "psi2_bwd", (k_psi2_bwd<QC>), dim3(R, G), P2_THREADS, P2Cfg<QC>::BWD_SMEM, rows,
                 s.Mp, s.nt, s.nblocks, s.qk, Zt, Ct, w, HP, lam, Wq, ACCp, qoff, (double*)nullptr);
    }
  }
  return 0;
}

// ---- small inducing sets (psi2_small.cuh): one CTA holds the whole pair matrix of a row -------------------
// One kernel configuration of the small family: CTA size, supertile slots per warp, buffers of L, k split, and the
// work table (which warp computes which supertiles / jobs).
struct SmallVariant {
  int warps = PS_WARPS, s1 = 2, nbuf = 2, KS = 1, JMAX = 1;
  SmallSched sched;
};
struct SmallPlan {
  bool ok = false;
  int Ms = 0, Mp16 = 0, QT = 0;
  SmallVariant fwd, bwd, fused;     // forward only / backward only / backward + Psi2
};

// DMMAs per row of the 64 x 64 block kernels / of the small kernel (both passes), used to choose between them
static inline double block_dmma_per_row(const Shape& s, bool pipelined) {
  const int ks1 = s.qk / 4;
  int qs = s.QC > 64 ? 64 : s.QC;                       // stage-2 columns per pass
  if (pipelined && s.QC == 64 && s.Q <= 48) qs = 48;
  const int passes = s.QC > 64 ? 2 : 1;
  const double diag = 36.0 * ks1 + 8.0 * (qs / 8) * 16, off = 64.0 * ks1 + 2 * 8.0 * (qs / 8) * 16;
  const double fwd = s.nt * 36.0 * ks1 + (s.nblocks - s.nt) * 64.0 * ks1;
  return fwd + passes * (s.nt * diag + (s.nblocks - s.nt) * off);
}
static inline int small_valid_tiles(int si, int sj, int M8) {
  const int vi = 16 * si + 8 < M8 ? 2 : 1, vj = 16 * sj + 8 < M8 ? 2 : 1;
  return si == sj ? (vi == 2 ? 4 : 1) : vi * vj;     // 8x8 tiles the kernel computes on this supertile
}
static inline int small_col_ksteps(int sk, int M8) { return 16 * sk + 8 < M8 ? 4 : 2; }   // k-steps of a supertile column
static inline double small_dmma_per_row(int M, int Ms, int QT, int qk) {
  const int M8 = (M + 7) & ~7;
  double s1 = 0, s2 = 0;
  int ksteps = 0;
  for (int sk = 0; sk < Ms; ++sk) ksteps += small_col_ksteps(sk, M8);
  for (int i = 0; i < Ms; ++i)
    for (int j = i; j < Ms; ++j) s1 += small_valid_tiles(i, j, M8) * (qk / 4);
  for (int sp = 0; sp < Ms; ++sp) s2 += (16 * sp + 8 < M8 ? 2 : 1) * QT * ksteps;
  return 2 * s1 + s2;
}

// The work table of the small kernels.  Items: the supertiles of the upper triangle (stage 1 + exp) and, for the
// backward pass, jobs = (16-row strip, 1 / KS of the k-steps).
// Warp w issues on SM sub-partition w % 4, so the items are dealt greedily (largest first) to the least loaded
// sub-partition, then to its least loaded warp with a free slot.  Costs are FP64-pipe cycles: 16 per DMMA plus
// the scalar epilogue / fold work.
static void small_schedule(int M, int Ms, int QT, int qk, int KS, int JMAX, int warps, int s1cap, bool bwd, SmallSched* sc) {
  memset(sc, 0, sizeof(*sc));
  const int M8 = (M + 7) & ~7;
  struct Item { double cost; int kind, a, b, c, d; };   // kind 0: supertile index a; kind 1: job (strip a, columns [b, c), slot d)
  std::vector<Item> items;
  int u = 0;
  for (int i = 0; i < Ms; ++i)
    for (int j = i; j < Ms; ++j, ++u) {
      const int tiles = small_valid_tiles(i, j, M8);
      items.push_back({tiles * ((qk / 4) * 16.0 + 45.0) + 30.0, 0, u, 0, 0, 0});
    }
  int kslots = 1;
  if (bwd) {
    // k-steps are numbered 4 * (supertile column) + (0..3); a column whose second half is padding has only 2, so the
    // valid ones are listed first and each strip's list is cut into KS contiguous ranges of equal length
    int ks_list[4 * PS_MS_MAX], nks = 0;
    for (int sk = 0; sk < Ms; ++sk)
      for (int kk = 0; kk < small_col_ksteps(sk, M8); ++kk) ks_list[nks++] = 4 * sk + kk;
    for (int sp = 0; sp < Ms; ++sp) {
      int slot = 0;
      const int rowsets = 16 * sp + 8 < M8 ? 2 : 1;
      const int ksp = rowsets == 2 ? KS : (KS + 1) / 2;      // a half strip (8 valid rows) is half the work per k-step
      for (int i = 0; i < ksp; ++i) {
        const int a = nks * i / ksp, b = nks * (i + 1) / ksp;
        if (b <= a) continue;
        const int kb = ks_list[a], ke = ks_list[b - 1] + 1;
        items.push_back({(b - a) * (rowsets * QT * 16.0 + 4.0) + ((ke + 3) / 4 - kb / 4) * 12.0 + 80.0 + 14.0 * QT, 1, sp,
                         kb, ke, slot});
        ++slot;
      }
      kslots = std::max(kslots, slot);
    }
  }
  std::stable_sort(items.begin(), items.end(), [](const Item& x, const Item& y) { return x.cost > y.cost; });
  double wload[PS_WARPS] = {0}, pload[4] = {0};
  int njobs = 0;
  for (const Item& it : items) {
    int best = -1;
    for (int w = 0; w < warps; ++w) {
      if (it.kind == 0 ? sc->ns[w] >= s1cap : sc->nj[w] >= JMAX) continue;
      if (best < 0 || pload[w & 3] < pload[best & 3] - 1e-9 ||
          (pload[w & 3] < pload[best & 3] + 1e-9 && wload[w] < wload[best] - 1e-9))
        best = w;
    }
    if (best < 0) best = 0;     // cannot happen: 16 * PS_S1 >= 28 supertiles, 16 * JMAX >= jobs
    wload[best] += it.cost;
    pload[best & 3] += it.cost;
    if (it.kind == 0) {
      sc->su[best][sc->ns[best]++] = (signed char)it.a;
    } else {
      sc->jsp[njobs] = (signed char)it.a;
      sc->jkb[njobs] = (signed char)it.b;
      sc->jke[njobs] = (signed char)it.c;
      sc->jslot[njobs] = (signed char)it.d;
      sc->jw[best][sc->nj[best]++] = (signed char)njobs;
      ++njobs;
    }
  }
  sc->njobs = (signed char)njobs;
  sc->kslots = (signed char)kslots;
}

// small_m: 0 = never, 1 = whenever the shape fits, 2 (default) = where they measured faster (rule below)
static SmallPlan small_plan(const rgp_psi_ctx* h, const Shape& s) {
  SmallPlan p;
  if (h->small_m == 0) return p;
  const int Ms = (s.M + 15) / 16;
  int QT = s.Q / 8 + 1;                                 // stage-2 columns: Q and the ones column, in tiles of 8
  if (QT == 5) QT = 6;                                  // Q > 23: two column passes of QT / 2 tiles (4 or 6 tiles)
  if (Ms > PS_MS_MAX || QT > 6) return p;
  // measured (profiles/small_ab_r02.jsonl): with 6 - 7 super rows (M = 81 ... 112) the small kernels win up to ~0.8 of the
  // block kernels' DMMAs (M = 100 / 112 at Q = 7 ... 46); with fewer super rows only when they save more
  // (M = 33, Q = 20 wins at 0.37 and M = 50, Q = 20 at 0.69; M = 50, Q = 40 loses at 0.81, M = 64, Q = 16 at 1.3)
  const double ratio = small_dmma_per_row(s.M, Ms, QT, s.qk) / block_dmma_per_row(s, s.QC == 64 && s.Q <= 48);
  if (h->small_m == 2 && ratio > (Ms >= 6 ? 0.82 : 0.72)) return p;
  p.ok = true;
  p.Ms = Ms;
  p.Mp16 = 16 * Ms;
  p.QT = QT;
  // Variants (measured in profiles/small_ab_r02.jsonl; option small_warps = 16 / 8 forces one family for A/B runs):
  //  * 16 warps, one CTA per SM, L double-buffered (one barrier per row): serves everything;
  //  * 8 warps, two CTAs per SM with independent barriers (one CTA's stage 2 overlaps the other's stage 1), when two CTAs
  //    fit in shared memory (used for M <= 64, where it was measured): backward 5.02 -> 4.00 ms at (50, 20),
  //    3.21 -> 2.33 ms at (33, 20), forward 1.89 -> 1.61 / 1.68 -> 1.10 ms;
  //  * measured and NOT used (experiment builds only, small_warps = 8): 8-warp CTAs with 4 supertile slots per warp for
  //    M = 81 ... 112 - forward 4.33 -> 4.63 ms, and a single-buffered L with a second barrier per row so that two such
  //    CTAs fit - backward 10.30 -> 12.10 ms at (100, 20).
  const int want = h->small_warps;
  const int half_sm = 113 * 1024;
  auto make = [&](SmallVariant* v, int mode) -> bool {
    const bool bwd = mode != 0;
    v->warps = PS_WARPS; v->s1 = 2; v->nbuf = 2;
    if (want != 16 && Ms <= 4 && small_smem_doubles(Ms, QT, bwd, 2) * 8 <= half_sm) v->warps = 8;   // measured at M = 33 ... 64
#ifdef RGP_DEBUG
    if (want == 8 && v->warps == PS_WARPS) {
      if (mode == 0) { v->warps = 8; v->s1 = 4; }
      else if (mode == 1 && QT <= 3 && small_smem_doubles(Ms, QT, true, 1) * 8 <= half_sm) { v->warps = 8; v->s1 = 4; v->nbuf = 1; }
    }
#endif
    const int jmax_allowed = (QT > 3 || v->warps == 8) ? 1 : 2;     // one job per warp for wide Q and in the 8-warp CTAs
    v->KS = h->small_ks > 0 ? h->small_ks : (Ms >= 5 ? 2 : 4);      // ~13 ... 16 jobs per row with 16 warps
    while (v->KS > 1 && Ms * v->KS > jmax_allowed * v->warps) v->KS /= 2;
    if (bwd && Ms * v->KS > jmax_allowed * v->warps) return false;
    v->JMAX = std::max(1, (Ms * v->KS + v->warps - 1) / v->warps);
    small_schedule(s.M, Ms, QT, s.qk, v->KS, v->JMAX, v->warps, v->s1, bwd, &v->sched);
    return true;
  };
  if (!make(&p.fwd, 0) || !make(&p.bwd, 1) || !make(&p.fused, 2)) return SmallPlan();
  return p;
}

template <int QT, int MODE>
static int launch_small_qt(rgp_psi_ctx* h, cudaStream_t st, const Shape& s, const SmallPlan& p, int64_t rows, int Rs,
                           const double* Zt, const double* Ct, const double* w, const double* HP, double* lam,
                           double* Wq, double* ACCp, double* P2s) {
  const char* name = MODE == 0 ? "psi2_fwd" : (MODE == 1 ? "psi2_bwd" : "psi2_bwd_fused");
  const SmallVariant& v = MODE == 0 ? p.fwd : (MODE == 1 ? p.bwd : p.fused);
  const int smem = small_smem_doubles(p.Ms, QT, MODE != 0, v.nbuf) * 8;
#define RGP_SMALL_ARGS rows, s.M, s.Q, s.Mp, p.Ms, s.nt, s.qk, s.QC, s.RS, v.sched, Zt, Ct, w, HP, lam, Wq, ACCp, P2s
  if (v.s1 == 4) {
#ifdef RGP_DEBUG
    if constexpr (MODE == 0) RGP_LAUNCH(h, st, name, (k_psi2_small<QT, 0, 1, 4, 2>), Rs, 32 * v.warps, smem, RGP_SMALL_ARGS);
    else if constexpr (MODE == 1 && QT <= 3) RGP_LAUNCH(h, st, name, (k_psi2_small<QT, 1, 1, 4, 1>), Rs, 32 * v.warps, smem, RGP_SMALL_ARGS);
#endif
  } else if constexpr (MODE == 0) {
    RGP_LAUNCH(h, st, name, (k_psi2_small<QT, 0, 1, 2, 2>), Rs, 32 * v.warps, smem, RGP_SMALL_ARGS);
  } else if (v.JMAX == 1 || QT > 3) {
    RGP_LAUNCH(h, st, name, (k_psi2_small<QT, MODE, 1, 2, 2>), Rs, 32 * v.warps, smem, RGP_SMALL_ARGS);
  } else if constexpr (QT <= 3) {
    RGP_LAUNCH(h, st, name, (k_psi2_small<QT, MODE, 2, 2, 2>), Rs, 32 * v.warps, smem, RGP_SMALL_ARGS);
  }
#undef RGP_SMALL_ARGS
  return 0;
}

template <int MODE>
static int launch_small(rgp_psi_ctx* h, cudaStream_t st, const Shape& s, const SmallPlan& p, int64_t rows, int Rs,
                        const double* Zt, const double* Ct, const double* w, const double* HP, double* lam,
                        double* Wq, double* ACCp, double* P2s) {
  if (p.QT == 1) return launch_small_qt<1, MODE>(h, st, s, p, rows, Rs, Zt, Ct, w, HP, lam, Wq, ACCp, P2s);
  if (p.QT == 2) return launch_small_qt<2, MODE>(h, st, s, p, rows, Rs, Zt, Ct, w, HP, lam, Wq, ACCp, P2s);
  if (p.QT == 3) return launch_small_qt<3, MODE>(h, st, s, p, rows, Rs, Zt, Ct, w, HP, lam, Wq, ACCp, P2s);
  if (p.QT == 4) return launch_small_qt<4, MODE>(h, st, s, p, rows, Rs, Zt, Ct, w, HP, lam, Wq, ACCp, P2s);
  return launch_small_qt<6, MODE>(h, st, s, p, rows, Rs, Zt, Ct, w, HP, lam, Wq, ACCp, P2s);
}

// CTAs of a small-kernel launch: one (16 warps) or two (8 warps) per SM, and at least 4 rows per CTA so that the
// prologue (Z' tile, C fragments, exp table) is amortised at the N ~ 500 of the system-identification models
static inline int small_grid(const rgp_psi_ctx* h, const SmallVariant& v, int64_t rows) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((v.warps == 8 ? 2 : 1) * h->sm_count, (rows + 3) / 4));
}

// lam[0] += sum_{g>=1} lam[g]  (and the same for Wq); only launched when G > 1
__global__ void k_collapse(int64_t count, int G, double* __restrict__ buf) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double v = buf[i];
  for (int g = 1; g < G; ++g) v += buf[(int64_t)g * count + i];
  buf[i] = v;
}

static int static_prep(rgp_psi_ctx* h, cudaStream_t st, const Shape& s, const double* Z, double* o,
                       double* Zt, double* ZB) {
  RGP_LAUNCH(h, st, "center", k_center, s.Q, 128, 0, s.M, s.Q, Z, o);
  RGP_LAUNCH(h, st, "build_Z", k_build_Z, ceil_div((int64_t)s.Mp * s.RS, 256), 256, 0, s.M, s.Mp, s.Q,
             s.QC, s.RS, Z, o, Zt, ZB);
  return 0;
}

// HP[tile][rows][64] = b2 + A2 . ZB^T  (and Psi1 / L1 through the same GEMM)
static int gemm_nt(rgp_psi_ctx* h, cudaStream_t st, const char* name, const Shape& s, int64_t rows,
                   const double* A, const double* ZB, GemmEpi e) {
  dim3 grid(ceil_div(rows, 64), s.nt, 1);
  RGP_LAUNCH(h, st, name, (k_gemm<true, true>), grid, 256, 0, op(A, 2 * s.QC, 1, rows),
             op(ZB, 2 * s.QC, 1, s.Mp), (int64_t)2 * s.QC, e);
  return 0;
}

static int forward(rgp_psi_ctx* h, cudaStream_t st, int64_t N, int M, int Q, const double* mu,
                   const double* S, const double* Z, const double* ell, double variance,
                   double* psi0, double* psi1, double* psi2) {
  const Shape s = make_shape(h, N, M, Q);
  int R, G;
  pick_grid(s.rc, s.nblocks, 2 * h->sm_count, &R, &G);
  const int QC = s.QC;
  const SmallPlan sp = small_plan(h, s);
  const int Rs = small_grid(h, sp.fwd, s.rc);
  const size_t p2_count = sp.ok ? (size_t)Rs * sp.Mp16 * sp.Mp16 : (size_t)s.nblocks * R * 4096;
  size_t need = bump_size(Q, 8) + bump_size((size_t)s.Mp * s.RS, 8) + bump_size((size_t)s.Mp * 2 * QC, 8) +
                bump_size(p2_count, 8) + bump_size(s.rc * QC, 8) +
                bump_size(s.rc * 2 * QC, 8) * 2 + bump_size(s.rc, 8) * 2 + bump_size(s.rc * s.Mp, 8);
  RGP_TRY(arena_reserve(&h->ws, &h->ws_bytes, need));
  Bump b(h->ws, h->ws_bytes);
  double* o = b.take<double>(Q);
  double* Zt = b.take<double>((size_t)s.Mp * s.RS);
  double* ZB = b.take<double>((size_t)s.Mp * 2 * QC);
  double* P2p = b.take<double>(p2_count);
  double* w = b.take<double>(s.rc * QC);
  double* A2 = b.take<double>(s.rc * 2 * QC);
  double* A1 = b.take<double>(s.rc * 2 * QC);
  double* b2 = b.take<double>(s.rc);
  double* b1 = b.take<double>(s.rc);
  double* HP = b.take<double>(s.rc * s.Mp);

  RGP_TRY(static_prep(h, st, s, Z, o, Zt, ZB));
  if (psi0) RGP_LAUNCH(h, st, "fill_psi0", k_fill, ceil_div(N, 256), 256, 0, N, variance, psi0);
  int chunk = 0;
  for (int64_t r0 = 0; r0 < N; r0 += s.rc, ++chunk) {
    const int64_t rows = std::min(s.rc, N - r0);
    RGP_LAUNCH(h, st, "rowprep", k_rowprep, ceil_div(rows, 4), 128, 0, rows, Q, QC, mu + r0 * Q,
               S + r0 * Q, ell, o, w, A2, b2, psi1 ? A1 : (double*)nullptr, psi1 ? b1 : (double*)nullptr);
    GemmEpi e;
    e.mode = EPI_HP; e.bias = b2; e.variance = variance; e.scale = nullptr; e.scale_ld = 0; e.M = M;
    e.out = HP; e.out_ld = rows; e.split_stride = 0;
    RGP_TRY(gemm_nt(h, st, "hprime_gemm", s, rows, A2, ZB, e));
    if (psi1) {
      e.mode = EPI_PSI1; e.bias = b1; e.out = psi1 + r0 * M; e.out_ld = M;
      RGP_TRY(gemm_nt(h, st, "psi1_fwd", s, rows, A1, ZB, e));
    }
    if (sp.ok) {
      const int Rr = std::min(Rs, small_grid(h, sp.fwd, rows));
      RGP_TRY(launch_small<0>(h, st, s, sp, rows, Rr, Zt, nullptr, w, HP, nullptr, nullptr, nullptr, P2p));
      RGP_LAUNCH(h, st, "psi2_reduce", k_psi2_reduce_small, ceil_div((int64_t)M * M, 256), 256, 0, M, sp.Mp16, Rr,
                 variance * variance, P2p, (chunk > 0 || h->accumulate) ? 1 : 0, psi2);
      continue;
    }
    int Rc, Gc;
    pick_grid(rows, s.nblocks, 2 * h->sm_count, &Rc, &Gc);
    Rc = std::min(Rc, R);
    if (QC == 16) RGP_TRY(launch_fwd<16>(h, st, s, rows, Rc, Gc, Zt, w, HP, P2p));
    else if (QC == 32) RGP_TRY(launch_fwd<32>(h, st, s, rows, Rc, Gc, Zt, w, HP, P2p));
    else if (QC == 64) RGP_TRY(launch_fwd<64>(h, st, s, rows, Rc, Gc, Zt, w, HP, P2p));
    else RGP_TRY(launch_fwd<128>(h, st, s, rows, Rc, Gc, Zt, w, HP, P2p));
    RGP_LAUNCH(h, st, "psi2_reduce", k_psi2_reduce, dim3(s.nblocks, 16), 256, 0, M, s.nt, Rc, variance * variance,
               P2p, (chunk > 0 || h->accumulate) ? 1 : 0, psi2);
  }
  return 0;
}

static int backward(rgp_psi_ctx* h, cudaStream_t st, int64_t N, int M, int Q, const double* mu,
                    const double* S, const double* Z, const double* ell, double variance,
                    const double* dL0, double dL0c, const double* dL1, const double* dL2,
                    double* dmu, double* dS, double* dZ, double* dell, double* dvar,
                    double* psi1_out = nullptr, double* psi2_out = nullptr) {
  // psi1_out / psi2_out (optional): fused evaluation - the statistics come out of the same pass
  // (Psi1 from one more small GEMM, Psi2 from the backward kernel itself, see k_psi2_bwd FUSE)
  Shape s = make_shape(h, N, M, Q);
  if (use_pipelined(h, s.QC, Q, psi2_out != nullptr)) s.RS = s.QC + 4;   // the Z' tile layout follows the kernel
  const SmallPlan sp = small_plan(h, s);
  const SmallVariant& sv = psi2_out ? sp.fused : sp.bwd;
  const int Rs = small_grid(h, sv, s.rc);
  int R, G;
  pick_grid(s.rc, s.nblocks, h->sm_count, &R, &G);
  if (sp.ok) G = 1;
  const int QC = s.QC, Mp = s.Mp;
  const int ncta = sp.ok ? Rs * sv.sched.kslots : R * G;
  const size_t p2_count = sp.ok ? (size_t)Rs * sp.Mp16 * sp.Mp16 : (size_t)s.nblocks * R * 4096;
  const int tn_tiles = s.nt * (2 * QC / 64 > 0 ? (2 * QC + 63) / 64 : 1);
  const int splits = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)64, (int64_t)(2 * h->sm_count / std::max(1, tn_tiles)),
                                                                 (s.rc + 63) / 64}));
  const int nfin = (int)std::min<int64_t>(ceil_div(s.rc, 4), (int64_t)8 * h->sm_count);
  size_t need = bump_size(Q, 8) + bump_size((size_t)Mp * s.RS, 8) + bump_size((size_t)Mp * 2 * QC, 8) +
                bump_size((size_t)s.nblocks * 4096, 8) + bump_size(s.rc * QC, 8) +
                bump_size(s.rc * 2 * QC, 8) * 4 + bump_size(s.rc, 8) * 2 + bump_size(s.rc * Mp, 8) * 2 +
                bump_size((size_t)G * s.rc * Mp, 8) + bump_size((size_t)G * s.rc * QC, 8) +
                bump_size((size_t)ncta * Mp * QC, 8) + bump_size((size_t)splits * Mp * 2 * QC, 8) * 2 +
                bump_size((size_t)nfin * (QC + 1), 8) +
                (psi2_out ? bump_size(p2_count, 8) : 0);
  RGP_TRY(arena_reserve(&h->ws, &h->ws_bytes, need));
  Bump b(h->ws, h->ws_bytes);
  double* o = b.take<double>(Q);
  double* Zt = b.take<double>((size_t)Mp * s.RS);
  double* ZB = b.take<double>((size_t)Mp * 2 * QC);
  double* Ct = b.take<double>((size_t)s.nblocks * 4096);
  double* P2p = psi2_out ? b.take<double>(p2_count) : nullptr;
  double* w = b.take<double>(s.rc * QC);
  double* A2 = b.take<double>(s.rc * 2 * QC);
  double* A1 = b.take<double>(s.rc * 2 * QC);
  double* R2 = b.take<double>(s.rc * 2 * QC);
  double* R1 = b.take<double>(s.rc * 2 * QC);
  double* b2 = b.take<double>(s.rc);
  double* b1 = b.take<double>(s.rc);
  double* HP = b.take<double>(s.rc * Mp);
  double* L1 = b.take<double>(s.rc * Mp);
  double* lam = b.take<double>((size_t)G * s.rc * Mp);
  double* Wq = b.take<double>((size_t)G * s.rc * QC);
  double* ACCp = b.take<double>((size_t)ncta * Mp * QC);
  double* Gl = b.take<double>((size_t)splits * Mp * 2 * QC);
  double* GL = b.take<double>((size_t)splits * Mp * 2 * QC);
  double* part = b.take<double>((size_t)nfin * (QC + 1));

  if (!h->accumulate) {
    RGP_CUDA(cudaMemsetAsync(dZ, 0, sizeof(double) * M * Q, st));
    RGP_CUDA(cudaMemsetAsync(dell, 0, sizeof(double) * Q, st));
    RGP_CUDA(cudaMemsetAsync(dvar, 0, sizeof(double), st));
  }
  RGP_TRY(static_prep(h, st, s, Z, o, Zt, ZB));
  RGP_LAUNCH(h, st, "build_C", k_build_C, s.nblocks, 256, 0, M, s.nt, dL2, variance * variance, Ct);

  int chunk = 0;
  for (int64_t r0 = 0; r0 < N; r0 += s.rc, ++chunk) {
    const int64_t rows = std::min(s.rc, N - r0);
    int Rc, Gc;
    pick_grid(rows, s.nblocks, h->sm_count, &Rc, &Gc);
    Rc = std::min(Rc, R);
    Gc = std::min(Gc, G);
    const int Rr = std::min(Rs, small_grid(h, sv, rows));
    const int nc = sp.ok ? Rr * sv.sched.kslots : Rc * Gc;
    RGP_CUDA(cudaMemsetAsync(lam, 0, sizeof(double) * (size_t)Gc * rows * Mp, st));
    RGP_CUDA(cudaMemsetAsync(Wq, 0, sizeof(double) * (size_t)Gc * rows * QC, st));
    RGP_CUDA(cudaMemsetAsync(ACCp, 0, sizeof(double) * (size_t)nc * Mp * QC, st));
    RGP_LAUNCH(h, st, "rowprep", k_rowprep, ceil_div(rows, 4), 128, 0, rows, Q, QC, mu + r0 * Q,
               S + r0 * Q, ell, o, w, A2, b2, (dL1 || psi1_out) ? A1 : (double*)nullptr,
               (dL1 || psi1_out) ? b1 : (double*)nullptr);
    GemmEpi e;
    e.mode = EPI_HP; e.bias = b2; e.variance = variance; e.scale = nullptr; e.scale_ld = 0; e.M = M;
    e.out = HP; e.out_ld = rows; e.split_stride = 0;
    RGP_TRY(gemm_nt(h, st, "hprime_gemm", s, rows, A2, ZB, e));
    if (psi1_out) {
      e.mode = EPI_PSI1; e.bias = b1; e.out = psi1_out + r0 * M; e.out_ld = M;
      RGP_TRY(gemm_nt(h, st, "psi1_fwd", s, rows, A1, ZB, e));
    }
    if (dL1) {
      e.mode = EPI_L1; e.bias = b1; e.scale = dL1 + r0 * M; e.scale_ld = M; e.out = L1; e.out_ld = Mp;
      RGP_TRY(gemm_nt(h, st, "psi1_L1", s, rows, A1, ZB, e));
    }
    if (sp.ok) {
      if (P2p) RGP_TRY(launch_small<2>(h, st, s, sp, rows, Rr, Zt, Ct, w, HP, lam, Wq, ACCp, P2p));
      else RGP_TRY(launch_small<1>(h, st, s, sp, rows, Rr, Zt, Ct, w, HP, lam, Wq, ACCp, P2p));
    } else if (QC == 16) RGP_TRY(launch_bwd<16>(h, st, s, rows, Rc, Gc, Zt, Ct, w, HP, lam, Wq, ACCp, P2p));
    else if (QC == 32) RGP_TRY(launch_bwd<32>(h, st, s, rows, Rc, Gc, Zt, Ct, w, HP, lam, Wq, ACCp, P2p));
    else if (QC == 64) RGP_TRY(launch_bwd<64>(h, st, s, rows, Rc, Gc, Zt, Ct, w, HP, lam, Wq, ACCp, P2p));
    else RGP_TRY(launch_bwd<128>(h, st, s, rows, Rc, Gc, Zt, Ct, w, HP, lam, Wq, ACCp, P2p));
    if (P2p && sp.ok)
      RGP_LAUNCH(h, st, "psi2_reduce", k_psi2_reduce_small, ceil_div((int64_t)M * M, 256), 256, 0, M, sp.Mp16, Rr,
                 variance * variance, P2p, (chunk > 0 || h->accumulate) ? 1 : 0, psi2_out);
    else if (P2p)
      RGP_LAUNCH(h, st, "psi2_reduce", k_psi2_reduce, dim3(s.nblocks, 16), 256, 0, M, s.nt, Rc, variance * variance, P2p,
                 (chunk > 0 || h->accumulate) ? 1 : 0, psi2_out);
    if (Gc > 1) {
      RGP_LAUNCH(h, st, "collapse", k_collapse, ceil_div(rows * Mp, 256), 256, 0, rows * Mp, Gc, lam);
      RGP_LAUNCH(h, st, "collapse", k_collapse, ceil_div(rows * QC, 256), 256, 0, rows * QC, Gc, Wq);
    }
    // R2 = lam [rows x Mp] . ZB [Mp x 2QC]   (U | V);  R1 likewise from L1
    dim3 gnn(ceil_div(rows, 64), ceil_div(2 * QC, 64), 1);
    RGP_LAUNCH(h, st, "rows_gemm", (k_gemm<true, false>), gnn, 256, 0, op(lam, Mp, 1, rows),
               op(ZB, 1, 2 * QC, 2 * QC), (int64_t)Mp, epi_plain(R2, 2 * QC, 0));
    if (dL1)
      RGP_LAUNCH(h, st, "rows_gemm", (k_gemm<true, false>), gnn, 256, 0, op(L1, Mp, 1, rows),
                 op(ZB, 1, 2 * QC, 2 * QC), (int64_t)Mp, epi_plain(R1, 2 * QC, 0));
    const int nf = (int)std::min<int64_t>(ceil_div(rows, 4), (int64_t)8 * h->sm_count);
    RGP_LAUNCH(h, st, "rows_finalize", k_rows_finalize, nf, 128, sizeof(double) * 4 * (QC + 1), rows, M, Mp,
               Q, QC, 1, mu + r0 * Q, S + r0 * Q, ell, o, variance, lam, Wq, R2,
               dL1 ? L1 : (const double*)nullptr, dL1 ? R1 : (const double*)nullptr,
               dL0 ? dL0 + r0 : (const double*)nullptr, dL0c, dmu + r0 * Q, dS + r0 * Q, part);
    // Gl = lam^T [Mp x rows] . A2 [rows x 2QC]  (split over rows);  GL = L1^T . A1
    const int nsplit = (int)std::max<int64_t>(1, std::min<int64_t>(splits, (rows + 63) / 64));   // short K per split at small N
    dim3 gtn(s.nt, ceil_div(2 * QC, 64), nsplit);
    RGP_LAUNCH(h, st, "dz_gemm", (k_gemm<false, false>), gtn, 256, 0, op(lam, 1, Mp, Mp),
               op(A2, 1, 2 * QC, 2 * QC), rows, epi_plain(Gl, 2 * QC, (int64_t)Mp * 2 * QC));
    if (dL1)
      RGP_LAUNCH(h, st, "dz_gemm", (k_gemm<false, false>), gtn, 256, 0, op(L1, 1, Mp, Mp),
                 op(A1, 1, 2 * QC, 2 * QC), rows, epi_plain(GL, 2 * QC, (int64_t)Mp * 2 * QC));
    RGP_LAUNCH(h, st, "final_small", k_final_small, ceil_div((int64_t)M * Q + Q + 1, 128), 128, 0, M, Mp, Q,
               QC, ZB, Gl, nsplit, dL1 ? GL : (const double*)nullptr, nsplit, ACCp, nc, part, nf, dZ, dell, dvar);
  }
  return 0;
}

}  // namespace fast
}  // namespace rgp
