// psi2_kernels.cuh - the two O(N * M^2 * Q) kernels: Psi2 forward and Psi2 backward.
//
// Work decomposition.  The M x M pair matrix is cut into 64 x 64 blocks (I <= J, upper
// triangle).  A CTA owns a contiguous row range [r0, r1) and walks the blocks of its
// block group; for every block it streams its rows ONE ROW AT A TIME:
//
//   stage 1   E[m,m'] = H_nm + H_nm' + sum_q (ws_nq Z'_mq) Z'_m'q   64 x 64 x Q   (DMMA)
//             (ws = -w >= 0, so E is the exponent itself)
//   epilogue  p = exp(E)                                           4096 exps
//             forward : Psi2 tile += p                (register accumulators)
//             backward: L = C[m,m'] * p -> shared     (C = s2^2 * sym(dL_dpsi2), registers)
//   stage 2-I T[m,q]   = sum_m' L[m,m'] Z'_m'q                     64 x Q x 64   (DMMA)
//             accI[m,q] += ws_nq T[m,q];   Wq[n,q] += sum_m Z'_mq T[m,q]
//   stage 2-J accJ[m',q] += sum_m L[m,m'] (ws_nq Z'_mq)             64 x Q x 64   (DMMA)
//
// (diagonal blocks skip stage 2-J).  Row sums / column sums of L give lambda_nm.  The
// O(M*Q)-per-row remainder of the gradient algebra is in fast_prep.cuh.
//
// Why mma.sync m8n8k4 f64 (DMMA) and not DFMA register tiles: measured on this B200
// (profiles/microbench_r01.jsonl) DMMA and DFMA share the FP64 pipe and peak at the
// same 36.9 TFLOP/s, but an LDS.128 costs >= 2.5 SM-cycles even fully broadcast, so an
// 8x4 DFMA register tile fed from shared memory needs ~190 B/clk of LSU bandwidth and
// is LSU-bound; DMMA fragments need 0.5-1 B per FMA.  The arithmetic is still IEEE
// fp64 on the FP64 pipe; the roofline denominator is the DFMA-chain peak.
//
// Shared-memory tiles are [64][QC+4] / [64][68] doubles: row stride == 4 (mod 16)
// doubles makes every fragment load (rows by lane/4, cols by lane%4, or transposed)
// hit each bank exactly twice, the minimum for 256 B.
#pragma once
#include "common.cuh"

namespace rgp {
namespace fast {

constexpr int P2_THREADS = 256;
constexpr int RSL = 68;     // L tiles: stride == 4 (mod 16) for their 8-byte fragment loads, whatever the Z' tiles use

template <int QC>
struct P2Cfg {
  static constexpr int RS = QC + tile_pad(QC);
  // stage-2 output width per pass: Q in (64, 128] runs the backward kernel twice, once per 64-wide
  // q half (stage 1 + exp are recomputed; registers cannot hold 128-wide dZ accumulators)
  static constexpr int QS = QC > 64 ? 64 : QC;
  static constexpr int NJ = QS / 16;        // 8-wide q tiles per warp in stage 2
  static constexpr int VB = QC + 128;       // per-row vector slot: w[QC] | H_I[64] | H_J[64]
  static constexpr int VR = QC > 64 ? 1 : 8; // rows per TMA batch (QC = 128: shared memory is full)
  static constexpr int FWD_SMEM = (2 * 64 * RS + 2 * VR * VB + 256 + 2) * 8;
  static constexpr int BWD_SMEM =
      (2 * 64 * RS + 2 * 64 * RSL + 2 * VR * VB + 2 * 4 * QS + 2 * 2 * 64 + 2 * 4 * 64 + 256 + 2) * 8;
  // fused variant (also accumulates the Psi2 tile of the block in shared memory): + [64][RSL]
  static constexpr int BWD_FUSED_SMEM = BWD_SMEM + 64 * RSL * 8;
};

RGP_DEVINL void red_add(double* addr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

// ---- mbarrier / bulk-copy wrappers (PTX ISA: mbarrier, cp.async.bulk) -----------------------------
RGP_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
RGP_DEVINL void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
RGP_DEVINL void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
RGP_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
RGP_DEVINL void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
RGP_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}


// ---- TMA staging of the per-row vectors -------------------------------------------------------------
// The vectors a row needs - ws[QC], H_I[64], H_J[64] - are three contiguous runs in HBM.  One elected thread
// fetches them for VR consecutive rows with cp.async.bulk (UBLKCP in SASS) into a slot of a two-slot ring; an
// mbarrier per slot carries the byte count, every thread waits on it once per batch.  Row w of a batch sits at
// slot + w * VB in the layout the compute code reads ([ws | H_I | H_J]), so no compute thread issues a global
// load for operands inside the row loop.
template <int QC, int VR>
struct RowVecStage {
  static constexpr int VB = QC + 128;
  double* ring;              // 2 slots of VR * VB doubles
  uint64_t* mbar;            // 2 mbarriers
  uint32_t phase_bits = 0;   // per thread: parity of the next completion of each slot
  int64_t r0 = 0, r1 = 0;
  const double* wrow = nullptr;
  const double* hI = nullptr;
  const double* hJ = nullptr;
  int64_t next = 0;          // batches issued so far (meaningful on the issuing thread)
  int64_t nb = 0;

  RGP_DEVINL void init_barriers(int tid) {
    if (tid == 0) {
      mbar_init(&mbar[0], 1);
      mbar_init(&mbar[1], 1);
      mbar_fence_init();
    }
  }
  // start a block: (called by all threads after the CTA barrier that ends the previous block)
  RGP_DEVINL void begin(int64_t r0_, int64_t r1_, const double* w, const double* hi, const double* hj, int tid) {
    r0 = r0_; r1 = r1_; wrow = w; hI = hi; hJ = hj;
    nb = r1 > r0 ? (r1 - r0 + VR - 1) / VR : 0;
    next = 0;
    if (tid == 0 && nb > 0) issue();
  }
  RGP_DEVINL void issue() {                      // issuing thread only: batch `next` -> slot next & 1
    const int64_t n0 = r0 + next * VR;
    const int rows = (int)((r1 - n0 < VR) ? r1 - n0 : VR);
    const int slot = (int)(next & 1);
    double* dst = ring + slot * VR * VB;
    mbar_expect_tx(&mbar[slot], (uint32_t)(rows * VB * 8));
    for (int w = 0; w < rows; ++w) {
      bulk_g2s(dst + w * VB, wrow + (n0 + w) * QC, QC * 8, &mbar[slot]);
      bulk_g2s(dst + w * VB + QC, hI + (n0 + w) * 64, 512, &mbar[slot]);
      bulk_g2s(dst + w * VB + QC + 64, hJ + (n0 + w) * 64, 512, &mbar[slot]);
    }
    ++next;
  }
  // issuing thread, at a point where every thread has finished all rows < r0 + idx: refill free slots.
  // Batch j may overwrite the slot of batch j - 2 once all rows of batch j - 2 are done: (j - 1) VR <= idx.
  RGP_DEVINL void refill(int64_t idx) {
    while (next < nb && next <= idx / VR + 1 && (next - 1) * VR <= idx) issue();
  }
  // all threads, before the first use of row r0 + idx; returns the row's [ws | H_I | H_J]
  RGP_DEVINL const double* row(int64_t idx) {
    const int slot = (int)((idx / VR) & 1);
    if (idx % VR == 0) {
      mbar_wait(&mbar[slot], (phase_bits >> slot) & 1u);
      phase_bits ^= 1u << slot;
    }
    return ring + slot * VR * VB + (idx % VR) * VB;
  }
};

RGP_DEVINL void block_ij(int b, int nt, int& I, int& J) {
  int i = 0, rem = b;
  while (rem >= nt - i) { rem -= nt - i; ++i; }
  I = i;
  J = i + rem;
}

// copy a [64][RS] tile (contiguous in global) into shared memory, 16 B per thread-step
template <int COUNT>
RGP_DEVINL void copy_tile(double* dst, const double* __restrict__ src, int tid) {
  const double2* s2 = reinterpret_cast<const double2*>(src);
  double2* d2 = reinterpret_cast<double2*>(dst);
  for (int i = tid; i < COUNT / 2; i += P2_THREADS) d2[i] = s2[i];
}

// stage 1: acc[2][4][2] = H_m + H_m' + sum_q (ws_q ZI[m][q]) ZJ[m'][q]  (= the exponent) for this
// warp's 16 x 32 sub-block; ws = -w = S/(l^2 (2S+l^2)) >= 0 is what the row vector holds.
template <int QC>
RGP_DEVINL void stage1(const double* __restrict__ sZI, const double* __restrict__ sZJ,
                       const double* __restrict__ v, int qk, int wr, int wc, int lane,
                       double (&acc)[2][4][2]) {
  constexpr int RS = P2Cfg<QC>::RS;
  const int g = lane >> 2, t = lane & 3;
  const double* pa = sZI + (16 * wr + g) * RS + t;
  const double* pb = sZJ + (32 * wc + g) * RS + t;
  const double* sw = v;
  const double* vI = v + QC + 16 * wr + g;
  const double* vJ = v + QC + 64 + 32 * wc + 2 * t;
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
  if constexpr (tile_pad(QC) == 8) {
  const int qk8 = qk & ~7;
#pragma unroll 2
  for (int k0 = 0; k0 < qk8; k0 += 8) {         // two k-steps per trip, fragments by LDS.128 (see common.cuh)
    const double2 wv = *reinterpret_cast<const double2*>(sw + k0 + 2 * t);
    double2 a[2], b[4];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      a[i] = *reinterpret_cast<const double2*>(pa + i * 8 * RS + k0 + t);   // pa already points at column t: + t more = 2 t
      a[i].x *= wv.x;
      a[i].y *= wv.y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double2*>(pb + j * 8 * RS + k0 + t);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
  }
  if (qk & 4) {                                 // odd number of k-steps: one ordinary step at the end
    const int k0 = qk8;
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
  } else {
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
}

// Diagonal blocks (I == J) are symmetric: only the 36 upper-triangle 8x8 tiles of the 8x8
// tile grid are computed (warps 0-3 own 5 tiles, warps 4-7 own 4) and mirrored on store.
RGP_DEVINL void diag_tiles(int wid, int (&ti)[5], int (&tj)[5], int& cnt) {
  const int first = wid < 4 ? 5 * wid : 20 + 4 * (wid - 4);
  cnt = wid < 4 ? 5 : 4;
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    int idx = first + (s < cnt ? s : cnt - 1);
    int r = 0, off = 0;
    while (idx >= off + 8 - r) { off += 8 - r; ++r; }
    ti[s] = r;
    tj[s] = r + idx - off;
  }
}

template <int QC, int CNT>
RGP_DEVINL void stage1_diag_n(const double* __restrict__ sZ, const double* __restrict__ v, int qk,
                              const int (&ti)[5], const int (&tj)[5], int lane, double (&acc)[5][2]) {
  constexpr int RS = P2Cfg<QC>::RS;
  const int g = lane >> 2, t = lane & 3;
  const double* pa[CNT];
  const double* pb[CNT];
#pragma unroll
  for (int s = 0; s < CNT; ++s) {
    pa[s] = sZ + (8 * ti[s] + g) * RS + t;
    pb[s] = sZ + (8 * tj[s] + g) * RS + t;
    const double hi = v[QC + 8 * ti[s] + g];
    const double2 hj = *reinterpret_cast<const double2*>(v + QC + 8 * tj[s] + 2 * t);
    acc[s][0] = hi + hj.x;
    acc[s][1] = hi + hj.y;
  }
  if constexpr (tile_pad(QC) == 8) {
  const int qk8 = qk & ~7;
#pragma unroll 2
  for (int k0 = 0; k0 < qk8; k0 += 8) {
    const double2 wv = *reinterpret_cast<const double2*>(v + k0 + 2 * t);
    double2 a[CNT], b[CNT];
#pragma unroll
    for (int s = 0; s < CNT; ++s) {
      a[s] = *reinterpret_cast<const double2*>(pa[s] + k0 + t);
      a[s].x *= wv.x;
      a[s].y *= wv.y;
      b[s] = *reinterpret_cast<const double2*>(pb[s] + k0 + t);
    }
#pragma unroll
    for (int s = 0; s < CNT; ++s) dmma(acc[s][0], acc[s][1], a[s].x, b[s].x);
#pragma unroll
    for (int s = 0; s < CNT; ++s) dmma(acc[s][0], acc[s][1], a[s].y, b[s].y);
  }
  if (qk & 4) {
    const int k0 = qk8;
    const double wv = v[k0 + t];
    double a[CNT], b[CNT];
#pragma unroll
    for (int s = 0; s < CNT; ++s) {
      a[s] = pa[s][k0] * wv;
      b[s] = pb[s][k0];
    }
#pragma unroll
    for (int s = 0; s < CNT; ++s) dmma(acc[s][0], acc[s][1], a[s], b[s]);
  }
  } else {
#pragma unroll 2
  for (int k0 = 0; k0 < qk; k0 += 4) {
    const double wv = v[k0 + t];
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
}

// the tile count is warp-uniform; branching outside the k loop keeps it a straight-line,
// software-pipelined loop
template <int QC>
RGP_DEVINL void stage1_diag(const double* __restrict__ sZ, const double* __restrict__ v, int qk,
                            const int (&ti)[5], const int (&tj)[5], int cnt, int lane,
                            double (&acc)[5][2]) {
  if (cnt == 5) stage1_diag_n<QC, 5>(sZ, v, qk, ti, tj, lane, acc);
  else stage1_diag_n<QC, 4>(sZ, v, qk, ti, tj, lane, acc);
}

// =====================================================================================
// Forward: partial Psi2 tiles.  grid = (R row ranges, G block groups).
//   P2p[b][r][64][64] = sum_{n in range r} exp(E_n[m,m'])                    (no s2^2 yet)
// =====================================================================================
template <int QC>
__global__ void __launch_bounds__(P2_THREADS, 2)
k_psi2_fwd(int64_t rc, int nt, int nblocks, int qk, const double* __restrict__ Zt,
           const double* __restrict__ wrow, const double* __restrict__ HP,
           double* __restrict__ P2p) {
  using C = P2Cfg<QC>;
  constexpr int RS = C::RS, VB = C::VB, VR = C::VR;
  extern __shared__ __align__(16) double smem[];
  double* sZI = smem;
  double* sZJ = sZI + 64 * RS;
  double* sV = sZJ + 64 * RS;                     // row-vector ring: 2 slots of VR rows (TMA, see RowVecStage)
  double* sT = sV + 2 * VR * VB;                  // exp table, 256 entries
  RowVecStage<QC, VR> rv;
  rv.ring = sV;
  rv.mbar = reinterpret_cast<uint64_t*>(sT + 256);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int wr = wid >> 1, wc = wid & 1, g = lane >> 2, t = lane & 3;
  const int R = gridDim.x, G = gridDim.y;
  const int64_t per = (rc + R - 1) / R;
  const int64_t r0 = per * blockIdx.x, r1 = (r0 + per < rc) ? r0 + per : rc;
  exp_table_init(sT, tid);
  rv.init_barriers(tid);

  int curI = -1, curJ = -1;
  for (int b = blockIdx.y; b < nblocks; b += G) {
    int I, J;
    block_ij(b, nt, I, J);
    const bool diag = (I == J);
    __syncthreads();                              // previous block fully done with smem
    if (I != curI) copy_tile<64 * RS>(sZI, Zt + (size_t)I * 64 * RS, tid);
    if (J != curJ) copy_tile<64 * RS>(sZJ, Zt + (size_t)J * 64 * RS, tid);
    curI = I;
    curJ = J;
    const double* hI = HP + (size_t)I * rc * 64;
    const double* hJ = HP + (size_t)J * rc * 64;
    rv.begin(r0, r1, wrow, hI, hJ, tid);
    __syncthreads();
    double* out = P2p + ((size_t)b * R + blockIdx.x) * 4096;

    if (!diag) {
      double pacc[2][4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) pacc[i][j][0] = pacc[i][j][1] = 0.0;
      for (int64_t n = r0; n < r1; ++n) {
        if (tid == 0) rv.refill(n - r0);          // every thread is past the barrier of row n-1
        const double* v = rv.row(n - r0);
        double acc[2][4][2];
        stage1<QC>(sZI, sZJ, v, qk, wr, wc, lane, acc);
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            pacc[i][j][0] += exp_tab(acc[i][j][0], sT);
            pacc[i][j][1] += exp_tab(acc[i][j][1], sT);
          }
        __syncthreads();                          // all reads of row n's vectors done
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int m = 16 * wr + 8 * i + g, mp = 32 * wc + 8 * j + 2 * t;
          *reinterpret_cast<double2*>(out + m * 64 + mp) = make_double2(pacc[i][j][0], pacc[i][j][1]);
        }
    } else {
      int ti[5], tj[5], cnt;
      diag_tiles(wid, ti, tj, cnt);
      double pacc[5][2];
#pragma unroll
      for (int s = 0; s < 5; ++s) pacc[s][0] = pacc[s][1] = 0.0;
      for (int64_t n = r0; n < r1; ++n) {
        if (tid == 0) rv.refill(n - r0);
        const double* v = rv.row(n - r0);
        double acc[5][2];
        stage1_diag<QC>(sZI, v, qk, ti, tj, cnt, lane, acc);
#pragma unroll
        for (int s = 0; s < 5; ++s)
          if (s < cnt) {
            pacc[s][0] += exp_tab(acc[s][0], sT);
            pacc[s][1] += exp_tab(acc[s][1], sT);
          }
        __syncthreads();
      }
#pragma unroll
      for (int s = 0; s < 5; ++s)
        if (s < cnt) {
          const int m = 8 * ti[s] + g, mp = 8 * tj[s] + 2 * t;
          *reinterpret_cast<double2*>(out + m * 64 + mp) = make_double2(pacc[s][0], pacc[s][1]);
          if (ti[s] != tj[s]) {
            out[mp * 64 + m] = pacc[s][0];
            out[(mp + 1) * 64 + m] = pacc[s][1];
          }
        }
    }
  }
}

// =====================================================================================
// Backward.  grid = (R, G).  Outputs
//   lam [g][rc][Mp]  += row / column sums of L        (red.global.add, buffers pre-zeroed)
//   Wq  [g][rc][QC]  += (1 or 2) * sum_m Z'_mq T[m,q]
//   ACCp[cta][Mp][QC] += sum_n ws_nq (L_n Z')[m,q]     (CTA-private, plain RMW)
// FUSE: the kernel also produces the forward result - it has p = exp(E) in hand for every row, so
//   P2p[b][r][64][64] = sum_n p (the partial tiles k_psi2_fwd writes, same layout)
// is accumulated in a CTA-private shared tile (every element is owned by one thread: plain
// read-modify-write, 16 adds per thread and row) and the separate forward pass with its own
// stage 1 + exp is not needed.  Usable when the upstream gradients do not depend on the statistics
// of the same evaluation (the SVI bound, autoreg/inference/svi_vardtc.py:162-169).
// =====================================================================================
// OPT bits (A/B experiments, default 3): 1 = stage 2-J scales by ws AFTER the MMA (16 dense FMAs instead of
// 64 DMULs inside the DMMA loop), 2 = stage 1 scales its A fragments in batches of four k-steps
// DBG (RGP_DEBUG builds only; results are WRONG when set): ablation bits for timing experiments on the
// off-diagonal path - 1 no exp (p = x), 2 no lambda sums / flush, 4 no stage 2-I folds / W flush,
// 8 no epilogue at all (no exp, no C multiply, no L store), 16 no per-row barrier, 32 no stage 2-J,
// 64 no stage 2-I, 128 no stage 1
template <int QC, bool FUSE = false, int DBG = 0>
__global__ void __launch_bounds__(P2_THREADS, 1)
k_psi2_bwd(int64_t rc, int Mp, int nt, int nblocks, int qk, const double* __restrict__ Zt,
           const double* __restrict__ Ct, const double* __restrict__ wrow,
           const double* __restrict__ HP, double* __restrict__ lam, double* __restrict__ Wq,
           double* __restrict__ ACCp, int qoff, double* __restrict__ P2p = nullptr) {
  // qoff: first q column of this pass's stage-2 outputs (0, or 64 for the second pass of QC = 128);
  // lambda is flushed by the qoff == 0 pass only.
  using C = P2Cfg<QC>;
  constexpr int RS = C::RS, VB = C::VB, NJ = C::NJ, QS = C::QS;
  extern __shared__ __align__(16) double smem[];
  double* sZI = smem;
  double* sZJ = sZI + 64 * RS;
  double* sL = sZJ + 64 * RS;                     // 2 slots of 64*RSL
  double* sV = sL + 2 * 64 * RSL;                 // row-vector ring: 2 slots of VR rows (TMA, see RowVecStage)
  double* sWq = sV + 2 * C::VR * VB;              // [2][4][QS]
  double* sLr = sWq + 2 * 4 * QS;                 // [2][2][64]  row-sum partials (per wc)
  double* sLc = sLr + 2 * 2 * 64;                 // [2][4][64]  col-sum partials (per wr)
  double* sT = sLc + 2 * 4 * 64;                  // exp table, 256 entries
  double* sP = sT + 256 + 2;                      // FUSE: Psi2 tile of the block [64][RSL]
  RowVecStage<QC, C::VR> rv;
  rv.ring = sV;
  rv.mbar = reinterpret_cast<uint64_t*>(sT + 256);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int wr = wid >> 1, wc = wid & 1, g = lane >> 2, t = lane & 3;
  const int R = gridDim.x, G = gridDim.y;
  const int64_t per = (rc + R - 1) / R;
  const int64_t r0 = per * blockIdx.x, r1 = (r0 + per < rc) ? r0 + per : rc;
  const int cta = blockIdx.y * R + blockIdx.x;
  double* lamg = lam + (size_t)blockIdx.y * rc * Mp;
  double* Wqg = Wq + (size_t)blockIdx.y * rc * QC;
  double* accp = ACCp + (size_t)cta * Mp * QC;
  const int qbase = wc * (QS / 2);                // this warp's q columns in stage 2 (relative to qoff)
  exp_table_init(sT, tid);
  rv.init_barriers(tid);

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
    const double* cb = Ct + (size_t)b * 4096;     // C = s2^2 sym(dL_dpsi2) on this block
    rv.begin(r0, r1, wrow, hI, hJ, tid);
    if constexpr (FUSE)
      for (int i = tid; i < 64 * RSL; i += P2_THREADS) sP[i] = 0.0;
    double accI[2][NJ][2], accJ[2][NJ][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) accI[i][j][0] = accI[i][j][1] = accJ[i][j][0] = accJ[i][j][1] = 0.0;
    __syncthreads();

    // Wq partials of row n are written after the barrier of row n and flushed after the
    // barrier of row n+1 (two slots; hazard analysis in DESIGN.md "Row pipeline").
    auto flush_wq = [&](int64_t n) {
      const int s = (int)(n & 1);
      if (tid < QS) {
        const double* p = sWq + s * 4 * QS + tid;
        double v = p[0] + p[QS] + p[2 * QS] + p[3 * QS];
        red_add(Wqg + n * QC + qoff + tid, diag ? v : 2.0 * v);
      }
    };

    // stage 2-I: T = L ZJ ; accI += ws T ; Wq partial.  A = L (rows m, k = m'), B = ZJ (k = m', cols q)
    auto stage2I = [&](const double* v, const double* Lb, int s) {
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
      if constexpr (DBG & 4) {                     // keep T alive with the cheapest possible use
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            accI[i][j][0] = T[i][j][0];
            accI[i][j][1] = T[i][j][1];
          }
        return;
      }
      double wp[2 * NJ];                           // Wq partials, index c = 2 j + e
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
      for (int64_t n = r0; n < r1; ++n) {
        const int s = (int)(n & 1);
        const double* v = rv.row(n - r0);
        double* Lb = sL + s * 64 * RSL;
        {
          double acc[2][4][2];
          if constexpr (DBG & 128) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = v[i + j];
          } else stage1<QC>(sZI, sZJ, v, qk, wr, wc, lane, acc);
          if constexpr (DBG & 8) {                 // no epilogue: fold the exponents into one register so stage 1 stays
            double sink = 0.0;
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) sink += acc[i][j][0] + acc[i][j][1];
            if (sink == 1.2345e300) Lb[tid] = sink;
          } else {
          double rs[2] = {0.0, 0.0};
          double cs[8];                            // column partials, index c = 2 j + e
#pragma unroll
          for (int c = 0; c < 8; ++c) cs[c] = 0.0;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const double p0 = (DBG & 1) ? acc[i][j][0] : exp_tab(acc[i][j][0], sT);
              const double p1 = (DBG & 1) ? acc[i][j][1] : exp_tab(acc[i][j][1], sT);
              if constexpr (FUSE) {
                double2* pp = reinterpret_cast<double2*>(sP + (16 * wr + 8 * i + g) * RSL + 32 * wc + 8 * j + 2 * t);
                double2 o = *pp;
                o.x += p0;
                o.y += p1;
                *pp = o;
              }
              double l0 = creg[i][j][0] * p0;
              double l1 = creg[i][j][1] * p1;
              *reinterpret_cast<double2*>(Lb + (16 * wr + 8 * i + g) * RSL + 32 * wc + 8 * j + 2 * t) =
                  make_double2(l0, l1);
              if constexpr (!(DBG & 2)) {
                rs[i] += l0 + l1;
                cs[2 * j] += l0;
                cs[2 * j + 1] += l1;
              }
            }
          }
          // row sums: reduce over the 4 lanes of a quad (t); col sums: over the 8 quads (g)
          if constexpr (!(DBG & 2)) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 1);
            rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 2);
          }
          if (t == 0) {
            sLr[s * 128 + wc * 64 + 16 * wr + g] = rs[0];
            sLr[s * 128 + wc * 64 + 16 * wr + 8 + g] = rs[1];
          }
          {
            const double tot = reduce8_over_g(cs, lane);
            const int c = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            sLc[s * 256 + wr * 64 + 32 * wc + 8 * (c >> 1) + 2 * t + (c & 1)] = tot;
          }
          }
          }
        }
        if constexpr (!(DBG & 16)) __syncthreads();   // L tile + lambda partials of row n complete; row n-1 fully finished
        if (tid == 0) rv.refill(n - r0);              // ... so the slot of a batch that ended before row n is free
        if constexpr (DBG & 2) {
        } else if (qoff != 0) {
        } else if (tid >= 64 && tid < 128) {
          const int m = tid - 64;
          const double* p = sLr + s * 128 + m;
          red_add(lamg + n * Mp + I * 64 + m, p[0] + p[64]);
        } else if (tid >= 128 && tid < 192) {
          const int m = tid - 128;
          const double* p = sLc + s * 256 + m;
          red_add(lamg + n * Mp + J * 64 + m, p[0] + p[64] + p[128] + p[192]);
        }
        if constexpr (!(DBG & 4)) if (n > r0) flush_wq(n - 1);
        if constexpr (!(DBG & 64)) stage2I(v, Lb, s);
        // stage 2-J: accJ[m',q] += ws_q sum_m L[m,m'] ZI[m,q].  A = L^T, B = ZI; ws is applied AFTER the MMA
        // (16 FMAs per thread instead of 64 DMULs inside the DMMA loop: measured 1.4 % faster)
        if constexpr (DBG & 32) {
        } else {
          const double* pa = Lb + t * RSL + 16 * wr + g;
          const double* pb = sZI + t * RS + qoff + qbase + g;
          double TJ[2][NJ][2];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) TJ[i][j][0] = TJ[i][j][1] = 0.0;
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
            const double2 wq = *reinterpret_cast<const double2*>(v + qoff + qbase + 8 * j + 2 * t);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              accJ[i][j][0] = fma(wq.x, TJ[i][j][0], accJ[i][j][0]);
              accJ[i][j][1] = fma(wq.y, TJ[i][j][1], accJ[i][j][1]);
            }
          }
        }
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
      for (int64_t n = r0; n < r1; ++n) {
        const int s = (int)(n & 1);
        const double* v = rv.row(n - r0);
        double* Lb = sL + s * 64 * RSL;
        {
          double acc[5][2];
          stage1_diag<QC>(sZI, v, qk, ti, tj, cnt, lane, acc);
#pragma unroll
          for (int s5 = 0; s5 < 5; ++s5)
            if (s5 < cnt) {
              const double p0 = exp_tab(acc[s5][0], sT), p1 = exp_tab(acc[s5][1], sT);
              const double l0 = creg[s5][0] * p0;
              const double l1 = creg[s5][1] * p1;
              const int m = 8 * ti[s5] + g, mp = 8 * tj[s5] + 2 * t;
              if constexpr (FUSE) {
                double2* pp = reinterpret_cast<double2*>(sP + m * RSL + mp);
                double2 o = *pp;
                o.x += p0;
                o.y += p1;
                *pp = o;
              }
              *reinterpret_cast<double2*>(Lb + m * RSL + mp) = make_double2(l0, l1);
              if (ti[s5] != tj[s5]) {
                Lb[mp * RSL + m] = l0;
                Lb[(mp + 1) * RSL + m] = l1;
              }
            }
        }
        __syncthreads();
        if (tid == 0) rv.refill(n - r0);
        if (qoff == 0 && tid >= 64 && tid < 128) { // lambda_m = full row sum of the symmetric tile
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
        stage2I(v, Lb, s);
      }
    }
    __syncthreads();
    if (r1 > r0) flush_wq(r1 - 1);
    if constexpr (FUSE) {
      // every thread writes the tile elements it owns (the ones it accumulated), forward layout;
      // diagonal blocks mirror their upper tiles like k_psi2_fwd
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
    // flush the CTA-private dZ accumulators of this block
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
