// psi2_small.cuh - Psi2 forward / backward for SMALL inducing sets (M <= 112, Q <= 47): the shapes of the
// reference's own models (M = 50 ... 100, Q = 10 ... 40; autoreg/benchmark/tasks.py:141-177,
// examples/walk_run_2_alex.py:364-393, svi_experiments/rgp_experiments.py).
//
// The 64 x 64 block kernels of psi2_kernels.cuh pay for padding there: M = 100 is computed as 128 (136 8x8 tiles
// of the pair matrix instead of 91), Q = 20 as 32 stage-2 columns, lambda / W go through 3 block passes of
// red.global.add, and 8 warps (2 per scheduler) cannot hide the DMMA issue latency of the short k loops.  Here ONE
// CTA (16 warps; 8 warps and two CTAs per SM at M <= 64) holds the whole problem of a row:
//
//   Z' [Mp16][Qp + 4] in shared memory (Mp16 = M rounded up to 16, Qp = 8 (Q / 8 + 1): at least one spare column,
//   the last one holds ONES so that T[:, Qp - 1] = L 1 = lambda comes out of the stage-2 MMAs for free) and the UPPER
//   TRIANGLE of the symmetric L_n as packed 16 x 16 supertiles (row stride 20), double-buffered by row parity;
//   stage 1   E = H_m + H_m' + sum_q (ws_q Z'_mq) Z'_m'q on the supertiles of the upper triangle (8x8 tiles that lie
//             entirely in the padding are skipped), p = exp(E), Psi2 += p (registers), L = C p -> shared;
//   stage 2   T = L Z' as jobs (16-row strip x Qp columns x a range of k-steps); a job reads L[strip][k]
//             from the supertile (strip, k) directly or from (k, strip) transposed - both fragment patterns are
//             conflict free at stride 20, the four k-steps of a supertile column are unrolled with immediate
//             offsets - and folds  acc[m,q] += ws_q T[m,q],  W_q += sum_m Z'_mq T[m,q],  lambda_m = T[m, Qp - 1];
//   per row   ONE CTA barrier (L_n complete); the partials of row n are combined after the barrier of row n + 1 and
//             lambda_n, W_n are written once with plain stores, in a fixed order (deterministic, no atomics).
//
// Which warp computes which supertiles and jobs is a small table made on the host (SmallSched, fast_path.cuh): it
// balances the FP64-pipe load per SM sub-partition (warp w runs on sub-partition w % 4).  Where two CTAs fit in shared
// memory (M <= 64) the CTA has 8 warps instead of 16 and two CTAs share an SM: their barriers are independent, so one
// CTA's stage 2 overlaps the other's stage 1.
// The per-row vectors (ws[Qp], H[Mp16]) arrive by TMA bulk copies into a two-slot ring (SmallRowStage), as in the
// block kernels.  The forward-only kernel needs a barrier only when a ring slot is recycled (every 4 rows).
#pragma once
#include <type_traits>

#include "psi2_kernels.cuh"

namespace rgp {
namespace fast {

constexpr int PS_THREADS = 512;
constexpr int PS_WARPS = PS_THREADS / 32;
constexpr int PS_MS_MAX = 7;      // 16-row super rows: M <= 112
constexpr int PS_S1 = 4;          // supertile slots per warp in the work table (the product kernels use 2; 4 in experiment builds)
constexpr int PS_JOBS = 32;       // job list length
// jobs per row that get a partial slot (wide Q and the single-buffer variant: one per warp)
__host__ __device__ constexpr int PS_JOBS_OF(int QT, int nbuf = 2) { return (QT > 3 || nbuf == 1) ? 16 : 32; }
// rows per TMA batch (the single-buffer variant has two CTAs per SM and half the ring)
__host__ __device__ constexpr int PS_VR_OF(int nbuf) { return nbuf == 1 ? 2 : 4; }
constexpr int PS_ST = 320;        // doubles per packed supertile: 16 rows x stride 20
constexpr int PS_LAM = 16 * PS_MS_MAX;   // lambda partials: [parity][k slot <= 4][PS_LAM]

struct SmallSched {
  signed char ns[PS_WARPS];            // supertiles of warp w ...
  signed char su[PS_WARPS][PS_S1];     // ... and their indices in the row-major enumeration of the upper triangle
  signed char nj[PS_WARPS];            // jobs of warp w ...
  signed char jw[PS_WARPS][2];         // ... and their indices into the job list
  signed char njobs, kslots;           // job list length; ACCp slots per CTA (largest number of k ranges of a strip)
  signed char jsp[PS_JOBS], jkb[PS_JOBS], jke[PS_JOBS], jslot[PS_JOBS];   // strip, k-steps [kb, ke), k slot
};

// shared-memory size in doubles (host and device agree through this one function)
__host__ __device__ constexpr int small_smem_doubles(int Ms, int QT, bool bwd, int nbuf = 2) {
  const int Mp16 = 16 * Ms, Qp = 8 * QT;
  int d = Mp16 * (Qp + 4) + 2 * PS_VR_OF(nbuf) * (Qp + Mp16) + 258;
  if (bwd) d += nbuf * PS_JOBS_OF(QT, nbuf) * Qp + nbuf * 4 * PS_LAM + nbuf * (Ms * (Ms + 1) / 2) * PS_ST;
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
  int vr = 4;                 // rows per batch
  int QC = 0, Qp = 0, Mp16 = 0, VB = 0;   // the ring slot of a row is [ws (Qp) | H (Mp16)]; min(Qp, QC) of ws are copied

  RGP_DEVINL void init_barriers(int tid) {
    if (tid == 0) {
      mbar_init(&mbar[0], 1);
      mbar_init(&mbar[1], 1);
      mbar_fence_init();
    }
  }
  RGP_DEVINL void begin(int64_t r0_, int64_t r1_, int tid) {
    r0 = r0_; r1 = r1_;
    nb = r1 > r0 ? (r1 - r0 + vr - 1) / vr : 0;
    next = 0;
    if (tid == 0 && nb > 0) issue();
  }
  RGP_DEVINL void issue() {
    const int64_t n0 = r0 + next * vr;
    const int rows = (int)((r1 - n0 < vr) ? r1 - n0 : vr);
    const int slot = (int)(next & 1);
    double* dst = ring + slot * vr * VB;
    const int h0 = Mp16 < 64 ? Mp16 : 64, qc = Qp < QC ? Qp : QC;
    mbar_expect_tx(&mbar[slot], (uint32_t)(rows * (qc + Mp16) * 8));
    for (int w = 0; w < rows; ++w) {
      bulk_g2s(dst + w * VB, wrow + (n0 + w) * QC, qc * 8, &mbar[slot]);
      bulk_g2s(dst + w * VB + Qp, hp + (n0 + w) * 64, h0 * 8, &mbar[slot]);
      if (Mp16 > 64) bulk_g2s(dst + w * VB + Qp + 64, hp + htile + (n0 + w) * 64, (Mp16 - 64) * 8, &mbar[slot]);
    }
    ++next;
  }
  // issuing thread, when every thread has finished all rows < r0 + idx
  RGP_DEVINL void refill(int64_t idx) {
    while (next < nb && next <= idx / vr + 1 && (next - 1) * vr <= idx) issue();
  }
  RGP_DEVINL const double* row(int64_t idx) {
    const int slot = (int)((idx / vr) & 1);
    if (idx % vr == 0) {
      mbar_wait(&mbar[slot], (phase_bits >> slot) & 1u);
      phase_bits ^= 1u << slot;
    }
    return ring + slot * vr * VB + (idx % vr) * VB;
  }
};

// index of supertile (lo, hi), lo <= hi, in the row-major enumeration of the upper triangle
RGP_DEVINL int st_index(int lo, int hi, int Ms) { return lo * Ms - lo * (lo - 1) / 2 + (hi - lo); }

// MODE 0: forward only (Psi2 partials), 1: backward only, 2: backward + Psi2 partials (fused SVI pass).
// Outputs (strides of the block path, so the small GEMMs and combiners downstream are shared):
//   lam [rc][Mp]   Wq [rc][QC]                 complete per row, plain stores
//   ACCp[cta * kslots + slot][Mp][QC]          sum_n ws (L_n Z') over this CTA's rows and one k range (plain stores
//                                              where a job exists; the buffer is zeroed before the launch)
//   P2s [cta][Mp16][Mp16]                      sum_n p over this CTA's rows (MODE 0 / 2), full symmetric
// S1: supertile slots per warp.  NBUF: buffers of L - 2 = double-buffered by row parity, one barrier per row.
// The product uses S1 = 2, NBUF = 2.  (S1 = 4 with 8-warp CTAs, and NBUF = 1 = a single buffer with a second barrier per
// row so that two 8-warp CTAs fit at M = 81 ... 112, are measured negative results: forward 4.33 -> 4.63 ms, backward
// 10.30 -> 12.10 ms at M = 100, Q = 20; they are instantiated in -DRGP_DEBUG experiment builds only.)
template <int QT, int MODE, int JMAX, int S1 = 2, int NBUF = 2>
__global__ void __launch_bounds__(PS_THREADS, 1)
k_psi2_small(int64_t rc, int M, int Q, int Mp, int Ms, int nt, int qk, int QC, int RSz,
             const __grid_constant__ SmallSched sc,
             const double* __restrict__ Zt, const double* __restrict__ Ct, const double* __restrict__ wrow,
             const double* __restrict__ HP, double* __restrict__ lam, double* __restrict__ Wq,
             double* __restrict__ ACCp, double* __restrict__ P2s) {
  constexpr int Qp = 8 * QT, RS = Qp + 4;
  constexpr int NH = QT > 3 ? 2 : 1, QH = QT / NH;     // stage-2 column passes and 8-wide tiles per pass
  static_assert(QT % NH == 0 && 2 * QH <= 8, "stage-2 halves");
  constexpr bool BWD = MODE != 0, FWD = MODE != 1;
  const int Mp16 = 16 * Ms, VB = Qp + Mp16, M8 = (M + 7) & ~7, NS = Ms * (Ms + 1) / 2;
  extern __shared__ __align__(16) double smem[];
  double* sZ = smem;                         // [Mp16][RS]
  constexpr int VR = PS_VR_OF(NBUF), JOBS = PS_JOBS_OF(QT, NBUF);
  double* sV = sZ + Mp16 * RS;               // row-vector ring: 2 slots x VR rows x VB
  double* sT = sV + 2 * VR * VB;             // exp table (256) + 2 mbarriers
  double* sW = sT + 258;                     // [2][jobs][Qp]      W partials per job, by row parity
  double* sLam = sW + NBUF * JOBS * Qp;      // [NBUF][4][PS_LAM]  lambda partials per k slot
  double* sL = sLam + NBUF * 4 * PS_LAM;     // [NBUF][NS][PS_ST]  packed supertiles of L_n, by row parity

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
  rv.QC = QC; rv.Qp = Qp; rv.Mp16 = Mp16; rv.VB = VB; rv.vr = VR;
  rv.init_barriers(tid);
  for (int idx = tid; idx < Mp16 * Qp; idx += blockDim.x) {
    const int m = idx / Qp, c = idx - m * Qp;
    sZ[m * RS + c] = c < Q ? Zt[(size_t)m * RSz + c] : ((c == Qp - 1 && m < M) ? 1.0 : 0.0);   // last column: ones (lambda)
  }
  if constexpr (BWD)
    for (int idx = tid; idx < NBUF * 4 * PS_LAM; idx += blockDim.x) sLam[idx] = 0.0;   // (slot, strip) pairs without a job stay 0

  // this warp's supertiles
  const int ns = sc.ns[wid];
  int su[S1], si[S1], sj[S1];
#pragma unroll
  for (int s = 0; s < S1; ++s) {
    su[s] = s < ns ? sc.su[wid][s] : 0;
    int i = 0, rem = su[s];
    while (rem >= Ms - i) { rem -= Ms - i; ++i; }
    si[s] = i;
    sj[s] = i + rem;
  }
  double creg[S1][2][2][2], pacc[S1][2][2][2];
#pragma unroll
  for (int s = 0; s < S1; ++s)
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
  // this warp's stage-2 jobs
  const int nj = BWD ? sc.nj[wid] : 0;
  double accZ[JMAX][2][QT][2];
#pragma unroll
  for (int jj = 0; jj < JMAX; ++jj)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < QT; ++j) accZ[jj][i][j][0] = accZ[jj][i][j][1] = 0.0;

  // lambda_n, W_n of a finished row: fixed-order sums of the partials
  const int njobs = sc.njobs, kslots = sc.kslots;
  auto flush_row = [&](int64_t n) {
    const int par = NBUF == 2 ? (int)(n & 1) : 0;
    if (tid < Qp) {
      if (tid < QC) {
        const double* p = sW + par * JOBS * Qp + tid;
        double s = 0.0;
        for (int jb = 0; jb < njobs; ++jb) s += p[jb * Qp];
        Wq[n * QC + tid] = s;
      }
    } else if (tid >= 64 && tid < 64 + Mp16) {
      const int m = tid - 64;
      const double* p = sLam + par * 4 * PS_LAM + m;
      double s = p[0];
      for (int k = 1; k < kslots; ++k) s += p[k * PS_LAM];
      lam[n * Mp + m] = s;
    }
  };

  rv.begin(r0, r1, tid);
  if constexpr (!BWD)
    if (tid == 0) rv.refill(0);                       // forward only: both slots in flight from the start
  __syncthreads();

  for (int64_t n = r0; n < r1; ++n) {
    const double* v = rv.row(n - r0);
    const double* H = v + Qp;
    double* Lb = sL + (NBUF == 2 ? (int)(n & 1) : 0) * NS * PS_ST;
    // ------------------------------------------------------------------ stage 1 + exp (+ L)
#pragma unroll
    for (int s = 0; s < S1; ++s) {
      if (s < ns) {
        double acc[2][2][2];
        const int mi = 16 * si[s] + g, mj = 16 * sj[s] + 2 * t;
        const bool vi1 = 16 * si[s] + 8 < M8, vj1 = 16 * sj[s] + 8 < M8;     // second tile row / column not all padding
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
        if (vi1 && vj1) {
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
        } else {
          for (int k0 = 0; k0 < qk; k0 += 4) {
            const double wv = v[k0 + t];
            const double a0 = pa[k0] * wv, a1 = pa[8 * RS + k0] * wv;
            const double b0 = pb[k0], b1 = pb[8 * RS + k0];
            dmma(acc[0][0][0], acc[0][0][1], a0, b0);
            if (vj1) dmma(acc[0][1][0], acc[0][1][1], a0, b1);
            if (vi1) dmma(acc[1][0][0], acc[1][0][1], a1, b0);
          }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if ((i == 0 || vi1) && (j == 0 || vj1)) {
              const double p0 = exp_tab(acc[i][j][0], sT), p1 = exp_tab(acc[i][j][1], sT);
              if constexpr (FWD) {
                pacc[s][i][j][0] += p0;
                pacc[s][i][j][1] += p1;
              }
              if constexpr (BWD)
                *reinterpret_cast<double2*>(Lb + su[s] * PS_ST + (8 * i + g) * 20 + 8 * j + 2 * t) =
                    make_double2(creg[s][i][j][0] * p0, creg[s][i][j][1] * p1);
            }
          }
      }
    }
    if constexpr (!BWD) {
      // the ring slot of this batch is recycled once every thread has consumed its last row
      const int64_t idx = n - r0;
      if (idx % VR == VR - 1 && n + 1 < r1) {
        __syncthreads();
        if (tid == 0) rv.refill(idx + 1);
      }
    } else {
      __syncthreads();                                // L_n complete; every thread has finished row n - 1
      if (tid == 0) rv.refill(n - r0);
      if constexpr (NBUF == 2)
        if (n > r0) flush_row(n - 1);
      // ---------------------------------------------------------------- stage 2: T = L Z' by jobs
      const int par = NBUF == 2 ? (int)(n & 1) : 0;
#pragma unroll
      for (int jj = 0; jj < JMAX; ++jj) {
        if (jj < nj) {
          const int jb = sc.jw[wid][jj];
          const int sp = sc.jsp[jb], kb = sc.jkb[jb], ke = sc.jke[jb];   // k-steps [kb, ke): k-step 4 sk + kk = columns 16 sk + 4 kk ...
          const bool two = 16 * sp + 8 < M8;          // the strip's second 8 rows are not all padding
          double* myW = sW + (par * JOBS + jb) * Qp;
          // Q > 23: the stage-2 columns go in two halves of QH tiles (registers hold one half of T at a time)
#pragma unroll
          for (int h = 0; h < NH; ++h) {
          double T[2][QH][2];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < QH; ++j) T[i][j][0] = T[i][j][1] = 0.0;
          const double* pbz = sZ + t * RS + 8 * QH * h + g;
          // one supertile column (4 k-steps, [klo, khi) of them): T[i][j] += L[strip rows 8 i + g][k] Z'[k][8 j ...].
          // KST / IOFF: fragment strides of the packed supertile - (4, 160) read as stored, (80, 8) read transposed;
          // TWO: the strip's second 8 rows exist.  (A generic lambda, so T stays in registers.)
          auto column = [&](auto kst_c, auto ioff_c, auto two_c, const double* pa, const double* pb, int klo, int khi) {
            constexpr int KST = decltype(kst_c)::value, IOFF = decltype(ioff_c)::value;
            constexpr bool TWO = decltype(two_c)::value;
            auto kstep = [&](int kk) {
              const double a0 = pa[kk * KST];
              double a1 = 0.0;
              if constexpr (TWO) a1 = pa[kk * KST + IOFF];
              double bq[QH];
#pragma unroll
              for (int j = 0; j < QH; ++j) bq[j] = pb[kk * 4 * RS + 8 * j];
#pragma unroll
              for (int j = 0; j < QH; ++j) dmma(T[0][j][0], T[0][j][1], a0, bq[j]);
              if constexpr (TWO) {
#pragma unroll
                for (int j = 0; j < QH; ++j) dmma(T[1][j][0], T[1][j][1], a1, bq[j]);
              }
            };
            if (klo <= 0 && khi >= 4) {            // a whole column: straight-line code, loads of the next k-step overlap the MMAs
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) kstep(kk);
            } else {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                if (kk >= klo && kk < khi) kstep(kk);
            }
          };
          using I4 = std::integral_constant<int, 4>;
          using I8 = std::integral_constant<int, 8>;
          using I80 = std::integral_constant<int, 80>;
          using I160 = std::integral_constant<int, 160>;
          for (int sk = kb >> 2; 4 * sk < ke; ++sk) {
            // L[strip sp][k in supertile column sk]: from supertile (sp, sk) as stored, or from (sk, sp) transposed
            const bool tr = sk < sp;
            const double* pa = Lb + (tr ? st_index(sk, sp, Ms) : st_index(sp, sk, Ms)) * PS_ST + (tr ? t * 20 + g : g * 20 + t);
            const double* pb = pbz + sk * 16 * RS;
            const int klo = kb - 4 * sk, khi = ke - 4 * sk;     // (k-steps of padding columns are not in any job)
            if (tr) {
              if (two) column(I80{}, I8{}, std::true_type{}, pa, pb, klo, khi);
              else column(I80{}, I8{}, std::false_type{}, pa, pb, klo, khi);
            } else {
              if (two) column(I4{}, I160{}, std::true_type{}, pa, pb, klo, khi);
              else column(I4{}, I160{}, std::false_type{}, pa, pb, klo, khi);
            }
          }
          // folds: acc += ws T ; W partial = sum over this strip's rows of Z' T
          double wp[2 * QH];
#pragma unroll
          for (int j = 0; j < QH; ++j) {
            const int q = 8 * (QH * h + j) + 2 * t;
            const double2 wq = *reinterpret_cast<const double2*>(v + q);
            double w0 = 0.0, w1 = 0.0;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const double2 z = *reinterpret_cast<const double2*>(sZ + (16 * sp + 8 * i + g) * RS + q);
              accZ[jj][i][QH * h + j][0] = fma(wq.x, T[i][j][0], accZ[jj][i][QH * h + j][0]);
              accZ[jj][i][QH * h + j][1] = fma(wq.y, T[i][j][1], accZ[jj][i][QH * h + j][1]);
              w0 = fma(z.x, T[i][j][0], w0);
              w1 = fma(z.y, T[i][j][1], w1);
            }
            wp[2 * j] = w0;
            wp[2 * j + 1] = w1;
          }
          {
            double v8[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) v8[c] = (c < 2 * QH) ? wp[(c < 2 * QH) ? c : 0] : 0.0;
            const double tot = reduce8_over_g(v8, lane);
            const int cc = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            if (cc < 2 * QH) myW[8 * (QH * h + (cc >> 1)) + 2 * t + (cc & 1)] = tot;
          }
          // lambda of the strip's rows over this job's k range = the last column of T (the ones column of Z')
          if (h == NH - 1 && t == 3) {
            const double l0 = T[0][QH - 1][1], l1 = T[1][QH - 1][1];
            double* pl = sLam + (par * 4 + sc.jslot[jb]) * PS_LAM + 16 * sp + g;
            pl[0] = l0;
            pl[8] = l1;
          }
          }
        }
      }
      if constexpr (NBUF == 1) {
        __syncthreads();                              // L_n consumed (stage 1 of row n + 1 may overwrite it), partials complete
        flush_row(n);
      }
    }
  }

  if constexpr (BWD) {
    if constexpr (NBUF == 2) {
      __syncthreads();                                // partials of the last row complete
      if (r1 > r0) flush_row(r1 - 1);
    }
#pragma unroll
    for (int jj = 0; jj < JMAX; ++jj) {
      if (jj < nj) {
        const int jb = sc.jw[wid][jj];
        const int sp = sc.jsp[jb];
        double* out = ACCp + (size_t)(blockIdx.x * sc.kslots + sc.jslot[jb]) * Mp * QC;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < QT; ++j)
            if (8 * j + 2 * t < QC)
              *reinterpret_cast<double2*>(out + (size_t)(16 * sp + 8 * i + g) * QC + 8 * j + 2 * t) =
                  make_double2(accZ[jj][i][j][0], accZ[jj][i][j][1]);
      }
    }
  }
  if constexpr (FWD) {
    double* out = P2s + (size_t)blockIdx.x * Mp16 * Mp16;
#pragma unroll
    for (int s = 0; s < S1; ++s)
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
