// psi2_bwds.cuh - "strip" variant of the Psi2 backward kernel: split-phase hand-off of the L tile.
//
// k_psi2_bwd / k_psi2_bwd16 hand the 64 x 64 tile L = C . exp(E) from the warps that produced it to
// the warps that consume it through shared memory behind ONE CTA barrier per row with no slack:
// every warp waits for the slowest one, 1-2 K cycles out of ~16 K (profiles/SUMMARY_r01.md, sec. 4).
// Here a warp owns an 8-row STRIP of the block:
//
//   stage 1    E[strip, 0..63]  = H_m + H_m' + sum_q (ws_q Z'_mq) Z'_m'q      8 tiles x QC/4 DMMA
//   epilogue   L = C . exp(E)   (registers, accumulator layout) ; row sums -> lambda_I
//              strip of L -> shared tile (slot n & 1)
//   stage 2-I  T[strip, q]      = sum_m' L[m,m'] Z'_m'q : A = the warp's OWN L, moved from the
//              accI += ws . T ;  W_q partial -> shared     accumulator to the A layout by 2 shuffles
//   ARRIVE on FULL[n & 1], then WAIT for it - the other warps only have to have finished their
//              stage 1 + epilogue + stage 2-I of this row, i.e. be no more than a third of a row
//              behind in their own stage 2-J of the previous row
//   stage 2-J  T'[m' in strip of J, q] = sum_m L[m,m'] Z'_mq   (A = L^T from the shared tile)
//              accJ += ws . T' ; column sums of L (the A fragments it loads anyway) -> lambda_J
//   ARRIVE on FREE[n & 1] ; the write of row n + 2 into the slot WAITs on it (a whole row later)
//
// so producers and consumers of a tile are coupled by mbarriers with slack instead of meeting at a
// barrier, stage 2-I never touches the shared tile, and the pre-weighted tile of the 16-warp kernel
// is gone (stage 1 scales its one A fragment per k-step, stage 2 weights after the MMA).
// 8 warps x <= 255 registers, 1 CTA / SM.
//
// Diagonal blocks: warp r computes the tiles (r, tt >= r); the tile on the diagonal is halved after
// the lambda sums.  Stage 2-I then covers the tiles to the right of the diagonal, stage 2-J (column
// strip r, tiles (rr <= r, r)) the ones above it, both see half of the diagonal tile, and together
// they give the full symmetric product for the rows of strip r; 2 x sum_m Z'_mq T_mq is the full
// quadratic form.  Warps w and w + 4 share a scheduler and own strips r and 7 - r: 17 - r tile
// passes each, 27 per scheduler.
//
// Row vectors (ws[QC], H_J[64], H_I[strip]) are staged per warp with cp.async, double buffered; the
// W_q partials of the 8 strips go through a small shared buffer that rides on the same FULL
// barrier, so global atomics stay at 24 per warp and row.
#pragma once
#include "psi2_kernels.cuh"

namespace rgp {
namespace fast {

constexpr int SW = 8;   // warps per CTA

template <int QC>
struct P2CfgS {
  static constexpr int RS = QC + 4;
  static constexpr int QS = QC > 64 ? 64 : QC;
  static constexpr int NU = QS / 8;             // 8-wide q tiles of stage 2
  static constexpr int VS = QC + 64 + 8;        // per-warp row-vector slot: ws | H_J | H_I strip
  // shared-memory map (offsets in doubles)
  static constexpr int ZI = 0;
  static constexpr int ZJ = 64 * RS;
  static constexpr int CC = 2 * 64 * RS;                  // C tile           [64][RSL]
  static constexpr int LL = CC + 64 * RSL;                // L tile, 2 slots  [2][64][RSL]
  static constexpr int WW = LL + 2 * 64 * RSL;            // W partials       [2][SW][QS]
  static constexpr int VV = WW + 2 * SW * QS;             // row vectors      [SW][2][VS]
  static constexpr int TT = VV + SW * 2 * VS;             // exp table        [256]
  static constexpr int MB = TT + 256;                     // 4 mbarriers: FULL[2], FREE[2]
  static constexpr int SMEM = (MB + 4) * 8;
};

RGP_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
RGP_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

RGP_DEVINL void mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
RGP_DEVINL void mbar_arrive(unsigned bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(bar) : "memory");
}
RGP_DEVINL void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n"
      "WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra DONE_%=;\n bra WAIT_%=;\n"
      "DONE_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}

// Shared-memory reads of the row loop are volatile asm on 32-bit shared addresses with immediate
// offsets: together with the volatile DMMA they keep the order written here (fragments of step
// s + 1 are requested, then the eight DMMAs of step s issue), which bounds the live registers, and
// one 32-bit base replaces a dozen 64-bit pointers.
RGP_DEVINL unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int OFF>
RGP_DEVINL double lds_off(unsigned base) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(base), "n"(OFF * 8));
  return v;
}
template <int OFF>
RGP_DEVINL double2 lds2_off(unsigned base) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(base), "n"(OFF * 8));
  return v;
}
RGP_DEVINL double lds(unsigned a) { return lds_off<0>(a); }
RGP_DEVINL void sts2(unsigned a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
RGP_DEVINL void sts(unsigned a, double x) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(x) : "memory"); }

// exp_tab (common.cuh) with the table addressed through a 32-bit shared address
RGP_DEVINL double exp_tab_s(double x, unsigned tab) {
  const double INV = 369.32993046757463;
  const double STEP = 2.7076061740622863e-03;
  const double MAGIC = 6755399441055744.0;
  double kd = fma(x, INV, MAGIC);
  int n = __double2loint(kd);
  double nf = kd - MAGIC;
  double r = fma(nf, -STEP, x);
  double p = fma(r, 4.16666666666666666667e-02, 1.66666666666666666667e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  double res = lds(tab + 8u * (unsigned)(n & 255)) * p;
  int hi = __double2hiint(res) + ((n >> 8) << 20);
  res = __hiloint2double(hi, __double2loint(res));
  return ((unsigned)__double2hiint(x) > 0xC0862000u) ? 0.0 : res;
}

template <int STRIDE, int I, int N>
RGP_DEVINL void load8(unsigned base, double (&b)[8]) {        // b[i] = base[i * STRIDE], i < N
  if constexpr (I < N) {
    b[I] = lds_off<I * STRIDE>(base);
    load8<STRIDE, I + 1, N>(base, b);
  }
}

// b[i] = 16-byte load at base + (i * STRIDE doubles), i < N
template <int STRIDE, int I, int N>
RGP_DEVINL void load_pairs(unsigned base, double2 (&b)[8]) {
  if constexpr (I < N) {
    b[I] = lds2_off<I * STRIDE>(base);
    load_pairs<STRIDE, I + 1, N>(base, b);
  }
}

// Every contraction index and every output column of the three products is free to be permuted, and
// the permutations below are chosen so that each lane's operands for TWO DMMAs sit in one aligned
// 16-byte word: one LDS.128 per two DMMAs instead of one LDS.64 per DMMA (same wavefronts, half the
// load instructions - with two warps per scheduler the issue slots and load latencies of a warp
// are what bound it, profiles/SUMMARY_r01.md).
//   stage 1:  k-steps (2 s, 2 s + 1) contract q = 8 s + 2 t and 8 s + 2 t + 1   (t = lane % 4)
//   stage 2:  q tiles (2 w, 2 w + 1): column n of tile 2 w + e' is q = 16 w + 2 n + e', so lane g
//             loads Z'[.][16 w + 2 g .. + 1], and the accumulator pair (tile 2 w + e', register e)
//             of lane t holds q = 16 w + 4 t + 2 e + e'
// The scalar work that depends on freshly loaded or shuffled data is placed AFTER the DMMAs of the
// step before it, so the in-order warp never waits for it in front of independent DMMAs.

// stage 2-I, k-steps S..15 (tile S / 2, half S % 2).  x0, x1: tile registers of step S as shuffled
// from lane 4 g + 2 J + (t >> 1); the A fragment is register t & 1 of that.
template <int S>
RGP_DEVINL void s2i_shuffle(const double (&L)[8][2], int lane, double& x0, double& x1) {
  constexpr int TT = S >> 1, J = S & 1;
  const int src = (lane & ~3) | (2 * J + ((lane & 3) >> 1));
  x0 = __shfl_sync(0xffffffffu, L[TT][0], src);
  x1 = __shfl_sync(0xffffffffu, L[TT][1], src);
}

template <int QC, int S, bool DIAG>
RGP_DEVINL void s2i_steps(unsigned bz, const double (&L)[8][2], int lane, int t0, double x0, double x1,
                          double (&T)[P2CfgS<QC>::NU][2], double2 (&bc)[8], double2 (&bn)[8]) {
  constexpr int RS = P2CfgS<QC>::RS, NP = P2CfgS<QC>::NU / 2;
  if constexpr (S < 16) {
    double y0 = 0.0, y1 = 0.0;
    if constexpr (S + 1 < 16) {
      load_pairs<16, 0, NP>(bz + 8u * ((8 * ((S + 1) >> 1) + 4 * ((S + 1) & 1)) * RS), bn);
      s2i_shuffle<S + 1>(L, lane, y0, y1);
    }
    if (!DIAG || (S >> 1) >= t0) {
      const double a = (lane & 1) ? x1 : x0;
#pragma unroll
      for (int w = 0; w < NP; ++w) {
        dmma(T[2 * w][0], T[2 * w][1], a, bc[w].x);
        dmma(T[2 * w + 1][0], T[2 * w + 1][1], a, bc[w].y);
      }
    }
    s2i_steps<QC, S + 1, DIAG>(bz, L, lane, t0, y0, y1, T, bn, bc);
  }
}

// stage 2-J, k-steps S..15 over m = 4 S + t: A = L^T from the shared tile (a = L[4 S + t][8 r + g]),
// B = Z'_I pairs.  Tiles above the diagonal one (RR < r) also feed the column sums.
template <int QC, int S, bool DIAG>
RGP_DEVINL void s2j_steps(unsigned la, unsigned bz, int r, double a, double& csum,
                          double (&T)[P2CfgS<QC>::NU][2], double2 (&bc)[8], double2 (&bn)[8]) {
  constexpr int RS = P2CfgS<QC>::RS, NP = P2CfgS<QC>::NU / 2;
  if constexpr (S < 16) {
    double an = 0.0;
    if constexpr (S + 1 < 16) {
      load_pairs<16, 0, NP>(bz + 8u * (4 * (S + 1) * RS), bn);
      an = lds_off<4 * (S + 1) * RSL>(la);
    }
    constexpr int RR = S >> 1;                      // tile row (strip) the k-step reads from
    if (!DIAG || RR <= r) {
#pragma unroll
      for (int w = 0; w < NP; ++w) {
        dmma(T[2 * w][0], T[2 * w][1], a, bc[w].x);
        dmma(T[2 * w + 1][0], T[2 * w + 1][1], a, bc[w].y);
      }
      if (!DIAG || RR < r) csum += a;               // the diagonal tile's columns are its rows
    }
    s2j_steps<QC, S + 1, DIAG>(la, bz, r, an, csum, T, bn, bc);
  }
}

// acc[2 w + e'][e] += wv[q] * T[2 w + e'][e] with q = 16 w + 4 t + 2 e + e' (pw points at q = 4 t)
template <int QC>
RGP_DEVINL void fold_ws(unsigned pw, const double (&T)[P2CfgS<QC>::NU][2], double (&acc)[P2CfgS<QC>::NU][2]) {
  constexpr int NP = P2CfgS<QC>::NU / 2;
#pragma unroll
  for (int w = 0; w < NP; ++w) {
    const double2 w01 = lds2_off<0>(pw + 128u * w);
    const double2 w23 = lds2_off<2>(pw + 128u * w);
    acc[2 * w][0] = fma(w01.x, T[2 * w][0], acc[2 * w][0]);
    acc[2 * w + 1][0] = fma(w01.y, T[2 * w + 1][0], acc[2 * w + 1][0]);
    acc[2 * w][1] = fma(w23.x, T[2 * w][1], acc[2 * w][1]);
    acc[2 * w + 1][1] = fma(w23.y, T[2 * w + 1][1], acc[2 * w + 1][1]);
  }
}

// One row of one block for one warp.  sb = shared address of the CTA's dynamic smem, vb = shared
// address of this row's vector slot, slot = it & 1 (L tile / W buffer) with its mbarrier parities.
template <int QC, bool DIAG>
RGP_DEVINL void strip_row(unsigned sb, unsigned vb, int qk, int r, int lane, int wid, int qoff, int slot,
                          unsigned full_parity, bool wait_free, unsigned free_parity,
                          double* __restrict__ lamI, double* __restrict__ lamJ, double* __restrict__ wq_out,
                          double (&accI)[P2CfgS<QC>::NU][2], double (&accJ)[P2CfgS<QC>::NU][2]) {
  using C = P2CfgS<QC>;
  constexpr int RS = C::RS, NU = C::NU, NP = NU / 2, QS = C::QS;
  static_assert(NU == 8 || NU == 4, "strip kernel is instantiated for QC = 32, 64");
  const int g = lane >> 2, t = lane & 3;
  const int t0 = DIAG ? r : 0;
  const unsigned ws = vb, HJ = vb + 8u * QC, HI = vb + 8u * (QC + 64);
  const unsigned zi_g = sb + 8u * (C::ZI + (8 * r + g) * RS);    // Z'_I row 8 r + g
  const unsigned lt = sb + 8u * (C::LL + slot * 64 * RSL);       // this row's L tile
  const unsigned wb = sb + 8u * (C::WW + slot * SW * QS);        // this row's W partials [SW][QS]
  const unsigned full = sb + 8u * C::MB + 8u * slot, freeb = sb + 8u * C::MB + 16u + 8u * slot;

  // ---- stage 1: exponent of the strip ------------------------------------------------------
  double L[8][2];
  {
    const double hi = lds(HI + 8u * g);
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) {
      const double2 hj = lds2_off<0>(HJ + 8u * (8 * tt + 2 * t));
      L[tt][0] = hi + hj.x;
      L[tt][1] = hi + hj.y;
    }
    // (stage 1 keeps 8-byte fragment loads: with this row stride the 16-byte pattern of the paired
    // k-steps has 2-way bank conflicts, measured slower)
    unsigned pa = zi_g + 8u * t;
    unsigned pw = ws + 8u * t;
    unsigned pb = sb + 8u * (C::ZJ + g * RS + t);
    double bc[8], bn[8];
    load8<8 * RS, 0, 8>(pb, bc);
    double za = lds(pa), wa = lds(pw);
    for (int k0 = 0; k0 < qk; k0 += 8) {           // two k-steps per trip (qk is a multiple of 4)
      // operands of k0 + 4 are requested before the DMMAs of k0 issue, those of k0 + 8 before k0 + 4
      // (pad columns of Z' are zero; over-reads stay inside the padded tile / vector slot)
      load8<8 * RS, 0, 8>(pb + 32, bn);
      const double zb = lds_off<4>(pa), wb2 = lds_off<4>(pw);
      {
        const double a0 = za * wa;
#pragma unroll
        for (int tt = 0; tt < 8; ++tt)
          if (!DIAG || tt >= t0) dmma(L[tt][0], L[tt][1], a0, bc[tt]);
      }
      load8<8 * RS, 0, 8>(pb + 64, bc);
      za = lds_off<8>(pa);
      wa = lds_off<8>(pw);
      if (k0 + 4 < qk) {
        const double a1 = zb * wb2;
#pragma unroll
        for (int tt = 0; tt < 8; ++tt)
          if (!DIAG || tt >= t0) dmma(L[tt][0], L[tt][1], a1, bn[tt]);
      }
      pa += 64;
      pw += 64;
      pb += 64;
    }
  }

  // ---- epilogue: L = C exp(E), row sums, strip -> shared tile --------------------------------
  {
    const unsigned pc = sb + 8u * (C::CC + (8 * r + g) * RSL + 2 * t);
    const unsigned tab = sb + 8u * C::TT;
    double r0 = 0.0, r1 = 0.0;
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) {
      if (!DIAG || tt >= t0) {
        const double2 c2 = lds2_off<0>(pc + 64u * tt);
        const double l0 = c2.x * exp_tab_s(L[tt][0], tab);
        const double l1 = c2.y * exp_tab_s(L[tt][1], tab);
        r0 += l0;
        r1 += l1;
        const double h = (DIAG && tt == t0) ? 0.5 : 1.0;   // halve the diagonal tile (header comment)
        L[tt][0] = h * l0;
        L[tt][1] = h * l1;
      }
    }
    if (wait_free) mbar_wait(freeb, free_parity);   // stage 2-J of row n - 2 has left the slot
    const unsigned pl = lt + 8u * ((8 * r + g) * RSL + 2 * t);
#pragma unroll
    for (int tt = 0; tt < 8; ++tt)
      if (!DIAG || tt >= t0) sts2(pl + 64u * tt, L[tt][0], L[tt][1]);
    double rs = r0 + r1;
    rs += __shfl_xor_sync(0xffffffffu, rs, 1);
    rs += __shfl_xor_sync(0xffffffffu, rs, 2);
    if (qoff == 0 && t == 0) red_add(lamI + 8 * r + g, rs);
  }

  // ---- stage 2-I: T = L Z'_J from registers, accI += ws T, W partial -> shared ------------------
  double T[NU][2];
  {
#pragma unroll
    for (int u = 0; u < NU; ++u) T[u][0] = T[u][1] = 0.0;
    const unsigned bz = sb + 8u * (C::ZJ + t * RS + qoff + 2 * g);
    double2 bc[8], bn[8];
    load_pairs<16, 0, NP>(bz, bc);
    double x0, x1;
    s2i_shuffle<0>(L, lane, x0, x1);
    s2i_steps<QC, 0, DIAG>(bz, L, lane, t0, x0, x1, T, bc, bn);
    fold_ws<QC>(ws + 8u * (qoff + 4 * t), T, accI);
    // W partial: wp[2 u + e] = Z'_I[8 r + g][q(u, e)] T[u][e], summed over the strip's rows (lanes g)
    double wp0[8], wp1[8];
    const unsigned pz = zi_g + 8u * (qoff + 4 * t);
#pragma unroll
    for (int w = 0; w < NP; ++w) {
      const double2 z01 = lds2_off<0>(pz + 128u * w);
      const double2 z23 = lds2_off<2>(pz + 128u * w);
      double (&wp)[8] = (w < 2) ? wp0 : wp1;
      const int o = 4 * (w & 1);                    // tiles 2 w, 2 w + 1 -> entries 2 u + e of their half
      wp[o + 0] = z01.x * T[2 * w][0];
      wp[o + 1] = z23.x * T[2 * w][1];
      wp[o + 2] = z01.y * T[2 * w + 1][0];
      wp[o + 3] = z23.y * T[2 * w + 1][1];
    }
    // reduce8_over_g leaves entry c8 = 2 (u % 4) + e in this lane: q = 16 w + 4 t + 2 e + e'
    const int c8 = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    const int uu = c8 >> 1, e = c8 & 1;
    const int col = 16 * (uu >> 1) + 4 * t + 2 * e + (uu & 1);
    sts(wb + 8u * (wid * QS + col), reduce8_over_g(wp0, lane));
    if constexpr (NU == 8) sts(wb + 8u * (wid * QS + 32 + col), reduce8_over_g(wp1, lane));
  }
  __syncwarp();
  if (lane == 0) mbar_arrive(full);                 // L strip and W partials of this warp are in place
  mbar_wait(full, full_parity);                     // ... and everybody else's

  // ---- stage 2-J: T' = L^T Z'_I for column strip r of J; accJ += ws T'; column sums -------------
  {
#pragma unroll
    for (int u = 0; u < NU; ++u) T[u][0] = T[u][1] = 0.0;
    const unsigned la = lt + 8u * (t * RSL + 8 * r + g);
    const unsigned bz = sb + 8u * (C::ZI + t * RS + qoff + 2 * g);
    double csum = 0.0;
    double2 bc[8], bn[8];
    load_pairs<16, 0, NP>(bz, bc);
    s2j_steps<QC, 0, DIAG>(la, bz, r, lds(la), csum, T, bc, bn);
    fold_ws<QC>(ws + 8u * (qoff + 4 * t), T, accJ);
    csum += __shfl_xor_sync(0xffffffffu, csum, 1);
    csum += __shfl_xor_sync(0xffffffffu, csum, 2);
    if (qoff == 0 && t == 0 && (!DIAG || r > 0)) red_add(lamJ + 8 * r + g, csum);
    // W_q of the whole block: this warp adds up the eight strips for its QS / 8 columns
    if (lane < QS / SW) {
      const unsigned p = wb + 8u * (wid * (QS / SW) + lane);
      double w = 0.0;
#pragma unroll
      for (int s = 0; s < SW; ++s) w += lds(p + 8u * (s * QS));
      red_add(wq_out + qoff + wid * (QS / SW) + lane, 2.0 * w);
    }
  }
  __syncwarp();
  if (lane == 0) mbar_arrive(freeb);                // this warp is done with the slot
}

// grid = (R row ranges, G block groups), 256 threads; same arguments as k_psi2_bwd.
template <int QC>
__global__ void __launch_bounds__(SW * 32, 1)
k_psi2_bwds(int64_t rc, int Mp, int nt, int nblocks, int qk, const double* __restrict__ Zt,
            const double* __restrict__ Ct, const double* __restrict__ wrow,
            const double* __restrict__ HP, double* __restrict__ lam, double* __restrict__ Wq,
            double* __restrict__ ACCp, int qoff) {
  using C = P2CfgS<QC>;
  constexpr int RS = C::RS, NU = C::NU, VS = C::VS;
  extern __shared__ __align__(16) double smem[];
  const unsigned sb = smem_addr(smem);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int r = wid < 4 ? wid : 11 - wid;         // strip; warps w and w + 4 share a scheduler
  const int R = gridDim.x, G = gridDim.y;
  const int64_t per = (rc + R - 1) / R;
  const int64_t r0 = per * blockIdx.x, r1 = (r0 + per < rc) ? r0 + per : rc;
  const int cta = blockIdx.y * R + blockIdx.x;
  double* lamg = lam + (size_t)blockIdx.y * rc * Mp;
  double* Wqg = Wq + (size_t)blockIdx.y * rc * QC;
  double* accp = ACCp + (size_t)cta * Mp * QC;
  const unsigned myV = sb + 8u * (C::VV + wid * 2 * VS);
  exp_table_init(smem + C::TT, tid);
  if (tid < 4) mbar_init(sb + 8u * C::MB + 8u * tid, SW);
  unsigned it = 0;                                // rows this CTA has processed (all blocks): slot = it & 1

  int curI = -1, curJ = -1;
  for (int b = blockIdx.y; b < nblocks; b += G) {
    int I, J;
    block_ij(b, nt, I, J);
    const bool diag = (I == J);
    __syncthreads();                              // every warp is done with the previous tiles (and mbarrier init)
    if (I != curI) copy_tile<64 * RS>(smem + C::ZI, Zt + (size_t)I * 64 * RS, tid);
    if (J != curJ) copy_tile<64 * RS>(smem + C::ZJ, Zt + (size_t)J * 64 * RS, tid);
    {
      const double2* src = reinterpret_cast<const double2*>(Ct + (size_t)b * 4096);
      for (int i = tid; i < 2048; i += SW * 32) {
        const int m = i >> 5, c2 = i & 31;
        *reinterpret_cast<double2*>(smem + C::CC + m * RSL + 2 * c2) = src[i];
      }
    }
    curI = I;
    curJ = J;
    __syncthreads();
    const double* hI = HP + (size_t)I * rc * 64 + 8 * r;
    const double* hJ = HP + (size_t)J * rc * 64;

    double accI[NU][2], accJ[NU][2];
#pragma unroll
    for (int u = 0; u < NU; ++u) accI[u][0] = accI[u][1] = accJ[u][0] = accJ[u][1] = 0.0;

    // row vectors of row n -> slot n & 1 of this warp's staging area (16-byte cp.async chunks)
    auto stage = [&](int64_t n) {
      const unsigned dst = myV + 8u * ((unsigned)(n & 1) * VS);
      for (int c = lane; c < VS / 2; c += 32) {
        const double* src = c < QC / 2 ? wrow + n * QC + 2 * c
                            : (c < QC / 2 + 32 ? hJ + n * 64 + 2 * (c - QC / 2) : hI + n * 64 + 2 * (c - QC / 2 - 32));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst + 16u * c), "l"(src) : "memory");
      }
      cp_async_commit();
    };
    if (r0 < r1) stage(r0);
    for (int64_t n = r0; n < r1; ++n, ++it) {
      if (n + 1 < r1) {
        stage(n + 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
      const unsigned vb = myV + 8u * ((unsigned)(n & 1) * VS);
      double* lamI = lamg + n * Mp + I * 64;
      double* lamJ = lamg + n * Mp + J * 64;
      // mbarrier phases: slot s = it & 1 is used for the (it >> 1)-th time
      const int slot = (int)(it & 1u);
      const unsigned use = it >> 1;
      if (diag)
        strip_row<QC, true>(sb, vb, qk, r, lane, wid, qoff, slot, use & 1u, use > 0, (use - 1u) & 1u, lamI, lamJ,
                            Wqg + n * QC, accI, accJ);
      else
        strip_row<QC, false>(sb, vb, qk, r, lane, wid, qoff, slot, use & 1u, use > 0, (use - 1u) & 1u, lamI, lamJ,
                             Wqg + n * QC, accI, accJ);
      __syncwarp();                               // vector slot n & 1 is rewritten by the prefetch of row n + 2
    }

    // flush the dZ partials of this block: rows of strip r in I (accI) and in J (accJ); no two warps
    // of the CTA touch the same element, the partial is CTA-private.  Lane t of row g holds
    // q = 16 w + 4 t + {0, 1, 2, 3} in (tile 2 w, reg 0), (2 w + 1, 0), (2 w, 1), (2 w + 1, 1).
#pragma unroll
    for (int w = 0; w < NU / 2; ++w) {
      const int q = qoff + 16 * w + 4 * t;
      double2* pI = reinterpret_cast<double2*>(accp + (size_t)(I * 64 + 8 * r + g) * QC + q);
      double2 o0 = pI[0], o1 = pI[1];
      o0.x += accI[2 * w][0];
      o0.y += accI[2 * w + 1][0];
      o1.x += accI[2 * w][1];
      o1.y += accI[2 * w + 1][1];
      pI[0] = o0;
      pI[1] = o1;
      double2* pJ = reinterpret_cast<double2*>(accp + (size_t)(J * 64 + 8 * r + g) * QC + q);
      double2 x0 = pJ[0], x1 = pJ[1];
      x0.x += accJ[2 * w][0];
      x0.y += accJ[2 * w + 1][0];
      x1.x += accJ[2 * w][1];
      x1.y += accJ[2 * w + 1][1];
      pJ[0] = x0;
      pJ[1] = x1;
    }
  }
}

}  // namespace fast
}  // namespace rgp
