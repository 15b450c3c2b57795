// mlp_kernels.cuh - MLP back-constraint of a hidden layer: free-run recurrence and its
// back-propagation (autoreg/layers.py:623-715 with the network of autoreg/mlp.py).
//
// The first Xwin latent means of a sequence are free parameters; every later mean is the output of a
// small tanh MLP applied to the Xwin means before it and the aligned control window.  The recurrence
// is sequential in time (N ~ 500 steps, a few thousand flops each) and the reference runs it as a
// Python loop around a theano function per step.  Here one CTA walks one sequence with the weights
// resident in shared memory: the cost is N x (a few barriers), sequences run in parallel.
//
//   packed parameters: for layer l (up = u[l] inputs, down = u[l+1] outputs): W[down][up] row-major, b[down]
//   hidden layers use tanh, the last one is linear (mlp.py:127, positive_obs = False)
//   seq[s] = {row_start, nrows, lat_start, lat_len, ctl_start, ctl_len} as in lag_kernels.cuh
#pragma once
#include "common.cuh"
#include "lag_kernels.cuh"

namespace rgp {
namespace mlp {

constexpr int MAXL = 8;
constexpr int THREADS = 128;

struct Shape {
  int nl;                 // layers
  int u[MAXL + 1];        // units: u[0] = Q inputs ... u[nl] = Dx outputs
  int woff[MAXL];         // offset of W_l in the packed parameter vector (b_l follows W_l)
  int hoff[MAXL];         // offset of hidden layer l's outputs in a row of `acts`
  int nparams, nhid, maxu;
};

inline bool make_shape(int nl, const int* units, Shape* s) {
  if (nl < 1 || nl > MAXL) return false;
  s->nl = nl;
  int off = 0, hid = 0, mx = 0;
  for (int l = 0; l <= nl; ++l) {
    if (units[l] <= 0) return false;
    s->u[l] = units[l];
    mx = units[l] > mx ? units[l] : mx;
  }
  for (int l = 0; l < nl; ++l) {
    s->woff[l] = off;
    off += units[l + 1] * units[l] + units[l + 1];
    s->hoff[l] = hid;
    if (l < nl - 1) hid += units[l + 1];
  }
  s->nparams = off;
  s->nhid = hid;
  s->maxu = mx;
  return true;
}

inline size_t fwd_smem(const Shape& s) { return sizeof(double) * ((size_t)s.nparams + 4 * s.maxu) + MAXL * 48; }
inline size_t bwd_smem(const Shape& s, int Xwin, int Dx) {
  return sizeof(double) * (2 * (size_t)s.nparams + s.u[0] + s.nhid + 2 * s.maxu + (size_t)(Xwin + 1) * Dx + 1) + MAXL * 48;
}

// shared copy of the parameters with every W transposed (WT[i][j] = W[j][i]): thread j of a layer
// then reads consecutive words
__device__ __forceinline__ void load_params_T(const Shape& sh, const double* __restrict__ params, double* sP) {
  for (int l = 0; l < sh.nl; ++l) {
    const int up = sh.u[l], down = sh.u[l + 1], o = sh.woff[l];
    for (int idx = threadIdx.x; idx < down * up; idx += blockDim.x) {
      const int j = idx / up, i = idx - j * up;
      sP[o + i * down + j] = params[o + idx];
    }
    for (int j = threadIdx.x; j < down; j += blockDim.x) sP[o + down * up + j] = params[o + down * up + j];
  }
}

// tanh through the library costs a ~150-instruction dependent chain per hidden layer and step - in a
// recurrence that is pure latency.  tanh(x) = sign(x) * (-em / (em + 2)), em = expm1(-2 |x|): for
// 2|x| < 0.34 em comes from the exp polynomial of common.cuh without its constant term (no
// cancellation), otherwise from exp_neg() - 1.  ~30 dependent operations; max relative error 7e-15
// against the library over [-20, 20] and down to 1e-12 (absolute 3e-16).
RGP_DEVINL double tanh_fast(double x) {
  const double y = -2.0 * fabs(x);
  double em;
  if (y > -0.34) {
    double p = 2.76263572414472227e-07;
    p = fma(p, y, 2.76401807962098502e-06);
    p = fma(p, y, 2.48015043469976862e-05);
    p = fma(p, y, 1.98411702704400671e-04);
    p = fma(p, y, 1.38888889324885988e-03);
    p = fma(p, y, 8.33333338566778249e-03);
    p = fma(p, y, 4.16666666665731419e-02);
    p = fma(p, y, 1.66666666665544055e-01);
    p = fma(p, y, 5.00000000000000555e-01);
    p = fma(p, y, 1.00000000000000666e+00);
    em = p * y;
  } else {
    em = exp_neg(y) - 1.0;
  }
  return copysign(-em / (em + 2.0), x);
}

// out[j] = sum_i WT[i][j] * in[i] (+ bias[j]) for j < down, with the dot product of one output split
// over P adjacent lanes (P = 4, 2 or 1, so that down * P fits the block when it can) and two
// accumulators per lane: the dependent-FMA chain of a step is what bounds the recurrence, not the
// flop count.  Every thread of the block must call it; no barrier inside.
__host__ __device__ inline int log2_split_for(int rows) { return rows * 4 <= THREADS ? 2 : (rows * 2 <= THREADS ? 1 : 0); }

// Per-layer constants, computed once per kernel (the step loop is pure latency: no integer division,
// no indexed access to the by-value Shape inside it).
struct LayerK {
  int up, down, off;        // widths, offset of W in the packed vector
  int lpF, itF;             // forward mat-vec: log2 of the lanes per output, trips over the outputs
  int lpT, itT;             // transposed mat-vec (back-propagation): outputs = the layer's inputs
  int hin, hout;            // offsets of the layer's input / output activations in sAct (backward), -1 = none
};

__device__ __forceinline__ void layer_constants(const Shape& sh, LayerK* lk) {
  if ((int)threadIdx.x < sh.nl) {
    const int l = threadIdx.x;
    LayerK k;
    k.up = sh.u[l];
    k.down = sh.u[l + 1];
    k.off = sh.woff[l];
    k.lpF = log2_split_for(k.down);
    k.itF = (k.down + (THREADS >> k.lpF) - 1) / (THREADS >> k.lpF);
    k.lpT = log2_split_for(k.up);
    k.itT = (k.up + (THREADS >> k.lpT) - 1) / (THREADS >> k.lpT);
    k.hin = l == 0 ? 0 : sh.u[0] + sh.hoff[l - 1];
    k.hout = l < sh.nl - 1 ? sh.u[0] + sh.hoff[l] : -1;
    lk[l] = k;
  }
}

template <bool TRANSPOSED>
__device__ __forceinline__ void matvec(const double* __restrict__ Wm, int rows, int cols, const double* __restrict__ bias,
                                       const double* __restrict__ in, double* __restrict__ out, int lp, int trips) {
  // TRANSPOSED = false: out[j] = bias[j] + sum_i Wm[i * rows + j] * in[i]   (j < rows outputs, i < cols inputs)
  // TRANSPOSED = true : out[i] = sum_j Wm[i * cols + j] * in[j]             (i < rows outputs, j < cols inputs)
  // lp = log2(lanes per output), trips = ceil(rows / (THREADS >> lp)); every thread runs every trip
  const int P = 1 << lp, per = THREADS >> lp;
  const int part = threadIdx.x & (P - 1);
  int o = threadIdx.x >> lp;
  for (int trip = 0; trip < trips; ++trip, o += per) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (o < rows) {
      const double* w = TRANSPOSED ? Wm + o * cols : Wm + o;
      const int ws = TRANSPOSED ? 1 : rows;
      int i = part;
      for (; i + 3 * P < cols; i += 4 * P) {
        a0 = fma(w[i * ws], in[i], a0);
        a1 = fma(w[(i + P) * ws], in[i + P], a1);
        a2 = fma(w[(i + 2 * P) * ws], in[i + 2 * P], a2);
        a3 = fma(w[(i + 3 * P) * ws], in[i + 3 * P], a3);
      }
      for (; i < cols; i += P) a0 = fma(w[i * ws], in[i], a0);
    }
    double a = (a0 + a1) + (a2 + a3);
    if (lp >= 1) a += __shfl_xor_sync(0xffffffffu, a, 1);
    if (lp >= 2) a += __shfl_xor_sync(0xffffffffu, a, 2);
    if (o < rows && part == 0) out[o] = bias ? a + bias[o] : a;
  }
}

// grid = sequences.  The input vector of a step lives in shared memory: [Xwin latent means | Uwin
// controls]; after a step the window shifts by one latent step and takes the new mean, so the
// recurrence never waits for a global round trip.  The control part of the next step is fetched
// while the layers of the current one run.
__global__ void __launch_bounds__(THREADS)
k_freerun(Shape sh, const int64_t* __restrict__ seq, int Xwin, int Dx, int Uwin, int Du,
          const double* __restrict__ params, double* __restrict__ lat, const double* __restrict__ ctl,
          double* __restrict__ acts, int dbg) {
  // dbg (timing experiments only, results wrong): 1 no mat-vec, 2 no tanh, 4 no global stores, 8 no control prefetch
  extern __shared__ __align__(16) double smem[];
  double* sP = smem;
  double* win0 = sP + sh.nparams;                 // two input vectors (ping-pong), each maxu long
  double* win1 = win0 + sh.maxu;
  double* bufA = win1 + sh.maxu;
  double* bufB = bufA + sh.maxu;
  LayerK* lk = reinterpret_cast<LayerK*>(bufB + sh.maxu);
  layer_constants(sh, lk);
  const int nl = sh.nl, nhid = sh.nhid;
  const int s = blockIdx.x;
  const int Qx = Xwin * Dx, Qu = Uwin * Du;
  const int64_t row0 = seq[s * lag::DESC + 0], N = seq[s * lag::DESC + 1];
  const int64_t lat0 = seq[s * lag::DESC + 2], ctl0 = seq[s * lag::DESC + 4];
  load_params_T(sh, params, sP);
  for (int i = threadIdx.x; i < Qx + Qu; i += blockDim.x)
    win0[i] = i < Qx ? lat[lat0 * Dx + i] : ctl[ctl0 * Du + (i - Qx)];
  __syncthreads();
  for (int64_t n = 0; n < N; ++n) {
    double* in = (n & 1) ? win1 : win0;
    double* nxt = (n & 1) ? win0 : win1;
    double cnext = 0.0;                           // this thread's control input of the next step (Q <= THREADS, host-checked)
    if (!(dbg & 8) && (int)threadIdx.x >= Qx && (int)threadIdx.x < Qx + Qu && n + 1 < N)
      cnext = ctl[(ctl0 + n + 1) * Du + ((int)threadIdx.x - Qx)];
    const double* x = in;
    double* out = bufA;
    for (int l = 0; l < nl; ++l) {
      const LayerK k = lk[l];
      const int up = k.up, down = k.down, o = k.off;
      if (!(dbg & 1)) matvec<false>(sP + o, down, up, sP + o + down * up, x, out, k.lpF, k.itF);
      __syncthreads();
      if (l < nl - 1) {
        for (int j = threadIdx.x; j < down; j += blockDim.x) {
          const double a = (dbg & 2) ? out[j] : tanh_fast(out[j]);
          out[j] = a;
          if (!(dbg & 4)) acts[(row0 + n) * nhid + (k.hout - sh.u[0]) + j] = a;
        }
        __syncthreads();
      }
      x = out;
      out = (out == bufA) ? bufB : bufA;
    }
    // x = the new mean (Dx values): store it, shift the window, append the prefetched controls
    for (int i = threadIdx.x; i < Qx + Qu; i += blockDim.x) {
      double v;
      if (i < Qx - Dx) v = in[i + Dx];
      else if (i < Qx) v = x[i - (Qx - Dx)];
      else v = cnext;
      nxt[i] = v;
    }
    if (!(dbg & 4))
      for (int j = threadIdx.x; j < Dx; j += blockDim.x) lat[(lat0 + Xwin + n) * Dx + j] = x[j];
    __syncthreads();
  }
}

// lat_g: dL/d mean of every step on entry (read only for rows >= Xwin).  A shared ring holds the
// back-propagated contributions to the Xwin + 1 newest-but-unfinished rows, so the step loop reads
// nothing it wrote itself from global memory.  On exit rows < Xwin of lat_g hold the initial-mean
// gradients (objective part + back-propagated part); other rows are unchanged.
// pgrad[s][nparams]: parameter gradients of sequence s in the packed (untransposed) layout.
__global__ void __launch_bounds__(THREADS)
k_freerun_bwd(Shape sh, const int64_t* __restrict__ seq, int Xwin, int Dx, int Uwin, int Du,
              const double* __restrict__ params, const double* __restrict__ lat, const double* __restrict__ ctl,
              const double* __restrict__ acts, double* __restrict__ lat_g, double* __restrict__ ctl_g,
              double* __restrict__ pgrad) {
  extern __shared__ __align__(16) double smem[];
  double* sP = smem;                              // transposed weights (WT[i][j]) + biases
  double* sG = sP + sh.nparams;                   // gradients, packed untransposed layout
  double* sAct = sG + sh.nparams;                 // [Q inputs | hidden outputs] of the step
  double* d0 = sAct + sh.u[0] + sh.nhid;
  double* d1 = d0 + sh.maxu;
  double* ring = d1 + sh.maxu;                    // [(Xwin + 1) * Dx] pending contributions, row t at slot t % (Xwin + 1)
  LayerK* lk = reinterpret_cast<LayerK*>(ring + (Xwin + 1) * Dx + ((Xwin + 1) * Dx & 1));
  layer_constants(sh, lk);
  const int nl = sh.nl, nhid = sh.nhid;
  const int s = blockIdx.x;
  const int Qx = Xwin * Dx, Q = sh.u[0], RW = Xwin + 1;
  const int64_t row0 = seq[s * lag::DESC + 0], N = seq[s * lag::DESC + 1];
  const int64_t lat0 = seq[s * lag::DESC + 2], ctl0 = seq[s * lag::DESC + 4];
  load_params_T(sh, params, sP);
  for (int i = threadIdx.x; i < sh.nparams; i += blockDim.x) sG[i] = 0.0;
  for (int i = threadIdx.x; i < RW * Dx; i += blockDim.x) ring[i] = 0.0;
  __syncthreads();
  int slot_out = (int)((Xwin + N) % RW);
  const int my_k = (int)threadIdx.x / Dx, my_j = (int)threadIdx.x - my_k * Dx;    // latent row / dim of input element tid
  for (int64_t n = N - 1; n >= 0; --n) {
    // inputs, activations and the objective gradient of this step: nothing here was written by this kernel
    for (int i = threadIdx.x; i < Q; i += blockDim.x)
      sAct[i] = i < Qx ? lat[(lat0 + n) * Dx + i] : ctl[(ctl0 + n) * Du + (i - Qx)];
    for (int i = threadIdx.x; i < nhid; i += blockDim.x) sAct[Q + i] = acts[(row0 + n) * nhid + i];
    slot_out = slot_out == 0 ? RW - 1 : slot_out - 1;           // slot of row Xwin + n, walking down with n
    for (int j = threadIdx.x; j < Dx; j += blockDim.x) {
      d0[j] = lat_g[(lat0 + Xwin + n) * Dx + j] + ring[slot_out * Dx + j];
      ring[slot_out * Dx + j] = 0.0;              // the slot is reused by row n - 1
    }
    __syncthreads();
    double* dl = d0;
    double* dn = d1;
    for (int l = nl - 1; l >= 0; --l) {
      const LayerK k = lk[l];
      const int up = k.up, down = k.down, o = k.off;
      const double* in = sAct + k.hin;
      if (k.hout >= 0) {
        const double* out = sAct + k.hout;
        for (int j = threadIdx.x; j < down; j += blockDim.x) dl[j] *= 1.0 - out[j] * out[j];
        __syncthreads();
      }
      // outer product dl in^T -> W gradient: thread owns columns i = tid % up of consecutive rows, no division
      // in the loop (i and j advance by the fixed stride THREADS = qd * up + rm)
      {
        const int qd = THREADS / up, rm = THREADS - qd * up;
        int j = threadIdx.x / up, i = threadIdx.x - j * up;
        for (int idx = threadIdx.x; idx < down * up; idx += THREADS) {
          sG[o + idx] = fma(dl[j], in[i], sG[o + idx]);
          j += qd;
          i += rm;
          if (i >= up) {
            i -= up;
            ++j;
          }
        }
      }
      for (int j = threadIdx.x; j < down; j += blockDim.x) sG[o + down * up + j] += dl[j];
      matvec<true>(sP + o, up, down, nullptr, dl, dn, k.lpT, k.itT);     // dn[i] = sum_j WT[i][j] dl[j]
      __syncthreads();
      double* tmp = dl;
      dl = dn;
      dn = tmp;
    }
    {                                             // Q <= THREADS (host-checked): one input element per thread
      const int i = threadIdx.x;
      if (i < Qx) {                               // latent row n + my_k lives in slot (slot_out + 1 + my_k) mod RW
        int sl = slot_out + 1 + my_k;
        sl -= sl >= RW ? RW : 0;
        ring[sl * Dx + my_j] += dl[i];
      } else if (i < Q && ctl_g) {
        ctl_g[(ctl0 + n) * Du + (i - Qx)] += dl[i];
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < Qx; i += blockDim.x) {
    const int k = i / Dx, j = i - k * Dx;
    lat_g[(lat0 + k) * Dx + j] += ring[(k % RW) * Dx + j];
  }
  for (int i = threadIdx.x; i < sh.nparams; i += blockDim.x) pgrad[(size_t)s * sh.nparams + i] = sG[i];
}

}  // namespace mlp
}  // namespace rgp
