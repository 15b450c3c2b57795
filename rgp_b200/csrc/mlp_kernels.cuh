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

inline size_t fwd_smem(const Shape& s) { return sizeof(double) * ((size_t)s.nparams + 2 * s.maxu); }
inline size_t bwd_smem(const Shape& s) {
  return sizeof(double) * (2 * (size_t)s.nparams + s.u[0] + s.nhid + 2 * s.maxu);
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

__device__ __forceinline__ void gather_input(int n, int Xwin, int Dx, int Uwin, int Du, int64_t lat0, int64_t ctl0,
                                             const double* __restrict__ lat, const double* __restrict__ ctl,
                                             double* in) {
  const int Qx = Xwin * Dx, Q = Qx + Uwin * Du;
  for (int i = threadIdx.x; i < Q; i += blockDim.x)
    in[i] = i < Qx ? lat[(lat0 + n) * Dx + i] : ctl[(ctl0 + n) * Du + (i - Qx)];
}

__global__ void __launch_bounds__(THREADS)
k_freerun(Shape sh, const int64_t* __restrict__ seq, int Xwin, int Dx, int Uwin, int Du,
          const double* __restrict__ params, double* __restrict__ lat, const double* __restrict__ ctl,
          double* __restrict__ acts) {
  extern __shared__ __align__(16) double smem[];
  double* sP = smem;
  double* bufA = sP + sh.nparams;
  double* bufB = bufA + sh.maxu;
  const int s = blockIdx.x;
  const int64_t row0 = seq[s * lag::DESC + 0], N = seq[s * lag::DESC + 1];
  const int64_t lat0 = seq[s * lag::DESC + 2], ctl0 = seq[s * lag::DESC + 4];
  load_params_T(sh, params, sP);
  __syncthreads();
  for (int64_t n = 0; n < N; ++n) {
    double* in = bufA;
    double* out = bufB;
    gather_input((int)n, Xwin, Dx, Uwin, Du, lat0, ctl0, lat, ctl, in);
    __syncthreads();
    for (int l = 0; l < sh.nl; ++l) {
      const int up = sh.u[l], down = sh.u[l + 1], o = sh.woff[l];
      const bool hidden = l < sh.nl - 1;
      for (int j = threadIdx.x; j < down; j += blockDim.x) {
        double a = sP[o + down * up + j];
        for (int i = 0; i < up; ++i) a = fma(sP[o + i * down + j], in[i], a);
        a = hidden ? tanh(a) : a;
        out[j] = a;
        if (hidden) acts[(row0 + n) * sh.nhid + sh.hoff[l] + j] = a;
      }
      __syncthreads();
      double* tmp = in;
      in = out;
      out = tmp;
    }
    for (int j = threadIdx.x; j < Dx; j += blockDim.x) lat[(lat0 + Xwin + n) * Dx + j] = in[j];
    __syncthreads();              // the new mean is an input of the next steps
  }
}

// lat_g: dL/d mean of every step on entry; each step adds its input gradient onto the Xwin rows it read
// (newest step first), so on exit the first Xwin rows of a sequence are the initial-mean gradients.
// pgrad[s][nparams]: parameter gradients of sequence s in the packed (untransposed) layout.
__global__ void __launch_bounds__(THREADS)
k_freerun_bwd(Shape sh, const int64_t* __restrict__ seq, int Xwin, int Dx, int Uwin, int Du,
              const double* __restrict__ params, const double* __restrict__ lat, const double* __restrict__ ctl,
              const double* __restrict__ acts, double* __restrict__ lat_g, double* __restrict__ ctl_g,
              double* __restrict__ pgrad) {
  extern __shared__ __align__(16) double smem[];
  double* sP = smem;
  double* sG = sP + sh.nparams;
  double* sAct = sG + sh.nparams;                 // [Q inputs | hidden outputs]
  double* d0 = sAct + sh.u[0] + sh.nhid;
  double* d1 = d0 + sh.maxu;
  const int s = blockIdx.x;
  const int Qx = Xwin * Dx, Q = sh.u[0];
  const int64_t row0 = seq[s * lag::DESC + 0], N = seq[s * lag::DESC + 1];
  const int64_t lat0 = seq[s * lag::DESC + 2], ctl0 = seq[s * lag::DESC + 4];
  load_params_T(sh, params, sP);
  for (int i = threadIdx.x; i < sh.nparams; i += blockDim.x) sG[i] = 0.0;
  __syncthreads();
  for (int64_t n = N - 1; n >= 0; --n) {
    gather_input((int)n, Xwin, Dx, Uwin, Du, lat0, ctl0, lat, ctl, sAct);
    for (int i = threadIdx.x; i < sh.nhid; i += blockDim.x) sAct[Q + i] = acts[(row0 + n) * sh.nhid + i];
    for (int j = threadIdx.x; j < Dx; j += blockDim.x) d0[j] = lat_g[(lat0 + Xwin + n) * Dx + j];
    __syncthreads();
    double* dl = d0;
    double* dn = d1;
    for (int l = sh.nl - 1; l >= 0; --l) {
      const int up = sh.u[l], down = sh.u[l + 1], o = sh.woff[l];
      const double* in = l == 0 ? sAct : sAct + Q + sh.hoff[l - 1];
      if (l < sh.nl - 1) {
        const double* out = sAct + Q + sh.hoff[l];
        for (int j = threadIdx.x; j < down; j += blockDim.x) dl[j] *= 1.0 - out[j] * out[j];
        __syncthreads();
      }
      for (int idx = threadIdx.x; idx < down * up; idx += blockDim.x) {
        const int j = idx / up, i = idx - j * up;
        sG[o + idx] = fma(dl[j], in[i], sG[o + idx]);
      }
      for (int j = threadIdx.x; j < down; j += blockDim.x) sG[o + down * up + j] += dl[j];
      for (int i = threadIdx.x; i < up; i += blockDim.x) {
        double a = 0.0;
        for (int j = 0; j < down; ++j) a = fma(sP[o + i * down + j], dl[j], a);
        dn[i] = a;
      }
      __syncthreads();
      double* tmp = dl;
      dl = dn;
      dn = tmp;
    }
    for (int i = threadIdx.x; i < Q; i += blockDim.x) {
      if (i < Qx) lat_g[(lat0 + n) * Dx + i] += dl[i];
      else if (ctl_g) ctl_g[(ctl0 + n) * Du + (i - Qx)] += dl[i];
    }
    __syncthreads();              // older steps read the rows just updated
  }
  for (int i = threadIdx.x; i < sh.nparams; i += blockDim.x) pgrad[(size_t)s * sh.nparams + i] = sG[i];
}

}  // namespace mlp
}  // namespace rgp
