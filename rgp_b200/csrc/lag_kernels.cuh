// lag_kernels.cuh - lag-window gather / scatter-add.
//
// The layer input matrix X (N x Q) is a strided view of the latent sequences: row n of
// sequence s is [x_n ... x_{n+Xwin-1}] (lag major, then latent dim) followed by the Uwin
// control / upper-layer lags (autoreg/layers.py:510-526 via get_conv_1D, autoreg/util.py:6-12),
// sequences stacked row-wise (layers.py:481-489).  The backward direction adds each X-row
// gradient back onto the Xwin (Uwin) latent steps it was built from
// (update_latent_gradients, layers.py:552-571: a Python double loop in the reference).
// Both are pure HBM traffic; the scatter is written as a gather over the <= Xwin rows that
// touch a latent step, so it needs no atomics and is deterministic.
//
// seq[s] = {row_start, nrows, lat_start, lat_len, ctl_start, ctl_len}; ctl_start already
// includes the reference's "-N-U_win+1" alignment offset.
#pragma once
#include "common.cuh"

namespace rgp {
namespace lag {

constexpr int DESC = 6;

__device__ __forceinline__ int find_seq(const int64_t* __restrict__ seq, int nseq, int field, int64_t x) {
  int lo = 0, hi = nseq - 1;             // last s with seq[s][field] <= x
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (seq[mid * DESC + field] <= x) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// One warp per output row.  Row n of sequence s is two contiguous source segments:
// lat[(lat_start+n)*Dx ... +Xwin*Dx) (lag-major-then-dim IS memory order) and
// ctl[(ctl_start+n)*Du ... +Uwin*Du), so the gather is a pair of coalesced segment copies;
// consecutive rows re-read overlapping windows out of L1/L2, DRAM sees each source byte once.
__global__ void k_gather(int nseq, const int64_t* __restrict__ seq, int64_t N, int Xwin, int Dx, int Uwin,
                         int Du, const double* __restrict__ lat, const double* __restrict__ ctl,
                         double* __restrict__ out) {
  const int Qx = Xwin * Dx, Qu = Uwin * Du, Q = Qx + Qu;
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < N; row += warps) {
    const int s = find_seq(seq, nseq, 0, row);
    const int64_t n = row - seq[s * DESC + 0];
    double* o = out + row * Q;
    if (Qx) {
      const double* src = lat + (seq[s * DESC + 2] + n) * Dx;
      for (int c = lane; c < Qx; c += 32) o[c] = src[c];
    }
    if (Qu) {
      const double* src = ctl + (seq[s * DESC + 4] + n) * Du;
      for (int c = lane; c < Qu; c += 32) o[Qx + c] = src[c];
    }
  }
}

// grad[t, j] += sum_{w} dX[row_start + (t - w), colbase + w*D + j] over 0 <= t - w < nrows
__global__ void k_scatter(int nseq, const int64_t* __restrict__ seq, int win, int D, int colbase, int Q,
                          int start_field, const double* __restrict__ dX, int64_t total,
                          double* __restrict__ grad) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total * D) return;
  int64_t tg = idx / D;
  int j = (int)(idx - tg * D);
  int s = find_seq(seq, nseq, start_field, tg);
  int64_t t = tg - seq[s * DESC + start_field];
  if (t >= seq[s * DESC + start_field + 1]) return;       // padding between sequences / unused tail
  const int64_t row0 = seq[s * DESC + 0], nrows = seq[s * DESC + 1];
  double acc = grad[idx];
  for (int w = win - 1; w >= 0; --w) {      // rows in increasing n: the reference's += order, bit for bit
    int64_t n = t - w;
    if (n >= 0 && n < nrows) acc += dX[(row0 + n) * Q + colbase + w * D + j];
  }
  grad[idx] = acc;
}

// Latent-state terms of a hidden layer (Layer_new._prepare_gradients, autoreg/layers.py:582-615
// with autoreg/variational.py:4-24): every latent step t of every sequence, every dim j:
//   t <  Xwin : -NormalPrior      value -(m^2 + v - log v)/2 + 1/2,  d/dm = -m,  d/dv = -(1 - 1/v)/2
//   t >= Xwin : -NormalEntropy    value (1 + log 2pi + log v)/2,     d/dv = dL_dYvar + 1/(2v),
//               d/dm = dL_dYmean  (the step is also output row t - Xwin of the layer)
// The gradients are WRITTEN (the reference zeroes them first, :596-597), the scatter of the
// row gradients adds onto them afterwards.  Per-block partial sums of the value go to
// `partial` and are added up in block order by k_sum_partials: deterministic.
// dyvar_cols: 1 = dL_dYvar is [N] (VarDTC, vardtc.py:206), D = [N, D] (SVI, svi_vardtc.py:193).
__global__ void k_latent_terms(int nseq, const int64_t* __restrict__ seq, int Xwin, int D,
                               const double* __restrict__ mean, const double* __restrict__ var,
                               const double* __restrict__ dYmean, const double* __restrict__ dYvar,
                               int dyvar_cols, int64_t total, double* __restrict__ gmean,
                               double* __restrict__ gvar, double* __restrict__ partial) {
  __shared__ double scratch[33];
  const double LOG_2_PI = 1.8378770664093454836;
  double val = 0.0;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total * D;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tg = idx / D;
    const int j = (int)(idx - tg * D);
    const int s = find_seq(seq, nseq, 2, tg);
    const int64_t t = tg - seq[s * DESC + 2];
    if (t >= seq[s * DESC + 3]) continue;
    const double m = mean[idx], v = var[idx];
    if (t < Xwin) {
      val += -0.5 * (m * m + (v - log(v))) + 0.5;
      gmean[idx] = -m;
      gvar[idx] = -((1.0 - 1.0 / v) * 0.5);
    } else {
      const int64_t row = seq[s * DESC + 0] + (t - Xwin);
      val += 0.5 * (1.0 + LOG_2_PI + log(v));
      gmean[idx] = dYmean[row * D + j];
      gvar[idx] = dYvar[dyvar_cols == 1 ? row : row * D + j] + 1.0 / (v * 2.0);
    }
  }
  val = block_sum(val, scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = val;
}

__global__ void k_sum_partials(int n, const double* __restrict__ partial, double* __restrict__ out) {
  __shared__ double scratch[33];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += partial[i];
  v = block_sum(v, scratch);
  if (threadIdx.x == 0) out[0] = v;
}

}  // namespace lag
}  // namespace rgp
