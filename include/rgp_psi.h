/*
 * rgp_psi.h - C ABI of librgp_psi.so: RBF-ARD psi-statistics under Gaussian q(X) and
 * their gradients, fp64, hand-written CUDA for sm_100a (B200).
 *
 * This is the drop-in boundary for the ONE hot path of zhenwendai/RGP.  The reference
 * reaches the path through GPy's RBF kernel (a `psicomp` plugin object); each entry
 * point below cites the reference interface it stands behind.  GPy itself is a
 * third-party dependency that is not part of /root/reference.
 *
 *   forward  : GPy  PSICOMP_RBF.psicomputations(kern, Z, variational_posterior)
 *              <- kern.psi0/psi1/psi2(Z, X)   autoreg/inference/vardtc.py:59-61
 *                                             autoreg/inference/svi_vardtc.py:48-50
 *   backward : GPy  PSICOMP_RBF.psiDerivativecomputations(kern, dL_dpsi0, dL_dpsi1,
 *                                                         dL_dpsi2, Z, variational_posterior)
 *              <- kern.update_gradients_expectations   autoreg/layers.py:98-102
 *                 kern.gradients_Z_expectations         autoreg/layers.py:127-132
 *                 kern.gradients_qX_expectations        autoreg/layers.py:574-580
 *
 * Conventions
 *   - all matrices row-major (C order) fp64:  mu,S [N,Q]  Z [M,Q]  ell [Q]
 *     psi1, dL_dpsi1 [N,M]   psi2, dL_dpsi2 [M,M]   dmu,dS [N,Q]  dZ [M,Q]  dell [Q]
 *   - `ell` always has Q entries (a non-ARD kernel passes its single lengthscale
 *     replicated; the host wrapper sums dell back to a scalar as GPy does).
 *   - `*_dev` entry points take DEVICE pointers and enqueue on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream) without
 *     synchronising; `*_host` entry points take HOST pointers, copy in, run, copy
 *     out and synchronise before returning.
 *   - every function returns 0 on success or a negative rgp_psi_status; the message
 *     is available from rgp_psi_last_error() (thread local).  Nothing aborts.
 *   - a handle is bound to one device and may be used from one thread at a time;
 *     handles are independent (no hidden globals), one per (process, GPU).
 *   - there is NO CPU fallback: without a CUDA device rgp_psi_create fails.
 */
#ifndef RGP_PSI_H_
#define RGP_PSI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGP_PSI_ABI_VERSION 4   /* 2: + fused, latent-terms, MLP free-run entry points; 3: + rgp_host_digest; 4: + rgp_psi_small_schedule */

typedef struct rgp_psi_ctx* rgp_psi_handle_t;

typedef enum {
  RGP_PSI_OK = 0,
  RGP_PSI_ERR_INVALID = -1,   /* bad argument (null pointer, non-positive size, ...) */
  RGP_PSI_ERR_CUDA = -2,      /* a CUDA runtime call or kernel launch failed */
  RGP_PSI_ERR_NOMEM = -3,     /* device or pinned-host allocation failed */
  RGP_PSI_ERR_NODEVICE = -4   /* no usable sm_100 device */
} rgp_psi_status;

/* Largest input dimension either kernel family handles; larger Q is rejected with
 * RGP_PSI_ERR_INVALID by every entry point (the reference's configurations use Q = 10 ... 40,
 * the sweep of BASELINE.json stops at 128). */
#define RGP_PSI_MAX_Q 128

/* Which kernels serve a call.  AUTO = FAST (the tiled DMMA kernels; every supported shape).
 * REFERENCE = the simple one-thread-per-output kernels kept as an independent on-device
 * cross-check (they accumulate with atomicAdd, so their sums are not bit-reproducible; they
 * are never the served path). */
typedef enum { RGP_PSI_IMPL_AUTO = 0, RGP_PSI_IMPL_FAST = 1, RGP_PSI_IMPL_REFERENCE = 2 } rgp_psi_impl;

int rgp_psi_abi_version(void);
const char* rgp_psi_last_error(void);

/* Handle lifetime.  `device` is a CUDA ordinal. */
int rgp_psi_create(int device, rgp_psi_handle_t* out);
int rgp_psi_destroy(rgp_psi_handle_t h);

/* Options: "impl" (rgp_psi_impl), "row_chunk" (rows per internal device pass, 0 = 2^20),
 * "host_chunk" (rows per pipelined host<->device chunk of the *_host calls, 0 = 262144 for page-locked
 * caller buffers, 131072 when pageable buffers go through the pinned staging ring), "host_threads"
 * (threads of the pageable <-> pinned copies, 0 = min(8, cores / 2)),
 * "bwd_pipe" (Psi2 backward kernel: 0 = row-at-a-time, 1 = software-pipelined with TMA row-vector
 * staging, 2 (default) = row-at-a-time for the plain backward pass and pipelined for the fused pass,
 * the measured faster choice for each), "small_m" (kernels for small inducing sets, M <= 112 and Q <= 47,
 * where one CTA holds the whole pair matrix of a row: 0 = never, 1 = whenever the shape fits, 2 (default) =
 * when they also save work against the 64 x 64 block kernels), "small_ks" (their stage-2 k split: 0 =
 * default, or 1 / 2 / 4), "small_warps" (their CTA size: 0 = default, 16 = 16 warps and one CTA per SM, 8 = 8 warps
 * and two CTAs per SM where that fits, M <= 64), "profile" (1 = record a CUDA-event pair around every kernel
 * launch).  Experiment knobs that change results or occupancy ("debug_skip", "fwd_smem_pad") exist only
 * in libraries compiled with -DRGP_DEBUG. */
int rgp_psi_set_option(rgp_psi_handle_t h, const char* key, int64_t value);

/* ---- forward: Psi0 (optional N-vector), Psi1 (optional), Psi2 -------------------
 * Replaces psicomputations().  psi0_out / psi1_out may be NULL to skip them;
 * psi2_out must hold M*M doubles.  Psi0[n] = variance (the caller sums it,
 * vardtc.py:68). */
int rgp_psi_forward_dev(rgp_psi_handle_t h, void* stream, int64_t N, int M, int Q,
                        const double* mu, const double* S, const double* Z,
                        const double* ell, double variance,
                        double* psi0_out, double* psi1_out, double* psi2_out);

/* ---- backward: all five gradient blocks --------------------------------------
 * Replaces psiDerivativecomputations().  dL_dpsi0 may be NULL, in which case every
 * row uses dL_dpsi0_const (vardtc.py:175 passes a constant vector).  dL_dpsi1 may be
 * NULL (treated as zero).  dL_dpsi2 is symmetrised internally as GPy does.
 * Outputs: dmu,dS [N,Q], dZ [M,Q], dell [Q], dvar [1]; all overwritten. */
int rgp_psi_backward_dev(rgp_psi_handle_t h, void* stream, int64_t N, int M, int Q,
                         const double* mu, const double* S, const double* Z,
                         const double* ell, double variance,
                         const double* dL_dpsi0, double dL_dpsi0_const,
                         const double* dL_dpsi1, const double* dL_dpsi2,
                         double* dmu_out, double* dS_out, double* dZ_out,
                         double* dell_out, double* dvar_out);

/* ---- content digest of a host buffer (memoisation key of the plugin) -----------------------
 * GPy wraps psicomputations / psiDerivativecomputations in Cache_this(limit=10), valid because the
 * cacher observes paramz change notifications.  Without paramz the key must be the CONTENT: the layer
 * rewrites X.mean / X.variance in place on every evaluation (autoreg/layers.py:528-550; with the same
 * rows in a new order in testing/minibatch_tests.py:281-296).  out[0..1] = 128-bit order-sensitive
 * digest of data[0..nbytes); threads = 0 picks min(32, cores).  Pure host code, no device needed. */
int rgp_host_digest(const void* data, int64_t nbytes, int threads, uint64_t out[2]);

/* Work table of the small-inducing-set kernels (pure host code; introspection for tests and DESIGN.md).
 * For shape (M, Q), stage-2 k split `ks` (0 = default) and pass (`backward`: 0 forward only, 1 backward only,
 * 2 backward + Psi2) writes, as signed bytes,
 *   [0..15]  supertiles per warp        [16..79]  their indices, 4 per warp (row-major upper triangle of the
 *   16 x 16 supertile grid)             [80..95]  jobs per warp          [96..127] their job indices, 2 per warp
 *   [128] number of jobs  [129] k slots [130..161] strip of job j        [162..193] first k-step   [194..225] end
 *   k-step   [226..257] accumulator slot
 *   [258] warps per CTA (16: one CTA per SM, 8: two)   [259] supertile slots per warp   [260] buffers of L (2: one
 *   barrier per row, 1: two)   [261] job slots per warp   [262] 1 if the default rule ("small_m" = 2) serves this
 *   shape with the small kernels, 0 if it leaves it to the block kernels
 * and returns the number of bytes (263), or -1 (message set) when the small kernels do not serve the shape
 * or `out_bytes` is too small. */
int rgp_psi_small_schedule(int M, int Q, int ks, int backward, signed char* out, int out_bytes);

/* ---- host-buffer wrappers (the numpy-in / numpy-out plugin path) -------------------
 * Rows are streamed through double-buffered device mirrors on three streams (copy-in, compute,
 * copy-out) so the copies overlap the kernels.  Caller buffers may be ordinary pageable memory (what
 * numpy / GPy hand over): those are bounced through a pinned staging ring owned by the handle, filled
 * and drained by the calling thread (a few helper threads) while the GPU works on the neighbouring
 * chunk; page-locked caller buffers are used in place.  Device and pinned memory are bounded by
 * "host_chunk", not by N.  Synchronises before returning. */
int rgp_psi_forward_host(rgp_psi_handle_t h, int64_t N, int M, int Q,
                         const double* mu, const double* S, const double* Z,
                         const double* ell, double variance,
                         double* psi0_out, double* psi1_out, double* psi2_out);
int rgp_psi_backward_host(rgp_psi_handle_t h, int64_t N, int M, int Q,
                          const double* mu, const double* S, const double* Z,
                          const double* ell, double variance,
                          const double* dL_dpsi0, double dL_dpsi0_const,
                          const double* dL_dpsi1, const double* dL_dpsi2,
                          double* dmu_out, double* dS_out, double* dZ_out,
                          double* dell_out, double* dvar_out);

/* Fused evaluation: the statistics AND the gradients from one pass over the rows, for callers whose
 * upstream gradients do not depend on the statistics of the same evaluation - the uncollapsed SVI
 * bound, where dL_dpsi1 = beta Y (Kuu^-1 mu)^T and dL_dpsi2 = beta Lm^-T (D I - ...) Lm^-1 / 2 are
 * functions of q(U) and Kuu only (autoreg/inference/svi_vardtc.py:162-169).  The Psi2 backward kernel
 * has exp(E_n) in hand for every row and accumulates Psi2 on the side, so the separate forward pass
 * (a quarter of a two-phase evaluation) disappears.  Arguments as rgp_psi_backward_dev plus
 * psi1_out [N, M] (may be NULL) and psi2_out [M, M]. */
int rgp_psi_fused_dev(rgp_psi_handle_t h, void* stream, int64_t N, int M, int Q,
                      const double* mu, const double* S, const double* Z, const double* ell,
                      double variance, const double* dL_dpsi0, double dL_dpsi0_const,
                      const double* dL_dpsi1, const double* dL_dpsi2, double* psi1_out,
                      double* psi2_out, double* dmu_out, double* dS_out, double* dZ_out,
                      double* dell_out, double* dvar_out);

/* ---- lag-window gather / scatter-add (the callers either side of the path) ----------
 * Builds the layer input rows from the stacked latent sequences and adds X-row gradients
 * back onto latent steps: autoreg/layers.py:510-526 (_update_conv via get_conv_1D,
 * autoreg/util.py:6-12) and :552-571 (update_latent_gradients).
 *   seq_desc: device int64 [nseq][6] = {row_start, nrows, lat_start, lat_len, ctl_start,
 *             ctl_len}; ctl_start includes the reference's -N-U_win+1 alignment offset.
 *   lat [lat_total, Dx], ctl [ctl_total, Du] (ctl may be NULL when Uwin == 0), X [N, Q],
 *   Q = Xwin*Dx + Uwin*Du.  Call gather once for the means and once for the variances.
 *   scatter ACCUMULATES into lat_grad / ctl_grad (the reference uses +=). */
int rgp_lag_gather_dev(rgp_psi_handle_t h, void* stream, int nseq, const int64_t* seq_desc,
                       int64_t N, int Xwin, int Dx, int Uwin, int Du, const double* lat,
                       const double* ctl, double* X_out);
int rgp_lag_scatter_dev(rgp_psi_handle_t h, void* stream, int nseq, const int64_t* seq_desc,
                        int64_t N, int Xwin, int Dx, int Uwin, int Du, const double* dX,
                        int64_t lat_total, double* lat_grad, int64_t ctl_total, double* ctl_grad);

/* Latent-state terms of a hidden layer: Layer_new._prepare_gradients (autoreg/layers.py:582-615)
 * with NormalPrior / NormalEntropy (autoreg/variational.py:4-24).  For the stacked latent
 * series lat_mean / lat_var [lat_total, D] it WRITES lat_gmean / lat_gvar (the reference zeroes
 * them first) = the output-side gradients dL_dYmean [N, D] / dL_dYvar on the steps t >= Xwin
 * plus the entropy gradient there and the prior gradient on the first Xwin steps, and stores
 * the value the layer adds to its bound (-prior - entropy, "delta" at :594-615) in value_out
 * (one device double).  dyvar_cols = 1 when dL_dYvar is [N] (VarDTC), D when it is [N, D]
 * (SVI).  rgp_lag_scatter_dev then adds the input-side row gradients on top. */
int rgp_latent_terms_dev(rgp_psi_handle_t h, void* stream, int nseq, const int64_t* seq_desc,
                         int Xwin, int D, const double* lat_mean, const double* lat_var,
                         int64_t lat_total, const double* dL_dYmean, const double* dL_dYvar,
                         int dyvar_cols, double* lat_gmean, double* lat_gvar, double* value_out);

/* MLP back-constraint of a hidden layer (autoreg/layers.py:623-715, network of autoreg/mlp.py): the
 * first Xwin latent means of a sequence are free, every later one is the output of a tanh MLP on the
 * window before it and the aligned control window.  One CTA per sequence, weights in shared memory.
 *   units[0 .. nlayers] (HOST array): layer widths, units[0] = Xwin*Dx + Uwin*Du, units[nlayers] = Dx;
 *   params (device): for each layer W[down][up] row-major then b[down]; tanh on hidden layers, linear last;
 *   seq_desc as for the lag-window calls (row_start numbers the generated steps).
 * freerun     reads lat_mean rows < Xwin of every sequence and WRITES the rest; hidden_acts [N, sum hidden
 *             widths] keeps the activations for the backward call.
 * freerun_bwd lat_gmean holds dL/d mean of every step on entry; on exit its rows < Xwin of every sequence
 *             hold the initial-mean gradients (objective part + back-propagated part), the other rows
 *             are unchanged; the call ADDS onto ctl_gmean (may be NULL) and writes param_grads
 *             [nseq, nparams] (sum over the first axis for the total). */
int rgp_mlp_freerun_dev(rgp_psi_handle_t h, void* stream, int nseq, const int64_t* seq_desc, int Xwin,
                        int Dx, int Uwin, int Du, int nlayers, const int* units, const double* params,
                        double* lat_mean, const double* ctl_mean, double* hidden_acts);
int rgp_mlp_freerun_bwd_dev(rgp_psi_handle_t h, void* stream, int nseq, const int64_t* seq_desc, int Xwin,
                            int Dx, int Uwin, int Du, int nlayers, const int* units, const double* params,
                            const double* lat_mean, const double* ctl_mean, const double* hidden_acts,
                            double* lat_gmean, double* ctl_gmean, double* param_grads);

/* ---- measurement support ------------------------------------------------------- */
/* Kernel launches issued through this handle since creation (or the last reset). */
int64_t rgp_psi_launch_count(rgp_psi_handle_t h);
int rgp_psi_reset_counters(rgp_psi_handle_t h);
/* With option "profile"=1: per-kernel accumulated device time since the last reset.
 * Fills up to `cap` entries; names are static strings.  Synchronises the device.
 * Returns the number of distinct kernels (may exceed cap) or a negative status. */
int rgp_psi_kernel_times(rgp_psi_handle_t h, int cap, const char** names,
                         double* total_ms, int64_t* launches);
/* DFMA-chain microbenchmark: achieved fp64 FMA throughput of this GPU in TFLOP/s
 * (2 flop per FMA), best of `reps` launches.  The roofline denominator. */
int rgp_psi_fp64_peak(rgp_psi_handle_t h, void* stream, int reps, double* tflops_out);
/* Device workspace currently held by the handle, in bytes. */
int64_t rgp_psi_workspace_bytes(rgp_psi_handle_t h);

#ifdef __cplusplus
}
#endif
#endif /* RGP_PSI_H_ */
