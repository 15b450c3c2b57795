// fast_path.cuh - placeholder; replaced by the tiled kernels.
#pragma once
#include "context.cuh"
namespace rgp { namespace fast {
static inline bool supported(int, int) { return false; }
static inline int init(rgp_psi_ctx*) { return 0; }
static inline int forward(rgp_psi_ctx*, cudaStream_t, int64_t, int, int, const double*, const double*, const double*, const double*, double, double*, double*, double*) { return set_error(RGP_PSI_ERR_INVALID, "fast path not built"); }
static inline int backward(rgp_psi_ctx*, cudaStream_t, int64_t, int, int, const double*, const double*, const double*, const double*, double, const double*, double, const double*, const double*, double*, double*, double*, double*, double*) { return set_error(RGP_PSI_ERR_INVALID, "fast path not built"); }
}}
