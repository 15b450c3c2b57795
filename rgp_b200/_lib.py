"""ctypes binding of librgp_psi.so (the C ABI declared in include/rgp_psi.h).

Loading never falls back to a CPU implementation: if the shared library is missing the
import of the *library* raises, and if no CUDA device is present ``Handle()`` raises
``PsiError`` with the library's message.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Tuple

from . import _build

c_double_p = C.POINTER(C.c_double)

# name -> (restype, argtypes); must list every symbol include/rgp_psi.h declares.
SIGNATURES = {
    "rgp_psi_abi_version": (C.c_int, []),
    "rgp_psi_last_error": (C.c_char_p, []),
    "rgp_psi_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "rgp_psi_destroy": (C.c_int, [C.c_void_p]),
    "rgp_psi_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "rgp_psi_forward_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "rgp_psi_backward_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                       C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rgp_psi_forward_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "rgp_psi_backward_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                        C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rgp_psi_fused_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                    C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rgp_lag_gather_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rgp_lag_scatter_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                      C.c_void_p]),
    "rgp_latent_terms_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "rgp_mlp_freerun_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rgp_mlp_freerun_bwd_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rgp_host_digest": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_uint64)]),
    "rgp_psi_launch_count": (C.c_int64, [C.c_void_p]),
    "rgp_psi_reset_counters": (C.c_int, [C.c_void_p]),
    "rgp_psi_kernel_times": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), c_double_p,
                                       C.POINTER(C.c_int64)]),
    "rgp_psi_small_schedule": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]),
    "rgp_psi_fp64_peak": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_double_p]),
    "rgp_psi_workspace_bytes": (C.c_int64, [C.c_void_p]),
}

IMPL_AUTO, IMPL_FAST, IMPL_REFERENCE = 0, 1, 2

_lib: Optional[C.CDLL] = None


class PsiError(RuntimeError):
    """A librgp_psi call returned a non-zero status."""

    def __init__(self, status: int, message: str):
        super().__init__(f"librgp_psi status {status}: {message}")
        self.status = status


def library_path() -> str:
    """The in-tree product library; RGP_PSI_LIB points the timing probes at the -DRGP_DEBUG build."""
    return os.environ.get("RGP_PSI_LIB") or _build.LIB


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the in-tree librgp_psi.so, binding every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if path != _build.LIB:
        if not os.path.exists(path):
            raise OSError(f"RGP_PSI_LIB={path} does not exist")
    elif not os.path.exists(path) or (build_if_missing and not _build.is_current()):
        if not build_if_missing:
            raise OSError(f"{path} is missing; run `python -m rgp_b200._build` "
                          "(there is no CPU fallback)")
        _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.rgp_psi_abi_version() != 4:
        raise OSError("librgp_psi ABI version mismatch")
    _lib = lib
    return lib


def host_digest(a) -> bytes:
    """128-bit order-sensitive content digest of a C-contiguous numpy array (rgp_host_digest)."""
    out = (C.c_uint64 * 2)()
    check(load().rgp_host_digest(C.c_void_p(a.ctypes.data), a.nbytes, 0, out))
    return bytes(out)


def check(status: int) -> None:
    if status != 0:
        msg = load().rgp_psi_last_error()
        raise PsiError(status, msg.decode("utf-8", "replace") if msg else "")


class Handle:
    """Owns one rgp_psi_handle_t (one per process and GPU).  Pickle/deepcopy safe: the
    device handle is dropped and lazily re-created (the reference deep-copies models,
    testing/minibatch_tests.py:91)."""

    def __init__(self, device: int = 0):
        self.device = int(device)
        self._h: Optional[C.c_void_p] = None
        self._options: Dict[str, int] = {}

    # -- lifetime
    def _ensure(self) -> C.c_void_p:
        if self._h is None:
            lib = load()
            h = C.c_void_p()
            check(lib.rgp_psi_create(self.device, C.byref(h)))
            self._h = h
            for k, v in self._options.items():
                check(lib.rgp_psi_set_option(h, k.encode(), int(v)))
        return self._h

    def close(self) -> None:
        if self._h is not None and _lib is not None:
            _lib.rgp_psi_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __getstate__(self):
        return {"device": self.device, "_options": dict(self._options)}

    def __setstate__(self, state):
        self.device = state["device"]
        self._options = state["_options"]
        self._h = None

    # -- options / measurement
    def set_option(self, key: str, value: int) -> None:
        if self._h is not None:
            check(load().rgp_psi_set_option(self._h, key.encode(), int(value)))
        self._options[key] = int(value)          # remembered only once accepted (or until first use)

    def launch_count(self) -> int:
        return int(load().rgp_psi_launch_count(self._ensure()))

    def reset_counters(self) -> None:
        check(load().rgp_psi_reset_counters(self._ensure()))

    def kernel_times(self) -> Dict[str, Tuple[float, int]]:
        cap = 64
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        cnt = (C.c_int64 * cap)()
        n = load().rgp_psi_kernel_times(self._ensure(), cap, names, ms, cnt)
        if n < 0:
            check(n)
        return {names[i].decode(): (float(ms[i]), int(cnt[i])) for i in range(min(n, cap))}

    def fp64_peak(self, stream: int = 0, reps: int = 5) -> float:
        out = C.c_double(0.0)
        check(load().rgp_psi_fp64_peak(self._ensure(), C.c_void_p(stream), reps, C.byref(out)))
        return float(out.value)

    def workspace_bytes(self) -> int:
        return int(load().rgp_psi_workspace_bytes(self._ensure()))

    # -- raw pointer calls (device or host pointers as integers)
    def forward_dev(self, stream, N, M, Q, mu, S, Z, ell, variance, psi0, psi1, psi2) -> None:
        check(load().rgp_psi_forward_dev(self._ensure(), C.c_void_p(stream), N, M, Q, mu, S, Z, ell,
                                         float(variance), psi0, psi1, psi2))

    def backward_dev(self, stream, N, M, Q, mu, S, Z, ell, variance, dL0, dL0c, dL1, dL2,
                     dmu, dS, dZ, dell, dvar) -> None:
        check(load().rgp_psi_backward_dev(self._ensure(), C.c_void_p(stream), N, M, Q, mu, S, Z, ell,
                                          float(variance), dL0, float(dL0c), dL1, dL2,
                                          dmu, dS, dZ, dell, dvar))

    def fused_dev(self, stream, N, M, Q, mu, S, Z, ell, variance, dL0, dL0c, dL1, dL2, psi1, psi2,
                  dmu, dS, dZ, dell, dvar) -> None:
        check(load().rgp_psi_fused_dev(self._ensure(), C.c_void_p(stream), N, M, Q, mu, S, Z, ell,
                                       float(variance), dL0, float(dL0c), dL1, dL2, psi1, psi2,
                                       dmu, dS, dZ, dell, dvar))

    def lag_gather(self, stream, nseq, seq_desc, N, Xwin, Dx, Uwin, Du, lat, ctl, out) -> None:
        check(load().rgp_lag_gather_dev(self._ensure(), C.c_void_p(stream), nseq, seq_desc, N, Xwin, Dx, Uwin, Du,
                                        lat, ctl, out))

    def lag_scatter(self, stream, nseq, seq_desc, N, Xwin, Dx, Uwin, Du, dX, lat_total, lat_grad,
                    ctl_total, ctl_grad) -> None:
        check(load().rgp_lag_scatter_dev(self._ensure(), C.c_void_p(stream), nseq, seq_desc, N, Xwin, Dx, Uwin, Du,
                                         dX, lat_total, lat_grad, ctl_total, ctl_grad))

    def latent_terms(self, stream, nseq, seq_desc, Xwin, D, lat_mean, lat_var, lat_total, dYmean, dYvar,
                     dyvar_cols, gmean, gvar, value_out) -> None:
        check(load().rgp_latent_terms_dev(self._ensure(), C.c_void_p(stream), nseq, seq_desc, Xwin, D, lat_mean,
                                          lat_var, lat_total, dYmean, dYvar, dyvar_cols, gmean, gvar, value_out))

    def mlp_freerun(self, stream, nseq, seq_desc, Xwin, Dx, Uwin, Du, units, params, lat, ctl, acts) -> None:
        u = (C.c_int * len(units))(*units)
        check(load().rgp_mlp_freerun_dev(self._ensure(), C.c_void_p(stream), nseq, seq_desc, Xwin, Dx, Uwin, Du,
                                         len(units) - 1, u, params, lat, ctl, acts))

    def mlp_freerun_bwd(self, stream, nseq, seq_desc, Xwin, Dx, Uwin, Du, units, params, lat, ctl, acts,
                        lat_g, ctl_g, pgrad) -> None:
        u = (C.c_int * len(units))(*units)
        check(load().rgp_mlp_freerun_bwd_dev(self._ensure(), C.c_void_p(stream), nseq, seq_desc, Xwin, Dx, Uwin, Du,
                                             len(units) - 1, u, params, lat, ctl, acts, lat_g, ctl_g, pgrad))

    def forward_host(self, N, M, Q, mu, S, Z, ell, variance, psi0, psi1, psi2) -> None:
        check(load().rgp_psi_forward_host(self._ensure(), N, M, Q, mu, S, Z, ell, float(variance),
                                          psi0, psi1, psi2))

    def backward_host(self, N, M, Q, mu, S, Z, ell, variance, dL0, dL0c, dL1, dL2,
                      dmu, dS, dZ, dell, dvar) -> None:
        check(load().rgp_psi_backward_host(self._ensure(), N, M, Q, mu, S, Z, ell, float(variance),
                                           dL0, float(dL0c), dL1, dL2, dmu, dS, dZ, dell, dvar))
