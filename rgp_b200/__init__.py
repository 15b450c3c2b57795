"""rgp_b200 - B200-native RBF-ARD psi-statistics (the one hot path of zhenwendai/RGP).

Public surface:
  PSICOMP_RBF_B200   drop-in for GPy's psicomp plugin (numpy in / numpy out)
  DevicePsi          device-resident API on torch CUDA tensors   (rgp_b200.device)
  ShardedPsi         row-sharded multi-GPU driver                 (rgp_b200.sharded)
  gpy_compat         duck-typed GPy RBF / NormalPosterior for tests
  DeviceBound        VarDTC / SVI bounds on the device, row-sharded  (rgp_b200.inference)
  LagWindow          lag-window gather / scatter, latent terms        (rgp_b200.lagwindow)
  DeviceDeepAutoreg  one whole model objective on the device          (rgp_b200.layer)
  RecognitionEncoder, MLPBackConstraint, deep_autoreg_objective       (encoder, backconstraint, autograd)
  load_checkpoint    numpy reader for the reference's HDF5 checkpoints (rgp_b200.checkpoint)
(the torch-based modules are imported on demand; importing rgp_b200 needs numpy only)

The arithmetic lives in rgp_b200/_lib/librgp_psi.so, built in-tree by
``python -m rgp_b200._build`` (nvcc, sm_100a).  There is no CPU fallback.
"""
from .psicomp import PSICOMP_RBF_B200  # noqa: F401
from ._lib import Handle, PsiError, load as load_library, library_path  # noqa: F401

__all__ = ["PSICOMP_RBF_B200", "Handle", "PsiError", "load_library", "library_path"]
__version__ = "0.1.0"
