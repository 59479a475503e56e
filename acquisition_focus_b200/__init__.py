"""acquisition_focus_b200 - B200-native (sm_100a) implementation of the differentiable
view-acquisition hot path of multimodallearning/acquisition-focus.

Layout
  csrc/            hand-written CUDA kernels + the C ABI (include/afb200.h) -> lib/libafb200.so
  _lib.py          ctypes binding (fails loudly when the library is missing; no CPU fallback)
  functional.py    autograd entry points (slice extraction, view-parameter chain, embedding)
  utils/, models/  drop-in mirrors of the reference's call signatures on this path
  parallel.py      batch x view sharding over the GPUs of one box + NCCL all-reduce of dTheta
  install.py       monkey-patch an importable reference checkout to run on these kernels
"""
from . import functional  # noqa: F401
from .functional import (acquire_views, acquire_views_from_labels, affine_grid_sample, embed_slices,  # noqa: F401
                         embed_slices_multi, r6_to_matrix, slice_with_pre_affine, volume_min)
from .clinical_cardiac_views import get_clinical_cardiac_view_affines  # noqa: F401
from .models.hybrid_unet import SkipConnector  # noqa: F401
from .models.learnable_transform import AffineTransformModule, ATModulesContainer  # noqa: F401
from .utils.nifti_utils import nifti_grid_sample  # noqa: F401
from .utils.transform_utils import compute_rotation_matrix_from_ortho6d  # noqa: F401

__version__ = "0.1.0"
