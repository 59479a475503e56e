"""Route an importable ``acquisition_focus`` (the reference) through the B200 kernels.

    import acquisition_focus_b200.install as afb_install
    afb_install.install()        # after `import acquisition_focus...`

Every reference module on the training hot path that did ``from ...nifti_utils import
nifti_grid_sample`` holds its own name binding (``models/learnable_transform.py:6``,
``running/run_dl.py:30``), so each one is patched, as are ``compute_rotation_matrix_from_ortho6d``
(``models/learnable_transform.py:8``), ``hybrid_unet.SkipConnector`` and, when ``running.run_dl`` is imported, its per-batch
callers ``get_transformed`` / ``get_reconstruction_model_input`` (same signatures; with the reference's own
``ATModulesContainer`` they take the view-by-view route, with this package's container the fused all-views one).  The offline, host-side
callers (``datasets/base_dataset.py:19``, ``functional/clinical_cardiac_views.py:4``,
``utils/nnunetv2_utils.py:19``) are deliberately left on the reference's own code: they run once at
dataset-preparation time on CPU tensors and are outside this path.
The patched functions accept CUDA tensors only and raise otherwise - there is no CPU fallback.
"""
from __future__ import annotations

import sys

from .models.hybrid_unet import SkipConnector
from .running.model_input import get_reconstruction_model_input, get_transformed
from .utils.nifti_utils import nifti_grid_sample
from .utils.transform_utils import compute_rotation_matrix_from_ortho6d

_PATCH_SAMPLE = ["acquisition_focus.models.learnable_transform", "acquisition_focus.running.run_dl"]
_PATCH_R6 = ["acquisition_focus.models.learnable_transform"]

_originals = {}


def install() -> list:
    """Patch every already-imported hot-path module of the reference; returns what was patched."""
    done = []
    for name in _PATCH_SAMPLE:
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "nifti_grid_sample"):
            _originals.setdefault((name, "nifti_grid_sample"), mod.nifti_grid_sample)
            mod.nifti_grid_sample = nifti_grid_sample
            done.append(f"{name}.nifti_grid_sample")
    for name in _PATCH_R6:
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "compute_rotation_matrix_from_ortho6d"):
            _originals.setdefault((name, "compute_rotation_matrix_from_ortho6d"), mod.compute_rotation_matrix_from_ortho6d)
            mod.compute_rotation_matrix_from_ortho6d = compute_rotation_matrix_from_ortho6d
            done.append(f"{name}.compute_rotation_matrix_from_ortho6d")
    mod = sys.modules.get("acquisition_focus.running.run_dl")          # the per-batch callers (run_dl.py:146-204, 238-329)
    if mod is not None:
        for attr, fn in (("get_transformed", get_transformed), ("get_reconstruction_model_input", get_reconstruction_model_input)):
            if hasattr(mod, attr):
                _originals.setdefault(("acquisition_focus.running.run_dl", attr), getattr(mod, attr))
                setattr(mod, attr, fn)
                done.append(f"acquisition_focus.running.run_dl.{attr}")
    mod = sys.modules.get("acquisition_focus.models.hybrid_unet")
    if mod is not None:
        _originals.setdefault(("acquisition_focus.models.hybrid_unet", "SkipConnector"), mod.SkipConnector)
        mod.SkipConnector = SkipConnector
        done.append("acquisition_focus.models.hybrid_unet.SkipConnector")
    return done


def uninstall() -> None:
    for (name, attr), fn in _originals.items():
        setattr(sys.modules[name], attr, fn)
    _originals.clear()
