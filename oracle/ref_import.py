"""ORACLE (test infrastructure only): import the *unmodified* reference.

The reference (multimodallearning/acquisition-focus) is pure Python on top of
torch, so its own implementation of the hot path can be imported and run on CPU
in the build container, where ``/root/reference`` is mounted read-only.  That
is how the golden vectors under ``tests/golden/`` were minted
(``oracle/make_golden.py``) and how ``oracle/af_oracle.py`` (the restatement
that *can* travel to the GPU box) is validated.

``/root/reference`` does not exist on the GPU box: nothing under ``-m gpu``,
``__graft_entry__.smoke()`` or ``bench.py`` may call :func:`load_reference`.

Two ``sys.modules`` stubs are needed (plot-only / base-class-only imports):
``matplotlib.pyplot`` (``functional/clinical_cardiac_views.py:3``) and
``dynamic_network_architectures.architectures.unet.PlainConvUNet``
(``models/hybrid_unet.py:3``).
"""
from __future__ import annotations

import os
import sys
import types
import warnings
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("AFB_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "acquisition_focus"))


def _install_stubs() -> None:
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    try:
        import dynamic_network_architectures.architectures.unet  # noqa: F401
    except Exception:
        import torch

        root = types.ModuleType("dynamic_network_architectures")
        arch = types.ModuleType("dynamic_network_architectures.architectures")
        unet = types.ModuleType("dynamic_network_architectures.architectures.unet")

        class PlainConvUNet(torch.nn.Module):  # base class of HybridUnet only
            def __init__(self, *a, **k):
                super().__init__()

        unet.PlainConvUNet = PlainConvUNet
        root.architectures = arch
        arch.unet = unet
        sys.modules["dynamic_network_architectures"] = root
        sys.modules["dynamic_network_architectures.architectures"] = arch
        sys.modules["dynamic_network_architectures.architectures.unet"] = unet


_CACHE = None


def load_reference() -> SimpleNamespace:
    """Return the reference's hot-path callables (imported, not copied)."""
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    warnings.filterwarnings("ignore", category=UserWarning)
    warnings.filterwarnings("ignore", category=FutureWarning)
    from acquisition_focus.utils import nifti_utils, transform_utils
    from acquisition_focus.models import learnable_transform, hybrid_unet
    from acquisition_focus.functional import clinical_cardiac_views

    _CACHE = SimpleNamespace(
        nifti_utils=nifti_utils,
        transform_utils=transform_utils,
        learnable_transform=learnable_transform,
        hybrid_unet=hybrid_unet,
        clinical_cardiac_views=clinical_cardiac_views,
        nifti_grid_sample=nifti_utils.nifti_grid_sample,
        compute_rotation_matrix_from_ortho6d=transform_utils.compute_rotation_matrix_from_ortho6d,
        get_random_affine=transform_utils.get_random_affine,
        AffineTransformModule=learnable_transform.AffineTransformModule,
        SkipConnector=hybrid_unet.SkipConnector,
        get_clinical_cardiac_view_affines=clinical_cardiac_views.get_clinical_cardiac_view_affines,
    )
    return _CACHE
