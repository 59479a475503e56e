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


_RUN_DL = None


def load_run_dl():
    """Import the reference's per-batch caller module ``running/run_dl.py`` (``get_reconstruction_model_input`` :238-329,
    ``get_transformed`` :146-204, ``apply_affine_augmentation`` :208-223, ``get_input_affine_for_atm`` :227-234) unmodified.

    Its module-level imports pull in packages that are absent here and irrelevant to those four functions; they are
    replaced by empty ``sys.modules`` stubs: ``wandb``, ``dill``, ``monai`` (logging / metrics), ``pytorch_run_on_
    recommended_gpu`` (GPU picker, ``run_dl.py:10-11``) and the reference's own ``utils/nnunetv2_utils`` (needs nnunetv2;
    only ``DC_and_CE_loss`` is imported from it, for the loss)."""
    global _RUN_DL
    if _RUN_DL is not None:
        return _RUN_DL
    load_reference()

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    for name in ("wandb", "dill", "monai"):
        try:
            __import__(name)
        except Exception:
            stub(name)
    try:
        import pytorch_run_on_recommended_gpu.run_on_recommended_gpu  # noqa: F401
    except Exception:
        pkg = stub("pytorch_run_on_recommended_gpu")
        pkg.run_on_recommended_gpu = stub("pytorch_run_on_recommended_gpu.run_on_recommended_gpu",
                                          get_cuda_environ_vars=lambda *a, **k: {})
    if "acquisition_focus.utils.nnunetv2_utils" not in sys.modules:
        try:
            import acquisition_focus.utils.nnunetv2_utils  # noqa: F401
        except Exception:
            stub("acquisition_focus.utils.nnunetv2_utils", DC_and_CE_loss=object)
    from acquisition_focus.running import run_dl
    from acquisition_focus.utils.python_utils import DotDict
    _RUN_DL = SimpleNamespace(module=run_dl, DotDict=DotDict,
                              get_reconstruction_model_input=run_dl.get_reconstruction_model_input,
                              get_transformed=run_dl.get_transformed,
                              apply_affine_augmentation=run_dl.apply_affine_augmentation,
                              get_input_affine_for_atm=run_dl.get_input_affine_for_atm)
    return _RUN_DL
