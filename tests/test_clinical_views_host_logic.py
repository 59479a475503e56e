"""Host logic of the GPU drop-in for get_clinical_cardiac_view_affines (functional/clinical_cardiac_views.py:223-364), checked
on CPU: the three device passes (group moments, extent bisection, nearest slice) are replaced by numpy / oracle restatements of
what the kernels compute, everything else (frames, eigenvectors, projections) is the product's own host code.  Against the
golden affines minted by the reference (acquisition_focus_b200/data/phantom_view_affines.json) and, where /root/reference is
mounted, against the reference function live on a second phantom."""
import numpy as np
import pytest
import torch

from oracle import af_oracle as O
from oracle import cases
from oracle.ref_import import load_reference, reference_available


from oracle.clinical_np import cpu_extent as _cpu_extent, cpu_moments as _cpu_moments


@pytest.fixture
def cpu_views(monkeypatch):
    from acquisition_focus_b200 import clinical_cardiac_views as CV
    from acquisition_focus_b200 import _lib as L
    monkeypatch.setattr(L, "require_cuda", lambda *a, **k: None)
    monkeypatch.setattr(CV, "_moments", _cpu_moments)
    monkeypatch.setattr(CV, "_extent_along_axis", _cpu_extent)
    monkeypatch.setattr(CV, "nifti_grid_sample", O.nifti_grid_sample)
    return CV


def test_host_logic_vs_reference_minted_golden(cpu_views):
    syn = cases.synthetic
    lab = torch.from_numpy(syn.heart_phantom(128))
    nii = torch.diag(torch.tensor([1.5, 1.5, 1.5, 1.0]))
    got = cpu_views.get_clinical_cardiac_view_affines(lab, nii, syn.CLASS_DICT, num_sa_slices=3, return_unrolled=True)
    want = syn.phantom_view_affines()
    assert list(got.keys()) == list(want.keys())
    for k in want:
        assert (got[k] - want[k]).abs().max().item() <= 2e-5, k


@pytest.mark.skipif(not reference_available(), reason="reference not mounted")
def test_host_logic_vs_reference_live(cpu_views):
    R = load_reference()
    syn = cases.synthetic
    base = syn.heart_phantom(64)
    lab = torch.from_numpy(np.ascontiguousarray(np.roll(base, 3, axis=1)))
    nii = torch.diag(torch.tensor([3.0, 3.0, 3.0, 1.0]))
    want = R.get_clinical_cardiac_view_affines(lab, nii, syn.CLASS_DICT, num_sa_slices=5, return_unrolled=False)
    got = cpu_views.get_clinical_cardiac_view_affines(lab, nii, syn.CLASS_DICT, num_sa_slices=5, return_unrolled=False)
    assert list(got.keys()) == list(want.keys()) and len(got["ALL_SA"]) == 5
    for k in want:
        a, b = (torch.stack(got[k]), torch.stack(want[k])) if k == "ALL_SA" else (got[k], want[k])
        assert (a - b).abs().max().item() <= 2e-5, k
    # a label map without a needed structure returns {} like the reference
    lab2 = lab.clone()
    lab2[(lab2 == syn.CLASS_DICT["MYO"]) | (lab2 == syn.CLASS_DICT["LV"])] = 0
    assert R.get_clinical_cardiac_view_affines(lab2, nii, syn.CLASS_DICT, 3) == {}
    assert cpu_views.get_clinical_cardiac_view_affines(lab2, nii, syn.CLASS_DICT, 3) == {}
