"""Drop-in for ``get_clinical_cardiac_view_affines`` of the reference's ``functional/clinical_cardiac_views.py:223-364``
(last item): axial / sagittal / coronal / p2CH / p4CH / SA-k / 4CH / 2CH torch-grid affines from a 3-D label map.

The reference runs this once per volume at dataset-load time on sparse CPU tensors; for on-line ``from-segmented`` use it has
to run on the device.  Here every pass over voxels is a CUDA kernel:

* centre and inertia tensor of the four label groups (MYO+LV, MYO+LV+LA, MYO+LV+RV, whole heart) in ONE pass over the integer
  label map (``afb_label_group_moments``: exact 64-bit integer moments; reference ``utils/torch_sparse_tensor_utils.py:34-56``);
* the extent of MYO+LV along the LV axis (``afb_label_extent_search``: the bisection of reference ``:36-62`` in one launch);
* the three in-plane inertia analyses (reference ``:178-204``): a nearest-neighbour slice of the label map through the CUDA
  sampler (``nifti_grid_sample(is_label=True)``, FOV 300 x 300 x 1 mm at 128 x 128 x 1) + the same moments kernel on the slice.

The 3x3 eigenproblems (``torch.linalg.eig``, which also pins the reference's eigenvector sign convention) and the frame algebra
in between are a few dozen flops per volume and run on the host in fp32 with the reference's op order, six small D2H reads per
volume.  Same signature, same return dict (CPU fp32 4x4 tensors, like the reference).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .utils.nifti_utils import nifti_grid_sample


def _group_mask(values) -> int:
    m = 0
    for v in values:
        v = int(v)
        assert 0 < v < 32, "label values 1..31"
        m |= 1 << v
    return m


def _moments(label_dev: torch.Tensor, masks):
    """Exact integer moments of the voxel index cloud per group -> (count[G], centre[G,3] fp32, inertia[G,3,3] fp32) on the host.
    The inertia tensor is taken about the fp32-rounded centre, as the reference does (``dists = idxs - center``)."""
    D, H, W = label_dev.shape
    dev = label_dev.device
    G = len(masks)
    with torch.cuda.device(dev):
        m = torch.tensor(masks, dtype=torch.int64).to(torch.int32).to(dev)       # bit patterns of uint32 masks
        out = torch.zeros((G, 10), dtype=torch.int64, device=dev)
        L.check(L.lib().afb_label_group_moments(L.ptr(label_dev), L.DTYPES[label_dev.dtype], D, H, W, L.ptr(m), G, L.ptr(out),
                                                L.stream_ptr(dev)), "afb_label_group_moments")
    s = out.cpu().numpy().astype(np.float64)
    cnt = s[:, 0]
    centers = torch.zeros(G, 3)
    inertia = torch.zeros(G, 3, 3)
    for g in range(G):
        n = cnt[g]
        if n == 0:
            continue
        s1 = s[g, 1:4]
        s2 = np.array([[s[g, 4], s[g, 5], s[g, 6]], [s[g, 5], s[g, 7], s[g, 8]], [s[g, 6], s[g, 8], s[g, 9]]])
        c32 = torch.tensor(s1 / n, dtype=torch.float32)
        c = c32.double().numpy()
        cov = s2 - np.outer(c, s1) - np.outer(s1, c) + n * np.outer(c, c)       # sum (x - c)_i (x - c)_j
        r2 = np.trace(cov)
        centers[g] = c32
        inertia[g] = torch.from_numpy(r2 * np.eye(3) - cov).float()
    return cnt, centers, inertia


def get_main_principal_axes(inertia: torch.Tensor):
    """Eigenvectors of the inertia tensor sorted by eigenvalue: (min, mid, max) - reference torch_sparse_tensor_utils.py:79-85."""
    eig = torch.linalg.eig(inertia)
    vecs = eig.eigenvectors.real.T[eig.eigenvalues.real.argsort()]
    return vecs[0], vecs[1], vecs[2]


def get_vector_projection(projectee, base_vect, orthogonal_to_base=False):
    if orthogonal_to_base:
        return projectee - projectee @ base_vect * base_vect
    return projectee @ base_vect * base_vect


def get_angle_between_vectors(v1, v2):
    v1 = v1 / torch.linalg.norm(v1, 2)
    v2 = v2 / torch.linalg.norm(v2, 2)
    return torch.acos(v1 @ v2)


def get_torch_grid_affine_from_pix_affine(pix_affine, shape):
    """Pixel-space frame -> torch grid affine (reference :66-71)."""
    pt = pix_affine.clone()
    pt[:3, :3] = pt[:3, :3].flip(0, 1).T
    pt[:3, -1] = (2.0 * pt[:3, -1] / torch.as_tensor(shape) - 1.0).flip(0)
    return pt


def get_pix_affine_from_center_and_plane_vects(px_center, main_plane_vect, plane_vect_two, px_center_projected=None,
                                               do_return_normal_three=False):
    """Orthonormal frame (rows: second in-plane axis, main axis, normal) through a centre (reference :75-100).  Like the
    reference it normalises the two given vectors IN PLACE (later calls see the normalised vectors)."""
    main_plane_vect /= torch.linalg.norm(main_plane_vect, 2)
    plane_vect_two /= torch.linalg.norm(plane_vect_two, 2)
    normal_three = torch.linalg.cross(main_plane_vect, plane_vect_two)
    normal_three /= torch.linalg.norm(normal_three, 2)
    plane_vect_two = torch.linalg.cross(normal_three, main_plane_vect)
    affine = torch.eye(4)
    affine[:3, :3] = torch.stack([plane_vect_two, main_plane_vect, normal_three], dim=0)
    if px_center_projected is not None:
        delta_center = px_center_projected - px_center
        affine[:3, -1] = px_center + get_vector_projection(delta_center[:3], normal_three[:3], orthogonal_to_base=True)
    else:
        affine[:3, -1] = px_center
    if do_return_normal_three:
        return affine, normal_three
    return affine


def _extent_along_axis(label_dev, mask, center, direction):
    """``get_min_max_extent_along_axis`` (reference :51-62): the two end points of the group along +dir / -dir."""
    D, H, W = label_dev.shape
    dev = label_dev.device
    init_end = torch.linalg.vector_norm(torch.as_tensor(label_dev.shape, dtype=torch.float), 2).item()
    with torch.cuda.device(dev):
        c = center.to(dev, torch.float32).contiguous()
        d = direction.to(dev, torch.float32).contiguous()
        out = torch.zeros(2, dtype=torch.float64, device=dev)
        L.check(L.lib().afb_label_extent_search(L.ptr(label_dev), L.DTYPES[label_dev.dtype], D, H, W, C.c_uint(mask), L.ptr(c), L.ptr(d),
                                                float(init_end), L.ptr(out), L.stream_ptr(dev)), "afb_label_extent_search")
    f = out.cpu().tolist()
    return center + f[0] * direction, center + f[1] * (-direction)


def get_slice_center_inertia_in_volume_space(label_dev, mask, volume_affine, pix_affine, label_shape):
    """Principal axes of a group inside the slice given by ``pix_affine``, mapped back to volume space (reference :178-204)."""
    fov_mm, fov_vox = torch.tensor([300.0, 300.0, 1.0]), torch.tensor([128, 128, 1])
    slicing = get_torch_grid_affine_from_pix_affine(pix_affine, label_shape)
    dev = label_dev.device
    lbl_slice, *_ = nifti_grid_sample(label_dev[None, None], volume_affine[None].to(dev), target_fov_mm=fov_mm, target_fov_vox=fov_vox,
                                      is_label=True, pre_grid_sample_affine=slicing[None].to(dev))
    # nearest sampling commutes with the group filter: slice the full label map, filter inside the moments kernel
    _, _, slc_inertia = _moments(lbl_slice[0, 0].contiguous(), [mask])
    mn, md, mx = get_main_principal_axes(slc_inertia[0])
    inv = pix_affine.inverse()[:3, :3]
    return inv @ mn, inv @ md, inv @ mx


def get_clinical_cardiac_view_affines(label: torch.Tensor, volume_affine, class_dict: dict, num_sa_slices: int = 3,
                                      return_unrolled=False, debug=False):
    """``label`` [D,H,W] integer label map on a CUDA device, ``volume_affine`` [4,4] NIfTI affine -> dict view name -> 4x4 torch
    grid affine (CPU fp32), ``'ALL_SA'`` a list (or ``'SA-k'`` entries with ``return_unrolled``); ``{}`` if a needed structure is
    missing.  Same contract as the reference."""
    L.require_cuda(label, "label")
    assert label.dim() == 3 and not label.dtype.is_floating_point
    for k in ("LV", "RV", "MYO", "LA"):
        assert k in class_dict
    assert num_sa_slices % 2 == 1
    if label.dtype not in L.DTYPES:
        label = label.to(torch.int32)
    label = label.contiguous()
    label_shape = list(label.shape)
    volume_affine = torch.as_tensor(volume_affine).detach().cpu()

    m_myolv = _group_mask((class_dict["MYO"], class_dict["LV"]))
    m_myolvla = _group_mask((class_dict["MYO"], class_dict["LV"], class_dict["LA"]))
    m_myolvrv = _group_mask((class_dict["MYO"], class_dict["LV"], class_dict["RV"]))
    m_heart = _group_mask(class_dict.values())
    cnt, centers, inertia = _moments(label, [m_myolv, m_myolvla, m_myolvrv, m_heart])
    if (cnt == 0).any():
        return {}
    myolv_center, myolvla_center, heart_center = centers[0], centers[1], centers[3]

    # 0. axial, sagittal, coronal (reference :249-261)
    sagittal_vect, coronal_vect, axial_vect = torch.tensor([1.0, 0, 0]), torch.tensor([0, 1.0, 0]), torch.tensor([0, 0, 1.0])
    pix_axial = get_pix_affine_from_center_and_plane_vects(heart_center, sagittal_vect, coronal_vect)
    pix_coronal = get_pix_affine_from_center_and_plane_vects(heart_center, axial_vect, sagittal_vect)
    pix_sagittal = get_pix_affine_from_center_and_plane_vects(heart_center, coronal_vect, axial_vect)
    views = {"axial": get_torch_grid_affine_from_pix_affine(pix_axial, label_shape),
             "sagittal": get_torch_grid_affine_from_pix_affine(pix_sagittal, label_shape),
             "coronal": get_torch_grid_affine_from_pix_affine(pix_coronal, label_shape)}

    # 1. LV + MYO centre line, pointing to the base (:263-270)
    lv_min_principal, *_ = get_main_principal_axes(inertia[0])
    if get_angle_between_vectors(lv_min_principal[:3], sagittal_vect[:3]) < np.pi / 2:
        lv_min_principal = -1 * lv_min_principal

    # 2. pseudo 2CH / 4CH (:276-291)
    pix_p2ch, ortho_p2ch = get_pix_affine_from_center_and_plane_vects(myolv_center, lv_min_principal, axial_vect,
                                                                     px_center_projected=heart_center, do_return_normal_three=True)
    views["p2CH"] = get_torch_grid_affine_from_pix_affine(pix_p2ch, label_shape)
    pix_p4ch, ortho_p4ch = get_pix_affine_from_center_and_plane_vects(myolv_center, lv_min_principal, ortho_p2ch,
                                                                     px_center_projected=heart_center, do_return_normal_three=True)
    views["p4CH"] = get_torch_grid_affine_from_pix_affine(pix_p4ch, label_shape)

    # 4. short-axis stack from base to apex (:293-306)
    p1, p2 = _extent_along_axis(label, m_myolv, myolv_center, lv_min_principal)
    delta_p = p2 - p1
    sa = []
    for k in range(num_sa_slices):
        p_along = p1 + delta_p * k / (num_sa_slices - 1)
        pix_sa = get_pix_affine_from_center_and_plane_vects(p_along, ortho_p2ch, ortho_p4ch, px_center_projected=heart_center)
        sa.append(get_torch_grid_affine_from_pix_affine(pix_sa, label_shape))
    views["ALL_SA"] = sa

    # 5. 4CH from the in-plane axes of MYO+LV+RV in the centre SA slice and of MYO+LV+LA in the p2CH slice (:308-328)
    pix_center_sa = get_pix_affine_from_center_and_plane_vects(p1 + 0.5 * delta_p, ortho_p2ch, ortho_p4ch, px_center_projected=heart_center)
    sa_min, sa_mid = get_slice_center_inertia_in_volume_space(label, m_myolvrv, volume_affine, pix_center_sa, label_shape)[:2]
    p2ch_min = get_slice_center_inertia_in_volume_space(label, m_myolvla, volume_affine, pix_p2ch, label_shape)[0]
    pix_4ch = get_pix_affine_from_center_and_plane_vects(myolv_center, sa_min, p2ch_min, px_center_projected=heart_center)
    views["4CH"] = get_torch_grid_affine_from_pix_affine(pix_4ch, label_shape)

    # 6. 2CH (:330-343)
    fch_min = get_slice_center_inertia_in_volume_space(label, m_myolvla, volume_affine, pix_4ch, label_shape)[0]
    pix_2ch = get_pix_affine_from_center_and_plane_vects(myolvla_center, sa_mid, fch_min, px_center_projected=heart_center)
    views["2CH"] = get_torch_grid_affine_from_pix_affine(pix_2ch, label_shape)

    ordered = {k: views[k] for k in ("axial", "sagittal", "coronal", "p2CH", "p4CH", "ALL_SA", "4CH", "2CH")}
    if return_unrolled:
        out = {}
        for name, aff in ordered.items():
            if name == "ALL_SA":
                for i, a in enumerate(aff):
                    out[f"SA-{i}"] = a
            else:
                out[name] = aff
        return out
    return ordered
