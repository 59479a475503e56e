"""ORACLE / test infrastructure: numpy / torch-CPU restatements of the three DEVICE passes of the GPU drop-in for
get_clinical_cardiac_view_affines (acquisition_focus_b200/clinical_cardiac_views.py): group moments
(utils/torch_sparse_tensor_utils.py:34-56), extent bisection (functional/clinical_cardiac_views.py:36-62).  The nearest slice is
``af_oracle.nifti_grid_sample``.  With these patched in, the product's host logic runs on CPU and is checked against the
reference function itself (tests/test_clinical_views_host_logic.py)."""
import numpy as np
import torch


def cpu_moments(label, masks):
    lab = label.cpu().numpy().astype(np.int64)
    D, H, W = lab.shape
    G = len(masks)
    cnt = np.zeros(G)
    centers = torch.zeros(G, 3)
    inertia = torch.zeros(G, 3, 3)
    for g, m in enumerate(masks):
        sel = np.zeros_like(lab, dtype=bool)
        for l in range(1, 32):
            if (m >> l) & 1:
                sel |= lab == l
        idx = np.argwhere(sel).astype(np.float64)
        n = len(idx)
        cnt[g] = n
        if n == 0:
            continue
        s1 = idx.sum(0)
        s2 = idx.T @ idx
        c32 = torch.tensor(s1 / n, dtype=torch.float32)
        c = c32.double().numpy()
        cov = s2 - np.outer(c, s1) - np.outer(s1, c) + n * np.outer(c, c)
        centers[g] = c32
        inertia[g] = torch.from_numpy(np.trace(cov) * np.eye(3) - cov).float()
    return cnt, centers, inertia


def cpu_extent(label, mask, center, direction):
    lab = label.cpu().numpy().astype(np.int64)
    sel = np.zeros_like(lab, dtype=bool)
    for l in range(1, 32):
        if (mask >> l) & 1:
            sel |= lab == l
    idx = torch.from_numpy(np.argwhere(sel)).float()
    init_end = torch.linalg.vector_norm(torch.as_tensor(label.shape, dtype=torch.float), 2).item()
    out = []
    for d in (direction, -direction):
        start, end = 0.0, init_end
        while (end - start) > 1.73 / 2:
            new_end = end - (end - start) / 2.0
            dist = torch.linalg.vector_norm(idx - (center + new_end * d).view(1, 3), 2, dim=1).min()
            if dist > 1.73 / 2:
                end = new_end
            else:
                start += (end - start) / 2.0
        out.append(center + (start + end) / 2.0 * d)
    return out[0], out[1]
