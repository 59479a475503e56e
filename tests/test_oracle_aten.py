"""oracle/aten_np.py (numpy restatement of ATen affine_grid / grid_sampler_3d) must equal
torch-CPU ATen bitwise in forward and within tolerance in backward."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import aten_np as A

SHAPES = [(2, 3, 20, 24, 28, 16, 12, 1), (1, 2, 32, 32, 32, 32, 32, 1), (1, 1, 16, 16, 16, 8, 9, 10),
          (1, 1, 5, 6, 7, 1, 1, 1)]


@pytest.mark.parametrize("K", [1, 2, 3, 7, 16, 32, 64, 100, 128, 256])
def test_linspace_and_base_bitwise(K):
    assert np.array_equal(A.linspace_m1_p1(K), torch.linspace(-1, 1, K).numpy())
    g = F.affine_grid(torch.eye(3, 4)[None], [1, 1, 1, 1, K], align_corners=False)[0, 0, 0, :, 0].numpy()
    assert np.array_equal(A.base_coords(K), g)


@pytest.mark.parametrize("shape", SHAPES)
def test_grid_and_sample_bitwise(shape):
    N, C, D, H, W, Do, Ho, Wo = shape
    torch.manual_seed(sum(shape))
    vol = torch.randn(N, C, D, H, W)
    th = torch.eye(3, 4)[None].repeat(N, 1, 1) + 0.3 * torch.randn(N, 3, 4)
    g = F.affine_grid(th, [N, C, Do, Ho, Wo], align_corners=False)
    gn = A.affine_grid_3d(th.numpy(), (Do, Ho, Wo))
    assert np.array_equal(g.numpy(), gn)
    o = F.grid_sample(vol, g, mode="bilinear", padding_mode="zeros", align_corners=False).numpy()
    assert np.array_equal(o, A.grid_sample_3d(vol.numpy(), gn, "bilinear"))
    lab = torch.randint(0, 7, (N, C, D, H, W))
    o = F.grid_sample(lab.float(), g, mode="nearest", padding_mode="zeros", align_corners=False).long().numpy()
    assert np.array_equal(o, A.grid_sample_3d(lab.numpy(), gn, "nearest"))


@pytest.mark.parametrize("shape", SHAPES[:3])
def test_backward_tolerance(shape):
    N, C, D, H, W, Do, Ho, Wo = shape
    torch.manual_seed(sum(shape) + 1)
    vol = torch.randn(N, C, D, H, W, requires_grad=True)
    th = (torch.eye(3, 4)[None].repeat(N, 1, 1) + 0.3 * torch.randn(N, 3, 4)).requires_grad_(True)
    g = F.affine_grid(th, [N, C, Do, Ho, Wo], align_corners=False)
    g.retain_grad()
    out = F.grid_sample(vol, g, mode="bilinear", padding_mode="zeros", align_corners=False)
    go = torch.randn_like(out)
    out.backward(go)
    dv, dg = A.grid_sample_3d_backward(go.numpy(), vol.detach().numpy(), g.detach().numpy())
    assert np.abs(dv - vol.grad.numpy()).max() <= 1e-5 * max(1.0, vol.grad.abs().max().item())
    assert np.abs(dg - g.grad.numpy()).max() <= 1e-5 * g.grad.abs().max().item()
    dth = A.affine_grid_3d_backward(g.grad.numpy())
    assert np.abs(dth - th.grad.numpy()).max() <= 1e-5 * th.grad.abs().max().item()


STRUCTURED = {
    "identity": [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]],
    "flip_x": [[-1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]],
    "swap_xz_flip": [[0, 0, -1, 0], [0, 1, 0, 0], [1, 0, 0, 0]],
    "half_voxel_shift": [[1, 0, 0, 1.0 / 32], [0, 1, 0, -1.0 / 32], [0, 0, 1, 0]],
    "zoom2": [[0.5, 0, 0, 0], [0, 0.5, 0, 0], [0, 0, 0.5, 0]],
}


@pytest.mark.parametrize("name", sorted(STRUCTURED))
@pytest.mark.parametrize("sizes", [((32, 32, 32), (32, 32, 32)), ((32, 32, 32), (8, 8, 8)), ((64, 64, 64), (16, 16, 1))])
def test_structured_affines_tie_policy_bitwise(name, sizes):
    """Coordinates exactly on integers / halves (axis-aligned views, exact down-sampling, half-voxel shifts): the numpy
    restatement takes the same floor / round-half-even decisions as ATen (the CUDA kernels are checked the same way in
    tests/test_gpu_slice.py)."""
    (D, H, W), (Do, Ho, Wo) = sizes
    th = torch.tensor(STRUCTURED[name], dtype=torch.float32)[None]
    torch.manual_seed(D + Do)
    vol = torch.randn(1, 2, D, H, W)
    g = F.affine_grid(th, [1, 2, Do, Ho, Wo], align_corners=False)
    gn = A.affine_grid_3d(th.numpy(), (Do, Ho, Wo))
    assert np.array_equal(g.numpy(), gn)
    o = F.grid_sample(vol, g, mode="bilinear", padding_mode="zeros", align_corners=False).numpy()
    assert np.array_equal(o, A.grid_sample_3d(vol.numpy(), gn, "bilinear"))
    lab = torch.randint(0, 50, (1, 2, D, H, W))
    o = F.grid_sample(lab.float(), g, mode="nearest", padding_mode="zeros", align_corners=False).long().numpy()
    assert np.array_equal(o, A.grid_sample_3d(lab.numpy(), gn, "nearest"))
