"""CPU-side checks of the host logic around the kernels (no CUDA needed): layout detection, sharding, the clinical
composition / augmentation helpers of running/model_input.py against the oracle restatement, and the module mirror's
parameter bookkeeping."""
import pytest
import torch

from oracle import af_oracle as O


def test_is_dense_layout_detection():
    from acquisition_focus_b200.functional import _is_dense
    x = torch.zeros(2, 8, 5, 6, 7)
    assert _is_dense(x)
    assert _is_dense(x.permute(0, 2, 3, 4, 1))                        # channels-last view of a compact block
    assert _is_dense(torch.nn.functional.one_hot(torch.zeros(2, 3, 4, 5, dtype=torch.long), 8).permute(0, 4, 1, 2, 3))
    assert _is_dense(torch.zeros(2, 8, 5, 6, 1).expand(2, 8, 5, 6, 1)) and _is_dense(torch.zeros(2, 1, 5, 6, 7)[:, 0][:, None])
    assert not _is_dense(x[:, ::2])                                   # strided channel slice: holes
    assert not _is_dense(x[..., :3])                                  # cropped last axis: holes
    assert not _is_dense(x[:, :1].expand(2, 8, 5, 6, 7))              # broadcast (overlapping) channels


def test_shard_range_partitions_exactly():
    from acquisition_focus_b200.parallel import shard_range
    for n in (1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_random_aug_affine_matches_oracle_stream():
    """product copy (synthetic.random_aug_affine, used by apply_affine_augmentation) == oracle restatement of
    utils/transform_utils.py:6-23 for the same generator state."""
    from acquisition_focus_b200.synthetic import random_aug_affine
    for seed in (0, 5):
        a = random_aug_affine(torch.Generator().manual_seed(seed), 0.3, 0.2, 0.1)
        b = O.random_aug_affine(torch.Generator().manual_seed(seed), 0.3, 0.2, 0.1)
        assert torch.equal(a, b)


def test_input_affine_and_augmentation_helpers():
    from acquisition_focus_b200.running import model_input as MI

    class Atm:
        view_id = "p2CH"
        random_grid_affine = torch.eye(4)[None] * 2.0
    B = 3
    gen = torch.Generator().manual_seed(1)
    base = torch.eye(4)[None].repeat(B, 1, 1) + 0.1 * torch.randn(B, 4, 4, generator=gen)
    base[:, 3] = torch.tensor([0.0, 0, 0, 1])
    view = torch.eye(4)[None].repeat(B, 1, 1) + 0.1 * torch.randn(B, 4, 4, generator=gen)
    got = MI.get_input_affine_for_atm(Atm(), base, {"p2CH": view})
    assert torch.allclose(got, O.input_affine_for_view(base, view), atol=1e-6)
    rnd = Atm(); rnd.view_id = "RND"
    assert torch.equal(MI.get_input_affine_for_atm(rnd, base, {}), (torch.eye(4)[None] * 2.0).repeat(B, 1, 1))
    # one random affine per batch element, shared by every affine of the list (run_dl.py:208-223)
    lst = [torch.eye(4)[None].repeat(B, 1, 1), 2.0 * torch.eye(4)[None].repeat(B, 1, 1)]
    out = MI.apply_affine_augmentation(lst, generator=torch.Generator().manual_seed(3))
    assert torch.allclose(out[1], 2.0 * out[0]) and not torch.allclose(out[0][0], out[0][1])


def test_module_mirror_bookkeeping():
    """constructor arguments, vox_range / arra (learnable_transform.py:112-116) and ap_space per optim_method."""
    import acquisition_focus_b200 as afb
    fov = torch.tensor([192.0, 192.0, 192.0])
    for method, ap in (("R6-vector", 6), ("angle-axis", 3), ("normal-vector", 3)):
        m = afb.AffineTransformModule(8, fov, torch.tensor([128, 128, 128]), torch.tensor([192.0, 192.0, 1.5]),
                                      torch.tensor([128, 128, 1]), optim_method=method, offset_clip_value=0.2, zoom_clip_value=0.0,
                                      localization_net=torch.nn.Identity())
        assert m.ap_space == ap and m.vox_range == O.offset_vox_range(0.2, 128) == 26
        assert torch.equal(m.arra, O.offset_positions(128, 26))
        assert m.init_theta_ap.numel() == ap and not m.init_theta_ap.requires_grad
    with pytest.raises(AssertionError):
        afb.AffineTransformModule(8, fov, torch.tensor([128, 128, 128]), fov, torch.tensor([128, 128, 1]), optim_method="euler")
    with pytest.raises(NotImplementedError):
        afb.AffineTransformModule(8, fov, torch.tensor([128, 128, 128]), fov, torch.tensor([128, 128, 1]), optim_method="R6-vector",
                                  align_corners=True)
