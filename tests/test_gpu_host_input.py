"""One-hot materialisation fused with the min record (run_dl.py:261-264 + nifti_utils.py:200) and the pipelined host
upload built on it: bit-exact against torch's one_hot / .float(), and the acquisition that takes the record instead of
re-scanning the soft volume must give bitwise the same slices and dVolume as the plain path."""
import pytest
import torch
import torch.nn.functional as F

from oracle import cases

pytestmark = pytest.mark.gpu
INIT = torch.tensor([[1e-2, 0, 0, 0, 1e-2, 0, 0, 0, 0, 1.0]])


@pytest.fixture(scope="module")
def afb():
    import acquisition_focus_b200 as m
    return m


@pytest.mark.parametrize("shape,C,dt", [((2, 16, 16, 16), 8, torch.int64), ((1, 7, 9, 11), 8, torch.uint8),
                                        ((3, 8, 8, 8), 5, torch.int32), ((1, 5, 5, 5), 3, torch.int16),
                                        ((2, 32, 32, 32), 2, torch.int64)])
def test_expand_bit_exact_vs_torch(afb, shape, C, dt):
    from acquisition_focus_b200 import functional as AF
    gen = torch.Generator().manual_seed(sum(shape) + C)
    lab = torch.randint(0, C, shape, generator=gen).to(dt).cuda()
    oh = torch.empty(shape + (C,), dtype=torch.int64, device="cuda")
    soft = torch.empty(shape + (C,), dtype=torch.float32, device="cuda")
    rec = AF.min_record_alloc(soft.numel(), "cuda")
    AF.onehot_expand(lab, C, out_label=oh, out_soft=soft, record=rec)
    ref = F.one_hot(lab.long(), C)
    assert torch.equal(oh, ref) and torch.equal(soft, ref.float())
    mc = AF.min_count_from_record(rec, soft.numel())
    ref_mc = AF.volume_min(soft)
    assert torch.equal(mc, ref_mc)
    assert mc[0].item() == 0.0 and mc[1].item() == float((ref == 0).sum().item())
    # either output alone
    oh2, soft2 = torch.empty_like(oh), torch.empty_like(soft)
    AF.onehot_expand(lab, C, out_label=oh2)
    AF.onehot_expand(lab, C, out_soft=soft2)
    assert torch.equal(oh2, ref) and torch.equal(soft2, ref.float())


def test_expand_all_one_class_and_out_of_range(afb):
    """C = 1: every value is 1, so the minimum is 1 with full multiplicity; labels outside [0, C) give zero voxels."""
    from acquisition_focus_b200 import functional as AF
    lab = torch.zeros(1, 8, 8, 8, dtype=torch.int64, device="cuda")
    soft = torch.empty(1, 8, 8, 8, 1, device="cuda")
    rec = AF.min_record_alloc(soft.numel(), "cuda")
    AF.onehot_expand(lab, 1, out_soft=soft, record=rec)
    assert AF.min_count_from_record(rec, soft.numel()).tolist() == [1.0, 512.0]
    lab = torch.tensor([-1, 0, 3, 9], dtype=torch.int64, device="cuda").view(1, 1, 2, 2)
    soft = torch.empty(1, 1, 2, 2, 4, device="cuda")
    AF.onehot_expand(lab, 4, out_soft=soft)
    assert soft.view(4, 4).tolist() == [[0, 0, 0, 0], [1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 0, 0]]


def test_ranged_expand_equals_whole(afb):
    from acquisition_focus_b200 import functional as AF
    C, shape = 8, (6, 16, 16, 16)
    lab = torch.randint(0, C, shape, generator=torch.Generator().manual_seed(3)).cuda()
    soft_a, soft_b = (torch.empty(shape + (C,), device="cuda") for _ in range(2))
    rec_a, rec_b = (AF.min_record_alloc(soft_a.numel(), "cuda").zero_() for _ in range(2))
    AF.onehot_expand(lab, C, out_soft=soft_a, record=rec_a)
    per = lab[0].numel() * C
    for b0, b1 in ((0, 2), (2, 3), (3, 6)):
        AF.onehot_expand(lab[b0:b1], C, out_soft=soft_b[b0:b1], record=rec_b, total_elements=soft_b.numel(), elem_offset=b0 * per)
    assert torch.equal(soft_a, soft_b) and torch.equal(rec_a, rec_b)
    with pytest.raises(Exception):      # a range that is not chunk-aligned is refused
        AF.onehot_expand(lab[:1], C, out_soft=soft_b[:1], record=rec_b, total_elements=soft_b.numel(), elem_offset=100)


@pytest.mark.parametrize("group", [1, 2, 4])
def test_upload_pipeline_matches_plain_path(afb, group):
    """Pinned host batch -> upload_one_hot -> acquire_views(soft_pad=record-derived pad) must equal (slices bitwise,
    gradients up to the order of the atomics) the plain path on tensors built with torch's one_hot on the device."""
    from acquisition_focus_b200.running.host_input import upload_one_hot
    S, B, V = 32, 4, 3
    case = cases.atm_case(S, B, V, seed=11)
    kw = dict(offset_clip=case["offset_clip"], zoom_clip=case["zoom_clip"], spat=S, slice_fov_mm=case["slice_fov_mm"].tolist(),
              slice_fov_vox=case["slice_fov_vox"].tolist())
    gpre = torch.stack(case["gpre"], 1).cuda()
    init = INIT.repeat(V, 1).cuda()
    p1 = torch.stack(case["params"], 1).cuda().requires_grad_(True)
    p2 = p1.detach().clone().requires_grad_(True)
    soft1 = case["soft"].cuda().requires_grad_(True)
    plain = afb.acquire_views(soft1, case["label"].cuda(), case["image"].cuda(), case["nii"].cuda(), gpre, p1, init, **kw)

    db = upload_one_hot(case["lab"].pin_memory(), case["image"].pin_memory(), 8, "cuda", group_volumes=group)
    assert torch.equal(db.label, case["label"].cuda()) and torch.equal(db.soft_label, case["soft"].cuda())
    assert db.soft_label.stride() == case["soft"].cuda().stride() and torch.equal(db.image, case["image"].cuda())
    soft2 = db.soft_label.requires_grad_(True)
    piped = afb.acquire_views(soft2, db.label, db.image, case["nii"].cuda(), gpre, p2, init, soft_pad=db.soft_pad,
                              image_pad=db.image_pad, **kw)
    for a, b in zip(plain[:4], piped[:4]):
        assert torch.equal(a, b)
    go = cases.pattern(plain[0].shape, 1.0).cuda()
    (plain[0] * go).sum().backward()
    (piped[0] * go).sum().backward()
    # same kernels on the same data; only the order of the floating-point atomics may differ between two launches
    assert (soft1.grad - soft2.grad).abs().max().item() <= 1e-6 * soft1.grad.abs().max().item()
    assert torch.equal(soft1.grad == 0, soft2.grad == 0)
    assert (p1.grad - p2.grad).abs().max().item() <= 1e-6 * p1.grad.abs().max().item()


@pytest.mark.parametrize("n", [512 * 7, 4099, 2 ** 20 + 13])
def test_record_of_arbitrary_fp32_volume(afb, n):
    """[min, multiplicity] rebuilt from the record alone equals the direct pass, ties and ragged tails included."""
    from acquisition_focus_b200 import functional as AF
    gen = torch.Generator().manual_seed(n)
    x = torch.randint(-3, 9, (n,), generator=gen).float().cuda()          # many ties at the minimum
    direct = AF.volume_min(x, with_mask=True)
    again = AF.min_count_from_record(direct._afb_mask, n)
    assert torch.equal(direct, again) and direct[0].item() == -3.0 and direct[1].item() == float((x == -3).sum().item())


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_half_dvolume_cast_is_round_to_nearest_even(afb, dt):
    """dVolume of a half-precision volume: fp32 accumulation, then ONE flat cast kernel == torch's .to(dtype), strides kept."""
    import ctypes as C
    from acquisition_focus_b200 import _lib as L
    x = (torch.randn(3, 5, 7, 11, 8, generator=torch.Generator().manual_seed(9)) * 300).cuda()      # 9240 elements: ragged tail
    x[0, 0, 0, 0, :4] = torch.tensor([65504.0, 1e-8, -0.0, 3.0e38]).cuda()
    out = torch.empty_like(x, dtype=dt)
    L.check(L.lib().afb_cast_from_f32(L.ptr(x), L.ptr(out), L.DTYPES[dt], x.numel(), L.stream_ptr(x.device)), "cast")
    assert torch.equal(out, x.to(dt))


@pytest.mark.parametrize("narrow", [True, False, "split"])
def test_double_buffered_pipeline_equals_single_upload(afb, narrow):
    """HostInputPipeline (next batch uploaded + expanded while the current one is used; slots reused two submits later) hands
    over, batch after batch, bitwise what upload_one_hot returns, also when a slow consumer delays the slot release - with the
    int64 label maps packed to uint8 on the host by a worker thread (the default) and with the plain int64 upload."""
    from acquisition_focus_b200.running.host_input import HostInputPipeline, upload_one_hot
    C, B, S = 8, 4, 32
    batches = []
    for k in range(5):
        lab = cases.randint(0, C, (B, S, S, S), 700 + k).pin_memory()
        img = cases.randn((B, 1, S, S, S), 800 + k).pin_memory()
        batches.append((lab, img))
    pipe = HostInputPipeline(C, "cuda", depth=2, group_volumes=2, narrow_labels=narrow, narrow_threads=3)
    pipe.submit(*batches[0])
    if narrow != "split":
        assert pipe.h2d_bytes_last == batches[0][0].numel() * (1 if narrow else 8) + batches[0][1].numel() * 4
    else:
        pipe.pack_fraction = 0.5                # two of the four volumes packed, two uploaded as int64
    spin = torch.empty(64 * 1024 * 1024, device="cuda")
    for k in range(5):
        if k + 1 < 5:
            pipe.submit(*batches[k + 1])
        db = pipe.get()
        assert db.label_map.dtype == (torch.uint8 if narrow else torch.int64)
        got = [db.label_map.long(), db.label.clone(), db.soft_label.clone(), db.image.clone(), db.soft_pad.clone(), db.image_pad.clone()]
        for _ in range(3):
            spin.add_(1.0)                      # the consumer keeps the compute stream busy before it releases the slot
        chk = db.soft_label.sum()               # ... and still reads the slot afterwards
        pipe.release(db)
        ref = upload_one_hot(batches[k][0], batches[k][1], C, "cuda", group_volumes=2)
        want = [ref.label_map, ref.label, ref.soft_label, ref.image, ref.soft_pad, ref.image_pad]
        for a, b in zip(got, want):
            assert torch.equal(a, b)
        assert chk.item() == ref.soft_label.sum().item()


def test_pipeline_rejects_labels_that_do_not_fit_a_byte(afb):
    from acquisition_focus_b200.running.host_input import HostInputPipeline
    lab = cases.randint(0, 8, (2, 16, 16, 16), 5).pin_memory()
    lab[1, 3, 4, 5] = 300
    pipe = HostInputPipeline(8, "cuda", depth=2, group_volumes=1)
    pipe.submit(lab, None)
    with pytest.raises(ValueError, match="outside"):
        pipe.get()
