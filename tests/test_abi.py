"""CPU-side checks of the boundary: the library builds for sm_100a, loads, and exports every symbol the
header declares (no compute calls without a GPU); the product refuses CPU tensors."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_build_and_exports():
    from acquisition_focus_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.lib()
    header = open(os.path.join(ROOT, "include", "afb200.h")).read()
    declared = set(re.findall(r"\b(afb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in afb200.h but not exported"
    assert declared == set(_lib.exported_symbols())
    assert lib.afb_version() == int(re.search(r"#define AFB_VERSION (\d+)", header).group(1))
    assert lib.afb_error_string(-1).decode().startswith("invalid argument")


def test_struct_layout_matches_header():
    import ctypes as C
    from acquisition_focus_b200 import _lib
    assert C.sizeof(_lib.AfbVolume) == 8 + 4 * 6 + 8 * 5
    # afb_views: 2 int, 2 ptr, int(+pad), 3 ptr, 2 int, 2 float, ptr, 3 double
    assert C.sizeof(_lib.AfbViews) == 8 + 16 + 8 + 24 + 8 + 8 + 8 + 24 + 8


def test_argument_errors_without_gpu():
    from acquisition_focus_b200 import _lib
    lib = _lib.lib()
    assert lib.afb_slice_fwd(None, None, 1, 1, 1, 0, 0, 0.0, None, None, None) == -1
    assert lib.afb_embed_fwd(None, None, 1, 1, 1, 4, None, None, None) == -1
    assert lib.afb_r6_fwd(None, 1, None, None) == -1
    assert lib.afb_volume_min(None, 0, 10, None, None, None) == -1


def test_no_cpu_fallback():
    import acquisition_focus_b200 as afb
    from acquisition_focus_b200._lib import AfbError
    with pytest.raises(AfbError):
        afb.nifti_grid_sample(torch.zeros(1, 1, 4, 4, 4), torch.eye(4)[None].double())
    with pytest.raises(AfbError):
        afb.compute_rotation_matrix_from_ortho6d(torch.randn(2, 6))
    with pytest.raises(AfbError):
        afb.SkipConnector(1)(torch.zeros(1, 1, 4, 4), [torch.eye(4)[None]])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "acquisition_focus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_host_label_packing_runs_without_a_gpu():
    """afb_host_narrow_labels is host code: int64 / int32 / int16 -> uint8 over several threads, ragged lengths, unaligned
    sources, and the out-of-range report (negative values and values above 255)."""
    import ctypes as C
    import torch
    from acquisition_focus_b200 import _lib as L
    from acquisition_focus_b200.running.host_input import narrow_labels_host
    g = torch.Generator().manual_seed(3)
    for dt in (torch.int64, torch.int32, torch.int16):
        for n in (1, 15, 16, 1000003):
            src = torch.randint(0, 256, (n + 1,), generator=g, dtype=dt)
            for off in (0, 1):
                x = src[off:off + n].contiguous() if off == 0 else src[off:off + n]
                out = torch.empty(n, dtype=torch.uint8)
                narrow_labels_host(x, out, n_threads=5)
                assert torch.equal(out, x.to(torch.uint8))
    x = torch.zeros(200000, dtype=torch.int64)
    out = torch.empty(200000, dtype=torch.uint8)
    for bad_value in (256, -1, 1 << 40):
        x[150001] = bad_value
        with pytest.raises(ValueError):
            narrow_labels_host(x, out, n_threads=4)
    bad = C.c_int(7)
    assert L.lib().afb_host_narrow_labels(x.data_ptr(), L.DTYPES[torch.float32], 10, out.data_ptr(), 1, C.byref(bad)) != 0


def test_split_upload_share_balances_cores_and_link():
    """HostInputPipeline's split upload: at the returned share the packing pass and the link take the same time (or the share
    saturates at 0 / 1)."""
    from acquisition_focus_b200.running.host_input import balanced_pack_fraction
    L_, I_, e = 1.07e9, 0.54e9, 8
    for r_pack, r_link in ((41e9, 55e9), (15e9, 16.5e9), (5e9, 55e9), (20e9, 20e9)):
        f = balanced_pack_fraction(I_, L_, e, r_pack, r_link)
        assert 0.0 < f < 1.0
        t_cores = f * L_ / r_pack
        t_link = (I_ + (1 - f) * L_ + f * L_ / e) / r_link
        assert abs(t_cores - t_link) < 1e-9 * max(t_cores, 1.0) + 1e-6 * t_link
    assert balanced_pack_fraction(I_, L_, e, 92e9, 50e9) == 1.0          # enough cores: pack everything (N = 1 on the B200 box)
    assert balanced_pack_fraction(I_, L_, e, 0.0, 50e9) == 0.0
    assert balanced_pack_fraction(0.0, 0.0, e, 1e9, 1e9) == 0.0
