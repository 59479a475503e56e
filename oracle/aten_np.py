"""ORACLE (test infrastructure, never shipped, never measured as the product).

Bit-level numpy/float32 restatement of the three PyTorch ATen primitives the
reference's view-acquisition path delegates its arithmetic to.  The reference
(`/root/reference`, multimodallearning/acquisition-focus) holds no arithmetic
of its own for sampling: it calls

* ``F.affine_grid``  at ``acquisition_focus/utils/nifti_utils.py:182`` and
  ``acquisition_focus/models/hybrid_unet.py:85``
* ``F.grid_sample``  at ``acquisition_focus/utils/nifti_utils.py:90,93`` and
  ``acquisition_focus/models/hybrid_unet.py:88``

which live in a third-party dependency that is not vendored under
``/root/reference``: **PyTorch ATen, pinned ``torch==2.0.0``**
(``pyproject.toml:11``, ``poetry.lock:3249-3250``); this image carries torch
2.11.0.  The published algorithm restated here is ATen's
``affine_grid_generator`` (5-D, ``align_corners=False``:
``linspace(-1,1,K)*(K-1)/K`` base grid times ``theta^T``) and
``grid_sampler_3d`` forward/backward (``padding_mode='zeros'``,
``align_corners=False``, modes ``bilinear`` and ``nearest``); the semantics
headers shipped in the wheel are ``torch/include/ATen/native/GridSampler.h``
(unnormalize ``((x+1)*size-1)/2``, ``:27-36``; source index ``:164-171``).

Every float32 operation below is performed as a separate IEEE round-to-nearest
numpy float32 operation (no fused multiply-add) *except* where torch's CPU
build fuses: ``linspace`` (``start + step*i`` is one FMA) and the ``bmm`` inside
``affine_grid`` (an FMA chain in k order, accumulator starting at ``0``).
Those two facts were established empirically against torch-CPU in this
container (``tests/test_oracle_aten.py`` keeps the check alive: the functions
here must equal torch-CPU **bitwise**).

Parity pinning: the reference ships no tests/golden vectors (``tests/`` holds
an empty ``__init__.py``), so this restatement is pinned against outputs of
the reference itself (run here on CPU, ``oracle/make_golden.py`` ->
``tests/golden/*.npz``) and against torch-CPU ATen at test time.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def _fma(a, b, c):
    """float32 fused multiply-add emulated in float64 (a*b is exact in f64)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    return (a * b + c).astype(F32)


def linspace_m1_p1(K: int) -> np.ndarray:
    """``torch.linspace(-1, 1, K)`` in float32, bitwise.

    ATen ``RangeFactories``: ``step=(end-start)/(K-1)``; first half
    ``start + step*i``, second half ``end - step*(K-1-i)``, each compiled to a
    single FMA (verified: the non-fused form differs in 42/128 positions).
    """
    if K == 1:
        return np.array([-1.0], dtype=F32)
    step = F32(F32(2.0) / F32(K - 1))
    i = np.arange(K)
    lo = _fma(step, i.astype(F32), F32(-1.0))
    hi = _fma(-step, (K - 1 - i).astype(F32), F32(1.0))
    return np.where(i < K // 2, lo, hi).astype(F32)


def base_coords(K: int) -> np.ndarray:
    """Normalised base coordinate of ``affine_grid(align_corners=False)``.

    ``linspace(-1,1,K) * (K-1) / K`` as two separate float32 ops (ATen
    ``make_base_grid``); mathematically ``(2k+1)/K - 1`` but *not* bitwise.
    """
    lin = linspace_m1_p1(K)
    return ((lin * F32(K - 1)).astype(F32) / F32(K)).astype(F32)


def affine_grid_3d(theta: np.ndarray, size_dhw) -> np.ndarray:
    """``F.affine_grid(theta[N,3,4], [N,C,D,H,W], align_corners=False)``.

    Returns ``grid[N,D,H,W,3]`` (last dim x,y,z -> W,H,D) float32, bitwise equal
    to torch-CPU: ``g_r = fma(1,t_r3, fma(z,t_r2, fma(y,t_r1, fma(x,t_r0, 0))))``.
    """
    theta = np.asarray(theta, dtype=F32)
    N = theta.shape[0]
    D, H, W = (int(s) for s in size_dhw)
    x = np.broadcast_to(base_coords(W)[None, None, :], (D, H, W))
    y = np.broadcast_to(base_coords(H)[None, :, None], (D, H, W))
    z = np.broadcast_to(base_coords(D)[:, None, None], (D, H, W))
    grid = np.empty((N, D, H, W, 3), dtype=F32)
    for n in range(N):
        for r in range(3):
            acc = (x * theta[n, r, 0]).astype(F32)          # fma(x, a, 0)
            acc = _fma(y, theta[n, r, 1], acc)
            acc = _fma(z, theta[n, r, 2], acc)
            acc = (acc + theta[n, r, 3]).astype(F32)        # fma(1, d, acc)
            grid[n, ..., r] = acc
    return grid


def unnormalize(coord: np.ndarray, size: int) -> np.ndarray:
    """``grid_sampler_unnormalize(align_corners=False)``: ((c+1)*size-1)/2."""
    c = np.asarray(coord, dtype=F32)
    return ((((c + F32(1.0)).astype(F32) * F32(size)).astype(F32) - F32(1.0)).astype(F32) / F32(2.0)).astype(F32)


def _corner_setup(grid, D, H, W):
    ix = unnormalize(grid[..., 0], W)
    iy = unnormalize(grid[..., 1], H)
    iz = unnormalize(grid[..., 2], D)
    x0 = np.floor(ix).astype(np.int64)
    y0 = np.floor(iy).astype(np.int64)
    z0 = np.floor(iz).astype(np.int64)
    return ix, iy, iz, x0, y0, z0


# ATen corner order: tnw, tne, tsw, tse, bnw, bne, bsw, bse  (dx, dy, dz)
_CORNERS = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)]


def _weights(ix, iy, iz, x0, y0, z0):
    """Corner weights as ATen forms them: products of corner-coordinate
    differences, ``(wx*wy)*wz`` left to right, opposite corner as float."""
    x1f = (x0 + 1).astype(F32); y1f = (y0 + 1).astype(F32); z1f = (z0 + 1).astype(F32)
    x0f = x0.astype(F32); y0f = y0.astype(F32); z0f = z0.astype(F32)
    wx = [(x1f - ix).astype(F32), (ix - x0f).astype(F32)]
    wy = [(y1f - iy).astype(F32), (iy - y0f).astype(F32)]
    wz = [(z1f - iz).astype(F32), (iz - z0f).astype(F32)]
    return wx, wy, wz


def grid_sample_3d(vol: np.ndarray, grid: np.ndarray, mode: str = "bilinear") -> np.ndarray:
    """``F.grid_sample(vol[N,C,D,H,W], grid[N,Do,Ho,Wo,3], mode, 'zeros', False)``.

    Bitwise restatement of ATen's CPU ``grid_sampler_3d`` (non-fused mul/add,
    accumulation order tnw..bse, out-of-bounds corners skipped); ``nearest``
    uses ``nearbyint`` (round-half-to-even).
    """
    vol = np.asarray(vol)
    N, C, D, H, W = vol.shape
    ix, iy, iz, x0, y0, z0 = _corner_setup(np.asarray(grid, dtype=F32), D, H, W)
    out_shape = (N, C) + grid.shape[1:4]
    if mode == "nearest":
        xn = np.rint(ix).astype(np.int64); yn = np.rint(iy).astype(np.int64); zn = np.rint(iz).astype(np.int64)
        inb = (xn >= 0) & (xn < W) & (yn >= 0) & (yn < H) & (zn >= 0) & (zn < D)
        out = np.zeros(out_shape, dtype=vol.dtype)
        for n in range(N):
            m = inb[n]
            out[n][:, m] = vol[n][:, zn[n][m], yn[n][m], xn[n][m]]
        return out
    assert mode == "bilinear"
    vol = vol.astype(F32, copy=False)
    wx, wy, wz = _weights(ix, iy, iz, x0, y0, z0)
    out = np.zeros(out_shape, dtype=F32)
    for dx, dy, dz in _CORNERS:
        w = ((wx[dx] * wy[dy]).astype(F32) * wz[dz]).astype(F32)
        xc = x0 + dx; yc = y0 + dy; zc = z0 + dz
        inb = (xc >= 0) & (xc < W) & (yc >= 0) & (yc < H) & (zc >= 0) & (zc < D)
        for n in range(N):
            m = inb[n]
            v = vol[n][:, zc[n][m], yc[n][m], xc[n][m]]                     # [C, K]
            out[n][:, m] = (out[n][:, m] + (v * w[n][m][None]).astype(F32)).astype(F32)
    return out


def grid_sample_3d_backward(grad_out: np.ndarray, vol: np.ndarray, grid: np.ndarray):
    """Backward of bilinear ``grid_sample`` (ATen ``grid_sampler_3d_backward``).

    Returns ``(d_vol[N,C,D,H,W], d_grid[N,Do,Ho,Wo,3])``.  Accumulates in
    float64 and rounds once (ATen accumulates in float32 in a data-dependent
    order; gradients are tolerance-checked, not bitwise).
    """
    vol = np.asarray(vol, dtype=F32)
    N, C, D, H, W = vol.shape
    go = np.asarray(grad_out, dtype=np.float64)
    ix, iy, iz, x0, y0, z0 = _corner_setup(np.asarray(grid, dtype=F32), D, H, W)
    wx, wy, wz = _weights(ix, iy, iz, x0, y0, z0)
    wx = [a.astype(np.float64) for a in wx]; wy = [a.astype(np.float64) for a in wy]; wz = [a.astype(np.float64) for a in wz]
    d_vol = np.zeros(vol.shape, dtype=np.float64)
    gx = np.zeros(ix.shape, dtype=np.float64); gy = np.zeros_like(gx); gz = np.zeros_like(gx)
    for dx, dy, dz in _CORNERS:
        xc = x0 + dx; yc = y0 + dy; zc = z0 + dz
        inb = (xc >= 0) & (xc < W) & (yc >= 0) & (yc < H) & (zc >= 0) & (zc < D)
        sx = 1.0 if dx else -1.0; sy = 1.0 if dy else -1.0; sz = 1.0 if dz else -1.0
        for n in range(N):
            m = inb[n]
            g = go[n][:, m]                                                  # [C,K]
            w = (wx[dx][n][m] * wy[dy][n][m] * wz[dz][n][m])[None]
            flat = (zc[n][m] * H + yc[n][m]) * W + xc[n][m]
            for c in range(C):
                np.add.at(d_vol[n, c].reshape(-1), flat, (g[c] * w[0]))
            v = vol[n][:, zc[n][m], yc[n][m], xc[n][m]].astype(np.float64)    # [C,K]
            s = (v * g).sum(0)
            gx[n][m] += sx * s * wy[dy][n][m] * wz[dz][n][m]
            gy[n][m] += sy * s * wx[dx][n][m] * wz[dz][n][m]
            gz[n][m] += sz * s * wx[dx][n][m] * wy[dy][n][m]
    d_grid = np.stack([gx * (W / 2.0), gy * (H / 2.0), gz * (D / 2.0)], axis=-1)
    return d_vol.astype(F32), d_grid.astype(F32)


def affine_grid_3d_backward(d_grid: np.ndarray) -> np.ndarray:
    """``d_theta[N,3,4] = sum_voxels d_grid[.,r] * (x,y,z,1)`` (the dTheta reduce)."""
    d_grid = np.asarray(d_grid, dtype=np.float64)
    N, D, H, W, _ = d_grid.shape
    x = np.broadcast_to(base_coords(W)[None, None, :], (D, H, W)).astype(np.float64)
    y = np.broadcast_to(base_coords(H)[None, :, None], (D, H, W)).astype(np.float64)
    z = np.broadcast_to(base_coords(D)[:, None, None], (D, H, W)).astype(np.float64)
    base = np.stack([x, y, z, np.ones_like(x)], axis=-1)                      # [D,H,W,4]
    return np.einsum("ndhwr,dhwk->nrk", d_grid, base).astype(F32)
