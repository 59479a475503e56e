#!/usr/bin/env python
"""CPU simulation of the slicer's memory footprint on the bench workload (no GPU needed).

For the views `bench.py` generates (augmented p2CH composed with random R6 / offset parameters) it rebuilds the sampling
coordinates exactly as the kernels do (F.affine_grid of the grid affine G', un-normalise, floor) and reports, for the
channels-last C = 8 fp32 volume (32-byte voxels, four voxels per 128-byte line):

  * per warp-wide corner gather (8x4 pixel patch, what `pixel_of_tile` assigns to a warp): distinct 128-byte lines and
    32-byte sectors - the quantity the L1 wavefront model of DESIGN.md section 5 is built on;
  * per 16x16 tile: unique voxels touched, number of (y,z) rows, and what a row-run staging scheme (round-2 plan,
    DESIGN.md section 10) would have to hold in shared memory: sum over rows of (max x - min x + 1), in voxels and bytes.

Usage: python profiles/experiments/gather_footprint_sim.py [n_volumes]   (default 8 volumes x 6 views)
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import af_oracle as O  # noqa: E402  (analysis script: test-infrastructure side, not the product)

S, C = bench.S, bench.NUM_CLASSES
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
V = 6
h = bench.make_host_inputs(nv, V, seed=1000)
fov_in = torch.tensor([S, S, S], dtype=torch.float64)
fov_mm_out = torch.tensor([192.0, 192.0, 1.5], dtype=torch.float64)
fov_out = torch.tensor([S, S, 1], dtype=torch.float64)

lines_per_req, sectors_per_req, uniq_vox, rows_per_tile, run_vox, fully_out = [], [], [], [], [], 0
ana_vox, ana_rows, ana_missed = [], [], 0      # analytic row runs (what a kernel can compute without looking at the pixels)


def analytic_runs(ixt, iyt, izt):
    """Row runs of one 16x16 tile from its affine geometry alone.  Within the slicing plane x is a linear function of (y, z):
    x = g + a*y + b*z, so the voxels of row (y, z) that any pixel's 2x2x2 corner block can touch lie in
    [x_c - (|a|+|b|) - 1, x_c + (|a|+|b|) + 1] with x_c = g + a*y + b*z (|iy - y| <= 1, |iz - z| <= 1), clipped to the tile's own
    x-range; rows whose (y, z) box maps outside the tile's pixel rectangle are skipped.  Conservative by construction."""
    p00 = np.array([ixt[0, 0], iyt[0, 0], izt[0, 0]], dtype=np.float64)
    u = (np.array([ixt[15, 0], iyt[15, 0], izt[15, 0]]) - p00) / 15.0
    v = (np.array([ixt[0, 15], iyt[0, 15], izt[0, 15]]) - p00) / 15.0
    xt_lo, xt_hi = int(np.floor(ixt.min())), int(np.floor(ixt.max())) + 1
    ylo, yhi = int(np.floor(iyt.min() - 1e-3)), int(np.floor(iyt.max() + 1e-3)) + 1
    zlo, zhi = int(np.floor(izt.min() - 1e-3)), int(np.floor(izt.max() + 1e-3)) + 1
    det = u[1] * v[2] - v[1] * u[2]
    runs = {}
    for z in range(max(zlo, 0), min(zhi, S - 1) + 1):
        for y in range(max(ylo, 0), min(yhi, S - 1) + 1):
            lo, hi = xt_lo, xt_hi
            if abs(det) > 1e-3:
                minv = np.array([[v[2], -v[1]], [-u[2], u[1]]]) / det          # (iy - Y0, iz - Z0) -> (row idx, col idx)
                d = np.array([y - p00[1], z - p00[2]])
                ab = minv @ d
                slack = np.abs(minv).sum(1) * 1.001 + 1e-3
                if ab[0] + slack[0] < 0 or ab[0] - slack[0] > 15 or ab[1] + slack[1] < 0 or ab[1] - slack[1] > 15:
                    continue                                                    # no pixel of this tile can touch the row
                coef = np.array([u[0], v[0]]) @ minv                            # dx/d(iy), dx/d(iz) inside the plane
                xc = p00[0] + coef @ d
                half = np.abs(coef).sum() * 1.001 + 1e-3
                lo, hi = max(lo, int(np.floor(xc - half))), min(hi, int(np.floor(xc + half)) + 1)
            lo, hi = max(lo, 0), min(hi, S - 1)
            if hi >= lo:
                runs[(z, y)] = (lo, hi)
    return runs
for b in range(nv):
    for v in range(V):
        theta = O.view_theta(h["params"][b:b + 1, v], h["init"][v:v + 1, :6], h["init"][v, 6:9], h["init"][v:v + 1, 9:],
                             bench.OFFSET_CLIP, bench.ZOOM_CLIP, S)
        pre = (h["gpre"][b:b + 1, v] @ theta).double()
        g, _ = O.grid_and_nii_affine(h["nii"][b:b + 1].double(), fov_in, fov_mm_out, fov_out, pre)
        grid = F.affine_grid(g[:, :3].float(), [1, 1, S, S, 1], align_corners=False)[0, :, :, 0]      # [Do,Ho,3] (x,y,z)
        ix = ((grid[..., 0] + 1) * S - 1) / 2
        iy = ((grid[..., 1] + 1) * S - 1) / 2
        iz = ((grid[..., 2] + 1) * S - 1) / 2
        x0, y0, z0 = (torch.floor(t).long().numpy() for t in (ix, iy, iz))
        for tr in range(S // 16):
            for tc in range(S // 16):
                sl = (slice(tr * 16, tr * 16 + 16), slice(tc * 16, tc * 16 + 16))
                X, Y, Z = x0[sl], y0[sl], z0[sl]
                vox = set()
                rows = {}
                any_in = False
                for dz in (0, 1):
                    for dy in (0, 1):
                        for dx in (0, 1):
                            xx, yy, zz = X + dx, Y + dy, Z + dz
                            ok = (xx >= 0) & (xx < S) & (yy >= 0) & (yy < S) & (zz >= 0) & (zz < S)
                            any_in |= bool(ok.any())
                            # warp patches: rows of 4, columns of 8 inside the tile
                            for wr in range(4):
                                for wc in range(2):
                                    m = ok[wr * 4:wr * 4 + 4, wc * 8:wc * 8 + 8]
                                    if not m.any():
                                        continue
                                    a = xx[wr * 4:wr * 4 + 4, wc * 8:wc * 8 + 8][m]
                                    bb = yy[wr * 4:wr * 4 + 4, wc * 8:wc * 8 + 8][m]
                                    cc = zz[wr * 4:wr * 4 + 4, wc * 8:wc * 8 + 8][m]
                                    lines_per_req.append(len(set(zip(cc.tolist(), bb.tolist(), (a // 4).tolist()))))
                                    sectors_per_req.append(len(set(zip(cc.tolist(), bb.tolist(), a.tolist()))))
                            for a, bb, cc in zip(xx[ok].tolist(), yy[ok].tolist(), zz[ok].tolist()):
                                vox.add((cc, bb, a))
                                lo, hi = rows.get((cc, bb), (a, a))
                                rows[(cc, bb)] = (min(lo, a), max(hi, a))
                if not any_in:
                    fully_out += 1
                    continue
                uniq_vox.append(len(vox))
                rows_per_tile.append(len(rows))
                run_vox.append(sum(hi - lo + 1 for lo, hi in rows.values()))
                ar = analytic_runs(ix[sl].numpy(), iy[sl].numpy(), iz[sl].numpy())
                ana_vox.append(sum(hi - lo + 1 for lo, hi in ar.values()))
                ana_rows.append(len(ar))
                ana_missed += sum(1 for (cc, bb, a) in vox if (cc, bb) not in ar or not (ar[(cc, bb)][0] <= a <= ar[(cc, bb)][1]))


def q(a):
    a = np.asarray(a, dtype=np.float64)
    return f"mean {a.mean():.1f}  p50 {np.percentile(a, 50):.0f}  p95 {np.percentile(a, 95):.0f}  max {a.max():.0f}"


print(f"{nv * V} slices, {len(uniq_vox)} tiles with at least one in-bounds corner, {fully_out} tiles completely outside the volume")
print("per warp-wide corner gather (32 pixels):  distinct 128-byte lines:", q(lines_per_req))
print("                                          distinct 32-byte sectors:", q(sectors_per_req))
print("per 16x16 tile: unique voxels touched:   ", q(uniq_vox), f"  (= {np.mean(uniq_vox) / 256:.2f} per pixel, 8 corner reads per pixel)")
print("                (y,z) rows:              ", q(rows_per_tile))
print("                row-run staging, voxels: ", q(run_vox), f"  (= {np.mean(run_vox) * 32 / 1024:.1f} KiB mean, {np.max(run_vox) * 32 / 1024:.1f} KiB max)")
print(f"                staged / unique = {np.sum(run_vox) / np.sum(uniq_vox):.2f};  tiles over 40 KiB: {np.mean(np.asarray(run_vox) * 32 > 40 * 1024) * 100:.1f} %")
print("analytic row runs (from the tile's affine geometry only, no per-pixel pass):")
print("                rows:                    ", q(ana_rows))
print("                staged voxels:           ", q(ana_vox), f"  (= {np.mean(ana_vox) * 32 / 1024:.1f} KiB mean, {np.max(ana_vox) * 32 / 1024:.1f} KiB max)")
print(f"                staged / unique = {np.sum(ana_vox) / np.sum(uniq_vox):.2f};  needed voxels NOT covered (would take the global-load fallback): "
      f"{ana_missed} of {int(np.sum(uniq_vox))};  tiles over 48 KiB: {np.mean(np.asarray(ana_vox) * 32 > 48 * 1024) * 100:.1f} %")
