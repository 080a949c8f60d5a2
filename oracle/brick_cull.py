"""CPU restatement of the brick-level cull rule of docs/ROUND2_BRICK_CULL.md (section 2) -- test infrastructure for
the round-2 kernel (`k_brick_cull`), usable today to check the RULE itself: a (brick, frame, LiDAR) triple may be
masked only if the reference's visibility test (tools/occ/occ_annotate.py:541-547) fails for every voxel centre of
the brick through that (frame, LiDAR).

All geometry in f64 from the reference's own chain (centre -> ego :490-498 -> sensor :161-164); the footprint is
derived from the 8 corners of the FULL brick only, as a kernel would:
  range    r_lo = min_i v_i.u  <=  |p|  <=  max_i |v_i| = r_hi          (u = unit vector to the brick centre)
  columns  azimuth extremes of a convex body clear of the sensor axis are at its corners
  rows     z is linear and sin(inc) = z / |p|, bracketed with r_lo / r_hi by the sign of z
"""
from __future__ import annotations

import numpy as np

from . import oracle


def _grid_geometry(res, voxel_size):
    size = res["size"].astype(np.float32)
    mb = np.array([-size[0] * np.float32(0.5), -size[1] * np.float32(0.5), 0], np.float32).astype(np.float64)
    return mb, float(voxel_size)


def centre_tests(batch, t, res):
    """Exact per-voxel results for every NON-OCCUPIED voxel of tracklet ``t`` (``res`` = its oracle result):
    -> (vox int64 [n,3], rows [B,L,n], cols [B,L,n], ranges [B,L,n], free bool [B,L,n])."""
    trk = batch.tracklets[t]
    seg = batch.segments[trk.segment]
    mb, vs = _grid_geometry(res, batch.voxel_size)
    vox = np.stack(np.nonzero(res["occ"] != 1), 1).astype(np.int64)
    cen = vox.astype(np.float64) * vs + mb + vs / 2                                  # :467-471
    B, L = len(trk), len(seg.inclinations)
    n = len(vox)
    rows = np.zeros((B, L, n), np.int64)
    cols = np.zeros((B, L, n), np.int64)
    rng = np.zeros((B, L, n), np.float64)
    free = np.zeros((B, L, n), bool)
    for i in range(B):
        box = trk.boxes[i]
        s, c = np.float64(np.sin(np.float32(box[6]))), np.float64(np.cos(np.float32(box[6])))
        ego = cen @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]) + box[:3].astype(np.float64)      # :490-498
        f = int(trk.frame_ids[i])
        for li in range(L):
            ri = seg.range_images[li][f]
            H, W = ri.shape
            idx, r = oracle.point_cloud_to_range_image_idx(ego[None], seg.extrinsics[f:f + 1, li],
                                                           np.ascontiguousarray(seg.inclinations[li][::-1])[None], (H, W))
            rows[i, li], cols[i, li], rng[i, li] = idx[0, :, 0], idx[0, :, 1], r[0]
            free[i, li] = ri[idx[0, :, 0], idx[0, :, 1]].astype(np.float64) >= r[0]                 # :541-547
    return vox, rows, cols, rng, free


def brick_cull_masks(batch, t, res, brick=4, tile=(2, 8)):
    """-> (brick_of_voxel-compatible brick coordinates int64 [nb,3], masked bool [nb, B, L]) for every brick of the
    grid of tracklet ``t``: masked[b, i, c] = the rule of section 2 proves that no centre of brick b can be free
    through frame i / LiDAR c, using a max pyramid of ``tile`` (rows x cols) pixels."""
    trk = batch.tracklets[t]
    seg = batch.segments[trk.segment]
    mb, vs = _grid_geometry(res, batch.voxel_size)
    X, Y, Z = res["occ"].shape
    nbx, nby, nbz = -(-X // brick), -(-Y // brick), -(-Z // brick)
    bricks = np.stack(np.meshgrid(np.arange(nbx), np.arange(nby), np.arange(nbz), indexing="ij"), -1).reshape(-1, 3)
    B, L = len(trk), len(seg.inclinations)
    masked = np.zeros((len(bricks), B, L), bool)
    corner_off = np.array([[x, y, z] for x in (0, brick - 1) for y in (0, brick - 1) for z in (0, brick - 1)], np.float64)
    Rb = 0.5 * np.sqrt(3.0) * (brick - 1) * vs + 1e-3
    tr, tc = tile
    for i in range(B):
        box = trk.boxes[i]
        s, c = np.float64(np.sin(np.float32(box[6]))), np.float64(np.cos(np.float32(box[6])))
        rot_t = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        f = int(trk.frame_ids[i])
        for li in range(L):
            ri = seg.range_images[li][f]
            H, W = ri.shape
            E = seg.extrinsics[f, li].astype(np.float32)
            v2l = np.linalg.inv(E).astype(np.float64)[:3]
            azc = float(np.arctan2(E[1, 0], E[0, 0]))
            incl = seg.inclinations[li][::-1].astype(np.float64)
            for b, (bx, by, bz) in enumerate(bricks):
                lo3 = np.array([bx, by, bz], np.float64) * brick
                cidx = lo3 + 0.5 * (brick - 1)
                pb = v2l[:, :3] @ ((cidx * vs + mb + vs / 2) @ rot_t + box[:3].astype(np.float64)) + v2l[:, 3]
                d, rho = np.linalg.norm(pb), np.hypot(pb[0], pb[1])
                if not (d > 1.25 * Rb and rho > 1.05 * Rb):
                    continue                                            # guard: keep the pair
                pc = (((corner_off + lo3) * vs + mb + vs / 2) @ rot_t + box[:3].astype(np.float64)) @ v2l[:, :3].T + v2l[:, 3]
                u = pb / d
                r_lo, r_hi = float((pc @ u).min()), float(np.linalg.norm(pc, axis=1).max())
                zmin, zmax = float(pc[:, 2].min()), float(pc[:, 2].max())
                inc_hi = np.arcsin(np.clip(zmax / (r_lo if zmax > 0 else r_hi), -1, 1)) + 1e-4
                inc_lo = np.arcsin(np.clip(zmin / (r_hi if zmin > 0 else r_lo), -1, 1)) - 1e-4
                ra = int(np.argmin(np.abs(inc_hi - incl)))
                rb_ = int(np.argmin(np.abs(inc_lo - incl)))
                r0, r1 = max(min(ra, rb_) - 1, 0), min(max(ra, rb_) + 1, H - 1)
                az_c = np.arctan2(pb[1], pb[0])
                azs = az_c + ((np.arctan2(pc[:, 1], pc[:, 0]) - az_c + np.pi) % (2 * np.pi) - np.pi)
                cfa = (W - 0.5) - (azs.max() + azc + 1e-4 + np.pi) / (2 * np.pi) * W
                cfb = (W - 0.5) - (azs.min() + azc - 1e-4 + np.pi) / (2 * np.pi) * W
                q0, q1 = int(np.floor(cfa)) - 1, int(np.ceil(cfb)) + 1
                if q1 - q0 + 1 > W // 2:
                    continue
                cols_idx = np.arange((q0 // tc) * tc, (q1 // tc + 1) * tc) % W
                rr0, rr1 = (r0 // tr) * tr, min((r1 // tr + 1) * tr, H)
                masked[b, i, li] = ri[rr0:rr1][:, cols_idx].max() < r_lo - 1e-3
    return bricks, masked
