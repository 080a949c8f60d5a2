#!/usr/bin/env python
"""CPU estimate for a round-2 idea: how many visibility tests of the voxels that stay
UNKNOWN could be removed by an exact, conservative cull of (4x4x4-voxel brick, frame, LiDAR) triples?

For a brick and a (frame, LiDAR) pair: project its 64 voxel centres exactly (oracle), take the bounding pixel
rectangle of their (row, col), padded by one pixel and snapped outwards to `tile` (rows x cols) pyramid tiles;
if the largest return in that rectangle is below the smallest range of the brick's centres, `ri >= range` is
false for every centre of the brick through that pair.  Prints the fraction of (unknown voxel, pair) tests that
such a cull removes, for several pyramid tile sizes.  Uses only the oracle (no GPU).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from objectcentricocccompletion_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402


def main(n_trk=3, n_frames=20, brick=4, seed=0):
    batch = synth.make_batch(n_trk, n_frames, 0.2, "vehicle", seed=seed)
    res = oracle.annotate_batch(batch, threads=8)
    vs = batch.voxel_size
    tiles = [(1, 1), (2, 8), (4, 16), (8, 32)]
    tot = 0
    culled = {t: 0 for t in tiles}
    culled_ball = {t: 0 for t in tiles}
    culled_corner = {t: 0 for t in tiles}
    pair_culled = 0
    dense_vox = all_vox = 0
    for trk, r in zip(batch.tracklets, res):
        if r["occ"] is None:
            continue
        seg = batch.segments[trk.segment]
        occ = r["occ"]
        X, Y, Z = occ.shape
        size = r["size"].astype(np.float32)
        mb = np.array([-size[0] * np.float32(0.5), -size[1] * np.float32(0.5), 0], np.float32).astype(np.float64)
        ux, uy, uz = np.nonzero(occ == 0)                       # voxels that stay unknown
        bid = ((ux // brick) * 1000 + (uy // brick)) * 1000 + (uz // brick)
        order = np.argsort(bid, kind="stable")
        ux, uy, uz, bid = ux[order], uy[order], uz[order], bid[order]
        starts = np.flatnonzero(np.r_[True, bid[1:] != bid[:-1]])
        ends = np.r_[starts[1:], len(bid)]
        cen = np.stack([ux, uy, uz], 1).astype(np.float64) * vs + mb + vs / 2
        fill = ends - starts
        dense_vox += int(fill[fill >= 3 * brick ** 3 // 4].sum())
        all_vox += int(fill.sum())
        for i in range(len(trk)):
            box = trk.boxes[i]
            s, c = np.float64(np.sin(np.float32(box[6]))), np.float64(np.cos(np.float32(box[6])))
            rot_t = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
            ego = cen @ rot_t + box[:3].astype(np.float64)
            f = int(trk.frame_ids[i])
            for li in range(len(seg.inclinations)):
                ri = seg.range_images[li][f]
                H, W = ri.shape
                E = seg.extrinsics[f, li].astype(np.float32)
                v2l = np.linalg.inv(E).astype(np.float64)[:3]
                azc = float(np.arctan2(E[1, 0], E[0, 0]))
                incl_flip = seg.inclinations[li][::-1].astype(np.float64)
                idx, rng = oracle.point_cloud_to_range_image_idx(
                    ego[None], seg.extrinsics[f:f + 1, li], np.ascontiguousarray(seg.inclinations[li][::-1])[None], (H, W))
                rows, cols, rng = idx[0, :, 0], idx[0, :, 1] % W, rng[0]
                tot += len(rows)
                # tracklet-level cull (what k_pair_build already does, here with the exact footprint of all centres)
                R0, R1 = max(((rows.min() - 1) // 8) * 8, 0), min(((rows.max() + 1) // 8 + 1) * 8, H)
                C0, C1 = max(((cols.min() - 1) // 32) * 32, 0), min(((cols.max() + 1) // 32 + 1) * 32, W)
                whole = (cols.max() - cols.min() <= W // 2) and ri[R0:R1, C0:C1].max() < rng.min()
                if whole:
                    pair_culled += len(rows)
                    continue
                for a, b in zip(starts, ends):
                    r0, r1 = rows[a:b].min() - 1, rows[a:b].max() + 1
                    c0, c1 = cols[a:b].min() - 1, cols[a:b].max() + 1
                    if c1 - c0 > W // 2:                        # footprint wraps around the image seam: keep
                        continue
                    dmin = rng[a:b].min()
                    # ball variant (what k_pair_build does per tracklet, here per brick): everything from the
                    # brick's centre and the half diagonal of a full brick -- no corner projections
                    k = (a + b) // 2
                    bx, by, bz = ux[a] // brick, uy[a] // brick, uz[a] // brick
                    cidx = (np.array([bx, by, bz], np.float64) + 0.5) * brick - 0.5       # centre of the full brick
                    cb = (cidx * vs + mb + vs / 2) @ rot_t + box[:3].astype(np.float64)
                    pb = v2l[:, :3] @ cb + v2l[:, 3]
                    Rb = 0.5 * np.sqrt(3.0) * (brick - 1) * vs + 1e-3
                    d = np.linalg.norm(pb)
                    rho = np.hypot(pb[0], pb[1])
                    if d > 1.25 * Rb and rho > 1.05 * Rb:
                        dlt = np.arcsin(Rb / d) + 1e-4
                        inc_c = np.arctan2(pb[2], rho)
                        fl = incl_flip
                        rr_hi = int(np.argmin(np.abs(min(inc_c + dlt, 1.57) - fl)))     # smaller row index = higher beam
                        rr_lo = int(np.argmin(np.abs(max(inc_c - dlt, -1.57) - fl)))
                        daz = np.arcsin(Rb / rho) + 1e-4
                        az_c = np.arctan2(pb[1], pb[0]) + azc
                        cf0 = (W - 0.5) - (az_c + daz + np.pi) / (2 * np.pi) * W
                        cf1 = (W - 0.5) - (az_c - daz + np.pi) / (2 * np.pi) * W
                        q0, q1 = int(np.floor(cf0)) - 2, int(np.ceil(cf1)) + 2
                        ball_rows = (max(min(rr_hi, rr_lo) - 1, 0), min(max(rr_hi, rr_lo) + 1, H - 1))
                        for (tr, tc) in tiles:
                            r_a, r_b = (ball_rows[0] // tr) * tr, min((ball_rows[1] // tr + 1) * tr, H)
                            cols_idx = np.arange((q0 // tc) * tc, (q1 // tc + 1) * tc) % W
                            if ri[r_a:r_b][:, cols_idx].max() < d - Rb - 1e-3:
                                culled_ball[(tr, tc)] += b - a
                    # corner variant (rigorous, 8 projections): azimuth extremes of a convex body that stays clear
                    # of the sensor axis are at its corners; z is linear and |p| >= p.u >= min corner.u (u = unit vector
                    # to the brick centre), |p| <= max corner norm, so sin(inc) = z/|p| is bracketed by the corner extremes
                    if d > 1.25 * Rb and rho > 1.05 * Rb:
                        lo3 = np.array([bx, by, bz], np.float64) * brick
                        cor = np.array([[x, y, z] for x in (0, brick - 1) for y in (0, brick - 1) for z in (0, brick - 1)],
                                       np.float64) + lo3
                        pc = ((cor * vs + mb + vs / 2) @ rot_t + box[:3].astype(np.float64)) @ v2l[:, :3].T + v2l[:, 3]
                        u_c = pb / d
                        rmin_c, rmax_c = float((pc @ u_c).min()), float(np.linalg.norm(pc, axis=1).max())
                        zmin, zmax = float(pc[:, 2].min()), float(pc[:, 2].max())
                        s_hi = zmax / (rmin_c if zmax > 0 else rmax_c)
                        s_lo = zmin / (rmax_c if zmin > 0 else rmin_c)
                        inc_hi = np.arcsin(np.clip(s_hi, -1, 1)) + 1e-4
                        inc_lo = np.arcsin(np.clip(s_lo, -1, 1)) - 1e-4
                        rr_hi = int(np.argmin(np.abs(inc_hi - incl_flip)))
                        rr_lo = int(np.argmin(np.abs(inc_lo - incl_flip)))
                        azs = np.arctan2(pc[:, 1], pc[:, 0])
                        azs = az_c - azc + ((azs - (az_c - azc) + np.pi) % (2 * np.pi) - np.pi)    # same branch as the centre
                        cfa = (W - 0.5) - (azs.max() + azc + 1e-4 + np.pi) / (2 * np.pi) * W
                        cfb = (W - 0.5) - (azs.min() + azc - 1e-4 + np.pi) / (2 * np.pi) * W
                        q0, q1 = int(np.floor(cfa)) - 1, int(np.ceil(cfb)) + 1
                        cr_rows = (max(min(rr_hi, rr_lo) - 1, 0), min(max(rr_hi, rr_lo) + 1, H - 1))
                        # sanity: the rigorous footprint must contain the exact one
                        assert cr_rows[0] <= rows[a:b].min() and cr_rows[1] >= rows[a:b].max(), "row footprint not conservative"
                        assert rmin_c - 1e-3 <= dmin, "range bound not conservative"
                        for (tr, tc) in tiles:
                            r_a, r_b = (cr_rows[0] // tr) * tr, min((cr_rows[1] // tr + 1) * tr, H)
                            cols_idx = np.arange((q0 // tc) * tc, (q1 // tc + 1) * tc) % W
                            assert np.isin(cols[a:b], cols_idx).all(), "column footprint not conservative"
                            if ri[r_a:r_b][:, cols_idx].max() < rmin_c - 1e-3:
                                culled_corner[(tr, tc)] += b - a
                    for (tr, tc) in tiles:
                        rr0, rr1 = max((r0 // tr) * tr, 0), min((r1 // tr + 1) * tr, H)
                        cc0, cc1 = max((c0 // tc) * tc, 0), min((c1 // tc + 1) * tc, W)
                        if ri[rr0:rr1, cc0:cc1].max() < dmin:
                            culled[(tr, tc)] += b - a
    left = tot - pair_culled
    print(f"unknown-voxel tests: {tot}; removed by the tracklet-level pair cull: {100.0 * pair_culled / tot:.1f} %; left: {left}")
    print(f"unknown voxels sitting in bricks that are at least 3/4 unknown: {100.0 * dense_vox / max(all_vox, 1):.1f} %")
    for t in tiles:
        print(f"  brick {brick}^3, pyramid tile {t[0]}x{t[1]}: {100.0 * culled[t] / left:5.1f} % of the remaining tests culled "
              f"(exact footprint) / {100.0 * culled_corner[t] / left:5.1f} % (8 corners, rigorous) / "
              f"{100.0 * culled_ball[t] / left:5.1f} % (ball around the brick centre)")


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
