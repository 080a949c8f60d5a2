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
                    for (tr, tc) in tiles:
                        rr0, rr1 = max((r0 // tr) * tr, 0), min((r1 // tr + 1) * tr, H)
                        cc0, cc1 = max((c0 // tc) * tc, 0), min((c1 // tc + 1) * tc, W)
                        if ri[rr0:rr1, cc0:cc1].max() < dmin:
                            culled[(tr, tc)] += b - a
    left = tot - pair_culled
    print(f"unknown-voxel tests: {tot}; removed by the tracklet-level pair cull: {100.0 * pair_culled / tot:.1f} %; left: {left}")
    print(f"unknown voxels sitting in bricks that are at least 3/4 unknown: {100.0 * dense_vox / max(all_vox, 1):.1f} %")
    for t in tiles:
        print(f"  brick {brick}^3, pyramid tile {t[0]}x{t[1]}: {100.0 * culled[t] / left:5.1f} % of the remaining tests culled")


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
