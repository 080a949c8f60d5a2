"""Pin oracle/occ_oracle.c against the real reference (runs only where /root/reference exists).

    python -m oracle.validate_oracle [--quick]

Checks, all bit-exact unless noted:
  1. annotate path: C oracle vs oracle/torch_ref.py (reference's own
     point_cloud_to_range_image_idx + reference-compiled points_in_boxes_cpu +
     the torch-CPU op sequence of annotate_trk) -- labels, dims, local points.
  2. point_cloud_to_range_image_idx alone on random f64 points: indices and ranges.
  3. points_in_boxes: C oracle vs reference-compiled points_in_boxes_cpu + golden vectors
     of tests/test_models/test_common_modules/test_roiaware_pool3d.py:43-120.
  4. dynamic / hard voxelize: C oracle vs reference-compiled voxel_layer CPU; KAT of
     tests/test_models/test_voxel_encoder/test_voxel_generator.py:6-22.
  5. scatter: C oracle vs torch.unique + index_add_/index_reduce_ (the reference has no CPU
     kernel, voxelization.h:106; tolerance 1e-5 rel on features, exact on coords/maps).
  6. --save-mean-var grids, 7. MirrorOccLabel: vs the reference's torch op sequences.
  8. occ_ops: oracle quantize_points / generate_dense_voxel_centers vs the reference's own
     mmdet3d/ops/occ/occ_ops.py loaded by file path (it needs nothing but torch).
"""
from __future__ import annotations

import sys
import time

import numpy as np
import torch

from objectcentricocccompletion_b200 import synth
from . import build as _build
from . import oracle, torch_ref


def check_annotate(quick):
    cases = [dict(n=3, b=12, vs=0.2, kind="vehicle", seed=1, small=True),
             dict(n=4, b=14, vs=0.2, kind="vehicle", seed=2, small=False),
             dict(n=2, b=10, vs=0.1, kind="large", seed=3, small=False),
             dict(n=3, b=20, vs=0.25, kind="vehicle", seed=4, small=True)]
    if not quick:
        cases += [dict(n=6, b=20, vs=0.2, kind="vehicle", seed=10 + i, small=False) for i in range(4)]
        cases += [dict(n=2, b=12, vs=0.1, kind="large", seed=20 + i, small=False) for i in range(2)]
    nvox = nbad = npts = nptbad = 0
    for cs in cases:
        b = synth.make_batch(cs["n"], cs["b"], cs["vs"], cs["kind"], cs["seed"], small=cs["small"])
        ref = torch_ref.annotate_batch(b)
        orc = oracle.annotate_batch(b)
        for t, (r, o) in enumerate(zip(ref, orc)):
            assert r["status"] == o["status"], (cs, t, r["status"], o["status"])
            if r["occ"] is None:
                continue
            assert r["occ"].shape == o["occ"].shape and (r["dims"] == o["dims"]).all()
            assert (r["size"] == o["size"]).all()
            nvox += r["occ"].size
            nbad += int((r["occ"] != o["occ"]).sum())
            dbg = oracle.annotate_tracklet_debug(b, t)
            loc = dbg["loc"][dbg["keep"]]
            # torch_ref's loc is after the q<dims filter; compare the common prefix semantics via set equality
            vsf = np.float32(cs["vs"])
            mb = np.array([-0.5 * r["size"][0], -0.5 * r["size"][1], 0], np.float32)
            q = np.floor((loc - mb) / vsf).astype(np.int64)
            kept = loc[(q < r["dims"][None]).all(1)]
            npts += len(kept)
            assert kept.shape == r["loc"].shape, (kept.shape, r["loc"].shape)
            nptbad += int((kept.view(np.uint32) != r["loc"].view(np.uint32)).any(1).sum())
    print(f"[1] annotate: {nvox} voxels, {nbad} label mismatches; {npts} local points, {nptbad} bit mismatches")
    return nbad == 0 and nptbad == 0


def check_projection(quick):
    fn = torch_ref.reference_projection_fn()
    rng = np.random.default_rng(0)
    bad = tot = 0
    for (H, W, n) in [(64, 2650, 20000), (200, 600, 20000), (16, 331, 5000)] + ([] if quick else [(64, 2650, 200000)]):
        rig = synth.lidar_rig(rng)
        B = 5
        E = np.stack([r["extrinsic"] for r in rig], 0)
        lo, hi = (-17.6, 2.4) if H != 200 else (-90, 30)
        incl = np.sort(np.deg2rad(rng.uniform(lo, hi, (B, H))).astype(np.float32), 1)[:, ::-1].copy()
        pts = rng.uniform(-60, 60, (B, n, 3))
        pts[..., 2] = rng.uniform(-3, 6, (B, n))
        idx_r, rng_r = fn(torch.from_numpy(pts), torch.from_numpy(E), torch.from_numpy(incl), (H, W))
        idx_o, rng_o = oracle.point_cloud_to_range_image_idx(pts, E, incl, (H, W))
        bad += int((idx_r.numpy() != idx_o).any(-1).sum()) + int((rng_r.numpy().view(np.uint64) != rng_o.view(np.uint64)).sum())
        tot += B * n
    print(f"[2] point_cloud_to_range_image_idx: {tot} points, {bad} mismatches (index or range bits)")
    return bad == 0


def check_points_in_boxes(quick):
    ext = _build.load_ref("ref_points_in_boxes")
    rng = np.random.default_rng(1)
    bad = tot = 0
    for _ in range(3 if quick else 10):
        T, M = 6, 50000
        boxes = np.concatenate([rng.uniform(-20, 20, (T, 3)), rng.uniform(1, 8, (T, 3)), rng.uniform(-4, 4, (T, 1))], 1).astype(np.float32)
        pts = (boxes[rng.integers(0, T, M), :3] + rng.normal(0, 3, (M, 3))).astype(np.float32)
        out = torch.zeros((T, M), dtype=torch.int32)
        ext.points_in_boxes_cpu(torch.from_numpy(boxes), torch.from_numpy(pts), out)
        mine = oracle.points_in_boxes_cpu(pts, boxes)
        bad += int((out.numpy() != mine).sum())
        tot += T * M
    # golden vectors of the reference's tests (test_roiaware_pool3d.py:43-120)
    boxes = np.array([[1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 0.3], [-10.0, 23.0, 16.0, 10, 20, 20, 0.5]], np.float32)
    pts = np.array([[1, 2, 3.3], [1.2, 2.5, 3.0], [0.8, 2.1, 3.5], [1.6, 2.6, 3.6], [0.8, 1.2, 3.9], [-9.2, 21.0, 18.2],
                    [3.8, 7.9, 6.3], [4.7, 3.5, -12.2], [3.8, 7.6, -2], [-10.6, -12.9, -20], [-16, -18, 9],
                    [-21.3, -52, -5], [0, 0, 0], [6, 7, 8], [-2, -3, -4]], np.float32)
    exp = np.array([[1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0]], np.int32)
    ok = (oracle.points_in_boxes_cpu(pts, boxes) == exp).all()
    print(f"[3] points_in_boxes: {tot} tests vs reference-compiled, {bad} mismatches; golden vector ok={bool(ok)}")
    return bad == 0 and ok


def check_voxelize(quick):
    ext = _build.load_ref("ref_voxel_layer")
    rng = np.random.default_rng(2)
    bad = 0
    for C, n in [(4, 100000), (5, 300000)]:
        pts = np.concatenate([rng.uniform(-80, 80, (n, 2)), rng.uniform(-6, 10, (n, 1)), rng.random((n, C - 3))], 1).astype(np.float32)
        vs, cr = [0.2, 0.2, 0.2], [-74.88, -74.88, -4, 74.88, 74.88, 8]
        coors = torch.zeros((n, 3), dtype=torch.int32)
        ext.dynamic_voxelize(torch.from_numpy(pts), coors, vs, cr, 3)
        bad += int((coors.numpy() != oracle.dynamic_voxelize(pts, vs, cr)).sum())
        vs, cr = [0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1]
        mp, mv = 8, 5000
        voxels = torch.zeros((mv, mp, C)); co = torch.zeros((mv, 3), dtype=torch.int32); num = torch.zeros((mv,), dtype=torch.int32)
        m = ext.hard_voxelize(torch.from_numpy(pts), voxels, co, num, vs, cr, mp, mv, 3)
        v2, c2, n2 = oracle.hard_voxelize(pts, vs, cr, mp, mv)
        bad += int(m != len(c2)) + int((voxels[:m].numpy() != v2).sum()) + int((co[:m].numpy() != c2).sum()) + int((num[:m].numpy() != n2).sum())
    # KAT test_voxel_generator.py:6-22
    np.random.seed(0)
    points = np.random.rand(1000, 4).astype(np.float32)
    v, c, nn = oracle.hard_voxelize(points, [0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1], 1000, 20000)
    exp_c = np.array([[7, 81, 1], [6, 81, 0], [7, 80, 1], [6, 81, 1], [7, 81, 0], [6, 80, 1], [7, 80, 0], [6, 80, 0]])
    exp_n = np.array([120, 121, 127, 134, 115, 127, 125, 131])
    kat = (c == exp_c).all() and (nn == exp_n).all()
    print(f"[4] voxelize: {bad} mismatches vs reference-compiled voxel_layer; VoxelGenerator KAT ok={bool(kat)}")
    return bad == 0 and kat


def check_scatter(quick):
    rng = np.random.default_rng(3)
    ok = True
    n, c = 200000, 3
    feats = (rng.random((n, c)) * 100 - 50).astype(np.float32)
    coors = rng.integers(-1, 20, (n, 3)).astype(np.int32)
    for mode in ("mean", "max", "sum"):
        vf, vc, mp, cnt = oracle.dynamic_scatter_fwd(feats, coors, mode)
        tc = torch.from_numpy(coors)
        clean = tc.masked_fill(tc.lt(0).any(-1, True), -1)
        u, inv, cn = torch.unique(clean, dim=0, return_inverse=True, return_counts=True)
        u, cn, inv = u[1:], cn[1:], inv - 1
        ok &= bool((u.numpy() == vc).all() and (cn.numpy() == cnt).all() and (inv.numpy() == mp).all())
        valid = inv >= 0
        tf = torch.from_numpy(feats).double()
        if mode == "max":
            ref = torch.full((len(u), c), -np.inf, dtype=torch.float64).index_reduce_(0, inv[valid], tf[valid], "amax")
        else:
            ref = torch.zeros((len(u), c), dtype=torch.float64).index_add_(0, inv[valid], tf[valid])
            if mode == "mean":
                ref = ref / cn[:, None]
        ok &= bool(np.allclose(vf, ref.numpy(), rtol=1e-5, atol=1e-3))
    c64 = rng.integers(0, 12, (n, 4)).astype(np.int64)
    for mode in ("mean", "max", "sum"):
        nf, nc, inv = oracle.scatter_v2(feats, c64, mode)
        u, ti = torch.unique(torch.from_numpy(c64), dim=0, return_inverse=True)
        ok &= bool((u.numpy() == nc).all() and (ti.numpy() == inv).all())
    print(f"[5] scatter: DynamicScatter fwd (drop-first) + scatter_v2 vs torch.unique/index ops ok={ok}")
    return ok


def check_mean_var(quick):
    """--save-mean-var grids (occ_annotate.py:627-645): numpy restatement vs the torch op sequence on the reference
    glue's own kept points and quantised coordinates."""
    from objectcentricocccompletion_b200 import synth
    from .make_golden import edge_batch

    ncell = bad = 0
    batches = [synth.make_batch(3, 12, 0.2, seed=1, small=True), edge_batch()]
    if not quick:
        batches.append(synth.make_batch(2, 14, 0.2, seed=6))
    for b in batches:
        mine = oracle.annotate_mean_var(b)
        for m, r in zip(mine, torch_ref.annotate_batch(b)):
            if r["occ"] is None:
                bad += m is not None
                continue
            e = torch_ref.mean_var(r)
            ncell += int((e != 0).any(-1).sum())
            bad += int((m.view(np.uint32) != e.view(np.uint32)).any(-1).sum())
    print(f"[6] mean/var grids: {ncell} filled cells, {bad} cells with a bit mismatch")
    return bad == 0


def check_mirror(quick):
    """MirrorOccLabel (occ_pinelines.py:88-126): numpy restatement vs the reference's torch op sequence."""
    rng = np.random.default_rng(5)
    ok = True
    for shape in [(11, 24, 9), (12, 23, 10), (1, 3, 2), (29, 120, 35)]:
        g = torch.from_numpy(rng.integers(0, 3, shape).astype(np.int32))
        XS, YS, ZS = g.shape
        flat = g.clone().view(-1)
        unknown = flat == 0
        mid = XS // 2
        vx, vy, vz = torch.meshgrid(torch.arange(XS), torch.arange(YS), torch.arange(ZS), indexing="ij")
        coors = torch.stack([((vx + 0.5 - mid) * -1.0 + mid).long(), vy, vz], -1).view(-1, 3)
        flat[unknown] = g[coors[unknown][:, 0], coors[unknown][:, 1], coors[unknown][:, 2]]
        ok &= bool((oracle.mirror_occ_label(g.numpy()) == flat.view(XS, YS, ZS).numpy()).all())
    print(f"[7] MirrorOccLabel vs the reference's torch ops ok={ok}")
    return ok


def check_occ_ops(quick):
    """A10 against the reference module itself (mmdet3d/ops/occ/occ_ops.py:5-93, imports only torch)."""
    import importlib.util
    import os

    path = os.path.join(_build.REFERENCE, "mmdet3d", "ops", "occ", "occ_ops.py")
    spec = importlib.util.spec_from_file_location("ref_occ_ops", path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(3)
    n_el = bad = 0
    for rep in range(3 if quick else 12):
        R, N = 40, 200_000 if quick else 1_000_000
        rois = np.concatenate([rng.integers(0, 4, (R, 1)), rng.uniform(-40, 40, (R, 3)), rng.uniform(1.0, 12.0, (R, 3)),
                               rng.uniform(-3, 3, (R, 1 + 2 * (rep % 2)))], 1).astype(np.float32)
        idx = rng.integers(0, R, N)
        pts = ((rng.random((N, 3)) - 0.5) * rois[idx, 4:7] * 1.3).astype(np.float32)
        vs = [0.2, 0.1, 0.25, 0.3][rep % 4]
        scale, offset = ([1.0, 1.0, 1.0], [0.0, 0.0, 0.0]) if rep % 3 == 0 else ([1.1, 1.05, 1.2], [0.4, 0.2, 0.1])
        for to_center in (False, True):
            want = ref.quantize_points(torch.from_numpy(pts), torch.from_numpy(rois), torch.from_numpy(idx), vs, scale,
                                       offset, to_center).numpy()
            got = oracle.quantize_points(pts, rois, idx, vs, scale, offset, to_center)
            n_el += want.size
            bad += int((np.ascontiguousarray(got).view(np.uint8) != np.ascontiguousarray(want).view(np.uint8)).any(-1).sum()) \
                if got.shape == want.shape and got.dtype == want.dtype else want.size
        sizes = rois[:8, 4:7]
        want_c = ref.generate_dense_voxel_centers(torch.from_numpy(sizes), vs, scale, offset)
        got_c = oracle.generate_dense_voxel_centers(sizes, vs, scale, offset)
        for w_, g_ in zip(want_c, got_c):
            w_ = w_.numpy()
            n_el += w_.size
            bad += int((g_.view(np.uint32) != w_.view(np.uint32)).sum()) if g_.shape == w_.shape else w_.size
    print(f"[8] occ_ops vs the reference's own occ_ops.py: {n_el} elements, {bad} bit mismatches")
    return bad == 0


def main():
    quick = "--quick" in sys.argv
    assert torch_ref.available(), "needs /root/reference"
    _build.build_oracle()
    _build.build_ref()
    t0 = time.time()
    res = [check_annotate(quick), check_projection(quick), check_points_in_boxes(quick), check_voxelize(quick), check_scatter(quick),
           check_mean_var(quick), check_mirror(quick), check_occ_ops(quick)]
    print(f"oracle pinned: {all(res)}  ({time.time() - t0:.1f}s)")
    return 0 if all(res) else 1


if __name__ == "__main__":
    sys.exit(main())
