"""Python face of the CPU parity oracle (TEST INFRASTRUCTURE, never the product).

Loads ``oracle/_build/liboccoracle.so`` (the C restatement in occ_oracle.c) and
adds the few *host* operations the reference itself performs with torch-CPU
before/around its tensor code -- ``torch.sin/cos`` of the box yaw
(lidar_box3d.py:163-164, occ_annotate.py:490-491), ``torch.linalg.inv`` of the
f32 extrinsic (occ_annotate.py:158-160), ``torch.atan2`` of its first column
(occ_annotate.py:175) and the inclination flip (occ_annotate.py:528).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import numpy as np

from . import build as _build

STATUS = {0: "ok", 1: "skip_short", 2: "no_points", 3: "empty_after_filter", 4: "index_error"}
REDUCE = {"sum": 0, "mean": 1, "max": 2}

_lib = None


class SensorT(C.Structure):
    _fields_ = [("ri_off", C.c_int64), ("incl_off", C.c_int64), ("H", C.c_int32), ("W", C.c_int32),
                ("v2l", C.c_float * 12), ("azc", C.c_float), ("pad", C.c_int32)]


class TrkT(C.Structure):
    _fields_ = [("B", C.c_int32), ("pad", C.c_int32), ("frame0", C.c_int64), ("label_off", C.c_int64),
                ("label_cap", C.c_int64)]


SENSOR_DTYPE = np.dtype([("ri_off", "<i8"), ("incl_off", "<i8"), ("H", "<i4"), ("W", "<i4"),
                         ("v2l", "<f4", (12,)), ("azc", "<f4"), ("pad", "<i4")])
TRK_DTYPE = np.dtype([("B", "<i4"), ("pad", "<i4"), ("frame0", "<i8"), ("label_off", "<i8"), ("label_cap", "<i8")])
assert SENSOR_DTYPE.itemsize == C.sizeof(SensorT) == 80
assert TRK_DTYPE.itemsize == C.sizeof(TrkT) == 32


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build_oracle())
    return _lib


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


# ---------------------------------------------------------------- A1
def points_in_boxes_cpu(points, boxes):
    """points [M,3], boxes [T,7] -> int32 [T,M] (points_in_boxes.py:52-83)."""
    pts, bx = _c(points, np.float32), _c(boxes, np.float32)
    out = np.zeros((bx.shape[0], pts.shape[0]), np.int32)
    lib().orc_points_in_boxes_cpu(C.c_int(bx.shape[0]), C.c_int(pts.shape[0]), _p(bx), _p(pts), _p(out))
    return out


def points_in_boxes_gpu(points, boxes):
    """points [B,M,3], boxes [B,T,7] -> int32 [B,M], -1 background (points_in_boxes.py:6-50)."""
    pts, bx = _c(points, np.float32), _c(boxes, np.float32)
    B, M, _ = pts.shape
    out = np.full((B, M), -1, np.int32)
    lib().orc_points_in_boxes_gpu(C.c_int(B), C.c_int(bx.shape[1]), C.c_int(M), _p(bx), _p(pts), _p(out))
    return out


def points_in_boxes_batch(points, boxes):
    """points [B,M,3], boxes [B,T,7] -> int32 [B,M,T] (points_in_boxes.py:86-123)."""
    pts, bx = _c(points, np.float32), _c(boxes, np.float32)
    B, M, _ = pts.shape
    T = bx.shape[1]
    out = np.zeros((B, M, T), np.int32)
    lib().orc_points_in_boxes_batch(C.c_int(B), C.c_int(T), C.c_int(M), _p(bx), _p(pts), _p(out))
    return out


# ---------------------------------------------------------------- A7 / A8
def dynamic_voxelize(points, voxel_size, coors_range):
    pts = _c(points, np.float32)
    coors = np.zeros((pts.shape[0], 3), np.int32)
    vs, cr = _c(voxel_size, np.float32), _c(coors_range, np.float32)
    lib().orc_dynamic_voxelize(_p(pts), C.c_long(pts.shape[0]), C.c_int(pts.shape[1]), _p(vs), _p(cr), _p(coors))
    return coors


def hard_voxelize(points, voxel_size, coors_range, max_points, max_voxels):
    pts = _c(points, np.float32)
    n, c = pts.shape
    voxels = np.zeros((max_voxels, max_points, c), np.float32)
    coors = np.zeros((max_voxels, 3), np.int32)
    num = np.zeros((max_voxels,), np.int32)
    vs, cr = _c(voxel_size, np.float32), _c(coors_range, np.float32)
    f = lib().orc_hard_voxelize
    f.restype = C.c_int
    m = f(_p(pts), C.c_long(n), C.c_int(c), _p(vs), _p(cr), C.c_int(max_points), C.c_int(max_voxels),
          _p(voxels), _p(coors), _p(num))
    return voxels[:m], coors[:m], num[:m]


# ---------------------------------------------------------------- A6 / A9
def unique_rows(rows):
    r = _c(rows, np.int64)
    n, k = r.shape
    uniq = np.zeros((max(n, 1), k), np.int64)
    inv = np.zeros((n,), np.int64)
    cnt = np.zeros((max(n, 1),), np.int64)
    f = lib().orc_unique_rows
    f.restype = C.c_long
    m = f(_p(r), C.c_long(n), C.c_int(k), _p(uniq), _p(inv), _p(cnt))
    return uniq[:m], inv, cnt[:m]


def scatter_v2(feat, coors, mode, return_inv=True, min_points=0, unq_inv=None, new_coors=None):
    """sst_ops.py:150-181 (torch_scatter restated: mean = sum / clamp(count,1); max -> values)."""
    feat = _c(feat, np.float32)
    coors = _c(coors, np.int64)
    assert feat.shape[0] == coors.shape[0]
    if mode == "avg":
        mode = "mean"
    if mode not in REDUCE:
        raise NotImplementedError(mode)
    if unq_inv is None:
        new_coors, unq_inv, unq_cnt = unique_rows(coors)
    else:
        assert new_coors is not None
        unq_inv = _c(unq_inv, np.int64)
        unq_cnt = np.bincount(unq_inv, minlength=len(new_coors)).astype(np.int64)
    if min_points > 0:
        valid = unq_cnt[unq_inv] >= min_points
        feat, coors = feat[valid], coors[valid]
        new_coors, unq_inv, unq_cnt = unique_rows(coors)
    m, c = len(new_coors), feat.shape[1]
    out = np.zeros((m, c), np.float32)
    cnt = np.ascontiguousarray(unq_cnt, np.int64)
    feat = np.ascontiguousarray(feat)
    lib().orc_segment_reduce(_p(feat), C.c_long(feat.shape[0]), C.c_int(c), _p(unq_inv), C.c_long(m),
                             C.c_int(REDUCE[mode]), _p(out), _p(cnt))
    if not return_inv:
        return out, new_coors
    return out, new_coors, unq_inv


def dynamic_scatter_fwd(feats, coors, reduce_type):
    """scatter_points_cuda.cu:183-234 -> (voxel_feats, voxel_coors, point2voxel_map, count)."""
    feats, coors = _c(feats, np.float32), _c(coors, np.int32)
    n, c = feats.shape
    if n == 0:
        return feats.copy(), coors.copy(), np.zeros((0,), np.int32), np.zeros((0,), np.int32)
    vf = np.zeros((n, c), np.float32)
    vc = np.zeros((n, 3), np.int32)
    mp = np.zeros((n,), np.int32)
    cnt = np.zeros((n,), np.int32)
    f = lib().orc_dynamic_scatter_fwd
    f.restype = C.c_long
    m = f(_p(feats), _p(coors), C.c_long(n), C.c_int(c), C.c_int(REDUCE[reduce_type]), _p(vf), _p(vc), _p(mp), _p(cnt))
    return vf[:m].copy(), vc[:m].copy(), mp, cnt[:m].copy()


def dynamic_scatter_bwd(grad_voxel, feats, voxel_feats, p2v, cnt, reduce_type):
    feats = _c(feats, np.float32)
    n, c = feats.shape
    g = np.zeros((n, c), np.float32)
    gv, vf = _c(grad_voxel, np.float32), _c(voxel_feats, np.float32)
    p2v, cnt = _c(p2v, np.int32), _c(cnt, np.int32)
    lib().orc_dynamic_scatter_bwd(_p(g), _p(gv), _p(feats), _p(vf), _p(p2v), _p(cnt), C.c_long(n),
                                  C.c_long(vf.shape[0]), C.c_int(c), C.c_int(REDUCE[reduce_type]))
    return g


def dynamic_scatter_batched(feats, coors, reduce_type):
    """DynamicScatter.forward with 4-col coords (scatter_points.py:83-99)."""
    feats, coors = _c(feats, np.float32), _c(coors, np.int32)
    if coors.shape[1] == 3:
        vf, vc, _, _ = dynamic_scatter_fwd(feats, coors, reduce_type)
        return vf, vc
    bs = int(coors[-1, 0]) + 1
    vfs, vcs = [], []
    for i in range(bs):
        sel = coors[:, 0] == i
        vf, vc, _, _ = dynamic_scatter_fwd(feats[sel], coors[sel][:, 1:], reduce_type)
        vcs.append(np.concatenate([np.full((len(vc), 1), i, np.int32), vc], 1))
        vfs.append(vf)
    return np.concatenate(vfs, 0), np.concatenate(vcs, 0)


# ---------------------------------------------------------------- A10
def quantize_points(points, rois, rois_points_idx, voxel_size, scale_wlh=(1.0, 1.0, 1.0),
                    offset_wlh=(0.0, 0.0, 0.0), to_center=False):
    pts, rois = _c(points, np.float32), _c(rois, np.float32)
    idx = _c(rois_points_idx, np.int64)
    n = pts.shape[0]
    coor = np.zeros((n, 3), np.int64)
    cen = np.zeros((n, 3), np.float32)
    sc, of = _c(scale_wlh, np.float32), _c(offset_wlh, np.float32)
    lib().orc_quantize_points(_p(pts), C.c_long(n), _p(rois), C.c_int(rois.shape[1]), _p(idx),
                              C.c_float(voxel_size), _p(sc), _p(of), C.c_int(int(to_center)), _p(coor), _p(cen))
    return cen if to_center else coor


def generate_dense_voxel_centers(bbox_sizes, voxel_size, scale_wlh=(1.0, 1.0, 1.0),
                                 offset_wlh=(0.0, 0.0, 0.0), as_volume=False):
    sizes = _c(bbox_sizes, np.float32)
    sc, of = _c(scale_wlh, np.float32), _c(offset_wlh, np.float32)
    f = lib().orc_dense_voxel_centers
    f.restype = C.c_long
    out = []
    for s in sizes:
        s = np.ascontiguousarray(s)
        dims = np.zeros(3, np.int32)
        n = f(_p(s), C.c_float(voxel_size), _p(sc), _p(of), _p(dims), None)
        cen = np.zeros((n, 3), np.float32)
        f(_p(s), C.c_float(voxel_size), _p(sc), _p(of), _p(dims), _p(cen))
        out.append(cen.reshape(dims[0], dims[1], dims[2], 3) if as_volume else cen)
    return out


# ---------------------------------------------------------------- A5
def host_calib(extrinsics):
    """inv(extrinsic) in f32 and atan2(E[1,0],E[0,0]) in f32 with torch-CPU (occ_annotate.py:158-160,175).

    extrinsics f32 [...,4,4] -> v2l f32 [...,12] (rows 0..2 of the inverse), azc f32 [...].
    """
    import torch

    E = torch.from_numpy(np.ascontiguousarray(extrinsics, np.float32))
    shape = E.shape[:-2]
    E = E.reshape(-1, 4, 4)
    inv = torch.linalg.inv(E)
    azc = torch.atan2(E[:, 1, 0], E[:, 0, 0])
    return (inv[:, :3, :].reshape(*shape, 12).numpy().copy(), azc.reshape(shape).numpy().copy())


def host_trig(rz):
    """cos(-rz), sin(-rz), cos(rz), sin(rz) in torch-CPU f32 (lidar_box3d.py:163-164; occ_annotate.py:490-491)."""
    import torch

    r = torch.from_numpy(np.ascontiguousarray(rz, np.float32))
    m = -r
    return torch.stack([torch.cos(m), torch.sin(m), torch.cos(r), torch.sin(r)], -1).numpy().copy()


def point_cloud_to_range_image_idx(points, extrinsics, inclinations, range_image_size):
    """Reference signature (occ_annotate.py:141-201): points f64 [B,N,3], extrinsics f32 [B,4,4],
    inclinations f32 [B,H] (already flipped) -> (ri_indices int64 [B,N,2], ri_range f64 [B,N])."""
    pts = _c(points, np.float64)
    B, N, _ = pts.shape
    H, W = range_image_size
    v2l, azc = host_calib(extrinsics)
    incl = _c(inclinations, np.float32)
    assert incl.shape == (B, H)
    idx = np.zeros((B, N, 2), np.int64)
    rng = np.zeros((B, N), np.float64)
    lib().orc_point_cloud_to_range_image_idx(_p(pts), C.c_int(B), C.c_long(N), _p(_c(v2l, np.float32)),
                                             _p(_c(azc, np.float32)), _p(incl), C.c_int(H), C.c_int(W), _p(idx), _p(rng))
    return idx, rng


# ---------------------------------------------------------------- A2-A5 annotate
class PackedBatch:
    """Flat arrays the C oracle consumes, built from a synth.TrackletBatch-shaped object."""

    def __init__(self, batch, pack_override: Optional[dict] = None):
        trks = batch.tracklets
        segs = batch.segments
        L = len(segs[0].inclinations) if segs else 5
        self.L = L
        self.T = len(trks)
        self.voxel_size = float(batch.voxel_size)
        # sensors per (segment frame, lidar)
        sf_base = np.cumsum([0] + [s.num_frames for s in segs])
        SF = int(sf_base[-1])
        sensors = np.zeros((SF, L), SENSOR_DTYPE)
        incl_pool, ri_pool = [], []
        incl_off = ri_off = 0
        for si, s in enumerate(segs):
            v2l, azc = host_calib(s.extrinsics)            # [B,L,12], [B,L]
            for c in range(L):
                fl = np.ascontiguousarray(s.inclinations[c][::-1], np.float32)   # flip (:528)
                incl_pool.append(fl)
                img = np.ascontiguousarray(s.range_images[c], np.float32)
                Bf, H, W = img.shape
                ri_pool.append(img.reshape(-1))
                sl = slice(sf_base[si], sf_base[si + 1])
                sensors["incl_off"][sl, c] = incl_off
                sensors["ri_off"][sl, c] = ri_off + np.arange(Bf, dtype=np.int64) * H * W
                sensors["H"][sl, c] = H
                sensors["W"][sl, c] = W
                sensors["v2l"][sl, c] = v2l[:, c]
                sensors["azc"][sl, c] = azc[:, c]
                incl_off += H
                ri_off += Bf * H * W
        self.sensors = sensors
        self.incl_pool = np.concatenate(incl_pool) if incl_pool else np.zeros(0, np.float32)
        self.ri_pool = np.concatenate(ri_pool) if ri_pool else np.zeros(0, np.float32)
        nfr = [len(t) for t in trks]
        self.trk_frame_off = np.cumsum([0] + nfr).astype(np.int64)
        self.boxes = (np.concatenate([t.boxes for t in trks], 0).astype(np.float32) if trks
                      else np.zeros((0, 7), np.float32))
        self.frame_sf = (np.concatenate([sf_base[t.segment] + t.frame_ids for t in trks]).astype(np.int32)
                         if trks else np.zeros(0, np.int32))
        # only xyz enters the path (occ_annotate.py:97 slices [:, :3]); KITTI rows carry 3 more columns
        allp = [np.asarray(p)[:, :3] for t in trks for p in t.points]
        self.pt_off = np.cumsum([0] + [len(p) for p in allp]).astype(np.int64)
        self.points = (np.concatenate(allp, 0).astype(np.float32) if allp and self.pt_off[-1] > 0
                       else np.zeros((0, 3), np.float32))
        self.trig = host_trig(self.boxes[:, 6]) if len(self.boxes) else np.zeros((0, 4), np.float32)
        if pack_override:          # golden fixtures carry the host-derived values they were made with
            if "trig" in pack_override:
                self.trig = np.ascontiguousarray(pack_override["trig"], np.float32)
            if "v2l" in pack_override:
                self.sensors["v2l"] = pack_override["v2l"].reshape(self.sensors["v2l"].shape)
            if "azc" in pack_override:
                self.sensors["azc"] = pack_override["azc"].reshape(self.sensors["azc"].shape)
        # label slots sized by the per-tracklet upper bound max-over-all-frames box size
        vsf = np.float32(self.voxel_size)
        caps = []
        for t in trks:
            if len(t) == 0:
                caps.append(0)
                continue
            d = np.ceil(t.boxes[:, 3:6].max(0).astype(np.float32) / vsf).astype(np.int64)
            caps.append(int(d[0] * d[1] * d[2]))
        self.label_off = np.cumsum([0] + caps).astype(np.int64)


def annotate_batch(batch, threads: int = 1, packed: Optional[PackedBatch] = None):
    """occ_annotate.py annotate_trk for every tracklet -> list of dict(status, occ[X,Y,Z] int32 | None, ...)."""
    pk = packed or PackedBatch(batch)
    T = pk.T
    trk = np.zeros(T, TRK_DTYPE)
    trk["B"] = np.diff(pk.trk_frame_off)
    trk["frame0"] = pk.trk_frame_off[:-1]
    trk["label_off"] = pk.label_off[:-1]
    trk["label_cap"] = np.diff(pk.label_off)
    labels = np.zeros(int(pk.label_off[-1]), np.int32)
    dims = np.zeros((T, 3), np.int32)
    size = np.zeros((T, 3), np.float32)
    status = np.zeros(T, np.int32)
    nunk = np.zeros(T, np.int64)
    lib().orc_annotate_batch(C.c_int(T), _p(trk), _p(pk.boxes), _p(pk.trig), _p(pk.points), _p(pk.pt_off),
                             _p(pk.frame_sf), _p(pk.sensors), C.c_int(pk.L), _p(pk.incl_pool), _p(pk.ri_pool),
                             C.c_double(pk.voxel_size), _p(dims), _p(size), _p(labels), _p(status), _p(nunk),
                             C.c_int(threads))
    out = []
    for t in range(T):
        st = int(status[t])
        if st != 0:
            out.append(dict(status=STATUS.get(st, str(st)), occ=None, dims=dims[t].copy(), size=size[t].copy(), n_unknown=0))
            continue
        X, Y, Z = (int(v) for v in dims[t])
        occ = labels[pk.label_off[t]: pk.label_off[t] + X * Y * Z].reshape(X, Y, Z).copy()
        out.append(dict(status="ok", occ=occ, dims=dims[t].copy(), size=size[t].copy(), n_unknown=int(nunk[t])))
    return out


def annotate_tracklet_debug(batch, t: int):
    """Single tracklet with the intermediate local points / in-box flags exposed."""
    sub = type(batch)(segments=batch.segments, tracklets=[batch.tracklets[t]], voxel_size=batch.voxel_size)
    pk = PackedBatch(sub)
    B = int(pk.trk_frame_off[1])
    L = pk.L
    sn = np.ascontiguousarray(pk.sensors[pk.frame_sf])      # [B, L]
    P = int(pk.pt_off[-1])
    loc = np.full((max(P, 1), 3), np.nan, np.float32)
    keep = np.zeros(max(P, 1), np.int8)
    cap = int(pk.label_off[-1])
    labels = np.zeros(max(cap, 1), np.int32)
    dims = np.zeros(3, np.int32)
    size = np.zeros(3, np.float32)
    nunk = np.zeros(1, np.int64)
    f = lib().orc_annotate_tracklet
    f.restype = C.c_int
    st = f(C.c_int(B), _p(pk.boxes), _p(pk.trig), _p(pk.points), _p(pk.pt_off), _p(sn), C.c_int(L),
           _p(pk.incl_pool), _p(pk.ri_pool), C.c_double(pk.voxel_size), _p(dims), _p(size), _p(labels),
           C.c_long(cap), _p(nunk), _p(loc), _p(keep))
    return dict(status=STATUS.get(st, str(st)), dims=dims, size=size, labels=labels, n_unknown=int(nunk[0]),
                loc=loc[:P], keep=keep[:P].astype(bool), packed=pk)


def mean_var_grid(loc, q, dims):
    """``--save-mean-var`` block of annotate_trk (tools/occ/occ_annotate.py:627-645) with scatter_v2
    (mmdet3d/ops/sst/sst_ops.py:150-181) restated in numpy: rows of ``q`` (raw quantised coordinates, int64
    [n,3], possibly negative) are grouped by ``np.unique(axis=0)`` (sorted, like torch.unique(dim=0)); mean =
    f32 sum in point order / count (torch_scatter's CPU order); the second pass averages the squared deviations
    from the gathered means; the dense f32 [X,Y,Z,6] grid is filled by index assignment, so a negative coordinate
    wraps and the later group of the sorted list wins a shared cell (CPU index_put order)."""
    X, Y, Z = (int(v) for v in dims)
    out = np.zeros((X, Y, Z, 6), np.float32)
    if len(q) == 0:
        return out
    loc = np.ascontiguousarray(loc, np.float32)
    unq, inv = np.unique(np.asarray(q, np.int64), axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    cnt = np.bincount(inv, minlength=len(unq)).astype(np.float32)[:, None]
    s = np.zeros((len(unq), 3), np.float32)
    np.add.at(s, inv, loc)                                   # sequential f32 accumulation in point order
    mean = s / np.maximum(cnt, np.float32(1))
    dev2 = (loc - mean[inv]) ** 2
    v = np.zeros((len(unq), 3), np.float32)
    np.add.at(v, inv, dev2)
    var = v / np.maximum(cnt, np.float32(1))
    both = np.concatenate([mean, var], 1)
    for k in range(len(unq)):                                # sequential assignment: later rows overwrite
        out[unq[k, 0], unq[k, 1], unq[k, 2]] = both[k]
    return out


def annotate_mean_var(batch):
    """Per tracklet the [X,Y,Z,6] mean/variance grid of ``--save-mean-var`` (None where no file is written)."""
    res = []
    vsf = np.float32(batch.voxel_size)
    for t in range(len(batch.tracklets)):
        d = annotate_tracklet_debug(batch, t)
        if d["status"] != "ok":
            res.append(None)
            continue
        loc = d["loc"][d["keep"]]
        size = d["size"].astype(np.float32)
        min_bound = np.array([-size[0] * np.float32(0.5), -size[1] * np.float32(0.5), 0], np.float32)
        q = np.floor((loc - min_bound) / vsf).astype(np.int64)            # occ_annotate.py:425
        ok = (q < d["dims"][None].astype(np.int64)).all(1)                # :430-431
        res.append(mean_var_grid(loc[ok], q[ok], d["dims"]))
    return res


def build_range_image(points, extrinsic, inclination, size):
    """waymo_open_dataset ... build_range_image_from_point_cloud as called at
    tools/data_converter/waymo_converter.py:653-668 for one LiDAR of one frame (C restatement:
    ``orc_build_range_image``; the TF package is not vendored -> parity unpinned against TF itself).

    points f32 [n,>=3] vehicle frame; extrinsic [4,4]; inclination f32 [H] as stored (reversed here, :659)
    -> (ri f32 [H,W] min range per pixel / 0, rows int32 [n], cols int32 [n], ranges f32 [n], n_bad)."""
    H, W = (int(v) for v in size)
    E = np.asarray(extrinsic).astype(np.float64)
    v2l = np.ascontiguousarray(np.linalg.inv(E)[:3, :].reshape(12))
    incl = np.ascontiguousarray(np.asarray(inclination, np.float32)[::-1])
    pts = np.ascontiguousarray(points, np.float32)
    n = len(pts)
    ri = np.zeros((H, W), np.float32)
    rows, cols, rng = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float32)
    f = lib().orc_build_range_image
    f.restype = C.c_long
    bad = f(_p(pts), C.c_long(n), C.c_int(pts.shape[1] if pts.ndim == 2 else 3), _p(v2l),
            C.c_double(float(np.arctan2(E[1, 0], E[0, 0]))), _p(incl), C.c_int(H), C.c_int(W), _p(ri), _p(rows),
            _p(cols), _p(rng))
    return ri, rows[:n], cols[:n], rng[:n], int(bad)


def mirror_occ_label(occ):
    """MirrorOccLabel (mmdet3d/datasets/pipelines/occ_pinelines.py:88-126) on one int grid [X,Y,Z], with the
    reference's own float arithmetic for the mirror index (f32, truncation toward zero)."""
    occ = np.asarray(occ)
    X = occ.shape[0]
    mid = X // 2
    x = np.arange(X, dtype=np.int64)
    mx = ((x.astype(np.float32) + np.float32(0.5) - np.float32(mid)) * np.float32(-1.0) + np.float32(mid)).astype(np.int64)
    out = occ.copy()
    mirrored = occ[mx]                      # [X,Y,Z]: row x holds the grid's mirror row (negative index would wrap)
    unknown = occ == 0
    out[unknown] = mirrored[unknown]
    return out
