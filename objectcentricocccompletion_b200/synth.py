"""Seeded synthetic tracklet scenes for the point -> occupancy path.

There is no Waymo data (and no network) in this environment, so tests and the
benchmark run on synthetic *segments* that have the same structure as what
``tools/occ/occ_annotate.py`` reads (reference file:line in brackets):

* a segment is ``B`` frames, each with 5 LiDARs ``TOP, FRONT, SIDE_LEFT,
  SIDE_RIGHT, REAR`` [occ_annotate.py:235], each LiDAR with an f32 4x4
  extrinsic, an f32 beam-inclination table (ascending as stored; the annotate
  path flips it [occ_annotate.py:528]) and an f32 range image ``H x W`` with
  ``0`` = no return [occ_annotate.py:512-514];
* a tracklet is one box ``(x, y, z_bottom, x_size, y_size, z_size, yaw)`` per
  frame in that frame's ego coordinates [tools/ctrl/utils.py:36-43] plus the
  candidate LiDAR returns of the frame around it (f32 xyz, ego frame)
  [occ_annotate.py:101-107].

Range images are rendered analytically (ground plane, far cylinder wall,
two-box vehicle shells, occasional occluder slabs) with the *inverse* of the
pixel model of ``point_cloud_to_range_image_idx`` [occ_annotate.py:141-201] so
occupied / free / unknown all occur.  Points are the returns of the rendered
pixels that land inside the tracklet box enlarged by 1 m.

Everything is NumPy on the host and depends only on the seed.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

LIDAR_NAMES = ["TOP", "FRONT", "SIDE_LEFT", "SIDE_RIGHT", "REAR"]
NUM_LIDARS = len(LIDAR_NAMES)


# --------------------------------------------------------------------------
# data containers
# --------------------------------------------------------------------------
@dataclass
class Segment:
    """Per-frame sensor data shared by every tracklet of a segment."""

    extrinsics: np.ndarray                 # f32 [B, L, 4, 4]
    inclinations: List[np.ndarray]         # L x f32 [H_c]  (ascending, as stored)
    range_images: List[np.ndarray]         # L x f32 [B, H_c, W_c]

    @property
    def num_frames(self) -> int:
        return self.extrinsics.shape[0]


@dataclass
class Tracklet:
    """One object track: a box and candidate points per frame."""

    boxes: np.ndarray                      # f32 [B, 7]
    points: List[np.ndarray]               # B x f32 [n_i, 3]
    segment: int                           # index into TrackletBatch.segments
    frame_ids: np.ndarray                  # i32 [B] frame index inside the segment
    kind: str = "vehicle"

    def __len__(self) -> int:
        return self.boxes.shape[0]


@dataclass
class TrackletBatch:
    segments: List[Segment]
    tracklets: List[Tracklet]
    voxel_size: float = 0.2
    meta: dict = field(default_factory=dict)

    def __len__(self) -> int:
        return len(self.tracklets)


# --------------------------------------------------------------------------
# LiDAR rig
# --------------------------------------------------------------------------
def _rot_zyx(yaw: float, pitch: float, roll: float) -> np.ndarray:
    cz, sz = np.cos(yaw), np.sin(yaw)
    cy, sy = np.cos(pitch), np.sin(pitch)
    cx, sx = np.cos(roll), np.sin(roll)
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]])
    ry = np.array([[cy, 0, sy], [0, 1.0, 0], [-sy, 0, cy]])
    rx = np.array([[1.0, 0, 0], [0, cx, -sx], [0, sx, cx]])
    return rz @ ry @ rx


def lidar_rig(rng: np.random.Generator, small: bool = False):
    """Waymo-like 5-LiDAR rig.  ``small`` shrinks the images for unit tests."""
    mounts = [
        # name, xyz, yaw, H, W, incl range (deg), max range, az half-fov (rad)
        ("TOP", (1.43, 0.0, 2.184), 0.0, 64, 2650, (-17.6, 2.4), 75.0, np.pi),
        ("FRONT", (4.07, 0.0, 0.691), 0.0, 200, 600, (-90.0, 30.0), 20.0, 1.75),
        ("SIDE_LEFT", (3.245, 1.025, 0.981), np.pi / 2, 200, 600, (-90.0, 30.0), 20.0, 1.75),
        ("SIDE_RIGHT", (3.245, -1.025, 0.981), -np.pi / 2, 200, 600, (-90.0, 30.0), 20.0, 1.75),
        ("REAR", (-1.154, 0.0, 0.466), np.pi, 200, 600, (-90.0, 30.0), 20.0, 1.75),
    ]
    rig = []
    for name, xyz, yaw, H, W, (lo, hi), max_range, fov in mounts:
        if small:
            H, W = (16, 331) if name == "TOP" else (25, 75)
        yaw_j = yaw + rng.uniform(-0.02, 0.02)
        pitch = rng.uniform(-0.02, 0.02)
        roll = rng.uniform(-0.02, 0.02)
        E = np.eye(4)
        E[:3, :3] = _rot_zyx(yaw_j, pitch, roll)
        E[:3, 3] = np.asarray(xyz) + rng.uniform(-0.01, 0.01, 3)
        if name == "TOP":
            # non-uniform beams, denser towards the horizon
            t = np.linspace(0.0, 1.0, H) ** 0.8
            incl = np.deg2rad(lo + (hi - lo) * t)
        else:
            incl = np.deg2rad((np.arange(H) + 0.5) / H * (hi - lo) + lo)
        rig.append(
            dict(name=name, extrinsic=E.astype(np.float32), incl=incl.astype(np.float32),
                 H=H, W=W, max_range=max_range, az_fov=fov)
        )
    return rig


def _pixel_rays(lidar):
    """Vehicle-frame origin and unit directions [H, W, 3] of every pixel.

    Inverse of the pixel model in occ_annotate.py:165-193: row r looks along
    inclination ``flip(incl)[r]``; column c along azimuth
    ``2*pi*(W-0.5-c)/W - pi`` minus the sensor yaw ``atan2(E[1,0], E[0,0])``.
    """
    E = lidar["extrinsic"].astype(np.float64)
    H, W = lidar["H"], lidar["W"]
    incl = lidar["incl"].astype(np.float64)[::-1]
    azc = np.arctan2(E[1, 0], E[0, 0])
    az = 2.0 * np.pi * (W - 0.5 - np.arange(W)) / W - np.pi - azc
    ci, si = np.cos(incl)[:, None], np.sin(incl)[:, None]
    d_s = np.stack([ci * np.cos(az)[None], ci * np.sin(az)[None], np.broadcast_to(si, (H, W))], -1)
    d_v = d_s @ E[:3, :3].T
    in_fov = np.abs((az + np.pi) % (2 * np.pi) - np.pi) <= lidar["az_fov"]
    return E[:3, 3].copy(), d_v, in_fov


def _base_image(origin, dirs, in_fov, max_range):
    """Ground plane z=0 and a cylinder wall of radius 80 m around the ego origin."""
    dz = dirs[..., 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        t_ground = np.where(dz < -1e-9, -origin[2] / dz, np.inf)
        a = dirs[..., 0] ** 2 + dirs[..., 1] ** 2
        b = 2.0 * (origin[0] * dirs[..., 0] + origin[1] * dirs[..., 1])
        c = origin[0] ** 2 + origin[1] ** 2 - 80.0 ** 2
        disc = np.maximum(b * b - 4 * a * c, 0.0)
        t_wall = np.where(a > 1e-12, (-b + np.sqrt(disc)) / (2 * a), np.inf)
    wall_z = origin[2] + t_wall * dz
    t_wall = np.where(wall_z > 8.0, np.inf, t_wall)       # sky: no return
    t = np.minimum(t_ground, t_wall)
    t = np.where(np.isfinite(t) & (t <= max_range) & in_fov[None, :], t, 0.0)
    return t


# --------------------------------------------------------------------------
# object shells
# --------------------------------------------------------------------------
def _to_box_frame(p, box, is_dir=False):
    """Ego -> box frame, same sense as occ_annotate.py:117-122 / lidar_box3d.py:163-184.

    ``p`` [..., 3]; ``box`` [..., 7] broadcastable.  x' runs along x_size, y' along y_size.
    """
    rz = box[..., 6]
    c, s = np.cos(rz), np.sin(rz)
    if is_dir:
        tx, ty, tz = p[..., 0], p[..., 1], p[..., 2]
    else:
        tx, ty, tz = p[..., 0] - box[..., 0], p[..., 1] - box[..., 1], p[..., 2] - box[..., 2]
    return np.stack([tx * c - ty * s, tx * s + ty * c, tz], -1)


def _slab_hit(o, d, lo, hi):
    """Ray/axis-aligned-box entry distance (inf when missed).  o,d [...,3]; lo,hi [...,3]."""
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / np.where(np.abs(d) < 1e-12, 1e-12, d)
        t0 = (lo - o) * inv
        t1 = (hi - o) * inv
    lo_t = np.minimum(t0, t1)
    hi_t = np.maximum(t0, t1)
    tn = np.maximum(np.maximum(lo_t[..., 0], lo_t[..., 1]), lo_t[..., 2])
    tf = np.minimum(np.minimum(hi_t[..., 0], hi_t[..., 1]), hi_t[..., 2])
    hit = (tf >= np.maximum(tn, 0.0))
    return np.where(hit, np.where(tn > 0, tn, tf), np.inf)


def _shell_hit(o_l, d_l, size, shape):
    """Two-box vehicle shell (body + cabin) in the box frame.  size [...,3]."""
    w, l, h = size[..., 0], size[..., 1], size[..., 2]
    sh = shape  # dict of scalars
    hb = h * sh["body_h"]
    lo1 = np.stack([-0.5 * w * sh["shrink"], -0.5 * l * sh["shrink"], np.zeros_like(h) + 0.02], -1)
    hi1 = np.stack([0.5 * w * sh["shrink"], 0.5 * l * sh["shrink"], hb], -1)
    lo2 = np.stack([-0.5 * w * sh["cab_w"], -0.5 * l * sh["cab_back"], hb], -1)
    hi2 = np.stack([0.5 * w * sh["cab_w"], 0.5 * l * sh["cab_front"], h * sh["shrink"]], -1)
    return np.minimum(_slab_hit(o_l, d_l, lo1, hi1), _slab_hit(o_l, d_l, lo2, hi2))


def _project_window(lidar, pts_v):
    """(row, col) float indices of vehicle-frame points [..., 3] for one LiDAR."""
    E = lidar["extrinsic"].astype(np.float64)
    H, W = lidar["H"], lidar["W"]
    p = (pts_v - E[:3, 3]) @ E[:3, :3]
    azc = np.arctan2(E[1, 0], E[0, 0])
    az = np.arctan2(p[..., 1], p[..., 0]) + azc
    az = (az + np.pi) % (2 * np.pi) - np.pi
    col = W - 0.5 - (az + np.pi) / (2 * np.pi) * W
    inc = np.arctan2(p[..., 2], np.hypot(p[..., 0], p[..., 1]))
    rng = np.linalg.norm(p, axis=-1)
    return inc, col, rng


def _box_corners(box, pad_xy=0.0, pad_z=0.0):
    """Ego-frame corners [..., 8, 3] of boxes [..., 7] (optionally enlarged)."""
    w = box[..., 3] + 2 * pad_xy
    l = box[..., 4] + 2 * pad_xy
    h = box[..., 5] + 2 * pad_z
    sx = np.array([-1, -1, -1, -1, 1, 1, 1, 1]) * 0.5
    sy = np.array([-1, -1, 1, 1, -1, -1, 1, 1]) * 0.5
    sz = np.array([0, 1, 0, 1, 0, 1, 0, 1.0])
    xl = w[..., None] * sx
    yl = l[..., None] * sy
    zl = h[..., None] * sz - pad_z
    rz = box[..., 6][..., None]
    c, s = np.cos(rz), np.sin(rz)
    # inverse of _to_box_frame:  tx = x' c + y' s ; ty = -x' s + y' c
    x = xl * c + yl * s + box[..., 0][..., None]
    y = -xl * s + yl * c + box[..., 1][..., None]
    z = zl + box[..., 2][..., None]
    return np.stack([x, y, z], -1)


class _Renderer:
    """Renders object shells of one segment into its range images, frame-vectorised."""

    def __init__(self, rig, num_frames):
        self.rig = rig
        self.B = num_frames
        self.rays = [_pixel_rays(l) for l in rig]
        self.images = []
        for lidar, (o, d, fov) in zip(rig, self.rays):
            base = _base_image(o, d, fov, lidar["max_range"])
            self.images.append(np.repeat(base[None], num_frames, 0))
        self.incl_flip = [l["incl"].astype(np.float64)[::-1] for l in rig]

    def _window(self, li, boxes, pad):
        """Common (rows, cols[B, w]) pixel window covering ``boxes`` [B,7] in every frame."""
        lidar = self.rig[li]
        H, W = lidar["H"], lidar["W"]
        corners = _box_corners(boxes, pad_xy=pad, pad_z=pad)             # B,8,3
        inc, col, rng = _project_window(lidar, corners)
        ctr = boxes[:, :3].copy()
        ctr[:, 2] += 0.5 * boxes[:, 5]
        _, col_c, rng_c = _project_window(lidar, ctr)
        if rng.min() < 1.0:
            return None
        if rng_c.min() - 0.5 * np.linalg.norm(boxes[:, 3:6], axis=1).max() - pad > lidar["max_range"]:
            return "far"
        dcol = (col - col_c[:, None] + W / 2) % W - W / 2                  # B,8
        half = int(np.ceil(np.abs(dcol).max())) + 2
        if 2 * half + 1 >= W:
            cols = np.broadcast_to(np.arange(W)[None], (self.B, W))
        else:
            cols = (np.round(col_c).astype(np.int64)[:, None] + np.arange(-half, half + 1)[None]) % W
        incf = self.incl_flip[li]
        step = np.abs(np.diff(incf)).max()
        rows = np.nonzero((incf >= inc.min() - step) & (incf <= inc.max() + step))[0]
        if rows.size == 0:
            return "far"
        return rows, cols

    def _window_rays(self, li, rows, cols):
        o, d, _ = self.rays[li]
        dirs = d[rows[None, :, None], cols[:, None, :]]                    # B,h,w,3
        return o, dirs

    def add_object(self, boxes, size_true, shape, frames=None):
        """Render a shell that follows ``boxes`` [B,7]; ``frames`` bool [B] limits where."""
        for li, lidar in enumerate(self.rig):
            win = self._window(li, boxes, 0.0)
            if win is None or isinstance(win, str):
                continue
            rows, cols = win
            o, dirs = self._window_rays(li, rows, cols)
            bb = boxes[:, None, None, :]
            o_l = _to_box_frame(np.broadcast_to(o, dirs.shape), bb)
            d_l = _to_box_frame(dirs, bb, is_dir=True)
            t = _shell_hit(o_l, d_l, size_true[:, None, None, :], shape)
            t = np.where(t <= lidar["max_range"], t, np.inf)
            if frames is not None:
                t = np.where(frames[:, None, None], t, np.inf)
            img = self.images[li]
            fi = np.arange(self.B)[:, None, None]
            ri = rows[None, :, None]
            ci = cols[:, None, :]
            cur = img[fi, ri, ci]
            cur_inf = np.where(cur > 0, cur, np.inf)
            new = np.minimum(cur_inf, t)
            img[fi, ri, ci] = np.where(np.isfinite(new), new, 0.0)

    def candidate_points(self, boxes, pad=1.0):
        """Returns of rendered pixels inside ``boxes`` enlarged by ``pad`` — list of B [n,3] f32."""
        per_frame = [[] for _ in range(self.B)]
        for li, lidar in enumerate(self.rig):
            win = self._window(li, boxes, pad)
            if win is None or isinstance(win, str):
                continue
            rows, cols = win
            o, dirs = self._window_rays(li, rows, cols)
            fi = np.arange(self.B)[:, None, None]
            r = self.images[li][fi, rows[None, :, None], cols[:, None, :]].astype(np.float32).astype(np.float64)
            pts = o + dirs * r[..., None]
            loc = _to_box_frame(pts, boxes[:, None, None, :])
            inside = (
                (r > 0)
                & (np.abs(loc[..., 0]) <= 0.5 * boxes[:, 3, None, None] + pad)
                & (np.abs(loc[..., 1]) <= 0.5 * boxes[:, 4, None, None] + pad)
                & (loc[..., 2] >= -0.5 * pad)
                & (loc[..., 2] <= boxes[:, 5, None, None] + 0.5 * pad)
            )
            for b in range(self.B):
                m = inside[b]
                if m.any():
                    per_frame[b].append(pts[b][m])
        out = []
        for b in range(self.B):
            if per_frame[b]:
                out.append(np.concatenate(per_frame[b], 0).astype(np.float32))
            else:
                out.append(np.zeros((0, 3), np.float32))
        return out


# --------------------------------------------------------------------------
# trajectories
# --------------------------------------------------------------------------
_KINDS = {
    # (w, l, h) mean; relative sigma
    "vehicle": ((2.1, 4.8, 1.8), 0.05),
}


def _sample_size(rng, kind):
    if kind == "vehicle":
        mean, sig = _KINDS["vehicle"]
        return np.asarray(mean) * (1.0 + sig * rng.standard_normal(3).clip(-2, 2))
    if kind == "large":      # truck / bus
        return np.array([rng.uniform(2.5, 3.0), rng.uniform(8.0, 14.0), rng.uniform(3.0, 4.0)])
    raise ValueError(kind)


def _sample_track(rng, B, kind, extent, others, min_range=9.0, max_range=40.0):
    """Per-frame boxes f64 [B,7] that stay between ``min_range`` and ``max_range`` of the ego origin."""
    for _ in range(2000):
        size = _sample_size(rng, kind)
        xy0 = rng.uniform(-extent, extent, 2)
        yaw0 = rng.uniform(-np.pi, np.pi)
        speed = rng.uniform(0.0, 15.0) * (rng.random() < 0.7)
        drift = rng.uniform(-0.02, 0.02)
        zb = rng.uniform(-0.1, 0.2)
        yaw = yaw0 + drift * np.arange(B) + 0.003 * rng.standard_normal(B)
        # y_size (length) axis of the box in the ego frame (see _to_box_frame): (sin rz, cos rz)
        step = speed * 0.1 * np.stack([np.sin(yaw), np.cos(yaw)], -1)
        xy = xy0 + np.cumsum(step, 0) - step[0]
        dist = np.linalg.norm(xy, axis=1)
        if dist.min() < min_range + 0.5 * size[1] or dist.max() > max_range:
            continue
        if others and min(np.linalg.norm(xy - o[:, :2], axis=1).min() for o in others) < 0.5 * size[1] + 3.0:
            continue
        boxes = np.zeros((B, 7))
        boxes[:, :2] = xy
        boxes[:, 2] = zb + 0.01 * rng.standard_normal(B)
        boxes[:, 3:6] = size[None] * (1.0 + 0.02 * rng.uniform(-1, 1, (B, 3)))
        boxes[:, 6] = (yaw + np.pi) % (2 * np.pi) - np.pi
        return boxes, size
    raise RuntimeError("could not place a track; lower the density")


# --------------------------------------------------------------------------
# public builders
# --------------------------------------------------------------------------
def make_segment(rng, num_objects, num_frames, kind="vehicle", extent=40.0,
                 small=False, occluder_prob=0.3):
    """One segment with ``num_objects`` tracklets.  Returns (Segment, [boxes f32 [B,7]], [points])."""
    rig = lidar_rig(rng, small=small)
    ren = _Renderer(rig, num_frames)
    tracks, sizes, shapes = [], [], []
    for _ in range(num_objects):
        boxes, size = _sample_track(rng, num_frames, kind, extent, tracks)
        tracks.append(boxes)
        sizes.append(size)
        shapes.append(dict(shrink=rng.uniform(0.88, 0.97), body_h=rng.uniform(0.45, 0.65),
                           cab_w=rng.uniform(0.75, 0.9), cab_back=rng.uniform(0.5, 0.9),
                           cab_front=rng.uniform(0.1, 0.5)))
    boxes32 = [b.astype(np.float32) for b in tracks]
    for b32, size, shape in zip(boxes32, sizes, shapes):
        b64 = b32.astype(np.float64)
        ren.add_object(b64, np.broadcast_to(size, (num_frames, 3)), shape)
    # occluder slabs between the TOP LiDAR and a random object, in a subset of frames
    top_o = rig[0]["extrinsic"][:3, 3].astype(np.float64)
    for b32 in boxes32:
        frames = rng.random(num_frames) < occluder_prob
        if not frames.any():
            continue
        b64 = b32.astype(np.float64)
        frac = rng.uniform(0.35, 0.65)
        occ = np.zeros((num_frames, 7))
        occ[:, :2] = top_o[:2] + frac * (b64[:, :2] - top_o[:2])
        occ[:, 2] = 0.0
        occ[:, 3:6] = np.array([0.3, rng.uniform(1.0, 3.0), rng.uniform(1.0, 2.5)])
        occ[:, 6] = rng.uniform(-np.pi, np.pi)
        if np.linalg.norm(occ[:, :2], axis=1).min() < 5.5:
            continue
        ren.add_object(occ, occ[:, 3:6], dict(shrink=1.0, body_h=1.0, cab_w=0.0, cab_back=0.0, cab_front=0.0), frames)
    points = [ren.candidate_points(b32.astype(np.float64)) for b32 in boxes32]
    E = np.stack([l["extrinsic"] for l in rig], 0)
    seg = Segment(
        extrinsics=np.ascontiguousarray(np.broadcast_to(E[None], (num_frames, NUM_LIDARS, 4, 4))).copy(),
        inclinations=[l["incl"].copy() for l in rig],
        range_images=[img.astype(np.float32) for img in ren.images],
    )
    return seg, boxes32, points


def make_batch(num_tracklets, num_frames, voxel_size=0.2, kind="vehicle", seed=0,
               tracklets_per_segment=None, small=False, extent=None, occluder_prob=0.3):
    """``num_tracklets`` tracklets of ``num_frames`` frames spread over shared segments."""
    rng = np.random.default_rng(seed)
    tps = tracklets_per_segment or num_tracklets
    if extent is None:
        extent = 40.0
    segments, tracklets = [], []
    remaining = num_tracklets
    while remaining > 0:
        k = min(tps, remaining)
        seg, boxes, points = make_segment(rng, k, num_frames, kind=kind, extent=extent,
                                          small=small, occluder_prob=occluder_prob)
        si = len(segments)
        segments.append(seg)
        for b, p in zip(boxes, points):
            tracklets.append(Tracklet(boxes=b, points=p, segment=si,
                                      frame_ids=np.arange(num_frames, dtype=np.int32), kind=kind))
        remaining -= k
    return TrackletBatch(segments=segments, tracklets=tracklets, voxel_size=float(voxel_size),
                         meta=dict(seed=seed, kind=kind, num_frames=num_frames))


# BASELINE.json configs (SURVEY.md section 8 sizes)
def config_batch(name: str, seed: int = 0, small: bool = False) -> TrackletBatch:
    name = name.lower()
    if name == "c1":      # 1 vehicle tracklet, 20 frames, 0.2 m
        return make_batch(1, 20, 0.2, "vehicle", seed, small=small)
    if name == "c2":      # 64 vehicle tracklets x 40 frames, 0.2 m, one segment
        return make_batch(64, 40, 0.2, "vehicle", seed, tracklets_per_segment=64, small=small)
    if name == "c3":      # truck / bus boxes at 0.1 m
        return make_batch(16, 40, 0.1, "large", seed, tracklets_per_segment=16, small=small)
    if name == "c5":      # 10k C2-shaped tracklets
        return make_batch(10000, 40, 0.2, "vehicle", seed, tracklets_per_segment=64, small=small)
    if name == "c5s":     # 1/10 of C5 for quick scaling checks
        return make_batch(1024, 40, 0.2, "vehicle", seed, tracklets_per_segment=64, small=small)
    raise ValueError(f"unknown config {name}")


def scatter_inputs(num_tracklets=32, num_frames=32, max_pts=1024, channels=5, seed=0):
    """C4: aggregated tracklet points for Voxelization + DynamicScatter.

    Returns points f32 [N, channels] (xyz, intensity, elongation) in a +-204.8 m / -4..8 m
    range [configs/ococc/ococcnet.py:8-9] and the per-point tracklet (batch) index, sorted.
    """
    rng = np.random.default_rng(seed)
    pts, bidx = [], []
    for t in range(num_tracklets):
        ctr = np.array([rng.uniform(-180, 180), rng.uniform(-180, 180), rng.uniform(-1, 1)])
        size = _sample_size(rng, "vehicle")
        for _ in range(num_frames):
            n = int(rng.integers(max_pts // 4, max_pts + 1))
            xyz = ctr + (rng.random((n, 3)) - 0.5) * size * 1.2
            feat = np.tanh(rng.standard_normal((n, max(channels - 3, 0))))
            pts.append(np.concatenate([xyz, feat], 1)[:, :channels])
            bidx.append(np.full(n, t, np.int32))
    return np.concatenate(pts, 0).astype(np.float32), np.concatenate(bidx, 0)
