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

import ctypes as C
import os
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
    flat: Optional[np.ndarray] = None      # f32 [sum n_i, 3]: `points` as one contiguous array (views into it)

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
# the same renderer with its per-pixel loops in C (csrc/synth_render.c, OpenMP): bit-identical output,
# ~20x faster -- what makes the 10 000-tracklet job (BASELINE configs[4]) generable in seconds
# --------------------------------------------------------------------------
class _CLidar(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("max_range", C.c_double), ("o", C.c_double * 3),
                ("dirs", C.c_void_p), ("img", C.c_void_p), ("img32", C.c_void_p)]


_WIN_DTYPE = np.dtype([("ok", "<i4"), ("r_lo", "<i4"), ("r_hi", "<i4"), ("ncols", "<i4")])
_SYNTH_LIB = None


def fast_lib():
    """csrc/libocc_synth.so (built by `make` next to libocc_b200.so) or None."""
    global _SYNTH_LIB
    if _SYNTH_LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libocc_synth.so")
        if not os.path.exists(path) or os.environ.get("OCCB200_SYNTH_NUMPY"):
            _SYNTH_LIB = False
        else:
            L = C.CDLL(path)
            vp = C.c_void_p
            L.synth_render_objects.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]
            L.synth_render_objects.restype = None
            L.synth_candidate_points.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_double, vp, vp, vp, vp, vp]
            L.synth_candidate_points.restype = None
            L.synth_f64_to_f32.argtypes = [vp, vp, C.c_int64]
            L.synth_f64_to_f32.restype = None
            L.synth_repeat.argtypes = [vp, vp, C.c_int64, C.c_int]
            L.synth_repeat.restype = None
            L.synth_set_threads.argtypes = [C.c_int]
            L.synth_set_threads.restype = None
            _SYNTH_LIB = L
    return _SYNTH_LIB or None


def set_threads(n: int) -> None:
    """OpenMP threads of the C pixel loops (torchrun pins OMP_NUM_THREADS=1 for every rank)."""
    L = fast_lib()
    if L is not None:
        L.synth_set_threads(int(n))


class _FastRenderer:
    """`_Renderer` with all objects of a segment handled per call: the windows of every object are derived
    with the array expressions of `_Renderer._window` (vectorised over objects), the pixel loops run in C."""

    def __init__(self, rig, num_frames):
        self.lib = fast_lib()
        self.rig = rig
        self.B = num_frames
        self.rays = [_pixel_rays(l) for l in rig]
        self.dirs = [np.ascontiguousarray(d) for _, d, _ in self.rays]
        self.images = []
        for lidar, (o, d, fov) in zip(rig, self.rays):
            base = np.ascontiguousarray(_base_image(o, d, fov, lidar["max_range"]))
            img = np.empty((num_frames,) + base.shape, np.float64)
            self.lib.synth_repeat(base.ctypes.data, img.ctypes.data, base.size, num_frames)
            self.images.append(img)
        self.images32 = None
        self.incl_flip = [l["incl"].astype(np.float64)[::-1] for l in rig]

    def _lidars(self):
        arr = (_CLidar * len(self.rig))()
        for li, lidar in enumerate(self.rig):
            a = arr[li]
            a.H, a.W, a.max_range = lidar["H"], lidar["W"], float(lidar["max_range"])
            for k in range(3):
                a.o[k] = float(self.rays[li][0][k])
            a.dirs = self.dirs[li].ctypes.data
            a.img = self.images[li].ctypes.data if self.images[li] is not None else None
            a.img32 = self.images32[li].ctypes.data if self.images32 is not None else None
        return arr

    def _windows(self, boxes, pad):
        """`_Renderer._window` for boxes [O,B,7] and every LiDAR: (win [O,L] records, col0 [O,L,B])."""
        O, B = boxes.shape[:2]
        L = len(self.rig)
        win = np.zeros((O, L), _WIN_DTYPE)
        col0 = np.zeros((O, L, B), np.int64)
        if O == 0:
            return win, col0
        corners = _box_corners(boxes, pad_xy=pad, pad_z=pad)                  # O,B,8,3
        ctr = boxes[..., :3].copy()
        ctr[..., 2] += 0.5 * boxes[..., 5]
        diag = 0.5 * np.linalg.norm(boxes[..., 3:6], axis=-1).max(-1)          # O
        for li, lidar in enumerate(self.rig):
            H, W = lidar["H"], lidar["W"]
            inc, col, rng = _project_window(lidar, corners)                    # O,B,8
            _, col_c, rng_c = _project_window(lidar, ctr)                      # O,B
            ok = ~(rng.min((1, 2)) < 1.0)
            ok &= ~(rng_c.min(1) - diag - pad > lidar["max_range"])
            dcol = (col - col_c[..., None] + W / 2) % W - W / 2
            half = np.ceil(np.abs(dcol).max((1, 2))).astype(np.int64) + 2      # O
            full = 2 * half + 1 >= W
            incf = self.incl_flip[li]
            step = np.abs(np.diff(incf)).max()
            lo, hi = inc.min((1, 2)) - step, inc.max((1, 2)) + step            # O
            sel = (incf[None] >= lo[:, None]) & (incf[None] <= hi[:, None])    # O,H (incf is monotone: one run)
            ok &= sel.any(1)
            win["ok"][:, li] = ok
            win["r_lo"][:, li] = sel.argmax(1)
            win["r_hi"][:, li] = H - 1 - sel[:, ::-1].argmax(1)
            win["ncols"][:, li] = np.where(full, W, 2 * half + 1)
            col0[:, li] = np.where(full[:, None], 0, np.round(col_c).astype(np.int64) - half[:, None])
        return win, col0

    def add_objects(self, objs):
        """objs: list of (boxes f64 [B,7], size_true [3], shape dict, frames bool [B] or None)."""
        if not objs:
            return
        boxes = np.ascontiguousarray(np.stack([o[0] for o in objs]), np.float64)
        rz = boxes[:, :, None, None, :][..., 6]
        cs = np.ascontiguousarray(np.stack([np.cos(rz), np.sin(rz)], -1).reshape(len(objs), self.B, 2))
        size = np.ascontiguousarray(np.stack([np.asarray(o[1], np.float64).reshape(3) for o in objs]))
        keys = ("shrink", "body_h", "cab_w", "cab_back", "cab_front")
        shape = np.ascontiguousarray([[float(o[2][k]) for k in keys] for o in objs], np.float64)
        frames = np.ascontiguousarray(np.stack([np.ones(self.B, bool) if o[3] is None else o[3] for o in objs])
                                      .astype(np.uint8))
        win, col0 = self._windows(boxes, 0.0)
        lid = self._lidars()
        self.lib.synth_render_objects(C.addressof(lid), len(self.rig), self.B, len(objs),
                                      boxes.ctypes.data, cs.ctypes.data, size.ctypes.data, shape.ctypes.data,
                                      frames.ctypes.data, win.ctypes.data, col0.ctypes.data)

    def finish(self):
        """f64 working images -> the segment's f32 range images."""
        self.images32 = []
        for img in self.images:
            out = np.empty(img.shape, np.float32)
            self.lib.synth_f64_to_f32(img.ctypes.data, out.ctypes.data, img.size)
            self.images32.append(out)
        self.images = [None] * len(self.images32)      # the f64 copies are not needed any more
        return self.images32

    def candidate_points_all(self, boxes, pad=1.0):
        """boxes f64 [O,B,7] -> (flat f32 [P,3], counts i64 [O,B]) in (object, frame, LiDAR, row, col) order."""
        boxes = np.ascontiguousarray(boxes, np.float64)
        O = boxes.shape[0]
        rz = boxes[:, :, None, None, :][..., 6]
        cs = np.ascontiguousarray(np.stack([np.cos(rz), np.sin(rz)], -1).reshape(O, self.B, 2))
        win, col0 = self._windows(boxes, pad)
        counts = np.zeros((O, self.B), np.int64)
        lid = self._lidars()
        args = (C.addressof(lid), len(self.rig), self.B, O, boxes.ctypes.data, cs.ctypes.data, float(pad),
                win.ctypes.data, col0.ctypes.data)
        self.lib.synth_candidate_points(*args, counts.ctypes.data, None, None)
        off = np.zeros(O * self.B + 1, np.int64)
        np.cumsum(counts.reshape(-1), out=off[1:])
        flat = np.empty((int(off[-1]), 3), np.float32)
        self.lib.synth_candidate_points(*args, None, off.ctypes.data, flat.ctypes.data)
        return flat, counts


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
        if len(others) and np.linalg.norm(xy[None] - others[:, :, :2], axis=2).min() < 0.5 * size[1] + 3.0:
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
                 small=False, occluder_prob=0.3, fast=None):
    """One segment with ``num_objects`` tracklets.  Returns (Segment, [boxes f32 [B,7]], [points], [flat]).

    ``fast`` selects the C pixel loops (default: when csrc/libocc_synth.so is built); both paths consume the
    random stream identically and produce the same bits."""
    if fast is None:
        fast = fast_lib() is not None
    rig = lidar_rig(rng, small=small)
    tracks, sizes, shapes = [], [], []
    placed = np.zeros((0, num_frames, 7))
    for _ in range(num_objects):
        boxes, size = _sample_track(rng, num_frames, kind, extent, placed)
        tracks.append(boxes)
        placed = np.concatenate([placed, boxes[None]], 0)
        sizes.append(size)
        shapes.append(dict(shrink=rng.uniform(0.88, 0.97), body_h=rng.uniform(0.45, 0.65),
                           cab_w=rng.uniform(0.75, 0.9), cab_back=rng.uniform(0.5, 0.9),
                           cab_front=rng.uniform(0.1, 0.5)))
    boxes32 = [b.astype(np.float32) for b in tracks]
    objs = [(b32.astype(np.float64), size, shape, None) for b32, size, shape in zip(boxes32, sizes, shapes)]
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
        objs.append((occ, occ[0, 3:6].copy(), dict(shrink=1.0, body_h=1.0, cab_w=0.0, cab_back=0.0, cab_front=0.0),
                     frames))
    if fast:
        ren = _FastRenderer(rig, num_frames)
        ren.add_objects(objs)
        images = ren.finish()
        flat, counts = ren.candidate_points_all(np.stack([o[0] for o in objs[:num_objects]]) if num_objects
                                                else np.zeros((0, num_frames, 7)))
        bounds = np.concatenate([[0], np.cumsum(counts.reshape(-1))])
        points, flats = [], []
        for k in range(num_objects):
            lo = k * num_frames
            points.append([flat[bounds[lo + b]:bounds[lo + b + 1]] for b in range(num_frames)])
            flats.append(flat[bounds[lo]:bounds[lo + num_frames]])
    else:
        ren = _Renderer(rig, num_frames)
        for b64, size, shape, frames in objs:
            ren.add_object(b64, np.broadcast_to(size, (num_frames, 3)), shape, frames)
        points = [ren.candidate_points(b32.astype(np.float64)) for b32 in boxes32]
        flats = [None] * num_objects
        images = [img.astype(np.float32) for img in ren.images]
    E = np.stack([l["extrinsic"] for l in rig], 0)
    seg = Segment(
        extrinsics=np.ascontiguousarray(np.broadcast_to(E[None], (num_frames, NUM_LIDARS, 4, 4))).copy(),
        inclinations=[l["incl"].copy() for l in rig],
        range_images=images,
    )
    return seg, boxes32, points, flats


def segment_rng(seed, index):
    """Random stream of segment ``index`` of a batch: segments are independent, so a rank can generate just its own."""
    return np.random.default_rng(seed) if index == 0 else np.random.default_rng([int(seed), int(index)])


def make_batch(num_tracklets, num_frames, voxel_size=0.2, kind="vehicle", seed=0,
               tracklets_per_segment=None, small=False, extent=None, occluder_prob=0.3, fast=None,
               only_segments=None):
    """``num_tracklets`` tracklets of ``num_frames`` frames spread over shared segments.

    ``only_segments``: iterable of segment indices to generate (the others are skipped entirely; the tracklets
    returned are those of the generated segments, ``Tracklet.segment`` indexing the returned segment list, and
    ``meta['global_tracklets']`` their indices in the full batch)."""
    tps = tracklets_per_segment or num_tracklets
    if extent is None:
        extent = 40.0
    nseg = (num_tracklets + tps - 1) // tps if num_tracklets else 0
    want = list(range(nseg)) if only_segments is None else sorted(int(i) for i in only_segments)
    segments, tracklets, gidx = [], [], []
    for i in want:
        k = min(tps, num_tracklets - i * tps)
        for attempt in range(16):      # a crowded draw may leave no room for the last tracks: redraw the segment
            rng = segment_rng(seed, i) if attempt == 0 else np.random.default_rng([int(seed), int(i), attempt])
            try:
                seg, boxes, points, flats = make_segment(rng, k, num_frames, kind=kind, extent=extent, small=small,
                                                         occluder_prob=occluder_prob, fast=fast)
                break
            except RuntimeError:
                if attempt == 15:
                    raise
        si = len(segments)
        segments.append(seg)
        for j, (b, p, fl) in enumerate(zip(boxes, points, flats)):
            tracklets.append(Tracklet(boxes=b, points=p, segment=si,
                                      frame_ids=np.arange(num_frames, dtype=np.int32), kind=kind, flat=fl))
            gidx.append(i * tps + j)
    return TrackletBatch(segments=segments, tracklets=tracklets, voxel_size=float(voxel_size),
                         meta=dict(seed=seed, kind=kind, num_frames=num_frames, num_segments=nseg,
                                   segment_ids=want, global_tracklets=gidx))


# BASELINE.json configs (SURVEY.md section 8 sizes)
C5_TRACKLETS, C5_PER_SEGMENT = 10000, 64


def config_batch(name: str, seed: int = 0, small: bool = False, only_segments=None) -> TrackletBatch:
    name = name.lower()
    if name == "c1":      # 1 vehicle tracklet, 20 frames, 0.2 m
        return make_batch(1, 20, 0.2, "vehicle", seed, small=small)
    if name == "c2":      # 64 vehicle tracklets x 40 frames, 0.2 m, one segment
        return make_batch(64, 40, 0.2, "vehicle", seed, tracklets_per_segment=64, small=small)
    if name == "c3":      # truck / bus boxes at 0.1 m
        return make_batch(16, 40, 0.1, "large", seed, tracklets_per_segment=16, small=small)
    if name == "c5":      # 10k C2-shaped tracklets
        return make_batch(C5_TRACKLETS, 40, 0.2, "vehicle", seed, tracklets_per_segment=C5_PER_SEGMENT, small=small,
                          only_segments=only_segments)
    if name == "c5s":     # 1/10 of C5 for quick scaling checks
        return make_batch(1024, 40, 0.2, "vehicle", seed, tracklets_per_segment=64, small=small,
                          only_segments=only_segments)
    raise ValueError(f"unknown config {name}")


def scatter_inputs(num_tracklets=32, num_frames=32, max_pts=1024, channels=5, seed=0, full=False):
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
            n = max_pts if full else int(rng.integers(max_pts // 4, max_pts + 1))   # full: the config's upper bound
            xyz = ctr + (rng.random((n, 3)) - 0.5) * size * 1.2
            feat = np.tanh(rng.standard_normal((n, max(channels - 3, 0))))
            pts.append(np.concatenate([xyz, feat], 1)[:, :channels])
            bidx.append(np.full(n, t, np.int32))
    return np.concatenate(pts, 0).astype(np.float32), np.concatenate(bidx, 0)
