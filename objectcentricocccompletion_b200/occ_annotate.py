"""Host side of the batched tracklet annotation -- mirrors ``tools/occ/occ_annotate.py``.

The reference annotates one tracklet at a time with ~500 small torch launches
(``OccAnnotator.annotate_trk``, occ_annotate.py:319-647).  Here the same result
for a whole batch of tracklets comes from ONE call into libocc_b200.so
(``occb200_annotate_batch``): the host only *packs* -- it evaluates the handful
of scalar operations the reference itself performs on the host side of its
tensor code, with the same torch-CPU / libm calls, so that the device never has
to guess their bits:

* ``torch.sin/cos(+-yaw)`` f32 per frame  (lidar_box3d.py:163-164, occ_annotate.py:490-491)
* ``cosf/sinf(yaw + pi/2)`` per frame     (points_in_boxes_cpu.cpp:19-20)
* ``torch.linalg.inv(extrinsic)`` f32 and ``torch.atan2(E[1,0], E[0,0])`` per (frame, LiDAR)
  (occ_annotate.py:158-160, 175)
* the inclination flip                     (occ_annotate.py:528)

Everything per point / per voxel runs in the CUDA kernels (csrc/annotate.cu).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib

LiDAR_NAME_LIST = ["TOP", "FRONT", "SIDE_LEFT", "SIDE_RIGHT", "REAR"]     # occ_annotate.py:235
STATUS_NAMES = {0: "ok", 1: "skip_short", 2: "no_points", 3: "empty_after_filter", 4: "index_error",
                5: "work_overflow", -1: "slot_too_small"}
FLAG_FORCE_F64 = 1
FLAG_NO_CULL = 2
FLAG_TINY_QUEUE = 8        # tests: forces the recheck-queue overflow path
FLAG_NO_BRICK_CULL = 16    # A/B: pair-level culling only
FLAG_CUDA_ARITH = 32       # scalar divisions as torch-CUDA evaluates them (x * (1/vs)); default: torch-CPU's x / vs


# ------------------------------------------------------------------------------------------------
# packing (host)
# ------------------------------------------------------------------------------------------------
def _host_trig(rz: np.ndarray) -> np.ndarray:
    r = torch.from_numpy(np.ascontiguousarray(rz, np.float32))
    m = -r
    return torch.stack([torch.cos(m), torch.sin(m), torch.cos(r), torch.sin(r)], -1).numpy()


def _host_calib(extrinsics: np.ndarray):
    E = torch.from_numpy(np.ascontiguousarray(extrinsics, np.float32)).reshape(-1, 4, 4)
    inv = torch.linalg.inv(E)                       # same call as occ_annotate.py:158-160 (on CPU, f32)
    azc = torch.atan2(E[:, 1, 0], E[:, 0, 0])       # occ_annotate.py:175
    return inv[:, :3, :].reshape(-1, 12).numpy(), azc.numpy()


def _mono(table: np.ndarray) -> int:
    d = np.diff(table.astype(np.float64))
    if d.size == 0 or (d >= 0).all():
        return 1
    if (d <= 0).all():
        return -1
    return 0


@dataclass
class PackedTracklets:
    """Host arrays in the layout of ``occb200_annotate_args_t`` (include/occ_b200.h).

    The two large inputs are NOT concatenated on the host: ``ri_parts`` / ``pt_parts`` list the source arrays
    (range images of one LiDAR of one segment; candidate points of one tracklet) with their offsets in the
    device pools, and ``upload`` copies each straight into its slot.  ``ri_pool`` / ``points`` materialise
    the concatenation on demand (tests, fixtures)."""

    T: int
    L: int
    F: int
    voxel_size: float
    trk_frame_off: np.ndarray      # i64 [T+1]
    poses: np.ndarray              # POSE_DTYPE [F]
    frame_sf: np.ndarray           # i32 [F]
    frame_trk: np.ndarray          # i32 [F]
    frame_pt_off: np.ndarray       # i64 [F+1]
    sensors: np.ndarray            # SENSOR_DTYPE [SF, L]
    incl_pool: np.ndarray          # f32
    label_off: np.ndarray          # i64 [T+1]
    brick_off: np.ndarray          # i64 [T+1]
    pyr_off: np.ndarray            # i64 [SF*L+1]
    table_off: np.ndarray          # i64 [n_tables]
    table_H: np.ndarray            # i32 [n_tables]
    max_pairs: int
    point_stride: int
    n_points: int
    ri_len: int
    pt_parts: list                 # [(row offset, f32 [n, stride] array)]
    ri_parts: list                 # [(float offset, f32 [nb, H, W] array)]
    trk_smax: Optional[np.ndarray] = None    # f32 [T,3] max box size over all frames (the grid's upper bound)

    @property
    def total_slots(self) -> int:
        return int(self.label_off[-1])

    @property
    def points(self) -> np.ndarray:
        out = np.zeros((self.n_points, self.point_stride), np.float32)
        for off, a in self.pt_parts:
            out[off:off + len(a)] = a
        return out

    @property
    def ri_pool(self) -> np.ndarray:
        out = np.zeros(self.ri_len, np.float32)
        for off, a in self.ri_parts:
            out[off:off + a.size] = a.reshape(-1)
        return out

    def input_bytes(self) -> int:
        return (sum(getattr(self, n).nbytes for n in _SMALL_FIELDS) + 4 * self.ri_len
                + 4 * self.n_points * self.point_stride)


_SMALL_FIELDS = ("trk_frame_off", "poses", "frame_sf", "frame_trk", "frame_pt_off", "sensors", "incl_pool",
                 "label_off", "brick_off", "pyr_off", "table_off", "table_H", "trk_smax")


def _same_memory(flat, pieces, n_rows) -> bool:
    """``flat`` is the concatenation of ``pieces`` (views into it, in order): same start, same end, same rows."""
    if flat is None or flat.ndim != 2 or not flat.flags.c_contiguous or flat.dtype != np.float32 or not pieces:
        return False
    first, last = pieces[0], pieces[-1]
    if n_rows != len(flat) or first.dtype != np.float32 or first.ndim != 2 or first.shape[1] != flat.shape[1]:
        return False
    a0 = flat.__array_interface__["data"][0]
    return (first.__array_interface__["data"][0] == a0 and last.flags.c_contiguous
            and last.__array_interface__["data"][0] + last.nbytes == a0 + flat.nbytes)


RI_BLOCK = 16      # floats per block of the windowed range-image upload (csrc/ri_windows.cu)


def window_mask(pk: "PackedTracklets", sub_edge: float = 1e9) -> np.ndarray:
    """u8 per 8-float block of ``ri_pool``: 1 where a visibility test of the batch can read
    (``occb200_host_window_mark``: corner-bracketed footprints; ``sub_edge`` <= 0: the device path's 0.8 m sub-boxes)."""
    n8 = (pk.ri_len + 7) // 8
    mask = np.zeros(max(n8, 1), np.uint8)
    if pk.T and pk.F and n8:
        rc = _lib.lib().occb200_host_window_mark(
            pk.T, pk.L, pk.trk_frame_off.ctypes.data, pk.poses.ctypes.data, pk.frame_sf.ctypes.data,
            pk.sensors.ctypes.data, pk.sensors.shape[0], pk.incl_pool.ctypes.data, pk.trk_smax.ctypes.data,
            float(pk.voxel_size), pk.ri_len, mask.ctypes.data, float(sub_edge))
        _lib.check(rc, "occb200_host_window_mark")
    return mask[:n8]


def window_blocks(pk: "PackedTracklets") -> np.ndarray:
    """The 16-float blocks of ``ri_pool`` the "host" upload mode copies (ascending u32)."""
    mask = window_mask(pk)
    out = np.empty((mask.size + 1) // 2 + 1, np.uint32)
    n = _lib.lib().occb200_host_mask_to_blocks(mask.ctypes.data, mask.size, out.ctypes.data)
    return out[:n].copy()


def pack_tracklets(batch, pack_override: Optional[dict] = None) -> PackedTracklets:
    """Pack a batch (``segments`` + ``tracklets``; see synth.TrackletBatch) for the device.

    ``pack_override`` may carry host-derived values (``trig`` [F,4], ``pib`` [F,2], ``v2l``, ``azc``)
    recorded with a golden fixture so a different host libm/torch cannot perturb a parity check.
    """
    segs, trks = batch.segments, batch.tracklets
    L = len(segs[0].inclinations) if segs else len(LiDAR_NAME_LIST)
    T = len(trks)
    sf_base = np.cumsum([0] + [s.num_frames for s in segs]).astype(np.int64)
    SF = int(sf_base[-1])
    sensors = np.zeros((SF, L), _lib.SENSOR_DTYPE)
    incl_parts, ri_parts, table_off, table_H = [], [], [], []
    incl_off = ri_off = 0
    ext = np.concatenate([np.asarray(s.extrinsics, np.float32).reshape(s.num_frames, L, 4, 4) for s in segs], 0) \
        if segs else np.zeros((0, L, 4, 4), np.float32)
    v2l_all, azc_all = _host_calib(ext) if SF else (np.zeros((0, 12), np.float32), np.zeros(0, np.float32))
    sensors["v2l"] = v2l_all.reshape(SF, L, 12)
    sensors["azc"] = azc_all.reshape(SF, L)
    for si, s in enumerate(segs):
        sl = slice(int(sf_base[si]), int(sf_base[si + 1]))
        for c in range(L):
            table = np.ascontiguousarray(s.inclinations[c][::-1], np.float32)      # flip (:528)
            img = s.range_images[c]
            if img.dtype != np.float32 or not img.flags.c_contiguous:
                img = np.ascontiguousarray(img, np.float32)
            nb, H, W = img.shape
            assert table.shape[0] == H and nb == s.num_frames
            sensors["incl_off"][sl, c] = incl_off
            sensors["incl_mono"][sl, c] = _mono(table)
            sensors["ri_off"][sl, c] = ri_off + np.arange(nb, dtype=np.int64) * H * W
            sensors["H"][sl, c] = H
            sensors["W"][sl, c] = W
            incl_parts.append(table)
            table_off.append(incl_off)
            table_H.append(H)
            ri_parts.append((ri_off, img))
            incl_off += H
            ri_off += -(-(nb * H * W) // RI_BLOCK) * RI_BLOCK      # parts start on block boundaries
    L_ = _lib.lib()
    sn = sensors.reshape(-1)
    tiles = np.zeros(sn.size + 1, np.int64)
    if sn.size:
        hw, inv = np.unique(np.stack([sn["H"], sn["W"]], 1), axis=0, return_inverse=True)
        per = np.array([L_.occb200_pyramid_tiles(int(h), int(w_)) for h, w_ in hw], np.int64)
        np.cumsum(per[inv.reshape(-1)], out=tiles[1:])
    nfr = np.array([len(t) for t in trks], np.int64)
    trk_frame_off = np.concatenate([[0], np.cumsum(nfr)]).astype(np.int64)
    F = int(trk_frame_off[-1])
    boxes = np.concatenate([t.boxes for t in trks], 0).astype(np.float32) if F else np.zeros((0, 7), np.float32)
    frame_sf = (np.concatenate([sf_base[t.segment] + np.asarray(t.frame_ids, np.int64) for t in trks]).astype(np.int32)
                if F else np.zeros(0, np.int32))
    frame_trk = np.repeat(np.arange(T, dtype=np.int32), nfr)
    # candidate points: one part per tracklet when its frames are views of one contiguous array, else per frame
    pt_parts, lens, stride, row = [], [], None, 0
    for t in trks:
        n_t = [len(p_) for p_ in t.points]
        lens.extend(n_t)
        if len(t.points) and stride is None:
            stride = int(t.points[0].shape[1])
        if getattr(t, "flat", None) is not None and _same_memory(t.flat, t.points, sum(n_t)):
            if len(t.flat):
                pt_parts.append((row, t.flat))
            row += len(t.flat)
        else:
            for p_ in t.points:
                if len(p_):
                    pt_parts.append((row, np.ascontiguousarray(p_, np.float32)))
                row += len(p_)
    stride = stride or 3
    frame_pt_off = np.concatenate([[0], np.cumsum(np.asarray(lens, np.int64))]).astype(np.int64)
    trig = _host_trig(boxes[:, 6]) if F else np.zeros((0, 4), np.float32)
    if pack_override and "trig" in pack_override:
        trig = np.ascontiguousarray(pack_override["trig"], np.float32)
    poses = np.zeros(F, _lib.POSE_DTYPE)
    if F:
        boxes = np.ascontiguousarray(boxes)
        trig = np.ascontiguousarray(trig, np.float32)
        L_.occb200_host_pose_pack(boxes.ctypes.data, trig.ctypes.data, F, poses.ctypes.data)
    if pack_override:
        if "pib" in pack_override:
            poses["cos_pib"] = pack_override["pib"][:, 0]
            poses["sin_pib"] = pack_override["pib"][:, 1]
        if "v2l" in pack_override:
            sensors["v2l"] = np.asarray(pack_override["v2l"], np.float32).reshape(SF, L, 12)
        if "azc" in pack_override:
            sensors["azc"] = np.asarray(pack_override["azc"], np.float32).reshape(SF, L)
    # label slots: upper bound of the grid = ceil(max over ALL frames of the box size / vs) in f32; the reciprocal
    # form (torch-CUDA arithmetic, flag bit 5) may round one voxel higher, so the slot takes the larger of the two
    vsf = np.float32(batch.voxel_size)
    caps = np.zeros((T, 3), np.int64)
    smax = np.zeros((T, 3), np.float32)
    if F:
        nz = nfr > 0
        mx = np.maximum.reduceat(boxes[:, 3:6], trk_frame_off[:-1][nz], axis=0)
        smax[nz] = mx
        d = np.maximum(np.ceil(mx / vsf), np.ceil(mx * (np.float32(1.0) / vsf))).astype(np.int64)
        caps[nz] = np.maximum(d, 0)
    label_off = np.concatenate([[0], np.cumsum(caps.prod(1))]).astype(np.int64)
    brick_off = np.concatenate([[0], np.cumsum(((caps + 3) // 4).prod(1))]).astype(np.int64)
    return PackedTracklets(
        T=T, L=L, F=F, voxel_size=float(batch.voxel_size), trk_frame_off=trk_frame_off, poses=poses,
        frame_sf=frame_sf, frame_trk=frame_trk, frame_pt_off=frame_pt_off, sensors=sensors,
        incl_pool=np.concatenate(incl_parts) if incl_parts else np.zeros(0, np.float32), label_off=label_off,
        brick_off=brick_off, pyr_off=tiles, table_off=np.asarray(table_off, np.int64),
        table_H=np.asarray(table_H, np.int32), max_pairs=int(nfr.max()) * L if T else 0, point_stride=stride,
        n_points=int(frame_pt_off[-1]), ri_len=int(ri_off), pt_parts=pt_parts, ri_parts=ri_parts, trk_smax=smax)


# ------------------------------------------------------------------------------------------------
# device side
# ------------------------------------------------------------------------------------------------
def _as_bytes(a: np.ndarray) -> torch.Tensor:
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a.view(np.uint8).reshape(-1)) if a.size else torch.zeros(0, dtype=torch.uint8)


_ARENA: dict = {}


def _pinned_arena(nbytes: int) -> torch.Tensor:
    """One process-wide pinned staging buffer for the one-shot API, grown geometrically and reused across calls (pinning
    memory costs milliseconds; a fresh pageable staging buffer costs its page faults on every call)."""
    buf = _ARENA.get("buf")
    if buf is None or buf.numel() < nbytes:
        _ARENA["buf"] = buf = torch.empty(int(max(nbytes, 1) * 1.25) + 4096, dtype=torch.uint8, pin_memory=True)
    return buf


class HostBuffers:
    """Host side of the H2D copies of a PackedTracklets.

    The small fields travel as ONE buffer.  The range images -- 104 MB per segment, of which the visibility test
    reads a small window per (tracklet-frame, LiDAR) -- go up in one of three ways (``ri_mode``):

    ``"pull"``  (``windows`` and ``pin``) the images sit in ONE pinned pool in the device layout (where a loader that
                reads into pinned memory would put them); the device marks the blocks the batch can read and pulls
                exactly those over PCIe (``occb200_pull_windows``).  No host work per step.
    ``"host"``  (``windows``, pageable memory) the host derives the windows (``occb200_host_ri_window_blocks``),
                gathers the blocks from the source arrays into a staging buffer, and the device scatters them.
    ``"whole"`` every image is copied whole, as the reference loads them (occ_annotate.py:502-533).

    ``pin=True`` also puts the candidate points into one pinned buffer; ``pin=False`` uploads everything from where it
    lies; ``pin="arena"`` (the one-shot ``annotate_batch``) stages small fields, points, window blocks and block list
    in the process-wide pinned arena with parallel host copies (``occb200_host_copy_parts``): valid until the next
    arena user, i.e. for one synchronous call."""

    def __init__(self, pk: PackedTracklets, pin=True, windows=True):
        self.pk = pk
        self.arena = pin == "arena"
        if self.arena:
            assert windows is True or windows == "host", "the arena stages the host-gathered windows"
            pin = False
        self.pin = pin
        if windows in ("pull", "host", "whole"):
            assert windows != "pull" or pin, "the device can only pull from pinned memory"
            self.ri_mode = windows
        else:
            self.ri_mode = ("pull" if pin else "host") if windows else "whole"
        if not (pk.ri_len and pk.F):
            self.ri_mode = "whole"

        def host(a):
            t = _as_bytes(a)
            return t.pin_memory() if (pin and t.numel()) else t

        # small fields: one buffer, 256-byte aligned slots (one copy instead of a dozen)
        self.small_off, off = _small_layout(pk)
        small = torch.zeros(off, dtype=torch.uint8)
        for name in _SMALL_FIELDS:
            src = _as_bytes(getattr(pk, name))
            small[self.small_off[name]: self.small_off[name] + src.numel()] = src
        self.small = small.pin_memory() if pin else small
        if pin and pk.n_points:
            pool = torch.empty(4 * pk.n_points * pk.point_stride, dtype=torch.uint8).pin_memory()
            for o, a in pk.pt_parts:
                b = _as_bytes(a)
                pool[o * pk.point_stride * 4: o * pk.point_stride * 4 + b.numel()] = b
            self.pt_parts = [(0, pool)]
        else:
            self.pt_parts = [(o * pk.point_stride * 4, _as_bytes(a)) for o, a in pk.pt_parts]
        self.ri_parts, self.ri_staging, self.ri_idx, self.ri_blocks, self.ri_pinned = [], None, None, None, None
        if self.ri_mode == "pull":
            pool = torch.zeros(pk.ri_len, dtype=torch.float32).pin_memory()
            for o, a in pk.ri_parts:
                pool[o: o + a.size] = torch.from_numpy(np.ascontiguousarray(a, np.float32).reshape(-1))
            self.ri_pinned = pool
        elif self.ri_mode == "host":
            self.ri_blocks = window_blocks(pk)
            n = int(self.ri_blocks.size)
            self.ri_idx = host(self.ri_blocks)
            self.ri_staging = torch.empty(max(n, 1) * RI_BLOCK, dtype=torch.float32)
            if pin:
                self.ri_staging = self.ri_staging.pin_memory()
            srcs = [np.ascontiguousarray(a, np.float32) for _, a in pk.ri_parts]
            self._src = srcs                                       # keeps the source arrays alive
            self._part_off = np.asarray([o for o, _ in pk.ri_parts], np.int64)
            self._part_len = np.asarray([a.size for a in srcs], np.int64)
            self._part_ptr = np.asarray([a.ctypes.data for a in srcs], np.uint64)
            if self.arena:
                self._stage_in_arena()
            self.gather_windows()
        else:
            self.ri_parts = [(o * 4, host(a)) for o, a in pk.ri_parts]
            if self.arena:
                self._stage_in_arena()

    def _stage_in_arena(self):
        """Move small fields, points, block list and the window staging area into the pinned arena (one buffer)."""
        pk = self.pk

        def up(n):
            return -(-n // 256) * 256

        n_small, n_pts = self.small.numel(), 4 * pk.n_points * pk.point_stride
        n_idx = self.ri_idx.numel() if self.ri_idx is not None else 0
        n_stage = 4 * self.ri_staging.numel() if self.ri_staging is not None else 0
        o_small, o_pts = 0, up(n_small)
        o_idx = o_pts + up(n_pts)
        o_stage = o_idx + up(n_idx)
        buf = _pinned_arena(o_stage + up(n_stage))
        srcs, sizes, offs, keep = [], [], [], []
        for off, t in [(o_small, self.small)] + [(o_pts + o, t) for o, t in self.pt_parts] + \
                ([(o_idx, self.ri_idx)] if n_idx else []):
            if t.numel():
                keep.append(t)
                srcs.append(t.data_ptr()); sizes.append(t.numel()); offs.append(off)
        if srcs:
            a_src = np.asarray(srcs, np.uint64)
            a_sz, a_off = np.asarray(sizes, np.int64), np.asarray(offs, np.int64)
            rc = _lib.lib().occb200_host_copy_parts(a_src.ctypes.data, a_sz.ctypes.data, a_off.ctypes.data, len(srcs),
                                                    buf.data_ptr())
            _lib.check(rc, "occb200_host_copy_parts")
        self.small = buf[o_small: o_small + n_small]
        self.pt_parts = [(0, buf[o_pts: o_pts + n_pts])] if n_pts else []
        if n_idx:
            self.ri_idx = buf[o_idx: o_idx + n_idx]
        if n_stage:
            self.ri_staging = buf[o_stage: o_stage + n_stage].view(torch.float32)

    def gather_windows(self):
        """"host" mode: copy the window blocks from the source range images into the staging buffer (OpenMP)."""
        n = int(self.ri_blocks.size)
        rc = _lib.lib().occb200_host_gather_blocks(self.ri_blocks.ctypes.data, n, self._part_off.ctypes.data,
                                                   self._part_len.ctypes.data, self._part_ptr.ctypes.data,
                                                   len(self._src), self.ri_staging.data_ptr())
        _lib.check(rc, "occb200_host_gather_blocks")

    def nbytes(self) -> int:
        """Bytes one ``DeviceTracklets.upload`` copies with cudaMemcpy ("pull" mode: plus what the device pulls,
        ``DeviceTracklets.pulled_bytes()``)."""
        ri = sum(t.numel() for _, t in self.ri_parts)
        if self.ri_mode == "host":
            ri = 4 * RI_BLOCK * int(self.ri_blocks.size) + self.ri_idx.numel()
        return self.small.numel() + sum(t.numel() for _, t in self.pt_parts) + ri


def _small_layout(pk: PackedTracklets):
    offs, off = {}, 0
    for name in _SMALL_FIELDS:
        offs[name] = off
        off += -(-max(getattr(pk, name).nbytes, 16) // 256) * 256
    return offs, off


class DeviceTracklets:
    """Device-resident inputs + outputs + workspace of one batch; reusable across calls.

    ``labels="i32"`` (default) allocates the reference's int32 labels, ``"u8"`` one byte per voxel (what a job
    keeps on the device and gathers; the int32 of occ_annotate.py:581 is produced at the file boundary),
    ``"both"`` both.  ``labels_u8`` may be a caller-owned uint8 tensor slice (a rank's common label buffer)."""

    def __init__(self, pk: PackedTracklets, device=None, labels: str = "i32", labels_u8: Optional[torch.Tensor] = None):
        _lib.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.pk = pk
        dev = self.device
        self.small_off, off = _small_layout(pk)       # the same slots as HostBuffers.small
        self.small = torch.empty(off, dtype=torch.uint8, device=dev)
        self.bufs = {name: self.small[self.small_off[name]:] for name in _SMALL_FIELDS}
        self.bufs["points"] = torch.empty(max(4 * pk.n_points * pk.point_stride, 16), dtype=torch.uint8, device=dev)
        # a windowed upload fills only the blocks the batch can read: everything else reads as "no return"
        self.bufs["ri_pool"] = torch.zeros(max(4 * pk.ri_len, 16), dtype=torch.uint8, device=dev)
        self.pulled = torch.zeros(1, dtype=torch.int64, device=dev)      # 32-byte blocks pulled so far ("pull" mode)
        T, total = pk.T, pk.total_slots
        self.labels = torch.zeros(max(total, 1), dtype=torch.int32, device=dev) if labels in ("i32", "both") else None
        if labels_u8 is not None:
            assert labels_u8.dtype == torch.uint8 and labels_u8.numel() >= total and labels_u8.is_cuda
            self.labels_u8 = labels_u8
        else:
            self.labels_u8 = (torch.zeros(max(total, 1), dtype=torch.uint8, device=dev)
                              if labels in ("u8", "both") else None)
        self.dims = torch.zeros((max(T, 1), 3), dtype=torch.int32, device=dev)
        self.sizes = torch.zeros((max(T, 1), 3), dtype=torch.float32, device=dev)
        self.status = torch.zeros(max(T, 1), dtype=torch.int32, device=dev)
        self.n_unknown = torch.zeros(max(T, 1), dtype=torch.int64, device=dev)
        self.n_steps = torch.zeros(max(T, 1), dtype=torch.int64, device=dev)
        self.pyr_tiles = int(pk.pyr_off[-1])
        ws = _lib.lib().occb200_annotate_workspace_bytes(T, pk.F, total, pk.sensors.shape[0], pk.L, pk.incl_pool.size,
                                                         self.pyr_tiles, int(pk.brick_off[-1]), pk.max_pairs)
        self.workspace = torch.empty(max(ws, 16), dtype=torch.uint8, device=dev)
        # pull mode: the device also marks the pyramid tiles the batch can read (OCCB200_NO_TILE_LIVE=1: A/B knob)
        self.use_tile_live = os.environ.get("OCCB200_NO_TILE_LIVE", "0") != "1"
        self.tile_live = None          # set by a pulled upload, consumed by args()

    def upload(self, host: HostBuffers):
        """Asynchronous H2D of every input on the current stream; returns the bytes copied."""
        self.small.copy_(host.small, non_blocking=True)
        n = host.small.numel()
        for dst, parts in ((self.bufs["points"], host.pt_parts), (self.bufs["ri_pool"], host.ri_parts)):
            for off, src in parts:
                dst[off: off + src.numel()].copy_(src, non_blocking=True)
                n += src.numel()
        if host.ri_mode == "host":
            n += self._upload_blocks(host.ri_staging, host.ri_idx, int(host.ri_blocks.size))
        elif host.ri_mode == "pull":
            self._pull(host.ri_pinned)
        if host.ri_mode != "pull" and self.tile_live is not None:
            self.tile_live = None          # the flags of an earlier pull say nothing about this upload,
            self._graphs = {}              # and a captured graph still points at them: capture again
        return n

    def _upload_blocks(self, staging: torch.Tensor, idx: torch.Tensor, nblk: int) -> int:
        """"host" mode: H2D of the gathered blocks and their indices + the scatter into the dense pool."""
        if not nblk:
            return 0
        if "ri_blocks" not in self.bufs or self.bufs["ri_idx"].numel() < 4 * nblk:
            self.bufs["ri_blocks"] = torch.empty(4 * RI_BLOCK * nblk, dtype=torch.uint8, device=self.device)
            self.bufs["ri_idx"] = torch.empty(4 * nblk, dtype=torch.uint8, device=self.device)
        sb = staging.view(torch.uint8)[: 4 * RI_BLOCK * nblk]
        self.bufs["ri_blocks"][: sb.numel()].copy_(sb, non_blocking=True)
        self.bufs["ri_idx"][: idx.numel()].copy_(idx, non_blocking=True)
        with torch.cuda.device(self.device):
            rc = _lib.lib().occb200_scatter_blocks(self.bufs["ri_blocks"].data_ptr(), self.bufs["ri_idx"].data_ptr(),
                                                   nblk, self.bufs["ri_pool"].data_ptr(), self.pk.ri_len,
                                                   _lib.stream_ptr(self.device))
        _lib.check(rc, "occb200_scatter_blocks")
        return sb.numel() + idx.numel()

    def _pull(self, ri_pinned: torch.Tensor):
        """"pull" mode: the device marks the blocks it can read and fetches them from the pinned host pool."""
        L_ = _lib.lib()
        if "ri_mask" not in self.bufs:
            self.bufs["ri_mask"] = torch.empty(4 * L_.occb200_window_mask_words(self.pk.ri_len), dtype=torch.uint8,
                                               device=self.device)
        assert ri_pinned.is_pinned() and ri_pinned.numel() == self.pk.ri_len
        if self.pyr_tiles and "ri_tile_live" not in self.bufs:
            self.bufs["ri_tile_live"] = torch.zeros(self.pyr_tiles, dtype=torch.uint8, device=self.device)
        live = self.bufs.get("ri_tile_live") if self.use_tile_live else None
        a = self.args(0)
        with torch.cuda.device(self.device):
            rc = L_.occb200_pull_windows(C.byref(a), self.bufs["trk_smax"].data_ptr(), ri_pinned.data_ptr(),
                                         self.bufs["ri_pool"].data_ptr(), self.pk.ri_len,
                                         self.bufs["ri_mask"].data_ptr(), self.pulled.data_ptr(),
                                         live.data_ptr() if live is not None else None,
                                         _lib.stream_ptr(self.device))
        _lib.check(rc, "occb200_pull_windows")
        # the pyramid of later calls on this batch skips the tiles no test can read (args.ri_tile_live)
        self.tile_live = live

    def pulled_bytes(self) -> int:
        """Bytes the device has pulled from pinned host memory so far (synchronises)."""
        return 32 * int(self.pulled.item())

    def window_mask(self) -> np.ndarray:
        """bool per 32-byte block of ri_pool: the blocks the last ``_pull`` marked (tests; synchronises)."""
        m = self.bufs["ri_mask"].cpu().numpy().view(np.uint32)
        return np.unpackbits(m.view(np.uint8), bitorder="little").astype(bool)[: self.pk.ri_len // 8]

    def args(self, flags: int = 0) -> _lib.AnnotateArgs:
        pk, b = self.pk, self.bufs
        a = _lib.AnnotateArgs()
        a.T, a.L, a.F = pk.T, pk.L, pk.F
        a.trk_frame_off = b["trk_frame_off"].data_ptr()
        a.poses = b["poses"].data_ptr()
        a.frame_sf = b["frame_sf"].data_ptr()
        a.points = b["points"].data_ptr()
        a.point_stride = pk.point_stride
        a.frame_pt_off = b["frame_pt_off"].data_ptr()
        a.sensors = b["sensors"].data_ptr()
        a.SF = pk.sensors.shape[0]
        a.incl_pool = b["incl_pool"].data_ptr()
        a.incl_len = pk.incl_pool.size
        a.ri_pool = b["ri_pool"].data_ptr()
        a.pyr_tiles = self.pyr_tiles
        a.items_cap = 0
        a.voxel_size = pk.voxel_size
        a.label_off = b["label_off"].data_ptr()
        a.labels = self.labels.data_ptr() if self.labels is not None else None
        a.labels_u8 = self.labels_u8.data_ptr() if self.labels_u8 is not None else None
        a.frame_trk = b["frame_trk"].data_ptr()
        a.pyr_off = b["pyr_off"].data_ptr()
        a.table_off = b["table_off"].data_ptr()
        a.table_H = b["table_H"].data_ptr()
        a.n_tables = int(pk.table_off.size)
        a.max_pairs = pk.max_pairs
        a.brick_off = b["brick_off"].data_ptr()
        a.bricks = int(pk.brick_off[-1])
        a.n_points = int(pk.n_points)
        a.ri_tile_live = self.tile_live.data_ptr() if self.tile_live is not None else None
        a.dims = self.dims.data_ptr()
        a.sizes = self.sizes.data_ptr()
        a.status = self.status.data_ptr()
        a.n_unknown = self.n_unknown.data_ptr()
        a.n_steps = self.n_steps.data_ptr()
        a.workspace = self.workspace.data_ptr()
        a.workspace_bytes = self.workspace.numel()
        a.flags = flags
        a.max_label_slots = int(np.diff(pk.label_off).max()) if pk.T else 0
        return a

    def run(self, flags: int = 0):
        """Launch the whole annotate pipeline on the current stream (no synchronisation)."""
        a = self.args(flags)
        with torch.cuda.device(self.device):
            rc = _lib.lib().occb200_annotate_batch(C.byref(a), self.pk.total_slots, _lib.stream_ptr(self.device))
        _lib.check(rc, "occb200_annotate_batch")

    def capture(self, flags: int = 0):
        """Record the whole pipeline -- every kernel, the side-stream fork / join included -- into a CUDA graph;
        ``replay(flags)`` then launches it with a single call (one graph launch instead of a dozen kernel launches
        plus event traffic).  The graph is tied to this object's buffers; ``upload`` refreshes their contents."""
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        self.run(flags)                               # the library creates its side stream outside the capture
        torch.cuda.synchronize(self.device)
        n0 = _lib.launch_count()
        g = torch.cuda.CUDAGraph()
        # thread_local: other threads of the process (e.g. NCCL's watchdog) may touch CUDA during the capture
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self.run(flags)
        self._graphs[flags] = (g, _lib.launch_count() - n0)
        return self._graphs[flags][1]

    def replay(self, flags: int = 0) -> int:
        """Launch the captured pipeline on the current stream; returns the number of kernels in the graph."""
        g, n = self._graphs[flags]
        g.replay()
        return n

    def queue_stats(self, flags: int = 0):
        """(tests sent to the exact f64 recheck, queue capacity) of the last run (synchronises)."""
        a = self.args(flags)
        out = (C.c_int64 * 2)()
        with torch.cuda.device(self.device):
            rc = _lib.lib().occb200_annotate_queue_stats(C.byref(a), self.pk.total_slots, C.addressof(out),
                                                         _lib.stream_ptr(self.device))
        _lib.check(rc, "occb200_annotate_queue_stats")
        return int(out[0]), int(out[1])

    def point_voxels(self, flags: int = 0):
        """After ``run``: (loc f32 [N,3], rows int32 [N,4]) for the N candidate points -- box-frame coordinates and
        (tracklet, qx, qy, qz) with the final grids; rows of points the reference drops start with -1."""
        n = max(int(self.pk.n_points), 1)
        loc = torch.empty((n, 3), dtype=torch.float32, device=self.device)
        rows = torch.empty((n, 4), dtype=torch.int32, device=self.device)
        a = self.args(flags)
        with torch.cuda.device(self.device):
            rc = _lib.lib().occb200_annotate_point_voxels(C.byref(a), self.pk.total_slots, loc.data_ptr(),
                                                          rows.data_ptr(), _lib.stream_ptr(self.device))
        _lib.check(rc, "occb200_annotate_point_voxels")
        n = int(self.pk.n_points)
        return loc[:n], rows[:n]

    def mean_var(self, flags: int = 0) -> List[Optional[np.ndarray]]:
        """``--save-mean-var`` (occ_annotate.py:627-645): per tracklet a dense f32 [X,Y,Z,6] grid with the mean of
        the box-frame points of each voxel and the mean of their squared deviations, zeros elsewhere (None where
        the reference writes no file).  ``scatter_v2`` does the reductions, exactly as in the reference; groups
        are formed on the RAW quantised coordinates and a negative one lands on its wrapped cell (:641), the
        last group in sorted order winning where two land on the same cell."""
        from .sst_ops import scatter_v2

        pk = self.pk
        dims = self.dims.cpu().numpy()
        status = self.status.cpu().numpy()
        loc, rows = self.point_voxels(flags)
        keep = rows[:, 0] >= 0
        loc, rows = loc[keep].contiguous(), rows[keep].long().contiguous()
        dense = torch.zeros((max(pk.total_slots, 1), 6), dtype=torch.float32, device=self.device)
        if rows.shape[0]:
            mean, new_coors, inv = scatter_v2(loc, rows, "mean", return_inv=True)
            var = ((loc - mean[inv]) ** 2).contiguous()
            var_m, _, _ = scatter_v2(var, rows, "mean", unq_inv=inv, new_coors=new_coors)
            d = self.dims.long()[new_coors[:, 0]]                       # [M,3] dims of each group's tracklet
            q = new_coors[:, 1:]
            q = torch.where(q < 0, q + d, q)                            # PyTorch negative-index wrap (:641)
            off = torch.from_numpy(pk.label_off[:-1]).to(self.device)[new_coors[:, 0]]
            cell = off + (q[:, 0] * d[:, 1] + q[:, 1]) * d[:, 2] + q[:, 2]
            # duplicates (a wrapped and an unwrapped group on one cell): the later row of the sorted list wins
            last = torch.full((dense.shape[0],), -1, dtype=torch.long, device=self.device)
            last.scatter_reduce_(0, cell, torch.arange(cell.shape[0], device=self.device), "amax")
            sel = last[cell] == torch.arange(cell.shape[0], device=self.device)
            dense[cell[sel]] = torch.cat([mean, var_m], 1)[sel]
        dense = dense.cpu().numpy()
        out = []
        for t in range(pk.T):
            if int(status[t]) != 0:
                out.append(None)
                continue
            X, Y, Z = (int(v) for v in dims[t])
            o = int(pk.label_off[t])
            out.append(dense[o:o + X * Y * Z].reshape(X, Y, Z, 6).copy())
        return out

    def results(self) -> List[dict]:
        """D2H of labels / dims / status and per-tracklet reshape (synchronises)."""
        pk = self.pk
        labels = (self.labels if self.labels is not None else self.labels_u8).cpu().numpy()
        dims = self.dims.cpu().numpy()
        sizes = self.sizes.cpu().numpy()
        status = self.status.cpu().numpy()
        nunk = self.n_unknown.cpu().numpy()
        nsteps = self.n_steps.cpu().numpy()
        out = []
        for t in range(pk.T):
            st = int(status[t])
            if st != 0:
                out.append(dict(status=STATUS_NAMES.get(st, str(st)), occ=None, dims=dims[t].copy(),
                                size=sizes[t].copy(), n_unknown=0, n_steps=0))
                continue
            X, Y, Z = (int(v) for v in dims[t])
            o = int(pk.label_off[t])
            out.append(dict(status="ok", occ=labels[o:o + X * Y * Z].reshape(X, Y, Z).astype(np.int32), dims=dims[t].copy(),
                            size=sizes[t].copy(), n_unknown=int(nunk[t]), n_steps=int(nsteps[t])))
        return out


def annotate_batch(batch, flags: int = 0, pack_override: Optional[dict] = None, device=None,
                   save_mean_var: bool = False, windows: bool = True) -> List[dict]:
    """Annotate every tracklet of ``batch``: the batched equivalent of ``OccAnnotator.annotate_trk``.

    Returns one dict per tracklet: ``status`` (``ok`` or the reason the reference produces no file),
    ``occ`` int32 [X,Y,Z] with 0 unknown / 1 occupied / 2 free (occ_annotate.py:558-563, 581), ``dims``,
    ``size``, ``n_unknown`` (U) and ``n_steps`` (visibility tests evaluated).
    """
    pk = pack_tracklets(batch, pack_override)
    # one-shot call: the pageable inputs are staged in the process-wide pinned arena (windows gathered by the host)
    host = HostBuffers(pk, pin="arena" if windows is True and pk.ri_len and pk.F else False, windows=windows)
    dev = DeviceTracklets(pk, device)
    dev.upload(host)
    dev.run(flags)
    res = dev.results()
    if save_mean_var:                                        # occ_annotate.py:627-645
        for r, mv in zip(res, dev.mean_var(flags)):
            r["mean_var"] = mv
    return res


def point_cloud_to_range_image_idx(points, extrinsics, inclinations, range_image_size):
    """Drop-in for occ_annotate.py:141-201 on CUDA tensors.

    points f64 [B,N,3]; extrinsics f32 [B,4,4]; inclinations f32 [B,H] (already flipped);
    returns (ri_indices int64 [B,N,2], ri_range f64 [B,N]).  As in the reference the extrinsic
    inverse and the azimuth correction are evaluated on the CPU (occ_annotate.py:158-160).
    """
    _lib.require_cuda(points)
    H, W = range_image_size
    B, N, _ = points.shape
    v2l, azc = _host_calib(extrinsics.detach().cpu().numpy())
    dev = points.device
    v2l_d = torch.from_numpy(np.ascontiguousarray(v2l)).to(dev)
    azc_d = torch.from_numpy(np.ascontiguousarray(azc)).to(dev)
    pts = points.to(torch.float64).contiguous()
    incl = inclinations.to(device=dev, dtype=torch.float32).contiguous()
    assert incl.shape == (B, H)
    idx = torch.empty((B, N, 2), dtype=torch.int64, device=dev)
    rng = torch.empty((B, N), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().occb200_point_cloud_to_range_image_idx(pts.data_ptr(), B, N, v2l_d.data_ptr(), azc_d.data_ptr(),
                                                                incl.data_ptr(), H, W, idx.data_ptr(), rng.data_ptr(),
                                                                _lib.stream_ptr(dev))
    _lib.check(rc, "occb200_point_cloud_to_range_image_idx")
    return idx, rng


class OccAnnotator:
    """Batched counterpart of the reference's ``OccAnnotator`` (occ_annotate.py:228-671).

    Keeps the reference's knobs that affect the result (``voxel_size``, ``overwrite``, the
    ``<out_dir>/<split>/<segment>/<trk_id>.npz`` layout with key ``occ``) and replaces the
    per-tracklet python loop by one device call per batch.
    """

    def __init__(self, out_dir: Optional[str] = None, split: str = "training", voxel_size: float = 0.2,
                 overwrite: bool = False, device=None):
        self.out_dir = out_dir
        self.split = split
        self.voxel_size = voxel_size
        self.overwrite = overwrite
        self.device = device

    def annotate(self, batch) -> List[dict]:
        batch.voxel_size = self.voxel_size
        return annotate_batch(batch, device=self.device)

    def out_name(self, segment_name: str, trk_id: str) -> str:
        return os.path.join(self.out_dir, self.split, segment_name, f"{trk_id}.npz")      # :330-332

    def annotate_and_save(self, batch, names: Sequence[tuple]) -> List[Optional[str]]:
        """``names[t] = (segment_name, trk_id)``.  Existing loadable files are skipped unless
        ``overwrite`` (:335-343); tracklets the reference would not write are skipped as well."""
        res = self.annotate(batch)
        written = []
        for r, (seg, tid) in zip(res, names):
            path = self.out_name(seg, tid)
            if r["occ"] is None:
                written.append(None)
                continue
            if os.path.isfile(path) and not self.overwrite:
                try:
                    np.load(path)
                    written.append(path)
                    continue
                except Exception:
                    pass
            os.makedirs(os.path.dirname(path), exist_ok=True)
            np.savez(path, occ=r["occ"])                                                   # :647
            written.append(path)
        return written
