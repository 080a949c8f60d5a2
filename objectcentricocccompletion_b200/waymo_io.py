"""On-disk formats either side of the annotate path.

Reads what ``tools/occ/occ_annotate.py`` reads and writes what it writes, so the batched CUDA path can run on a
converted Waymo directory and feed ``LoadAnnotationsOcc`` unchanged:

* ``<data_root>/kitti_format/idx2timestamp.pkl``: ``{idx: timestamp}``           (occ_annotate.py:256, 280-282)
* ``<data_root>/kitti_format/<split>/velodyne/<idx>.bin``: f32 ``[N, 6]`` rows   (tools/ctrl/utils.py:60-66)
* ``<data_root>/waymo_raw/<split>/<idx>.pkl``: dict with, per LiDAR name,
  ``<NAME>_BEAM_INCLINATION [H]``, ``<NAME>_LIDAR_EXTRINSIC [4,4]``,
  ``<NAME>_RANGE_IMAGE_MERGE_VIRTUAL [H,W]``                                     (occ_annotate.py:502-519)
* output ``<out_dir>/<split>/<segment>/<trk_id>.npz`` with key ``occ`` int32 ``[X,Y,Z]`` (:330-332, 647)

Tracklets come as plain records (``TrackletRecord``); a reference ``LiDARTracklet`` converts with
``TrackletRecord(t.segment_name, t.id, t.type, torch.cat([b.tensor for b in t.box_list]).numpy(), t.ts_list)``
(unpickling ``LiDARTracklet`` itself needs mmdet3d, which is outside this package).

The hot path is unchanged: this module only builds ``synth.Segment`` / ``synth.Tracklet`` objects from files.
Candidate points of a tracklet-frame are the frame's points inside a sphere around the box (a superset of the
in-box points; the exact in-box test of the reference runs on the device).
"""
from __future__ import annotations

import os
import pickle
from dataclasses import dataclass
from typing import Tuple, Dict, List, Optional, Sequence

import numpy as np

from .occ_annotate import LiDAR_NAME_LIST, annotate_batch
from .synth import Segment, Tracklet, TrackletBatch


@dataclass
class TrackletRecord:
    segment_name: str
    id: str
    type: int
    boxes: np.ndarray            # f32 [B, 7]  (x, y, z_bottom, x_size, y_size, z_size, yaw)
    ts_list: Sequence[int]       # frame timestamps, one per box

    def __len__(self):
        return len(self.ts_list)


def read_velodyne_bin(path: str) -> np.ndarray:
    """KITTI-format Waymo point cloud: f32 rows of 6 (tools/ctrl/utils.py:60-66)."""
    return np.fromfile(path, dtype=np.float32).reshape(-1, 6)


def load_idx2timestamp(kitti_format_root: str) -> Dict[int, str]:
    """ts -> idx, as occ_annotate.py:280-282 builds it."""
    with open(os.path.join(kitti_format_root, "idx2timestamp.pkl"), "rb") as fr:
        idx2ts = pickle.load(fr)
    return {ts: idx for idx, ts in idx2ts.items()}


def load_raw_frame(raw_root: str, idx: str) -> dict:
    with open(os.path.join(raw_root, f"{idx}.pkl"), "rb") as fr:
        return pickle.load(fr)


def out_name(out_dir: str, split: str, segment_name: str, trk_id: str) -> str:
    return os.path.join(out_dir, split, segment_name, f"{trk_id}.npz")        # occ_annotate.py:330-332


def _needs_work(path: str, overwrite: bool) -> bool:
    """occ_annotate.py:335-343: an existing file that loads is kept unless --overwrite."""
    if overwrite or not os.path.isfile(path):
        return True
    try:
        np.load(path)
        return False
    except Exception:
        return True


def build_segment_batch(records: Sequence[TrackletRecord], ts2idx: Dict, data_root: str, split: str,
                        voxel_size: float, candidate_margin: float = 0.5, device_select: Optional[bool] = None,
                        device=None) -> Tuple[Optional[TrackletBatch], List[int]]:
    """All tracklets of ONE segment -> a TrackletBatch sharing that segment's frames (each frame's point cloud and
    range images are read once per segment, like ``cache_segment_pcs``, occ_annotate.py:312).

    A tracklet that touches a timestamp whose raw frame file is missing is dropped -- the reference aborts exactly
    those tracklets (:503-510), not the segment.  Returns (batch or None if nothing is left, indices of the kept
    records).  ``device_select`` (default: when CUDA is available) picks each tracklet-frame's candidate points
    from the whole-frame clouds on the GPU (candidates.select_candidates: every cloud is read once); otherwise a
    numpy sphere test per (tracklet, frame)."""
    kitti_root = os.path.join(data_root, "kitti_format")
    raw_root = os.path.join(data_root, "waymo_raw", split)
    all_ts = sorted({ts for r in records for ts in r.ts_list})
    missing = {ts for ts in all_ts if not os.path.isfile(os.path.join(raw_root, f"{ts2idx[ts]}.pkl"))}
    kept = [i for i, r in enumerate(records) if not (set(r.ts_list) & missing)]
    records = [records[i] for i in kept]
    if not records:
        return None, []
    all_ts = sorted({ts for r in records for ts in r.ts_list})
    frame_of = {ts: i for i, ts in enumerate(all_ts)}
    clouds, extr, incl, ris = [], [], None, None
    for ts in all_ts:
        idx = ts2idx[ts]
        fd = load_raw_frame(raw_root, idx)
        clouds.append(read_velodyne_bin(os.path.join(kitti_root, split, "velodyne", f"{idx}.bin")))
        extr.append(np.stack([np.asarray(fd[f"{n}_LIDAR_EXTRINSIC"], np.float32) for n in LiDAR_NAME_LIST], 0))
        tabs = [np.asarray(fd[f"{n}_BEAM_INCLINATION"], np.float32) for n in LiDAR_NAME_LIST]
        imgs = [np.asarray(fd[f"{n}_RANGE_IMAGE_MERGE_VIRTUAL"], np.float32) for n in LiDAR_NAME_LIST]
        if incl is None:
            incl = tabs
            ris = [[im] for im in imgs]
        else:
            for c in range(len(LiDAR_NAME_LIST)):
                if not np.array_equal(tabs[c], incl[c]):
                    raise ValueError("beam inclinations change inside a segment: split it into per-table segments")
                ris[c].append(imgs[c])
    seg = Segment(extrinsics=np.stack(extr, 0), inclinations=incl, range_images=[np.stack(r, 0) for r in ris])
    boxes = [np.asarray(r.boxes, np.float32).reshape(-1, 7) for r in records]
    frames = [np.array([frame_of[ts] for ts in r.ts_list], np.int32) for r in records]
    if device_select is None:
        import torch

        device_select = torch.cuda.is_available()
    trks = []
    if device_select:
        from .candidates import select_candidates, split_candidates

        pts, cnt = select_candidates(clouds, boxes, frames, margin=candidate_margin, device=device)
        for b, f, (flat, per_frame) in zip(boxes, frames, split_candidates(pts, cnt, [len(b) for b in boxes])):
            trks.append(Tracklet(boxes=b, points=per_frame, segment=0, frame_ids=f, flat=flat))
    else:
        for b7, f in zip(boxes, frames):
            pts = []
            for b, fi in zip(b7, f):
                pc = clouds[int(fi)]
                ctr = b[:3] + np.array([0, 0, 0.5 * b[5]], np.float32)
                rad = np.float32(0.5) * np.linalg.norm(b[3:6]).astype(np.float32) + np.float32(candidate_margin)
                d = pc[:, :3] - ctr
                pts.append(np.ascontiguousarray(pc[(d * d).sum(1) <= rad * rad]))
            trks.append(Tracklet(boxes=b7, points=pts, segment=0, frame_ids=f))
    return TrackletBatch(segments=[seg], tracklets=trks, voxel_size=float(voxel_size)), kept


def save_tracklet_records(path: str, records: Sequence[TrackletRecord]) -> None:
    """Plain-numpy tracklet file (npz): the stand-in for the reference's ``<name>_tracklets.pkl`` cache
    (occ_annotate.py:270-273), whose pickled ``LiDARTracklet`` objects need mmdet3d to load."""
    np.savez(path, n=np.int64(len(records)),
             **{f"seg{i}": np.array(r.segment_name) for i, r in enumerate(records)},
             **{f"id{i}": np.array(r.id) for i, r in enumerate(records)},
             **{f"type{i}": np.int64(r.type) for i, r in enumerate(records)},
             **{f"boxes{i}": np.asarray(r.boxes, np.float32).reshape(-1, 7) for i, r in enumerate(records)},
             **{f"ts{i}": np.asarray(r.ts_list, np.int64) for i, r in enumerate(records)})


def load_tracklet_records(path: str) -> List[TrackletRecord]:
    z = np.load(path, allow_pickle=False)
    return [TrackletRecord(str(z[f"seg{i}"]), str(z[f"id{i}"]), int(z[f"type{i}"]), z[f"boxes{i}"],
                           [int(t) for t in z[f"ts{i}"]]) for i in range(int(z["n"]))]


def segments_of_rank(segment_names: Sequence[str], rank: int, world: int) -> List[str]:
    """The reference deals sorted segments to its worker processes, worker ``wid`` on GPU ``wid % ngpus``
    (occ_annotate.py:278, 320-322, 659-665); here one process per GPU takes every ``world``-th segment."""
    return [s for i, s in enumerate(sorted(set(segment_names))) if i % world == rank]


def annotate_from_disk(records: Sequence[TrackletRecord], data_root: str, out_dir: str, split: str = "training",
                       voxel_size: float = 0.2, overwrite: bool = False, annotate_fn=None, save_mean_var: bool = False,
                       device=None, device_select: Optional[bool] = None) -> List[Optional[str]]:
    """The job of ``OccAnnotator.annotate_segment`` (occ_annotate.py:649-671) on a converted Waymo directory:
    group by segment, skip finished / short tracklets (:335-345), annotate each segment's batch, write npz files
    (``save_mean_var``: the ``--save-mean-var`` grids beside the labels, :627-645).
    Returns the written (or kept) path per record, None where the reference writes nothing."""
    if annotate_fn is None:
        def annotate_fn(batch):
            return annotate_batch(batch, device=device, save_mean_var=save_mean_var)
    ts2idx = load_idx2timestamp(os.path.join(data_root, "kitti_format"))
    result: List[Optional[str]] = [None] * len(records)
    by_seg: Dict[str, List[int]] = {}
    for i, r in enumerate(records):
        by_seg.setdefault(r.segment_name, []).append(i)
    for seg_name in sorted(by_seg):                                            # :278
        todo = []
        for i in by_seg[seg_name]:
            path = out_name(out_dir, split, seg_name, records[i].id)
            if not _needs_work(path, overwrite):
                result[i] = path
            elif len(records[i]) >= 10:                                        # :344
                todo.append(i)
        if not todo:
            continue
        batch, kept = build_segment_batch([records[i] for i in todo], ts2idx, data_root, split, voxel_size,
                                          device_select=device_select, device=device)
        if batch is None:
            continue
        for k, res in zip(kept, annotate_fn(batch)):
            i = todo[k]
            if res["occ"] is None:
                continue
            path = out_name(out_dir, split, seg_name, records[i].id)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            if res.get("mean_var") is not None:                                # --save-mean-var, :643-645
                np.savez(path, occ=res["occ"].astype(np.int32), mean_var=res["mean_var"])
            else:
                np.savez(path, occ=res["occ"].astype(np.int32))                # :647
            result[i] = path
    return result


def write_synthetic_dataset(batch: TrackletBatch, data_root: str, split: str = "training",
                            segment_names: Optional[Sequence[str]] = None) -> List[TrackletRecord]:
    """Test helper: lay a synthetic batch out on disk in the reference's formats and return its records."""
    kitti_root = os.path.join(data_root, "kitti_format")
    raw_root = os.path.join(data_root, "waymo_raw", split)
    os.makedirs(os.path.join(kitti_root, split, "velodyne"), exist_ok=True)
    os.makedirs(raw_root, exist_ok=True)
    idx2ts, records = {}, []
    for si, seg in enumerate(batch.segments):
        name = segment_names[si] if segment_names else f"segment-{si:04d}"
        trks = [t for t in batch.tracklets if t.segment == si]
        for f in range(seg.num_frames):
            idx = f"{si:04d}{f:03d}"
            ts = 1_000_000 * (si + 1) + 100_000 * f
            idx2ts[idx] = ts
            pts = [t.points[int(np.nonzero(t.frame_ids == f)[0][0])] for t in trks if (t.frame_ids == f).any()]
            cloud = np.concatenate(pts, 0) if pts else np.zeros((0, 3), np.float32)
            cloud6 = np.concatenate([cloud[:, :3], np.zeros((len(cloud), 3), np.float32)], 1).astype(np.float32)
            cloud6.tofile(os.path.join(kitti_root, split, "velodyne", f"{idx}.bin"))
            fd = {}
            for c, n in enumerate(LiDAR_NAME_LIST):
                fd[f"{n}_BEAM_INCLINATION"] = seg.inclinations[c]
                fd[f"{n}_LIDAR_EXTRINSIC"] = seg.extrinsics[f, c]
                fd[f"{n}_RANGE_IMAGE_MERGE_VIRTUAL"] = seg.range_images[c][f]
            with open(os.path.join(raw_root, f"{idx}.pkl"), "wb") as fw:
                pickle.dump(fd, fw)
        for k, t in enumerate(trks):
            records.append(TrackletRecord(name, f"obj{k}", 1, t.boxes,
                                          [1_000_000 * (si + 1) + 100_000 * int(f) for f in t.frame_ids]))
    with open(os.path.join(kitti_root, "idx2timestamp.pkl"), "wb") as fw:
        pickle.dump(idx2ts, fw)
    return records
