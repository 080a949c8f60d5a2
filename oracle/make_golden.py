"""Generate tests/golden/*.npz with the REAL reference (runs only where /root/reference exists).

    python -m oracle.make_golden

Fixtures (small enough to commit):
  annotate_small.npz   3 vehicle tracklets x 12 frames, reduced-resolution range images, 0.2 m voxels
  annotate_large.npz   2 truck/bus tracklets x 10 frames, 0.1 m voxels (reduced-resolution images)
  annotate_edge.npz    short tracklet / tracklet without in-box points / frames without points
  projection.npz       point_cloud_to_range_image_idx on random f64 points, 3 LiDAR shapes
Inputs are synthetic (objectcentricocccompletion_b200.synth, seeded); expected outputs come from
oracle/torch_ref.py, i.e. the reference's own point_cloud_to_range_image_idx exec'd from
tools/occ/occ_annotate.py, its compiled points_in_boxes_cpu, and the torch-CPU op sequence of
annotate_trk.  The host-derived pack values (torch sin/cos, f32 inverse, atan2, libm trig) used to make
them are stored too, so a different host libm cannot perturb the check on another machine.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from objectcentricocccompletion_b200 import synth
from . import oracle, torch_ref

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def batch_to_arrays(batch):
    d = dict(voxel_size=np.float64(batch.voxel_size), num_segments=np.int64(len(batch.segments)),
             num_tracklets=np.int64(len(batch.tracklets)))
    for si, s in enumerate(batch.segments):
        d[f"seg{si}_extrinsics"] = s.extrinsics
        for c in range(len(s.inclinations)):
            d[f"seg{si}_incl{c}"] = s.inclinations[c]
            d[f"seg{si}_ri{c}"] = s.range_images[c]
    for ti, t in enumerate(batch.tracklets):
        d[f"trk{ti}_boxes"] = t.boxes
        d[f"trk{ti}_segment"] = np.int64(t.segment)
        d[f"trk{ti}_frame_ids"] = t.frame_ids
        d[f"trk{ti}_pt_off"] = np.cumsum([0] + [len(p) for p in t.points]).astype(np.int64)
        d[f"trk{ti}_points"] = (np.concatenate(t.points, 0) if sum(len(p) for p in t.points) else np.zeros((0, 3), np.float32))
    return d


def arrays_to_batch(d):
    segs = []
    for si in range(int(d["num_segments"])):
        L = d[f"seg{si}_extrinsics"].shape[1]
        segs.append(synth.Segment(extrinsics=d[f"seg{si}_extrinsics"],
                                  inclinations=[d[f"seg{si}_incl{c}"] for c in range(L)],
                                  range_images=[d[f"seg{si}_ri{c}"] for c in range(L)]))
    trks = []
    for ti in range(int(d["num_tracklets"])):
        off = d[f"trk{ti}_pt_off"]
        pts = d[f"trk{ti}_points"]
        trks.append(synth.Tracklet(boxes=d[f"trk{ti}_boxes"], points=[pts[off[i]:off[i + 1]] for i in range(len(off) - 1)],
                                   segment=int(d[f"trk{ti}_segment"]), frame_ids=d[f"trk{ti}_frame_ids"]))
    return synth.TrackletBatch(segments=segs, tracklets=trks, voxel_size=float(d["voxel_size"]))


def add_expected(d, batch):
    ref = torch_ref.annotate_batch(batch)
    pk = oracle.PackedBatch(batch)
    d["pack_trig"] = pk.trig
    d["pack_v2l"] = pk.sensors["v2l"].copy()
    d["pack_azc"] = pk.sensors["azc"].copy()
    from objectcentricocccompletion_b200 import occ_annotate
    pp = occ_annotate.pack_tracklets(batch)
    d["pack_pib"] = np.stack([pp.poses["cos_pib"], pp.poses["sin_pib"]], 1)
    assert (pp.sensors["v2l"] == pk.sensors["v2l"]).all() and (pp.sensors["azc"] == pk.sensors["azc"]).all()
    d["exp_status"] = np.array([r["status"] for r in ref])
    for ti, r in enumerate(ref):
        if r["occ"] is not None:
            d[f"exp_occ{ti}"] = r["occ"].astype(np.int8)
    # the C oracle must agree before anything is written
    orc = oracle.annotate_batch(batch)
    for r, o in zip(ref, orc):
        assert r["status"] == o["status"]
        if r["occ"] is not None:
            assert (r["occ"] == o["occ"]).all()
    return [r["status"] for r in ref]


def edge_batch():
    b = synth.make_batch(5, 12, 0.2, seed=31, small=True)
    t = b.tracklets
    t[0] = synth.Tracklet(boxes=t[0].boxes[:9], points=t[0].points[:9], segment=0, frame_ids=t[0].frame_ids[:9])   # < 10 frames
    far = [p + np.float32(500.0) for p in t[1].points]                                                         # nothing in the box
    t[1] = synth.Tracklet(boxes=t[1].boxes, points=far, segment=0, frame_ids=t[1].frame_ids)
    pts = [p.copy() for p in t[2].points]
    for i in (0, 3, 4, 11):
        pts[i] = np.zeros((0, 3), np.float32)                                                                    # empty frames still used for visibility
    t[2] = synth.Tracklet(boxes=t[2].boxes, points=pts, segment=0, frame_ids=t[2].frame_ids)
    # the largest box of tracklet 3 sits on a frame without points: size must ignore it (box_mode max over KEPT frames)
    bx = t[3].boxes.copy()
    bx[5, 3:6] *= np.float32(1.3)
    pts = [p.copy() for p in t[3].points]
    pts[5] = np.zeros((0, 3), np.float32)
    t[3] = synth.Tracklet(boxes=bx, points=pts, segment=0, frame_ids=t[3].frame_ids)
    return b


def main():
    assert torch_ref.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    for name, batch in [("annotate_small", synth.make_batch(3, 12, 0.2, "vehicle", seed=11, small=True)),
                        ("annotate_large", synth.make_batch(2, 10, 0.1, "large", seed=12, small=True)),
                        ("annotate_edge", edge_batch())]:
        d = batch_to_arrays(batch)
        st = add_expected(d, batch)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, st, f"{os.path.getsize(path) / 1024:.0f} KiB")
    # projection
    fn = torch_ref.reference_projection_fn()
    rng = np.random.default_rng(5)
    d = {}
    for k, (H, W, n) in enumerate([(64, 2650, 3000), (200, 600, 3000), (16, 331, 2000)]):
        rig = synth.lidar_rig(rng)
        E = np.stack([r["extrinsic"] for r in rig], 0)[:3]
        lo, hi = (-17.6, 2.4) if H != 200 else (-90, 30)
        incl = np.sort(np.deg2rad(rng.uniform(lo, hi, (3, H))).astype(np.float32), 1)[:, ::-1].copy()
        pts = rng.uniform(-60, 60, (3, n, 3))
        pts[..., 2] = rng.uniform(-3, 6, (3, n))
        idx, r = fn(torch.from_numpy(pts), torch.from_numpy(E), torch.from_numpy(incl), (H, W))
        v2l, azc = oracle.host_calib(E)
        d.update({f"p{k}_points": pts, f"p{k}_extrinsics": E, f"p{k}_incl": incl, f"p{k}_hw": np.array([H, W]),
                  f"p{k}_v2l": v2l, f"p{k}_azc": azc, f"p{k}_idx": idx.numpy(), f"p{k}_range": r.numpy()})
    path = os.path.join(OUT, "projection.npz")
    np.savez_compressed(path, **d)
    print("projection", f"{os.path.getsize(path) / 1024:.0f} KiB")


def make_mean_var():
    """--save-mean-var grids (occ_annotate.py:627-645) of the committed annotate fixtures, from the torch
    reference glue, into tests/golden/mean_var.npz (keys <fixture>_<tracklet>)."""
    assert torch_ref.available(), "needs /root/reference"
    d = {}
    for name in ("annotate_small", "annotate_edge"):
        batch = arrays_to_batch(dict(np.load(os.path.join(OUT, name + ".npz"), allow_pickle=False)))
        for t, r in enumerate(torch_ref.annotate_batch(batch)):
            if r["occ"] is not None:
                d[f"{name}_{t}"] = torch_ref.mean_var(r)
    path = os.path.join(OUT, "mean_var.npz")
    np.savez_compressed(path, **d)
    print("mean_var", sorted(d), f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    import sys

    if "--mean-var-only" in sys.argv:
        make_mean_var()
    else:
        main()
        make_mean_var()
