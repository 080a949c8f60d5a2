"""On-disk formats: round trip through the reference's file layout."""
import os

import numpy as np
import pytest


def _dataset(tmp_path):
    from objectcentricocccompletion_b200 import synth, waymo_io

    batch = synth.make_batch(3, 12, 0.2, seed=21, small=True)
    # one short tracklet: the reference skips it (occ_annotate.py:344)
    t = batch.tracklets[2]
    batch.tracklets[2] = synth.Tracklet(boxes=t.boxes[:8], points=t.points[:8], segment=0, frame_ids=t.frame_ids[:8])
    recs = waymo_io.write_synthetic_dataset(batch, str(tmp_path / "data"))
    return batch, recs


def test_formats_round_trip_cpu(tmp_path):
    """Files -> Segment / Tracklet objects -> npz files.  The loaded candidate sets are supersets of the synthetic
    ones (the frame cloud holds every object's returns, as real data does), so expectations are taken on the
    loaded batch."""
    from objectcentricocccompletion_b200 import waymo_io
    from oracle import oracle

    batch, recs = _dataset(tmp_path)
    root = str(tmp_path / "data")
    pc = waymo_io.read_velodyne_bin(os.path.join(root, "kitti_format/training/velodyne/0000000.bin"))
    assert pc.dtype == np.float32 and pc.shape[1] == 6
    ts2idx = waymo_io.load_idx2timestamp(os.path.join(root, "kitti_format"))
    assert len(ts2idx) == 12
    loaded, kept = waymo_io.build_segment_batch(recs[:2], ts2idx, root, "training", 0.2, device_select=False)
    assert kept == [0, 1]
    seg, seg0 = loaded.segments[0], batch.segments[0]
    assert (seg.extrinsics == seg0.extrinsics).all()
    assert all((a == b).all() for a, b in zip(seg.range_images, seg0.range_images))
    assert all((a == b).all() for a, b in zip(seg.inclinations, seg0.inclinations))
    for t0, t1 in zip(batch.tracklets[:2], loaded.tracklets):
        assert (t0.boxes == t1.boxes).all() and (t0.frame_ids == t1.frame_ids).all()
        assert all(b.shape[1] == 6 and b.dtype == np.float32 for b in t1.points)
    exp = oracle.annotate_batch(loaded)
    assert all(e["status"] == "ok" for e in exp)
    # driver semantics with the oracle as the annotate function (no GPU needed)
    out = str(tmp_path / "out")
    paths = waymo_io.annotate_from_disk(recs, root, out, annotate_fn=oracle.annotate_batch)
    assert paths[2] is None and all(p and os.path.isfile(p) for p in paths[:2])
    assert paths[0].endswith(os.path.join("training", "segment-0000", "obj0.npz"))
    z = np.load(paths[0])
    assert list(z.keys()) == ["occ"] and z["occ"].dtype == np.int32 and (z["occ"] == exp[0]["occ"]).all()
    # resume: existing files are kept, not recomputed (occ_annotate.py:335-343)
    calls = []
    waymo_io.annotate_from_disk(recs, root, out, annotate_fn=lambda b: calls.append(1) or oracle.annotate_batch(b))
    assert calls == []
    waymo_io.annotate_from_disk(recs, root, out, overwrite=True,
                                annotate_fn=lambda b: calls.append(1) or oracle.annotate_batch(b))
    assert calls == [1]


@pytest.mark.gpu
def test_annotate_from_disk_gpu(tmp_path):
    from objectcentricocccompletion_b200 import waymo_io
    from oracle import oracle

    batch, recs = _dataset(tmp_path)
    root = str(tmp_path / "data")
    paths = waymo_io.annotate_from_disk(recs, root, str(tmp_path / "out"))
    ts2idx = waymo_io.load_idx2timestamp(os.path.join(root, "kitti_format"))
    host, _ = waymo_io.build_segment_batch(recs[:2], ts2idx, root, "training", 0.2, device_select=False)
    exp = oracle.annotate_batch(host)
    for p, e in zip(paths[:2], exp):
        assert (np.load(p)["occ"] == e["occ"]).all()
    assert paths[2] is None
    # candidate selection on the device picks exactly the host's candidates, in cloud order
    dev, _ = waymo_io.build_segment_batch(recs[:2], ts2idx, root, "training", 0.2, device_select=True)
    for th, td in zip(host.tracklets, dev.tracklets):
        assert len(th.points) == len(td.points)
        for a, b in zip(th.points, td.points):
            assert a.shape == b.shape and (a == b).all()


def test_missing_raw_frame_drops_only_its_tracklets(tmp_path):
    """A missing raw frame file aborts the tracklets that touch its timestamp, not the segment
    (occ_annotate.py:503-510)."""
    from objectcentricocccompletion_b200 import synth, waymo_io
    from oracle import oracle

    batch = synth.make_batch(3, 12, 0.2, seed=22, small=True)
    t = batch.tracklets[1]                                         # tracklet 1 only lives in frames 0..9
    batch.tracklets[1] = synth.Tracklet(boxes=t.boxes[:10], points=t.points[:10], segment=0, frame_ids=t.frame_ids[:10])
    root = str(tmp_path / "data")
    recs = waymo_io.write_synthetic_dataset(batch, root)
    os.remove(os.path.join(root, "waymo_raw", "training", "0000011.pkl"))      # the last frame's raw file
    paths = waymo_io.annotate_from_disk(recs, root, str(tmp_path / "out"), annotate_fn=oracle.annotate_batch,
                                        device_select=False)
    assert paths[0] is None and paths[2] is None and paths[1] is not None and os.path.isfile(paths[1])


@pytest.mark.gpu
def test_candidate_selection_large_frames_gpu():
    """180 k-point frame clouds, 64 boxes per frame: the device selection equals the numpy sphere test for every
    tracklet-frame (content and order), and each cloud is read once."""
    import time

    from objectcentricocccompletion_b200.candidates import candidate_spheres, select_candidates

    rng = np.random.default_rng(0)
    NF, T, M = 6, 64, 180_000
    clouds = [np.concatenate([rng.uniform(-60, 60, (M, 2)), rng.uniform(-2, 4, (M, 1)), rng.random((M, 3))], 1).astype(np.float32)
              for _ in range(NF)]
    clouds[3] = clouds[3][:1000]                                   # ragged cloud sizes
    boxes, frames = [], []
    for t in range(T):
        fr = np.sort(rng.choice(NF, size=rng.integers(1, NF + 1), replace=False))
        b = np.concatenate([rng.uniform(-50, 50, (len(fr), 2)), rng.uniform(-1, 1, (len(fr), 1)),
                            rng.uniform(1.5, 8, (len(fr), 3)), rng.uniform(-3, 3, (len(fr), 1))], 1).astype(np.float32)
        boxes.append(b)
        frames.append(fr)
    t0 = time.perf_counter()
    pts, cnt = select_candidates(clouds, boxes, frames, margin=0.5)
    dt = time.perf_counter() - t0
    off = np.concatenate([[0], np.cumsum(cnt)])
    i = 0
    for b7, fr in zip(boxes, frames):
        sph = candidate_spheres(b7, 0.5)
        for s4, f in zip(sph, fr):
            pc = clouds[int(f)]
            d = pc[:, :3] - s4[:3]
            want = pc[(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]) <= s4[3]]
            got = pts[off[i]: off[i + 1]]
            assert got.shape == want.shape and (got == want).all(), (i, got.shape, want.shape)
            i += 1
    print(f"select_candidates: {NF} frames x {M} points x {T} boxes in {dt * 1e3:.1f} ms (incl. upload / download)")


def test_job_driver_cpu(tmp_path, monkeypatch):
    """tools/occ_annotate_job.py: reference flags, type filter, segment dealing over two ranks, resume; the
    annotate function is the oracle here (the CUDA one is the default)."""
    import importlib.util

    from objectcentricocccompletion_b200 import synth, waymo_io
    from oracle import oracle

    spec = importlib.util.spec_from_file_location("job", os.path.join(os.path.dirname(__file__), "..", "tools",
                                                                      "occ_annotate_job.py"))
    job = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(job)

    root = str(tmp_path / "data")
    batch = synth.make_batch(4, 10, 0.2, seed=40, tracklets_per_segment=2, small=True)      # two segments
    recs = waymo_io.write_synthetic_dataset(batch, root)
    recs[1].type = 2                                              # a pedestrian: filtered out by --object-type vehicle
    trk_file = str(tmp_path / "tracklets.npz")
    waymo_io.save_tracklet_records(trk_file, recs)
    back = waymo_io.load_tracklet_records(trk_file)
    assert [(r.segment_name, r.id, r.type, list(r.ts_list)) for r in back] == \
        [(r.segment_name, r.id, r.type, list(r.ts_list)) for r in recs]
    assert all((a.boxes == b.boxes).all() for a, b in zip(recs, back))
    out = str(tmp_path / "out")
    argv = ["--data-root", root, "--out-dir", out, "--tracklets", trk_file, "--voxel-size", "0.2"]
    written = []
    for rank in range(2):
        monkeypatch.setenv("RANK", str(rank))
        monkeypatch.setenv("WORLD_SIZE", "2")
        written.append(job.main(argv, annotate_fn=oracle.annotate_batch))
    assert len(written[0]) + len(written[1]) == 3                 # 4 tracklets, one filtered by type
    files = [p for w in written for p in w if p]
    assert len(files) == 3 and all(os.path.isfile(p) for p in files)
    assert {os.path.basename(os.path.dirname(p)) for p in written[0] if p} == {"segment-0000"}
    assert {os.path.basename(os.path.dirname(p)) for p in written[1] if p} == {"segment-0001"}

