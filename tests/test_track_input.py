"""Tracklet point extraction (tools/ctrl/generate_track_input.py:69-117)."""
import numpy as np
import pytest


def test_enlarged_boxes_cpu():
    """lidar_box3d.py:269-285, incl. the negative-width guard."""
    import torch

    from objectcentricocccompletion_b200.track_input import enlarged_boxes

    b = np.array([[1, 2, 3, 2.0, 4.0, 1.5, 0.3], [0, 0, 0, 0.3, 4.0, 1.5, -1.0]], np.float32)
    for ew in (0.5, -0.2):
        t = torch.from_numpy(b).clone()
        e = t.clone()
        e[:, 3:6] += ew * 2
        e[:, 2] -= ew
        if ew < 0:
            bad = (e[:, 3:6] <= 0).any(1)
            e[bad] = t[bad]
        assert (enlarged_boxes(b, ew) == e.numpy()).all()
    assert (enlarged_boxes(b, -0.2)[1] == b[1]).all()


def test_pc_list_file_round_trip_cpu(tmp_path):
    from objectcentricocccompletion_b200.track_input import load_pc_list, save_pc_list

    pcs = [np.arange(12, dtype=np.float32).reshape(2, 6), np.zeros((0, 6), np.float32), np.ones((5, 6), np.float32)]
    p = save_pc_list(str(tmp_path), "segment-1", "obj7", pcs)
    assert p.endswith("segment-1--obj7.npy")
    back = load_pc_list(p)
    assert len(back) == 3 and all((a == b).all() and a.shape == b.shape for a, b in zip(pcs, back))


@pytest.mark.gpu
def test_extract_segment_vs_oracle_gpu():
    """Frame clouds = all candidate points of a synthetic segment; every tracklet gets, per timestamp, exactly
    the rows the one-box-at-a-time in-box test of the oracle keeps (host trig: bit-identical to the CPU twin),
    in order; overlapping enlarged boxes share points."""
    from objectcentricocccompletion_b200 import synth, track_input
    from oracle import oracle

    batch = synth.make_batch(4, 12, 0.2, seed=5, small=True)
    B = 12
    frames = {}
    for f in range(B):
        rows = [np.concatenate([t.points[k][:, :3], np.full((len(t.points[k]), 3), 0.5, np.float32)], 1)
                for t in batch.tracklets for k, fid in enumerate(t.frame_ids) if fid == f]
        frames[1000 + f] = np.concatenate(rows, 0).astype(np.float32)
    trks = [dict(ts=[1000 + int(f) for f in t.frame_ids], boxes=t.boxes) for t in batch.tracklets]
    # a second tracklet on top of the first one: its enlarged boxes overlap -> shared points
    trks.append(dict(ts=trks[0]["ts"][:6], boxes=trks[0]["boxes"][:6] + np.float32([0.3, 0, 0, 0, 0, 0, 0.05])))
    got = track_input.extract_segment(trks, frames, extra_width=0.5, host_trig=True)
    n_shared = 0
    for t, pcl in zip(trks, got):
        for k, ts in enumerate(t["ts"]):
            pc = frames[ts]
            box = track_input.enlarged_boxes(t["boxes"][k:k + 1], 0.5)
            keep = oracle.points_in_boxes_batch(pc[None, :, :3], box[None])[0, :, 0] == 1
            assert pcl[k].shape == (int(keep.sum()), 6)
            assert (pcl[k] == pc[keep]).all()
    first = {tuple(r) for r in np.concatenate(got[0][:6], 0)[:, :3]}
    n_shared = sum(tuple(r) in first for r in np.concatenate(got[-1], 0)[:, :3])
    assert n_shared > 0
    assert sum(len(p) for p in got[0]) > 0
