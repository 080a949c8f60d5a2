"""GPU tests of dynamic_point_pool_mixed / TrackletPointRoIExtractor (parity unpinned: the TorchEx
kernel is not in the reference tree).  Checked: the reference extractor's own assertions (run inside ``forward``),
and a numpy restatement of the documented semantics (csrc/point_pool.cu) including both truncations."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(seed, n_batch=2, n_frames=3, n_pts=6000):
    rng = np.random.default_rng(seed)
    rois, rframes = [], []
    for b in range(n_batch):
        for f in range(n_frames):
            c = rng.uniform(-20, 20, 2)
            rois.append([b, c[0], c[1], rng.uniform(-1, 0.5), rng.uniform(1.8, 2.4), rng.uniform(4, 5.5),
                         rng.uniform(1.5, 2.0), rng.uniform(-3.1, 3.1)])
            rframes.append(f)
    rois = np.asarray(rois, np.float32)
    rframes = np.asarray(rframes, np.int64)
    # points: around ROI centres (many inside) and far away, random batch / frame
    k = rng.integers(0, len(rois), n_pts)
    pts = np.stack([rois[k, 1] + rng.normal(0, 2.0, n_pts), rois[k, 2] + rng.normal(0, 2.0, n_pts),
                    rois[k, 3] + rng.uniform(-0.5, 2.5, n_pts)], 1).astype(np.float32)
    batch = np.where(rng.random(n_pts) < 0.8, rois[k, 0], rng.integers(0, n_batch, n_pts)).astype(np.int64)
    frame = np.where(rng.random(n_pts) < 0.8, rframes[k], rng.integers(0, n_frames, n_pts)).astype(np.int64)
    return rois, rframes, pts, batch, frame


def _restate(rois7, roi_idx, pts, pts_idx, extra, max_inbox, max_all):
    out = []
    for r, box in enumerate(rois7):
        a = np.float32(box[6] + np.float32(1.5707963267948966))
        ca, sa = np.cos(a, dtype=np.float32), np.sin(a, dtype=np.float32)
        hw, hl, hh = np.float32(0.5) * box[3], np.float32(0.5) * box[4], np.float32(0.5) * box[5]
        dx, dy, lz = pts[:, 0] - box[0], pts[:, 1] - box[1], pts[:, 2] - (box[2] + hh)
        lx, ly = dx * ca - dy * sa, dx * sa + dy * ca
        big = (np.abs(lz) < hh + np.float32(0.5 * extra[2])) & (np.abs(lx) < hl + np.float32(0.5 * extra[0])) & \
              (np.abs(ly) < hw + np.float32(0.5 * extra[1])) & (pts_idx == roi_idx[r])
        small = (np.abs(lz) < hh) & (np.abs(lx) < hl) & (np.abs(ly) < hw)
        for p in np.flatnonzero(big)[:max_inbox]:
            out.append((p, r, lx[p], ly[p], lz[p], float(not small[p])))
    return out[:max_all]


@pytest.mark.parametrize("max_inbox,max_all", [(512, 200000), (20, 200000), (512, 150)])
def test_point_pool_vs_restatement(max_inbox, max_all):
    import torch

    from objectcentricocccompletion_b200.point_pool import dynamic_point_pool_mixed

    rois, rframes, pts, batch, frame = _scene(3)
    mf = int(rframes.max()) + 1
    roi_idx = rois[:, 0].astype(np.int64) * mf + rframes
    pts_idx = batch * mf + frame
    extra = [0.5, 0.5, 0.3]
    pi, ri, feats = dynamic_point_pool_mixed(torch.from_numpy(rois[:, 1:]).cuda(), torch.from_numpy(roi_idx).cuda(),
                                             torch.from_numpy(pts).cuda(), torch.from_numpy(pts_idx).cuda(), extra,
                                             max_inbox, max_all)
    exp = _restate(rois[:, 1:], roi_idx, pts, pts_idx, extra, max_inbox, max_all)
    assert len(exp) > 100
    assert pi.cpu().tolist() == [e[0] for e in exp] and ri.cpu().tolist() == [e[1] for e in exp]
    f = feats.cpu().numpy()
    assert (f[:, :3] == pts[[e[0] for e in exp]]).all()
    assert np.allclose(f[:, 3:6], np.array([e[2:5] for e in exp], np.float32), atol=2e-5)
    assert (f[:, 12] == np.array([e[5] for e in exp], np.float32)).all()
    assert 0 < f[:, 12].sum() < len(f)


def test_tracklet_point_roi_extractor_asserts_hold():
    import torch

    from objectcentricocccompletion_b200.point_pool import TrackletPointRoIExtractor

    rois, rframes, pts, batch, frame = _scene(5)
    for combined in (False, True):
        ext = TrackletPointRoIExtractor(debug=True, extra_wlh=[0.4, 0.4, 0.2], max_inbox_point=256, combined=combined)
        inds, roi_inds, info = ext(torch.from_numpy(pts).cuda(), torch.from_numpy(batch).cuda(), torch.from_numpy(frame).cuda(),
                                   torch.from_numpy(rois).cuda(), torch.from_numpy(rframes).cuda())
        assert inds.numel() > 100 and info["local_xyz"].shape == (inds.numel(), 3)
        assert info["boundary_offset"].shape == (inds.numel(), 6) and info["is_in_margin"].shape == (inds.numel(),)
        assert (roi_inds[1:] >= roi_inds[:-1]).all()              # sorted by ROI, as the reference's output "automatically" is


def test_point_pool_no_hit_returns_the_fake_row():
    import torch

    from objectcentricocccompletion_b200.point_pool import dynamic_point_pool_mixed

    rois = torch.tensor([[0., 0., 0., 2., 4., 1.5, 0.3]]).cuda()
    pts = torch.tensor([[50., 50., 0.], [60., -3., 1.]]).cuda()
    pi, ri, f = dynamic_point_pool_mixed(rois, torch.zeros(1).int().cuda(), pts, torch.zeros(2).int().cuda(), [0, 0, 0], 16)
    assert pi.tolist() == [-1] and ri.tolist() == [-1] and f.shape == (1, 13) and not f.any()
