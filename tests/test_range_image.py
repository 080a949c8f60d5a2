"""Range-image builder (waymo_converter.py:632-670)."""
import numpy as np
import pytest


def _segment_points(seg, frame, li, jitter=0.0, seed=0):
    """Vehicle-frame returns of one range image: every pixel with a return, un-projected along its ray (the
    inverse of the pixel model, occ_annotate.py:165-193), optionally moved off the pixel centre."""
    E = seg.extrinsics[frame, li].astype(np.float64)
    ri = seg.range_images[li][frame]
    H, W = ri.shape
    incl = seg.inclinations[li].astype(np.float64)[::-1]
    azc = np.arctan2(E[1, 0], E[0, 0])
    rows, cols = np.nonzero(ri > 0)
    rng = np.random.default_rng(seed)
    inc = incl[rows] + jitter * rng.uniform(-1, 1, len(rows)) * 1e-3
    az = 2.0 * np.pi * (W - 0.5 - cols) / W - np.pi - azc + jitter * rng.uniform(-1, 1, len(rows)) * 2e-3
    r = ri[rows, cols].astype(np.float64)
    p_s = np.stack([np.cos(inc) * np.cos(az), np.cos(inc) * np.sin(az), np.sin(inc)], -1) * r[:, None]
    p_v = p_s @ E[:3, :3].T + E[:3, 3]
    return p_v.astype(np.float32)


def _segment():
    from objectcentricocccompletion_b200 import synth

    return synth.make_batch(2, 10, 0.2, seed=3, small=True).segments[0]


def test_oracle_round_trip_cpu():
    """Un-projecting a synthetic range image and building it again gives the image back (pixel centres)."""
    from oracle import oracle

    seg = _segment()
    for li in range(5):
        ri0 = seg.range_images[li][0]
        pts = _segment_points(seg, 0, li)
        ri, rows, cols, rng, bad = oracle.build_range_image(pts, seg.extrinsics[0, li], seg.inclinations[li], ri0.shape)
        assert bad == 0 and ri.dtype == np.float32
        assert ((ri > 0) == (ri0 > 0)).all()
        np.testing.assert_allclose(ri, ri0, rtol=2e-6)


@pytest.mark.gpu
def test_build_range_images_vs_oracle_gpu():
    """All LiDARs of two frames in one call, points off the pixel centres and a duplicated, farther second
    return: rows / columns as the oracle decides them, min range per pixel, zeros elsewhere -- bit for bit."""
    from objectcentricocccompletion_b200 import range_image
    from oracle import oracle

    seg = _segment()
    pts, ext, inc, sizes, exp = [], [], [], [], []
    for frame in (0, 3):
        for li in range(5):
            p = _segment_points(seg, frame, li, jitter=1.0, seed=10 * frame + li)
            far = p[::3] * np.float32(1.01)                      # a second, farther return along nearly the same rays
            p = np.concatenate([p, far], 0)
            pts.append(p)
            ext.append(seg.extrinsics[frame, li])
            inc.append(seg.inclinations[li])
            sizes.append(seg.range_images[li][frame].shape)
            exp.append(oracle.build_range_image(p, ext[-1], inc[-1], sizes[-1]))
    got = range_image.build_range_images(pts, np.stack(ext), inc, sizes)
    assert len(got) == 10
    for g, (ri, rows, cols, rng, bad) in zip(got, exp):
        assert bad == 0
        g = g.cpu().numpy()
        assert g.shape == ri.shape and g.dtype == np.float32
        assert (g.view(np.uint32) == ri.view(np.uint32)).all()
    # 6-column rows (KITTI layout) and an image without points
    p6 = [np.concatenate([pts[0], np.ones((len(pts[0]), 3), np.float32)], 1), np.zeros((0, 6), np.float32)]
    g6 = range_image.build_range_images(p6, np.stack(ext[:2]), inc[:2], sizes[:2])
    assert (g6[0].cpu().numpy().view(np.uint32) == exp[0][0].view(np.uint32)).all()
    assert float(g6[1].abs().max()) == 0.0


@pytest.mark.gpu
def test_merge_virtual_frame_dict_gpu():
    """The converter's inner loop (waymo_converter.py:632-668) on a frame dictionary."""
    from objectcentricocccompletion_b200 import range_image
    from oracle import oracle

    seg = _segment()
    names = ["TOP", "FRONT"]
    fd = {}
    for li, name in enumerate(names):
        ri0 = seg.range_images[li][1]
        H, W = ri0.shape
        pts = _segment_points(seg, 1, li)
        xyz = np.zeros((H, W, 3), np.float32)
        xyz[ri0 > 0] = pts
        first = np.concatenate([ri0[..., None], np.zeros((H, W, 2), np.float32), xyz], -1)
        second = np.zeros_like(first)
        fd[f"{name}_RANGE_IMAGE_FIRST_RETURN"] = first
        fd[f"{name}_RANGE_IMAGE_SECOND_RETURN"] = second
        fd[f"{name}_LIDAR_EXTRINSIC"] = seg.extrinsics[1, li]
        fd[f"{name}_BEAM_INCLINATION"] = seg.inclinations[li]
    out = range_image.merge_virtual(fd, names)
    for li, name in enumerate(names):
        ri0 = seg.range_images[li][1]
        exp = oracle.build_range_image(_segment_points(seg, 1, li), seg.extrinsics[1, li], seg.inclinations[li], ri0.shape)[0]
        got = out[f"{name}_RANGE_IMAGE_MERGE_VIRTUAL"]
        assert (got.view(np.uint32) == exp.view(np.uint32)).all()
        np.testing.assert_allclose(got, ri0, rtol=2e-6)
