"""The C pixel loops of the synthetic generator (csrc/synth_render.c) reproduce the NumPy renderer bit for bit, so
the fixtures, the parity tests and the benchmark workloads do not depend on which of the two produced them."""
import numpy as np
import pytest


def _same(a, b):
    assert len(a.segments) == len(b.segments) and len(a.tracklets) == len(b.tracklets)
    for sa, sb in zip(a.segments, b.segments):
        assert (sa.extrinsics == sb.extrinsics).all()
        assert all((x == y).all() for x, y in zip(sa.inclinations, sb.inclinations))
        for x, y in zip(sa.range_images, sb.range_images):
            assert x.dtype == y.dtype == np.float32 and x.shape == y.shape and (x.view(np.uint32) == y.view(np.uint32)).all()
    for ta, tb in zip(a.tracklets, b.tracklets):
        assert (ta.boxes == tb.boxes).all() and (ta.frame_ids == tb.frame_ids).all() and ta.segment == tb.segment
        assert len(ta.points) == len(tb.points)
        for x, y in zip(ta.points, tb.points):
            assert x.shape == y.shape and (x.view(np.uint32) == y.view(np.uint32)).all()


@pytest.mark.parametrize("cfg", [
    dict(n=64, b=40, vs=0.2, kind="vehicle", seed=0, small=False),          # BASELINE config 2, full size
    dict(n=4, b=12, vs=0.1, kind="large", seed=2, small=False),
    dict(n=9, b=14, vs=0.2, kind="vehicle", seed=5, small=True, tps=3),     # several segments
])
def test_c_renderer_equals_numpy_renderer(cfg):
    from objectcentricocccompletion_b200 import synth

    if synth.fast_lib() is None:
        pytest.skip("libocc_synth.so not built")
    kw = dict(num_tracklets=cfg["n"], num_frames=cfg["b"], voxel_size=cfg["vs"], kind=cfg["kind"], seed=cfg["seed"],
              small=cfg["small"], tracklets_per_segment=cfg.get("tps"))
    _same(synth.make_batch(fast=False, **kw), synth.make_batch(fast=True, **kw))


def test_only_segments_generates_the_same_segments():
    """A rank that generates only its own segments of a job gets exactly the data the full job holds."""
    from objectcentricocccompletion_b200 import synth

    full = synth.make_batch(20, 10, 0.2, seed=3, small=True, tracklets_per_segment=4)
    part = synth.make_batch(20, 10, 0.2, seed=3, small=True, tracklets_per_segment=4, only_segments=[1, 3])
    assert part.meta["global_tracklets"] == [4, 5, 6, 7, 12, 13, 14, 15]
    pick = synth.TrackletBatch(segments=[full.segments[1], full.segments[3]],
                               tracklets=[full.tracklets[i] for i in part.meta["global_tracklets"]], voxel_size=0.2)
    for t in pick.tracklets:
        t.segment = {1: 0, 3: 1}[t.segment]
    _same(pick, part)
    for t in part.tracklets:                                        # the contiguous copy pack_tracklets uploads from
        if t.flat is not None:
            assert len(t.flat) == sum(len(p) for p in t.points)
