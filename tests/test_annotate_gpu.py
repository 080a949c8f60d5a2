"""GPU parity tests of the annotate path (-m gpu): CUDA through the C ABI vs the CPU oracle and the
committed reference fixtures.  Labels, dims, statuses and counts must be bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cuda(batch, **kw):
    import objectcentricocccompletion_b200 as occ

    return occ.annotate_batch(batch, **kw)


@pytest.mark.parametrize("name", ["annotate_small", "annotate_large", "annotate_edge"])
@pytest.mark.parametrize("flags", [0, 1])
def test_reference_fixture(name, flags):
    """Outputs of the reference itself (oracle/make_golden.py) -- with the fixture's host-derived values."""
    from tests.util import load_golden

    batch, override, status, occ = load_golden(name)
    got = _cuda(batch, pack_override=override, flags=flags)
    assert [g["status"] for g in got] == status
    for t, e in occ.items():
        assert got[t]["occ"].shape == e.shape
        assert int((got[t]["occ"] != e).sum()) == 0


@pytest.mark.parametrize("cfg", [
    dict(n=1, b=20, vs=0.2, kind="vehicle", seed=0, small=False),      # BASELINE config 1
    dict(n=6, b=14, vs=0.2, kind="vehicle", seed=1, small=False),
    dict(n=3, b=12, vs=0.1, kind="large", seed=2, small=False),        # config 3 shape (long rays, 8x grid)
    dict(n=5, b=25, vs=0.25, kind="vehicle", seed=3, small=True),
    dict(n=4, b=10, vs=0.15, kind="vehicle", seed=4, small=True),
])
@pytest.mark.parametrize("flags", [0, 1, 2])     # default / all-f64 / no pair culling
def test_vs_oracle(cfg, flags):
    from objectcentricocccompletion_b200 import synth
    from oracle import oracle
    from tests.util import assert_same_results

    batch = synth.make_batch(cfg["n"], cfg["b"], cfg["vs"], cfg["kind"], cfg["seed"], small=cfg["small"])
    exp = oracle.annotate_batch(batch, threads=8)
    got = _cuda(batch, flags=flags)
    assert_same_results(got, exp, str(cfg))
    lab = np.concatenate([e["occ"].ravel() for e in exp if e["occ"] is not None])
    assert (np.bincount(lab, minlength=3) > 0).all()


def test_edge_cases_vs_oracle():
    """Short tracklets, no in-box points, empty frames, points on the far boundary, negative wrap."""
    from objectcentricocccompletion_b200 import synth
    from oracle import oracle
    from oracle.make_golden import edge_batch
    from tests.util import assert_same_results

    b = edge_batch()
    # a tracklet whose candidate points sit exactly on / beyond the far faces and just below the floor
    t = b.tracklets[4]
    pts = [p.copy() for p in t.points]
    for i in range(len(pts)):
        if len(pts[i]):
            pts[i][: len(pts[i]) // 3, 2] -= np.float32(0.05)          # some q_z = -1 -> wraps to the top layer
    b.tracklets[4] = synth.Tracklet(boxes=t.boxes, points=pts, segment=0, frame_ids=t.frame_ids)
    exp = oracle.annotate_batch(b)
    got = _cuda(b)
    assert [e["status"] for e in exp][:2] == ["skip_short", "no_points"]
    assert_same_results(got, exp, "edge")


def test_empty_batch_and_stride6():
    import objectcentricocccompletion_b200 as occ
    from objectcentricocccompletion_b200 import synth
    from oracle import oracle
    from tests.util import assert_same_results

    assert _cuda(synth.TrackletBatch(segments=[], tracklets=[], voxel_size=0.2)) == []
    b = synth.make_batch(2, 10, 0.2, seed=9, small=True)
    exp = oracle.annotate_batch(b)
    # KITTI-format rows: xyz + 3 extra columns (tools/ctrl/utils.py:60-66) -> point_stride 6
    for t in b.tracklets:
        t.points = [np.concatenate([p, np.ones((len(p), 3), np.float32)], 1) for p in t.points]
    got = _cuda(b)
    assert_same_results(got, exp, "stride6")


def test_projection_operator_fixture():
    """occ.point_cloud_to_range_image_idx vs the reference function's recorded output."""
    import os

    import torch

    import objectcentricocccompletion_b200 as occ
    from tests.util import GOLDEN

    d = np.load(os.path.join(GOLDEN, "projection.npz"))
    for k in range(3):
        H, W = (int(v) for v in d[f"p{k}_hw"])
        idx, rng = occ.point_cloud_to_range_image_idx(torch.from_numpy(d[f"p{k}_points"]).cuda(),
                                                      torch.from_numpy(d[f"p{k}_extrinsics"]).cuda(),
                                                      torch.from_numpy(d[f"p{k}_incl"]).cuda(), (H, W))
        assert (idx.cpu().numpy() == d[f"p{k}_idx"]).all()
        assert (rng.cpu().numpy().view(np.uint64) == d[f"p{k}_range"].view(np.uint64)).all()


def test_full_size_properties():
    """BASELINE config 2 at full size: properties that do not need the oracle -- determinism,
    occupied voxels == voxels hit by the oracle-independent host re-voxelisation, U + occupied == V,
    and a sampled subset of tracklets against the oracle."""
    from objectcentricocccompletion_b200 import synth
    from oracle import oracle

    batch = synth.config_batch("c2", seed=0)
    a = _cuda(batch)
    b = _cuda(batch, flags=1)
    for x, y in zip(a, b):
        assert x["status"] == y["status"] == "ok"
        assert (x["occ"] == y["occ"]).all()
        assert x["n_unknown"] == int((x["occ"] != 1).sum())
    sub = synth.TrackletBatch(segments=batch.segments, tracklets=batch.tracklets[::16], voxel_size=batch.voxel_size)
    exp = oracle.annotate_batch(sub, threads=8)
    for e, g in zip(exp, a[::16]):
        assert (e["occ"] == g["occ"]).all()


def test_fast_path_margins():
    """The f32 arctangent stays inside the error the fast kernel's margins assume, and the fraction of
    tests that need the exact f64 recheck stays small (they are correct either way)."""
    import ctypes

    import torch

    from objectcentricocccompletion_b200 import _lib, occ_annotate, synth

    err = ctypes.c_double(0)
    _lib.check(_lib.lib().occb200_selftest_atan2(200_000_000, 12345, ctypes.addressof(err), _lib.stream_ptr()), "selftest")
    print('max |atan2_fast - atan2| =', err.value)
    assert 0 < err.value < 1.6e-6, err.value
    batch = synth.make_batch(8, 20, 0.2, seed=5)
    pk = occ_annotate.pack_tracklets(batch)
    d = occ_annotate.DeviceTracklets(pk)
    d.upload(occ_annotate.HostBuffers(pk))
    d.run()
    torch.cuda.synchronize()
    n_recheck, cap = d.queue_stats()
    res = d.results()
    executed = sum(r["n_steps"] for r in res if r["occ"] is not None)
    assert n_recheck < cap and n_recheck < 0.05 * executed, (n_recheck, executed)


def test_many_tracklets_two_pipes():
    """1024 tracklets over 16 segments (dozens of them take the size-correction redo pass): the fast path,
    the fast path without culling and the all-f64 path agree bit for bit, repeated runs on two buffers /
    two streams (bench.py's e2e pipeline) reproduce them, and a sampled subset matches the oracle."""
    import torch

    from objectcentricocccompletion_b200 import occ_annotate, synth
    from oracle import oracle

    batch = synth.config_batch("c5s", seed=0)
    pk = occ_annotate.pack_tracklets(batch)
    host = occ_annotate.HostBuffers(pk)
    keys = ("labels", "dims", "status", "n_unknown")
    d, d2 = occ_annotate.DeviceTracklets(pk), occ_annotate.DeviceTracklets(pk)
    d.upload(host)
    d.run(occ_annotate.FLAG_FORCE_F64)
    torch.cuda.synchronize()
    ref = {k: getattr(d, k).cpu().numpy().copy() for k in keys}
    want = d.results()
    pipes = [(torch.cuda.Stream(), d), (torch.cuda.Stream(), d2)]
    for flags in (0, occ_annotate.FLAG_NO_CULL):
        for rep in range(3):
            for i in range(6):
                st, dd = pipes[i % 2]
                with torch.cuda.stream(st):
                    dd.upload(host)
                    dd.run(flags)
            torch.cuda.synchronize()
            for _, dd in pipes:
                for k in keys:
                    assert (getattr(dd, k).cpu().numpy() == ref[k]).all(), (flags, rep, k)
    sel = list(range(0, len(batch.tracklets), 37))
    sub = synth.TrackletBatch(segments=batch.segments, tracklets=[batch.tracklets[i] for i in sel],
                              voxel_size=batch.voxel_size)
    exp = oracle.annotate_batch(sub, threads=8)
    for e, i in zip(exp, sel):
        assert e["status"] == want[i]["status"]
        if e["occ"] is not None:
            assert (e["occ"] == want[i]["occ"]).all()


def test_graph_replay_matches_direct_launch():
    """The captured CUDA graph (side-stream fork/join included) reproduces the direct launches, also after
    the inputs are replaced by ``upload``."""
    import torch

    from objectcentricocccompletion_b200 import occ_annotate, synth

    a = synth.make_batch(6, 14, 0.2, seed=1)
    pk = occ_annotate.pack_tracklets(a)
    d = occ_annotate.DeviceTracklets(pk)
    host = occ_annotate.HostBuffers(pk)
    d.upload(host)
    d.run()
    want = d.results()
    n = d.capture()
    assert n >= 8
    d.labels.zero_()
    d.upload(host)
    assert d.replay() == n
    torch.cuda.synchronize()
    got = d.results()
    for w, g in zip(want, got):
        assert w["status"] == g["status"] and (w["occ"] == g["occ"]).all() and w["n_unknown"] == g["n_unknown"]


@pytest.mark.parametrize("name", ["annotate_small", "annotate_edge"])
def test_mean_var_vs_reference_fixture(name):
    """--save-mean-var grids: means within 1e-5 relative; mean squared deviations within 1e-5 relative plus
    1e-7 m^2 (one ulp of a mean moves the deviation of a point that sits almost on it)."""
    import os

    from tests.util import GOLDEN, load_golden

    gold = np.load(os.path.join(GOLDEN, "mean_var.npz"))
    batch, override, status, _ = load_golden(name)
    got = _cuda(batch, pack_override=override, save_mean_var=True)
    for t, g in enumerate(got):
        key = f"{name}_{t}"
        assert (g["mean_var"] is None) == (key not in gold)
        if g["mean_var"] is None:
            continue
        m, e = g["mean_var"], gold[key]
        assert m.shape == e.shape and m.dtype == np.float32
        assert ((m != 0).any(-1) == (e != 0).any(-1)).all()            # same cells filled
        np.testing.assert_allclose(m[..., :3], e[..., :3], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(m[..., 3:], e[..., 3:], rtol=1e-5, atol=1e-7)


def test_mean_var_vs_oracle_larger():
    from objectcentricocccompletion_b200 import synth
    from oracle import oracle

    batch = synth.make_batch(6, 14, 0.2, seed=1)
    exp = oracle.annotate_mean_var(batch)
    got = _cuda(batch, save_mean_var=True)
    for g, e in zip(got, exp):
        assert (g["mean_var"] is None) == (e is None)
        if e is not None:
            np.testing.assert_allclose(g["mean_var"][..., :3], e[..., :3], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(g["mean_var"][..., 3:], e[..., 3:], rtol=1e-5, atol=1e-7)


def test_recheck_queue_overflow_path():
    """With a 64-entry recheck queue almost every undecided test is decided in place by the exact f64 routine:
    labels must not change."""
    from objectcentricocccompletion_b200 import occ_annotate, synth
    from tests.util import assert_same_results

    batch = synth.make_batch(6, 14, 0.2, seed=1)
    want = _cuda(batch)
    got = _cuda(batch, flags=occ_annotate.FLAG_TINY_QUEUE)
    assert_same_results(got, want, "tiny queue")
    got = _cuda(batch, flags=occ_annotate.FLAG_TINY_QUEUE | occ_annotate.FLAG_NO_CULL)
    assert_same_results(got, want, "tiny queue, no culling")
