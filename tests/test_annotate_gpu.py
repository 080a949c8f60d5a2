"""GPU parity tests of the annotate path (-m gpu): CUDA through the C ABI vs the CPU oracle and the
committed reference fixtures.  Labels, dims, statuses and counts must be bit-exact."""
import os

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
@pytest.mark.parametrize("flags", [0, 1, 2, 16])     # default / all-f64 / no culling at all / no brick culling
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


def test_full_size_c2_vs_oracle():
    """BASELINE config 2 at full size (64 tracklets x 40 frames, 0.2 m): every label of every tracklet against the
    CPU oracle, the fast path against the all-f64 path and the path without brick culling, determinism, and
    U + occupied == V."""
    from objectcentricocccompletion_b200 import synth
    from oracle import oracle
    from tests.util import assert_same_results

    batch = synth.config_batch("c2", seed=0)
    a = _cuda(batch)
    exp = oracle.annotate_batch(batch, threads=os.cpu_count())
    assert_same_results(a, exp, "c2 full")
    for flags in (1, 16):
        b = _cuda(batch, flags=flags)
        for x, y in zip(a, b):
            assert x["status"] == y["status"] == "ok"
            assert (x["occ"] == y["occ"]).all()
            assert x["n_unknown"] == int((x["occ"] != 1).sum())
    again = _cuda(batch)
    assert all((x["occ"] == y["occ"]).all() for x, y in zip(a, again))


def test_full_size_c3_vs_oracle():
    """BASELINE config 3 at full size: 16 truck / bus tracklets x 40 frames at 0.1 m voxels (8x grids, long rays)."""
    from objectcentricocccompletion_b200 import synth
    from oracle import oracle
    from tests.util import assert_same_results

    batch = synth.config_batch("c3", seed=0)
    got = _cuda(batch)
    exp = oracle.annotate_batch(batch, threads=os.cpu_count())
    assert_same_results(got, exp, "c3 full")
    assert sum(int(e["occ"].size) for e in exp if e["occ"] is not None) > 1_000_000


def test_brick_cull_effect_and_u8_labels():
    """The brick cull only removes tests that must fail: same labels with fewer executed tests; the one-byte label
    output equals the int32 one."""
    import torch

    from objectcentricocccompletion_b200 import occ_annotate, synth

    batch = synth.make_batch(16, 30, 0.2, seed=11)
    pk = occ_annotate.pack_tracklets(batch)
    host = occ_annotate.HostBuffers(pk)
    d = occ_annotate.DeviceTracklets(pk, labels="both")
    d.upload(host)
    out = {}
    for flags in (0, occ_annotate.FLAG_NO_BRICK_CULL):
        d.run(flags)
        torch.cuda.synchronize()
        out[flags] = (d.labels.cpu().numpy().copy(), d.labels_u8.cpu().numpy().copy(), int(d.n_steps.sum()))
        assert (out[flags][0] == out[flags][1].astype(np.int32)).all()
    assert (out[0][0] == out[16][0]).all()
    print("executed tests with / without the brick cull:", out[0][2], out[16][2])
    assert out[0][2] < 0.8 * out[16][2]
    only_u8 = occ_annotate.DeviceTracklets(pk, labels="u8")
    only_u8.upload(host)
    only_u8.run()
    torch.cuda.synchronize()
    assert only_u8.labels is None and (only_u8.labels_u8.cpu().numpy() == out[0][1]).all()


def test_work_overflow_is_reported():
    """A violated host bound (here: brick_off too small for one tracklet) must surface as a status, never as
    silently wrong labels."""
    import torch

    from objectcentricocccompletion_b200 import occ_annotate, synth

    batch = synth.make_batch(3, 12, 0.2, seed=7, small=True)
    pk = occ_annotate.pack_tracklets(batch)
    pk.brick_off = pk.brick_off.copy()
    pk.brick_off[2:] -= pk.brick_off[2] - pk.brick_off[1] - 1          # tracklet 1 gets a single brick
    d = occ_annotate.DeviceTracklets(pk)
    d.upload(occ_annotate.HostBuffers(pk))
    d.run()
    torch.cuda.synchronize()
    res = d.results()
    assert [r["status"] for r in res] == ["ok", "work_overflow", "ok"]
    good = _cuda(batch)
    assert (res[0]["occ"] == good[0]["occ"]).all() and (res[2]["occ"] == good[2]["occ"]).all()


def test_cuda_arith_flag_matches_torch_cuda_division():
    """Flag bit 5 evaluates the scalar divisions of the path the way torch evaluates them on CUDA tensors
    (x * (1 / vs)) and the in-box test the way the reference's CUDA kernel does (device trig, FMA contraction;
    pinned bit-exactly in tests/test_ref_cuda_gpu.py).  dims equal a torch-CUDA evaluation of
    occ_annotate.py:414-416; the occupied voxels equal a torch-CUDA evaluation of :425 on the oracle's kept points
    except where a point sits within an ulp of a box face (the two in-box arithmetics differ there: < 1e-4)."""
    import torch

    from objectcentricocccompletion_b200 import occ_annotate, synth
    from oracle import oracle

    batch = synth.make_batch(6, 14, 0.3, seed=3)                       # 0.3: 1/0.3f is not a float, the forms differ
    base = oracle.annotate_batch(batch, threads=8)
    got = _cuda(batch, flags=occ_annotate.FLAG_CUDA_ARITH)
    dbg = [oracle.annotate_tracklet_debug(batch, t) for t in range(len(batch.tracklets))]
    nvox = nbad = 0
    for t, (g, e) in enumerate(zip(got, base)):
        if e["occ"] is None:
            continue
        size = torch.from_numpy(e["size"]).cuda()
        dims = torch.ceil(size / batch.voxel_size).int().cpu().numpy()               # :414-416 on CUDA
        assert (g["dims"] == dims).all()
        loc = torch.from_numpy(np.ascontiguousarray(dbg[t]["loc"][dbg[t]["keep"]])).cuda()
        mb = torch.tensor([-e["size"][0] * np.float32(0.5), -e["size"][1] * np.float32(0.5), 0.0],
                          dtype=torch.float32).cuda()
        q = torch.floor((loc - mb) / batch.voxel_size).long()                         # :425 on CUDA
        keep = (q < torch.from_numpy(dims).cuda().long()).all(1)
        q = q[keep]
        occ = torch.zeros(tuple(int(v) for v in dims), dtype=torch.bool, device="cuda")
        occ[q[:, 0], q[:, 1], q[:, 2]] = True
        nvox += occ.numel()
        nbad += int(((g["occ"] == 1) != occ.cpu().numpy()).sum())
    assert nvox > 0 and nbad <= 1e-4 * nvox, (nbad, nvox)


def test_fast_path_margins():
    """The f32 arctangent stays inside the error the fast kernel's margins assume, and the fraction of
    tests that need the exact f64 recheck stays small (they are correct either way)."""
    import ctypes

    import torch

    from objectcentricocccompletion_b200 import _lib, occ_annotate, synth

    err = (ctypes.c_double * 2)()
    _lib.check(_lib.lib().occb200_selftest_atan2(200_000_000, 12345, ctypes.addressof(err), _lib.stream_ptr()), "selftest")
    print('max |atan2_fast - atan2| =', err[0], ' narrow path:', err[1])
    assert 0 < err[0] < 1.6e-6, err[0]          # the margins assume 2e-6
    assert 0 < err[1] < 0.9e-6, err[1]          # the margins assume 1e-6
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
