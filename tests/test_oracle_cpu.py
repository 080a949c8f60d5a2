"""CPU suite (-m "not gpu"): the oracle against the reference's golden vectors and the committed
fixtures, the host packing logic, and the C-ABI surface of libocc_b200.so (no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import oracle
from tests.util import GOLDEN, load_golden


# ---- reference golden vectors (tests/test_models/test_common_modules/test_roiaware_pool3d.py:43-120)
BOXES = np.array([[1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 0.3], [-10.0, 23.0, 16.0, 10, 20, 20, 0.5]], np.float32)
PTS15 = np.array([[1, 2, 3.3], [1.2, 2.5, 3.0], [0.8, 2.1, 3.5], [1.6, 2.6, 3.6], [0.8, 1.2, 3.9], [-9.2, 21.0, 18.2],
                  [3.8, 7.9, 6.3], [4.7, 3.5, -12.2], [3.8, 7.6, -2], [-10.6, -12.9, -20], [-16, -18, 9],
                  [-21.3, -52, -5], [0, 0, 0], [6, 7, 8], [-2, -3, -4]], np.float32)


def test_points_in_boxes_cpu_golden():
    exp = np.array([[1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0]], np.int32)
    assert (oracle.points_in_boxes_cpu(PTS15, BOXES) == exp).all()


def test_points_in_boxes_gpu_golden_semantics():
    boxes = BOXES[:, None, :]                                  # [2,1,7]
    pts = np.stack([PTS15[:8], np.array([[3.8, 7.6, -2], [-10.6, -12.9, -20], [-16, -18, 9], [-21.3, -52, -5],
                                          [0, 0, 0], [6, 7, 8], [-2, -3, -4], [6, 4, 9]], np.float32)], 0)
    exp = np.array([[0, 0, 0, 0, 0, -1, -1, -1], [-1] * 8], np.int32)
    assert (oracle.points_in_boxes_gpu(pts, boxes) == exp).all()
    expb = np.array([[[1, 0], [1, 0], [1, 0], [1, 0], [1, 0], [0, 1]] + [[0, 0]] * 9], np.int32)
    assert (oracle.points_in_boxes_batch(PTS15[None], BOXES[None]) == expb).all()


def test_voxel_generator_kat():
    """tests/test_models/test_voxel_encoder/test_voxel_generator.py:6-22"""
    np.random.seed(0)
    points = np.random.rand(1000, 4).astype(np.float32)
    v, c, n = oracle.hard_voxelize(points, [0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1], 1000, 20000)
    assert (c == np.array([[7, 81, 1], [6, 81, 0], [7, 80, 1], [6, 81, 1], [7, 81, 0], [6, 80, 1], [7, 80, 0], [6, 80, 0]])).all()
    assert (n == np.array([120, 121, 127, 134, 115, 127, 125, 131])).all()
    assert v.shape == (8, 1000, 4)


def test_dynamic_scatter_property():
    """tests/test_models/test_voxel_encoder/test_dynamic_scatter.py:55-84 (property test, reused)."""
    rng = np.random.default_rng(0)
    feats = (rng.random((20000, 3)) * 100 - 50).astype(np.float32)
    coors = rng.integers(-1, 20, (20000, 3)).astype(np.int32)
    ref_c = np.unique(coors, axis=0)
    ref_c = ref_c[ref_c.min(-1) >= 0]
    for mode, fn in (("mean", lambda a: a.mean(0)), ("max", lambda a: a.max(0))):
        vf, vc, mp, cnt = oracle.dynamic_scatter_fwd(feats, coors, mode)
        assert (vc == ref_c).all()
        sel = np.random.default_rng(1).choice(len(ref_c), 50, replace=False)
        for s in sel:
            m = (coors == ref_c[s]).all(-1)
            assert np.allclose(vf[s], fn(feats[m].astype(np.float64)), rtol=1e-5, atol=1e-2)
            assert cnt[s] == m.sum() and (mp[m] == s).all()
    # empty input and all-invalid input (:22-53)
    e = oracle.dynamic_scatter_fwd(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32), "mean")
    assert e[0].shape == (0, 3) and e[1].shape == (0, 3)
    inv = oracle.dynamic_scatter_fwd(feats, np.full_like(coors, -1), "max")
    assert inv[0].shape[0] == 0 and (inv[2] == -1).all()


def test_dynamic_scatter_drop_first_quirk():
    """scatter_points_cuda.cu:207-210: with no negative coordinate the smallest voxel is dropped."""
    coors = np.array([[0, 0, 1], [0, 0, 0], [0, 0, 1], [2, 1, 0]], np.int32)
    feats = np.arange(8, dtype=np.float32).reshape(4, 2)
    vf, vc, mp, cnt = oracle.dynamic_scatter_fwd(feats, coors, "sum")
    assert (vc == np.array([[0, 0, 1], [2, 1, 0]])).all() and (mp == np.array([0, -1, 0, 1])).all()
    assert (vf == np.array([[4, 6], [6, 7]], np.float32)).all() and (cnt == np.array([2, 1])).all()


@pytest.mark.parametrize("name", ["annotate_small", "annotate_large", "annotate_edge"])
def test_oracle_matches_reference_fixture(name):
    """Committed outputs of the reference run (oracle/make_golden.py) vs the C oracle."""
    batch, override, status, occ = load_golden(name)
    pk = oracle.PackedBatch(batch, pack_override=override)
    got = oracle.annotate_batch(batch, packed=pk)
    assert [g["status"] for g in got] == status
    seen = np.zeros(3, np.int64)
    for t, e in occ.items():
        assert got[t]["occ"].shape == e.shape and (got[t]["occ"] == e).all()
        seen += np.bincount(e.ravel(), minlength=3)
    assert (seen > 0).all(), "fixture must contain unknown, occupied and free voxels"


def test_oracle_projection_fixture():
    d = np.load(os.path.join(GOLDEN, "projection.npz"))
    for k in range(3):
        H, W = (int(v) for v in d[f"p{k}_hw"])
        idx, rng = oracle.point_cloud_to_range_image_idx(d[f"p{k}_points"], d[f"p{k}_extrinsics"], d[f"p{k}_incl"], (H, W))
        assert (idx == d[f"p{k}_idx"]).all()
        assert (rng.view(np.uint64) == d[f"p{k}_range"].view(np.uint64)).all()


def test_host_trig_matches_fixture():
    """The fixture stores the torch-CPU trig it was made with; this host must reproduce it (else the
    parity tests fall back to the stored values, which they do anyway -- this is the early warning)."""
    batch, override, _, _ = load_golden("annotate_small")
    pk = oracle.PackedBatch(batch)
    assert (pk.trig.view(np.uint32) == override["trig"].view(np.uint32)).mean() > 0.99


# ---- the product's host logic and C-ABI surface (no GPU needed)
def test_library_exports_every_declared_symbol():
    from objectcentricocccompletion_b200 import _lib

    header = open(_lib.HEADER).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(occb200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/occ_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.lib().occb200_abi_version() == _lib.ABI_VERSION


def test_pack_matches_oracle_pack():
    """Product packing (occ_annotate.pack_tracklets) and the oracle's independent packing agree."""
    from objectcentricocccompletion_b200 import occ_annotate, synth

    b = synth.make_batch(4, 11, 0.2, seed=3, small=True)
    pk = occ_annotate.pack_tracklets(b)
    ok = oracle.PackedBatch(b)
    assert (pk.trk_frame_off == ok.trk_frame_off).all() and (pk.frame_pt_off == ok.pt_off).all()
    assert (pk.label_off == ok.label_off).all() and (pk.frame_sf == ok.frame_sf).all()
    for f in ("incl_off", "H", "W", "v2l", "azc"):
        assert (pk.sensors[f] == ok.sensors[f]).all(), f
    assert (pk.poses["box"] == ok.boxes).all()
    assert (np.stack([pk.poses[k] for k in ("cos_m", "sin_m", "cos_p", "sin_p")], 1) == ok.trig).all()
    assert (pk.sensors["incl_mono"] == -1).all()          # flipped ascending tables are descending
    assert (pk.incl_pool == ok.incl_pool).all()
    # the product pool pads every (segment, LiDAR) part to a multiple of 16 floats (windowed upload): same images
    mine, theirs = pk.ri_pool, ok.ri_pool
    for a, b_ in zip(pk.sensors.reshape(-1), ok.sensors.reshape(-1)):
        n = int(a["H"]) * int(a["W"])
        assert (mine[a["ri_off"]: a["ri_off"] + n] == theirs[b_["ri_off"]: b_["ri_off"] + n]).all()
    assert pk.ri_len % 16 == 0 and (pk.trk_smax >= pk.poses["box"][pk.trk_frame_off[:-1], 3:6]).all()


def test_ops_fail_loudly_without_cuda():
    import torch

    import objectcentricocccompletion_b200 as occ

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        occ.Voxelization([0.2] * 3, [0, 0, 0, 1, 1, 1], -1)(torch.zeros(4, 3))
    with pytest.raises(RuntimeError):
        occ.scatter_v2(torch.zeros(4, 3), torch.zeros(4, 3, dtype=torch.long), "mean")
    with pytest.raises(RuntimeError):
        occ.points_in_boxes_gpu(torch.zeros(1, 4, 3), torch.zeros(1, 1, 7))


def test_oracle_ignores_extra_point_columns():
    """KITTI-format rows carry 6 columns; only xyz enters the path (occ_annotate.py:97)."""
    import numpy as np

    from objectcentricocccompletion_b200 import synth
    from oracle import oracle

    b = synth.make_batch(2, 10, 0.2, seed=9, small=True)
    exp = oracle.annotate_batch(b)
    for t in b.tracklets:
        t.points = [np.concatenate([p, np.full((len(p), 3), 7.0, np.float32)], 1) for p in t.points]
    got = oracle.annotate_batch(b)
    for e, g in zip(exp, got):
        assert e["status"] == g["status"] and (e["occ"] == g["occ"]).all()


def test_mean_var_oracle_matches_reference_fixture():
    """--save-mean-var (occ_annotate.py:627-645): the numpy restatement reproduces the grids recorded from the
    torch reference glue bit for bit (CPU accumulation order is the same)."""
    import os

    import numpy as np

    from oracle import oracle
    from tests.util import GOLDEN, load_golden

    gold = np.load(os.path.join(GOLDEN, "mean_var.npz"))
    seen = 0
    for name in ("annotate_small", "annotate_edge"):
        batch, _, status, _ = load_golden(name)
        mv = oracle.annotate_mean_var(batch)
        for t, m in enumerate(mv):
            key = f"{name}_{t}"
            assert (m is None) == (key not in gold)
            if m is not None:
                assert m.dtype == np.float32 and m.shape == gold[key].shape
                assert (m.view(np.uint32) == gold[key].view(np.uint32)).all()
                seen += 1
    assert seen >= 5


def test_mirror_occ_label_oracle_matches_reference_ops():
    """MirrorOccLabel (occ_pinelines.py:88-126): numpy restatement vs the reference's torch op sequence, copied
    as a sequence of torch calls (the pipeline class itself needs mmdet's registry to import)."""
    import numpy as np
    import torch

    from oracle import oracle

    rng = np.random.default_rng(0)
    for shape in [(11, 24, 9), (12, 23, 10), (1, 3, 2), (2, 2, 2)]:
        g = torch.from_numpy(rng.integers(0, 3, shape).astype(np.int32))
        XS, YS, ZS = g.shape
        flat = g.clone().view(-1)
        unknown = flat == 0
        mid = XS // 2
        vx, vy, vz = torch.meshgrid(torch.arange(XS), torch.arange(YS), torch.arange(ZS), indexing="ij")
        mxx = ((vx + 0.5 - mid) * -1.0 + mid).long()
        coors = torch.stack([mxx, vy, vz], -1).view(-1, 3)
        flat[unknown] = g[coors[unknown][:, 0], coors[unknown][:, 1], coors[unknown][:, 2]]
        assert (oracle.mirror_occ_label(g.numpy()) == flat.view(XS, YS, ZS).numpy()).all()


def test_brick_cull_rule_is_safe_and_useful():
    """docs/ROUND2_BRICK_CULL.md section 2 (round-2 design, CPU restatement only): a (brick, frame, LiDAR) triple the
    rule masks never contains a voxel centre whose reference visibility test succeeds, and the rule removes a
    sizeable share of the tests of the voxels that stay unknown."""
    import numpy as np

    from objectcentricocccompletion_b200 import synth
    from oracle import brick_cull, oracle

    batch = synth.make_batch(2, 10, 0.2, seed=4, small=True)
    res = oracle.annotate_batch(batch)
    unknown_tests = culled_tests = 0
    for t, r in enumerate(res):
        assert r["status"] == "ok"
        vox, rows, cols, rng, free = brick_cull.centre_tests(batch, t, r)
        for tile in [(1, 1), (2, 8)]:
            bricks, masked = brick_cull.brick_cull_masks(batch, t, r, brick=4, tile=tile)
            nby, nbz = int(bricks[:, 1].max()) + 1, int(bricks[:, 2].max()) + 1
            b_of_vox = ((vox[:, 0] // 4) * nby + vox[:, 1] // 4) * nbz + vox[:, 2] // 4
            m = masked[b_of_vox].transpose(1, 2, 0)                      # [B, L, n] like `free`
            assert not (m & free).any(), "a masked triple holds a centre the reference would free"
        still_unknown = r["occ"][vox[:, 0], vox[:, 1], vox[:, 2]] == 0
        unknown_tests += int(still_unknown.sum()) * m.shape[0] * m.shape[1]
        culled_tests += int(m[:, :, still_unknown].sum())
    assert culled_tests > 0.3 * unknown_tests, (culled_tests, unknown_tests)
