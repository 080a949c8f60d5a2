"""GPU parity tests of the operator surface (-m gpu): points_in_boxes, Voxelization, DynamicScatter,
scatter_v2, occ_ops -- CUDA through the C ABI vs the CPU oracle and the reference's golden vectors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _t(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ---------------------------------------------------------------- points_in_boxes
def test_points_in_boxes_reference_golden():
    """tests/test_models/test_common_modules/test_roiaware_pool3d.py:43-64, 97-120"""
    import objectcentricocccompletion_b200 as occ
    from tests.test_oracle_cpu import BOXES, PTS15

    boxes = _t(BOXES[:, None, :])
    pts = _t(np.stack([PTS15[:8], np.array([[3.8, 7.6, -2], [-10.6, -12.9, -20], [-16, -18, 9], [-21.3, -52, -5],
                                            [0, 0, 0], [6, 7, 8], [-2, -3, -4], [6, 4, 9]], np.float32)], 0))
    exp = np.array([[0, 0, 0, 0, 0, -1, -1, -1], [-1] * 8], np.int32)
    for host_trig in (False, True):
        out = occ.points_in_boxes_gpu(points=pts, boxes=boxes, host_trig=host_trig)
        assert out.shape == (2, 8) and out.dtype.is_floating_point is False
        assert (out.cpu().numpy() == exp).all()
    expb = np.array([[[1, 0], [1, 0], [1, 0], [1, 0], [1, 0], [0, 1]] + [[0, 0]] * 9], np.int32)
    outb = occ.points_in_boxes_batch(points=_t(PTS15[None]), boxes=_t(BOXES[None]))
    assert outb.shape == (1, 15, 2) and (outb.cpu().numpy() == expb).all()


@pytest.mark.parametrize("M", [1, 7, 1000, 100003])
def test_points_in_boxes_vs_oracle(M):
    import objectcentricocccompletion_b200 as occ
    from oracle import oracle

    rng = np.random.default_rng(M)
    B, T = 3, 9
    boxes = np.concatenate([rng.uniform(-20, 20, (B, T, 3)), rng.uniform(1, 8, (B, T, 3)), rng.uniform(-4, 4, (B, T, 1))], 2).astype(np.float32)
    pts = (boxes[np.arange(B)[:, None], rng.integers(0, T, (B, M)), :3] + rng.normal(0, 3, (B, M, 3))).astype(np.float32)
    exp = oracle.points_in_boxes_gpu(pts, boxes)
    got = occ.points_in_boxes_gpu(_t(pts), _t(boxes), host_trig=True).cpu().numpy()
    assert (got == exp).all()                                   # host libm trig: bit-identical to the CPU reference
    dev = occ.points_in_boxes_gpu(_t(pts), _t(boxes)).cpu().numpy()
    assert (dev != exp).mean() < 1e-4                           # device cosf/sinf may move a boundary by 1 ulp
    expb = oracle.points_in_boxes_batch(pts, boxes)
    gotb = occ.points_in_boxes_batch(_t(pts), _t(boxes), host_trig=True).cpu().numpy()
    assert (gotb == expb).all()


# ---------------------------------------------------------------- Voxelization
@pytest.mark.parametrize("C,N", [(3, 1), (4, 1000), (5, 100001), (6, 4097), (20, 300)])
def test_dynamic_voxelize_vs_oracle(C, N):
    import torch

    import objectcentricocccompletion_b200 as occ
    from oracle import oracle

    rng = np.random.default_rng(C * 1000 + N)
    pts = np.concatenate([rng.uniform(-220, 220, (N, 2)), rng.uniform(-6, 10, (N, 1)), rng.random((N, C - 3))], 1).astype(np.float32)
    vs, pcr = [0.2, 0.2, 0.2], [-204.8, -204.8, -4, 204.8, 204.8, 8]          # configs/ococc/ococcnet.py:8-9
    exp = oracle.dynamic_voxelize(pts, vs, pcr)
    vox = occ.Voxelization(vs, pcr, -1)
    got = vox(_t(pts))
    assert got.dtype == torch.int32 and got.shape == (N, 3)
    assert (got.cpu().numpy() == exp).all()
    # odd alignment (storage offset) takes the unstaged path
    base = torch.zeros(N * C + 1, device="cuda")
    base[1:] = _t(pts).reshape(-1)
    got2 = vox(base[1:].view(N, C))
    assert (got2.cpu().numpy() == exp).all()
    got64 = occ.voxelization(_t(pts.astype(np.float64)), vs, pcr, -1, -1)
    exp64 = np.clip(np.floor((pts[:, :3].astype(np.float64) - np.array(pcr[:3], np.float32)) / np.array(vs, np.float32)),
                    0, np.array([2048, 2048, 60]) - 1).astype(np.int32)[:, ::-1]
    assert (got64.cpu().numpy() == exp64).all()


def test_dynamic_voxelize_empty():
    import torch

    import objectcentricocccompletion_b200 as occ

    out = occ.voxelization(torch.zeros((0, 4), device="cuda"), [0.2] * 3, [0, 0, 0, 1, 1, 1], -1, -1)
    assert out.shape == (0, 3)


@pytest.mark.parametrize("mp,mv", [(8, 5000), (3, 50), (1000, 20000)])
def test_hard_voxelize_vs_oracle(mp, mv):
    import objectcentricocccompletion_b200 as occ
    from oracle import oracle

    rng = np.random.default_rng(mp)
    N, C = 60000, 4
    pts = np.concatenate([rng.uniform(-5, 75, (N, 1)), rng.uniform(-45, 45, (N, 1)), rng.uniform(-4, 2, (N, 1)), rng.random((N, 1))], 1).astype(np.float32)
    vs, pcr = [0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1]
    ev, ec, en = oracle.hard_voxelize(pts, vs, pcr, mp, mv)
    gv, gc, gn = occ.Voxelization(vs, pcr, mp, mv).eval()(_t(pts))
    assert gc.shape == ec.shape and (gc.cpu().numpy() == ec).all()
    assert (gn.cpu().numpy() == en).all() and (gv.cpu().numpy() == ev).all()


def test_hard_voxelize_kat():
    """tests/test_models/test_voxel_encoder/test_voxel_generator.py:6-22 through the CUDA path."""
    import objectcentricocccompletion_b200 as occ

    np.random.seed(0)
    points = np.random.rand(1000, 4).astype(np.float32)
    v, c, n = occ.voxelization(_t(points), [0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1], 1000, 20000)
    assert (c.cpu().numpy() == np.array([[7, 81, 1], [6, 81, 0], [7, 80, 1], [6, 81, 1], [7, 81, 0], [6, 80, 1], [7, 80, 0], [6, 80, 0]])).all()
    assert (n.cpu().numpy() == np.array([120, 121, 127, 134, 115, 127, 125, 131])).all()


# ---------------------------------------------------------------- DynamicScatter
def test_dynamic_scatter_reference_test():
    """tests/test_models/test_voxel_encoder/test_dynamic_scatter.py:8-93, ported to this package."""
    import torch
    from torch.autograd import gradcheck

    from objectcentricocccompletion_b200 import DynamicScatter

    torch.manual_seed(0)
    feats = torch.rand(size=(200000, 3), dtype=torch.float32, device='cuda') * 100 - 50
    coors = torch.randint(low=-1, high=20, size=(200000, 3), dtype=torch.int32, device='cuda')
    dsmean = DynamicScatter([0.32, 0.32, 6], [-74.88, -74.88, -2, 74.88, 74.88, 4], True)
    dsmax = DynamicScatter([0.32, 0.32, 6], [-74.88, -74.88, -2, 74.88, 74.88, 4], False)
    empty_feats = torch.empty(size=(0, 3), dtype=torch.float32, device='cuda').requires_grad_()
    empty_coors = torch.empty(size=(0, 3), dtype=torch.int32, device='cuda')
    for ds in (dsmean, dsmax):
        f, c = ds(empty_feats, empty_coors)
        f.sum().backward()
        assert f.shape == empty_feats.shape and c.shape == empty_coors.shape
    eo_feats = (torch.rand(size=(200000, 3), dtype=torch.float32, device='cuda') * 100 - 50).requires_grad_()
    eo_coors = torch.randint(low=-1, high=0, size=(200000, 3), dtype=torch.int32, device='cuda')
    for ds in (dsmean, dsmax):
        f, c = ds(eo_feats, eo_coors)
        f.sum().backward()
        assert f.shape[0] == 0 and (eo_feats.grad == 0).all()
    ref_voxel_coors = coors.unique(dim=0, sorted=True)
    ref_voxel_coors = ref_voxel_coors[ref_voxel_coors.min(dim=-1).values >= 0]
    fm, cm = dsmean(feats, coors)
    fx, cx = dsmax(feats, coors)
    assert (cm == ref_voxel_coors).all() and (cx == ref_voxel_coors).all()     # already in sorted order
    inv = {tuple(c): i for i, c in enumerate(ref_voxel_coors.cpu().numpy().tolist())}
    cf, cc = feats.cpu().double().numpy(), coors.cpu().numpy()
    key = (cc[:, 0] * 400 + cc[:, 1] * 20 + cc[:, 2])
    for s in np.random.default_rng(0).choice(len(ref_voxel_coors), 200, replace=False):
        c = ref_voxel_coors[s].cpu().numpy()
        m = (key == c[0] * 400 + c[1] * 20 + c[2]) & (cc.min(-1) >= 0)
        assert np.allclose(fm[s].cpu().numpy(), cf[m].mean(0), atol=1e-2, rtol=1e-5)
        assert np.allclose(fx[s].cpu().numpy(), cf[m].max(0), atol=1e-2, rtol=1e-5)
    feats = (torch.rand(size=(100, 4), dtype=torch.float32, device='cuda') * 100 - 50).requires_grad_()
    coors = torch.randint(low=-1, high=3, size=(100, 3), dtype=torch.int32, device='cuda')
    gradcheck(dsmean, (feats, coors), eps=1e-2, atol=1e-2, rtol=1e-5)
    gradcheck(dsmax, (feats, coors), eps=1e-2, atol=1e-2, rtol=1e-5)


@pytest.mark.parametrize("C", [3, 5, 16, 128])
@pytest.mark.parametrize("mode", ["mean", "max", "sum"])
def test_dynamic_scatter_vs_oracle(C, mode):
    """Forward (coords/map/count exact incl. the drop-first quirk, features 1e-5 rel) and backward."""
    import torch

    from objectcentricocccompletion_b200 import voxel
    from oracle import oracle

    rng = np.random.default_rng(C)
    N = 30011
    feats = rng.standard_normal((N, C)).astype(np.float32)
    for lowv in (-1, 0):                                           # with and without invalid rows
        coors = rng.integers(lowv, 14, (N, 3)).astype(np.int32)
        evf, evc, emp, ecnt = oracle.dynamic_scatter_fwd(feats, coors, mode)
        f = _t(feats).requires_grad_()
        gvf, gvc, gmp, gcnt, _ = voxel.dynamic_point_to_voxel_forward(f.detach(), _t(coors), mode)
        assert (gvc.cpu().numpy() == evc).all() and (gmp.cpu().numpy() == emp).all() and (gcnt.cpu().numpy() == ecnt).all()
        assert np.allclose(gvf.cpu().numpy(), evf, rtol=1e-5, atol=1e-6)
        out, _ = voxel.dynamic_scatter(f, _t(coors), mode)
        g = rng.standard_normal(evf.shape).astype(np.float32)
        out.backward(_t(g))
        eg = oracle.dynamic_scatter_bwd(g, feats, evf, emp, ecnt, mode)
        assert np.allclose(f.grad.cpu().numpy(), eg, rtol=1e-5, atol=1e-7)
        # reference-signature backward without the saved argmax
        g2 = torch.zeros_like(f)
        voxel.dynamic_point_to_voxel_backward(g2, _t(g), f.detach(), gvf, gmp, gcnt, mode)
        assert np.allclose(g2.cpu().numpy(), eg, rtol=1e-5, atol=1e-7)


def test_dynamic_scatter_batched_vs_oracle():
    """4-column coords: per-sample drop-first (scatter_points.py:83-99) in one sort."""
    from objectcentricocccompletion_b200 import DynamicScatter
    from oracle import oracle

    rng = np.random.default_rng(4)
    N, C = 50000, 5
    feats = rng.standard_normal((N, C)).astype(np.float32)
    b = np.sort(rng.integers(0, 6, N)).astype(np.int32)
    b[b == 3] = 4                                                   # an empty sample in the middle
    coors = np.concatenate([b[:, None], rng.integers(-1, 10, (N, 3)).astype(np.int32)], 1)
    coors[b == 2, 1:] = np.abs(coors[b == 2, 1:])                   # sample 2 has no invalid row -> its first voxel is dropped
    for avg in (True, False):
        ef, ec = oracle.dynamic_scatter_batched(feats, coors, "mean" if avg else "max")
        gf, gc = DynamicScatter([0.2] * 3, [0] * 6, avg)(_t(feats), _t(coors))
        assert gc.shape == ec.shape and (gc.cpu().numpy() == ec).all()
        assert np.allclose(gf.cpu().numpy(), ef, rtol=1e-5, atol=1e-6)


def test_unknown_reduce_type():
    import torch

    from objectcentricocccompletion_b200 import dynamic_scatter

    with pytest.raises(RuntimeError):
        dynamic_scatter(torch.zeros(4, 3, device="cuda"), torch.zeros(4, 3, dtype=torch.int32, device="cuda"), "median")


# ---------------------------------------------------------------- scatter_v2
@pytest.mark.parametrize("K,C", [(3, 3), (4, 128), (5, 6)])
def test_scatter_v2_vs_oracle(K, C):
    import torch

    from objectcentricocccompletion_b200 import scatter_v2
    from oracle import oracle

    rng = np.random.default_rng(K * 10 + C)
    N = 40009
    feats = rng.standard_normal((N, C)).astype(np.float32)
    coors = rng.integers(-3, 9, (N, K)).astype(np.int64)
    for mode in ("avg", "max", "sum"):
        ef, ec, ei = oracle.scatter_v2(feats, coors, mode)
        f = _t(feats).requires_grad_()
        gf, gc, gi = scatter_v2(f, _t(coors), mode)
        assert gc.dtype == torch.int64 and gi.dtype == torch.int64
        assert (gc.cpu().numpy() == ec).all() and (gi.cpu().numpy() == ei).all()
        assert np.allclose(gf.detach().cpu().numpy(), ef, rtol=1e-5, atol=1e-6)
        # reuse of unq_inv / new_coors (sir.py:70 unique_once) with and without the cached plan
        gf2, _, _ = scatter_v2(f, _t(coors), mode, unq_inv=gi, new_coors=gc)
        gf3, _, _ = scatter_v2(f, _t(coors), mode, unq_inv=gi.clone(), new_coors=gc)
        assert (gf2 == gf).all() and (gf3 == gf).all()
        gf.sum().backward()
        cnt = np.bincount(ei, minlength=len(ec)).astype(np.float32)
        if mode == "sum":
            assert np.allclose(f.grad.cpu().numpy(), 1.0)
        elif mode == "avg":
            assert np.allclose(f.grad.cpu().numpy(), (1.0 / cnt[ei])[:, None] * np.ones((1, C)), rtol=1e-6)
        else:
            assert np.allclose(f.grad.sum(0).cpu().numpy(), len(ec))       # one winner per (voxel, channel)
    ef, ec, ei = oracle.scatter_v2(feats, coors, "mean", min_points=3)
    gf, gc, gi = scatter_v2(_t(feats), _t(coors), "mean", min_points=3)
    assert (gc.cpu().numpy() == ec).all() and (gi.cpu().numpy() == ei).all()
    assert np.allclose(gf.cpu().numpy(), ef, rtol=1e-5, atol=1e-6)
    r = scatter_v2(_t(feats), _t(coors), "mean", return_inv=False)
    assert len(r) == 2
    with pytest.raises(NotImplementedError):
        scatter_v2(_t(feats), _t(coors), "median")


# ---------------------------------------------------------------- occ_ops
def test_occ_ops_vs_oracle():
    import objectcentricocccompletion_b200 as occ
    from oracle import oracle

    rng = np.random.default_rng(8)
    R, N = 7, 20000
    rois = np.concatenate([rng.integers(0, 2, (R, 1)), rng.uniform(-30, 30, (R, 3)), rng.uniform(1.5, 6, (R, 3)), rng.uniform(-3, 3, (R, 3))], 1).astype(np.float32)
    idx = rng.integers(0, R, N)
    pts = ((rng.random((N, 3)) - 0.5) * rois[idx, 4:7] * 1.3).astype(np.float32)
    sc, of = [1.1, 1.2, 1.0], [0.4, 0.4, 0.2]
    for to_center in (False, True):
        e = oracle.quantize_points(pts, rois, idx, 0.2, sc, of, to_center)
        g = occ.quantize_points(_t(pts), _t(rois), _t(idx), 0.2, sc, of, to_center).cpu().numpy()
        assert g.dtype == e.dtype and (g == e).all()
        if not to_center:
            e_coor = e
    # PyTorch index semantics of rois_points_idx (occ_ops.py:81): negative counts from the end, out of range raises
    neg = idx.copy()
    neg[::3] -= R
    g = occ.quantize_points(_t(pts), _t(rois), _t(neg), 0.2, sc, of).cpu().numpy()
    assert (g == oracle.quantize_points(pts, rois, idx, 0.2, sc, of)).all()
    bad = idx.copy()
    bad[5] = R
    with pytest.raises(IndexError):
        occ.quantize_points(_t(pts), _t(rois), _t(bad), 0.2, sc, of)
    g = occ.quantize_points(_t(pts), _t(rois), _t(bad), 0.2, sc, of, check_index=False).cpu().numpy()
    assert (g[5] == np.iinfo(np.int64).min).all() and (np.delete(g, 5, 0) == np.delete(e_coor, 5, 0)).all()
    for as_volume in (False, True):
        e = oracle.generate_dense_voxel_centers(rois[:, 4:7], 0.2, sc, of, as_volume)
        g = occ.generate_dense_voxel_centers(_t(rois[:, 4:7]), 0.2, sc, of, as_volume)
        assert len(e) == len(g)
        for a, b in zip(e, g):
            assert a.shape == tuple(b.shape) and (a == b.cpu().numpy()).all()


def test_mirror_occ_label_gpu():
    """occ_ops.mirror_occ_label vs the oracle on ragged grids (odd / even X, X = 1), one launch for the list."""
    import numpy as np
    import torch

    import objectcentricocccompletion_b200.occ_ops as occ_ops
    from oracle import oracle

    rng = np.random.default_rng(3)
    grids = [rng.integers(0, 3, s).astype(np.int32) for s in [(11, 24, 9), (12, 23, 10), (1, 3, 2), (29, 120, 35)]]
    got = occ_ops.mirror_occ_label([torch.from_numpy(g).cuda() for g in grids])
    for g, o in zip(grids, got):
        assert o.shape == g.shape and (o.cpu().numpy() == oracle.mirror_occ_label(g)).all()
    assert occ_ops.mirror_occ_label([]) == []


@pytest.mark.parametrize("mode", ["mean", "max", "sum"])
def test_dynamic_scatter_float64(mode):
    """float64 features (the reference dispatches AT_DISPATCH_FLOATING_TYPES, scatter_points_cuda.cu:215): forward
    against a numpy f64 evaluation, and gradcheck in double precision (test_dynamic_scatter.py:86-93)."""
    import torch
    from torch.autograd import gradcheck

    from objectcentricocccompletion_b200 import voxel

    rng = np.random.default_rng(5)
    N, C = 20011, 5
    feats = rng.standard_normal((N, C))
    coors = rng.integers(-1, 9, (N, 3)).astype(np.int32)
    vf, vc, mp, cnt, _ = voxel.dynamic_point_to_voxel_forward(_t(feats), _t(coors), mode)
    assert vf.dtype == torch.float64
    mp_h, vf_h = mp.cpu().numpy(), vf.cpu().numpy()
    for v in range(0, vf_h.shape[0], 7):
        sel = feats[mp_h == v]
        want = {"mean": sel.mean(0), "max": sel.max(0), "sum": sel.sum(0)}[mode]
        np.testing.assert_allclose(vf_h[v], want, rtol=1e-12, atol=1e-12)
    f = _t(rng.standard_normal((60, 3))).requires_grad_()
    c = _t(rng.integers(-1, 3, (60, 3)).astype(np.int32))
    assert gradcheck(lambda x: voxel.dynamic_scatter(x, c, mode)[0], (f,), eps=1e-6, atol=1e-6, rtol=1e-4)


@pytest.mark.parametrize("mode", [1, 2])
def test_dynamic_scatter_bounded_keys_equal_unbounded(mode):
    """occb200_unique_rows_bounded: with the grid's bounds the sort keys are laid out without the min/max pass; the
    result equals the unbounded call, also when a row lies beyond a bound (device flag -> unbounded repeat)."""
    import torch

    from objectcentricocccompletion_b200.voxel import _dynamic_scatter

    torch.manual_seed(3)
    N = 150_000
    grid = (40, 300, 280)                                  # (z, y, x)
    cols = [torch.randint(-1, g, (N,), dtype=torch.int32, device='cuda') for g in grid]
    if mode == 2:
        cols = [torch.sort(torch.randint(0, 4, (N,), dtype=torch.int32, device='cuda')).values] + cols
    coors = torch.stack(cols, 1).contiguous()
    feats = torch.rand((N, 5), device='cuda')
    cm = [g - 1 for g in grid] if mode == 1 else [0] + [g - 1 for g in grid]
    for red in ('mean', 'max'):
        a = _dynamic_scatter.apply(feats, coors, red, mode, None)
        b = _dynamic_scatter.apply(feats, coors, red, mode, cm)
        assert a[1].shape == b[1].shape and bool((a[1] == b[1]).all()) and bool((a[0] == b[0]).all())
        tight = [max(c // 2, 0) for c in cm]               # rows beyond these bounds exist: the fallback path
        c = _dynamic_scatter.apply(feats, coors, red, mode, tight)
        assert a[1].shape == c[1].shape and bool((a[1] == c[1]).all()) and bool((a[0] == c[0]).all())
