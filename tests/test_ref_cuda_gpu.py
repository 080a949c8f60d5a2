"""GPU tests against the REFERENCE'S OWN CUDA kernels (unmodified sources compiled for sm_100a by
oracle/build.py::build_ref_cuda into oracle/_ref; skipped when those prebuilt files are absent)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ext(name):
    from oracle import build as obuild

    m = obuild.load_ref(name)
    if m is None:
        pytest.skip(f"oracle/_ref/{name} not built")
    return m


def _t(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_voxelize_vs_reference_cuda():
    import torch

    import objectcentricocccompletion_b200 as occ

    ref = _ext("ref_voxel_layer_cuda")
    rng = np.random.default_rng(0)
    N, C = 200000, 5
    pts = np.concatenate([rng.uniform(-220, 220, (N, 2)), rng.uniform(-6, 10, (N, 1)), rng.random((N, C - 3))], 1).astype(np.float32)
    p = _t(pts)
    vs, pcr = [0.2, 0.2, 0.2], [-204.8, -204.8, -4, 204.8, 204.8, 8]
    exp = torch.zeros((N, 3), dtype=torch.int32, device="cuda")
    ref.dynamic_voxelize(p, exp, vs, pcr, 3)
    got = occ.voxelization(p, vs, pcr, -1, -1)
    assert (got == exp).all()
    # hard voxelization
    pts2 = np.concatenate([rng.uniform(-5, 75, (60000, 1)), rng.uniform(-45, 45, (60000, 1)), rng.uniform(-4, 2, (60000, 1)), rng.random((60000, 1))], 1).astype(np.float32)
    p2 = _t(pts2)
    vs2, pcr2, mp, mv = [0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1], 8, 5000
    voxels = torch.zeros((mv, mp, 4), device="cuda"); coors = torch.zeros((mv, 3), dtype=torch.int32, device="cuda")
    num = torch.zeros((mv,), dtype=torch.int32, device="cuda")
    m = ref.hard_voxelize(p2, voxels, coors, num, vs2, pcr2, mp, mv, 3)
    gv, gc, gn = occ.voxelization(p2, vs2, pcr2, mp, mv)
    assert gc.shape[0] == m and (gc == coors[:m]).all() and (gn == num[:m]).all() and (gv == voxels[:m]).all()


@pytest.mark.parametrize("mode", ["mean", "max", "sum"])
def test_dynamic_scatter_vs_reference_cuda(mode):
    import torch

    from objectcentricocccompletion_b200 import voxel

    ref = _ext("ref_voxel_layer_cuda")
    rng = np.random.default_rng(1)
    N, C = 100000, 16
    feats = _t(rng.standard_normal((N, C)).astype(np.float32))
    for lowv in (-1, 0):
        coors = _t(rng.integers(lowv, 25, (N, 3)).astype(np.int32))
        ef, ec, em, en = ref.dynamic_point_to_voxel_forward(feats, coors, mode)
        gf, gc, gm, gn, _ = voxel.dynamic_point_to_voxel_forward(feats, coors, mode)
        assert gc.shape == ec.shape and (gc == ec).all() and (gm == em).all() and (gn == en).all()
        assert torch.allclose(gf, ef, rtol=1e-5, atol=1e-5)
        g = _t(rng.standard_normal(tuple(ef.shape)).astype(np.float32))
        eg = torch.zeros_like(feats); gg = torch.zeros_like(feats)
        ref.dynamic_point_to_voxel_backward(eg, g, feats, ef, em, en, mode)
        voxel.dynamic_point_to_voxel_backward(gg, g, feats, gf, gm, gn, mode)
        assert torch.allclose(gg, eg, rtol=1e-5, atol=1e-6)


def test_points_in_boxes_vs_reference_cuda():
    import torch

    import objectcentricocccompletion_b200 as occ

    ref = _ext("ref_points_in_boxes_cuda")
    rng = np.random.default_rng(2)
    B, T, M = 2, 12, 300000
    boxes = np.concatenate([rng.uniform(-20, 20, (B, T, 3)), rng.uniform(1, 8, (B, T, 3)), rng.uniform(-4, 4, (B, T, 1))], 2).astype(np.float32)
    pts = (boxes[np.arange(B)[:, None], rng.integers(0, T, (B, M)), :3] + rng.normal(0, 3, (B, M, 3))).astype(np.float32)
    b, p = _t(boxes), _t(pts)
    exp = torch.full((B, M), -1, dtype=torch.int32, device="cuda")
    ref.points_in_boxes_gpu(b, p, exp)
    got = occ.points_in_boxes_gpu(p, b)
    # default arithmetic = the reference CUDA kernel's: device cosf/sinf and nvcc's FMA contraction of local_x / local_y
    assert (got == exp).all()
    expb = torch.zeros((B, M, T), dtype=torch.int32, device="cuda")
    ref.points_in_boxes_batch(b, p, expb)
    gotb = occ.points_in_boxes_batch(p, b)
    assert (gotb == expb).all()
    # the CPU arithmetic (host libm trig, unfused products) differs from it only for points within an ulp of a face
    cpu = occ.points_in_boxes_gpu(p, b, host_trig=True)
    assert 0 <= (cpu != exp).float().mean().item() < 1e-5
