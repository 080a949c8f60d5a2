#!/usr/bin/env python
"""BASELINE.json config 4: OcCo-Net input build -- dynamic Voxelization + DynamicScatter mean/max (+ scatter_v2)
over 32 tracklets of aggregated points (N ~ 1.05 M).  Prints one JSON object per operator:
device time (CUDA events, L2 flushed between iterations), algorithmic bytes,
fraction of the measured HBM peak, the torch-op sequence the reference runs on the same GPU
(torch.unique(dim=0) + index_add_/index_reduce_), and the CPU oracle time for voxelize."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import objectcentricocccompletion_b200 as occ  # noqa: E402
from objectcentricocccompletion_b200 import synth  # noqa: E402


def timed(fn, flush, iters=20, warm=3):
    for _ in range(warm):
        fn()
    evs = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs]))


def _ref(name):
    from oracle import build as obuild

    return obuild.load_ref(name)


def main():
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pts, bidx = synth.scatter_inputs(32, 32, 1024, 5, seed=0)
    N = pts.shape[0]
    p = torch.from_numpy(pts).to(dev)
    b = torch.from_numpy(bidx).to(dev)
    vs, pcr = [0.2, 0.2, 0.2], [-204.8, -204.8, -4, 204.8, 204.8, 8]
    out = []
    vox = occ.Voxelization(vs, pcr, -1)
    ms = timed(lambda: vox(p), flush)
    alg = 4 * 5 * N + 12 * N
    coors = vox(p)
    from oracle import oracle
    t0 = time.perf_counter(); oracle.dynamic_voxelize(pts, vs, pcr); cpu_ms = (time.perf_counter() - t0) * 1e3
    rv = _ref("ref_voxel_layer_cuda")
    ref_ms = None
    if rv is not None:
        rc = torch.zeros((N, 3), dtype=torch.int32, device=dev)
        ref_ms = timed(lambda: rv.dynamic_voxelize(p, rc, vs, pcr, 3), flush)
        assert (rc == coors).all()
    out.append(dict(op="Voxelization(dynamic)", N=N, C=5, ms=ms, alg_bytes=alg, gbs=alg / ms / 1e6, frac=alg / ms / 1e6 / peak,
                    cpu_oracle_ms=cpu_ms, reference_cuda_kernel_sm100a_ms=ref_ms))
    # points_in_boxes: one frame of 180k points against 64 boxes (what occ_annotate streams per tracklet-frame)
    rp = _ref("ref_points_in_boxes_cuda")
    rng = np.random.default_rng(0)
    M, T = 180000, 64
    boxes = torch.from_numpy(np.concatenate([rng.uniform(-60, 60, (1, T, 2)), rng.uniform(-1, 1, (1, T, 1)), rng.uniform(1.5, 6, (1, T, 3)),
                                             rng.uniform(-3, 3, (1, T, 1))], 2).astype(np.float32)).to(dev)
    fpts = torch.from_numpy(np.concatenate([rng.uniform(-75, 75, (1, M, 2)), rng.uniform(-2, 4, (1, M, 1))], 2).astype(np.float32)).to(dev)
    ms = timed(lambda: occ.points_in_boxes_gpu(fpts, boxes), flush)
    ref_ms = None
    if rp is not None:
        ro = torch.full((1, M), -1, dtype=torch.int32, device=dev)
        ref_ms = timed(lambda: rp.points_in_boxes_gpu(boxes, fpts, ro), flush)
    out.append(dict(op="points_in_boxes_gpu", M=M, T=T, ms=ms, alg_bytes=16 * M, gbs=16 * M / ms / 1e6,
                    reference_cuda_kernel_sm100a_ms=ref_ms))
    coors4 = torch.cat([b[:, None].int(), coors], 1).contiguous()
    for C, mode in [(3, "mean"), (5, "mean"), (128, "max")]:
        f = (p[:, :C].contiguous() if C <= 5 else torch.randn(N, C, device=dev))
        ds = occ.DynamicScatter(vs, pcr, mode == "mean")
        ms = timed(lambda: ds(f, coors4), flush)
        vf, vc = ds(f, coors4)
        M = vf.shape[0]
        alg = 4 * C * N + 16 * N + (4 * C + 16) * M

        def ref():
            u, inv = torch.unique(coors4, dim=0, return_inverse=True)
            if mode == "mean":
                o = torch.zeros((u.shape[0], C), device=dev).index_add_(0, inv, f)
                cnt = torch.zeros(u.shape[0], device=dev).index_add_(0, inv, torch.ones(N, device=dev))
                return o / cnt[:, None]
            return torch.full((u.shape[0], C), -float("inf"), device=dev).index_reduce_(0, inv, f, "amax")
        ref_ms = timed(ref, flush, iters=5, warm=1)
        rk_ms = None
        if rv is not None:      # the reference's own kernel takes 3-column coords: one sample's worth is the whole set here
            c3 = coors.contiguous()
            rk_ms = timed(lambda: rv.dynamic_point_to_voxel_forward(f, c3, mode), flush, iters=5, warm=1)
            ms3c = timed(lambda: occ.dynamic_scatter(f, c3, mode), flush)
        out.append(dict(op=f"DynamicScatter({mode})", N=N, C=C, M=M, ms=ms, alg_bytes=alg, gbs=alg / ms / 1e6,
                        frac=alg / ms / 1e6 / peak, torch_ops_same_gpu_ms=ref_ms,
                        ms_3col=ms3c if rv is not None else None, reference_cuda_kernel_sm100a_3col_ms=rk_ms))
        c64 = coors4.long()
        ms2 = timed(lambda: occ.scatter_v2(f, c64, mode), flush)
        nf, nc, inv = occ.scatter_v2(f, c64, mode)
        ms3 = timed(lambda: occ.scatter_v2(f, c64, mode, unq_inv=inv, new_coors=nc), flush)
        out.append(dict(op=f"scatter_v2({mode})", N=N, C=C, M=int(nc.shape[0]), ms=ms2, ms_reusing_unq_inv=ms3,
                        alg_bytes=alg + 16 * N, gbs=(alg + 16 * N) / ms2 / 1e6))
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
