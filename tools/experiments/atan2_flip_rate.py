#!/usr/bin/env python
"""How often does CUDA's f64 atan2 (the exact recheck, geom.cuh::project_exact) take a different row / column
decision than glibc's (the oracle = what the reference's torch-CPU run computes)?  DESIGN.md 1.1 (c) argues
~1e-13 per test from the 2-ulp bound; this measures it: N random ego-frame points through the TOP LiDAR of the
synthetic rig, row / column / range from occb200_point_cloud_to_range_image_idx (device) against the C oracle
(host, one thread per chunk).
    python tools/experiments/atan2_flip_rate.py [N=200000000]"""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import objectcentricocccompletion_b200 as occ  # noqa: E402
from objectcentricocccompletion_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
CH = 4_000_000                                     # points per device call
TH = min(16, os.cpu_count() or 1)
seg = synth.make_batch(1, 10, 0.2, seed=3).segments[0]
done = rows = cols = rngs = 0
t0 = time.time()
g = torch.Generator(device="cuda").manual_seed(1)
with ThreadPoolExecutor(TH) as pool:
    while done < N:
        n = min(CH, N - done)
        lidar = (done // CH) % len(seg.inclinations)          # all five LiDARs of the rig in turn
        frame = (done // CH) % seg.extrinsics.shape[0]
        E = np.ascontiguousarray(seg.extrinsics[frame, lidar][None], np.float32)
        incl = np.ascontiguousarray(seg.inclinations[lidar][::-1], np.float32)[None]
        H, W = seg.range_images[lidar][0].shape
        # ego-frame points: 3 .. 75 m away, heights -3 .. 4 m (what voxel centres of tracked objects look like)
        r = 3.0 + 72.0 * torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
        a = 6.283185307179586 * torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
        z = -3.0 + 7.0 * torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
        pts = torch.stack([r * torch.cos(a), r * torch.sin(a), z], -1)[None].contiguous()
        idx, rng = occ.point_cloud_to_range_image_idx(pts, torch.from_numpy(E), torch.from_numpy(incl), (H, W))
        ph = pts.cpu().numpy()
        parts = np.array_split(np.arange(n), TH)
        res = list(pool.map(lambda p: oracle.point_cloud_to_range_image_idx(ph[:, p[0]: p[-1] + 1], E, incl, (H, W)), parts))
        ridx = np.concatenate([x[0] for x in res], 1)
        rrng = np.concatenate([x[1] for x in res], 1)
        gi, gr = idx.cpu().numpy(), rng.cpu().numpy()
        rows += int((gi[..., 0] != ridx[..., 0]).sum())
        cols += int((gi[..., 1] != ridx[..., 1]).sum())
        rngs += int((gr != rrng).sum())
        done += n
print(f"{done} tests in {time.time() - t0:.0f} s: row flips {rows}, column flips {cols}, range values that differ {rngs}")
print(f"flip rate <= {3.0 / done:.1e} per test at 95 % confidence" if rows + cols == 0 else
      f"flip rate {(rows + cols) / done:.2e} per test")
