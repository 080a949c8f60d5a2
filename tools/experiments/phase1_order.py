#!/usr/bin/env python
"""How much of the freeable volume do the first 8 surviving pairs free, for different pair orders?  Uses the
oracle's exact per-voxel results (oracle/brick_cull.py::centre_tests).  Phase 1 of the ray-cast runs the first 8
surviving pairs on every voxel; what it does not free goes to phase 2.  CPU only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from objectcentricocccompletion_b200 import synth  # noqa: E402
from oracle import brick_cull, oracle  # noqa: E402


def strided(B, S=8):
    return [r + j * S for r in range(S) for j in range((B - r + S - 1) // S)]


def main(n_trk=3, n_frames=40, seed=0, P1=8):
    batch = synth.make_batch(n_trk, n_frames, 0.2, seed=seed)
    res = oracle.annotate_batch(batch, threads=8)
    out = {}
    for t, r in enumerate(res):
        if r["occ"] is None:
            continue
        vox, rows, cols, rng, free = brick_cull.centre_tests(batch, t, r)
        B, L, n = free.shape
        freeable = free.any((0, 1))
        alive = free.any(2)                                   # pairs that free at least one voxel (a lower bound on
        orders = {                                            # what survives the pair cull)
            "frames 0..B-1, LiDARs in order (round 1a)": [(i, c) for i in range(B) for c in range(L)],
            "frames strided by 8 (current)": [(i, c) for i in strided(B) for c in range(L)],
            "TOP LiDAR of strided frames first": [(i, 0) for i in strided(B)] + [(i, c) for i in strided(B) for c in range(1, L)],
        }
        for name, order in orders.items():
            surv = [pc for pc in order if alive[pc]][:P1]
            got = np.zeros(n, bool)
            for i, c in surv:
                got |= free[i, c]
            a, b = out.get(name, (0, 0))
            out[name] = (a + int(got.sum()), b + int(freeable.sum()))
        # greedy upper bound
        got = np.zeros(n, bool)
        for _ in range(P1):
            gains = (free & ~got[None, None]).sum(2)
            i, c = np.unravel_index(np.argmax(gains), gains.shape)
            got |= free[i, c]
        a, b = out.get("greedy (upper bound)", (0, 0))
        out["greedy (upper bound)"] = (a + int(got.sum()), b + int(freeable.sum()))
    for name, (a, b) in out.items():
        print(f"  {name:48s} frees {100.0 * a / b:5.1f} % of the freeable voxels in the first {P1} surviving pairs")


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
