#!/usr/bin/env python
"""Which margin sends a fast test to the exact f64 recheck?  Numpy f32 port of `fast_test` (annotate.cu) on the
non-occupied voxels of a few synthetic tracklets: share of tests that are undecided because of the row margin, the
column margin or the range margin, and what a tighter column evaluation would leave.  MUFU approximations are
replaced by exact f32 division / sqrt (their 2-ulp error sits inside the margins either way).  CPU only; informs
DESIGN.md section 8 item 1."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from objectcentricocccompletion_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402

f32 = np.float32
K_ATAN_ERR = f32(2.0e-6)


def u_of_angle(a):
    s, c = np.sin(a), np.cos(a)
    return s / (np.abs(s) + c)


def atan2_fast(y, x):
    ax, ay = np.abs(x), np.abs(y)
    mx, mn = np.maximum(ax, ay), np.minimum(ax, ay)
    a = (mn / mx).astype(f32)
    s = a * a
    p = np.full_like(a, f32(0.006811790633946657))
    for c in (-0.0336042158305645, 0.07962366938591003, -0.1323334276676178, 0.19807815551757812,
              -0.3331736922264099, 0.9999961256980896):
        p = (p * s + f32(c)).astype(f32)
    r = p * a
    r = np.where(ay > ax, f32(1.57079632679489661923) - r, r)
    r = np.where(x < 0, f32(3.14159265358979323846) - r, r)
    return np.where(y < 0, -r, r).astype(f32)


def main(n_trk=3, n_frames=16, seed=1):
    batch = synth.make_batch(n_trk, n_frames, 0.2, seed=seed)
    res = oracle.annotate_batch(batch, threads=8)
    tot = und_row = und_col = und_rng = und_any = und_col_tight = 0
    for t, r in enumerate(res):
        if r["occ"] is None:
            continue
        trk = batch.tracklets[t]
        seg = batch.segments[trk.segment]
        size = r["size"].astype(f32)
        mb = np.array([-size[0] * f32(0.5), -size[1] * f32(0.5), 0], f32)
        vs = batch.voxel_size
        dims = np.array(r["occ"].shape)
        idx = np.stack(np.nonzero(r["occ"] != 1), 1).astype(f32)
        pk = oracle.PackedBatch(type(batch)(segments=batch.segments, tracklets=[trk], voxel_size=vs))
        for i in range(len(trk)):
            rc, rs = np.float64(pk.trig[i, 2]), np.float64(pk.trig[i, 3])
            Rm = np.array([[rc, rs, 0], [-rs, rc, 0], [0, 0, 1]])
            f = int(trk.frame_ids[i])
            for c in range(len(seg.inclinations)):
                sn = pk.sensors[pk.frame_sf[i], c]
                V = sn["v2l"].astype(np.float64).reshape(3, 4)
                VR = V[:, :3] @ Rm
                b64 = VR @ (mb.astype(np.float64) + vs / 2) + V[:, :3] @ trk.boxes[i, :3].astype(np.float64) + V[:, 3]
                A, b = (vs * VR).astype(f32), b64.astype(f32)
                M = np.abs(b64) + (np.abs(vs * VR) * np.maximum(dims - 1, 0)).sum(1)
                eps = f32(6.0 * 5.9604644775390625e-08 * M.max())
                ri = seg.range_images[c][f]
                H, W = ri.shape
                tab = seg.inclinations[c][::-1].astype(np.float64)
                ub = np.concatenate([[2.0], u_of_angle(0.5 * (tab[:-1] + tab[1:])), [-2.0]]).astype(f32)   # sentinels
                p = (idx @ A.T + b).astype(f32)
                px, py, pz = p[:, 0], p[:, 1], p[:, 2]
                s2 = (py * py + px * px).astype(f32)
                r2 = (pz * pz + s2).astype(f32)
                inv_rho, inv_r = (f32(1) / np.sqrt(s2)).astype(f32), (f32(1) / np.sqrt(r2)).astype(f32)
                u = (pz / (s2 * inv_rho + np.abs(pz))).astype(f32)
                row = (ub[1:-1][None, :] > u[:, None]).sum(1)                         # number of boundaries above u
                above, below = ub[row], ub[row + 1]
                ok_row = np.minimum(u - below, above - u) > (f32(1.5) * eps * inv_r + f32(1.5e-6))
                kcol = f32(W / (2 * np.pi))
                c_col = f32(3.0) * f32(W) * f32(1.1920929e-07) + f32(W) * f32(4e-8)
                az = atan2_fast(py, px) + f32(sn["azc"])
                colf = (az * (-kcol) + (f32(0.5) * f32(W) - f32(0.5))).astype(f32)
                cr = np.rint(colf)
                frac = np.abs(colf - cr)
                ecol = (K_ATAN_ERR + f32(3e-7)) * kcol + c_col
                ok_col = frac + (f32(1.5) * eps * kcol * inv_rho + ecol) < f32(0.5)
                # tighter evaluation: small-angle arctangent after a rotation -> 3e-7 rad, colf rounding
                # at magnitude ~200 px instead of 1.5 W
                ecol_t = f32(3e-7) * kcol + f32(3.0) * f32(200.0) * f32(1.1920929e-07) + f32(W) * f32(4e-8)
                ok_col_t = frac + (f32(1.5) * eps * kcol * inv_rho + ecol_t) < f32(0.5)
                col = np.mod(cr.astype(np.int64), W)
                rv = ri[np.minimum(row, H - 1), col]
                rr = r2 * inv_r
                m = rr * (rr * f32(6.0e-7) + f32(2.01 * 1.7321) * eps) + f32(3.0003) * eps * eps
                ri2 = rv * rv
                ok_rng = (ri2 >= r2 + m) | (ri2 <= r2 - m)
                n = len(idx)
                tot += n
                und_row += int((~ok_row).sum())
                und_col += int((~ok_col).sum())
                und_rng += int((~ok_rng).sum())
                und_any += int((~(ok_row & ok_col & ok_rng)).sum())
                und_col_tight += int((~(ok_row & ok_col_t & ok_rng)).sum())
    pct = lambda v: f"{100.0 * v / tot:.3f} %"
    print(f"tests: {tot}; undecided: {pct(und_any)}  (row margin {pct(und_row)}, column margin {pct(und_col)}, "
          f"range margin {pct(und_rng)});  with the tighter column evaluation: {pct(und_col_tight)}")


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
