#!/usr/bin/env python
"""Numpy f32 emulation of the round-2 fast visibility test (csrc/annotate.cu: `make_pair`, `k_table_setup`,
`fast_test`, `k_brick_cull`) checked against the exact per-voxel results of the oracle.

What it proves on the CPU before any GPU time is spent:
  * every test the f32 path DECIDES agrees with the reference's f64 decision (row, column and free/not free);
  * the share of undecided tests (they go to the exact f64 recheck), split by cause;
  * a (brick, pair) the brick cull masks never contains a voxel the reference frees through that pair, and how
    many of the never-freed tests it removes.

MUFU approximations (rcp / rsqrt, <= 2 ulp) are replaced by correctly rounded f32 division / sqrt and perturbed by
+-2 ulp in a second pass, so the margins are exercised on both sides.  fma(a,b,c) is emulated through f64.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from objectcentricocccompletion_b200 import synth  # noqa: E402
from oracle import brick_cull, oracle  # noqa: E402

f32 = np.float32
U24 = 5.9604644775390625e-08          # 2^-24
K_LUT_PER_ROW = 64
ATAN_C = (0.006811790633946657, -0.0336042158305645, 0.07962366938591003, -0.1323334276676178,
          0.19807815551757812, -0.3331736922264099, 0.9999961256980896)
K_ATAN_NARROW = 1.0e-6                 # |t| <= 1 polynomial path
K_ATAN_WIDE = 2.0e-6                   # full-quadrant path
MSCALE = float(os.environ.get('EMU_MARGIN_SCALE', '1'))   # sanity: 0 must produce decided-but-wrong tests


def fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def u_of_angle(a):
    s, c = np.sin(a), np.cos(a)
    return s / (np.abs(s) + c)


def build_table(tab_flipped):
    """k_table_setup: boundaries in u space and the 8-byte lookup cells (b, h).  None if no fast path."""
    tab = tab_flipped.astype(np.float64)
    H = len(tab)
    if H < 2 or not (np.diff(tab) < 0).all() or not (np.abs(tab) < 1.5707).all():
        return None
    ub = u_of_angle(0.5 * (tab[:-1] + tab[1:])).astype(f32)              # H-1 boundaries, descending
    spacing = float((ub[:-1] - ub[1:]).min()) if H > 2 else 0.25
    if not (spacing > 1e-6):
        return None
    w = f32(0.5 * spacing)
    ncell = int(np.ceil(2.04 / w)) + 1
    if ncell > H * K_LUT_PER_ROW:
        return None
    inv_w = f32(1.0) / w
    cell0m = f32(1.02) * inv_w
    k = np.arange(ncell)
    uk = (k.astype(np.float64) - float(cell0m)) / float(inv_w)           # nominal centre of cell k
    ubd = ub.astype(np.float64)
    # boundary inside the extended cell [uk - 0.75 w, uk + 0.75 w] (at most one: spacing >= 2 w)
    j = np.abs(ubd[None, :] - uk[:, None]).argmin(1)
    inside = np.abs(ubd[j] - uk) <= 0.75 * float(w)
    b = np.where(inside, ub[j], f32(-4.0)).astype(f32)
    h = np.where(inside, j, (ubd[None, :] > uk[:, None]).sum(1)).astype(np.int32)
    return dict(ub=ub, w=w, inv_w=inv_w, cell0m=cell0m, ncell=ncell, b=b, h=h, H=H)


def make_pair(VR, V, box_o, mb, vs, dims, azc32, W):
    """Per (frame, LiDAR): the affine map voxel index -> sensor frame, rotated about the sensor's z axis so that the
    grid centre lies on +x; returns the f32 record fields and the f64 geometry the cull uses."""
    c0 = mb.astype(np.float64) + vs / 2
    b64 = VR @ c0 + V[:, :3] @ box_o + V[:, 3]
    A64 = vs * VR
    span = np.maximum(dims - 1, 0).astype(np.float64)
    cen = 0.5 * span
    pcen = b64 + A64 @ cen
    rho_c = np.hypot(pcen[0], pcen[1])
    ct, st = (pcen[0] / rho_c, pcen[1] / rho_c) if rho_c > 0 else (1.0, 0.0)
    Rz = np.array([[ct, st, 0], [-st, ct, 0], [0, 0, 1]])
    A2 = Rz @ A64
    bc = Rz @ pcen
    theta = np.arctan2(pcen[1], pcen[0])
    Rxy = 0.0                                                            # half extent of the lattice in the sensor frame
    R3 = 0.0
    for k in range(3):
        R3 += (A64[:, k] ** 2).sum() * (0.5 * span[k]) ** 2
    R = np.sqrt(R3) * 1.001 + 1e-3
    A32, bc32, cen32 = A2.astype(f32), bc.astype(f32), cen.astype(f32)
    M = np.abs(bc) + (np.abs(A2) * (0.5 * span)[None, :]).sum(1)
    eps_y = 6.0 * U24 * M[1]
    eps_xz = 6.0 * U24 * max(M[0], M[2])
    d_c = np.linalg.norm(pcen)
    narrow = rho_c > 2.1 * R and not os.environ.get('EMU_FORCE_WIDE')
    tmax = min(R / (rho_c - R), 1.0) if narrow else 1.0
    kcol = W / (2 * np.pi)
    C0 = (W - 0.5) - (theta + float(azc32) + np.pi) / (2 * np.pi) * W
    C0m = C0 % W
    cint = int(np.floor(C0m))
    c0f = f32(C0m - cint)
    phimax = np.arctan(tmax) if narrow else np.pi
    c_col = 1.5 * 2 * U24 * (kcol * phimax + 2.0) + W * 4e-8
    katan = K_ATAN_NARROW if narrow else K_ATAN_WIDE
    rec = dict(A=A32, bc=bc32, cen=cen32, narrow=narrow, eps_y=f32(eps_y), eps_xz=f32(eps_xz),
               e15z=f32(1.5 * max(eps_xz, eps_y)), ecolk=f32(1.5 * (eps_y + tmax * eps_xz) * kcol),
               ecol=f32((katan + 3e-7) * kcol + c_col), nkcol=f32(-kcol), c0f=c0f, cint=cint,
               c1=f32(1.6 * eps_xz + 1.1 * eps_y), W=W, d_min=max(d_c - R, 1e-3))
    return rec


def fast_test(p, tab, idx, ri, ulp_rcp=0):
    """-> (code [n]: 2 free / 0 not free / 1 undecided, row, col, cause flags)."""
    A, bc, cen = p["A"], p["bc"], p["cen"]
    d = (idx - cen[None, :]).astype(f32)
    dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
    px = fma(dz, A[0, 2], fma(dy, A[0, 1], fma(dx, A[0, 0], bc[0])))
    py = fma(dz, A[1, 2], fma(dy, A[1, 1], fma(dx, A[1, 0], bc[1])))
    pz = fma(dz, A[2, 2], fma(dy, A[2, 1], fma(dx, A[2, 0], bc[2])))
    s2 = fma(py, py, (px * px).astype(f32))
    r2 = fma(pz, pz, s2)

    def approx(v):                      # a MUFU result: the correctly rounded value moved by 2 ulps (ulp_rcp = +-1)
        v = v.astype(f32)
        if not ulp_rcp:
            return v
        tgt = f32(np.inf) if ulp_rcp > 0 else f32(-np.inf)
        return np.nextafter(np.nextafter(v, tgt), tgt)

    inv_rho = approx(f32(1) / np.sqrt(s2))
    inv_r = approx(f32(1) / np.sqrt(r2))
    # row
    u = (pz * approx(f32(1) / fma(s2, inv_rho, np.abs(pz)))).astype(f32)
    cellf = fma(u, tab["inv_w"], tab["cell0m"])
    cell = np.clip(np.rint(cellf).astype(np.int64), 0, tab["ncell"] - 1)
    b, h = tab["b"][cell], tab["h"][cell]
    row = h + (b > u)
    m_row = fma(p["e15z"], inv_r, f32(1.5e-6))
    ok_row = np.abs((u - b).astype(f32)) > m_row * f32(MSCALE)
    # column
    if p["narrow"]:
        t = (py * approx(f32(1) / px)).astype(f32)
        s = (t * t).astype(f32)
        q = np.full_like(t, f32(ATAN_C[0]))
        for c in ATAN_C[1:]:
            q = fma(q, s, f32(c))
        phi = (q * t).astype(f32)
    else:
        ax, ay = np.abs(px), np.abs(py)
        mx, mn = np.maximum(ax, ay), np.minimum(ax, ay)
        a = (mn * approx(f32(1) / mx)).astype(f32)
        s = (a * a).astype(f32)
        q = np.full_like(a, f32(ATAN_C[0]))
        for c in ATAN_C[1:]:
            q = fma(q, s, f32(c))
        r = (q * a).astype(f32)
        r = np.where(ay > ax, f32(1.57079632679489661923) - r, r).astype(f32)
        r = np.where(px < 0, f32(3.14159265358979323846) - r, r).astype(f32)
        phi = np.where(py < 0, -r, r).astype(f32)
    colf = fma(phi, p["nkcol"], p["c0f"])
    cr = np.rint(colf).astype(f32)
    ok_col = (np.abs((colf - cr).astype(f32)) + fma(p["ecolk"], inv_rho, p["ecol"]) * f32(MSCALE)).astype(f32) < f32(0.5)
    W = p["W"]
    col = p["cint"] + cr.astype(np.int64)
    col = np.where(col < 0, col + W, col)
    col = np.where(col >= W, col - W, col)
    col = np.clip(col, 0, W - 1)
    # range
    rv = ri[np.clip(row, 0, tab["H"] - 1), col]
    rr = (r2 * inv_r).astype(f32)
    dd = (rv - rr).astype(f32)
    ok_rng = np.abs(dd) > fma(rr, f32(6.5e-7), p["c1"]) * f32(MSCALE)
    yes = dd > 0
    code = np.where(ok_row & ok_col & ok_rng, np.where(yes, 2, 0), 1)
    return code, row, col, (ok_row, ok_col, ok_rng)


def main(n_trk=3, n_frames=16, seed=1, kind="vehicle", vs=0.2):
    batch = synth.make_batch(n_trk, n_frames, vs, kind, seed=seed)
    res = oracle.annotate_batch(batch, threads=8)
    tot = und = wrong = 0
    cause = np.zeros(3, np.int64)
    n_narrow = n_pairs = 0
    for t, r in enumerate(res):
        if r["occ"] is None:
            continue
        trk = batch.tracklets[t]
        seg = batch.segments[trk.segment]
        size = r["size"].astype(f32)
        mb = np.array([-size[0] * f32(0.5), -size[1] * f32(0.5), 0], f32)
        dims = np.array(r["occ"].shape)
        vox, rows, cols, rng, free = brick_cull.centre_tests(batch, t, r)
        idx = vox.astype(f32)
        pk = oracle.PackedBatch(type(batch)(segments=batch.segments, tracklets=[trk], voxel_size=vs))
        tabs = [build_table(np.ascontiguousarray(seg.inclinations[c][::-1])) for c in range(len(seg.inclinations))]
        for i in range(len(trk)):
            rc, rs = np.float64(pk.trig[i, 2]), np.float64(pk.trig[i, 3])
            Rm = np.array([[rc, rs, 0], [-rs, rc, 0], [0, 0, 1]])
            f = int(trk.frame_ids[i])
            for c in range(len(seg.inclinations)):
                sn = pk.sensors[pk.frame_sf[i], c]
                V = sn["v2l"].astype(np.float64).reshape(3, 4)
                ri = seg.range_images[c][f]
                H, W = ri.shape
                p = make_pair(V[:, :3] @ Rm, V, trk.boxes[i, :3].astype(np.float64), mb, vs, dims, sn["azc"], W)
                tab = tabs[c]
                # the pair qualifies for the fast path only if the row margin stays below a fifth of a cell
                m_row_max = float(p["e15z"]) / p["d_min"] + 1.5e-6
                assert tab is not None and m_row_max < 0.2 * float(tab["w"])
                n_pairs += 1
                n_narrow += bool(p["narrow"])
                for ulp in (0, 1, -1):
                    code, row, col, oks = fast_test(p, tab, idx, ri, ulp)
                    dec = code != 1
                    cref = np.where(cols[i, c] < 0, cols[i, c] + W, cols[i, c])
                    bad = dec & ((row != rows[i, c]) | (col != cref) | ((code == 2) != free[i, c]))
                    wrong += int(bad.sum())
                    if ulp == 0:
                        tot += len(idx)
                        und += int((~dec).sum())
                        cause += [int((~o).sum()) for o in oks]
    print(f"tests {tot}, pairs {n_pairs} (narrow {n_narrow}); decided-but-wrong {wrong}; undecided "
          f"{100.0 * und / tot:.4f} %  (row {100.0 * cause[0] / tot:.4f} %, col {100.0 * cause[1] / tot:.4f} %, "
          f"range {100.0 * cause[2] / tot:.4f} %)")
    return wrong


if __name__ == "__main__":
    a = sys.argv[1:]
    w = main(int(a[0]) if a else 3, int(a[1]) if len(a) > 1 else 16, int(a[2]) if len(a) > 2 else 1,
             a[3] if len(a) > 3 else "vehicle", float(a[4]) if len(a) > 4 else 0.2)
    sys.exit(1 if w else 0)
