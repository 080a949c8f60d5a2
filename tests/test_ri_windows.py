"""Windowed range-image upload (csrc/ri_windows.cu): the host-side window geometry must cover every pixel the
reference's visibility test reads (occ_annotate.py:141-201, 541-547), and the gather must reproduce the pool."""
import numpy as np
import pytest

from objectcentricocccompletion_b200 import occ_annotate, synth
from oracle import oracle


def _centres(dims, size, vs):
    """f64 voxel centres of the whole grid (occ_annotate.py:467-471), box frame."""
    X, Y, Z = (int(v) for v in dims)
    g = np.stack(np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij"), -1).reshape(-1, 3)
    mb = np.array([np.float32(size[0]) * np.float32(-0.5), np.float32(size[1]) * np.float32(-0.5), 0.0], np.float32)
    return g.astype(np.float64) * vs + mb.astype(np.float64) + vs / 2


def _check_cover(batch):
    pk = occ_annotate.pack_tracklets(batch)
    blocks = occ_annotate.window_blocks(pk)
    nblk = pk.ri_len // occ_annotate.RI_BLOCK
    assert pk.ri_len % occ_annotate.RI_BLOCK == 0
    mask = np.zeros(nblk, bool)
    mask[blocks] = True
    n_pix = check_cover(batch, pk, mask, occ_annotate.RI_BLOCK)
    return blocks, pk, n_pix


def check_cover(batch, pk, mask, blk):
    """Every pixel the reference's visibility test reads for the batch lies in a marked block of `blk` floats."""
    res = oracle.annotate_batch(batch, threads=4)
    sensors = pk.sensors
    n_pix = 0
    vsf = np.float32(batch.voxel_size)
    for t, (trk, r) in enumerate(zip(batch.tracklets, res)):
        if r["occ"] is not None:
            cen = _centres(r["dims"], r["size"], batch.voxel_size)
        else:        # no grid in the reference (e.g. no in-box point): the largest grid the tracklet could have had
            smax = pk.trk_smax[t]
            cen = _centres(np.ceil(smax / vsf).astype(np.int64), smax, batch.voxel_size)
        seg = batch.segments[trk.segment]
        trig = oracle.host_trig(trk.boxes[:, 6])
        for i in range(len(trk)):
            c, s = float(trig[i, 2]), float(trig[i, 3])
            R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float64)
            ego = cen @ R + trk.boxes[i, :3].astype(np.float64)          # :490-499
            fid = int(trk.frame_ids[i])
            sf = int(pk.frame_sf[pk.trk_frame_off[t] + i])
            for l in range(pk.L):
                H, W = int(sensors["H"][sf, l]), int(sensors["W"][sf, l])
                incl = np.ascontiguousarray(seg.inclinations[l][::-1], np.float32)[None]
                idx, _ = oracle.point_cloud_to_range_image_idx(ego[None], seg.extrinsics[fid, l][None], incl, (H, W))
                row, col = idx[0, :, 0], idx[0, :, 1]
                col = np.where(col < 0, col + W, col)                     # negative index wrap (:543)
                flat = int(sensors["ri_off"][sf, l]) + row * W + col
                hit = mask[flat // blk]
                assert hit.all(), f"tracklet {t} frame {i} lidar {l}: {int((~hit).sum())} pixels outside the windows"
                n_pix += flat.size
    return n_pix


@pytest.mark.parametrize("kind,vs", [("vehicle", 0.2), ("large", 0.1)])
def test_windows_cover_every_pixel_the_reference_reads(kind, vs):
    batch = synth.make_batch(4, 12, vs, kind=kind, seed=11, small=True)
    blocks, pk, n = _check_cover(batch)
    assert n > 0
    assert 0 < blocks.size < pk.ri_len // occ_annotate.RI_BLOCK      # a real subset


def test_windows_cover_close_and_overhead_objects():
    """Boxes next to / around the sensor axis (rho < R, d < R): whole rows / whole images are marked."""
    batch = synth.make_batch(2, 10, 0.2, seed=5, small=True)
    for trk in batch.tracklets:
        trk.boxes[:, :2] *= 0.02                                          # drag the track onto the ego vehicle
    _check_cover(batch)


def device_style_mask(pk, sub_edge=0.0):
    """The 8-float block mask of the device-side path, from the host build of the same footprint code."""
    return occ_annotate.window_mask(pk, sub_edge).astype(bool)


@pytest.mark.parametrize("kind,vs,drag", [("vehicle", 0.2, 1.0), ("large", 0.1, 1.0), ("vehicle", 0.2, 0.02)])
def test_subbox_footprints_cover_every_pixel_the_reference_reads(kind, vs, drag):
    """The footprint code of k_window_mark (host build): sub-box footprints of the tracklet's centre box."""
    batch = synth.make_batch(4, 12, vs, kind=kind, seed=13, small=True)
    for trk in batch.tracklets:
        trk.boxes[:, :2] *= drag
    pk = occ_annotate.pack_tracklets(batch)
    mask = device_style_mask(pk)
    assert check_cover(batch, pk, mask, 8) > 0
    if drag == 1.0:
        one = np.zeros(pk.ri_len // 16, bool)
        one[occ_annotate.window_blocks(pk)] = True
        assert mask.sum() * 8 <= one.sum() * 16                # sub-boxes are at least as tight as one box
        assert one.sum() * 16 < 0.6 * pk.ri_len                # and one box is already a fraction of the images


def test_gather_blocks_reproduces_the_pool():
    batch = synth.make_batch(3, 12, 0.2, seed=3, small=True)
    pk = occ_annotate.pack_tracklets(batch)
    host = occ_annotate.HostBuffers(pk, pin=False)
    assert host.ri_mode == "host"
    pool = pk.ri_pool
    st = host.ri_staging.numpy().reshape(-1, occ_annotate.RI_BLOCK)
    exp = pool.reshape(-1, occ_annotate.RI_BLOCK)[host.ri_blocks]
    assert (st[: host.ri_blocks.size] == exp).all()
    assert host.nbytes() < occ_annotate.HostBuffers(pk, pin=False, windows=False).nbytes()


def test_host_copy_parts():
    """occb200_host_copy_parts (host code of the library): parts of odd sizes, an empty one and one larger than a
    copy piece land at their offsets; bytes between them are untouched."""
    from objectcentricocccompletion_b200 import _lib

    rng = np.random.default_rng(0)
    sizes = [0, 1, 4097, 300_000, 700_001]
    parts = [rng.integers(0, 256, n, dtype=np.uint8) for n in sizes]
    offs, o = [], 16
    for n in sizes:
        offs.append(o)
        o += n + 7
    dst = np.full(o + 32, 0xAB, np.uint8)
    src = np.asarray([p.ctypes.data for p in parts], np.uint64)
    a_sz, a_off = np.asarray(sizes, np.int64), np.asarray(offs, np.int64)
    rc = _lib.lib().occb200_host_copy_parts(src.ctypes.data, a_sz.ctypes.data, a_off.ctypes.data, len(parts), dst.ctypes.data)
    assert rc == 0
    exp = np.full_like(dst, 0xAB)
    for p, off in zip(parts, offs):
        exp[off: off + p.size] = p
    assert (dst == exp).all()


def test_mask_to_blocks_matches_numpy():
    """occb200_host_mask_to_blocks (two-pass, OpenMP): the ascending list of 16-float blocks that hold a marked
    8-float block, for odd lengths, empty / full / sparse masks and lengths around the chunk size."""
    from objectcentricocccompletion_b200 import _lib

    def blocks(m):
        out = np.empty((m.size + 1) // 2 + 1, np.uint32)
        n = _lib.lib().occb200_host_mask_to_blocks(m.ctypes.data, m.size, out.ctypes.data)
        return out[:n].copy()

    rng = np.random.default_rng(5)
    for n, p in ((0, 0.5), (1, 1.0), (7, 0.5), (9, 1.0), (65535, 0.2), (65536, 0.01), (65537, 0.9), (200_003, 0.0),
                 (1_000_001, 0.22)):
        m = np.ascontiguousarray((rng.random(n) < p).astype(np.uint8))
        k = (n + 1) // 2
        pairs = np.concatenate([m, np.zeros(2 * k - n, np.uint8)]).reshape(k, 2)
        exp = np.flatnonzero(pairs.any(1)).astype(np.uint32)
        got = blocks(m)
        assert got.size == exp.size and (got == exp).all(), (n, p)
