"""GPU tests of the windowed range-image upload: the three upload modes give identical labels, the blocks the
device marks cover every pixel the reference reads, and the pull moves a fraction of the whole images."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(batch, pin, windows, flags=0):
    import torch

    from objectcentricocccompletion_b200 import occ_annotate

    pk = occ_annotate.pack_tracklets(batch)
    host = occ_annotate.HostBuffers(pk, pin=pin, windows=windows)
    dev = occ_annotate.DeviceTracklets(pk)
    dev.upload(host)
    dev.run(flags)
    torch.cuda.synchronize()
    return pk, host, dev, dev.results()


@pytest.mark.parametrize("cfg", [
    dict(n=6, b=14, vs=0.2, kind="vehicle", seed=1, small=False),
    dict(n=3, b=12, vs=0.1, kind="large", seed=2, small=False),
    dict(n=5, b=25, vs=0.25, kind="vehicle", seed=3, small=True),
])
def test_upload_modes_agree_with_the_oracle(cfg):
    from objectcentricocccompletion_b200 import synth
    from oracle import oracle
    from tests.util import assert_same_results

    batch = synth.make_batch(cfg["n"], cfg["b"], cfg["vs"], cfg["kind"], cfg["seed"], small=cfg["small"])
    exp = oracle.annotate_batch(batch, threads=8)
    modes = {}
    for pin, windows in ((True, True), (False, True), (True, False)):
        pk, host, dev, res = _run(batch, pin, windows)
        modes[host.ri_mode] = (host, dev)
        assert_same_results(res, exp, f"{cfg} {host.ri_mode}")
        if host.ri_mode == "pull":          # the all-f64 kernel reads the same pool
            _, _, _, res64 = _run(batch, pin, windows, flags=1)
            assert_same_results(res64, exp, f"{cfg} pull f64")
    assert set(modes) == {"pull", "host", "whole"}
    pulled = modes["pull"][1].pulled_bytes()
    assert 0 < pulled < 4 * pk.ri_len
    assert modes["host"][0].nbytes() < modes["whole"][0].nbytes()


def test_device_marks_cover_the_reference_pixels():
    from objectcentricocccompletion_b200 import synth
    from tests.test_ri_windows import check_cover, device_style_mask

    batch = synth.make_batch(5, 12, 0.2, seed=21, small=True)
    batch.tracklets[0].boxes[:, :2] *= 0.03              # one object on top of the ego vehicle: whole rows / images
    pk, host, dev, _ = _run(batch, True, True)
    mask = dev.window_mask()
    assert check_cover(batch, pk, mask, 8) > 0
    emu = device_style_mask(pk)                           # the same code compiled for the host
    assert (mask != emu).mean() < 0.01
    assert dev.pulled_bytes() == 32 * int(mask.sum())
    # the pulled blocks hold the source pixels, everything else is still zero
    pool = dev.bufs["ri_pool"].cpu().numpy().view(np.float32)[: pk.ri_len].reshape(-1, 8)
    src = pk.ri_pool.reshape(-1, 8)
    assert (pool[mask] == src[mask]).all() and not pool[~mask].any()


def test_pull_full_size_c2():
    """BASELINE config 2 through the pull path: labels equal to the whole-image upload, under a quarter of the bytes."""
    from objectcentricocccompletion_b200 import synth

    batch = synth.config_batch("c2", seed=0)
    pk, _, dev, res = _run(batch, True, True)
    _, _, _, ref = _run(batch, True, False)
    for x, y in zip(res, ref):
        assert x["status"] == y["status"]
        if y["occ"] is not None:
            assert (x["occ"] == y["occ"]).all() and x["n_unknown"] == y["n_unknown"]
    assert dev.pulled_bytes() < pk.ri_len


def test_tile_flags_cover_the_marked_blocks_and_keep_the_labels():
    """args.ri_tile_live (ABI v7): every block the pull marks lies in a pyramid tile flagged live, a good part of the
    tiles is dead (the pyramid skips them), and labels / dims / statuses are the same with and without the flags."""
    import torch

    from objectcentricocccompletion_b200 import occ_annotate, synth

    batch = synth.make_batch(6, 14, 0.2, seed=4)
    pk = occ_annotate.pack_tracklets(batch)
    host = occ_annotate.HostBuffers(pk, pin=True, windows=True)
    out = {}
    for use in (True, False):
        dev = occ_annotate.DeviceTracklets(pk)
        dev.use_tile_live = use
        dev.upload(host)
        assert (dev.tile_live is not None) == use
        for flags in (0, occ_annotate.FLAG_NO_BRICK_CULL):
            dev.run(flags)
            torch.cuda.synchronize()
            out[use, flags] = dev.results()
        if use:
            live = dev.tile_live.cpu().numpy().astype(bool)
            mask = dev.window_mask()
    for flags in (0, occ_annotate.FLAG_NO_BRICK_CULL):
        for x, y in zip(out[True, flags], out[False, flags]):
            assert x["status"] == y["status"] and (x["dims"] == y["dims"]).all()
            if y["occ"] is not None:
                assert (x["occ"] == y["occ"]).all()
    assert 0 < live.mean() < 0.7, live.mean()
    # every pixel the reference's visibility test reads lies in a live tile (the blocks of the mask are 8 floats of the
    # flat pool and may reach into a neighbouring tile; only the pixels a test can read matter)
    from tests.test_ri_windows import check_cover

    pix = np.zeros(pk.ri_len, bool)
    sens = pk.sensors.reshape(-1)
    for e in range(sens.size):
        H, W = int(sens["H"][e]), int(sens["W"][e])
        ntr, ntc = (H + 7) // 8, (W + 31) // 32
        lv = live[int(pk.pyr_off[e]): int(pk.pyr_off[e]) + ntr * ntc].reshape(ntr, ntc)
        img = np.repeat(np.repeat(lv, 8, 0), 32, 1)[:H, :W]
        o = int(sens["ri_off"][e])
        pix[o: o + H * W] |= img.reshape(-1)
    assert check_cover(batch, pk, pix, 1) > 0
    assert mask.any()
