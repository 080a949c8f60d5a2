"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    """-> (batch, pack_override, expected_status list, {tracklet index: occ int32 [X,Y,Z]})"""
    from oracle.make_golden import arrays_to_batch

    d = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    batch = arrays_to_batch(d)
    override = dict(trig=d["pack_trig"], pib=d["pack_pib"], v2l=d["pack_v2l"], azc=d["pack_azc"])
    status = [str(s) for s in d["exp_status"]]
    occ = {t: d[f"exp_occ{t}"].astype(np.int32) for t in range(len(status)) if f"exp_occ{t}" in d}
    return batch, override, status, occ


def assert_same_results(got, exp, what=""):
    assert len(got) == len(exp)
    for t, (g, e) in enumerate(zip(got, exp)):
        assert g["status"] == e["status"], f"{what} tracklet {t}: status {g['status']} != {e['status']}"
        if e["occ"] is None:
            assert g["occ"] is None
            continue
        assert g["occ"].shape == e["occ"].shape, f"{what} tracklet {t}: shape {g['occ'].shape} != {e['occ'].shape}"
        bad = int((g["occ"] != e["occ"]).sum())
        assert bad == 0, f"{what} tracklet {t}: {bad} of {e['occ'].size} labels differ"
        assert (g["dims"] == e["dims"]).all() and (g["size"] == e["size"]).all()
        assert g["n_unknown"] == e["n_unknown"]
