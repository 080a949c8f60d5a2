"""GPU tests of the label consumers with random sampling against fixtures produced by the reference's
own code (oracle/make_golden_samplers.py: RandomSampleOccPoints and sample_observation exec'd from their source
files on CPU tensors with a fixed torch seed).  With rng="cpu" the draws use the same torch CPU generator stream, so
everything -- indices, labels, centres -- must be bit-identical."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _golden():
    from tests.util import GOLDEN

    return dict(np.load(os.path.join(GOLDEN, "samplers.npz")))


@pytest.mark.parametrize("case", ["weighted", "balance", "mirror_unknown", "all", "oversample"])
def test_random_sample_occ_points_matches_the_reference(case):
    import torch

    from objectcentricocccompletion_b200 import occ_ops
    from oracle.make_golden_samplers import SAMPLER_CASES

    d = _golden()
    n = int(d["n_grids"])
    grids = [d[f"grid{i}"] for i in range(n)]
    scores = [0.0 if not (g > 0).any() else 0.8 for g in grids]
    kw = SAMPLER_CASES[case]
    torch.manual_seed(1234)
    res = dict(occ_infos=[dict(occ_label_name=f"g{i}") for i in range(n)],
               occ_label_list=[torch.from_numpy(g.copy()).cuda() for g in grids], occ_scores=torch.tensor(scores))
    out = occ_ops.RandomSampleOccPoints(rng="cpu", **kw)(res)
    if kw["num_sample_points"] == -1:
        for i in range(n):
            assert (out["sample_occs"][i].cpu().numpy() == d[f"rs_{case}_occ{i}"]).all()
            assert (out["sample_occ_centers"][i].cpu().numpy() == d[f"rs_{case}_cen{i}"]).all()
    else:
        assert out["sample_occs"].shape == d[f"rs_{case}_occs"].shape
        assert (out["sample_occs"].cpu().numpy() == d[f"rs_{case}_occs"]).all()
        assert (out["sample_occ_centers"].cpu().numpy() == d[f"rs_{case}_centers"]).all()
    assert (out["occ_sizes"].cpu().numpy() == d[f"rs_{case}_sizes"]).all()
    assert (out["occ_scores"].numpy() == d[f"rs_{case}_scores"]).all()
    for i in range(n):                                       # the mirror fill is written back, as in the reference
        assert (out["occ_label_list"][i].cpu().numpy() == d[f"rs_{case}_grid{i}"]).all()


@pytest.mark.parametrize("comp", [True, False])
@pytest.mark.parametrize("case", ["plain", "weighted_ds", "balance", "balance_ds"])
def test_sample_observation_matches_the_reference(case, comp):
    import torch

    from objectcentricocccompletion_b200 import occ_ops
    from oracle.make_golden_samplers import OBS_CASES, OBS_SELF

    d = _golden()
    torch.manual_seed(4321)
    xyz, lab, rid = occ_ops.sample_observation(torch.from_numpy(d["obs_pts"]).cuda(), torch.from_numpy(d["obs_rois"]).cuda(),
                                               torch.from_numpy(d["obs_idx"]).cuda(), compensate_encoder_coors=comp,
                                               rng="cpu", **OBS_SELF, **OBS_CASES[case])
    key = f"obs_{case}_{int(comp)}"
    assert lab.shape == d[key + "_lab"].shape
    assert (lab.cpu().numpy() == d[key + "_lab"]).all()
    assert (rid.cpu().numpy() == d[key + "_roi"]).all()
    assert (xyz.cpu().numpy() == d[key + "_xyz"]).all()


def test_sample_observation_device_rng():
    """rng="cuda" (where the reference's head runs it): reproducible under a seed, balanced as specified."""
    import torch

    from objectcentricocccompletion_b200 import occ_ops
    from oracle.make_golden_samplers import OBS_SELF

    d = _golden()
    args = (torch.from_numpy(d["obs_pts"]).cuda(), torch.from_numpy(d["obs_rois"]).cuda(), torch.from_numpy(d["obs_idx"]).cuda())
    outs = []
    for _ in range(2):
        torch.manual_seed(7)
        outs.append(occ_ops.sample_observation(*args, balance_sample=True, compensate_encoder_coors=True, **OBS_SELF))
    assert all((a == b).all() for a, b in zip(*outs))
    xyz, lab, rid = outs[0]
    for r in rid.unique().tolist():
        m = rid == r
        pos, neg = int((lab[m] == 1).sum()), int((lab[m] == 0).sum())
        assert neg == max(pos, 1)                            # as many negatives as positives (one if no positive)
    # the deterministic case equals the CPU-reference fixture whatever the rng
    xyz, lab, rid = occ_ops.sample_observation(*args, compensate_encoder_coors=True, **OBS_SELF)
    assert (lab.cpu().numpy() == d["obs_plain_1_lab"]).all() and (xyz.cpu().numpy() == d["obs_plain_1_xyz"]).all()
