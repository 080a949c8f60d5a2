"""Golden vectors of the label consumers with random sampling (TEST INFRASTRUCTURE; runs only where /root/reference
exists): the reference's own ``RandomSampleOccPoints`` (mmdet3d/datasets/pipelines/occ_pinelines.py:130-358) and
``OccAutoEncoder.sample_observation`` (mmdet3d/models/roi_heads/bbox_heads/occ_ae_head.py:65-201), AST-extracted from
their source files (the modules cannot be imported: mmcv / mmdet registries) and executed on CPU tensors with a
fixed torch seed.  Writes tests/golden/samplers.npz.

    python -m oracle.make_golden_samplers
"""
from __future__ import annotations

import ast
import importlib.util
import os
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "samplers.npz")


def _load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_random_sample_class():
    path = f"{REF}/mmdet3d/datasets/pipelines/occ_pinelines.py"
    tree = ast.parse(open(path).read())
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "RandomSampleOccPoints")
    node.decorator_list = []                                  # @PIPELINES.register_module(): the registry is mmdet's
    ns = {"torch": torch, "np": np}
    exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns["RandomSampleOccPoints"]


def reference_sample_observation():
    path = f"{REF}/mmdet3d/models/roi_heads/bbox_heads/occ_ae_head.py"
    tree = ast.parse(open(path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "OccAutoEncoder")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "sample_observation")
    occ_ops = _load_by_path("ref_occ_ops", f"{REF}/mmdet3d/ops/occ/occ_ops.py")
    utils = _load_by_path("ref_box_utils", f"{REF}/mmdet3d/core/bbox/structures/utils.py")
    ns = {"torch": torch, "np": np, "occ_ops": occ_ops, "rotation_3d_in_axis": utils.rotation_3d_in_axis}
    exec(compile(ast.Module([fn], []), path, "exec"), ns)
    return ns["sample_observation"]


def label_grids(seed=0):
    """A few int32 label grids shaped like annotate outputs (0 unknown / 1 occupied / 2 free), one all-unknown."""
    rng = np.random.default_rng(seed)
    grids = []
    for dims in [(11, 24, 9), (12, 25, 10), (7, 9, 5), (15, 30, 11)]:
        g = rng.choice(np.array([0, 1, 2], np.int32), size=dims, p=[0.35, 0.15, 0.5]).astype(np.int32)
        grids.append(g)
    grids.insert(2, np.zeros((1, 1, 1), np.int32))             # "fake an empty grid" (occ_pinelines.py:45-47)
    return grids


SAMPLER_CASES = {
    "weighted": dict(num_sample_points=256, pos_sample_weight=0.5),
    "balance": dict(num_sample_points=300, pos_sample_weight=0.4, balance_sample=True),
    "mirror_unknown": dict(num_sample_points=128, mirror_x=True, use_unknown=True, weighted_sample=False),
    "all": dict(num_sample_points=-1, mirror_x=True),
    "oversample": dict(num_sample_points=4096, weighted_sample=False),
}


def observation_inputs(seed=0):
    rng = np.random.default_rng(seed)
    R = 5
    rois = np.zeros((R, 8), np.float32)
    rois[:, 0] = np.arange(R) % 2
    rois[:, 1:4] = rng.normal(0, 10, (R, 3))
    rois[:, 4:7] = np.array([2.1, 4.8, 1.8], np.float32) * (1 + 0.1 * rng.standard_normal((R, 3)))
    rois[:, 7] = rng.uniform(-3, 3, R)
    n = 4000
    idx = rng.integers(0, R - 1, n)                            # the last ROI gets no point
    pts = (rng.uniform(-0.55, 0.55, (n, 3)) * rois[idx, 4:7][:, [1, 0, 2]]).astype(np.float32)   # some fall outside
    return pts, rois, idx.astype(np.int64)


OBS_CASES = {
    "plain": dict(downsample_size=-1, balance_sample=False),
    "weighted_ds": dict(downsample_size=500, balance_sample=False),
    "balance": dict(downsample_size=-1, balance_sample=True),
    "balance_ds": dict(downsample_size=300, balance_sample=True),
}
OBS_SELF = dict(voxel_size=0.2, scale_wlh=[1.1, 1.1, 1.1], offset_wlh=[0.2, 0.2, 0.2])


def main():
    d = {}
    cls = reference_random_sample_class()
    grids = label_grids()
    for gi, g in enumerate(grids):
        d[f"grid{gi}"] = g
    d["n_grids"] = np.int64(len(grids))
    scores = [0.0 if not (g > 0).any() else 0.8 for g in grids]
    for name, kw in SAMPLER_CASES.items():
        torch.manual_seed(1234)
        res = dict(occ_infos=[dict(occ_label_name=f"g{i}") for i in range(len(grids))],
                   occ_label_list=[torch.from_numpy(g.copy()) for g in grids], occ_scores=torch.tensor(scores))
        out = cls(**kw)(res)
        if kw["num_sample_points"] == -1:
            for i, (a, b) in enumerate(zip(out["sample_occs"], out["sample_occ_centers"])):
                d[f"rs_{name}_occ{i}"] = a.numpy()
                d[f"rs_{name}_cen{i}"] = b.numpy()
        else:
            d[f"rs_{name}_occs"] = out["sample_occs"].numpy()
            d[f"rs_{name}_centers"] = out["sample_occ_centers"].numpy()
        d[f"rs_{name}_sizes"] = out["occ_sizes"].numpy()
        d[f"rs_{name}_scores"] = out["occ_scores"].numpy()
        for i, g in enumerate(out["occ_label_list"]):
            d[f"rs_{name}_grid{i}"] = g.numpy()                # the mirror fill is written back
    fn = reference_sample_observation()
    pts, rois, idx = observation_inputs()
    d["obs_pts"], d["obs_rois"], d["obs_idx"] = pts, rois, idx
    for comp in (False, True):
        me = types.SimpleNamespace(compensate_encoder_coors=comp, **OBS_SELF)
        for name, kw in OBS_CASES.items():
            torch.manual_seed(4321)
            xyz, lab, rid = fn(me, torch.from_numpy(pts), torch.from_numpy(rois), torch.from_numpy(idx), **kw)
            key = f"obs_{name}_{int(comp)}"
            d[key + "_xyz"], d[key + "_lab"], d[key + "_roi"] = xyz.numpy(), lab.numpy(), rid.numpy()
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, os.path.getsize(OUT), "bytes", len(d), "arrays")


if __name__ == "__main__":
    main()
