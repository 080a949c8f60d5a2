"""``quantize_points`` / ``generate_dense_voxel_centers`` -- drop-ins for ``mmdet3d/ops/occ/occ_ops.py:5-93``."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _f3(v):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(-1))
    assert a.size == 3
    return a


def quantize_points(points, rois, rois_points_idx, voxel_size, scale_wlh=[1.0, 1.0, 1.0],
                    offset_wlh=[0.0, 0.0, 0.0], to_center=False, check_index=True):
    """points [N,3] (ROI-local), rois [R,8|10], rois_points_idx [N] -> int64 [N,3] voxel coords, or the f32
    voxel centres when ``to_center`` (occ_ops.py:53-93).  ``rois_points_idx`` follows PyTorch indexing (negative
    indices count from the end); an index outside [-R, R) raises IndexError as in the reference
    (``check_index=False`` skips the device->host read of the error counter; bad rows then hold INT64_MIN / NaN)."""
    _lib.require_cuda(points, rois, rois_points_idx)
    pts = points.float().contiguous()
    r = rois.float().contiguous()
    idx = rois_points_idx.long().contiguous()
    N = pts.size(0)
    coor = None if to_center else torch.empty((N, 3), dtype=torch.long, device=pts.device)
    cen = torch.empty((N, 3), dtype=torch.float32, device=pts.device) if to_center else None
    sc, of = _f3(scale_wlh), _f3(offset_wlh)
    n_bad = torch.zeros(1, dtype=torch.int64, device=pts.device) if check_index else None
    with torch.cuda.device(pts.device):
        rc = _lib.lib().occb200_quantize_points(pts.data_ptr(), N, r.data_ptr(), r.size(0), r.size(1), idx.data_ptr(),
                                                float(voxel_size), sc.ctypes.data, of.ctypes.data, int(to_center),
                                                _lib.ptr(coor), _lib.ptr(cen), _lib.ptr(n_bad),
                                                _lib.stream_ptr(pts.device))
    _lib.check(rc, "occb200_quantize_points")
    if check_index and int(n_bad) != 0:
        raise IndexError(f"rois_points_idx: {int(n_bad)} indices out of range for {r.size(0)} rois")
    return cen if to_center else coor


def generate_dense_voxel_centers(bbox_sizes, voxel_size, scale_wlh=[1.0, 1.0, 1.0], offset_wlh=[0.0, 0.0, 0.0],
                                 as_volume=False):
    """bbox_sizes [R,3] -> list of f32 [X*Y*Z,3] (or [X,Y,Z,3]) voxel centres in the gravity-centred object
    frame (occ_ops.py:5-50).  One kernel for all R boxes instead of a python loop of meshgrids."""
    _lib.require_cuda(bbox_sizes)
    sizes = bbox_sizes.float().contiguous()
    R = sizes.size(0)
    if R == 0:
        return []
    sc, of = _f3(scale_wlh), _f3(offset_wlh)
    # grid dims on the host with the reference's f32 ops (it syncs per ROI as well: torch.arange(XS))
    s_host = sizes.cpu() * torch.from_numpy(sc) + torch.from_numpy(of)
    dims = torch.ceil(s_host / voxel_size).to(torch.int32)
    nvox = dims.long().prod(1)
    off = torch.zeros(R + 1, dtype=torch.long)
    off[1:] = torch.cumsum(nvox, 0)
    total = int(off[-1])
    dev = sizes.device
    centers = torch.empty((total, 3), dtype=torch.float32, device=dev)
    dims_d, off_d = dims.to(dev), off.to(dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().occb200_dense_voxel_centers(sizes.data_ptr(), dims_d.data_ptr(), off_d.data_ptr(), R, total,
                                                    float(voxel_size), sc.ctypes.data, of.ctypes.data,
                                                    centers.data_ptr(), _lib.stream_ptr(dev))
    _lib.check(rc, "occb200_dense_voxel_centers")
    out = []
    for r in range(R):
        c = centers[int(off[r]):int(off[r + 1])]
        if as_volume:
            c = c.view(int(dims[r, 0]), int(dims[r, 1]), int(dims[r, 2]), 3)
        out.append(c)
    return out


def jitter_voxel_center(voxel_size, voxel_centers):
    """occ_ops.py:96-100 (plain torch RNG; not on the kernel path)."""
    return voxel_centers + (torch.rand_like(voxel_centers) * voxel_size - voxel_size / 2)


def mirror_occ_label(occ_label_list):
    """``MirrorOccLabel.__call__`` (mmdet3d/datasets/pipelines/occ_pinelines.py:88-126) for a list of CUDA label
    grids int32 [X,Y,Z]: unknown voxels take the label of their mirror image across the x mid-plane; all grids of
    the list in one launch.  Returns new tensors (the inputs are not modified)."""
    if len(occ_label_list) == 0:
        return []
    _lib.require_cuda(*occ_label_list)
    dev = occ_label_list[0].device
    dims = np.array([list(g.shape) for g in occ_label_list], np.int32).reshape(-1, 3)
    sizes = dims.astype(np.int64).prod(1)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    flat = torch.cat([g.to(torch.int32).reshape(-1) for g in occ_label_list]) if off[-1] else \
        torch.zeros(1, dtype=torch.int32, device=dev)
    out = torch.empty_like(flat)
    off_d = torch.from_numpy(off).to(dev)
    dims_d = torch.from_numpy(dims).to(dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().occb200_mirror_occ_label(flat.data_ptr(), off_d.data_ptr(), dims_d.data_ptr(), None,
                                                 len(occ_label_list), int(sizes.max()), out.data_ptr(),
                                                 _lib.stream_ptr(dev))
    _lib.check(rc, "occb200_mirror_occ_label")
    return [out[int(off[i]): int(off[i + 1])].view(*[int(v) for v in dims[i]]).to(occ_label_list[i].dtype)
            for i in range(len(occ_label_list))]


# ------------------------------------------------------------------------------------------------
# label consumers with random sampling
#
# RNG-parity policy.  Both routines below are deterministic except for their ``torch.multinomial`` calls.  Those
# are the reference's own calls -- same weights, same ``num_samples`` / ``replacement``, same order -- so with the
# same torch generator state the drawn indices, and therefore the outputs, are identical to the reference's.
# Everything around the draws (grids, mirror fill, compaction, centres, gathers) is batched on the device.
# ``rng="cpu"`` draws on the host (where the reference's dataloader pipeline runs its tensors; the weights of one
# grid are a few KB), ``rng="cuda"`` on the device (where the reference's head runs ``sample_observation``).
# ------------------------------------------------------------------------------------------------
def _multinomial(weights, num_samples, replacement, rng, generator):
    if rng == "cpu":
        return torch.multinomial(weights.cpu(), num_samples, replacement=replacement, generator=generator).to(weights.device)
    return torch.multinomial(weights, num_samples, replacement=replacement, generator=generator)


def observed_labels(pts_coors, pts_roi_inds, dims, off):
    """int64 [off[-1]]: the dense observation grids of all ROIs (occ_ae_head.py:100-127); ``dims`` int32 [R,3] and
    ``off`` int64 [R+1] on the host."""
    _lib.require_cuda(pts_coors, pts_roi_inds)
    dev = pts_coors.device
    total = int(off[-1])
    labels = torch.zeros(max(total, 1), dtype=torch.long, device=dev)
    c = pts_coors.long().contiguous()
    r = pts_roi_inds.long().contiguous()
    dims_d, off_d = dims.to(dev).contiguous(), off.to(dev).contiguous()
    with torch.cuda.device(dev):
        rc = _lib.lib().occb200_observed_labels(c.data_ptr(), r.data_ptr(), c.size(0), dims_d.data_ptr(), off_d.data_ptr(),
                                                dims.size(0), labels.data_ptr(), _lib.stream_ptr(dev))
    _lib.check(rc, "occb200_observed_labels")
    return labels[:total]


def sample_observation(local_xyz, rois, pts_roi_inds, voxel_size, scale_wlh=[1.0, 1.0, 1.0], offset_wlh=[0.0, 0.0, 0.0],
                       downsample_size=-1, balance_sample=False, compensate_encoder_coors=False, rng="cuda",
                       generator=None):
    """``OccAutoEncoder.sample_observation`` (mmdet3d/models/roi_heads/bbox_heads/occ_ae_head.py:65-201) with the head's
    attributes as arguments: per ROI the dense grid of voxel centres labelled 1 where a point was observed, then the
    reference's (balanced / weighted) down-sampling.  Returns (smp_pts_xyz_local f32 [K,3], obs_occ_labels int64 [K],
    smp_pts_roi_inds int64 [K]).  The grids of all ROIs are built by three launches (quantise, centres, observed
    labels) instead of a python loop over ROIs; only the ``torch.multinomial`` draws remain per ROI, in the
    reference's order."""
    _lib.require_cuda(local_xyz, rois, pts_roi_inds)
    assert rois.size(1) in (8, 10)
    if not compensate_encoder_coors:                      # :74-79, rotation_3d_in_axis(axis=2) by pi/2 (utils.py:21-61)
        ang = local_xyz.new_tensor((np.pi / 2,))
        s_, c_ = torch.sin(ang), torch.cos(ang)
        one, zero = torch.ones_like(c_), torch.zeros_like(c_)
        rot_t = torch.stack([torch.stack([c_, -s_, zero]), torch.stack([s_, c_, zero]), torch.stack([zero, zero, one])])
        local_xyz = torch.einsum('aij,jka->aik', (local_xyz[None, :, :], rot_t)).squeeze(0)
    pts_coors = quantize_points(local_xyz, rois, pts_roi_inds, voxel_size, scale_wlh=scale_wlh, offset_wlh=offset_wlh)
    volumes = generate_dense_voxel_centers(rois[:, 4:7], voxel_size, scale_wlh=scale_wlh, offset_wlh=offset_wlh,
                                           as_volume=True)
    R = len(volumes)
    dims = torch.tensor([list(v.shape[:3]) for v in volumes], dtype=torch.int32).reshape(-1, 3)
    off = torch.zeros(R + 1, dtype=torch.long)
    off[1:] = torch.cumsum(dims.long().prod(1), 0)
    labels_all = observed_labels(pts_coors, pts_roi_inds, dims, off)
    label_list, xyz_list, roi_list = [], [], []
    for i, volume in enumerate(volumes):
        new_labels = labels_all[int(off[i]): int(off[i + 1])]
        volume = volume.reshape(-1, 3)
        if balance_sample:                                # :131-160
            indexes = torch.arange(len(new_labels), device=new_labels.device)
            pos_indexes = indexes[new_labels == 1]
            neg_indexes = indexes[new_labels == 0]
            num_neg_sample = int(len(pos_indexes) * 1.0)
            if num_neg_sample > 0 and len(neg_indexes) > 0:
                choices = _multinomial(torch.ones_like(neg_indexes, dtype=torch.float32), num_neg_sample,
                                       num_neg_sample > len(neg_indexes), rng, generator)
                new_labels = torch.cat([new_labels[pos_indexes], new_labels[neg_indexes[choices]]])
                volume = torch.cat([volume[pos_indexes], volume[neg_indexes[choices]]])
            elif len(neg_indexes) > 0:
                new_labels = new_labels[:1]
                volume = volume[:1]
            else:
                continue
            if len(new_labels) > downsample_size > 0:
                choices = _multinomial(torch.ones_like(new_labels, dtype=torch.float32), downsample_size, False, rng,
                                       generator)
                new_labels = new_labels[choices]
                volume = volume[choices]
        elif len(new_labels) > downsample_size > 0:       # :162-174
            sample_weights = torch.ones_like(new_labels, dtype=torch.float32)
            sample_weights[new_labels == 1] = 100
            choices = _multinomial(sample_weights, downsample_size, False, rng, generator)
            new_labels = new_labels[choices]
            volume = volume[choices]
        label_list.append(new_labels)
        xyz_list.append(volume)
        roi_list.append(torch.full((new_labels.numel(),), i, dtype=torch.long, device=new_labels.device))
    if len(label_list) == 0:                              # :186-191
        return local_xyz.new_zeros((0, 3)), local_xyz.new_zeros((0,)), local_xyz.new_zeros((0,))
    return torch.cat(xyz_list, dim=0), torch.cat(label_list, dim=0), torch.cat(roi_list, dim=0)


class RandomSampleOccPoints(object):
    """``RandomSampleOccPoints`` (mmdet3d/datasets/pipelines/occ_pinelines.py:130-358) on CUDA label grids: same
    constructor, same ``results`` keys in and out (``occ_label_list``, ``occ_scores``, ``occ_infos`` ->
    ``sample_occs`` [N,K], ``sample_occ_centers`` [N,K,3], ``occ_sizes`` [N,3]).  The mirror fill of all grids is one
    launch (``mirror_occ_label``; as in the reference it is written back into ``occ_label_list``); the draws follow the
    RNG-parity policy above (default ``rng="cpu"``: the pipeline's tensors live on the host in the reference)."""

    def __init__(self, num_sample_points=1024, pos_sample_weight=0.5, voxel_size=0.2, use_unknown=False,
                 use_potential=False, mirror_x=False, balance_sample=False, weighted_sample=True, rng="cpu",
                 generator=None):
        self.num_sample_points = num_sample_points
        self.pos_sample_weight = pos_sample_weight
        self.voxel_size = voxel_size
        self.use_unknown = use_unknown
        self.use_potential = use_potential
        if use_potential:
            self.potential = {}
        self.mirror_x = mirror_x
        self.balance_sample = balance_sample
        self.weighted_sample = weighted_sample
        self.rng = rng
        self.generator = generator

    def _draw(self, weights, n, replacement):
        return _multinomial(weights, n, replacement, self.rng, self.generator)

    def __call__(self, results):
        if "occ_label_list" not in results:
            return results
        occ_infos, occ_grids, occ_scores = results["occ_infos"], results["occ_label_list"], results["occ_scores"]
        K = 0 if self.num_sample_points == -1 else self.num_sample_points
        if len(occ_grids) == 0:                                                   # :157-168
            results["sample_occs"] = torch.zeros((0, K))
            results["sample_occ_centers"] = torch.zeros((0, K, 3))
            results["occ_sizes"] = torch.zeros((0, 3))
            return results
        _lib.require_cuda(*occ_grids)
        dev = occ_grids[0].device
        annotated = [bool((g > 0).any()) for g in occ_grids]
        if self.mirror_x:                                                         # :205-219, written back in place
            live = [i for i, a in enumerate(annotated) if a]
            for i, m in zip(live, mirror_occ_label([occ_grids[i] for i in live])):
                occ_grids[i].copy_(m)
        sample_occs, sample_occ_centers, occ_sizes = [], [], []
        for i, (occ_grid, occ_score, info) in enumerate(zip(occ_grids, occ_scores, occ_infos)):
            if not annotated[i]:                                                  # :175-188
                assert occ_score == 0, "occ_score should be 0 if no occ grid is annotated"
                sample_centered = torch.zeros(K, 3, device=dev)
                sample_occ = torch.zeros(K, device=dev)
                w, l, h = 0.0, 0.0, 0.0
            else:
                XS, YS, ZS = occ_grid.shape
                occ_grid_flat = occ_grid.reshape(-1)
                flat = torch.arange(XS * YS * ZS, device=dev)
                voxel_coors = torch.stack([flat // (YS * ZS), (flat // ZS) % YS, flat % ZS], dim=-1)   # ij meshgrid
                if not self.use_unknown:                                          # :225-232
                    keep = occ_grid_flat > 0
                    valid_voxel_coors, valid_occ_grid = voxel_coors[keep], occ_grid_flat[keep]
                else:
                    valid_voxel_coors, valid_occ_grid = voxel_coors, occ_grid_flat.clone()
                w, l, h = float(XS) * self.voxel_size, float(YS) * self.voxel_size, float(ZS) * self.voxel_size
                min_bound = torch.tensor([-w / 2, -l / 2, -h / 2], dtype=torch.float32, device=dev)
                valid_voxel_centers = (valid_voxel_coors.to(torch.float) * self.voxel_size + min_bound
                                       + self.voxel_size / 2)                     # :245-249
                n_valid = len(valid_occ_grid)
                if self.num_sample_points == -1:
                    sample_idx = torch.arange(n_valid, device=dev)
                elif self.balance_sample:                                         # :253-288
                    num_pos = int(self.num_sample_points * self.pos_sample_weight)
                    num_neg = self.num_sample_points - num_pos
                    idxs = torch.arange(n_valid, device=dev)
                    pos_idxs, neg_idxs = idxs[valid_occ_grid == 1], idxs[valid_occ_grid != 1]
                    if len(pos_idxs) == 0 or len(neg_idxs) == 0:
                        sample_idx = self._draw(torch.ones_like(valid_occ_grid, dtype=torch.float),
                                                self.num_sample_points, n_valid < self.num_sample_points)
                        occ_scores[i] = 0.0
                    else:
                        pos_choice = self._draw(torch.ones_like(pos_idxs, dtype=torch.float), num_pos, len(pos_idxs) < num_pos)
                        neg_choice = self._draw(torch.ones_like(neg_idxs, dtype=torch.float), num_neg, len(neg_idxs) < num_neg)
                        sample_idx = torch.cat([pos_idxs[pos_choice], neg_idxs[neg_choice]], dim=0)
                elif self.use_potential:                                          # :289-305
                    potential = self.potential.get(info["occ_label_name"], torch.ones_like(valid_occ_grid, dtype=torch.float))
                    if n_valid < self.num_sample_points:
                        sample_idx = self._draw(1 / potential, self.num_sample_points, True)
                    else:
                        _, sample_idx = torch.topk(potential, self.num_sample_points, dim=0, largest=False)
                    potential[sample_idx] += 1
                    self.potential[info["occ_label_name"]] = potential
                elif self.weighted_sample:                                        # :306-335
                    try:
                        sample_weights = torch.ones_like(valid_occ_grid) * (1 - self.pos_sample_weight)
                        sample_weights[valid_occ_grid == 1] = self.pos_sample_weight
                        sample_idx = self._draw(sample_weights, self.num_sample_points, n_valid < self.num_sample_points)
                    except Exception:
                        sample_idx = self._draw(torch.ones_like(valid_occ_grid, dtype=torch.float),
                                                self.num_sample_points, n_valid < self.num_sample_points)
                else:
                    sample_idx = self._draw(torch.ones_like(valid_occ_grid, dtype=torch.float), self.num_sample_points,
                                            n_valid < self.num_sample_points)
                sample_centered = valid_voxel_centers[sample_idx]
                sample_occ = valid_occ_grid[sample_idx]
            sample_occs.append(sample_occ)
            sample_occ_centers.append(sample_centered)
            occ_sizes.append(torch.tensor([w, l, h], dtype=torch.float32, device=dev))
        if self.num_sample_points != -1:
            results["sample_occs"] = torch.stack(sample_occs, dim=0)
            results["sample_occ_centers"] = torch.stack(sample_occ_centers, dim=0)
        else:
            results["sample_occs"] = sample_occs
            results["sample_occ_centers"] = sample_occ_centers
        results["occ_sizes"] = torch.stack(occ_sizes, dim=0)
        return results
