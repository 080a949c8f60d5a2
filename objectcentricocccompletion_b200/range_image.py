"""Range-image builder: the ``*_RANGE_IMAGE_MERGE_VIRTUAL`` step of the reference's
converter (tools/data_converter/waymo_converter.py:632-670) on the GPU.

The reference calls ``waymo_open_dataset.utils.range_image_utils.build_range_image_from_point_cloud`` (TF,
v1.2.0, not vendored); ``build_range_images`` follows that function's published algorithm -- it is the forward
use of the projection the annotate path inverts.  As there, the extrinsic inverse and the azimuth correction
are evaluated on the host (here in f64, as TF does after its cast).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import _lib


def _mono(t: np.ndarray) -> int:
    d = np.diff(t.astype(np.float64))
    return -1 if (d < 0).all() else (1 if (d > 0).all() else 0)


def build_range_images(points: Sequence, extrinsics: np.ndarray, inclinations: Sequence[np.ndarray],
                       sizes: Sequence[Sequence[int]], device=None) -> List[torch.Tensor]:
    """One range image per entry.

    points[b]        f32 [n_b, >=3] vehicle-frame returns of image b (numpy or CUDA tensor; both returns of the
                     LiDAR concatenated, waymo_converter.py:641-652)
    extrinsics       [B,4,4] LiDAR -> vehicle (f32 as stored; cast to f64 like TF does)
    inclinations[b]  f32 [H_b] as stored (ascending); reversed here (:659)
    sizes[b]         (H_b, W_b)
    -> list of f32 [H_b, W_b] CUDA tensors: min range per pixel, 0 where no return (``ri`` of :662-668).
    Raises if a point's column falls outside [0, W) (the TF op asserts that).
    """
    _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    B = len(points)
    assert len(inclinations) == B and len(sizes) == B and len(extrinsics) == B
    if B == 0:
        return []
    desc = np.zeros(B, _lib.RI_DESC_DTYPE)
    E = np.asarray(extrinsics).astype(np.float64)
    v2l = np.linalg.inv(E)                                     # tf.linalg.inv after the f64 cast
    desc["v2l"] = v2l[:, :3, :].reshape(B, 12)
    desc["azc"] = np.arctan2(E[:, 1, 0], E[:, 0, 0])
    incl, incl_off, ri_off, pt_off = [], 0, 0, [0]
    for b in range(B):
        H, W = (int(v) for v in sizes[b])
        t = np.ascontiguousarray(np.asarray(inclinations[b], np.float32)[::-1])
        assert t.shape == (H,)
        incl.append(t)
        desc["incl_off"][b], desc["ri_off"][b], desc["H"][b], desc["W"][b], desc["mono"][b] = incl_off, ri_off, H, W, _mono(t)
        incl_off += H
        ri_off += H * W
        pt_off.append(pt_off[-1] + int(points[b].shape[0]))
    strides = {int(p.shape[1]) for p in points}
    assert len(strides) == 1, "all point arrays must have the same number of columns"
    pts = [p if torch.is_tensor(p) else torch.from_numpy(np.ascontiguousarray(p, np.float32)) for p in points]
    pts_d = torch.cat([p.to(device=dev, dtype=torch.float32) for p in pts], 0).contiguous()
    if pts_d.shape[0] == 0:
        pts_d = torch.zeros((1, strides.pop()), dtype=torch.float32, device=dev)
    desc_d = torch.from_numpy(desc.view(np.uint8)).to(dev)
    incl_d = torch.from_numpy(np.concatenate(incl)).to(dev)
    off_d = torch.tensor(pt_off, dtype=torch.int64, device=dev)
    ri = torch.empty(max(ri_off, 1), dtype=torch.float32, device=dev)
    n_bad = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().occb200_build_range_images(pts_d.data_ptr(), int(pts_d.shape[1]), off_d.data_ptr(),
                                                   int(np.diff(pt_off).max()), desc_d.data_ptr(), B,
                                                   incl_d.data_ptr(), ri.data_ptr(), ri_off, n_bad.data_ptr(),
                                                   _lib.stream_ptr(dev))
    _lib.check(rc, "occb200_build_range_images")
    if int(n_bad.item()):
        raise RuntimeError(f"{int(n_bad.item())} points project outside the image columns "
                           "(build_range_image_from_point_cloud asserts 0 <= col < W)")
    return [ri[int(desc["ri_off"][b]): int(desc["ri_off"][b]) + int(desc["H"][b]) * int(desc["W"][b])]
            .view(int(desc["H"][b]), int(desc["W"][b])) for b in range(B)]


def merge_virtual(frame_dict: dict, lidar_names: Sequence[str], device=None) -> dict:
    """``convert_one``'s inner loop (waymo_converter.py:632-668) for one frame dictionary: adds
    ``<LIDAR>_RANGE_IMAGE_MERGE_VIRTUAL`` (numpy f32 [H,W]) for every LiDAR, all images in one call."""
    pts, ext, inc, sizes = [], [], [], []
    for name in lidar_names:
        ri0, ri1 = frame_dict[f"{name}_RANGE_IMAGE_FIRST_RETURN"], frame_dict[f"{name}_RANGE_IMAGE_SECOND_RETURN"]
        h, w = ri0.shape[:2]
        p0 = ri0[..., 3:].reshape(-1, 3)[ri0[..., 0].reshape(-1) > 0]       # :641-646
        p1 = ri1[..., 3:].reshape(-1, 3)[ri1[..., 0].reshape(-1) > 0]
        pts.append(np.concatenate([p0, p1], 0).astype(np.float32))
        ext.append(frame_dict[f"{name}_LIDAR_EXTRINSIC"])
        inc.append(frame_dict[f"{name}_BEAM_INCLINATION"])
        sizes.append((h, w))
    imgs = build_range_images(pts, np.stack(ext, 0), inc, sizes, device)
    for name, img in zip(lidar_names, imgs):
        frame_dict[f"{name}_RANGE_IMAGE_MERGE_VIRTUAL"] = img.cpu().numpy()
    return frame_dict
