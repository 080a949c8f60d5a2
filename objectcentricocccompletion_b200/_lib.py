"""ctypes binding of ``csrc/libocc_b200.so`` -- the C ABI declared in ``include/occ_b200.h``.

There is no CPU fallback: if the shared library is missing, importing this module raises, and if
CUDA is missing every operator raises.  ``build()`` (re)compiles the library with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("OCCB200_LIB", os.path.join(CSRC, "libocc_b200.so"))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "occ_b200.h")

vp, i32, i64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
ABI_VERSION = 7


class Pose(C.Structure):
    _fields_ = [("box", f32 * 7), ("cos_pib", f32), ("sin_pib", f32), ("cos_m", f32), ("sin_m", f32),
                ("cos_p", f32), ("sin_p", f32), ("pad", f32 * 3)]


class Sensor(C.Structure):
    _fields_ = [("ri_off", i64), ("incl_off", i64), ("H", i32), ("W", i32), ("v2l", f32 * 12), ("azc", f32),
                ("incl_mono", i32)]


class AnnotateArgs(C.Structure):
    _fields_ = [("T", i32), ("L", i32), ("F", i64), ("trk_frame_off", vp), ("poses", vp), ("frame_sf", vp),
                ("points", vp), ("point_stride", i32), ("pad0", i32), ("frame_pt_off", vp), ("sensors", vp),
                ("SF", i64), ("incl_pool", vp), ("incl_len", i64), ("ri_pool", vp), ("pyr_tiles", i64), ("items_cap", i64), ("voxel_size", f64), ("label_off", vp), ("labels", vp),
                ("dims", vp), ("sizes", vp), ("status", vp), ("n_unknown", vp), ("n_steps", vp), ("workspace", vp),
                ("workspace_bytes", i64), ("flags", i32), ("pad1", i32), ("max_label_slots", i64),
                ("labels_u8", vp), ("frame_trk", vp), ("pyr_off", vp), ("table_off", vp), ("table_H", vp),
                ("n_tables", i32), ("max_pairs", i32), ("brick_off", vp), ("bricks", i64), ("n_points", i64),
                ("ri_tile_live", vp)]


POSE_DTYPE = np.dtype([("box", "<f4", (7,)), ("cos_pib", "<f4"), ("sin_pib", "<f4"), ("cos_m", "<f4"),
                       ("sin_m", "<f4"), ("cos_p", "<f4"), ("sin_p", "<f4"), ("pad", "<f4", (3,))])
SENSOR_DTYPE = np.dtype([("ri_off", "<i8"), ("incl_off", "<i8"), ("H", "<i4"), ("W", "<i4"), ("v2l", "<f4", (12,)),
                         ("azc", "<f4"), ("incl_mono", "<i4")])
assert POSE_DTYPE.itemsize == C.sizeof(Pose) == 64
assert SENSOR_DTYPE.itemsize == C.sizeof(Sensor) == 80

CAND_BOX_DTYPE = np.dtype([("cx", "<f4"), ("cy", "<f4"), ("cz", "<f4"), ("r2", "<f4"), ("tf", "<i4"), ("pad", "<i4", (3,))])
assert CAND_BOX_DTYPE.itemsize == 32

RI_DESC_DTYPE = np.dtype([("v2l", "<f8", (12,)), ("azc", "<f8"), ("incl_off", "<i8"), ("ri_off", "<i8"), ("H", "<i4"),
                          ("W", "<i4"), ("mono", "<i4"), ("pad", "<i4")])
assert RI_DESC_DTYPE.itemsize == 136

# name -> (restype, argtypes); must list every symbol include/occ_b200.h declares
SIGNATURES = {
    "occb200_abi_version": (C.c_int, []),
    "occb200_struct_sizes": (None, [vp]),
    "occb200_last_error": (C.c_char_p, []),
    "occb200_launch_count": (i64, []),
    "occb200_profile_enable": (None, [C.c_int]),
    "occb200_profile_kinds": (C.c_int, []),
    "occb200_profile_read": (C.c_int, [vp, vp]),
    "occb200_selftest_atan2": (C.c_int, [i64, C.c_uint64, vp, vp]),
    "occb200_points_in_boxes_gpu": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "occb200_points_in_boxes_batch": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "occb200_host_box_trig": (None, [vp, i64, vp]),
    "occb200_dynamic_voxelize": (C.c_int, [vp, C.c_int, i64, C.c_int, vp, vp, vp, vp]),
    "occb200_hard_voxelize_workspace_bytes": (i64, [i64]),
    "occb200_hard_voxelize": (C.c_int, [vp, i64, C.c_int, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, i64, vp, vp]),
    "occb200_unique_workspace_bytes": (i64, [i64, C.c_int]),
    "occb200_unique_rows": (C.c_int, [vp, C.c_int, i64, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, i64, vp, vp]),
    "occb200_unique_rows_bounded": (C.c_int, [vp, C.c_int, i64, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp]),
    "occb200_plan_workspace_bytes": (i64, [i64]),
    "occb200_plan_from_inverse": (C.c_int, [vp, i64, i64, vp, vp, vp, vp, i64, vp]),
    "occb200_segment_reduce": (C.c_int, [vp, i64, C.c_int, vp, vp, vp, i64, C.c_int, vp, vp, vp]),
    "occb200_segment_reduce_backward": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, i64, i64, C.c_int, C.c_int, vp]),
    "occb200_segment_reduce_f64": (C.c_int, [vp, i64, C.c_int, vp, vp, vp, i64, C.c_int, vp, vp, vp]),
    "occb200_segment_reduce_backward_f64": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, i64, i64, C.c_int, C.c_int, vp]),
    "occb200_quantize_points": (C.c_int, [vp, i64, vp, i64, C.c_int, vp, f32, vp, vp, C.c_int, vp, vp, vp, vp]),
    "occb200_dense_voxel_centers": (C.c_int, [vp, vp, vp, C.c_int, i64, f32, vp, vp, vp, vp]),
    "occb200_observed_labels": (C.c_int, [vp, vp, i64, vp, vp, i64, vp, vp]),
    "occb200_mirror_occ_label": (C.c_int, [vp, vp, vp, vp, i32, i64, vp, vp]),
    "occb200_candidate_chunk": (C.c_int, []),
    "occb200_candidate_warps": (C.c_int, []),
    "occb200_select_candidates": (C.c_int, [vp, C.c_int, vp, i32, i64, vp, vp, vp, vp, vp, vp, C.c_int, vp]),
    "occb200_annotate_workspace_bytes": (i64, [i32, i64, i64, i64, i32, i64, i64, i64, i32]),
    "occb200_grid_bricks": (i64, [i32, i32, i32]),
    "occb200_pyramid_tiles": (i64, [i32, i32]),
    "occb200_annotate_batch": (C.c_int, [C.POINTER(AnnotateArgs), i64, vp]),
    "occb200_annotate_queue_stats": (C.c_int, [C.POINTER(AnnotateArgs), i64, vp, vp]),
    "occb200_annotate_point_voxels": (C.c_int, [C.POINTER(AnnotateArgs), i64, vp, vp, vp]),
    "occb200_point_pool": (C.c_int, [vp, vp, vp, i64, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]),
    "occb200_host_pose_pack": (None, [vp, vp, i64, vp]),
    "occb200_host_mask_to_blocks": (i64, [vp, i64, vp]),
    "occb200_host_gather_blocks": (C.c_int, [vp, i64, vp, vp, vp, i32, vp]),
    "occb200_scatter_blocks": (C.c_int, [vp, vp, i64, vp, i64, vp]),
    "occb200_window_mask_words": (i64, [i64]),
    "occb200_host_window_mark": (C.c_int, [i32, i32, vp, vp, vp, vp, i64, vp, vp, f64, i64, vp, f32]),
    "occb200_host_copy_parts": (C.c_int, [vp, vp, vp, i32, vp]),
    "occb200_pull_windows": (C.c_int, [C.POINTER(AnnotateArgs), vp, vp, vp, i64, vp, vp, vp, vp]),
    "occb200_build_range_images": (C.c_int, [vp, C.c_int, vp, i64, vp, i32, vp, vp, i64, vp, vp]),
    "occb200_point_cloud_to_range_image_idx": (C.c_int, [vp, C.c_int, i64, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp]),
}


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libocc_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean"], stdout=subprocess.DEVNULL)
    out = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("nvcc build of libocc_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA library has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)            # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        if L.occb200_abi_version() != ABI_VERSION:
            raise ImportError("libocc_b200.so ABI version mismatch; rebuild")
        sizes = (i64 * 4)()
        L.occb200_struct_sizes(C.addressof(sizes))
        mine = (C.sizeof(Pose), C.sizeof(Sensor), C.sizeof(AnnotateArgs), RI_DESC_DTYPE.itemsize)
        if tuple(sizes) != mine:
            raise ImportError(f"struct layout mismatch between _lib.py {mine} and libocc_b200.so {tuple(sizes)}")
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {lib().occb200_last_error().decode()}")


def require_cuda(*tensors):
    """Operators run on CUDA tensors only -- fail loudly instead of falling back."""
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("objectcentricocccompletion_b200 needs a CUDA device (sm_100a); there is no CPU path")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("expected a CUDA tensor; there is no CPU path in this package")


def stream_ptr(device=None) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def ptr(t) -> int:
    """Device/host address of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


def launch_count() -> int:
    return int(lib().occb200_launch_count())
