"""objectcentricocccompletion_b200 -- B200 (sm_100a) kernels for the point -> object-centric
occupancy hot path of Ghostish/ObjectCentricOccCompletion, behind the reference's own operator
names.  See DESIGN.md (scope, kernels, rooflines) and include/occ_b200.h (the C ABI).

No CPU fallback: the operators need libocc_b200.so and a CUDA device and raise otherwise.
"""
import os as _os

# OCCB200_HOST_ONLY=1: only the host-side modules (synth, waymo_io's packing) are wanted -- bench.py's reference arm
# times the CPU port and must not map the CUDA library.  Nothing of the operator surface is importable then.
if _os.environ.get("OCCB200_HOST_ONLY", "0") == "1":
    __all__ = []
else:
    from ._lib import lib as _load_lib

    _load_lib()          # fail loudly at import time when the CUDA library has not been built

    from .occ_annotate import (OccAnnotator, annotate_batch, pack_tracklets,  # noqa: E402
                               point_cloud_to_range_image_idx)
    from .occ_ops import generate_dense_voxel_centers, quantize_points  # noqa: E402
    from .points_in_boxes import points_in_boxes_batch, points_in_boxes_gpu  # noqa: E402
    from .range_image import build_range_images  # noqa: E402
    from .sst_ops import scatter_v2  # noqa: E402
    from .voxel import DynamicScatter, Voxelization, dynamic_scatter, voxelization  # noqa: E402

    __all__ = ["Voxelization", "voxelization", "DynamicScatter", "dynamic_scatter", "points_in_boxes_gpu",
               "points_in_boxes_batch", "scatter_v2", "quantize_points", "generate_dense_voxel_centers",
               "OccAnnotator", "annotate_batch", "pack_tracklets", "point_cloud_to_range_image_idx", "build_range_images"]
