"""Build recipe for the parity oracle (TEST INFRASTRUCTURE, never the product).

* ``build_oracle()`` compiles ``oracle/occ_oracle.c`` (the C restatement) into
  ``oracle/_build/liboccoracle.so`` with gcc.  ``-ffp-contract=off`` so only the
  FMAs written explicitly in the source are fused.
* ``build_ref()`` compiles the reference's OWN CPU sources, from where they lie
  under ``/root/reference`` (nothing is copied into the repo), into
  ``oracle/_ref/``:
    - ``mmdet3d/ops/voxel/src/{voxelization.cpp,voxelization_cpu.cpp,scatter_points_cpu.cpp}``
      -> ``ref_voxel_layer``  (hard_voxelize / dynamic_voxelize CPU)
    - ``mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cpu.cpp`` + a 10-line
      pybind shim written here -> ``ref_points_in_boxes``
  They need only torch's headers (present in the image), no cmake and no
  generated code.  The rest of the path (tools/occ/occ_annotate.py) is Python and
  is exercised by oracle/torch_ref.py instead.
  ``/root/reference`` exists only in the build container; on the GPU box the
  prebuilt ``oracle/_ref/*.so`` files are used if present.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
REF_OUT = os.path.join(HERE, "_ref")
REFERENCE = "/root/reference"


def _newer(src, dst):
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def build_oracle(force: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, "occ_oracle.c")
    out = os.path.join(BUILD, "liboccoracle.so")
    if force or _newer(src, out):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=gnu11", "-ffp-contract=off", "-fno-fast-math",
               "-fopenmp", "-o", out, src, "-lm"]
        subprocess.check_call(cmd)
    return out


_PIB_SHIM = r"""
#include <torch/extension.h>
int points_in_boxes_cpu(at::Tensor boxes_tensor, at::Tensor pts_tensor, at::Tensor pts_indices_tensor);
PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("points_in_boxes_cpu", &points_in_boxes_cpu, "points_in_boxes_cpu (reference source)");
}
"""


def build_ref(verbose: bool = False):
    """Compile the reference's CPU ops into oracle/_ref (only where /root/reference exists)."""
    if not os.path.isdir(REFERENCE):
        return None
    os.makedirs(REF_OUT, exist_ok=True)
    from torch.utils.cpp_extension import load

    vdir = os.path.join(REFERENCE, "mmdet3d/ops/voxel/src")
    built = {}
    if not _has_ext("ref_voxel_layer"):
        load(name="ref_voxel_layer",
             sources=[os.path.join(vdir, f) for f in ("voxelization.cpp", "voxelization_cpu.cpp", "scatter_points_cpu.cpp")],
             build_directory=_mk(os.path.join(REF_OUT, "ref_voxel_layer")), verbose=verbose, extra_cflags=["-O2"])
    built["ref_voxel_layer"] = True
    if not _has_ext("ref_points_in_boxes"):
        shim_dir = _mk(os.path.join(REF_OUT, "ref_points_in_boxes"))
        shim = os.path.join(shim_dir, "pib_shim.cpp")
        with open(shim, "w") as f:
            f.write(_PIB_SHIM)
        load(name="ref_points_in_boxes",
             sources=[shim, os.path.join(REFERENCE, "mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cpu.cpp")],
             build_directory=shim_dir, verbose=verbose, extra_cflags=["-O2"])
    built["ref_points_in_boxes"] = True
    return built


_PIB_CUDA_SHIM = r"""
#include <torch/extension.h>
int points_in_boxes_gpu(at::Tensor boxes_tensor, at::Tensor pts_tensor, at::Tensor box_idx_of_points_tensor);
int points_in_boxes_batch(at::Tensor boxes_tensor, at::Tensor pts_tensor, at::Tensor box_idx_of_points_tensor);
PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("points_in_boxes_gpu", &points_in_boxes_gpu, "points_in_boxes_gpu (reference CUDA source)");
  m.def("points_in_boxes_batch", &points_in_boxes_batch, "points_in_boxes_batch (reference CUDA source)");
}
"""


def build_ref_cuda(verbose: bool = False):
    """The reference's own CUDA kernels, UNMODIFIED, compiled for sm_100a into oracle/_ref:
    `ref_voxel_layer_cuda` (voxelization_cuda.cu + scatter_points_cuda.cu + the CPU files, -DWITH_CUDA) and
    `ref_points_in_boxes_cuda` (points_in_boxes_cuda.cu + a shim).  They are the "reference on B200" bar of
    tools/bench_ops.py and a second oracle for the operators on the GPU box."""
    if not os.path.isdir(REFERENCE):
        return None
    os.makedirs(REF_OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load

    vdir = os.path.join(REFERENCE, "mmdet3d/ops/voxel/src")
    if not _has_ext("ref_voxel_layer_cuda"):
        load(name="ref_voxel_layer_cuda",
             sources=[os.path.join(vdir, f) for f in ("voxelization.cpp", "voxelization_cpu.cpp", "scatter_points_cpu.cpp",
                                                      "voxelization_cuda.cu", "scatter_points_cuda.cu")],
             build_directory=_mk(os.path.join(REF_OUT, "ref_voxel_layer_cuda")), verbose=verbose,
             extra_cflags=["-O2", "-DWITH_CUDA"], extra_cuda_cflags=["-O2", "-DWITH_CUDA"], with_cuda=True)
    if not _has_ext("ref_points_in_boxes_cuda"):
        d = _mk(os.path.join(REF_OUT, "ref_points_in_boxes_cuda"))
        shim = os.path.join(d, "pib_cuda_shim.cpp")
        with open(shim, "w") as f:
            f.write(_PIB_CUDA_SHIM)
        load(name="ref_points_in_boxes_cuda",
             sources=[shim, os.path.join(REFERENCE, "mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cuda.cu")],
             build_directory=d, verbose=verbose, extra_cflags=["-O2"], extra_cuda_cflags=["-O2"], with_cuda=True)
    return True


def _mk(d):
    os.makedirs(d, exist_ok=True)
    return d


def _has_ext(name):
    return os.path.exists(os.path.join(REF_OUT, name, name + ".so"))


def load_ref(name):
    """Import a prebuilt reference extension from oracle/_ref, or return None."""
    so = os.path.join(REF_OUT, name, name + ".so")
    if not os.path.exists(so):
        return None
    import importlib.util

    import torch  # noqa: F401  (the extension links against libtorch)

    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv))
    print(build_ref(verbose="-v" in sys.argv))
    if "--cuda" in sys.argv:
        print(build_ref_cuda(verbose="-v" in sys.argv))
