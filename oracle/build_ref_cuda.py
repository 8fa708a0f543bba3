"""oracle/build_ref_cuda.py -- compile the reference's own CUDA ops for sm_100a.  MEASUREMENT
INFRASTRUCTURE (the GPU-vs-GPU column): oracle/_ref/detmatch_ref_cuda.so from
  /root/reference/mmdet3d/ops/voxel/src/voxelization_cuda.cu          (hard / dynamic voxelize, :184-371)
  /root/reference/mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cuda.cu   (:51-203)
plus oracle/ref_cuda_shim.cpp (ours).  The sources are compiled where they lie, unmodified; nothing is
copied into the repository; oracle/_ref/ is git-ignored but travels to the GPU box.  nvcc cross-compiles
without a GPU.  Never linked into or imported by the product (detmatch_b200/)."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DETMATCH_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
NAME = "detmatch_ref_cuda"
SOURCES = [
    os.path.join(HERE, "ref_cuda_shim.cpp"),
    os.path.join(REF, "mmdet3d/ops/voxel/src/voxelization_cuda.cu"),
    os.path.join(REF, "mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cuda.cu"),
]


def built_path():
    return os.path.join(OUT, NAME + ".so")


def build(verbose=False):
    """Returns the path of the built module, or None when /root/reference is absent."""
    if not all(os.path.exists(s) for s in SOURCES):
        return built_path() if os.path.exists(built_path()) else None
    if os.path.exists(built_path()) and all(os.path.getmtime(built_path()) >= os.path.getmtime(s) for s in SOURCES):
        return built_path()
    bdir = os.path.join(OUT, "cuda")
    os.makedirs(bdir, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils import cpp_extension
    cpp_extension.load(name=NAME, sources=SOURCES, extra_cflags=["-O2", "-w"],
                       extra_cuda_cflags=["-O3", "-w", "-gencode", "arch=compute_100a,code=sm_100a"],
                       build_directory=bdir, verbose=verbose, is_python_module=False, with_cuda=True)
    shutil.copy(os.path.join(bdir, NAME + ".so"), built_path())
    return built_path()


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print(p if p else "reference tree not found; nothing built")
