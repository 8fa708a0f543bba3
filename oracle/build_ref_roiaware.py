"""oracle/build_ref_roiaware.py -- the reference's OWN roiaware_pool3d extension compiled for sm_100a.
TEST INFRASTRUCTURE: oracle/_ref/detmatch_ref_roiaware.so is the reference's pybind module
`roiaware_pool3d_ext` (forward, backward, points_in_boxes_{gpu,batch,cpu}) built from its unmodified sources
  /root/reference/mmdet3d/ops/roiaware_pool3d/src/{roiaware_pool3d.cpp, roiaware_pool3d_kernel.cu,
                                                    points_in_boxes_cpu.cpp, points_in_boxes_cuda.cu}
where they lie (no shim needed: roiaware_pool3d.cpp carries the PYBIND11_MODULE).  It pins the RoI-aware
pooling restatement of oracle/pcfe_oracle.c on the GPU box (tests/test_gpu_roiaware.py).  Nothing is copied
into the repository; oracle/_ref/ is git-ignored but travels with the gpurun snapshot."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DETMATCH_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
NAME = "detmatch_ref_roiaware"
SRC = os.path.join(REF, "mmdet3d/ops/roiaware_pool3d/src")
SOURCES = [os.path.join(SRC, f) for f in ("roiaware_pool3d.cpp", "roiaware_pool3d_kernel.cu", "points_in_boxes_cpu.cpp",
                                          "points_in_boxes_cuda.cu")]


def built_path():
    return os.path.join(OUT, NAME + ".so")


def build(verbose=False):
    if not all(os.path.exists(s) for s in SOURCES):
        return built_path() if os.path.exists(built_path()) else None
    if os.path.exists(built_path()) and all(os.path.getmtime(built_path()) >= os.path.getmtime(s) for s in SOURCES):
        return built_path()
    bdir = os.path.join(OUT, "roiaware")
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
