"""oracle/build_ref.py -- compile the reference's own CPU ops in place.  TEST INFRASTRUCTURE.

Builds oracle/_ref/detmatch_ref_cpu.so from
  /root/reference/mmdet3d/ops/voxel/src/voxelization_cpu.cpp
  /root/reference/mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cpu.cpp
plus oracle/ref_shim.cpp (ours), and oracle/_ref/pcdet_ref_cpu.so from
  /root/reference/thirdparty/Spconv-OpenPCDet/pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp
plus oracle/pcdet_stub.cpp (ours).  Nothing from /root/reference is copied into the repo;
oracle/_ref/ is git-ignored but travels to the GPU box with the gpurun snapshot.

The reference's setup.py cannot build these two extensions without CUDA sources
(setup.py:199-219 lists the .cu files unconditionally), so the translation units are
compiled directly.  Flags: -O2 (what torch's cpp_extension / the reference's setup.py use,
setup.py:41-47 adds no -O flag so distutils' default -O2/-O3 applies) and
-ffp-contract=off (a no-op on baseline x86-64, pinned so that -march changes or an aarch64
host cannot silently fuse the rotation).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DETMATCH_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
NAME = "detmatch_ref_cpu"

SOURCES = [
    os.path.join(HERE, "ref_shim.cpp"),
    os.path.join(REF, "mmdet3d/ops/voxel/src/voxelization_cpu.cpp"),
    os.path.join(REF, "mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cpu.cpp"),
]


# The OpenPCDet CPU op (second module: the file carries its own PYBIND11_MODULE).  Its CUDA
# launchers are declared in the file and defined in a .cu that is not compiled here; the stub
# (ours) satisfies the linker and aborts if ever called.
PCDET_NAME = "pcdet_ref_cpu"
PCDET_SOURCES = [
    os.path.join(HERE, "pcdet_stub.cpp"),
    os.path.join(REF, "thirdparty/Spconv-OpenPCDet/pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp"),
]


def built_path():
    return os.path.join(OUT, NAME + ".so")


def pcdet_built_path():
    return os.path.join(OUT, PCDET_NAME + ".so")


def build_pcdet(verbose=False):
    if not all(os.path.exists(s) for s in PCDET_SOURCES):
        return pcdet_built_path() if os.path.exists(pcdet_built_path()) else None
    if os.path.exists(pcdet_built_path()) and all(
            os.path.getmtime(pcdet_built_path()) >= os.path.getmtime(s) for s in PCDET_SOURCES):
        return pcdet_built_path()
    os.makedirs(os.path.join(OUT, "pcdet"), exist_ok=True)
    from torch.utils import cpp_extension
    cpp_extension.load(name=PCDET_NAME, sources=PCDET_SOURCES,
                       extra_cflags=["-O2", "-ffp-contract=off", "-fno-fast-math", "-w"],
                       build_directory=os.path.join(OUT, "pcdet"), verbose=verbose, is_python_module=False)
    import shutil
    shutil.copy(os.path.join(OUT, "pcdet", PCDET_NAME + ".so"), pcdet_built_path())
    return pcdet_built_path()


def build(verbose=False):
    """Returns the path of the built module, or None when /root/reference is absent."""
    if not all(os.path.exists(s) for s in SOURCES):
        return built_path() if os.path.exists(built_path()) else None
    if os.path.exists(built_path()) and all(
            os.path.getmtime(built_path()) >= os.path.getmtime(s) for s in SOURCES):
        return built_path()
    os.makedirs(OUT, exist_ok=True)
    from torch.utils import cpp_extension
    cpp_extension.load(
        name=NAME,
        sources=SOURCES,
        extra_cflags=["-O2", "-ffp-contract=off", "-fno-fast-math"],
        build_directory=OUT,
        verbose=verbose,
        is_python_module=False,  # only build; oracle/ref.py loads it by path
    )
    return built_path()


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print(p if p else "reference tree not found; nothing built")
    p = build_pcdet(verbose="-v" in sys.argv)
    print(p if p else "reference tree not found; pcdet module not built")
