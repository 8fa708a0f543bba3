"""oracle/ref.py -- loader for oracle/_ref/detmatch_ref_cpu.so: the reference's UNMODIFIED
voxelization_cpu.cpp + points_in_boxes_cpu.cpp compiled in place (oracle/build_ref.py).

TEST INFRASTRUCTURE ONLY.  Wrappers mirror mmdet3d/ops/voxel/voxelize.py:41-58 and
mmdet3d/ops/roiaware_pool3d/points_in_boxes.py:53-82 (torch CPU tensors in and out).
"""
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "detmatch_ref_cpu.so")
_mod = None


def available():
    return os.path.exists(_SO)


def module():
    global _mod
    if _mod is None:
        import torch  # noqa: F401  (libtorch symbols must be loaded first)
        spec = importlib.util.spec_from_file_location("detmatch_ref_cpu", _SO)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _mod = m
    return _mod


def voxelization(points, voxel_size, coors_range, max_points=35, max_voxels=20000):
    """voxelize.py:13-58 (_Voxelization.forward) on CPU tensors."""
    import torch
    ext = module()
    if max_points == -1 or max_voxels == -1:
        coors = points.new_zeros(size=(points.size(0), 3), dtype=torch.int)
        ext.dynamic_voxelize(points, coors, list(voxel_size), list(coors_range), 3)
        return coors
    voxels = points.new_zeros(size=(max_voxels, max_points, points.size(1)))
    coors = points.new_zeros(size=(max_voxels, 3), dtype=torch.int)
    num = points.new_zeros(size=(max_voxels,), dtype=torch.int)
    voxel_num = ext.hard_voxelize(points, voxels, coors, num, list(voxel_size), list(coors_range),
                                  max_points, max_voxels, 3)
    return voxels[:voxel_num], coors[:voxel_num], num[:voxel_num]


def points_in_boxes_cpu(points, boxes):
    """points_in_boxes.py:53-82."""
    import torch
    ext = module()
    out = points.new_zeros((boxes.shape[0], points.shape[0]), dtype=torch.int)
    ext.points_in_boxes_cpu(boxes.float().contiguous(), points.float().contiguous(), out)
    return out


_PCDET_SO = os.path.join(_HERE, "_ref", "pcdet_ref_cpu.so")
_pcdet = None


def pcdet_available():
    return os.path.exists(_PCDET_SO)


def pcdet_points_in_boxes_cpu(points, boxes):
    """thirdparty/Spconv-OpenPCDet/pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py:9-25 on the
    reference's unmodified roiaware_pool3d.cpp (compiled in place): (N,3),(T,7) -> (T,N) int32."""
    global _pcdet
    import torch
    if _pcdet is None:
        spec = importlib.util.spec_from_file_location("pcdet_ref_cpu", _PCDET_SO)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _pcdet = m
    out = points.new_zeros((boxes.shape[0], points.shape[0]), dtype=torch.int)
    _pcdet.points_in_boxes_cpu(boxes.float().contiguous(), points.float().contiguous(), out)
    return out


_CUDA_SO = os.path.join(_HERE, "_ref", "detmatch_ref_cuda.so")
_cuda = None


def cuda_available():
    return os.path.exists(_CUDA_SO)


def cuda_module():
    """oracle/_ref/detmatch_ref_cuda.so: the reference's UNMODIFIED voxelization_cuda.cu and
    points_in_boxes_cuda.cu compiled for sm_100a (oracle/build_ref_cuda.py) -- the GPU-vs-GPU column."""
    global _cuda
    if _cuda is None:
        import torch  # noqa: F401
        spec = importlib.util.spec_from_file_location("detmatch_ref_cuda", _CUDA_SO)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _cuda = m
    return _cuda


def cuda_voxelization(points, voxel_size, coors_range, max_points=35, max_voxels=20000):
    """voxelize.py:13-58 (_Voxelization.forward) on CUDA tensors through the reference's own CUDA op."""
    import torch
    ext = cuda_module()
    if max_points == -1 or max_voxels == -1:
        coors = points.new_zeros(size=(points.size(0), 3), dtype=torch.int)
        ext.dynamic_voxelize(points, coors, list(voxel_size), list(coors_range), 3)
        return coors
    voxels = points.new_zeros(size=(max_voxels, max_points, points.size(1)))
    coors = points.new_zeros(size=(max_voxels, 3), dtype=torch.int)
    num = points.new_zeros(size=(max_voxels,), dtype=torch.int)
    voxel_num = ext.hard_voxelize(points, voxels, coors, num, list(voxel_size), list(coors_range),
                                  max_points, max_voxels, 3)
    return voxels[:voxel_num], coors[:voxel_num], num[:voxel_num]


_ROIAWARE_SO = os.path.join(_HERE, "_ref", "detmatch_ref_roiaware.so")
_roiaware = None


def roiaware_available():
    return os.path.exists(_ROIAWARE_SO)


def roiaware_module():
    """oracle/_ref/detmatch_ref_roiaware.so: the reference's own roiaware_pool3d_ext (forward / backward /
    points_in_boxes_*) compiled for sm_100a from its unmodified sources (oracle/build_ref_roiaware.py)."""
    global _roiaware
    if _roiaware is None:
        import torch  # noqa: F401
        spec = importlib.util.spec_from_file_location("detmatch_ref_roiaware", _ROIAWARE_SO)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _roiaware = m
    return _roiaware
