"""Drop-in for the reference's pybind module ``mmdet3d.ops.voxel.voxel_layer``
(mmdet3d/ops/voxel/src/voxelization.cpp:6-11): same function names, argument order and
ownership rules -- the caller allocates every output and the op writes in place.

Differences, all deliberate:
  * CUDA tensors only; results follow the reference's CPU semantics bit for bit;
  * rows of ``voxels``/``coors``/``num_points_per_voxel`` at index >= voxel_num are left as
    the caller passed them (the reference leaves them zero only because its wrapper pre-zeroes);
  * the only host synchronisation is reading back ``voxel_num`` for the int return value.
"""
import torch

from ... import _cabi
from ..._torch_glue import ptr, stream_ptr, workspace


def _check_points(points):
    if not points.is_cuda:
        raise RuntimeError("voxel_layer: points must be a CUDA tensor (no CPU fallback)")
    if points.dtype != torch.float32:
        raise TypeError(f"voxel_layer: points must be float32, got {points.dtype}")
    if points.dim() != 2 or points.size(1) < 3:
        raise RuntimeError("voxel_layer: points must have shape (N, C>=3)")
    if not points.is_contiguous():
        # voxelization_cuda.cu:192 CHECK_INPUT(points)
        raise RuntimeError("points must be contiguous")


def dynamic_voxelize(points, coors, voxel_size, coors_range, NDim=3):
    """voxelization.h:71-83.  coors (N, 3) int32 <- (z, y, x) or (-1, -1, -1)."""
    assert NDim == 3
    _check_points(points)
    assert coors.is_cuda and coors.dtype == torch.int32 and coors.is_contiguous()
    assert coors.shape == (points.size(0), 3)
    dev = points.device
    rc = _cabi.lib().pcfe_dynamic_voxelize_f32(ptr(points), points.size(0), points.size(1),
                                               _cabi.f3(voxel_size), _cabi.f6(coors_range), ptr(coors),
                                               dev.index, stream_ptr(dev))
    _cabi.check(rc, "pcfe_dynamic_voxelize_f32")


def hard_voxelize_async(points, voxels, coors, num_points_per_voxel, voxel_num, voxel_size,
                        coors_range, max_points, max_voxels):
    """Enqueues the op; ``voxel_num`` is a device int32[1] tensor.  No host synchronisation."""
    _check_points(points)
    dev = points.device
    n, c = points.size(0), points.size(1)
    need = _cabi.lib().pcfe_hard_voxelize_workspace_bytes(n, 1, 1, _cabi.f3(voxel_size), _cabi.f6(coors_range),
                                                          max_points, max_voxels)
    ws = workspace(dev, need)
    rc = _cabi.lib().pcfe_hard_voxelize_f32(ptr(points), n, c, _cabi.f3(voxel_size), _cabi.f6(coors_range),
                                            max_points, max_voxels, ptr(voxels), ptr(coors),
                                            ptr(num_points_per_voxel), ptr(voxel_num), ptr(ws), ws.numel(),
                                            dev.index, stream_ptr(dev))
    _cabi.check(rc, "pcfe_hard_voxelize_f32")


def hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range, max_points,
                  max_voxels, NDim=3):
    """voxelization.h:51-69.  Returns voxel_num as a Python int like the reference."""
    assert NDim == 3
    for t, dt in ((voxels, torch.float32), (coors, torch.int32), (num_points_per_voxel, torch.int32)):
        assert t.is_cuda and t.dtype == dt and t.is_contiguous()
    if max_points == -1:  # the reference's "unbounded" (voxelization_cpu.cpp:90): bounded by the buffer
        max_points = voxels.size(1)
    if max_voxels == -1:  # voxelization_cpu.cpp:78
        max_voxels = voxels.size(0)
    assert voxels.size(0) >= max_voxels and coors.size(0) >= max_voxels and num_points_per_voxel.size(0) >= max_voxels
    assert voxels.size(1) == max_points and voxels.size(2) == points.size(1)
    voxel_num = torch.empty(1, dtype=torch.int32, device=points.device)
    hard_voxelize_async(points, voxels, coors, num_points_per_voxel, voxel_num, voxel_size, coors_range,
                        max_points, max_voxels)
    return int(voxel_num.item())
