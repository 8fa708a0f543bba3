"""Mirror of the hot-path part of mmdet3d/ops/__init__.py:19-23."""
from .roiaware_pool3d import points_in_boxes_batch, points_in_boxes_cpu, points_in_boxes_gpu
from .voxel import Voxelization, voxelization, voxelize_batch

__all__ = ["Voxelization", "voxelization", "voxelize_batch", "points_in_boxes_batch",
           "points_in_boxes_cpu", "points_in_boxes_gpu"]
