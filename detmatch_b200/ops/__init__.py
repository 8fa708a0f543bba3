"""Mirror of the hot-path part of mmdet3d/ops/__init__.py:19-23."""
from .roiaware_pool3d import RoIAwarePool3d, points_in_boxes_batch, points_in_boxes_cpu, points_in_boxes_gpu
from .voxel import (DynamicScatter, dynamic_scatter, HostVoxelizePipeline, Voxelization, voxelization, voxelize_batch,
                    voxelize_batch_host, voxelize_batch_packed)
from .voxel_encoders import HardSimpleVFE, hard_simple_vfe, voxelize_mean_batch

__all__ = ["DynamicScatter", "dynamic_scatter", "Voxelization", "voxelization", "voxelize_batch", "voxelize_batch_host", "voxelize_batch_packed", "HostVoxelizePipeline", "HardSimpleVFE",
           "hard_simple_vfe", "voxelize_mean_batch", "points_in_boxes_batch",
           "points_in_boxes_cpu", "points_in_boxes_gpu", "RoIAwarePool3d"]
