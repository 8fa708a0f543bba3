"""Mirror of the mean voxel encoder that consumes hard voxelization
(mmdet3d/models/voxel_encoders/voxel_encoder.py:12-44, HardSimpleVFE)."""
from .voxel_encoder import HardSimpleVFE, hard_simple_vfe, voxelize_mean_batch

__all__ = ["HardSimpleVFE", "hard_simple_vfe", "voxelize_mean_batch"]
