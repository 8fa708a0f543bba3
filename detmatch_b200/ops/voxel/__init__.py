"""Mirror of mmdet3d/ops/voxel/__init__.py (hot-path part)."""
from .scatter_points import DynamicScatter, dynamic_scatter
from .voxelize import HardVoxelizeBatchPlan, Voxelization, voxelization, voxelize_batch, voxelize_batch_packed

__all__ = ["DynamicScatter", "dynamic_scatter", "HardVoxelizeBatchPlan", "Voxelization", "voxelization", "voxelize_batch", "voxelize_batch_packed"]
