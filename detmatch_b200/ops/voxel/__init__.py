"""Mirror of mmdet3d/ops/voxel/__init__.py (hot-path part)."""
from .voxelize import HardVoxelizeBatchPlan, Voxelization, voxelization, voxelize_batch, voxelize_batch_packed

__all__ = ["HardVoxelizeBatchPlan", "Voxelization", "voxelization", "voxelize_batch", "voxelize_batch_packed"]
