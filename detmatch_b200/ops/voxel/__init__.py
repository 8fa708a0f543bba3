"""Mirror of mmdet3d/ops/voxel/__init__.py (hot-path part)."""
from .scatter_points import DynamicScatter, dynamic_scatter
from .voxelize import HardVoxelizeBatchPlan, Voxelization, voxelization, voxelize_batch, voxelize_batch_packed
from .host_batch import HostVoxelizePipeline, voxelize_batch_host

__all__ = ["DynamicScatter", "dynamic_scatter", "HardVoxelizeBatchPlan", "HostVoxelizePipeline", "Voxelization", "voxelization",
           "voxelize_batch", "voxelize_batch_host", "voxelize_batch_packed"]
