"""Mirror of mmdet3d/ops/roiaware_pool3d/__init__.py."""
from .box_adapters import depth_boxes_points_in_boxes, depth_boxes_to_lidar, lidar_boxes_points_in_boxes
from .roiaware_pool3d import RoIAwarePool3d
from .points_in_boxes import points_in_boxes_batch, points_in_boxes_cpu, points_in_boxes_gpu

__all__ = ["RoIAwarePool3d", "points_in_boxes_batch", "points_in_boxes_cpu", "points_in_boxes_gpu", "depth_boxes_points_in_boxes",
           "depth_boxes_to_lidar", "lidar_boxes_points_in_boxes"]
