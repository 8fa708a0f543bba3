"""Mirror of mmdet3d/ops/roiaware_pool3d/__init__.py (hot-path part; RoIAwarePool3d is a
"next" row of SURVEY.md section 8(f))."""
from .box_adapters import depth_boxes_points_in_boxes, depth_boxes_to_lidar, lidar_boxes_points_in_boxes
from .points_in_boxes import points_in_boxes_batch, points_in_boxes_cpu, points_in_boxes_gpu

__all__ = ["points_in_boxes_batch", "points_in_boxes_cpu", "points_in_boxes_gpu", "depth_boxes_points_in_boxes",
           "depth_boxes_to_lidar", "lidar_boxes_points_in_boxes"]
