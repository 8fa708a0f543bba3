"""Mirror of mmdet3d/ops/roiaware_pool3d/__init__.py (hot-path part; RoIAwarePool3d is a
"next" row of SURVEY.md section 8(f))."""
from .points_in_boxes import points_in_boxes_batch, points_in_boxes_cpu, points_in_boxes_gpu

__all__ = ["points_in_boxes_batch", "points_in_boxes_cpu", "points_in_boxes_gpu"]
