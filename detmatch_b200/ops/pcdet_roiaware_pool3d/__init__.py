"""Mirror of the point-in-box part of OpenPCDet's ``pcdet.ops.roiaware_pool3d``
(thirdparty/Spconv-OpenPCDet/pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py:9-41) -- the variant
PV-RCNN's point head calls (pcdet/models/dense_heads/point_head_template.py:82-89)."""
from . import roiaware_pool3d_cuda, roiaware_pool3d_utils
from .roiaware_pool3d_utils import points_in_boxes_cpu, points_in_boxes_gpu

__all__ = ["points_in_boxes_cpu", "points_in_boxes_gpu", "roiaware_pool3d_cuda", "roiaware_pool3d_utils"]
