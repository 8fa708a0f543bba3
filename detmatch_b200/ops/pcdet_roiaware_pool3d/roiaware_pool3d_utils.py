"""Mirror of thirdparty/Spconv-OpenPCDet/pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py:9-41:
same function names, argument order (points, boxes), asserts, numpy pass-through and outputs."""
import numpy as np
import torch

from ..._torch_glue import to_device
from . import roiaware_pool3d_cuda


def _check_numpy_to_torch(x):
    # pcdet/utils/common_utils.py check_numpy_to_torch
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x).float(), True
    return x, False


def points_in_boxes_cpu(points, boxes):
    """
    Args:
        points: (num_points, 3)
        boxes: [x, y, z, dx, dy, dz, heading], (x, y, z) is the box center, each box DO NOT overlaps
    Returns:
        point_indices: (N, num_points)
    """
    assert boxes.shape[1] == 7
    assert points.shape[1] == 3
    points, is_numpy = _check_numpy_to_torch(points)
    boxes, is_numpy = _check_numpy_to_torch(boxes)
    dpoints, home = to_device(points.float().contiguous())
    dboxes, _ = to_device(boxes.float().contiguous(), dpoints.device)
    point_indices = torch.empty((boxes.shape[0], points.shape[0]), dtype=torch.int, device=dpoints.device)
    roiaware_pool3d_cuda.points_in_boxes_cpu(dboxes, dpoints, point_indices)
    point_indices = point_indices.to(home)
    return point_indices.numpy() if is_numpy else point_indices


def points_in_boxes_gpu(points, boxes):
    """
    :param points: (B, M, 3)
    :param boxes: (B, T, 7), num_valid_boxes <= T
    :return box_idxs_of_pts: (B, M), default background = -1
    """
    assert boxes.shape[0] == points.shape[0]
    assert boxes.shape[2] == 7 and points.shape[2] == 3
    batch_size, num_points, _ = points.shape
    dpoints, home = to_device(points.float().contiguous())
    dboxes, _ = to_device(boxes.float().contiguous(), dpoints.device)
    box_idxs_of_pts = torch.empty((batch_size, num_points), dtype=torch.int, device=dpoints.device)
    roiaware_pool3d_cuda.points_in_boxes_gpu(dboxes, dpoints, box_idxs_of_pts)
    return box_idxs_of_pts.to(home)
