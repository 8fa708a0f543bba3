"""Mirror of mmdet3d/ops/roiaware_pool3d/points_in_boxes.py: same three functions, argument
order (points, boxes), shape asserts, output dtype/fill and docstrings' semantics.

All three compute the CPU op's inside test (points_in_boxes_cpu.cpp:25-40) bit for bit on the
GPU; the reference's own CUDA kernels differ from its CPU op by device trig + FMA contraction
(SURVEY.md Appendix D).
"""
import torch

from ..._torch_glue import to_device
from . import roiaware_pool3d_ext


def points_in_boxes_gpu(points, boxes):
    """Find points that are in boxes.

    Args:
        points (torch.Tensor): [B, M, 3], [x, y, z] in LiDAR coordinate
        boxes (torch.Tensor): [B, T, 7], [x, y, z, w, l, h, ry] in LiDAR coordinate,
            (x, y, z) is the bottom center

    Returns:
        box_idxs_of_pts (torch.Tensor): (B, M), default background = -1
    """
    assert boxes.shape[0] == points.shape[0], \
        f'Points and boxes should have the same batch size, ' \
        f'got {boxes.shape[0]} and {points.shape[0]}'
    assert boxes.shape[2] == 7, \
        f'boxes dimension should be 7, ' \
        f'got unexpected shape {boxes.shape[2]}'
    assert points.shape[2] == 3, \
        f'points dimension should be 3, ' \
        f'got unexpected shape {points.shape[2]}'
    batch_size, num_points, _ = points.shape
    assert points.get_device() == boxes.get_device(), \
        'Points and boxes should be put on the same device'
    dpoints, home = to_device(points.float().contiguous())
    dboxes, _ = to_device(boxes.float().contiguous(), dpoints.device)
    # every element is written by the kernel; no fill_(-1) pass (points_in_boxes.py:29-30)
    box_idxs_of_pts = torch.empty((batch_size, num_points), dtype=torch.int, device=dpoints.device)
    roiaware_pool3d_ext.points_in_boxes_gpu(dboxes, dpoints, box_idxs_of_pts)
    return box_idxs_of_pts.to(home)


def points_in_boxes_cpu(points, boxes):
    """Find points that are in boxes, in the CPU op's (N_boxes, npoints) layout.

    Accepts CPU tensors like the reference (points_in_boxes.py:53-82); they are uploaded,
    tested on the GPU and the result is returned on the inputs' device.

    Args:
        points (torch.Tensor): [npoints, 3]
        boxes (torch.Tensor): [N, 7], in LiDAR coordinate, (x, y, z) is the bottom center

    Returns:
        point_indices (torch.Tensor): (N, npoints) int32 0/1
    """
    assert boxes.shape[1] == 7, \
        f'boxes dimension should be 7, ' \
        f'got unexpected shape {boxes.shape[1]}'
    assert points.shape[1] == 3, \
        f'points dimension should be 3, ' \
        f'got unexpected shape {points.shape[1]}'
    dpoints, home = to_device(points.float().contiguous())
    dboxes, _ = to_device(boxes.float().contiguous(), dpoints.device)
    point_indices = torch.empty((boxes.shape[0], points.shape[0]), dtype=torch.int, device=dpoints.device)
    roiaware_pool3d_ext.points_in_boxes_cpu(dboxes, dpoints, point_indices)
    return point_indices.to(home)


def points_in_boxes_batch(points, boxes):
    """Find points that are in boxes, all boxes per point.

    Args:
        points (torch.Tensor): [B, M, 3], [x, y, z] in LiDAR coordinate
        boxes (torch.Tensor): [B, T, 7], [x, y, z, w, l, h, ry] in LiDAR coordinate,
            (x, y, z) is the bottom center.

    Returns:
        box_idxs_of_pts (torch.Tensor): (B, M, T), default background = 0
    """
    assert boxes.shape[0] == points.shape[0], \
        f'Points and boxes should have the same batch size, ' \
        f'got {boxes.shape[0]} and {points.shape[0]}'
    assert boxes.shape[2] == 7, \
        f'boxes dimension should be 7, ' \
        f'got unexpected shape {boxes.shape[2]}'
    assert points.shape[2] == 3, \
        f'points dimension should be 3, ' \
        f'got unexpected shape {points.shape[2]}'
    batch_size, num_points, _ = points.shape
    num_boxes = boxes.shape[1]
    assert points.get_device() == boxes.get_device(), \
        'Points and boxes should be put on the same device'
    dpoints, home = to_device(points.float().contiguous())
    dboxes, _ = to_device(boxes.float().contiguous(), dpoints.device)
    box_idxs_of_pts = torch.empty((batch_size, num_points, num_boxes), dtype=torch.int, device=dpoints.device)
    roiaware_pool3d_ext.points_in_boxes_batch(dboxes, dpoints, box_idxs_of_pts)
    return box_idxs_of_pts.to(home)
