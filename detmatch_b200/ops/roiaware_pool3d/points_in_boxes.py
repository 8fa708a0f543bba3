"""Python face of the point-in-box ops, mirroring mmdet3d/ops/roiaware_pool3d/points_in_boxes.py:6-123:
the same three function names, the (points, boxes) argument order, the shape checks (AssertionError)
and the output dtypes / background values.

The inside test is the CPU op's (points_in_boxes_cpu.cpp:25-40), evaluated bit for bit on the GPU; the
reference's own CUDA kernels differ from its CPU op by device trig + FMA contraction (SURVEY.md
Appendix D).  Boxes are (x, y, z_bottom, w, l, h, yaw) in LiDAR coordinates throughout.
"""
import torch

from ..._torch_glue import to_device
from . import roiaware_pool3d_ext


def _check_batched(points, boxes):
    """Shape checks of the two batched ops (points_in_boxes.py:17-25,98-106)."""
    assert boxes.shape[0] == points.shape[0], (
        f'batch sizes differ: boxes {boxes.shape[0]}, points {points.shape[0]}')
    assert boxes.shape[2] == 7, f'boxes need 7 values per row, got {boxes.shape[2]}'
    assert points.shape[2] == 3, f'points need 3 values per row, got {points.shape[2]}'
    assert points.get_device() == boxes.get_device(), 'points and boxes live on different devices'


def _upload(points, boxes):
    dpoints, home = to_device(points.float().contiguous())
    dboxes, _ = to_device(boxes.float().contiguous(), dpoints.device)
    return dpoints, dboxes, home


def points_in_boxes_gpu(points, boxes):
    """points (B, M, 3), boxes (B, T, 7) -> (B, M) int32: for every point the LOWEST index of a box
    that contains it, -1 for background (points_in_boxes.py:6-50)."""
    _check_batched(points, boxes)
    dpoints, dboxes, home = _upload(points, boxes)
    # every element is written by the kernel, so no fill_(-1) pass (points_in_boxes.py:29-30)
    first_hit = torch.empty(points.shape[:2], dtype=torch.int, device=dpoints.device)
    roiaware_pool3d_ext.points_in_boxes_gpu(dboxes, dpoints, first_hit)
    return first_hit.to(home)


def points_in_boxes_cpu(points, boxes):
    """points (N, 3), boxes (T, 7) -> (T, N) int32 0/1 flags in the CPU op's box-major layout
    (points_in_boxes.py:53-82).  CPU tensors are accepted like in the reference: they are uploaded,
    tested on the GPU and the result comes back on the inputs' device."""
    assert boxes.shape[1] == 7, f'boxes need 7 values per row, got {boxes.shape[1]}'
    assert points.shape[1] == 3, f'points need 3 values per row, got {points.shape[1]}'
    dpoints, dboxes, home = _upload(points, boxes)
    flags = torch.empty((boxes.shape[0], points.shape[0]), dtype=torch.int, device=dpoints.device)
    roiaware_pool3d_ext.points_in_boxes_cpu(dboxes, dpoints, flags)
    return flags.to(home)


def points_in_boxes_batch(points, boxes):
    """points (B, M, 3), boxes (B, T, 7) -> (B, M, T) int32 0/1: one flag per (point, box) pair
    (points_in_boxes.py:85-123)."""
    _check_batched(points, boxes)
    dpoints, dboxes, home = _upload(points, boxes)
    flags = torch.empty((*points.shape[:2], boxes.shape[1]), dtype=torch.int, device=dpoints.device)
    roiaware_pool3d_ext.points_in_boxes_batch(dboxes, dpoints, flags)
    return flags.to(home)
