"""The two call sites through which DetMatch's box structures reach the point-in-box ops, as plain
functions on the boxes' ``.tensor`` (the structures themselves -- corners, rotation, flipping -- are
not on the point -> box assignment path):

* ``LiDARInstance3DBoxes.points_in_boxes``  mmdet3d/core/bbox/structures/lidar_box3d.py:258-270
* ``DepthInstance3DBoxes.points_in_boxes``  mmdet3d/core/bbox/structures/depth_box3d.py:251-277
  (axis swap of the points, DEPTH -> LIDAR conversion of the boxes, box_3d_mode.py:124-150)

A maintainer keeps the reference's structures and only swaps the imported ops;
these functions exist so that the flow -- batch dimension added and removed, boxes moved to the
points' device, depth -> LiDAR conversion -- is exercised against the oracle in this repository.
"""
import torch

from .points_in_boxes import points_in_boxes_batch, points_in_boxes_gpu


def lidar_boxes_points_in_boxes(boxes_tensor, points):
    """lidar_box3d.py:258-270: points (N, 3) -> (N,) index of the first box containing the point, -1 = none."""
    box_idx = points_in_boxes_gpu(points.unsqueeze(0), boxes_tensor.unsqueeze(0).to(points.device)).squeeze(0)
    return box_idx


def depth_boxes_to_lidar(boxes_tensor):
    """Box3DMode.convert(DEPTH -> LIDAR) with the default matrix (box_3d_mode.py:124-150):
    xyz' = xyz @ [[0, 1, 0], [-1, 0, 0], [0, 0, 1]]^T, sizes (y, x, z), remaining columns unchanged."""
    arr = boxes_tensor.clone()
    rt_mat = arr.new_tensor([[0, 1, 0], [-1, 0, 0], [0, 0, 1]])
    xyz = arr[:, :3] @ rt_mat.t()
    xyz_size = torch.cat([arr[..., 4:5], arr[..., 3:4], arr[..., 5:6]], dim=-1)
    return torch.cat([xyz[:, :3], xyz_size, arr[..., 6:]], dim=-1)


def depth_boxes_points_in_boxes(boxes_tensor, points):
    """depth_box3d.py:251-277: points (M, 3) or (1, M, 3) in depth coordinates -> (M, T) 0/1 flags."""
    points_lidar = points.clone()
    points_lidar = points_lidar[..., [1, 0, 2]]
    points_lidar[..., 1] *= -1
    if points.dim() == 2:
        points_lidar = points_lidar.unsqueeze(0)
    else:
        assert points.dim() == 3 and points_lidar.shape[0] == 1
    boxes_lidar = depth_boxes_to_lidar(boxes_tensor).to(points.device).unsqueeze(0)
    box_idxs_of_pts = points_in_boxes_batch(points_lidar, boxes_lidar)
    return box_idxs_of_pts.squeeze(0)
