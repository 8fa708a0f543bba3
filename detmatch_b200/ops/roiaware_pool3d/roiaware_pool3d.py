"""Mirror of mmdet3d/ops/roiaware_pool3d/roiaware_pool3d.py:9-110: ``RoIAwarePool3d`` /
``RoIAwarePool3dFunction`` with the reference's constructor arguments, forward signature, saved
context and gradient (only ``pts_feature`` receives one), executed by roiaware_pool3d.cu.

The three work tensors are allocated with ``empty`` instead of ``new_zeros``: the kernels write every
element that is ever read (for out_size = 14, 128 RoIs and max_pts_per_voxel = 128 the reference
zero-fills 180 MB per call, roiaware_pool3d.py:72-78).
"""
import torch
from torch import nn as nn
from torch.autograd import Function

from . import roiaware_pool3d_ext


class RoIAwarePool3d(nn.Module):

    def __init__(self, out_size, max_pts_per_voxel=128, mode='max'):
        super().__init__()
        """RoIAwarePool3d module

        Args:
            out_size (int or tuple): n or [n1, n2, n3]
            max_pts_per_voxel (int): m
            mode (str): 'max' or 'avg'
        """
        self.out_size = out_size
        self.max_pts_per_voxel = max_pts_per_voxel
        assert mode in ['max', 'avg']
        pool_method_map = {'max': 0, 'avg': 1}
        self.mode = pool_method_map[mode]

    def forward(self, rois, pts, pts_feature):
        """rois [N, 7] in LiDAR coordinates ((x, y, z) the bottom centre), pts [npoints, 3],
        pts_feature [npoints, C] -> pooled_features [N, out_x, out_y, out_z, C]."""
        return RoIAwarePool3dFunction.apply(rois, pts, pts_feature, self.out_size, self.max_pts_per_voxel, self.mode)


class RoIAwarePool3dFunction(Function):

    @staticmethod
    def forward(ctx, rois, pts, pts_feature, out_size, max_pts_per_voxel, mode):
        if isinstance(out_size, int):
            out_x = out_y = out_z = out_size
        else:
            assert len(out_size) == 3
            assert all(isinstance(v, int) for v in out_size)  # mmcv.is_tuple_of(out_size, int)
            out_x, out_y, out_z = out_size

        num_rois = rois.shape[0]
        num_channels = pts_feature.shape[-1]
        num_pts = pts.shape[0]

        rois = rois.contiguous().float()
        pts = pts.contiguous().float()
        feats = pts_feature.contiguous().float()
        pooled_features = feats.new_empty((num_rois, out_x, out_y, out_z, num_channels))
        argmax = torch.empty((num_rois, out_x, out_y, out_z, num_channels), dtype=torch.int, device=feats.device)
        pts_idx_of_voxels = torch.empty((num_rois, out_x, out_y, out_z, max_pts_per_voxel), dtype=torch.int,
                                        device=feats.device)

        roiaware_pool3d_ext.forward(rois, pts, feats, argmax, pts_idx_of_voxels, pooled_features, mode)

        ctx.roiaware_pool3d_for_backward = (pts_idx_of_voxels, argmax, mode, num_pts, num_channels)
        return pooled_features

    @staticmethod
    def backward(ctx, grad_out):
        pts_idx_of_voxels, argmax, mode, num_pts, num_channels = ctx.roiaware_pool3d_for_backward
        grad_in = grad_out.new_empty((num_pts, num_channels))  # zeroed by the kernel sequence
        roiaware_pool3d_ext.backward(pts_idx_of_voxels, argmax, grad_out.contiguous(), grad_in, mode)
        return None, None, grad_in, None, None, None
