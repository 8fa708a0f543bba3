"""Mirror of mmdet3d/ops/voxel/scatter_points.py:9-99: ``dynamic_scatter`` (autograd Function with
the reference's forward / backward contract) and ``DynamicScatter`` with the same constructor,
``forward_single``, ``forward`` and repr, on the sm_100a kernels (csrc/scatter.cu).

Voxels come out in the order ``at::unique_dim(sorted=True)`` gives the reference
(scatter_points_cuda.cu:204-210): lexicographic in the coordinate columns.  ``max`` is bit-exact
and deterministic; ``sum`` / ``mean`` use float atomics like the reference, so their last bits
depend on arrival order there and here.
"""
import ctypes

import torch
from torch import nn
from torch.autograd import Function

from ... import _cabi
from ..._torch_glue import ptr, stream_ptr, workspace

_REDUCE = {"sum": 0, "mean": 1, "max": 2}  # scatter_points_cuda.cu:7 / voxelization.h:85-94


def dynamic_point_to_voxel_forward(feats, coors, reduce_type, dims=None):
    """voxelization.h:96-108 -> (voxel_feats, voxel_coors, point2voxel_map, voxel_points_count).

    ``dims`` (exclusive upper bounds of the coordinate columns) sizes the occupancy bitmap; by
    default the column maxima are read back from the device (one extra host synchronisation)."""
    if reduce_type not in _REDUCE:
        raise ValueError(f"reduce_type must be one of {sorted(_REDUCE)}, got {reduce_type!r}")
    if not (feats.is_cuda and coors.is_cuda):
        raise RuntimeError("dynamic_scatter: CUDA tensors only (detmatch_b200 has no CPU path)")
    if feats.dtype != torch.float32 or coors.dtype != torch.int32:
        raise TypeError("dynamic_scatter: feats must be float32 and coors int32")
    feats, coors = feats.contiguous(), coors.contiguous()
    n, c = feats.shape
    ndim = coors.size(1)
    assert coors.size(0) == n
    dev = feats.device
    if n == 0:  # scatter_points_cuda.cu:193-197
        return (feats.clone().detach(), coors.clone().detach(), coors.new_empty((0,), dtype=torch.int32),
                coors.new_empty((0,), dtype=torch.int32))
    if dims is None:
        dims = [max(int(v) + 1, 0) for v in coors.amax(dim=0).tolist()]
    cd = (ctypes.c_int32 * ndim)(*[int(d) for d in dims])
    L = _cabi.lib()
    ws = workspace(dev, L.pcfe_dynamic_scatter_workspace_bytes(cd, ndim, n))
    coors_map = torch.empty((n,), dtype=torch.int32, device=dev)
    num = torch.empty((1,), dtype=torch.int32, device=dev)
    _cabi.check(L.pcfe_dynamic_scatter_map_i32(ptr(coors), n, ndim, cd, ptr(coors_map), ptr(num), ptr(ws), ws.numel(),
                                               dev.index, stream_ptr(dev)), "pcfe_dynamic_scatter_map_i32")
    m = int(num.item())  # the outputs are sized by the voxel count (the reference syncs in unique_dim)
    voxel_feats = torch.empty((m, c), dtype=torch.float32, device=dev)
    voxel_coors = torch.empty((m, ndim), dtype=torch.int32, device=dev)
    count = torch.empty((m,), dtype=torch.int32, device=dev)
    _cabi.check(L.pcfe_dynamic_scatter_reduce_f32(ptr(feats), ptr(coors), ptr(coors_map), n, c, ndim, _REDUCE[reduce_type], m,
                                                  ptr(voxel_feats), ptr(voxel_coors), ptr(count), dev.index,
                                                  stream_ptr(dev)), "pcfe_dynamic_scatter_reduce_f32")
    return voxel_feats, voxel_coors, coors_map, count


def dynamic_point_to_voxel_backward(grad_feats, grad_voxel_feats, feats, voxel_feats, point2voxel_map,
                                    voxel_points_count, reduce_type):
    """voxelization.h:110-123: fills ``grad_feats`` (N, C) in place."""
    n, c = feats.shape
    m = voxel_feats.size(0)
    dev = feats.device
    if n == 0:
        return
    L = _cabi.lib()
    ws = workspace(dev, max(m * c * 4, 256))
    _cabi.check(L.pcfe_dynamic_scatter_backward_f32(ptr(grad_voxel_feats), ptr(feats), ptr(voxel_feats), ptr(point2voxel_map),
                                                    ptr(voxel_points_count), n, m, c, _REDUCE[reduce_type], ptr(grad_feats),
                                                    ptr(ws), ws.numel(), dev.index, stream_ptr(dev)),
                "pcfe_dynamic_scatter_backward_f32")


class _dynamic_scatter(Function):

    @staticmethod
    def forward(ctx, feats, coors, reduce_type='max'):
        """scatter_points.py:11-37: feats (N, C), coors (N, ndim) -> voxel_feats (M, C), voxel_coors (M, ndim)."""
        results = dynamic_point_to_voxel_forward(feats, coors, reduce_type)
        (voxel_feats, voxel_coors, point2voxel_map, voxel_points_count) = results
        ctx.reduce_type = reduce_type
        ctx.save_for_backward(feats, voxel_feats, point2voxel_map, voxel_points_count)
        ctx.mark_non_differentiable(voxel_coors)
        return voxel_feats, voxel_coors

    @staticmethod
    def backward(ctx, grad_voxel_feats, grad_voxel_coors=None):
        (feats, voxel_feats, point2voxel_map, voxel_points_count) = ctx.saved_tensors
        grad_feats = torch.empty_like(feats)  # every element is written by the kernels
        dynamic_point_to_voxel_backward(grad_feats, grad_voxel_feats.contiguous(), feats, voxel_feats, point2voxel_map,
                                        voxel_points_count, ctx.reduce_type)
        return grad_feats, None, None


dynamic_scatter = _dynamic_scatter.apply


class DynamicScatter(nn.Module):
    """scatter_points.py:50-99.  ``forward`` with batched (N, 4) coordinates runs ONE scatter over
    (batch, z, y, x) keys instead of the reference's per-sample loop: lexicographic order over the
    four columns is the concatenation of the per-sample results the loop produces."""

    def __init__(self, voxel_size, point_cloud_range, average_points: bool):
        super(DynamicScatter, self).__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.average_points = average_points

    def forward_single(self, points, coors):
        reduce = 'mean' if self.average_points else 'max'
        return dynamic_scatter(points.contiguous(), coors.contiguous(), reduce)

    def forward(self, points, coors):
        if coors.size(-1) == 3:
            return self.forward_single(points, coors)
        # scatter_points.py:82-94: batch_size = coors[-1, 0] + 1; samples beyond it are not visited.
        batch_size = int(coors[-1, 0]) + 1
        keep = coors[:, 0] < batch_size
        if not bool(keep.all()):
            points, coors = points[keep], coors[keep]
        return self.forward_single(points, coors)

    def __repr__(self):
        tmpstr = self.__class__.__name__ + '('
        tmpstr += 'voxel_size=' + str(self.voxel_size)
        tmpstr += ', point_cloud_range=' + str(self.point_cloud_range)
        tmpstr += ', average_points=' + str(self.average_points)
        tmpstr += ')'
        return tmpstr
