"""HardSimpleVFE (mmdet3d/models/voxel_encoders/voxel_encoder.py:12-44) on the sm_100a kernels, and
its fusion into hard voxelization.

Arithmetic (include/pcfe.h, "mean voxel feature encoder"): float32, slot-order sum over all
max_points slots, one IEEE divide by float32(num_points).  ATen leaves the association of sum()
unspecified; tests pin these kernels bit for bit to oracle/vfe_mean.py and bound the distance to
the reference expression's own result.
"""
import ctypes

import torch
from torch import nn

from ... import _cabi
from ..._torch_glue import ptr, stream_ptr, workspace
from ..voxel import voxel_layer


def hard_simple_vfe(features, num_points, num_features=None, voxel_num=None):
    """features (N, M, C) float32 CUDA, num_points (N,) int32 -> (N, num_features) means.

    ``voxel_num`` (device int32 scalar tensor) limits the rows computed to ``min(voxel_num, N)``
    without a host synchronisation; rows beyond are unspecified.
    """
    if features.dtype != torch.float32:
        raise TypeError(f"hard_simple_vfe: features must be float32, got {features.dtype}")
    if not features.is_cuda:
        raise RuntimeError("hard_simple_vfe: CUDA tensors only (detmatch_b200 has no CPU path)")
    if features.dim() != 3 or num_points.dim() != 1 or num_points.size(0) != features.size(0):
        raise ValueError("hard_simple_vfe: features (N, M, C) and num_points (N,) expected")
    n, m, c = features.shape
    nf = c if num_features is None else int(num_features)
    if not 1 <= nf <= c:
        raise ValueError(f"hard_simple_vfe: num_features {nf} outside [1, {c}]")
    features = features.contiguous()
    if nf != c:  # the kernel reads (N, M, nf) rows
        features = features[:, :, :nf].contiguous()
    num_points = num_points.to(torch.int32).contiguous()
    out = torch.empty((n, nf), dtype=torch.float32, device=features.device)
    dev = features.device
    rc = _cabi.lib().pcfe_voxel_mean_f32(ptr(features), ptr(num_points), ptr(voxel_num) if voxel_num is not None else None,
                                         n, m, nf, ptr(out), dev.index, stream_ptr(dev))
    _cabi.check(rc, "pcfe_voxel_mean_f32")
    return out


class HardSimpleVFE(nn.Module):
    """voxel_encoder.py:12-44: same constructor argument and forward signature."""

    def __init__(self, num_features=4):
        super(HardSimpleVFE, self).__init__()
        self.num_features = num_features
        self.fp16_enabled = False

    def forward(self, features, num_points, coors=None):
        with torch.no_grad():
            return hard_simple_vfe(features.float(), num_points, self.num_features)


def voxelize_mean_batch(points, voxel_size, coors_range, max_points, max_voxels, points_range=None):
    """Hard voxelization + HardSimpleVFE(num_features=C) of a list of (N_i, C) CUDA frames.

    Returns ``(means, coors, num_points, voxel_num)``: means (F, max_voxels, C), coors (F,
    max_voxels, 3), num_points (F, max_voxels), voxel_num device int32 (F,); rows >= voxel_num[f]
    are unspecified.  For max_points == 5 and C in (4, 5) the encoder runs inside the expansion
    kernel and the (max_voxels, max_points, C) tensor is never materialised; other shapes run the
    batched voxelization followed by the stand-alone encoder kernel (same results).
    """
    assert len(points) > 0
    dev = points[0].device
    c = points[0].size(1)
    for p in points:
        voxel_layer._check_points(p)
        assert p.device == dev and p.size(1) == c
    nf = len(points)
    L = _cabi.lib()
    vs, rg = _cabi.f3(voxel_size), _cabi.f6(coors_range)
    with torch.no_grad():
        means = torch.empty((nf, max_voxels, c), dtype=torch.float32, device=dev)
        coors = torch.empty((nf, max_voxels, 3), dtype=torch.int32, device=dev)
        num = torch.empty((nf, max_voxels), dtype=torch.int32, device=dev)
        voxel_num = torch.empty((nf,), dtype=torch.int32, device=dev)
        n_max = max(p.size(0) for p in points)
        need = L.pcfe_hard_voxelize_workspace_bytes(n_max, nf, 0, vs, rg, max_points, max_voxels)
        ws = workspace(dev, need)
        flt = None if points_range is None else ctypes.cast(_cabi.f6(points_range), ctypes.POINTER(ctypes.c_float))
        fused = max_points == 5 and c in (4, 5) and all(p.data_ptr() % 16 == 0 for p in points)
        if fused:
            frames = (_cabi.Frame * nf)()
            for i, p in enumerate(points):
                frames[i] = _cabi.Frame(p.data_ptr(), p.size(0), means[i].data_ptr(), coors[i].data_ptr(),
                                        num[i].data_ptr())
            rc = L.pcfe_hard_voxelize_mean_batch_f32(frames, nf, c, vs, rg, flt, max_points, max_voxels,
                                                     ptr(voxel_num), ptr(ws), ws.numel(), dev.index, stream_ptr(dev))
            if rc == 0:
                return means, coors, num, voxel_num
            if rc != _cabi.ERR_SHAPE:  # ERR_SHAPE: a test knob selected a path without the epilogue
                _cabi.check(rc, "pcfe_hard_voxelize_mean_batch_f32")
        from ..voxel.voxelize import voxelize_batch
        voxels, coors, num, voxel_num = voxelize_batch(points, voxel_size, coors_range, max_points, max_voxels,
                                                       sync=False, points_range=points_range)
        for i in range(nf):
            means[i] = hard_simple_vfe(voxels[i], num[i], c, voxel_num=voxel_num[i:i + 1])
        return means, coors, num, voxel_num
