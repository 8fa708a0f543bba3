"""Drop-in for the reference's pybind module
``mmdet3d.ops.roiaware_pool3d.roiaware_pool3d_ext`` (roiaware_pool3d.cpp:126-136): same names,
**boxes first** argument order, caller-allocated ``out`` written in place, returns 1.

Differences, all deliberate: launched on the tensors' device and the current stream without
touching the global current device (the reference needs torch.cuda.set_device,
points_in_boxes.py:32-44); never exit()s (points_in_boxes_cuda.cu:120-125); arithmetic is the
CPU op's, bit for bit.
"""
import torch

from ... import _cabi
from ..._torch_glue import ptr, stream_ptr, workspace


def _check(boxes, points, out):
    for t in (boxes, points, out):
        if not t.is_cuda:
            raise RuntimeError("roiaware_pool3d_ext: tensors must be CUDA tensors (no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError("roiaware_pool3d_ext: tensors must be contiguous")
    if boxes.dtype != torch.float32 or points.dtype != torch.float32 or out.dtype != torch.int32:
        raise TypeError("roiaware_pool3d_ext: boxes/points must be float32 and out int32")
    if not (boxes.device == points.device == out.device):
        raise RuntimeError("roiaware_pool3d_ext: tensors must be on the same device")


def _run(fn_name, boxes, points, out, b, t, m):
    dev = points.device
    L = _cabi.lib()
    ws = workspace(dev, L.pcfe_points_in_boxes_workspace_bytes(b, t))
    fn = getattr(L, fn_name)
    if fn_name == "pcfe_points_in_boxes_boxmajor_f32":
        rc = fn(ptr(boxes), ptr(points), t, m, ptr(out), ptr(ws), ws.numel(), dev.index, stream_ptr(dev))
    else:
        rc = fn(ptr(boxes), ptr(points), b, t, m, ptr(out), ptr(ws), ws.numel(), dev.index, stream_ptr(dev))
    _cabi.check(rc, fn_name)
    return 1


def points_in_boxes_gpu(boxes, points, out):
    """(B,T,7), (B,M,3) -> out (B,M) int32: lowest containing box index or -1."""
    _check(boxes, points, out)
    b, t = boxes.shape[0], boxes.shape[1]
    m = points.shape[1]
    assert boxes.shape[2] == 7 and points.shape[2] == 3 and points.shape[0] == b and out.shape == (b, m)
    return _run("pcfe_points_in_boxes_part_f32", boxes, points, out, b, t, m)


def points_in_boxes_batch(boxes, points, out):
    """(B,T,7), (B,M,3) -> out (B,M,T) int32 0/1."""
    _check(boxes, points, out)
    b, t = boxes.shape[0], boxes.shape[1]
    m = points.shape[1]
    assert boxes.shape[2] == 7 and points.shape[2] == 3 and points.shape[0] == b and out.shape == (b, m, t)
    return _run("pcfe_points_in_boxes_all_f32", boxes, points, out, b, t, m)


def points_in_boxes_cpu(boxes, points, out):
    """(T,7), (N,3) -> out (T,N) int32 0/1 (the CPU op's layout, computed on the device)."""
    _check(boxes, points, out)
    t, n = boxes.shape[0], points.shape[0]
    assert boxes.shape[1] == 7 and points.shape[1] == 3 and out.shape == (t, n)
    return _run("pcfe_points_in_boxes_boxmajor_f32", boxes, points, out, 1, t, n)


def forward(rois, pts, pts_feature, argmax, pts_idx_of_voxels, pooled_features, pool_method):
    """roiaware_pool3d.cpp:49-91 (roiaware_pool3d_gpu): rois (N,7), pts (npoints,3), pts_feature
    (npoints,C), argmax (N,ox,oy,oz,C) int32, pts_idx_of_voxels (N,ox,oy,oz,max_pts) int32,
    pooled_features (N,ox,oy,oz,C); pool_method 0 = max, 1 = avg.  Outputs need no pre-zeroing."""
    for t in (rois, pts, pts_feature, argmax, pts_idx_of_voxels, pooled_features):
        if not t.is_cuda:
            raise RuntimeError("roiaware_pool3d_ext.forward: tensors must be CUDA tensors (no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError("roiaware_pool3d_ext.forward: tensors must be contiguous")
    if not (rois.dtype == pts.dtype == pts_feature.dtype == pooled_features.dtype == torch.float32):
        raise TypeError("roiaware_pool3d_ext.forward: float32 only")
    if argmax.dtype != torch.int32 or pts_idx_of_voxels.dtype != torch.int32:
        raise TypeError("roiaware_pool3d_ext.forward: argmax / pts_idx_of_voxels must be int32")
    n, m, c = rois.size(0), pts.size(0), pts_feature.size(1)
    ox, oy, oz, mp = (pts_idx_of_voxels.size(k) for k in (1, 2, 3, 4))
    assert (ox < 256) and (oy < 256) and (oz < 256)  # roiaware_pool3d.cpp:72-73
    assert rois.size(1) == 7 and pts.size(1) == 3 and pts_feature.size(0) == m
    dev = pts.device
    rc = _cabi.lib().pcfe_roiaware_pool3d_forward_f32(ptr(rois), ptr(pts), ptr(pts_feature), n, m, c, mp, ox, oy, oz,
                                                      int(pool_method), ptr(argmax), ptr(pts_idx_of_voxels),
                                                      ptr(pooled_features), dev.index, stream_ptr(dev))
    _cabi.check(rc, "pcfe_roiaware_pool3d_forward_f32")
    return 1


def backward(pts_idx_of_voxels, argmax, grad_out, grad_in, pool_method):
    """roiaware_pool3d.cpp:93-123 (roiaware_pool3d_gpu_backward): grad_in (npoints, C) is overwritten."""
    for t in (pts_idx_of_voxels, argmax, grad_out, grad_in):
        if not t.is_cuda or not t.is_contiguous():
            raise RuntimeError("roiaware_pool3d_ext.backward: tensors must be contiguous CUDA tensors")
    n = pts_idx_of_voxels.size(0)
    ox, oy, oz, mp = (pts_idx_of_voxels.size(k) for k in (1, 2, 3, 4))
    c = grad_out.size(4)
    dev = grad_out.device
    rc = _cabi.lib().pcfe_roiaware_pool3d_backward_f32(ptr(pts_idx_of_voxels), ptr(argmax), ptr(grad_out), n, ox, oy, oz, c, mp,
                                                       int(pool_method), grad_in.size(0), ptr(grad_in), dev.index,
                                                       stream_ptr(dev))
    _cabi.check(rc, "pcfe_roiaware_pool3d_backward_f32")
    return 1
