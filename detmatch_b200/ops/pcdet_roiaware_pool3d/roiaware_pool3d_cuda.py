"""Drop-in for the point-in-box entries of OpenPCDet's pybind module ``roiaware_pool3d_cuda``
(thirdparty/Spconv-OpenPCDet/pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:98-118,143-168,175-176):
same names, **boxes first**, caller-allocated ``out`` written in place, returns 1.

Boxes are OpenPCDet's (x, y, z_centre, dx, dy, dz, heading).  The arithmetic is the CPU function's
(roiaware_pool3d.cpp:121-140: float32, glibc cosf/sinf of -heading, no contraction, double
right-hand sides) with each entry's own MARGIN -- 1e-2 for ``points_in_boxes_cpu`` (:131), 1e-5 for
``points_in_boxes_gpu`` (roiaware_pool3d_kernel.cu:27).  Every element of ``out`` is written (the
reference's GPU kernel writes hits only and relies on the caller's fill_(-1)).
"""
from ... import _cabi
from ..._torch_glue import ptr, stream_ptr, workspace
from ..roiaware_pool3d.roiaware_pool3d_ext import _check


def points_in_boxes_gpu(boxes, points, out):
    """(B,T,7), (B,M,3) -> out (B,M) int32: lowest containing box index or -1."""
    _check(boxes, points, out)
    b, t = boxes.shape[0], boxes.shape[1]
    m = points.shape[1]
    assert boxes.shape[2] == 7 and points.shape[2] == 3 and points.shape[0] == b and out.shape == (b, m)
    dev = points.device
    L = _cabi.lib()
    ws = workspace(dev, L.pcfe_points_in_boxes_workspace_bytes(b, t))
    _cabi.check(L.pcfe_pcdet_points_in_boxes_gpu_f32(ptr(boxes), ptr(points), b, t, m, ptr(out), ptr(ws), ws.numel(),
                                                     dev.index, stream_ptr(dev)), "pcfe_pcdet_points_in_boxes_gpu_f32")
    return 1


def points_in_boxes_cpu(boxes, points, out):
    """(T,7), (N,3) -> out (T,N) int32 0/1 (the CPU op's layout and MARGIN, computed on the device)."""
    _check(boxes, points, out)
    t, n = boxes.shape[0], points.shape[0]
    assert boxes.shape[1] == 7 and points.shape[1] == 3 and out.shape == (t, n)
    dev = points.device
    L = _cabi.lib()
    ws = workspace(dev, L.pcfe_points_in_boxes_workspace_bytes(1, t))
    _cabi.check(L.pcfe_pcdet_points_in_boxes_cpu_f32(ptr(boxes), ptr(points), t, n, ptr(out), ptr(ws), ws.numel(),
                                                     dev.index, stream_ptr(dev)), "pcfe_pcdet_points_in_boxes_cpu_f32")
    return 1
