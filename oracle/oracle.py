"""oracle/oracle.py -- numpy front end of the plain-C oracle (liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs, never by the product package.  Every function cites the reference lines the
C code restates (see pcfe_oracle.c); parity of the restatement itself is pinned in
tests/test_oracle.py against the reference's known-answer tests and against oracle/_ref.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u32p = ctypes.POINTER(ctypes.c_uint32)


def build():
    """Compile liboracle.so with gcc if missing or stale."""
    src = os.path.join(_HERE, "pcfe_oracle.c")
    hdr = os.path.join(_HERE, "pcfe_oracle.h")
    if (not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.pcfe_oracle_grid_size.argtypes = [_f32p, _f32p, _i32p]
        L.pcfe_oracle_grid_size.restype = None
        L.pcfe_oracle_dynamic_voxelize_f32.argtypes = [_f32p, ctypes.c_int64, ctypes.c_int, _f32p, _f32p, _i32p]
        L.pcfe_oracle_dynamic_voxelize_f64.argtypes = [_f64p, ctypes.c_int64, ctypes.c_int, _f32p, _f32p, _i32p]
        L.pcfe_oracle_hard_voxelize_f32.argtypes = [_f32p, ctypes.c_int64, ctypes.c_int, _f32p, _f32p,
                                                    ctypes.c_int, ctypes.c_int, _f32p, _i32p, _i32p]
        L.pcfe_oracle_hard_voxelize_f64.argtypes = [_f64p, ctypes.c_int64, ctypes.c_int, _f32p, _f32p,
                                                    ctypes.c_int, ctypes.c_int, _f64p, _i32p, _i32p]
        L.pcfe_oracle_roiaware_pool3d_forward.argtypes = [_f32p, ctypes.c_int, _f32p, ctypes.c_int64, _f32p, ctypes.c_int,
                                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                          _i32p, _i32p, _f32p]
        L.pcfe_oracle_roiaware_pool3d_backward.argtypes = [_i32p, _i32p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                           ctypes.c_int64, _f32p]
        L.pcfe_oracle_points_in_boxes_cpu.argtypes = [_f32p, ctypes.c_int, _f32p, ctypes.c_int64, _i32p]
        L.pcfe_oracle_points_in_boxes_restated.argtypes = [_f32p, ctypes.c_int, _f32p, ctypes.c_int64, _i32p]
        L.pcfe_oracle_pcdet_points_in_boxes.argtypes = [_f32p, ctypes.c_int, _f32p, ctypes.c_int64, ctypes.c_float, _i32p]
        L.pcfe_oracle_sincosf.argtypes = [ctypes.c_float, _f32p, _f32p]
        L.pcfe_oracle_sincosf.restype = None
        L.pcfe_oracle_host_sincosf.argtypes = [ctypes.c_float, _f32p, _f32p]
        L.pcfe_oracle_host_sincosf.restype = None
        L.pcfe_oracle_sincosf_sweep.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, _u32p]
        L.pcfe_oracle_sincosf_sweep.restype = ctypes.c_int64
        for name in ("pcfe_oracle_host_sincosf_array", "pcfe_oracle_sincosf_array"):
            getattr(L, name).argtypes = [_f32p, ctypes.c_int64, _f32p, _f32p]
            getattr(L, name).restype = None
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def _vs_range(voxel_size, coors_range):
    # pybind narrows python floats to std::vector<float> (voxelization.h:51-83)
    vs = np.asarray(voxel_size, dtype=np.float32).copy()
    rg = np.asarray(coors_range, dtype=np.float32).copy()
    assert vs.shape == (3,) and rg.shape == (6,)
    return vs, rg


def grid_size(voxel_size, coors_range):
    vs, rg = _vs_range(voxel_size, coors_range)
    g = np.zeros(3, dtype=np.int32)
    lib().pcfe_oracle_grid_size(_p(vs, _f32p), _p(rg, _f32p), _p(g, _i32p))
    return g


def dynamic_voxelize(points, voxel_size, coors_range):
    """voxelization_cpu.cpp:7-41,144-169.  points (N,C) float32|float64 -> coors (N,3) int32."""
    vs, rg = _vs_range(voxel_size, coors_range)
    pts = np.ascontiguousarray(points)
    n, c = pts.shape
    coors = np.zeros((n, 3), dtype=np.int32)
    if pts.dtype == np.float64:
        rc = lib().pcfe_oracle_dynamic_voxelize_f64(_p(pts, _f64p), n, c, _p(vs, _f32p), _p(rg, _f32p), _p(coors, _i32p))
    else:
        pts = pts.astype(np.float32, copy=False)
        rc = lib().pcfe_oracle_dynamic_voxelize_f32(_p(pts, _f32p), n, c, _p(vs, _f32p), _p(rg, _f32p), _p(coors, _i32p))
    assert rc == 0, rc
    return coors


def hard_voxelize(points, voxel_size, coors_range, max_points, max_voxels):
    """voxelization_cpu.cpp:43-99,105-142 called the way voxelize.py:46-58 does.

    Returns (voxels[:M], coors[:M], num_points_per_voxel[:M]).
    """
    vs, rg = _vs_range(voxel_size, coors_range)
    pts = np.ascontiguousarray(points)
    if pts.dtype != np.float64:
        pts = pts.astype(np.float32, copy=False)
    n, c = pts.shape
    voxels = np.zeros((max_voxels, max_points, c), dtype=pts.dtype)
    coors = np.zeros((max_voxels, 3), dtype=np.int32)
    num = np.zeros((max_voxels,), dtype=np.int32)
    if pts.dtype == np.float64:
        m = lib().pcfe_oracle_hard_voxelize_f64(_p(pts, _f64p), n, c, _p(vs, _f32p), _p(rg, _f32p),
                                                max_points, max_voxels, _p(voxels, _f64p), _p(coors, _i32p), _p(num, _i32p))
    else:
        m = lib().pcfe_oracle_hard_voxelize_f32(_p(pts, _f32p), n, c, _p(vs, _f32p), _p(rg, _f32p),
                                                max_points, max_voxels, _p(voxels, _f32p), _p(coors, _i32p), _p(num, _i32p))
    assert m >= 0, m
    return voxels[:m], coors[:m], num[:m]


def points_in_boxes_cpu(points, boxes, restated_trig=False):
    """points_in_boxes_cpu.cpp:42-69 via points_in_boxes.py:53-82: (N,3),(T,7) -> (T,N) int32."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    bx = np.ascontiguousarray(boxes, dtype=np.float32)
    assert bx.ndim == 2 and bx.shape[1] == 7 and pts.ndim == 2 and pts.shape[1] == 3
    out = np.zeros((bx.shape[0], pts.shape[0]), dtype=np.int32)
    fn = lib().pcfe_oracle_points_in_boxes_restated if restated_trig else lib().pcfe_oracle_points_in_boxes_cpu
    rc = fn(_p(bx, _f32p), bx.shape[0], _p(pts, _f32p), pts.shape[0], _p(out, _i32p))
    assert rc == 1, rc
    return out


def pcdet_points_in_boxes_cpu(points, boxes, margin=1e-2):
    """OpenPCDet points_in_boxes_cpu (roiaware_pool3d.cpp:121-168 via roiaware_pool3d_utils.py:9-25):
    (N,3),(T,7) -> (T,N) int32.  margin=1e-5 gives the arithmetic of the CUDA twin."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    bx = np.ascontiguousarray(boxes, dtype=np.float32)
    assert bx.ndim == 2 and bx.shape[1] == 7 and pts.ndim == 2 and pts.shape[1] == 3
    out = np.zeros((bx.shape[0], pts.shape[0]), dtype=np.int32)
    rc = lib().pcfe_oracle_pcdet_points_in_boxes(_p(bx, _f32p), bx.shape[0], _p(pts, _f32p), pts.shape[0],
                                                 ctypes.c_float(margin), _p(out, _i32p))
    assert rc == 1, rc
    return out


def pcdet_points_in_boxes_gpu(points, boxes):
    """(B,M,3),(B,T,7) -> (B,M) lowest containing box or -1 (roiaware_pool3d_kernel.cu:313-336,
    MARGIN 1e-5, with the CPU arithmetic)."""
    out = []
    for p, b in zip(points, boxes):
        m = pcdet_points_in_boxes_cpu(p, b, margin=1e-5)  # (T,N)
        hit = m.argmax(axis=0).astype(np.int32) if m.shape[0] else np.zeros(m.shape[1], np.int32)
        any_ = m.any(axis=0) if m.shape[0] else np.zeros(m.shape[1], bool)
        out.append(np.where(any_, hit, -1).astype(np.int32))
    return np.stack(out)


def points_in_boxes_batch(points, boxes):
    """(B,M,3),(B,T,7) -> (B,M,T) 0/1: points_in_boxes_cpu per frame, transposed
    (semantics of points_in_boxes_cuda.cu:79-105 with the CPU arithmetic)."""
    return np.stack([points_in_boxes_cpu(p, b).T.copy() for p, b in zip(points, boxes)])


def points_in_boxes_gpu(points, boxes):
    """(B,M,3),(B,T,7) -> (B,M): lowest box index containing the point or -1
    (semantics of points_in_boxes_cuda.cu:51-77 with the CPU arithmetic)."""
    out = []
    for p, b in zip(points, boxes):
        m = points_in_boxes_cpu(p, b)  # (T,N)
        if m.shape[0] == 0:
            out.append(np.full((m.shape[1],), -1, dtype=np.int32))
            continue
        first = np.argmax(m, axis=0).astype(np.int32)
        first[m.max(axis=0) == 0] = -1
        out.append(first)
    return np.stack(out)


def sincosf(x, host=False):
    s = ctypes.c_float()
    c = ctypes.c_float()
    fn = lib().pcfe_oracle_host_sincosf if host else lib().pcfe_oracle_sincosf
    fn(ctypes.c_float(x), ctypes.byref(s), ctypes.byref(c))
    return np.float32(s.value), np.float32(c.value)


def sincosf_array(x, host=True):
    """sinf/cosf of every element of x with the host libm (host=True) or the restatement."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    s, c = np.empty_like(x), np.empty_like(x)
    fn = lib().pcfe_oracle_host_sincosf_array if host else lib().pcfe_oracle_sincosf_array
    fn(_p(x, _f32p), x.size, _p(s, _f32p), _p(c, _f32p))
    return s, c


def sincosf_sweep(lo_bits, hi_bits, stride=1):
    """#floats in [lo_bits,hi_bits) where restated sinf/cosf != host sinf/cosf/sincosf."""
    bad = ctypes.c_uint32(0)
    n = lib().pcfe_oracle_sincosf_sweep(lo_bits, hi_bits, stride, ctypes.byref(bad))
    return int(n), int(bad.value)


def roiaware_pool3d_forward(rois, pts, pts_feature, out_size, max_pts_per_voxel, mode):
    """RoIAwarePool3dFunction.forward (roiaware_pool3d.py:46-88) on numpy arrays: returns
    (pooled_features, argmax, pts_idx_of_voxels).  mode 0 = max, 1 = avg."""
    ox, oy, oz = (out_size,) * 3 if isinstance(out_size, int) else out_size
    r = np.ascontiguousarray(rois, dtype=np.float32)
    p = np.ascontiguousarray(pts, dtype=np.float32)
    f = np.ascontiguousarray(pts_feature, dtype=np.float32)
    n, m, c = r.shape[0], p.shape[0], f.shape[1]
    pooled = np.zeros((n, ox, oy, oz, c), dtype=np.float32)
    argmax = np.zeros((n, ox, oy, oz, c), dtype=np.int32)
    lists = np.zeros((n, ox, oy, oz, max_pts_per_voxel), dtype=np.int32)
    rc = lib().pcfe_oracle_roiaware_pool3d_forward(_p(r, _f32p), n, _p(p, _f32p), m, _p(f, _f32p), c, max_pts_per_voxel,
                                                   ox, oy, oz, int(mode), _p(argmax, _i32p), _p(lists, _i32p), _p(pooled, _f32p))
    assert rc == 1, rc
    return pooled, argmax, lists


def roiaware_pool3d_backward(pts_idx_of_voxels, argmax, grad_out, num_pts, mode):
    """RoIAwarePool3dFunction.backward (roiaware_pool3d.py:90-106): grad_in (num_pts, C)."""
    lists = np.ascontiguousarray(pts_idx_of_voxels, dtype=np.int32)
    am = np.ascontiguousarray(argmax, dtype=np.int32)
    g = np.ascontiguousarray(grad_out, dtype=np.float32)
    n, ox, oy, oz, mp = lists.shape
    c = g.shape[4]
    grad_in = np.zeros((num_pts, c), dtype=np.float32)
    rc = lib().pcfe_oracle_roiaware_pool3d_backward(_p(lists, _i32p), _p(am, _i32p), _p(g, _f32p), n, ox, oy, oz, c, mp,
                                                    int(mode), num_pts, _p(grad_in, _f32p))
    assert rc == 1, rc
    return grad_in
