"""CPU tests: the C-ABI library builds/loads and exports every symbol include/pcfe.h declares;
host-side logic that needs no GPU (grid size, workspace sizing, argument errors)."""
import ctypes
import os
import re

import pytest

from detmatch_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _cabi.lib()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "pcfe.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pcfe_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 14
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"libpcfe.so does not export {name}"
    assert declared == set(_cabi.PROTOTYPES), "ctypes prototypes out of sync with include/pcfe.h"


def test_version_and_error_strings(lib):
    assert lib.pcfe_version() == 100
    assert lib.pcfe_error_string(0) == b"ok"
    assert b"workspace" in lib.pcfe_error_string(-4)
    assert lib.pcfe_error_string(1)  # a cudaError_t string


def test_grid_size_matches_reference_float32_rounding(lib):
    assert _cabi.grid_size([0.1, 0.1, 0.15], [-75.2, -75.2, -2, 75.2, 75.2, 4]) == [1504, 1504, 40]
    assert _cabi.grid_size([0.05, 0.05, 0.1], [0, -40, -3, 70.4, 40, 1]) == [1408, 1600, 40]
    assert _cabi.grid_size([0.25, 0.25, 8], [-50, -50, -5, 50, 50, 3]) == [400, 400, 1]
    assert _cabi.grid_size([0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1]) == [141, 160, 8]
    from oracle import oracle
    for vs, rg in (([0.1, 0.1, 0.15], [-75.2, -75.2, -2, 75.2, 75.2, 4]), ([0.3, 0.7, 0.11], [-1, -2, -3, 4.4, 5.5, 6.6])):
        assert _cabi.grid_size(vs, rg) == oracle.grid_size(vs, rg).tolist()


def test_workspace_sizing(lib):
    vs, rg = _cabi.f3([0.1, 0.1, 0.15]), _cabi.f6([-75.2, -75.2, -2, 75.2, 75.2, 4])
    one = lib.pcfe_hard_voxelize_workspace_bytes(180000, 1, 1, vs, rg, 5, 150000)
    assert one > 0 and one % 256 == 0
    # more frames than one wave: two wave buffers so that consecutive waves can overlap
    assert lib.pcfe_hard_voxelize_workspace_bytes(180000, 64, 4, vs, rg, 5, 150000) == 2 * 4 * one
    assert lib.pcfe_hard_voxelize_workspace_bytes(180000, 4, 4, vs, rg, 5, 150000) == 4 * one
    auto = lib.pcfe_hard_voxelize_workspace_bytes(180000, 64, 0, vs, rg, 5, 150000)
    assert auto % one == 0 and one <= auto <= 2 * 64 * one
    # empty grid -> 0 (error)
    assert lib.pcfe_hard_voxelize_workspace_bytes(10, 1, 1, vs, _cabi.f6([0, 0, 0, 0, 0, 0]), 5, 10) == 0
    # prepared boxes (32 B) + conservative-reject records (16 B) per box
    # prepared boxes + reject records + (first-hit grid) headers, cell starts, cell lists of 64 entries per box
    al = lambda x: (x + 255) // 256 * 256  # noqa: E731
    assert lib.pcfe_points_in_boxes_workspace_bytes(16, 200) == (
        16 * 200 * (32 + 16) + al(16 * 32) + al(16 * 4097 * 4) + al(16 * 200 * 64 * 2))


def test_argument_errors_without_gpu(lib):
    """Argument validation happens before any CUDA call, so it is testable on CPU."""
    vs, rg = _cabi.f3([0.5] * 3), _cabi.f6([0, -40, -3, 70.4, 40, 1])
    null = ctypes.c_void_p(0)
    assert lib.pcfe_dynamic_voxelize_f32(null, -1, 4, vs, rg, null, 0, null) == -2   # negative n
    assert lib.pcfe_dynamic_voxelize_f32(null, 10, 2, vs, rg, null, 0, null) == -2   # c < 3
    assert lib.pcfe_dynamic_voxelize_f32(null, 10, 4, vs, rg, null, 0, null) == -1   # NULL points
    assert lib.pcfe_hard_voxelize_f32(null, 10, 4, vs, rg, -1, 10, null, null, null, null, null, 0, 0, null) == -1
    fake = ctypes.c_void_p(0x1000)
    assert lib.pcfe_hard_voxelize_f32(fake, 10, 4, vs, rg, -1, 10, fake, fake, fake, fake, fake, 1 << 20, 0, null) == -6
    assert lib.pcfe_hard_voxelize_f32(fake, 10, 4, vs, _cabi.f6([0] * 6), 5, 10, fake, fake, fake, fake, fake, 1 << 20, 0, null) == -3
    assert lib.pcfe_points_in_boxes_all_f32(fake, fake, -1, 1, 1, fake, fake, 1 << 20, 0, null) == -2
    assert lib.pcfe_points_in_boxes_all_f32(fake, fake, 1, 4, 8, fake, fake, 16, 0, null) == -4   # workspace too small


def test_python_api_mirrors_reference_names():
    import detmatch_b200.ops as ops
    from detmatch_b200.ops.roiaware_pool3d import roiaware_pool3d_ext
    from detmatch_b200.ops.voxel import voxel_layer
    for name in ("Voxelization", "voxelization", "points_in_boxes_batch", "points_in_boxes_cpu", "points_in_boxes_gpu"):
        assert hasattr(ops, name)
    for name in ("hard_voxelize", "dynamic_voxelize"):
        assert hasattr(voxel_layer, name)
    for name in ("points_in_boxes_gpu", "points_in_boxes_batch", "points_in_boxes_cpu"):
        assert hasattr(roiaware_pool3d_ext, name)
    v = ops.Voxelization([0.05, 0.05, 0.1], [0, -40, -3, 70.4, 40, 1], 5, (16000, 40000))
    assert v.grid_size.tolist() == [1408, 1600, 40] and [int(x) for x in v.pcd_shape] == [1, 1600, 1408]
    assert repr(v) == ("Voxelization(voxel_size=[0.05, 0.05, 0.1], point_cloud_range=[0, -40, -3, 70.4, 40, 1], "
                       "max_num_points=5, max_voxels=(16000, 40000))")
    assert ops.Voxelization([0.5] * 3, [0, -40, -3, 70.4, 40, 1], 5, 123).max_voxels == (123, 123)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    import detmatch_b200.ops as ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.voxelization(torch.zeros(4, 4), [0.5] * 3, [0, -40, -3, 70.4, 40, 1], 5, 10)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.points_in_boxes_cpu(torch.zeros(4, 3), torch.zeros(1, 7))


def test_next_row_entries_argument_errors_without_gpu(lib):
    """Packed / mean voxelization, the encoder, DynamicScatter and the OpenPCDet point-in-box entries
    validate their arguments before any CUDA call."""
    vs, rg = _cabi.f3([0.5] * 3), _cabi.f6([0, -40, -3, 70.4, 40, 1])
    null, fake = ctypes.c_void_p(0), ctypes.c_void_p(0x1000)
    fr = (_cabi.Frame * 1)()
    assert lib.pcfe_hard_voxelize_mean_batch_f32(fr, 1, 5, vs, rg, None, 64, 10, null, null, 0, 0, null) == -2  # P != 5
    assert lib.pcfe_hard_voxelize_mean_batch_f32(fr, 1, 3, vs, rg, None, 5, 10, null, null, 0, 0, null) == -2   # C not 4 / 5
    pp, nn = (ctypes.c_void_p * 1)(0x1000), (ctypes.c_int64 * 1)(100)
    args = (vs, rg, None, 5, 50, 0)
    assert lib.pcfe_hard_voxelize_packed_batch_f32(pp, nn, 1, 5, *args, fake, fake, fake, 49, fake, fake, 1 << 20, 0, null) == -4
    assert lib.pcfe_hard_voxelize_packed_batch_f32(pp, nn, 1, 5, *args, fake, ctypes.c_void_p(0x1008), fake, 50, fake, fake,
                                                   1 << 20, 0, null) == -5
    assert lib.pcfe_voxel_mean_f32(null, null, null, -1, 5, 4, null, 0, null) == -2
    assert lib.pcfe_voxel_mean_f32(null, null, null, 0, 5, 4, null, 0, null) == 0      # nothing to do
    assert lib.pcfe_voxel_mean_f32(null, fake, null, 3, 5, 4, fake, 0, null) == -1
    dims = (ctypes.c_int32 * 3)(40, 1600, 1408)
    need = lib.pcfe_dynamic_scatter_workspace_bytes(dims, 3, 120000)
    # level 1: one bit per 256 cells; level 2: 32 bytes (+ 32 of prefix) per point at most
    assert 2 * 120000 * 32 <= need <= 2 * 120000 * 32 + 4 * (40 * 1600 * 1408 // 256 // 8) + (1 << 16)
    assert lib.pcfe_dynamic_scatter_workspace_bytes((ctypes.c_int32 * 3)(1 << 20, 1 << 20, 1 << 20), 3, 10) == 0  # too large
    assert lib.pcfe_dynamic_scatter_map_i32(fake, 10, 5, dims, fake, fake, fake, 1 << 30, 0, null) == -2  # ndim > 4
    assert lib.pcfe_dynamic_scatter_map_i32(fake, 10, 3, dims, fake, null, fake, 1 << 30, 0, null) == -1
    assert lib.pcfe_dynamic_scatter_reduce_f32(fake, fake, fake, 10, 4, 3, 7, 5, fake, fake, fake, 0, null) == -2  # reduce
    assert lib.pcfe_dynamic_scatter_backward_f32(fake, fake, fake, fake, fake, 10, 5, 0, 2, fake, fake, 0, 0, null) == -2
    assert lib.pcfe_pcdet_points_in_boxes_gpu_f32(fake, fake, -1, 1, 1, fake, fake, 1 << 20, 0, null) == -2
    assert lib.pcfe_pcdet_points_in_boxes_cpu_f32(fake, fake, 4, 8, fake, fake, 16, 0, null) == -4


def test_next_row_python_mirrors():
    import torch
    import detmatch_b200.ops as ops
    from detmatch_b200.ops.pcdet_roiaware_pool3d import roiaware_pool3d_cuda, roiaware_pool3d_utils
    for name in ("DynamicScatter", "dynamic_scatter", "HardSimpleVFE", "voxelize_batch_packed", "voxelize_mean_batch"):
        assert hasattr(ops, name)
    for mod in (roiaware_pool3d_cuda, roiaware_pool3d_utils):
        assert hasattr(mod, "points_in_boxes_gpu") and hasattr(mod, "points_in_boxes_cpu")
    assert repr(ops.DynamicScatter([0.32, 0.32, 6], [-74.88, -74.88, -2, 74.88, 74.88, 4], True)) == (
        "DynamicScatter(voxel_size=[0.32, 0.32, 6], point_cloud_range=[-74.88, -74.88, -2, 74.88, 74.88, 4], "
        "average_points=True)")
    assert ops.HardSimpleVFE(num_features=5).num_features == 5
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU"):
            ops.hard_simple_vfe(torch.zeros(2, 5, 4), torch.ones(2, dtype=torch.int32))
        with pytest.raises(RuntimeError, match="no CPU"):
            ops.dynamic_scatter(torch.zeros(2, 4), torch.zeros(2, 3, dtype=torch.int32), "max")
