"""Packed batched hard voxelization: the detectors' voxelize() (openpcdet.py:59-76, voxelnet.py:50-67
-- per-frame voxel_layer, F.pad(coors, (1, 0), value=i), torch.cat) written directly by the kernels.
Checked bit for bit against the oracle's per-frame results concatenated the reference's way."""
import numpy as np
import pytest
import torch

from detmatch_b200 import _cabi, synth
from detmatch_b200.ops import voxelize_batch_packed
from oracle import oracle, vfe_mean
from tests.helpers import assert_same_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["record", "map0", "map1", "cluster", "fallback", "general", "global", "waves"])
def pack_mode(request):
    """record: packed rows from the expansion kernel; cluster: the same with the frame's partition +
    grouping done by one thread-block cluster (hv_cluster.cuh); fallback: every frame through the two-stage
    overflow fallback; general / global: paths without packed output (the wrapper concatenates);
    waves: two frames per launch sequence, so offsets cross waves."""
    mode = request.param
    _cabi.debug_set("hv_path", 1 if mode == "global" else 0)
    _cabi.debug_set("hv_force_overflow", 1 if mode == "fallback" else 0)
    _cabi.debug_set("hv_bucket_variant", 1 if mode == "general" else 0)
    _cabi.debug_set("hv_wave", 2 if mode == "waves" else 0)
    _cabi.debug_set("hv_cluster", 1 if mode == "cluster" else 0)
    _cabi.debug_set("hv_expand_map", {"map0": 0, "map1": 1}.get(mode, 2))  # record expansion: which tiles a warp takes
    yield mode
    _cabi.debug_set("hv_expand_map", 2)
    for k in ("hv_path", "hv_force_overflow", "hv_bucket_variant", "hv_wave"):
        _cabi.debug_set(k, 0)
    _cabi.debug_set("hv_cluster", 0)


def _reference_flow(frames, vs, rg, P, V, mean):
    """voxelize() of the detectors on the oracle's per-frame outputs."""
    vox, num, coors = [], [], []
    for i, p in enumerate(frames):
        ev, ec, en = oracle.hard_voxelize(p, vs, rg, P, V)
        vox.append(vfe_mean.hard_simple_vfe(ev, en) if mean else ev)
        num.append(en)
        coors.append(np.concatenate([np.full((len(en), 1), i, np.int32), ec], axis=1))
    return np.concatenate(vox), np.concatenate(num), np.concatenate(coors)


def _check(frames, vs, rg, P, V, mean, what, points_range=None, filtered=None):
    vox, num, coors = voxelize_batch_packed([torch.from_numpy(p).cuda() for p in frames], vs, rg, P, V, mean=mean,
                                            points_range=points_range)
    ev, en, ec = _reference_flow(filtered if filtered is not None else frames, vs, rg, P, V, mean)
    assert coors.shape == (len(en), 4) and num.shape == (len(en),)
    assert_same_bits(coors.cpu().numpy(), ec, what + " coors_batch")
    assert_same_bits(num.cpu().numpy(), en, what + " num_points")
    assert_same_bits(vox.cpu().numpy(), ev, what + (" means" if mean else " voxels"))


@pytest.mark.parametrize("mean", [False, True])
@pytest.mark.parametrize("cfg_name,ci,cap", [("C4", 4, None), ("C1", 1, None), ("C4", 4, 2500), ("C5", 5, 1500)])
def test_packed_ragged_batch_vs_oracle(cfg_name, ci, cap, mean, pack_mode):
    cfg = synth.CONFIGS[cfg_name]
    V = cap or cfg["max_voxels"]
    frames = [synth.lidar_frame(n, cfg["c"], 8000 + 10 * ci + k, cfg["r_max"]).numpy()
              for k, n in enumerate((21000, 0, 1, 12345, 33, 18000))]
    frames[3][11, 1] = np.nan
    _check(frames, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], V, mean, f"{cfg_name} {pack_mode}")


@pytest.mark.parametrize("mean", [False, True])
def test_packed_with_a_really_overflowing_frame(mean, pack_mode):
    """Frame 1 puts every point into one voxel (one bucket takes them all: the frame overflows for
    real and is voxelized by the fallback) between frames on the fast path."""
    cfg = synth.CONFIGS["C4"]
    vs, rg = cfg["voxel_size"], cfg["point_cloud_range"]
    a = synth.lidar_frame(20000, 5, 8101, cfg["r_max"]).numpy()
    b = np.tile(np.float32([[10.01, 5.02, 0.3, 0.5, 0.1]]), (20000, 1))
    b[:, 3] = np.arange(20000, dtype=np.float32)
    b[::7, 0] = 30.0  # a second voxel, interleaved
    c = synth.lidar_frame(15000, 5, 8102, cfg["r_max"]).numpy()
    _check([a, b, c, b[:5000], a[:100]], vs, rg, 5, 20000, mean, f"overflow {pack_mode}")


def test_packed_with_points_range_filter(pack_mode):
    cfg = synth.CONFIGS["C4"]
    vs, rg = cfg["voxel_size"], cfg["point_cloud_range"]
    fr = [rg[0] + 3.0, rg[1] + 1.5, rg[2] + 0.2, rg[3] - 7.0, rg[4] - 2.5, rg[5] - 0.4]
    lo, hi = np.asarray(fr[:3], np.float32), np.asarray(fr[3:], np.float32)
    frames = [synth.lidar_frame(16000 + k, 5, 8200 + k, cfg["r_max"]).numpy() for k in range(4)]
    kept = [np.ascontiguousarray(p[np.all(p[:, :3] > lo, axis=1) & np.all(p[:, :3] < hi, axis=1)]) for p in frames]
    _check(frames, vs, rg, 5, 9000, False, f"filter {pack_mode}", points_range=fr, filtered=kept)


def test_packed_entry_argument_checks():
    L = _cabi.lib()
    import ctypes
    pp = (ctypes.c_void_p * 1)(None)
    nn = (ctypes.c_int64 * 1)(10)
    vs, rg = _cabi.f3([0.1, 0.1, 0.1]), _cabi.f6([0, 0, 0, 1, 1, 1])

    def call(c, p, cap, coors=None):
        return L.pcfe_hard_voxelize_packed_batch_f32(pp, nn, 1, c, vs, rg, None, p, 100, 0, None, coors, None, cap, None, None,
                                                     0, 0, None)
    assert call(3, 5, 100) == _cabi.ERR_SHAPE      # c not in (4, 5)
    assert call(5, 64, 100) == _cabi.ERR_SHAPE     # max_points != 5
    assert call(5, 5, 9) == -4                     # PCFE_ERR_WORKSPACE: cap_rows < min(n, max_voxels)
    assert call(5, 5, 100, coors=8) == -5          # PCFE_ERR_ALIGN: coors_batch rows are 16-byte stores


def test_full_c4_batch_packed_equals_concatenation():
    """BASELINE C4 at full size (64 x 180 000 x 5): the packed tensors equal the concatenation of
    the per-frame outputs of the plain batched call, bit for bit; means likewise."""
    from detmatch_b200.ops import hard_simple_vfe, voxelize_batch
    cfg = synth.CONFIGS["C4"]
    vs, rg, P, V = cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"]
    pts = [synth.lidar_frame(cfg["n"], 5, synth.seed_for(4, k), cfg["r_max"]).cuda() for k in range(cfg["frames"])]
    vox_cat, num_cat, coors_batch = voxelize_batch(pts, vs, rg, P, V, sync=True)
    vox, num, coors = voxelize_batch_packed(pts, vs, rg, P, V)
    assert torch.equal(coors, coors_batch) and torch.equal(num, num_cat)
    assert torch.equal(vox.view(torch.int32), vox_cat.view(torch.int32))
    del vox
    means, num2, coors2 = voxelize_batch_packed(pts, vs, rg, P, V, mean=True)
    assert torch.equal(coors2, coors_batch) and torch.equal(num2, num_cat)
    assert torch.equal(means.view(torch.int32), hard_simple_vfe(vox_cat, num_cat).view(torch.int32))


def test_packed_all_frames_empty(pack_mode):
    """The reference's voxelize() loop over empty frames returns empty concatenated tensors."""
    empty = [torch.zeros((0, 5)).cuda(), torch.zeros((0, 5)).cuda()]
    for mean in (False, True):
        v, n, c = voxelize_batch_packed(empty, [0.1, 0.1, 0.15], [-75.2, -75.2, -2, 75.2, 75.2, 4], 5, 1000, mean=mean)
        assert v.shape == ((0, 5) if mean else (0, 5, 5)) and n.shape == (0,) and c.shape == (0, 4)


@pytest.mark.parametrize("cfg_name,ci,chunk", [("C4", 4, 2), ("C1", 1, 3), ("C5", 5, 2)])
def test_host_buffer_pipeline_vs_oracle(cfg_name, ci, chunk, pack_mode):
    """voxelize_batch_host: HOST frames in, the reference's concatenated tensors out in HOST memory
    (chunks pipelined on three streams, packed outputs read back at their final offsets), every frame
    against the oracle; C5 (P = 64) takes the per-frame outputs + concatenation branch."""
    from detmatch_b200.ops import HostVoxelizePipeline, voxelize_batch_host
    cfg = synth.CONFIGS[cfg_name]
    sizes = [cfg["n"] // 3, 0, cfg["n"] // 5, 4097, cfg["n"] // 4]
    frames = [synth.lidar_frame(n, cfg["c"], synth.seed_for(ci, 40 + k), cfg["r_max"]) for k, n in enumerate(sizes)]
    V = min(cfg["max_voxels"], 30000)
    P = cfg["max_num_points"]
    exp = [oracle.hard_voxelize(p.numpy(), cfg["voxel_size"], cfg["point_cloud_range"], P, V) for p in frames]
    ev = np.concatenate([e[0] for e in exp])
    en = np.concatenate([e[2] for e in exp])
    ec = np.concatenate([np.pad(e[1], ((0, 0), (1, 0)), constant_values=k) for k, e in enumerate(exp)]).astype(np.int32)
    v, n, c = voxelize_batch_host(frames, cfg["voxel_size"], cfg["point_cloud_range"], P, V, chunk=chunk)
    assert not v.is_cuda and not n.is_cuda and not c.is_cuda
    assert_same_bits(v.numpy(), ev, "host voxels")
    assert_same_bits(n.numpy(), en, "host num")
    assert_same_bits(c.numpy(), ec, "host coors_batch")
    # the reusable pipeline: pinned inputs, two runs give the same views
    pipe = HostVoxelizePipeline(sizes, cfg["c"], cfg["voxel_size"], cfg["point_cloud_range"], P, V, chunk=chunk)
    pinned = [p.pin_memory() for p in frames]
    for _ in range(2):
        v2, n2, c2 = pipe.run(pinned)
        assert_same_bits(v2.numpy(), ev, "pipeline voxels")
        assert_same_bits(c2.numpy(), ec, "pipeline coors_batch")
    assert pipe.counts == [len(e[2]) for e in exp]
    assert pipe.d2h_bytes == sum(len(e[2]) for e in exp) * (P * cfg["c"] * 4 + 20) + 4 * len(sizes)
