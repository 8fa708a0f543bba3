"""Mean voxel encoder (HardSimpleVFE, voxel_encoder.py:12-44) on the GPU: stand-alone kernel and
the epilogue fused into hard voxelization, bit for bit against oracle/vfe_mean.py (slot-order
float32 sum, IEEE divide) and within the summation error bound of the reference expression's own
output (tests/golden/vfe_mean.npz)."""
import numpy as np
import pytest
import torch

from detmatch_b200 import _cabi, synth
from detmatch_b200.ops.voxel_encoders import HardSimpleVFE, hard_simple_vfe, voxelize_mean_batch
from oracle import oracle, vfe_mean
from tests.helpers import assert_same_bits, golden

pytestmark = pytest.mark.gpu

EPS = np.float32(2.0 ** -24)


def _within_sum_bound(got, expected, features, num_points, nf):
    """|got - expected| <= 2 (P - 1) eps sum_s |x_s| / n + 2 eps |expected|: two float32 summations
    of the same P terms in different association orders, each followed by one rounded divide."""
    p = features.shape[1]
    mag = np.abs(features[:, :, :nf]).sum(axis=1, dtype=np.float64) / num_points.reshape(-1, 1)
    bound = 2 * (p - 1) * float(EPS) * mag + 2 * float(EPS) * np.abs(expected.astype(np.float64))
    err = np.abs(got.astype(np.float64) - expected.astype(np.float64))
    assert np.all(err <= bound), f"max excess {np.max(err - bound)}"


@pytest.mark.parametrize("tag,nf", [("c4", 5), ("c1", 4), ("c4_nf4", 4), ("rand", 5)])
def test_standalone_vs_oracle_and_reference_golden(tag, nf):
    g = golden("vfe_mean")
    src = "c4" if tag == "c4_nf4" else tag
    f, n, exp = g[src + "_features"], g[src + "_num_points"], g[tag + "_expected"]
    got = HardSimpleVFE(num_features=nf)(torch.from_numpy(f).cuda(), torch.from_numpy(n).cuda(), None).cpu().numpy()
    assert_same_bits(got, vfe_mean.hard_simple_vfe(f, n, nf), f"{tag} vs oracle")
    _within_sum_bound(got, exp, f, n, nf)


def test_reference_test_shape():
    """tests/test_models/test_voxel_encoder/test_voxel_encoders.py:26-33 at its full size."""
    gen = torch.Generator().manual_seed(11)
    f = torch.rand([240000, 10, 5], generator=gen)
    n = torch.randint(1, 10, [240000], generator=gen)
    out = HardSimpleVFE(num_features=5)(f.cuda(), n.cuda(), None)
    assert out.shape == torch.Size([240000, 5])
    assert_same_bits(out.cpu().numpy(), vfe_mean.hard_simple_vfe(f.numpy(), n.numpy(), 5), "rand 240000")


def test_standalone_device_side_row_limit_and_empty():
    f = torch.rand([1000, 5, 4]).cuda()
    n = torch.randint(1, 6, [1000], dtype=torch.int32).cuda()
    lim = torch.tensor([300], dtype=torch.int32).cuda()
    out = torch.full((1000, 4), -7.0).cuda()
    rc = _cabi.lib().pcfe_voxel_mean_f32(f.data_ptr(), n.data_ptr(), lim.data_ptr(), 1000, 5, 4, out.data_ptr(), 0,
                                         torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    o = out.cpu().numpy()
    assert_same_bits(o[:300], vfe_mean.hard_simple_vfe(f.cpu().numpy()[:300], n.cpu().numpy()[:300]), "limited rows")
    assert np.all(o[300:] == -7.0)
    e = hard_simple_vfe(torch.empty((0, 5, 4)).cuda(), torch.empty((0,), dtype=torch.int32).cuda())
    assert e.shape == (0, 4)


@pytest.fixture(params=["record", "map0", "map1", "cluster", "fallback", "general", "global"])
def mean_mode(request):
    """record: the fused epilogue of the expansion kernel; cluster: the same with the frame's partition +
    grouping done by one thread-block cluster (hv_cluster.cuh); fallback: every frame forced through the
    overflow fallback (its own mean epilogue); general / global: paths without the epilogue, where
    the wrapper runs voxelization + the stand-alone kernel."""
    mode = request.param
    _cabi.debug_set("hv_path", 1 if mode == "global" else 0)
    _cabi.debug_set("hv_force_overflow", 1 if mode == "fallback" else 0)
    _cabi.debug_set("hv_bucket_variant", 1 if mode == "general" else 0)
    _cabi.debug_set("hv_cluster", 1 if mode == "cluster" else 0)
    _cabi.debug_set("hv_expand_map", {"map0": 0, "map1": 1}.get(mode, 2))  # record expansion: which tiles a warp takes
    yield mode
    _cabi.debug_set("hv_expand_map", 2)
    for k in ("hv_path", "hv_force_overflow", "hv_bucket_variant"):
        _cabi.debug_set(k, 0)
    _cabi.debug_set("hv_cluster", 0)


@pytest.mark.parametrize("cfg_name,ci,cap", [("C4", 4, None), ("C1", 1, None), ("C4", 4, 3000), ("C5", 5, 2000)])
def test_fused_batch_vs_oracle(cfg_name, ci, cap, mean_mode):
    cfg = synth.CONFIGS[cfg_name]
    vs, rg, P = cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"]
    V = cap or cfg["max_voxels"]
    frames = [synth.lidar_frame(n, cfg["c"], 7000 + 10 * ci + k, cfg["r_max"]).numpy()
              for k, n in enumerate((30000, 1, 0, 17001, 64))]
    frames[0][9, 2] = np.nan
    means, coors, num, vnum = voxelize_mean_batch([torch.from_numpy(p).cuda() for p in frames], vs, rg, P, V)
    counts = vnum.cpu().tolist()
    for k, p in enumerate(frames):
        ev, ec, en = oracle.hard_voxelize(p, vs, rg, P, V)
        m = counts[k]
        assert m == len(en), f"frame {k}"
        assert_same_bits(coors[k, :m].cpu().numpy(), ec, f"{cfg_name} frame {k} coors")
        assert_same_bits(num[k, :m].cpu().numpy(), en, f"{cfg_name} frame {k} num")
        assert_same_bits(means[k, :m].cpu().numpy(), vfe_mean.hard_simple_vfe(ev, en), f"{cfg_name} frame {k} means")


def test_fused_with_points_range_filter(mean_mode):
    cfg = synth.CONFIGS["C4"]
    vs, rg, P, V = cfg["voxel_size"], cfg["point_cloud_range"], 5, 20000
    fr = [rg[0] + 3.0, rg[1] + 1.5, rg[2] + 0.2, rg[3] - 7.0, rg[4] - 2.5, rg[5] - 0.4]
    lo, hi = np.asarray(fr[:3], np.float32), np.asarray(fr[3:], np.float32)
    frames = [synth.lidar_frame(25000 + k, 5, 7100 + k, cfg["r_max"]).numpy() for k in range(3)]
    means, coors, num, vnum = voxelize_mean_batch([torch.from_numpy(p).cuda() for p in frames], vs, rg, P, V,
                                                  points_range=fr)
    counts = vnum.cpu().tolist()
    for k, p in enumerate(frames):
        keep = np.all(p[:, :3] > lo, axis=1) & np.all(p[:, :3] < hi, axis=1)
        ev, ec, en = oracle.hard_voxelize(np.ascontiguousarray(p[keep]), vs, rg, P, V)
        m = counts[k]
        assert m == len(en)
        assert_same_bits(coors[k, :m].cpu().numpy(), ec, f"frame {k} coors")
        assert_same_bits(means[k, :m].cpu().numpy(), vfe_mean.hard_simple_vfe(ev, en), f"frame {k} means")


def test_fused_entry_rejects_other_shapes():
    """The C-ABI entry is the record path only (include/pcfe.h): P != 5 or C not in (4, 5) is
    PCFE_ERR_SHAPE before anything is launched."""
    L = _cabi.lib()
    fr = (_cabi.Frame * 1)()
    for c, p in ((3, 5), (5, 64), (4, 4)):
        rc = L.pcfe_hard_voxelize_mean_batch_f32(fr, 1, c, _cabi.f3([0.1, 0.1, 0.1]), _cabi.f6([0, 0, 0, 1, 1, 1]), None, p,
                                                 100, None, None, 0, 0, None)
        assert rc == _cabi.ERR_SHAPE


def test_full_c4_batch_mean_properties():
    """BASELINE C4 at full size (64 x 180 000 x 5): the fused means equal the stand-alone encoder
    applied to the plain voxelization of the same frames, bit for bit."""
    from detmatch_b200.ops import voxelize_batch
    cfg = synth.CONFIGS["C4"]
    vs, rg, P, V = cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"]
    pts = [synth.lidar_frame(cfg["n"], 5, synth.seed_for(4, k), cfg["r_max"]).cuda() for k in range(cfg["frames"])]
    means, coors, num, vnum = voxelize_mean_batch(pts, vs, rg, P, V)
    vox, coors2, num2, vnum2 = voxelize_batch(pts, vs, rg, P, V, sync=False)
    counts = vnum.cpu().tolist()
    assert counts == vnum2.cpu().tolist()
    for k, m in enumerate(counts):
        assert torch.equal(coors[k, :m], coors2[k, :m]) and torch.equal(num[k, :m], num2[k, :m])
        ref = hard_simple_vfe(vox[k, :m], num2[k, :m])
        assert torch.equal(means[k, :m].view(torch.int32), ref.view(torch.int32)), f"frame {k}"
    # one frame all the way down to the oracle
    ev, ec, en = oracle.hard_voxelize(pts[3].cpu().numpy(), vs, rg, P, V)
    assert_same_bits(means[3, :counts[3]].cpu().numpy(), vfe_mean.hard_simple_vfe(ev, en), "frame 3 means")
