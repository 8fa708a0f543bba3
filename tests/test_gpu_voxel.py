"""GPU parity tests for hard / dynamic voxelization: CUDA path (through the C ABI) vs the CPU
oracle and the committed golden vectors.  Bar: bit-exact voxels, coors, num_points_per_voxel."""
import numpy as np
import pytest
import torch

from detmatch_b200 import synth
from detmatch_b200.ops import Voxelization, voxelization, voxelize_batch
from detmatch_b200.ops.voxel import voxel_layer
from oracle import oracle
from tests.helpers import assert_same_bits, golden, golden_names

pytestmark = pytest.mark.gpu

KITTI = [0, -40, -3, 70.4, 40, 1]


@pytest.fixture(autouse=True, params=["launches", "map0", "map1", "dedup", "cluster", "bucket_general", "global", "fallback"])
def hv_mode(request):
    """Every test runs against all hard-voxelize implementations behind the one entry point: the
    bucket path as a launch sequence (default; record kernels where P == 5 and C = 4 / 5), the record
    same with warp-level __match_any_sync key de-duplication in front of the bucket table, the record
    path with one thread-block cluster per frame (hv_cluster.cuh: measured slower, kept as the
    round-2 DSMEM experiment), the general bucket kernels (register-sorted chains for P <= 8, bitonic ranks
    otherwise), the global-memory path, and the default path with every frame forced through its
    overflow fallback."""
    from detmatch_b200 import _cabi
    mode = request.param
    _cabi.debug_set("hv_path", 1 if mode == "global" else 0)
    _cabi.debug_set("hv_cluster", 1 if mode == "cluster" else 0)
    _cabi.debug_set("hv_warp_dedup", 1 if mode == "dedup" else 0)
    _cabi.debug_set("hv_force_overflow", 1 if mode == "fallback" else 0)
    _cabi.debug_set("hv_bucket_variant", 1 if mode == "bucket_general" else 0)
    _cabi.debug_set("hv_expand_map", {"map0": 0, "map1": 1}.get(mode, 2))  # record expansion: which tiles a warp takes
    yield mode
    _cabi.debug_set("hv_expand_map", 2)
    _cabi.debug_set("hv_path", 0)
    _cabi.debug_set("hv_cluster", 0)
    _cabi.debug_set("hv_warp_dedup", 0)
    _cabi.debug_set("hv_force_overflow", 0)
    _cabi.debug_set("hv_bucket_variant", 0)


def _oracle_frame(cfg_name, ci, k):
    """Oracle outputs of frame k of a BASELINE config (LiDAR-like generator; ~0.1 s per frame)."""
    cfg = synth.CONFIGS[cfg_name]
    p = synth.lidar_frame(cfg["n"], cfg["c"], synth.seed_for(ci, k), cfg["r_max"]).numpy()
    return oracle.hard_voxelize(p, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"])


def _gpu_hard(pts, vs, rg, p, v):
    out = voxelization(torch.from_numpy(np.ascontiguousarray(pts)).cuda(), list(vs), list(rg), int(p), int(v))
    return [o.cpu().numpy() for o in out]


def _check_hard(pts, vs, rg, p, v, what=""):
    ev, ec, en = oracle.hard_voxelize(pts, vs, rg, p, v)
    gv, gc, gn = _gpu_hard(pts, vs, rg, p, v)
    assert_same_bits(gc, ec, what + " coors")
    assert_same_bits(gn, en, what + " num")
    assert_same_bits(gv, ev, what + " voxels")
    return len(en)


@pytest.mark.parametrize("name", [n for n in golden_names("voxel_") if n != "voxel_generator_kat"])
def test_golden_hard_and_dynamic(name):
    g = golden(name)
    gv, gc, gn = _gpu_hard(g["points"], g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    assert_same_bits(gc, g["coors"], "coors")
    assert_same_bits(gn, g["num"], "num")
    assert_same_bits(gv, g["voxels"], "voxels")
    if "dyn_coors" in g:
        d = voxelization(torch.from_numpy(g["points"]).cuda(), list(g["voxel_size"]), list(g["range"]), -1, -1)
        assert_same_bits(d.cpu().numpy(), g["dyn_coors"], "dyn")


def test_voxel_generator_kat_float32():
    """tests/test_models/test_voxel_encoder/test_voxel_generator.py:6-22 literal (float32 input)."""
    g = golden("voxel_generator_kat")
    gv, gc, gn = _gpu_hard(g["points64"].astype(np.float32), g["voxel_size"], g["range"], g["max_points"], g["max_voxels"])
    assert_same_bits(gc, g["coors"], "coors")
    assert_same_bits(gn, g["num"], "num")
    assert_same_bits(gv, g["voxels32"], "voxels")


def test_reference_test_voxelization_flow():
    """tests/test_models/test_voxel_encoder/test_voxelize.py:15-83 with the nn.Module API."""
    g = golden("voxel_kitti_fixture")
    vs, rg = [0.5, 0.5, 0.5], KITTI
    hard = Voxelization(vs, rg, 1000)
    dyn = Voxelization(vs, rg, -1)
    points = torch.from_numpy(g["points"]).contiguous().to("cuda:0")
    voxels, coors, num = hard.forward(points)
    assert np.all(coors.cpu().numpy() == g["coors"])
    assert np.all(voxels.cpu().numpy() == g["voxels"])
    assert np.all(num.cpu().numpy() == g["num"])
    dc = dyn.forward(points).cpu().numpy()
    pts = g["points"]
    for i in range(g["voxels"].shape[0]):
        idx = np.all(dc == g["coors"][i], axis=1)
        k = pts[idx].shape[0]
        assert k > 0 and k == g["num"][i]
        assert np.all(pts[idx] == g["voxels"][i][:k])


@pytest.mark.parametrize("cfg_name,ci", [("C1", 1), ("C4", 4), ("C5", 5)])
@pytest.mark.parametrize("kind", ["lidar", "uniform"])
def test_full_size_frame_vs_oracle(cfg_name, ci, kind):
    cfg = synth.CONFIGS[cfg_name]
    if kind == "lidar":
        pts = synth.lidar_frame(cfg["n"], cfg["c"], synth.seed_for(ci, 0), cfg["r_max"]).numpy()
    else:
        pts = synth.uniform_frame(cfg["n"], cfg["c"], synth.seed_for(ci, 1), cfg["point_cloud_range"]).numpy()
    m = _check_hard(pts, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"],
                    f"{cfg_name}/{kind}")
    assert m > 1000
    d = voxelization(torch.from_numpy(pts).cuda(), cfg["voxel_size"], cfg["point_cloud_range"], -1, -1)
    assert_same_bits(d.cpu().numpy(), oracle.dynamic_voxelize(pts, cfg["voxel_size"], cfg["point_cloud_range"]), "dyn")


@pytest.mark.parametrize("p,v", [(1, 7), (3, 150), (5, 100000), (64, 50), (200, 3)])
def test_caps(p, v):
    rng = np.random.default_rng(p * 1000 + v)
    pts = np.concatenate([rng.uniform([0, -4, -3], [8, 4, 1], size=(20011, 3)), rng.uniform(0, 1, size=(20011, 1))],
                         axis=1).astype(np.float32)
    _check_hard(pts, [0.5, 0.5, 0.5], KITTI, p, v, f"caps P={p} V={v}")


@pytest.mark.parametrize("seed", range(24))
def test_random_small_configs(seed):
    """Randomised frames, grids and caps (seeded): sizes around the tile / warp / bucket boundaries,
    duplicated points, points on cell faces, out-of-range and non-finite rows, every row length the
    kernels specialise on -- against the CPU op on every hard-voxelize path."""
    rng = np.random.default_rng(7700 + seed)
    n = int(rng.choice([1, 2, 33, 257, 4095, 4096, 4097, 9000, 30011]))
    c = int(rng.choice([3, 4, 5, 6]))
    p = int(rng.choice([1, 2, 5, 5, 5, 8, 9, 35]))
    vs = [float(rng.choice([0.05, 0.1, 0.25, 0.5, 1.0])) for _ in range(2)] + [float(rng.choice([0.1, 0.5, 4.0]))]
    lo = np.array([rng.uniform(-60, 0), rng.uniform(-40, 0), rng.uniform(-5, -1)])
    span = np.array([rng.uniform(4, 80), rng.uniform(4, 80), 4.0])
    rg = [float(x) for x in np.concatenate([lo, lo + span])]
    xyz = rng.uniform(lo - 0.05 * span, lo + 1.05 * span, size=(n, 3))
    k = n // 3
    if k:  # duplicates (same cell, same coordinates) and points exactly on cell faces
        xyz[rng.integers(0, n, k)] = xyz[rng.integers(0, n, k)]
        on = rng.integers(0, n, k)
        xyz[on] = lo + np.round((xyz[on] - lo) / np.array(vs)) * np.array(vs)
    pts = np.concatenate([xyz, rng.uniform(0, 1, size=(n, c - 3))], axis=1).astype(np.float32)
    if n > 40:
        pts[7, 0] = np.nan
        pts[11, 1] = np.inf
        pts[13, 2] = -np.inf
        pts[17, 0] = 3e38
    cells = len(np.unique(oracle.dynamic_voxelize(pts, vs, rg), axis=0))
    v = int(rng.choice([1, max(1, cells // 2), cells + 5, 200000]))
    _check_hard(pts, vs, rg, p, v, f"random seed={seed} N={n} C={c} P={p} V={v}")
    d = voxelization(torch.from_numpy(pts).cuda(), vs, rg, -1, -1)
    assert_same_bits(d.cpu().numpy(), oracle.dynamic_voxelize(pts, vs, rg), "dyn")


def test_heavy_contention_single_voxel():
    """All points in very few voxels: the sorted per-voxel lists must still come out in index order."""
    rng = np.random.default_rng(5)
    pts = np.concatenate([rng.uniform([1.0, 1.0, -1.0], [1.9, 1.4, -0.6], size=(50000, 3)),
                          rng.uniform(0, 1, size=(50000, 2))], axis=1).astype(np.float32)
    for p in (1, 5, 64, 333):
        _check_hard(pts, [0.5, 0.5, 0.5], KITTI, p, 100, f"contention P={p}")


@pytest.mark.parametrize("c", [3, 4, 5, 7])
@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 4097])
def test_shapes(c, n):
    rng = np.random.default_rng(c * 100 + n)
    pts = np.concatenate([rng.uniform([-2, -42, -3.5], [72, 42, 1.5], size=(n, 3)), rng.uniform(0, 1, size=(n, c - 3))],
                         axis=1).astype(np.float32)
    _check_hard(pts, [0.4, 0.4, 0.8], KITTI, 3, 500, f"C={c} N={n}")
    d = voxelization(torch.from_numpy(pts).cuda(), [0.4, 0.4, 0.8], KITTI, -1, -1)
    assert_same_bits(d.cpu().numpy(), oracle.dynamic_voxelize(pts, [0.4, 0.4, 0.8], KITTI), "dyn")


def test_empty_and_all_out_of_range():
    vs = [0.5, 0.5, 0.5]
    v, c, n = voxelization(torch.zeros((0, 4), device="cuda"), vs, KITTI, 5, 10)
    assert v.shape == (0, 5, 4) and c.shape == (0, 3) and n.shape == (0,)
    assert voxelization(torch.zeros((0, 4), device="cuda"), vs, KITTI, -1, -1).shape == (0, 3)
    far = torch.full((100, 4), 1e6, device="cuda")
    v, c, n = voxelization(far, vs, KITTI, 5, 10)
    assert v.shape[0] == 0
    assert (voxelization(far, vs, KITTI, -1, -1) == -1).all()
    v, c, n = voxelization(torch.rand((100, 4), device="cuda"), vs, KITTI, 5, 0)
    assert v.shape[0] == 0


def test_special_values_nan_inf():
    g = golden("voxel_special")
    pts = torch.from_numpy(g["points"]).cuda()
    d = voxelization(pts, list(g["voxel_size"]), list(g["range"]), -1, -1).cpu().numpy()
    assert_same_bits(d, g["dyn_coors"], "special dyn")
    assert (d[8:16] == -1).all()  # NaN / Inf / huge rows


def test_voxel_layer_dropin_signature():
    """voxel_layer.hard_voxelize(points, voxels, coors, num, vs, range, P, V, 3) -> int, in place."""
    g = golden("voxel_caps")
    pts = torch.from_numpy(g["points"]).cuda()
    P, V = int(g["max_points"]), int(g["max_voxels"])
    voxels = pts.new_zeros((V, P, pts.size(1)))
    coors = pts.new_zeros((V, 3), dtype=torch.int)
    num = pts.new_zeros((V,), dtype=torch.int)
    m = voxel_layer.hard_voxelize(pts, voxels, coors, num, list(g["voxel_size"]), list(g["range"]), P, V, 3)
    assert isinstance(m, int) and m == len(g["num"])
    assert_same_bits(voxels[:m].cpu().numpy(), g["voxels"], "voxels")
    assert_same_bits(coors[:m].cpu().numpy(), g["coors"], "coors")
    assert_same_bits(num[:m].cpu().numpy(), g["num"], "num")
    with pytest.raises(RuntimeError, match="contiguous"):
        voxel_layer.hard_voxelize(pts.t().contiguous().t(), voxels, coors, num, [0.5] * 3, KITTI, P, V, 3)
    with pytest.raises(TypeError):
        voxelization(pts.double(), [0.5] * 3, KITTI, P, V)


def test_cpu_tensor_round_trip():
    g = golden("voxel_caps")
    v, c, n = voxelization(torch.from_numpy(g["points"]), list(g["voxel_size"]), list(g["range"]),
                           int(g["max_points"]), int(g["max_voxels"]))
    assert not v.is_cuda
    assert_same_bits(v.numpy(), g["voxels"], "voxels")
    assert_same_bits(c.numpy(), g["coors"], "coors")


@pytest.mark.parametrize("cfg_name,ci,frames", [("C1", 1, 3), ("C4", 4, 6), ("C5", 5, 3)])
def test_batched_ragged_vs_oracle(cfg_name, ci, frames):
    cfg = synth.CONFIGS[cfg_name]
    sizes = [cfg["n"] // 4 - 17 * k for k in range(frames)]
    sizes[1] = 0 if frames > 2 else sizes[1]  # an empty frame in the middle of the batch
    pts = [synth.lidar_frame(n, cfg["c"], synth.seed_for(ci, 30 + k), cfg["r_max"]) if n else torch.zeros((0, cfg["c"]))
           for k, n in enumerate(sizes)]
    mv = cfg["max_voxels"] // 4
    vox, coors, num, vnum = voxelize_batch([p.cuda() for p in pts], cfg["voxel_size"], cfg["point_cloud_range"],
                                           cfg["max_num_points"], mv, sync=False)
    counts = vnum.cpu().tolist()
    exp = [oracle.hard_voxelize(p.numpy(), cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], mv) for p in pts]
    for k, (ev, ec, en) in enumerate(exp):
        assert counts[k] == len(en)
        assert_same_bits(coors[k, :counts[k]].cpu().numpy(), ec, f"frame {k} coors")
        assert_same_bits(num[k, :counts[k]].cpu().numpy(), en, f"frame {k} num")
        assert_same_bits(vox[k, :counts[k]].cpu().numpy(), ev, f"frame {k} voxels")
    # detector-style concatenation (openpcdet.py:69-76)
    vc, nc, cb = voxelize_batch([p.cuda() for p in pts], cfg["voxel_size"], cfg["point_cloud_range"],
                                cfg["max_num_points"], mv, sync=True)
    assert_same_bits(vc.cpu().numpy(), np.concatenate([e[0] for e in exp]), "cat voxels")
    assert_same_bits(nc.cpu().numpy(), np.concatenate([e[2] for e in exp]), "cat num")
    ecb = np.concatenate([np.pad(e[1], ((0, 0), (1, 0)), constant_values=k) for k, e in enumerate(exp)])
    assert_same_bits(cb.cpu().numpy(), ecb.astype(np.int32), "coors_batch")


def test_many_frames_more_than_one_wave():
    """More frames than fit one wave / one kernel-parameter table (64)."""
    cfg = synth.CONFIGS["C1"]
    pts = [synth.lidar_frame(3000 + 7 * k, 4, 9000 + k, 80.0) for k in range(70)]
    vox, coors, num, vnum = voxelize_batch([p.cuda() for p in pts], cfg["voxel_size"], cfg["point_cloud_range"], 5, 800,
                                           sync=False)
    counts = vnum.cpu().tolist()
    for k in (0, 1, 33, 63, 64, 69):
        ev, ec, en = oracle.hard_voxelize(pts[k].numpy(), cfg["voxel_size"], cfg["point_cloud_range"], 5, 800)
        assert counts[k] == len(en)
        assert_same_bits(vox[k, :counts[k]].cpu().numpy(), ev, f"frame {k}")
        assert_same_bits(coors[k, :counts[k]].cpu().numpy(), ec, f"frame {k}")


def test_full_c4_batch_properties():
    """BASELINE config C4 at full size (64 x 180k x 5): size-independent properties on every frame
    (computed on the GPU with torch ops), plus the bit-exact oracle comparison of EVERY frame ."""
    cfg = synth.CONFIGS["C4"]
    F = 64
    pts = [synth.lidar_frame(cfg["n"], cfg["c"], synth.seed_for(4, k), cfg["r_max"]).cuda() for k in range(F)]
    P, V = cfg["max_num_points"], cfg["max_voxels"]
    vox, coors, num, vnum = voxelize_batch(pts, cfg["voxel_size"], cfg["point_cloud_range"], P, V, sync=False)
    counts = vnum.cpu().tolist()
    gx, gy, gz = 1504, 1504, 40
    for k in range(F):
        m = counts[k]
        dyn = voxelization(pts[k], cfg["voxel_size"], cfg["point_cloud_range"], -1, -1).long()
        valid = dyn[:, 0] >= 0
        keys = (dyn[:, 0] * gy + dyn[:, 1]) * gx + dyn[:, 2]
        uniq, inv = torch.unique(keys[valid], return_inverse=True)
        assert m == min(len(uniq), V)
        c = coors[k, :m].long()
        vkeys = (c[:, 0] * gy + c[:, 1]) * gx + c[:, 2]
        assert len(torch.unique(vkeys)) == m                      # every voxel once
        # first-occurrence order: first point index of consecutive voxels is increasing
        idx = torch.arange(len(keys), device="cuda")[valid]
        first = torch.full((len(uniq),), len(keys), device="cuda", dtype=torch.long).scatter_reduce(0, inv, idx, "amin")
        pos = torch.searchsorted(uniq, vkeys)
        assert (uniq[pos] == vkeys).all()
        fo = first[pos]
        assert (fo[1:] > fo[:-1]).all()
        # counts: min(#points in voxel, P); padding is exactly zero; slot 0 is the first point
        cnt = torch.bincount(inv, minlength=len(uniq))[pos].clamp(max=P)
        assert (num[k, :m].long() == cnt).all()
        slot = torch.arange(P, device="cuda")[None, :, None]
        pad = vox[k, :m] * (slot >= num[k, :m, None, None]).float()
        assert (pad == 0).all() and (vox[k, :m].view(torch.int32)[(slot >= num[k, :m, None, None]).expand(-1, -1, 5)] == 0).all()
        assert (vox[k, :m, 0] == pts[k][fo]).all()
    for k in range(F):
        ev, ec, en = _oracle_frame("C4", 4, k)
        assert counts[k] == len(en)
        assert_same_bits(vox[k, :counts[k]].cpu().numpy(), ev, f"frame {k} voxels")
        assert_same_bits(coors[k, :counts[k]].cpu().numpy(), ec, f"frame {k} coors")
        assert_same_bits(num[k, :counts[k]].cpu().numpy(), en, f"frame {k} num")


def test_full_c5_batch_16_frames_vs_oracle():
    """BASELINE config C5 at its per-GPU size at 8 GPUs (128 / 8 = 16 frames x 300k x 5, P = 64,
    V = 40 000): every frame bit-exact against the oracle, one batched call."""
    cfg = synth.CONFIGS["C5"]
    F = 16
    pts = [synth.lidar_frame(cfg["n"], cfg["c"], synth.seed_for(5, k), cfg["r_max"]).cuda() for k in range(F)]
    P, V = cfg["max_num_points"], cfg["max_voxels"]
    vox, coors, num, vnum = voxelize_batch(pts, cfg["voxel_size"], cfg["point_cloud_range"], P, V, sync=False)
    counts = vnum.cpu().tolist()
    for k in range(F):
        ev, ec, en = _oracle_frame("C5", 5, k)
        assert counts[k] == len(en) and counts[k] > 10000
        assert_same_bits(coors[k, :counts[k]].cpu().numpy(), ec, f"frame {k} coors")
        assert_same_bits(num[k, :counts[k]].cpu().numpy(), en, f"frame {k} num")
        assert_same_bits(vox[k, :counts[k]].cpu().numpy(), ev, f"frame {k} voxels")


@pytest.mark.parametrize("cfg_name,ci", [("C4", 4), ("C1", 1)])
def test_sweep_ordered_frames_vs_oracle(cfg_name, ci):
    """Un-shuffled frames in a spinning sensor's firing order (test-time input: consecutive points share
    voxels; the warp-level key de-duplication has real groups to merge here): full-size frames against the
    oracle on every path."""
    cfg = synth.CONFIGS[cfg_name]
    pts = [synth.lidar_frame(cfg["n"], cfg["c"], synth.seed_for(ci, 70 + k), cfg["r_max"], order="sweep") for k in range(3)]
    P, V = cfg["max_num_points"], cfg["max_voxels"]
    vox, coors, num, vnum = voxelize_batch([p.cuda() for p in pts], cfg["voxel_size"], cfg["point_cloud_range"], P, V, sync=False)
    counts = vnum.cpu().tolist()
    for k, p in enumerate(pts):
        ev, ec, en = oracle.hard_voxelize(p.numpy(), cfg["voxel_size"], cfg["point_cloud_range"], P, V)
        assert counts[k] == len(en)
        assert_same_bits(coors[k, :counts[k]].cpu().numpy(), ec, f"frame {k} coors")
        assert_same_bits(num[k, :counts[k]].cpu().numpy(), en, f"frame {k} num")
        assert_same_bits(vox[k, :counts[k]].cpu().numpy(), ev, f"frame {k} voxels")


def test_c2_dynamic_batch_16_frames():
    """BASELINE config C2: 16 KITTI-shape frames, dynamic voxelization, vs the oracle."""
    cfg = synth.CONFIGS["C2"]
    for k in range(cfg["frames"]):
        p = synth.lidar_frame(cfg["n"], cfg["c"], synth.seed_for(2, k), cfg["r_max"])
        d = voxelization(p.cuda(), cfg["voxel_size"], cfg["point_cloud_range"], -1, -1)
        assert_same_bits(d.cpu().numpy(), oracle.dynamic_voxelize(p.numpy(), cfg["voxel_size"], cfg["point_cloud_range"]),
                         f"C2 frame {k}")


def test_repeatable():
    """Atomics must not leak scheduling order into the result: two runs are bit-identical."""
    cfg = synth.CONFIGS["C5"]
    p = synth.lidar_frame(100000, 5, 4242, 60.0).cuda()
    a = voxelization(p, cfg["voxel_size"], cfg["point_cloud_range"], 64, 40000)
    b = voxelization(p, cfg["voxel_size"], cfg["point_cloud_range"], 64, 40000)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


@pytest.mark.parametrize("c", [4, 5])
def test_unaligned_buffers(c):
    """Points / voxels buffers that are only 4-byte aligned (views into larger tensors) must take
    the scalar paths and still be exact; also exercises the last-row guard of the vector loads."""
    from detmatch_b200 import _cabi
    from detmatch_b200._torch_glue import ptr, stream_ptr, workspace
    cfg = synth.CONFIGS["C4"]
    n, P, V = 30011, 5, 20000
    base = synth.lidar_frame(n + 1, c, 777, 80.0).cuda()
    pts = base.view(-1)[c:].view(n, c)                       # starts 4*c bytes into the allocation
    assert pts.data_ptr() % 16 != 0 or c == 4
    vox_store = torch.empty(V * P * c + 1, dtype=torch.float32, device="cuda")
    voxels = vox_store[1:].view(V, P, c)                       # 4-byte aligned only
    coors = torch.empty((V, 3), dtype=torch.int32, device="cuda")
    num = torch.empty((V,), dtype=torch.int32, device="cuda")
    vnum = torch.empty(1, dtype=torch.int32, device="cuda")
    L = _cabi.lib()
    vs, rg = _cabi.f3(cfg["voxel_size"]), _cabi.f6(cfg["point_cloud_range"])
    ws = workspace(pts.device, L.pcfe_hard_voxelize_workspace_bytes(n, 1, 1, vs, rg, P, V))
    rc = L.pcfe_hard_voxelize_f32(ptr(pts), n, c, vs, rg, P, V, ptr(voxels), ptr(coors), ptr(num), ptr(vnum),
                                  ptr(ws), ws.numel(), 0, stream_ptr(pts.device))
    assert rc == 0
    m = int(vnum.item())
    ev, ec, en = oracle.hard_voxelize(pts.cpu().numpy(), cfg["voxel_size"], cfg["point_cloud_range"], P, V)
    assert m == len(en)
    assert_same_bits(voxels[:m].cpu().numpy(), ev, "voxels")
    assert_same_bits(coors[:m].cpu().numpy(), ec, "coors")
    assert_same_bits(num[:m].cpu().numpy(), en, "num")
    # aligned buffers whose LAST point is kept: the two-chunk row load must not read past the end
    pts2 = synth.lidar_frame(4096, c, 778, 80.0)
    pts2[-1, :3] = torch.tensor([1.0, 1.0, 0.0])
    _check_hard(pts2.numpy(), cfg["voxel_size"], cfg["point_cloud_range"], P, V, "last row")


@pytest.mark.parametrize("lo,vs,hi", [(0.0, 0.05, 70.4), (-40.0, 0.05, 40.0), (-3.0, 0.1, 1.0), (-75.2, 0.1, 75.2),
                                      (-2.0, 0.15, 4.0), (-50.0, 0.25, 50.0), (-5.0, 8.0, 3.0), (-51.2, 0.2, 51.2),
                                      (-1.0, 1.0 / 3.0, 1.0), (0.5, 1e-3, 9.5), (-1e4, 7.77, 1e4)])
def test_fast_division_equals_ieee_divide_exhaustive(lo, vs, hi, hv_mode):
    """The bin kernel divides with a hoisted reciprocal + three fused steps (ptxas's own fast path)
    instead of __fdiv_rn per point: checked for ALL 2^32 float32 values of a coordinate."""
    if hv_mode != "launches":
        pytest.skip("path-independent")
    from detmatch_b200 import _cabi
    out = torch.zeros(2, dtype=torch.int64, device="cuda")
    rc = _cabi.lib().pcfe_debug_axis_sweep(lo, vs, hi, out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    bad, first = out.cpu().tolist()
    assert bad == 0, f"{bad} coordinates disagree with the IEEE divide, first bits {first & 0xFFFFFFFF:#x}"


@pytest.mark.parametrize("cfg_name,ci", [("C1", 1), ("C4", 4), ("C5", 5)])
def test_fused_points_range_filter(cfg_name, ci):
    """voxelize_batch(points_range=...) == PointsRangeFilter (base_points.py:223-228, strict on both
    sides, float32) followed by hard voxelization of the filtered frames: bit for bit, including
    points exactly on the filter faces, NaN rows and a filter range tighter than the voxel range."""
    cfg = synth.CONFIGS[cfg_name]
    rg = cfg["point_cloud_range"]
    frames = []
    for k in range(3):
        p = synth.lidar_frame(20000 + 1000 * k, cfg["c"], 5000 + 10 * ci + k, cfg["r_max"]).numpy().copy()
        p[5, 0] = rg[0]          # on the lower face: the filter drops it, plain voxelization keeps it
        p[6, 1] = rg[4]          # on the upper face
        p[7, 2] = np.nan
        p[8, 0] = np.float32(rg[3]) - np.float32(1e-3)
        frames.append(p)
    for fr in (rg, [rg[0] + 3.0, rg[1] + 1.5, rg[2] + 0.2, rg[3] - 7.0, rg[4] - 2.5, rg[5] - 0.4]):
        lo, hi = np.asarray(fr[:3], np.float32), np.asarray(fr[3:], np.float32)
        P, V = cfg["max_num_points"], min(cfg["max_voxels"], 9000)
        vox, coors, num, vnum = voxelize_batch([torch.from_numpy(p).cuda() for p in frames], cfg["voxel_size"], rg, P, V,
                                               sync=False, points_range=fr)
        counts = vnum.cpu().tolist()
        for k, p in enumerate(frames):
            keep = np.all(p[:, :3] > lo, axis=1) & np.all(p[:, :3] < hi, axis=1)
            ev, ec, en = oracle.hard_voxelize(np.ascontiguousarray(p[keep]), cfg["voxel_size"], rg, P, V)
            assert counts[k] == len(en)
            assert_same_bits(coors[k, :counts[k]].cpu().numpy(), ec, f"{cfg_name} frame {k} coors")
            assert_same_bits(num[k, :counts[k]].cpu().numpy(), en, f"{cfg_name} frame {k} num")
            assert_same_bits(vox[k, :counts[k]].cpu().numpy(), ev, f"{cfg_name} frame {k} voxels")


def test_empty_frames_in_their_own_wave():
    """A wave whose frames are all empty (hv_wave = 1 with empty frames in the middle and at the end)
    must not launch a zero-sized grid: voxel_num = 0 for those frames, the others unaffected."""
    from detmatch_b200 import _cabi
    cfg = synth.CONFIGS["C1"]
    full = synth.lidar_frame(6000, 4, 123, 80.0)
    empty = torch.zeros((0, 4))
    frames = [full, empty, full[:777], empty]
    _cabi.debug_set("hv_wave", 1)
    try:
        vox, coors, num, vnum = voxelize_batch([p.cuda() for p in frames], cfg["voxel_size"], cfg["point_cloud_range"], 5, 900,
                                               sync=False)
        counts = vnum.cpu().tolist()
    finally:
        _cabi.debug_set("hv_wave", 0)
    assert counts[1] == 0 and counts[3] == 0
    for k in (0, 2):
        ev, ec, en = oracle.hard_voxelize(frames[k].numpy(), cfg["voxel_size"], cfg["point_cloud_range"], 5, 900)
        assert counts[k] == len(en)
        assert_same_bits(vox[k, :counts[k]].cpu().numpy(), ev, f"frame {k}")
        assert_same_bits(coors[k, :counts[k]].cpu().numpy(), ec, f"frame {k}")


def test_workspace_query_covers_three_feature_rows():
    """The size query has no row length argument: with C = 3 and 257 <= max_points <= 341 the call
    selects the bucket plan (larger scratch) and must fit what the query returned."""
    rng = np.random.default_rng(5)
    pts = rng.uniform([0, -4, -3], [8, 4, 1], size=(5000, 3)).astype(np.float32)
    _check_hard(pts, [0.5, 0.5, 0.5], KITTI, 300, 50, "C=3 P=300")


@pytest.mark.parametrize("P,C", [(64, 5), (32, 4), (20, 4)])
def test_long_voxels_into_dirty_output_buffers(P, C):
    """Pillar shapes (the hvb_expand_words path) through a pre-allocated plan whose outputs start out as NaN
    bit patterns and are reused by a second run: every word of a row < voxel_num must have been written --
    short pillars, a pillar longer than P, frames with fewer points than max_voxels, an empty frame."""
    from detmatch_b200.ops.voxel import HardVoxelizeBatchPlan
    cfg = synth.CONFIGS["C5"]
    vs, rg, V = cfg["voxel_size"], cfg["point_cloud_range"], 3000
    frames = [synth.lidar_frame(n, C, 8800 + 13 * k + P, cfg["r_max"]).numpy() for k, n in enumerate((40000, 2500, 0, 9000))]
    rng = np.random.default_rng(5 + P)
    for fr, (lo, cnt) in zip((frames[0], frames[3]), ((100, 200), (5, 37))):  # one very dense and one medium pillar
        fr[lo:lo + cnt, 0] = 10.3 + 0.2 * rng.random(cnt).astype(np.float32)
        fr[lo:lo + cnt, 1] = -7.1 + 0.2 * rng.random(cnt).astype(np.float32)
    plan = HardVoxelizeBatchPlan([len(f) for f in frames], C, vs, rg, P, V, "cuda:0")
    for t in (plan.voxels, plan.coors, plan.num_points):
        t.view(torch.int32).fill_(-1)  # NaN bit pattern / -1
    plan.bind([torch.from_numpy(f).cuda() for f in frames])
    for _ in range(2):  # the second run sees the first one's leftovers
        vox, coors, num, vnum = plan.run()
    counts = vnum.cpu().tolist()
    for k, fr in enumerate(frames):
        ev, ec, en = oracle.hard_voxelize(fr, vs, rg, P, V)
        assert counts[k] == len(en)
        assert_same_bits(num[k, :counts[k]].cpu().numpy(), en, f"frame {k} num")
        assert_same_bits(coors[k, :counts[k]].cpu().numpy(), ec, f"frame {k} coors")
        assert_same_bits(vox[k, :counts[k]].cpu().numpy(), ev, f"frame {k} voxels")
    assert int(num[0, :counts[0]].max()) == P
