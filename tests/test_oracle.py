"""CPU tests that pin the oracle (oracle/pcfe_oracle.c) before anything trusts it:

1. against the golden vectors the reference itself produced (tests/golden/make_golden.py),
   including the literals of the reference's own known-answer tests;
2. against oracle/_ref (the reference's unmodified .cpp files compiled in place) on fresh
   random inputs, when that build is present;
3. the restated glibc sinf/cosf against the host libm.
"""
import struct

import numpy as np
import pytest

from detmatch_b200 import synth
from oracle import oracle, ref, scatter, vfe_mean
from tests.helpers import assert_same_bits, golden, golden_names


def _bits(f):
    return struct.unpack("<I", struct.pack("<f", f))[0]


@pytest.mark.parametrize("name", golden_names("voxel_"))
def test_oracle_voxel_golden(name):
    g = golden(name)
    if name == "voxel_generator_kat":
        # tests/test_models/test_voxel_encoder/test_voxel_generator.py:6-22 literal
        for pts, vox in ((g["points64"], g["voxels64"]), (g["points64"].astype(np.float32), g["voxels32"])):
            v, c, n = oracle.hard_voxelize(pts, g["voxel_size"], g["range"], int(g["max_points"]), int(g["max_voxels"]))
            assert_same_bits(c, g["coors"], "coors")
            assert_same_bits(n, g["num"], "num")
            assert_same_bits(v, vox, "voxels")
        return
    v, c, n = oracle.hard_voxelize(g["points"], g["voxel_size"], g["range"], int(g["max_points"]), int(g["max_voxels"]))
    assert_same_bits(v, g["voxels"], "voxels")
    assert_same_bits(c, g["coors"], "coors")
    assert_same_bits(n, g["num"], "num")
    if "dyn_coors" in g:
        assert_same_bits(oracle.dynamic_voxelize(g["points"], g["voxel_size"], g["range"]), g["dyn_coors"], "dyn")


def test_oracle_dynamic_matches_hard_content():
    """tests/test_models/test_voxel_encoder/test_voxelize.py:49-59: points grouped by dynamic
    coors, in order, equal the hard-voxelize content."""
    g = golden("voxel_kitti_fixture")
    coors = oracle.dynamic_voxelize(g["points"], g["voxel_size"], g["range"])
    for i in range(g["coors"].shape[0]):
        idx = np.all(coors == g["coors"][i], axis=1)
        k = int(idx.sum())
        assert k > 0 and k == g["num"][i]
        assert np.array_equal(g["points"][idx], g["voxels"][i][:k])


@pytest.mark.parametrize("name", golden_names("pib_"))
def test_oracle_pib_golden(name):
    g = golden(name)
    for restated in (False, True):
        out = oracle.points_in_boxes_cpu(g["points"], g["boxes"], restated_trig=restated)
        assert_same_bits(out, g["expected_cpu"], f"{name} restated={restated}")
    if name == "pib_kat":
        assert_same_bits(oracle.points_in_boxes_gpu(g["gpu_points"], g["gpu_boxes"]), g["expected_gpu"], "gpu literal")
        assert_same_bits(oracle.points_in_boxes_batch(g["batch_points"], g["batch_boxes"]), g["expected_batch"], "batch literal")


def test_oracle_empty_inputs():
    """the CPU ops accept N=0 / T=0."""
    vs, rg = [0.5] * 3, [0, -40, -3, 70.4, 40, 1]
    v, c, n = oracle.hard_voxelize(np.zeros((0, 4), np.float32), vs, rg, 5, 10)
    assert v.shape == (0, 5, 4) and c.shape == (0, 3) and n.shape == (0,)
    assert oracle.dynamic_voxelize(np.zeros((0, 4), np.float32), vs, rg).shape == (0, 3)
    assert oracle.points_in_boxes_cpu(np.zeros((0, 3), np.float32), np.zeros((2, 7), np.float32)).shape == (2, 0)
    assert oracle.points_in_boxes_cpu(np.zeros((5, 3), np.float32), np.zeros((0, 7), np.float32)).shape == (0, 5)


needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("cfg_name,ci", [("C1", 1), ("C4", 4), ("C5", 5)])
def test_oracle_vs_reference_voxelize(cfg_name, ci):
    import torch
    cfg = synth.CONFIGS[cfg_name]
    n = 30000
    for k, pts in enumerate((synth.lidar_frame(n, cfg["c"], synth.seed_for(ci, 10), cfg["r_max"]),
                             synth.uniform_frame(n, cfg["c"], synth.seed_for(ci, 11), cfg["point_cloud_range"]))):
        mv = cfg["max_voxels"] // (8 if k == 0 else 20)
        rv, rc, rn = ref.voxelization(pts, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], mv)
        v, c, m = oracle.hard_voxelize(pts.numpy(), cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], mv)
        assert_same_bits(v, rv.numpy(), "voxels")
        assert_same_bits(c, rc.numpy(), "coors")
        assert_same_bits(m, rn.numpy(), "num")
        rd = ref.voxelization(pts, cfg["voxel_size"], cfg["point_cloud_range"], -1, -1)
        assert_same_bits(oracle.dynamic_voxelize(pts.numpy(), cfg["voxel_size"], cfg["point_cloud_range"]), rd.numpy(), "dyn")
        assert isinstance(rd, torch.Tensor)


@needs_ref
def test_oracle_vs_reference_pib():
    import torch
    c3 = synth.CONFIGS["C3"]
    pts = synth.lidar_frame(20000, 3, synth.seed_for(3, 20), c3["r_max"])
    bxs = synth.random_boxes(100, synth.seed_for(3, 21), c3["point_cloud_range"])
    bxs[:30, 0:2] = pts[:30, 0:2]
    bxs[:30, 2] = pts[:30, 2] - 0.4
    fp = synth.face_points(bxs, 5, per_box=32)
    pts = torch.cat([pts, fp])
    exp = ref.points_in_boxes_cpu(pts, bxs).numpy()
    assert exp.sum() > 200
    assert_same_bits(oracle.points_in_boxes_cpu(pts.numpy(), bxs.numpy()), exp, "host trig")
    assert_same_bits(oracle.points_in_boxes_cpu(pts.numpy(), bxs.numpy(), restated_trig=True), exp, "restated trig")


def test_restated_sincosf_matches_host_libm():
    """The device computes cosa/sina with this algorithm; it must equal the host libm the
    reference calls (points_in_boxes_cpu.cpp:20).  Exhaustive over [2^-14, 16) takes ~2 s; the
    rest of the float line is strided."""
    sweeps = [
        (0x00000000, _bits(2.0 ** -14), 997),
        (_bits(2.0 ** -14), _bits(16.0), 1),
        (_bits(16.0), _bits(120.0), 3),
        (_bits(120.0), 0x7F800010, 211),
        (0x80000000, 0x80000000 + _bits(16.0), 13),
        (0x80000000 + _bits(16.0), 0xFF800010, 223),
    ]
    for lo, hi, stride in sweeps:
        bad, first = oracle.sincosf_sweep(lo, hi, stride)
        assert bad == 0, f"{bad} mismatches in [{lo:#x},{hi:#x}), first at bits {first:#x}"


def test_grid_size_float32_rounding():
    # (75.2+75.2)/0.1 = 1503.9999 in float32 -> round -> 1504
    assert oracle.grid_size([0.1, 0.1, 0.15], [-75.2, -75.2, -2, 75.2, 75.2, 4]).tolist() == [1504, 1504, 40]
    assert oracle.grid_size([0.05, 0.05, 0.1], [0, -40, -3, 70.4, 40, 1]).tolist() == [1408, 1600, 40]
    assert oracle.grid_size([0.25, 0.25, 8], [-50, -50, -5, 50, 50, 3]).tolist() == [400, 400, 1]


@pytest.mark.parametrize("tag,nf,exact", [("c4", 5, False), ("c1", 4, True), ("c4_nf4", 4, True), ("rand", 5, False)])
def test_vfe_mean_restatement_vs_reference_golden(tag, nf, exact):
    """oracle/vfe_mean.py against outputs of the reference expression itself (voxel_encoder.py:41-44
    evaluated by torch on the CPU, tests/golden/make_golden.py: vfe_cases).  ATen's association is
    layout dependent: the slot-order restatement reproduces it bit for bit for 4 features and
    stays within the float32 summation bound 2 (P - 1) eps sum|x| / n + 2 eps |mean| otherwise."""
    g = golden("vfe_mean")
    src = "c4" if tag == "c4_nf4" else tag
    f, n, exp = g[src + "_features"], g[src + "_num_points"], g[tag + "_expected"]
    got = vfe_mean.hard_simple_vfe(f, n, nf)
    assert got.shape == exp.shape and got.dtype == exp.dtype
    if exact:
        assert_same_bits(got, exp, tag)
        return
    eps = 2.0 ** -24
    mag = np.abs(f[:, :, :nf]).sum(axis=1, dtype=np.float64) / n.reshape(-1, 1)
    bound = 2 * (f.shape[1] - 1) * eps * mag + 2 * eps * np.abs(exp.astype(np.float64))
    err = np.abs(got.astype(np.float64) - exp.astype(np.float64))
    assert np.all(err <= bound)
    assert np.mean(got.view(np.uint32) == exp.view(np.uint32)) > 0.85


def test_oracle_pcdet_pib_golden():
    """The OpenPCDet restatement against outputs of the reference's own roiaware_pool3d.cpp
    (compiled in place, tests/golden/make_golden.py: pcdet_cases): LiDAR-like points, and points
    within a few ulps of the MARGIN-expanded faces, NaN / Inf points and boxes."""
    g = golden("pcdet_pib")
    assert_same_bits(oracle.pcdet_points_in_boxes_cpu(g["points"], g["boxes"]), g["expected_cpu"], "pcdet random")
    assert_same_bits(oracle.pcdet_points_in_boxes_cpu(g["face_points"], g["face_boxes"]), g["face_expected_cpu"],
                     "pcdet faces")
    # the two margins really separate on the face set (the GPU flavour's 1e-5 is not the CPU's 1e-2)
    tight = oracle.pcdet_points_in_boxes_cpu(g["face_points"], g["face_boxes"], margin=1e-5)
    assert (tight != g["face_expected_cpu"]).sum() > 100 and np.all(tight <= g["face_expected_cpu"])


@pytest.mark.skipif(not ref.pcdet_available(), reason="oracle/_ref/pcdet_ref_cpu.so not built")
def test_oracle_pcdet_vs_reference():
    import torch
    c3 = synth.CONFIGS["C3"]
    pts = synth.lidar_frame(30000, 3, 4242, c3["r_max"])
    bxs = synth.random_boxes(200, 4243, c3["point_cloud_range"])
    bxs[:60, 0:3] = pts[:60]
    exp = ref.pcdet_points_in_boxes_cpu(pts, bxs).numpy()
    assert exp.sum() > 1000
    assert_same_bits(oracle.pcdet_points_in_boxes_cpu(pts.numpy(), bxs.numpy()), exp, "pcdet vs _ref")


def test_oracle_dynamic_scatter_vs_reference_test_construction():
    """oracle/scatter.py against the expected values the reference's own test builds
    (test_dynamic_scatter.py:55-84, evaluated by torch; tests/golden/make_golden.py: scatter_cases):
    voxel order and max bit for bit, mean within the test's tolerance (atol 1e-2, rtol 1e-5 -- in
    fact within 4 float32 ulps of the sum)."""
    g = golden("dynamic_scatter")
    for r in ("mean", "max"):
        out, vc, mp, cnt = scatter.forward(g["feats"], g["coors"], r)
        assert_same_bits(vc, g["voxel_coors"], "voxel_coors")
        if r == "max":
            assert_same_bits(out, g["max"], "max")
        else:
            assert np.allclose(out, g["mean"], atol=1e-2, rtol=1e-5)
            assert np.abs(out - g["mean"]).max() < 1e-5
        assert cnt.sum() == (mp >= 0).sum() and (mp >= 0).sum() == (g["coors"] >= 0).all(axis=1).sum()


@needs_ref
@pytest.mark.parametrize("seed", range(24))
def test_oracle_vs_reference_random_configs(seed):
    """Randomised grids, caps and feature counts: the C restatement against the reference's own
    compiled CPU ops (hard and dynamic voxelization, point in box), bit for bit."""
    import torch
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(0, 4000))
    c = int(rng.integers(3, 7))
    vs = [float(rng.choice([0.05, 0.1, 0.16, 0.2, 0.32, 0.5, 1.0])) for _ in range(3)]
    lo = [float(rng.uniform(-60, 0)) for _ in range(3)]
    ext = [vs[j] * int(rng.integers(1, 400 if j < 2 else 40)) for j in range(3)]
    rg = lo + [lo[j] + ext[j] for j in range(3)]
    p = int(rng.choice([1, 2, 3, 5, 8, 35]))
    v = int(rng.integers(1, 600))
    pts = rng.uniform(-1.0, 1.0, size=(n, c)).astype(np.float32)
    for j in range(3):  # most points inside, some outside, some exactly on the faces
        pts[:, j] = (lo[j] + (pts[:, j] * 0.6 + 0.5) * ext[j]).astype(np.float32)
    if n > 10:
        pts[0, 0], pts[1, 1], pts[2, 2] = rg[3], rg[1], rg[5]
        pts[3, int(rng.integers(0, 3))] = np.nan
        pts[4:8] = pts[8:9]  # duplicates
    tp = torch.from_numpy(pts)
    rv, rc, rn = ref.voxelization(tp, vs, rg, p, v)
    ov, oc, on = oracle.hard_voxelize(pts, vs, rg, p, v)
    assert_same_bits(oc, rc.numpy(), "coors")
    assert_same_bits(on, rn.numpy(), "num")
    assert_same_bits(ov, rv.numpy(), "voxels")
    assert_same_bits(oracle.dynamic_voxelize(pts, vs, rg), ref.voxelization(tp, vs, rg, -1, -1).numpy(), "dynamic")
    t = int(rng.integers(1, 40))
    bxs = synth.random_boxes(t, 2000 + seed, rg).numpy()
    if n:
        k = min(t, n)
        bxs[:k, 0:3] = pts[:k, 0:3]
    exp = ref.points_in_boxes_cpu(torch.from_numpy(pts[:, :3].copy()), torch.from_numpy(bxs)).numpy()
    assert_same_bits(oracle.points_in_boxes_cpu(pts[:, :3].copy(), bxs), exp, "pib")
    if ref.pcdet_available():
        exp = ref.pcdet_points_in_boxes_cpu(torch.from_numpy(pts[:, :3].copy()), torch.from_numpy(bxs)).numpy()
        assert_same_bits(oracle.pcdet_points_in_boxes_cpu(pts[:, :3].copy(), bxs), exp, "pcdet pib")


def test_roiaware_pool3d_oracle_matches_reference_test_literals():
    """tests/test_models/test_common_modules/test_roiaware_pool3d.py:9-40: the restatement reproduces the
    sums the reference test asserts (rtol 1e-3) and puts exactly the points into the RoIs that the
    reference's compiled points_in_boxes_cpu does."""
    g = golden("roiaware_pool3d_kat")
    n = int(g["out_size"])
    mp = int(g["max_pts_per_voxel"])
    pm, am, lists = oracle.roiaware_pool3d_forward(g["rois"], g["pts"], g["pts"], n, mp, 0)
    pa, _, lists2 = oracle.roiaware_pool3d_forward(g["rois"], g["pts"], g["pts"], n, mp, 1)
    assert pm.shape == (2, n, n, n, 3)
    assert np.isclose(pm.sum(), g["expected_sum_max"], rtol=1e-3)
    assert np.isclose(pa.sum(), g["expected_sum_avg"], rtol=1e-3)
    assert np.array_equal(lists, lists2)
    for b in range(2):
        members = sorted(int(i) for v in lists[b].reshape(-1, mp) for i in v[1:1 + v[0]])
        assert members == list(np.nonzero(g["inside"][b])[0])
    # backward: max routes every pooled gradient to its argmax point, avg spreads it evenly
    go = np.ones_like(pm)
    gi = oracle.roiaware_pool3d_backward(lists, am, go, g["pts"].shape[0], 0)
    assert gi.sum() == (am != -1).sum()
    gi = oracle.roiaware_pool3d_backward(lists, am, go, g["pts"].shape[0], 1)
    assert np.isclose(gi.sum(), 3 * (lists[..., 0] > 0).sum())
