"""tests/golden/make_golden.py -- generate the committed golden vectors.

Run in the build container only (needs /root/reference):
    python oracle/build_ref.py && python tests/golden/make_golden.py

Every expected output below is produced by the REFERENCE ITSELF:
  * the unmodified voxelization_cpu.cpp / points_in_boxes_cpu.cpp compiled in place
    (oracle/_ref, driven exactly like voxelize.py:41-58 / points_in_boxes.py:53-82), and
  * the numba VoxelGenerator (mmdet3d/core/voxel/voxel_generator.py:136-207), loaded by file
    path, as a second opinion where the reference's own tests use it,
or is a literal copied from the reference's test-suite (cited per case).  The .npz files are
what travels to the GPU box; /root/reference does not exist there.
"""
import importlib.util
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("DETMATCH_REFERENCE", "/root/reference")

from detmatch_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

KITTI_RANGE = [0, -40, -3, 70.4, 40, 1]


def _voxel_generator():
    path = os.path.join(REF, "mmdet3d/core/voxel/voxel_generator.py")
    spec = importlib.util.spec_from_file_location("ref_voxel_generator", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.VoxelGenerator


def _hard(points, vs, rg, p, v):
    vo, co, nu = ref.voxelization(torch.from_numpy(points), vs, rg, p, v)
    return vo.numpy(), co.numpy(), nu.numpy()


def _dyn(points, vs, rg):
    return ref.voxelization(torch.from_numpy(points), vs, rg, -1, -1).numpy()


def _save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path)} bytes")


def voxel_cases():
    VG = _voxel_generator()
    # --- tests/test_models/test_voxel_encoder/test_voxelize.py:15-59 (KITTI fixture) -----------
    pts = np.fromfile(os.path.join(REF, "tests/data/kitti/training/velodyne_reduced/000000.bin"),
                      dtype=np.float32).reshape(-1, 4)
    vs, p, v = [0.5, 0.5, 0.5], 1000, 20000
    vo, co, nu = _hard(pts, vs, KITTI_RANGE, p, v)
    g_vo, g_co, g_nu = VG(vs, KITTI_RANGE, p).generate(pts)  # numba second opinion
    assert np.array_equal(g_vo, vo) and np.array_equal(g_co, co) and np.array_equal(g_nu, nu)
    _save("voxel_kitti_fixture", points=pts, voxel_size=np.float64(vs), range=np.float64(KITTI_RANGE),
          max_points=p, max_voxels=v, voxels=vo, coors=co, num=nu, dyn_coors=_dyn(pts, vs, KITTI_RANGE))

    # --- tests/test_models/test_voxel_encoder/test_voxel_generator.py:6-22 (literal KAT) -------
    np.random.seed(0)
    pts64 = np.random.rand(1000, 4)
    exp_coors = np.array([[7, 81, 1], [6, 81, 0], [7, 80, 1], [6, 81, 1], [7, 81, 0], [6, 80, 1],
                          [7, 80, 0], [6, 80, 0]], dtype=np.int32)
    exp_num = np.array([120, 121, 127, 134, 115, 127, 125, 131], dtype=np.int32)
    vo, co, nu = _hard(pts64, vs, KITTI_RANGE, p, v)  # float64 input through the C++ op
    assert np.array_equal(co, exp_coors) and np.array_equal(nu, exp_num)
    vo32, co32, nu32 = _hard(pts64.astype(np.float32), vs, KITTI_RANGE, p, v)
    assert np.array_equal(co32, exp_coors) and np.array_equal(nu32, exp_num)
    _save("voxel_generator_kat", points64=pts64, voxel_size=np.float64(vs), range=np.float64(KITTI_RANGE),
          max_points=p, max_voxels=v, coors=exp_coors, num=exp_num, voxels64=vo, voxels32=vo32)

    # --- Waymo fixture, columns 0..4; caps hit and not hit ------------------------
    w = np.fromfile(os.path.join(REF, "tests/data/waymo/kitti_format/training/velodyne/0000000.bin"),
                    dtype=np.float32).reshape(-1, 6)[:, :5].copy()
    c4 = synth.CONFIGS["C4"]
    for tag, pp, vv in (("nocap", 5, 150000), ("cap", 2, 50)):
        vo, co, nu = _hard(w, c4["voxel_size"], c4["point_cloud_range"], pp, vv)
        _save(f"voxel_waymo_fixture_{tag}", points=w, voxel_size=np.float64(c4["voxel_size"]),
              range=np.float64(c4["point_cloud_range"]), max_points=pp, max_voxels=vv,
              voxels=vo, coors=co, num=nu, dyn_coors=_dyn(w, c4["voxel_size"], c4["point_cloud_range"]))

    # --- behaviours no reference test pins: caps, special values -------------
    rng = np.random.default_rng(1234)
    pts = np.concatenate([rng.uniform([0, -4, -3], [8, 4, 1], size=(3000, 3)),
                          rng.uniform(0, 1, size=(3000, 1))], axis=1).astype(np.float32)
    vs2 = [0.5, 0.5, 0.5]
    vo, co, nu = _hard(pts, vs2, KITTI_RANGE, 3, 150)
    _save("voxel_caps", points=pts, voxel_size=np.float64(vs2), range=np.float64(KITTI_RANGE),
          max_points=3, max_voxels=150, voxels=vo, coors=co, num=nu)

    nan, inf = float("nan"), float("inf")
    special = np.array([
        [0.0, -40.0, -3.0, 1], [70.4, 0, 0, 2], [70.39999, 0, 0, 3], [-0.0, 0, 0, 4], [-1e-9, 0, 0, 5],
        [1.0, 0, 1.0, 6], [1.0, 0, 0.99999, 7], [1.0, 0, 0, 8], [nan, 0, 0, 9], [1, nan, 0, 10],
        [1, 0, nan, 11], [inf, 0, 0, 12], [-inf, 0, 0, 13], [1e20, 0, 0, 14], [3e9, 0, 0, 15],
        [-3e9, 0, 0, 16], [1, 40.0, 0, 17], [1, 39.99999, 0, 18], [1, -40.00001, 0, 19],
        [35.2, 0.05, -1.0, 20], [35.2, 0.05, -1.0, 21], [35.2, 0.05, -1.0, 22], [1.0, 0, 0, 23],
        [0.05, 0.1, -2.9, 24], [0.15, 0.1, -2.9, 25], [0.1, 0.1, -2.9, 26], [1e-45, 0, 0, 27],
    ], dtype=np.float32)
    vs3 = [0.05, 0.05, 0.1]
    vo, co, nu = _hard(special, vs3, KITTI_RANGE, 2, 100)
    _save("voxel_special", points=special, voxel_size=np.float64(vs3), range=np.float64(KITTI_RANGE),
          max_points=2, max_voxels=100, voxels=vo, coors=co, num=nu, dyn_coors=_dyn(special, vs3, KITTI_RANGE))

    # --- the five configs at reduced size (LiDAR-like + uniform adversarial) --------------------
    for ci, name in ((1, "C1"), (4, "C4"), (5, "C5")):
        cfg = synth.CONFIGS[name]
        n = 6000
        for kind in ("lidar", "uniform"):
            if kind == "lidar":
                pts = synth.lidar_frame(n, cfg["c"], synth.seed_for(ci, 0), cfg["r_max"]).numpy()
            else:
                pts = synth.uniform_frame(n, cfg["c"], synth.seed_for(ci, 1), cfg["point_cloud_range"]).numpy()
            mv = max(cfg["max_voxels"] // 50, 64)  # scaled so the cap still bites on 6 k points
            vo, co, nu = _hard(pts, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], mv)
            _save(f"voxel_{name}_{kind}_small", seed=synth.seed_for(ci, 0 if kind == "lidar" else 1), n=n,
                  points=pts, voxel_size=np.float64(cfg["voxel_size"]), range=np.float64(cfg["point_cloud_range"]),
                  max_points=cfg["max_num_points"], max_voxels=mv, voxels=vo, coors=co, num=nu,
                  dyn_coors=_dyn(pts, cfg["voxel_size"], cfg["point_cloud_range"]))


def pib_cases():
    # --- tests/test_models/test_common_modules/test_roiaware_pool3d.py:75-94 (literal KAT) -----
    boxes = np.array([[1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 0.3], [-10.0, 23.0, 16.0, 10, 20, 20, 0.5]], dtype=np.float32)
    pts = np.array([[1, 2, 3.3], [1.2, 2.5, 3.0], [0.8, 2.1, 3.5], [1.6, 2.6, 3.6], [0.8, 1.2, 3.9],
                    [-9.2, 21.0, 18.2], [3.8, 7.9, 6.3], [4.7, 3.5, -12.2], [3.8, 7.6, -2], [-10.6, -12.9, -20],
                    [-16, -18, 9], [-21.3, -52, -5], [0, 0, 0], [6, 7, 8], [-2, -3, -4]], dtype=np.float32)
    expected_cpu = np.array([[1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                             [0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0]], dtype=np.int32)
    got = ref.points_in_boxes_cpu(torch.from_numpy(pts), torch.from_numpy(boxes)).numpy()
    assert np.array_equal(got, expected_cpu)
    # :43-72 (points_in_boxes_gpu literal) and :97-128 (points_in_boxes_batch literal)
    gpu_boxes = np.array([[[1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 0.3]], [[-10.0, 23.0, 16.0, 10, 20, 20, 0.5]]], dtype=np.float32)
    gpu_pts = np.array([[[1, 2, 3.3], [1.2, 2.5, 3.0], [0.8, 2.1, 3.5], [1.6, 2.6, 3.6], [0.8, 1.2, 3.9],
                         [-9.2, 21.0, 18.2], [3.8, 7.9, 6.3], [4.7, 3.5, -12.2]],
                        [[3.8, 7.6, -2], [-10.6, -12.9, -20], [-16, -18, 9], [-21.3, -52, -5], [0, 0, 0],
                         [6, 7, 8], [-2, -3, -4], [6, 4, 9]]], dtype=np.float32)
    expected_gpu = np.array([[0, 0, 0, 0, 0, -1, -1, -1], [-1] * 8], dtype=np.int32)
    expected_batch = np.array([[[1, 0]] * 5 + [[0, 1]] + [[0, 0]] * 9], dtype=np.int32)
    # tests/test_utils/test_box3d.py:1202-1209 pins DepthInstance3DBoxes.points_in_boxes -> zeros (5,2);
    # it goes through geometry code that is out of scope, so only the op-level literals are kept.
    _save("pib_kat", boxes=boxes, points=pts, expected_cpu=expected_cpu, gpu_boxes=gpu_boxes, gpu_points=gpu_pts,
          expected_gpu=expected_gpu, batch_boxes=boxes[None], batch_points=pts[None], expected_batch=expected_batch)

    # --- random boxes on LiDAR-like points (C3 distribution, reduced) ---------------------------
    c3 = synth.CONFIGS["C3"]
    pts = synth.lidar_frame(4000, 3, synth.seed_for(3, 0), c3["r_max"]).numpy()
    bxs = synth.random_boxes(64, synth.seed_for(3, 0) + 500, c3["point_cloud_range"])
    # make sure a few boxes actually contain points: centre some on existing points
    bxs[:16, 0:2] = torch.from_numpy(pts[:16, 0:2])
    bxs[:16, 2] = torch.from_numpy(pts[:16, 2]) - 0.5
    bxs = bxs.numpy()
    out = ref.points_in_boxes_cpu(torch.from_numpy(pts), torch.from_numpy(bxs)).numpy()
    assert out.sum() > 100
    _save("pib_random", boxes=bxs, points=pts, expected_cpu=out)

    # --- face / edge / corner points with axis-aligned and diagonal yaw -------------------------
    yaws = [0.0, math.pi / 2, -math.pi / 2, math.pi, -math.pi, math.pi / 4, -math.pi / 4, 0.3, 2.5, -1.1]
    bxs = synth.random_boxes(len(yaws) * 3, 77, [-20, -20, -3, 20, 20, 1]).numpy()
    bxs[:, 6] = np.float32(yaws * 3)
    bxs[-1, 3:6] = 0.0  # zero-size (padded) box contains nothing
    pts = synth.face_points(torch.from_numpy(bxs), 78, per_box=96).numpy()
    nan = float("nan")
    extra = np.array([[0, 0, nan], [nan, 0, 0], [0, nan, 0], [np.inf, 0, 0], [0, 0, np.inf], [1e30, 1e30, 0]], dtype=np.float32)
    pts = np.concatenate([pts, extra, bxs[:, :3] + np.float32([0, 0, 0.25])]).astype(np.float32)
    bx2 = np.concatenate([bxs, np.float32([[0, 0, -1, 2, 4, 1, nan], [0, 0, -1, nan, 4, 1, 0.1],
                                            [0, 0, -1, 2, 4, 2, -math.pi / 2], [0, 0, -1, 2, 4, 2, 100.0],
                                            [0, 0, -1, 2, 4, 2, 1000.0], [0, 0, -1, 2, 4, 2, -12345.678]])])
    out = ref.points_in_boxes_cpu(torch.from_numpy(pts), torch.from_numpy(bx2)).numpy()
    _save("pib_faces", boxes=bx2, points=pts, expected_cpu=out)
    print("pib_faces inside pairs:", int(out.sum()), "of", out.size)


def _hard_simple_vfe(features, num_points, num_features):
    # mmdet3d/models/voxel_encoders/voxel_encoder.py:41-44, verbatim (the module itself needs mmcv,
    # which this image does not have; its forward is this one expression)
    points_mean = features[:, :, :num_features].sum(dim=1, keepdim=False) / num_points.type_as(features).view(-1, 1)
    return points_mean.contiguous()


def vfe_cases():
    out = {}
    # voxelized LiDAR-like frames: the tensors HardSimpleVFE sees in the detectors
    for tag, cfg, nfeat in (("c4", "C4", 5), ("c1", "C1", 4), ("c4_nf4", "C4", 4)):
        c = synth.CONFIGS[cfg]
        pts = synth.lidar_frame(6000, c["c"], synth.seed_for(9, 0), c["r_max"]).numpy()
        vo, co, nu = _hard(pts, c["voxel_size"], c["point_cloud_range"], c["max_num_points"], c["max_voxels"])
        if tag != "c4_nf4":  # same tensors as "c4", only the first 4 features used
            out[tag + "_features"] = vo
            out[tag + "_num_points"] = nu
        out[tag + "_expected"] = _hard_simple_vfe(torch.from_numpy(vo), torch.from_numpy(nu), nfeat).numpy()
    # tests/test_models/test_voxel_encoder/test_voxel_encoders.py:26-33 (shape reduced 240000 -> 2000)
    gen = torch.Generator().manual_seed(5)
    f = torch.rand([2000, 10, 5], generator=gen)
    n = torch.randint(1, 10, [2000], generator=gen)
    out["rand_features"], out["rand_num_points"] = f.numpy(), n.numpy().astype(np.int32)
    out["rand_expected"] = _hard_simple_vfe(f, n, 5).numpy()
    _save("vfe_mean", **out)


def pcdet_cases():
    """OpenPCDet points_in_boxes_cpu (roiaware_pool3d.cpp:121-168) from the reference's own file."""
    assert ref.pcdet_available(), "run oracle/build_ref.py first"
    c3 = synth.CONFIGS["C3"]
    pts = synth.lidar_frame(4000, 3, synth.seed_for(3, 1), c3["r_max"]).numpy()
    bxs = synth.random_boxes(64, synth.seed_for(3, 1) + 500, c3["point_cloud_range"]).numpy()
    bxs[:16, 0:3] = pts[:16]  # centred on existing points (z is the centre here)
    out = ref.pcdet_points_in_boxes_cpu(torch.from_numpy(pts), torch.from_numpy(bxs)).numpy()
    assert out.sum() > 100
    # points within a few ulps of the MARGIN-expanded faces (both margins), axis-aligned and odd headings
    rng = np.random.default_rng(91)
    yaws = [0.0, math.pi / 2, -math.pi / 2, math.pi, -math.pi, math.pi / 4, 0.3, 2.5, -1.1, 100.0, -12345.678, 1e-5]
    fb = synth.random_boxes(len(yaws) * 2, 92, [-20, -20, -3, 20, 20, 1]).numpy()
    fb[:, 6] = np.float32(yaws * 2)
    fb[-1, 3:6] = 0.0
    fp = []
    for b in fb.astype(np.float64):
        for margin in (1e-2, 1e-5):
            for _ in range(24):
                lx = (b[3] / 2 + margin) * rng.choice([-1, 1]) * (1 + rng.integers(-3, 4) * 2.0 ** -23)
                ly = (b[4] / 2 + margin) * rng.choice([-1, 1]) * (1 + rng.integers(-3, 4) * 2.0 ** -23)
                if rng.random() < 0.5:
                    lx *= rng.random()
                else:
                    ly *= rng.random()
                dz = (b[5] / 2) * rng.choice([-1, 1]) * (1 + rng.integers(-2, 3) * 2.0 ** -23) * (1 if rng.random() < 0.3 else rng.random())
                c, s_ = math.cos(b[6]), math.sin(b[6])
                fp.append([b[0] + lx * c - ly * s_, b[1] + lx * s_ + ly * c, b[2] + dz])
    nan = float("nan")
    fp = np.float32(fp + [[0, 0, nan], [nan, 0, 0], [np.inf, 0, 0], [0, 0, np.inf], [1e30, 1e30, 0]])
    fb2 = np.concatenate([fb, np.float32([[0, 0, 0, 2, 4, 1, nan], [0, 0, 0, nan, 4, 1, 0.1], [0, 0, 0, 2, 4, nan, 0.1],
                                          [0, 0, 0, -1, 4, 2, 0.2], [0, 0, 0, np.inf, np.inf, np.inf, 0.0]])])
    fout = ref.pcdet_points_in_boxes_cpu(torch.from_numpy(fp), torch.from_numpy(fb2)).numpy()
    print("pcdet faces inside pairs:", int(fout.sum()), "of", fout.size)
    _save("pcdet_pib", boxes=bxs, points=pts, expected_cpu=out, face_boxes=fb2, face_points=fp, face_expected_cpu=fout)


def scatter_cases():
    """DynamicScatter: the reference TEST's own expected values
    (tests/test_models/test_voxel_encoder/test_dynamic_scatter.py:13-16,55-66, N reduced 200000 -> 6000),
    evaluated by torch on the CPU."""
    gen = torch.Generator().manual_seed(17)
    feats = torch.rand(size=(6000, 3), dtype=torch.float32, generator=gen) * 100 - 50
    coors = torch.randint(low=-1, high=20, size=(6000, 3), dtype=torch.int32, generator=gen)
    ref_voxel_coors = coors.unique(dim=0, sorted=True)
    ref_voxel_coors = ref_voxel_coors[ref_voxel_coors.min(dim=-1).values >= 0]
    mean, mx = [], []
    for ref_voxel_coor in ref_voxel_coors:
        voxel_mask = (coors == ref_voxel_coor).all(dim=-1)
        mean.append(feats[voxel_mask].mean(dim=0))
        mx.append(feats[voxel_mask].max(dim=0).values)
    _save("dynamic_scatter", feats=feats.numpy(), coors=coors.numpy(), voxel_coors=ref_voxel_coors.numpy(),
          mean=torch.stack(mean).numpy(), max=torch.stack(mx).numpy())


def roiaware_cases():
    """RoIAwarePool3d: the reference TEST's literals and expectations
    (tests/test_models/test_common_modules/test_roiaware_pool3d.py:9-40).  The reference has no CPU
    implementation of the op (and this container has no GPU), so the golden holds the inputs, the two
    sums the reference test asserts (rtol 1e-3) and the membership of every point derived with the
    reference's compiled points_in_boxes_cpu -- the same inside-test text the pooling kernel carries
    (roiaware_pool3d_kernel.cu:26-42)."""
    rois = np.float32([[1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 0.3], [-10.0, 23.0, 16.0, 10, 20, 20, 0.5]])
    pts = np.float32([[1, 2, 3.3], [1.2, 2.5, 3.0], [0.8, 2.1, 3.5], [1.6, 2.6, 3.6], [0.8, 1.2, 3.9], [-9.2, 21.0, 18.2],
                      [3.8, 7.9, 6.3], [4.7, 3.5, -12.2], [3.8, 7.6, -2], [-10.6, -12.9, -20], [-16, -18, 9],
                      [-21.3, -52, -5], [0, 0, 0], [6, 7, 8], [-2, -3, -4]])
    inside = ref.points_in_boxes_cpu(torch.from_numpy(pts), torch.from_numpy(rois)).numpy()
    print("roiaware KAT: points inside per RoI:", inside.sum(axis=1))
    _save("roiaware_pool3d_kat", rois=rois, pts=pts, inside=inside, out_size=np.int32(4), max_pts_per_voxel=np.int32(128),
          expected_sum_max=np.float32(51.100), expected_sum_avg=np.float32(49.750))


if __name__ == "__main__":
    assert ref.available(), "run oracle/build_ref.py first"
    if len(sys.argv) > 1 and sys.argv[1] == "roiaware":
        roiaware_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "vfe":
        vfe_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "scatter":
        scatter_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pcdet":
        pcdet_cases()
        sys.exit(0)
    voxel_cases()
    pib_cases()
    vfe_cases()
    pcdet_cases()
    scatter_cases()
    roiaware_cases()
