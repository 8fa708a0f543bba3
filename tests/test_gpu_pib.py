"""GPU parity tests for points_in_boxes_{gpu,batch,cpu}: CUDA path (through the C ABI) vs the
CPU oracle (host libm trig, like the reference) and the golden vectors.  Bar: bit-exact."""
import math
import struct

import numpy as np
import pytest
import torch

from detmatch_b200 import _cabi, synth
from detmatch_b200.ops import points_in_boxes_batch, points_in_boxes_cpu, points_in_boxes_gpu
from oracle import oracle
from tests.helpers import assert_same_bits, golden, golden_names

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["grid", "brute"])
def pib_mode(request):
    """First-hit assignment (points_in_boxes_gpu) through the per-frame box grid (default) and by
    brute force over all boxes; the other two ops ignore the knob."""
    _cabi.debug_set("pib_grid", 1 if request.param == "grid" else 0)
    yield request.param
    _cabi.debug_set("pib_grid", 1)


def _all_three(pts, bxs, what):
    """pts (B,M,3), bxs (B,T,7) numpy -> checks the three ops against the oracle."""
    tp, tb = torch.from_numpy(pts).cuda(), torch.from_numpy(bxs).cuda()
    exp_batch = oracle.points_in_boxes_batch(pts, bxs)
    got = points_in_boxes_batch(tp, tb)
    assert got.dtype == torch.int32 and tuple(got.shape) == exp_batch.shape
    assert_same_bits(got.cpu().numpy(), exp_batch, what + " batch")
    assert_same_bits(points_in_boxes_gpu(tp, tb).cpu().numpy(), oracle.points_in_boxes_gpu(pts, bxs), what + " gpu")
    for b in range(pts.shape[0]):
        assert_same_bits(points_in_boxes_cpu(tp[b], tb[b]).cpu().numpy(), oracle.points_in_boxes_cpu(pts[b], bxs[b]),
                         what + f" cpu-layout b={b}")
    return int(exp_batch.sum())


@pytest.mark.parametrize("name", golden_names("pib_"))
def test_golden(name):
    g = golden(name)
    out = points_in_boxes_cpu(torch.from_numpy(g["points"]).cuda(), torch.from_numpy(g["boxes"]).cuda())
    assert_same_bits(out.cpu().numpy(), g["expected_cpu"], name)
    # same through CPU tensors (uploaded, computed on the GPU, returned on the CPU)
    out = points_in_boxes_cpu(torch.from_numpy(g["points"]), torch.from_numpy(g["boxes"]))
    assert not out.is_cuda
    assert_same_bits(out.numpy(), g["expected_cpu"], name + " via cpu tensors")
    b = points_in_boxes_batch(torch.from_numpy(g["points"])[None].cuda(), torch.from_numpy(g["boxes"])[None].cuda())
    assert_same_bits(b[0].t().contiguous().cpu().numpy(), g["expected_cpu"], name + " batch^T")


def test_reference_literals():
    """tests/test_models/test_common_modules/test_roiaware_pool3d.py:43-128."""
    g = golden("pib_kat")
    got = points_in_boxes_gpu(points=torch.from_numpy(g["gpu_points"]).cuda(), boxes=torch.from_numpy(g["gpu_boxes"]).cuda())
    assert got.shape == torch.Size([2, 8])
    assert_same_bits(got.cpu().numpy(), g["expected_gpu"], "points_in_boxes_gpu literal")
    got = points_in_boxes_batch(points=torch.from_numpy(g["batch_points"]).cuda(), boxes=torch.from_numpy(g["batch_boxes"]).cuda())
    assert got.shape == torch.Size([1, 15, 2])
    assert_same_bits(got.cpu().numpy(), g["expected_batch"], "points_in_boxes_batch literal")
    got = points_in_boxes_cpu(points=torch.from_numpy(g["points"]).cuda(), boxes=torch.from_numpy(g["boxes"]).cuda())
    assert got.shape == torch.Size([2, 15])
    assert_same_bits(got.cpu().numpy(), g["expected_cpu"], "points_in_boxes_cpu literal")


@pytest.mark.parametrize("b,m,t", [(1, 1, 1), (2, 1000, 4), (3, 777, 5), (2, 5000, 200), (1, 3000, 199), (4, 64, 36),
                                   (1, 2000, 1024), (1, 500, 1500), (1, 300, 2052), (2, 4099, 128)])
def test_shapes_vs_oracle(b, m, t):
    c3 = synth.CONFIGS["C3"]
    pts = np.stack([synth.lidar_frame(m, 3, 100 * b + k + m, c3["r_max"]).numpy() for k in range(b)])
    bxs = np.stack([synth.random_boxes(t, 7 * t + k, c3["point_cloud_range"]).numpy() for k in range(b)])
    k = min(t, m)
    bxs[:, :k, 0:2] = pts[:, :k, 0:2]     # make boxes hit something
    bxs[:, :k, 2] = pts[:, :k, 2] - 0.5
    inside = _all_three(pts, bxs, f"B={b} M={m} T={t}")
    assert inside > 0


def test_faces_edges_corners_and_special_values():
    g = golden("pib_faces")
    pts, bxs = g["points"][None], g["boxes"][None]
    _all_three(pts, bxs, "faces")


def test_reject_margin_adversarial():
    """The kernels skip the exact test when |x - cx| or |y - cy| exceeds 1.0001 * sqrt(hl^2 + hw^2).
    Points are placed on the world axes through each box centre at distances swept finely around
    that radius (where a box corner can lie), for yaws that put a corner on or next to an axis, for
    tiny / huge / degenerate / non-finite boxes: every flag must still equal the CPU op's."""
    rng = np.random.default_rng(20240607)
    T = 96
    bxs = np.zeros((T, 7), dtype=np.float32)
    bxs[:, 0:2] = rng.uniform(-30, 30, (T, 2))
    bxs[:, 2] = -1.0
    bxs[:, 3] = rng.uniform(0.3, 3.0, T)      # w
    bxs[:, 4] = rng.uniform(0.3, 6.0, T)      # l
    bxs[:, 5] = 2.0
    # yaw such that the corner (+hl, +hw) points along +x, +y, -x, -y (plus a little noise)
    corner = np.arctan2(bxs[:, 3], bxs[:, 4])
    bxs[:, 6] = (-np.pi / 2 - corner + (np.arange(T) % 4) * (np.pi / 2) + rng.normal(0, 1e-4, T)).astype(np.float32)
    bxs[90, 3:5] = [1e-3, 2e-3]
    bxs[91, 3:5] = [900.0, 1200.0]
    bxs[92, 3:5] = [0.0, 2.0]
    bxs[93, 3:5] = [-1.0, 2.0]
    bxs[94, 3:5] = [np.nan, 2.0]
    bxs[95, 3:5] = [np.inf, 2.0]
    rho = np.sqrt((bxs[:, 3].astype(np.float64) / 2) ** 2 + (bxs[:, 4].astype(np.float64) / 2) ** 2)
    rho = np.nan_to_num(rho, nan=1.0, posinf=1e6)
    f = np.concatenate([1.0 + np.linspace(-3e-4, 3e-4, 25), [0.5, 0.999, 1.001, 2.0]])
    pts = []
    for k in range(T):
        for sx, sy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            d = (rho[k] * f).astype(np.float64)
            p = np.zeros((len(f), 3))
            p[:, 0] = bxs[k, 0] + sx * d
            p[:, 1] = bxs[k, 1] + sy * d
            p[:, 2] = 0.0
            pts.append(p)
    pts = np.concatenate(pts).astype(np.float32)
    inside = _all_three(pts[None], bxs[None], "reject margin")
    assert inside > 0


def test_c3_frame_vs_oracle():
    """BASELINE config C3: 16 frames x 120k points x 200 boxes, every frame against the oracle, plus
    the cross-op consistency of the three layouts."""
    c3 = synth.CONFIGS["C3"]
    B = 16
    pts = torch.stack([synth.lidar_frame(c3["n"], 3, synth.seed_for(3, k), c3["r_max"]) for k in range(B)])
    bxs = torch.stack([synth.random_boxes(c3["boxes"], synth.seed_for(3, k) + 500, c3["point_cloud_range"]) for k in range(B)])
    bxs[:, :50, 0:2] = pts[:, :50, 0:2]
    bxs[:, :50, 2] = pts[:, :50, 2] - 0.5
    tp, tb = pts.cuda(), bxs.cuda()
    got = points_in_boxes_batch(tp, tb)
    assert got.shape == (B, c3["n"], c3["boxes"])
    for k in range(B):
        exp = oracle.points_in_boxes_cpu(pts[k].numpy(), bxs[k].numpy())
        assert exp.sum() > 1000
        assert_same_bits(got[k].t().contiguous().cpu().numpy(), exp, f"C3 frame {k}")
    # consistency of the three layouts on all frames
    first = points_in_boxes_gpu(tp, tb)
    any_hit = got.max(dim=2).values
    arg = got.argmax(dim=2).int()
    assert torch.equal(first, torch.where(any_hit > 0, arg, torch.full_like(arg, -1)))
    assert torch.equal(points_in_boxes_cpu(tp[3], tb[3]), got[3].t())


def test_empty():
    z = torch.zeros
    assert points_in_boxes_batch(z((2, 0, 3)).cuda(), z((2, 5, 7)).cuda()).shape == (2, 0, 5)
    assert points_in_boxes_batch(z((2, 9, 3)).cuda(), z((2, 0, 7)).cuda()).shape == (2, 9, 0)
    assert (points_in_boxes_gpu(z((2, 9, 3)).cuda(), z((2, 0, 7)).cuda()) == -1).all()
    assert points_in_boxes_cpu(z((0, 3)).cuda(), z((4, 7)).cuda()).shape == (4, 0)
    assert points_in_boxes_cpu(z((6, 3)).cuda(), z((0, 7)).cuda()).shape == (0, 6)


def test_asserts_like_reference():
    with pytest.raises(AssertionError):
        points_in_boxes_gpu(torch.zeros((2, 4, 3)).cuda(), torch.zeros((1, 4, 7)).cuda())
    with pytest.raises(AssertionError):
        points_in_boxes_batch(torch.zeros((1, 4, 3)).cuda(), torch.zeros((1, 4, 6)).cuda())
    with pytest.raises(AssertionError):
        points_in_boxes_cpu(torch.zeros((4, 2)).cuda(), torch.zeros((4, 7)).cuda())


def test_current_device_untouched():
    """points_in_boxes.py:32-44 changes the current device as a side effect; we must not."""
    before = torch.cuda.current_device()
    points_in_boxes_gpu(torch.zeros((1, 4, 3)).cuda(), torch.zeros((1, 2, 7)).cuda())
    assert torch.cuda.current_device() == before


def _bits(f):
    return struct.unpack("<I", struct.pack("<f", f))[0]


def test_device_trig_equals_host_libm():
    """cosa/sina come from a device restatement of glibc's sinf/cosf: compare with the HOST libm of
    this machine (what the reference would call here) on 6 M angles incl. every box-yaw-sized one."""
    lo, hi = _bits(2.0 ** -14), _bits(16.0)
    sweeps = [np.arange(lo, hi, 61, dtype=np.int64), np.arange(_bits(16.0), _bits(120.0), 17, dtype=np.int64),
              np.arange(_bits(120.0), 0x7F800000, 4099, dtype=np.int64), np.arange(0, lo, 8191, dtype=np.int64)]
    u = np.concatenate(sweeps).astype(np.uint32)
    u = np.concatenate([u, u | np.uint32(0x80000000), np.array([0x7F800000, 0xFF800000, 0x7FC00000], dtype=np.uint32)])
    x = torch.from_numpy(u.view(np.float32).copy()).cuda()
    s, c = torch.empty_like(x), torch.empty_like(x)
    rc = _cabi.lib().pcfe_debug_sincosf(x.data_ptr(), x.numel(), s.data_ptr(), c.data_ptr(), 0,
                                        torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    xs = x.cpu().numpy()
    hs, hc = oracle.sincosf_array(xs, host=True)   # this machine's libm sinf/cosf
    gs, gc = s.cpu().numpy(), c.cpu().numpy()
    nan = np.isnan(hs) & np.isnan(gs) & np.isnan(hc) & np.isnan(gc)
    bad = ((hs.view(np.uint32) != gs.view(np.uint32)) | (hc.view(np.uint32) != gc.view(np.uint32))) & ~nan
    assert bad.sum() == 0, f"{int(bad.sum())} device sin/cos values differ from the host libm, first x bits {u[np.argmax(bad)]:#x}"


def _first_hit(pts, bxs, what):
    got = points_in_boxes_gpu(torch.from_numpy(pts).cuda(), torch.from_numpy(bxs).cuda()).cpu().numpy()
    assert_same_bits(got, oracle.points_in_boxes_gpu(pts, bxs), what)
    return got


def test_first_hit_grid_adversarial():
    """Frames the box grid must get right or hand to the brute-force kernel: heavily overlapping
    boxes (lowest index wins), one huge box (spans more than 8 cells), non-finite boxes, all boxes
    identical, a single box, tiny boxes far from the origin, points far outside the boxes' extent,
    and more boxes than the grid lists (t > 4096)."""
    rng = np.random.default_rng(7)
    c3 = synth.CONFIGS["C3"]
    m = 6000
    pts = np.stack([synth.lidar_frame(m, 3, 4300 + k, c3["r_max"]).numpy() for k in range(2)])
    base = np.stack([synth.random_boxes(60, 4310 + k, c3["point_cloud_range"]).numpy() for k in range(2)])
    base[:, :40, 0:3] = pts[:, :40]
    base[:, :40, 2] -= 0.5

    # overlapping stacks: boxes 0..39 centred on points, boxes 40..59 copies of earlier ones (never win)
    ov = base.copy()
    ov[:, 40:60] = ov[:, 0:20]
    got = _first_hit(pts, ov, "overlap")
    assert (got >= 40).sum() == 0 and (got >= 0).sum() > 40

    huge = base.copy()
    huge[0, 5, 3:6] = [150.0, 160.0, 10.0]      # frame 0 falls back, frame 1 keeps its grid
    huge[0, 5, 0:3] = [0.0, 0.0, -5.0]
    got = _first_hit(pts, huge, "huge box")
    assert (got[0] == 5).sum() > 1000

    bad = base.copy()
    bad[1, 7, 0] = np.nan
    bad[1, 8, 3] = np.inf
    bad[0, 9, 6] = np.nan                        # NaN yaw: finite centre and radius, nothing inside
    _first_hit(pts, bad, "non-finite boxes")

    same = np.repeat(base[:, :1], 30, axis=1)
    got = _first_hit(pts, same, "identical boxes")
    assert set(np.unique(got)) <= {-1, 0}

    _first_hit(pts, base[:, :1], "one box")

    tiny = base.copy()
    tiny[:, :, 3:6] = 0.02
    tiny[:, :, 0:2] += 1.0e5                      # ulp(1e5) = 0.0078 > box size
    far_pts = pts.copy()
    far_pts[:, :2000, 0:2] += np.float32(1.0e5)
    far_pts[:, :40, 0:3] = tiny[:, :40, 0:3] + np.float32([0, 0, 0.01])
    _first_hit(far_pts, tiny, "tiny boxes far from the origin")

    outside = pts.copy()
    outside[:, ::3, 0] *= 50.0
    outside[:, 1::7, 1] = -np.inf
    outside[:, 2::11, 2] = np.nan
    _first_hit(outside, base, "points outside the grid")

    stretched = np.concatenate([np.repeat(base[:, :1], 140, axis=1), base], axis=1)  # 141 boxes in one cell
    stretched[0, 150, 0] = 1.0e6                                                      # an outlier stretches the grid
    _first_hit(pts, stretched, "long cell lists / stretched grid")

    many = np.stack([synth.random_boxes(4500, 4320 + k, c3["point_cloud_range"]).numpy() for k in range(2)])
    _first_hit(pts[:, :1500], many, "t > 4096")


@pytest.mark.parametrize("t", [1, 33, 200, 513, 1100])
def test_first_hit_random_vs_oracle(t):
    c3 = synth.CONFIGS["C3"]
    pts = np.stack([synth.lidar_frame(20000, 3, 4400 + k, c3["r_max"]).numpy() for k in range(3)])
    bxs = np.stack([synth.random_boxes(t, 4410 + k, c3["point_cloud_range"]).numpy() for k in range(3)])
    k = min(t, 64)
    bxs[:, :k, 0:3] = pts[:, :k]
    bxs[:, :k, 2] -= 0.4
    got = _first_hit(pts, bxs, f"t={t}")
    assert (got >= 0).sum() >= k


def test_depth_boxes_points_in_boxes_reference_literal():
    """tests/test_utils/test_box3d.py:1157-1209: DepthInstance3DBoxes(th_boxes, box_dim=6, with_yaw=False)
    rotated by -0.04599790655000615 (the tensor the reference test asserts at :1161-1173), queried with
    the test's five points: no point lies in a box."""
    from detmatch_b200.ops.roiaware_pool3d import depth_boxes_points_in_boxes
    boxes = torch.tensor([[0.64884546, 0.78390356, 0.10563634, 1.50373348, 0.23795205, 0.27956772, 0],
                          [1.45139421, 0.43169443, 0.93829232, 0.11967964, 0.93380373, 1.89191735, 0]])
    points = torch.tensor([[0.6762, 1.2559, -1.4658], [0.8784, 4.7814, -1.3857], [-0.2517, 6.7053, -0.9697],
                           [0.5520, 0.6533, -0.5265], [-0.5358, 4.5870, -1.4741]])
    got = depth_boxes_points_in_boxes(boxes, points.cuda())
    assert got.dtype == torch.int32 and got.device.type == "cuda"
    assert torch.all(got == torch.zeros((5, 2), dtype=torch.int32, device="cuda"))
    assert depth_boxes_points_in_boxes(boxes, points.cuda()[None]).shape == (5, 2)


def test_box_structure_adapters_vs_oracle():
    """depth_box3d.py:251-277 / lidar_box3d.py:258-270 flows on random boxes with points inside: the
    axis swap + DEPTH -> LIDAR conversion happen in torch, the masks equal points_in_boxes_cpu of the
    oracle on the converted inputs."""
    from detmatch_b200.ops.roiaware_pool3d import (depth_boxes_points_in_boxes, depth_boxes_to_lidar,
                                                   lidar_boxes_points_in_boxes)
    g = torch.Generator().manual_seed(31337)
    T, M = 40, 20000
    boxes = torch.cat([torch.rand((T, 3), generator=g) * 8 - 4, torch.rand((T, 3), generator=g) * 2 + 0.3,
                       (torch.rand((T, 1), generator=g) * 2 - 1) * 3.14159], dim=1)
    points = torch.rand((M, 3), generator=g) * 10 - 5
    points[:T] = boxes[:, :3] + torch.tensor([0.0, 0.0, 0.1])  # depth boxes are bottom-centred like LiDAR ones
    got = depth_boxes_points_in_boxes(boxes, points.cuda())
    pl = points[:, [1, 0, 2]].clone()
    pl[:, 1] *= -1
    exp = oracle.points_in_boxes_cpu(pl.numpy(), depth_boxes_to_lidar(boxes).numpy())
    assert exp.sum() > T
    assert_same_bits(got.t().contiguous().cpu().numpy(), exp, "depth flow")
    first = lidar_boxes_points_in_boxes(boxes, points.cuda())
    e = oracle.points_in_boxes_cpu(points.numpy(), boxes.numpy())  # (T, M)
    want = np.where(e.any(axis=0), e.argmax(axis=0), -1).astype(np.int32)
    assert_same_bits(first.cpu().numpy(), want, "lidar flow")
