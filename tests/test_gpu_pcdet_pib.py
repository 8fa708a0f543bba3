"""OpenPCDet points_in_boxes_{gpu,cpu} (thirdparty/Spconv-OpenPCDet/pcdet/ops/roiaware_pool3d) on the
GPU, bit for bit against the oracle's restatement of roiaware_pool3d.cpp:121-168 and the golden
outputs of the reference's own compiled file."""
import numpy as np
import pytest
import torch

from detmatch_b200 import synth
from detmatch_b200.ops.pcdet_roiaware_pool3d import points_in_boxes_cpu, points_in_boxes_gpu, roiaware_pool3d_cuda
from oracle import oracle
from tests.helpers import assert_same_bits, golden

pytestmark = pytest.mark.gpu


def test_golden_cpu_layout():
    g = golden("pcdet_pib")
    for pk, bk, ek in (("points", "boxes", "expected_cpu"), ("face_points", "face_boxes", "face_expected_cpu")):
        got = points_in_boxes_cpu(torch.from_numpy(g[pk]).cuda(), torch.from_numpy(g[bk]).cuda())
        assert_same_bits(got.cpu().numpy(), g[ek], pk)
        # numpy in, numpy out (roiaware_pool3d_utils.py:19-25)
        got_np = points_in_boxes_cpu(g[pk], g[bk])
        assert isinstance(got_np, np.ndarray)
        assert_same_bits(got_np, g[ek], pk + " numpy")


def test_golden_faces_gpu_margin():
    """points_in_boxes_gpu uses the CUDA kernel's MARGIN = 1e-5 (roiaware_pool3d_kernel.cu:27): first
    hit per point over the face set, where the two margins disagree on hundreds of pairs."""
    g = golden("pcdet_pib")
    pts, bxs = g["face_points"], g["face_boxes"]
    got = points_in_boxes_gpu(torch.from_numpy(pts[None]).cuda(), torch.from_numpy(bxs[None]).cuda())
    assert_same_bits(got.cpu().numpy(), oracle.pcdet_points_in_boxes_gpu(pts[None], bxs[None]), "faces first hit")


@pytest.mark.parametrize("b,m,t", [(1, 1, 1), (2, 1000, 7), (3, 20000, 200), (2, 5000, 513), (1, 300, 1100)])
def test_random_vs_oracle(b, m, t):
    c3 = synth.CONFIGS["C3"]
    pts = torch.stack([synth.lidar_frame(m, 3, 900 + k, c3["r_max"]) for k in range(b)])
    bxs = torch.stack([synth.random_boxes(t, 950 + k, c3["point_cloud_range"]) for k in range(b)])
    k = min(t, m, 40)
    bxs[:, :k, 0:3] = pts[:, :k]
    got = points_in_boxes_gpu(pts.cuda(), bxs.cuda()).cpu().numpy()
    exp = oracle.pcdet_points_in_boxes_gpu(pts.numpy(), bxs.numpy())
    assert_same_bits(got, exp, "gpu layout")
    if m > 100:
        assert (exp >= 0).sum() > 0
    got = points_in_boxes_cpu(pts[0].cuda(), bxs[0].cuda()).cpu().numpy()
    assert_same_bits(got, oracle.pcdet_points_in_boxes_cpu(pts[0].numpy(), bxs[0].numpy()), "cpu layout")


def test_point_head_usage():
    """point_head_template.py:82-89: (1, N, 3) points against (1, T, 7) gt boxes and the enlarged
    boxes; the foreground flags follow box_idxs >= 0."""
    c3 = synth.CONFIGS["C3"]
    pts = synth.lidar_frame(16384, 3, 77, c3["r_max"])
    gt = synth.random_boxes(40, 78, c3["point_cloud_range"])
    gt[:, 0:3] = pts[:40]
    ext = gt.clone()
    ext[:, 3:6] += 0.4  # box_utils.enlarge_box3d(extra_width=0.2) on each side
    idx = points_in_boxes_gpu(pts[None].cuda(), gt[None].cuda()).long().squeeze(0)
    ext_idx = points_in_boxes_gpu(pts[None].cuda(), ext[None].cuda()).long().squeeze(0)
    assert_same_bits(idx.int().cpu().numpy(), oracle.pcdet_points_in_boxes_gpu(pts[None].numpy(), gt[None].numpy())[0], "gt")
    assert_same_bits(ext_idx.int().cpu().numpy(), oracle.pcdet_points_in_boxes_gpu(pts[None].numpy(), ext[None].numpy())[0], "ext")
    fg, ext_fg = idx >= 0, ext_idx >= 0
    assert fg.sum() >= 40 and bool((ext_fg | ~fg).all())  # a point inside a box is inside its enlarged box


def test_empty_and_dropin_signature():
    out = torch.full((2, 5), 9, dtype=torch.int32).cuda()
    roiaware_pool3d_cuda.points_in_boxes_gpu(torch.empty((2, 0, 7)).cuda(), torch.zeros((2, 5, 3)).cuda(), out)
    assert bool((out == -1).all())  # no boxes: everything is background
    assert points_in_boxes_cpu(torch.empty((0, 3)).cuda(), torch.zeros((4, 7)).cuda()).shape == (4, 0)
    assert points_in_boxes_gpu(torch.empty((1, 0, 3)).cuda(), torch.zeros((1, 4, 7)).cuda()).shape == (1, 0)
    with pytest.raises(RuntimeError):
        roiaware_pool3d_cuda.points_in_boxes_gpu(torch.zeros((1, 2, 7)), torch.zeros((1, 5, 3)), torch.zeros((1, 5), dtype=torch.int32))
