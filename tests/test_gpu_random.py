"""Randomised cross-op parity on the GPU: every seed draws a grid, caps, a feature count, points with
NaN / face / duplicate rows and a set of boxes, and checks dynamic + hard voxelization (plain, packed,
mean) and the point-in-box ops of both conventions against the oracle, bit for bit."""
import numpy as np
import pytest
import torch

from detmatch_b200 import synth
from detmatch_b200.ops import (points_in_boxes_batch, points_in_boxes_cpu, points_in_boxes_gpu, voxelization,
                               voxelize_batch_packed)
from detmatch_b200.ops import pcdet_roiaware_pool3d as pcdet
from oracle import oracle, vfe_mean
from tests.helpers import assert_same_bits

pytestmark = pytest.mark.gpu


def _draw(seed):
    rng = np.random.default_rng(3000 + seed)
    n = int(rng.integers(1, 6000))
    c = int(rng.choice([4, 5])) if seed % 3 else int(rng.integers(3, 7))
    vs = [float(rng.choice([0.05, 0.1, 0.16, 0.2, 0.32, 0.5, 1.0])) for _ in range(3)]
    lo = [float(rng.uniform(-60, 0)) for _ in range(3)]
    ext = [vs[j] * int(rng.integers(1, 400 if j < 2 else 40)) for j in range(3)]
    rg = lo + [lo[j] + ext[j] for j in range(3)]
    p = 5 if seed % 2 else int(rng.choice([1, 2, 3, 8, 35]))
    v = int(rng.integers(1, 3000))
    pts = rng.uniform(-1.0, 1.0, size=(n, c)).astype(np.float32)
    for j in range(3):
        pts[:, j] = (lo[j] + (pts[:, j] * 0.6 + 0.5) * ext[j]).astype(np.float32)
    if n > 10:
        pts[0, 0], pts[1, 1], pts[2, 2] = rg[3], rg[1], rg[5]
        pts[3, int(rng.integers(0, 3))] = np.nan
        pts[4:8] = pts[8:9]
    t = int(rng.integers(1, 300))
    bxs = synth.random_boxes(t, 5000 + seed, rg).numpy()
    k = min(t, n)
    bxs[:k, 0:3] = pts[:k, 0:3]
    return pts, vs, rg, p, v, bxs


@pytest.mark.parametrize("seed", range(24))
def test_random_config_all_ops(seed):
    pts, vs, rg, p, v, bxs = _draw(seed)
    tp = torch.from_numpy(pts).cuda()
    # voxelization
    assert_same_bits(voxelization(tp, vs, rg, -1, -1).cpu().numpy(), oracle.dynamic_voxelize(pts, vs, rg), "dynamic")
    ev, ec, en = oracle.hard_voxelize(pts, vs, rg, p, v)
    gv, gc, gn = voxelization(tp, vs, rg, p, v)
    assert_same_bits(gc.cpu().numpy(), ec, "coors")
    assert_same_bits(gn.cpu().numpy(), en, "num")
    assert_same_bits(gv.cpu().numpy(), ev, "voxels")
    # the detectors' flow over two copies of the frame, plain and with the mean encoder
    for mean in (False, True):
        vox, num, cb = voxelize_batch_packed([tp, tp], vs, rg, p, v, mean=mean)
        exp_v = vfe_mean.hard_simple_vfe(ev, en) if mean else ev
        assert_same_bits(vox.cpu().numpy(), np.concatenate([exp_v, exp_v]), "packed " + ("means" if mean else "voxels"))
        assert_same_bits(num.cpu().numpy(), np.concatenate([en, en]), "packed num")
        exp_cb = np.concatenate([np.concatenate([np.full((len(en), 1), i, np.int32), ec], axis=1) for i in range(2)])
        assert_same_bits(cb.cpu().numpy(), exp_cb, "coors_batch")
    # point in box, both conventions
    xyz = np.ascontiguousarray(pts[:, :3])
    tx, tb = torch.from_numpy(xyz).cuda(), torch.from_numpy(bxs).cuda()
    assert_same_bits(points_in_boxes_cpu(tx, tb).cpu().numpy(), oracle.points_in_boxes_cpu(xyz, bxs), "pib cpu layout")
    assert_same_bits(points_in_boxes_gpu(tx[None], tb[None]).cpu().numpy(), oracle.points_in_boxes_gpu(xyz[None], bxs[None]),
                     "pib first hit")
    assert_same_bits(points_in_boxes_batch(tx[None], tb[None]).cpu().numpy(), oracle.points_in_boxes_batch(xyz[None], bxs[None]),
                     "pib batch")
    assert_same_bits(pcdet.points_in_boxes_cpu(tx, tb).cpu().numpy(), oracle.pcdet_points_in_boxes_cpu(xyz, bxs), "pcdet cpu")
    assert_same_bits(pcdet.points_in_boxes_gpu(tx[None], tb[None]).cpu().numpy(),
                     oracle.pcdet_points_in_boxes_gpu(xyz[None], bxs[None]), "pcdet first hit")
