"""Deterministic synthetic inputs for parity tests and bench.py.

There is no dataset on the GPU box, so frames are drawn from a LiDAR-like generator whose
statistics resemble what the reference's pipelines hand to ``Voxelization`` (points are
range-filtered and then shuffled: configs/detmatch/001/detmatch/split_0.py:584-586), and boxes
from the distribution SURVEY.md section 8(d) fixes for config C3.

Everything is generated on the CPU with a seeded ``torch.Generator`` so that the CPU oracle and
the CUDA path see byte-identical inputs on any machine.
"""
import math

import torch

# The five BASELINE.json configs (C1..C5).
CONFIGS = {
    "C1": dict(kind="hard", n=120_000, c=4, voxel_size=[0.05, 0.05, 0.1],
               point_cloud_range=[0, -40, -3, 70.4, 40, 1], max_num_points=5, max_voxels=16_000,
               frames=1, r_max=80.0),
    "C2": dict(kind="dynamic", n=120_000, c=4, voxel_size=[0.05, 0.05, 0.1],
               point_cloud_range=[0, -40, -3, 70.4, 40, 1], max_num_points=-1, max_voxels=-1,
               frames=16, r_max=80.0),
    "C3": dict(kind="points_in_boxes_batch", n=120_000, c=3, boxes=200, frames=16, r_max=80.0,
               point_cloud_range=[0, -40, -3, 70.4, 40, 1]),
    "C4": dict(kind="hard", n=180_000, c=5, voxel_size=[0.1, 0.1, 0.15],
               point_cloud_range=[-75.2, -75.2, -2, 75.2, 75.2, 4], max_num_points=5,
               max_voxels=150_000, frames=64, r_max=80.0),
    "C5": dict(kind="hard", n=300_000, c=5, voxel_size=[0.25, 0.25, 8],
               point_cloud_range=[-50, -50, -5, 50, 50, 3], max_num_points=64, max_voxels=40_000,
               frames=128, r_max=60.0),
}


def seed_for(config_index, frame):
    """seed = 1000 * config + frame index."""
    return 1000 * int(config_index) + int(frame)


def lidar_frame(n, c, seed, r_max=80.0, order="shuffled"):
    """One LiDAR-like frame, float32 (n, c), randomly permuted (mirrors PointShuffle; the default and
    what every BASELINE config uses) or, with ``order="sweep"``, in the firing order of a spinning
    sensor (azimuth step by azimuth step, the 64 beams of a step together: what an un-shuffled test-time
    frame looks like -- consecutive points are neighbours and often share a voxel).

    64 beams between -24.8 and +2 degrees, uniform azimuth; each return is the nearer of the
    ground plane (sensor 1.73 m above it) and an obstacle at r = 2 + (r_max-2)*u^2.
    """
    g = torch.Generator().manual_seed(int(seed))
    u = torch.rand(n, generator=g, dtype=torch.float64)
    az = (torch.rand(n, generator=g, dtype=torch.float64) * 2.0 - 1.0) * math.pi
    beam = torch.randint(0, 64, (n,), generator=g).to(torch.float64)
    elev = torch.deg2rad(-24.8 + beam * (26.8 / 63.0))
    r_ground = torch.where(elev < 0, 1.73 / torch.tan(-elev).clamp_min(1e-9),
                           torch.full_like(elev, float("inf")))
    r_obst = 2.0 + (r_max - 2.0) * u * u
    r = torch.minimum(r_ground, r_obst)
    x = r * torch.cos(elev) * torch.cos(az)
    y = r * torch.cos(elev) * torch.sin(az)
    z = r * torch.sin(elev) + 0.02 * torch.randn(n, generator=g, dtype=torch.float64)
    cols = [x, y, z]
    for _ in range(c - 3):
        cols.append(torch.rand(n, generator=g, dtype=torch.float64))
    pts = torch.stack(cols, dim=1).to(torch.float32)
    perm = torch.randperm(n, generator=g)
    if order == "sweep":
        step = torch.floor((az + math.pi) / (2.0 * math.pi) * 2650.0)  # ~0.136 degree azimuth steps
        perm = torch.argsort(step * 64.0 + beam, stable=True)
    else:
        assert order == "shuffled"
    return pts[perm].contiguous()


def uniform_frame(n, c, seed, point_cloud_range, inflate=0.05):
    """Adversarial frame: xyz uniform over the range inflated by 5 % per side (parity only)."""
    g = torch.Generator().manual_seed(int(seed))
    lo = torch.tensor(point_cloud_range[:3], dtype=torch.float64)
    hi = torch.tensor(point_cloud_range[3:], dtype=torch.float64)
    span = hi - lo
    lo = lo - inflate * span
    hi = hi + inflate * span
    xyz = lo + (hi - lo) * torch.rand(n, 3, generator=g, dtype=torch.float64)
    feats = torch.rand(n, max(c - 3, 0), generator=g, dtype=torch.float64)
    return torch.cat([xyz, feats], dim=1).to(torch.float32).contiguous()


def random_boxes(t, seed, point_cloud_range):
    """(t, 7) boxes (cx, cy, cz_bottom, w, l, h, rz) per SURVEY.md section 8(d), config C3."""
    g = torch.Generator().manual_seed(int(seed))
    lo = torch.tensor(point_cloud_range[:2], dtype=torch.float64)
    hi = torch.tensor(point_cloud_range[3:5], dtype=torch.float64)
    cxy = lo + (hi - lo) * torch.rand(t, 2, generator=g, dtype=torch.float64)
    cz = -3.0 + 2.0 * torch.rand(t, 1, generator=g, dtype=torch.float64)
    w = 0.5 + 1.5 * torch.rand(t, 1, generator=g, dtype=torch.float64)
    ln = 0.8 + 3.5 * torch.rand(t, 1, generator=g, dtype=torch.float64)
    h = 1.0 + 1.0 * torch.rand(t, 1, generator=g, dtype=torch.float64)
    rz = (torch.rand(t, 1, generator=g, dtype=torch.float64) * 2.0 - 1.0) * math.pi
    return torch.cat([cxy, cz, w, ln, h, rz], dim=1).to(torch.float32).contiguous()


def face_points(boxes, seed, per_box=64):
    """Points snapped onto faces, edges and corners of ``boxes`` (parity-only stress set).

    For every box the local coordinates are drawn from {-half, -half+ulp-ish, 0, +half, ...} and
    rotated back with float64 trig, so a good fraction lands within one float32 ulp of a face.
    """
    g = torch.Generator().manual_seed(int(seed))
    b = boxes.to(torch.float64)
    t = b.shape[0]
    sel = torch.tensor([-1.0, -0.999999, -0.5, 0.0, 0.5, 0.999999, 1.0, 1.000001], dtype=torch.float64)
    ix = torch.randint(0, len(sel), (t, per_box, 3), generator=g)
    f = sel[ix]  # (t, per_box, 3) multipliers of the half extents
    half = torch.stack([b[:, 4] / 2, b[:, 3] / 2, b[:, 5] / 2], dim=1)  # (l/2 along local x, w/2, h/2)
    loc = f * half[:, None, :]
    rot = b[:, 6] + math.pi / 2
    ca, sa = torch.cos(rot)[:, None], torch.sin(rot)[:, None]
    # inverse of local = R(rot) * shift  ->  shift = R(-rot) * local
    sx = loc[..., 0] * ca + loc[..., 1] * sa
    sy = -loc[..., 0] * sa + loc[..., 1] * ca
    x = sx + b[:, 0:1]
    y = sy + b[:, 1:2]
    z = loc[..., 2] + (b[:, 2:3] + b[:, 5:6] / 2)
    return torch.stack([x, y, z], dim=-1).reshape(-1, 3).to(torch.float32).contiguous()
