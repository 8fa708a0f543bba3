"""Small-batch latency of hard voxelization (not the headline metric): device time per step of the
pre-allocated batched plan at B = 1, 4, 8, 16, 64 frames for C1 / C4 / C5 frames, and the host wall
time of the unchanged per-frame API (Voxelization.forward incl. its voxel_num read-back).
    python tests/native/latency.py [--debug NAME=VALUE ...]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from detmatch_b200 import _cabi, synth  # noqa: E402
from detmatch_b200.ops import Voxelization  # noqa: E402
from detmatch_b200.ops.voxel import HardVoxelizeBatchPlan  # noqa: E402

for kv in sys.argv[1:]:
    if "=" in kv:
        k, v = kv.split("=")
        _cabi.debug_set(k, int(v))
dev = torch.device("cuda", 0)
out = {}
for name in ("C1", "C4", "C5"):
    cfg = synth.CONFIGS[name]
    ci = int(name[1])
    N, C, P, V = cfg["n"], cfg["c"], cfg["max_num_points"], cfg["max_voxels"]
    res = {}
    for B in (1, 4, 8, 16, 64):
        if name == "C5" and B > 16:
            continue
        pts = [synth.lidar_frame(N, C, synth.seed_for(ci, k), cfg["r_max"]).to(dev) for k in range(B)]
        plan = HardVoxelizeBatchPlan([N] * B, C, cfg["voxel_size"], cfg["point_cloud_range"], P, V, dev).bind(pts)
        for _ in range(10):
            plan.run()
        torch.cuda.synchronize()
        reps = 200
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            plan.run()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        m = plan.voxel_num.cpu().tolist()
        algo = sum(N * C * 4 + mm * (P * C * 4 + 16) for mm in m)
        # the same step replayed from a CUDA graph (no host launch cost)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        gms = None
        try:
            with torch.cuda.stream(s):
                plan.run()
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=s):
                    plan.run()
            for _ in range(5):
                g.replay()
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps):
                g.replay()
            b.record()
            torch.cuda.synchronize()
            gms = a.elapsed_time(b) / reps
        except Exception as e:  # noqa: BLE001
            gms = "capture failed: " + str(e)[:80]
        res[B] = {"ms_per_step": round(ms, 4), "graph_ms_per_step": gms if isinstance(gms, str) else round(gms, 4),
                  "us_per_frame": round(ms * 1e3 / B, 2), "GBps_algorithmic": round(algo / ms / 1e6, 1)}
        del plan, pts
    # per-frame API: Voxelization.forward (allocations + launch sequence + voxel_num read-back)
    layer = Voxelization(cfg["voxel_size"], cfg["point_cloud_range"], P, V).eval()
    p = synth.lidar_frame(N, C, synth.seed_for(ci, 0), cfg["r_max"]).to(dev)
    for _ in range(10):
        layer(p)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100):
        layer(p)
    torch.cuda.synchronize()
    res["forward_wall_us"] = round((time.perf_counter() - t0) / 100 * 1e6, 1)
    out[name] = res
print(json.dumps(out))
