"""Host-side enqueue cost of one batched call vs its device time: python tests/native/enqueue_cost.py"""
import sys, time, torch
sys.path.insert(0, '.')
from detmatch_b200 import synth
from detmatch_b200.ops.voxel import HardVoxelizeBatchPlan
for name, F in (("C1", 1), ("C1", 16), ("C4", 64)):
    cfg = synth.CONFIGS[name]
    pts = [synth.lidar_frame(cfg["n"], cfg["c"], 100 + k, cfg["r_max"]).cuda() for k in range(F)]
    plan = HardVoxelizeBatchPlan([cfg["n"]] * F, cfg["c"], cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"], "cuda:0").bind(pts)
    for _ in range(20): plan.run()
    torch.cuda.synchronize()
    n = 3000
    t0 = time.perf_counter()
    for _ in range(n): plan.run()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name} x{F}: enqueue {1e6*(t1-t0)/n:.1f} us/call, total {1e6*(t2-t0)/n:.1f} us/call")
    # host cost alone: 40 calls into an empty launch queue
    best = 1e9
    for _ in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(40): plan.run()
        best = min(best, (time.perf_counter() - t0) / 40)
        torch.cuda.synchronize()
    print(f"      host-side cost of one call (empty queue): {1e6*best:.1f} us")
