"""Device-time survey of every op on the five BASELINE configs (not the headline bench):
python tests/native/bench_ops.py  -> one line per config with ms, GB/s algorithmic, % of HBM peak."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from detmatch_b200 import _cabi, synth  # noqa: E402
from detmatch_b200.ops import points_in_boxes_batch, points_in_boxes_gpu  # noqa: E402
from detmatch_b200.ops.roiaware_pool3d import roiaware_pool3d_ext  # noqa: E402
from detmatch_b200.ops.voxel import HardVoxelizeBatchPlan  # noqa: E402
from detmatch_b200._torch_glue import ptr, stream_ptr  # noqa: E402

PEAK = 6549.1
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAK = float(json.load(open(p))["hbm_gbs"])


def timeit(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def report(name, ms, nbytes, units, unit_name):
    gbs = nbytes / ms / 1e6
    print(f"{name:44s} {ms:9.4f} ms  {units / ms / 1e3:10.1f} M{unit_name}/s  {gbs:8.1f} GB/s algorithmic  {gbs / PEAK:6.1%} of {PEAK:.0f}")


def hard(name, frames=None):
    cfg = synth.CONFIGS[name]
    F = frames or cfg["frames"]
    ci = int(name[1])
    pts = [synth.lidar_frame(cfg["n"], cfg["c"], synth.seed_for(ci, k), cfg["r_max"]).cuda() for k in range(F)]
    plan = HardVoxelizeBatchPlan([cfg["n"]] * F, cfg["c"], cfg["voxel_size"], cfg["point_cloud_range"],
                                 cfg["max_num_points"], cfg["max_voxels"], "cuda:0").bind(pts)
    ms = timeit(plan.run)
    m = plan.voxel_num.cpu().tolist()
    nbytes = sum(cfg["n"] * cfg["c"] * 4 + mm * (cfg["max_num_points"] * cfg["c"] * 4 + 16) for mm in m)
    report(f"{name} hard voxelize x{F} (M~{sum(m) // F})", ms, nbytes, F * cfg["n"], "pts")
    _cabi.profile(True)
    plan.run()
    torch.cuda.synchronize()
    rep = _cabi.profile_report()
    _cabi.profile(False)
    print("      ", {k: round(v[0], 4) for k, v in rep.items()})


def dynamic():
    cfg = synth.CONFIGS["C2"]
    F = cfg["frames"]
    pts = [synth.lidar_frame(cfg["n"], cfg["c"], synth.seed_for(2, k), cfg["r_max"]).cuda() for k in range(F)]
    coors = [torch.empty((cfg["n"], 3), dtype=torch.int32, device="cuda") for _ in range(F)]
    L = _cabi.lib()
    import ctypes
    pp = (ctypes.c_void_p * F)(*[t.data_ptr() for t in pts])
    cc = (ctypes.c_void_p * F)(*[t.data_ptr() for t in coors])
    nn = (ctypes.c_int64 * F)(*[cfg["n"]] * F)
    vs, rg = _cabi.f3(cfg["voxel_size"]), _cabi.f6(cfg["point_cloud_range"])
    dev = torch.device("cuda:0")

    def run():
        _cabi.check(L.pcfe_dynamic_voxelize_batch_f32(pp, nn, F, cfg["c"], vs, rg, cc, 0, stream_ptr(dev)), "dyn")
    ms = timeit(run)
    report(f"C2 dynamic voxelize x{F} (one batched launch)", ms, F * cfg["n"] * (cfg["c"] * 4 + 12), F * cfg["n"], "pts")


def pib():
    c3 = synth.CONFIGS["C3"]
    B, M, T = c3["frames"], c3["n"], c3["boxes"]
    pts = torch.stack([synth.lidar_frame(M, 3, synth.seed_for(3, k), c3["r_max"]) for k in range(B)]).cuda()
    bxs = torch.stack([synth.random_boxes(T, synth.seed_for(3, k) + 500, c3["point_cloud_range"]) for k in range(B)]).cuda()
    out = torch.empty((B, M, T), dtype=torch.int32, device="cuda")
    ms = timeit(lambda: roiaware_pool3d_ext.points_in_boxes_batch(bxs, pts, out))
    report(f"C3 points_in_boxes_batch {B}x{M}x{T}", ms, B * M * T * 4 + B * M * 12 + B * T * 28, B * M * T, "pairs")
    out2 = torch.empty((B, M), dtype=torch.int32, device="cuda")
    ms = timeit(lambda: roiaware_pool3d_ext.points_in_boxes_gpu(bxs, pts, out2))
    report(f"C3-shape points_in_boxes_gpu {B}x{M}x{T}", ms, B * M * 16 + B * T * 28, B * M * T, "pairs")
    out3 = torch.empty((T, M), dtype=torch.int32, device="cuda")
    ms = timeit(lambda: roiaware_pool3d_ext.points_in_boxes_cpu(bxs[0], pts[0], out3))
    report(f"C3-shape points_in_boxes_cpu-layout 1x{M}x{T}", ms, M * T * 4 + M * 12, M * T, "pairs")


def hard_mean(name):
    """Hard voxelization + mean encoder: fused epilogue vs voxelization followed by the stand-alone kernel."""
    import ctypes
    from detmatch_b200.ops.voxel_encoders import hard_simple_vfe
    cfg = synth.CONFIGS[name]
    F, ci, c, P, V = cfg["frames"], int(name[1]), cfg["c"], cfg["max_num_points"], cfg["max_voxels"]
    pts = [synth.lidar_frame(cfg["n"], c, synth.seed_for(ci, k), cfg["r_max"]).cuda() for k in range(F)]
    plan = HardVoxelizeBatchPlan([cfg["n"]] * F, c, cfg["voxel_size"], cfg["point_cloud_range"], P, V, "cuda:0").bind(pts)
    means = torch.empty((F, V, c), dtype=torch.float32, device="cuda")
    frames = (_cabi.Frame * F)()
    for i, t in enumerate(pts):
        frames[i] = _cabi.Frame(t.data_ptr(), t.size(0), means[i].data_ptr(), plan.coors[i].data_ptr(),
                                plan.num_points[i].data_ptr())
    L, dev = _cabi.lib(), torch.device("cuda:0")

    def fused():
        _cabi.check(L.pcfe_hard_voxelize_mean_batch_f32(frames, F, c, plan.vs, plan.rg, None, P, V, ptr(plan.voxel_num),
                                                        ptr(plan.ws), plan.ws.numel(), 0, stream_ptr(dev)), "mean")

    ms_f = timeit(fused)
    m = plan.voxel_num.cpu().tolist()
    nb_f = sum(cfg["n"] * c * 4 + mm * (c * 4 + 16) for mm in m)
    report(f"{name} voxelize + mean VFE, fused x{F}", ms_f, nb_f, F * cfg["n"], "pts")
    # the detectors' flow: voxelize, concatenate the frames' voxels (openpcdet.py:69-76), one encoder call
    plan.run()
    vox_cat = torch.cat([plan.voxels[i, :mm] for i, mm in enumerate(m)])
    num_cat = torch.cat([plan.num_points[i, :mm] for i, mm in enumerate(m)])
    ms_v = timeit(lambda: hard_simple_vfe(vox_cat, num_cat))
    report(f"{name} mean VFE kernel alone on the concatenated voxels", ms_v, sum(m) * ((P + 1) * c * 4 + 4), sum(m), "vox")
    ms_c = timeit(lambda: torch.cat([plan.voxels[i, :mm] for i, mm in enumerate(m)]))
    print(f"       unfused flow: voxelize {timeit(plan.run):.4f} + concatenate {ms_c:.4f} + encoder {ms_v:.4f} ms")
    _cabi.profile(True)
    fused()
    torch.cuda.synchronize()
    print("      ", {k: round(v[0], 4) for k, v in _cabi.profile_report().items()})
    _cabi.profile(False)


def dyn_scatter():
    """C2 frames: dynamic voxelization coordinates -> DynamicScatter over the whole batch (b, z, y, x)."""
    from detmatch_b200.ops import dynamic_scatter, voxelization
    cfg = synth.CONFIGS["C2"]
    F = cfg["frames"]
    pts = [synth.lidar_frame(cfg["n"], cfg["c"], synth.seed_for(2, k), cfg["r_max"]).cuda() for k in range(F)]
    coors = [voxelization(p, cfg["voxel_size"], cfg["point_cloud_range"], -1, -1) for p in pts]
    feats = torch.cat(pts)
    cb = torch.cat([torch.nn.functional.pad(c, (1, 0), value=i) for i, c in enumerate(coors)]).contiguous()
    n, c = feats.shape
    for r in ("max", "mean"):
        ms = timeit(lambda: dynamic_scatter(feats, cb, r), reps=20)
        out, oc = dynamic_scatter(feats, cb, r)
        m = out.size(0)
        nbytes = n * (c * 4 + 16) + m * (c * 4 + 16)
        report(f"C2-shape DynamicScatter {r} x{F} (M={m}, incl. 2 host syncs)", ms, nbytes, n, "pts")
    _cabi.profile(True)
    dynamic_scatter(feats, cb, "mean")
    torch.cuda.synchronize()
    print("      ", {k: round(v[0], 4) for k, v in _cabi.profile_report().items()})
    _cabi.profile(False)


def pcdet_pib():
    from detmatch_b200.ops.pcdet_roiaware_pool3d import roiaware_pool3d_cuda
    c3 = synth.CONFIGS["C3"]
    B, M, T = c3["frames"], c3["n"], c3["boxes"]
    pts = torch.stack([synth.lidar_frame(M, 3, synth.seed_for(3, k), c3["r_max"]) for k in range(B)]).cuda()
    bxs = torch.stack([synth.random_boxes(T, synth.seed_for(3, k) + 500, c3["point_cloud_range"]) for k in range(B)]).cuda()
    out = torch.empty((B, M), dtype=torch.int32, device="cuda")
    ms = timeit(lambda: roiaware_pool3d_cuda.points_in_boxes_gpu(bxs, pts, out))
    report(f"C3-shape OpenPCDet points_in_boxes_gpu {B}x{M}x{T}", ms, B * M * 16 + B * T * 28, B * M * T, "pairs")


def roiaware():
    """PartA2-shape RoI-aware pooling (128 RoIs, 16 384 points, 128 channels, 14^3 voxels, 128 slots), ours
    vs the reference's own kernel compiled for sm_100a (oracle/_ref/detmatch_ref_roiaware.so), wrapper
    allocations included on both sides (the reference zero-fills its three work tensors)."""
    from oracle import ref
    g = torch.Generator().manual_seed(1)
    n, m, c, o, mp = 128, 16384, 128, 14, 128
    rois = torch.cat([torch.rand((n, 2), generator=g) * 80 - 40, torch.rand((n, 1), generator=g) * 2 - 2,
                      torch.rand((n, 3), generator=g) * 3 + 1, (torch.rand((n, 1), generator=g) * 2 - 1) * 3.14], dim=1).cuda()
    pts = torch.cat([torch.rand((m, 2), generator=g) * 80 - 40, torch.rand((m, 1), generator=g) * 4 - 2], dim=1)
    own = torch.randint(0, n, (m // 2,), generator=g)
    pts[:m // 2, :2] = rois.cpu()[own, :2] + (torch.rand((m // 2, 2), generator=g) - 0.5)
    pts = pts.cuda()
    feats = torch.rand((m, c), generator=g).cuda()
    for mode in (0, 1):
        def ours():
            pooled = torch.empty((n, o, o, o, c), device="cuda")
            argmax = torch.empty((n, o, o, o, c), dtype=torch.int32, device="cuda")
            lists = torch.empty((n, o, o, o, mp), dtype=torch.int32, device="cuda")
            roiaware_pool3d_ext.forward(rois, pts, feats, argmax, lists, pooled, mode)
        ms = timeit(ours, reps=10, warm=2)
        nbytes = n * o ** 3 * c * (8 if mode == 0 else 4) + m * (12 + c * 4)
        report(f"RoIAwarePool3d fwd mode {mode} {n}x{m}x{c} out {o}", ms, nbytes, n * m, "pairs")
        if ref.roiaware_available():
            ext = ref.roiaware_module()

            def theirs():
                pooled = torch.zeros((n, o, o, o, c), device="cuda")
                argmax = torch.zeros((n, o, o, o, c), dtype=torch.int32, device="cuda")
                lists = torch.zeros((n, o, o, o, mp), dtype=torch.int32, device="cuda")
                ext.forward(rois, pts, feats, argmax, lists, pooled, mode)
            ms_r = timeit(theirs, reps=3, warm=1)
            print(f"       reference kernel (sm_100a build): {ms_r:.3f} ms  -> {ms_r / ms:.1f}x")


if __name__ == "__main__":
    torch.cuda.set_device(0)
    if len(sys.argv) > 1 and sys.argv[1] == "roiaware":
        roiaware()
        sys.exit(0)
    hard("C1", 16)
    dynamic()
    pib()
    pcdet_pib()
    dyn_scatter()
    hard("C4")
    hard_mean("C4")
    hard("C5", 16)
    roiaware()
