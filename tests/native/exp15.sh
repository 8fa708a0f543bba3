cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp15.log
: > $L
timeout 900 python -m pytest tests/test_gpu_pib.py tests/test_gpu_pcdet_pib.py tests/test_gpu_roiaware.py -m gpu -x -q 2>&1 | tail -3 >> $L
python - >> $L 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, '/root/repo')
from detmatch_b200 import synth, _cabi
from detmatch_b200.ops.roiaware_pool3d import roiaware_pool3d_ext
c3 = synth.CONFIGS["C3"]
for (B, M, T) in ((16, 120000, 200), (16, 120000, 64), (4, 120000, 200), (16, 120000, 203)):
    pts = torch.stack([synth.lidar_frame(M, 3, synth.seed_for(3, k), c3["r_max"]) for k in range(B)]).cuda()
    bxs = torch.stack([synth.random_boxes(T, 3000 + k, c3["point_cloud_range"]) for k in range(B)]).cuda()
    out = torch.empty((B, M, T), dtype=torch.int32, device="cuda")
    for mp in (0, 1, 0, 1):
        _cabi.debug_set("pib_map", mp)
        for _ in range(5): roiaware_pool3d_ext.points_in_boxes_batch(bxs, pts, out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50): roiaware_pool3d_ext.points_in_boxes_batch(bxs, pts, out)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 50
        nbytes = B * M * T * 4 + B * M * 12 + B * T * 28
        print(f"points_in_boxes_batch {B}x{M}x{T} pib_map={mp}: {ms:.4f} ms  {nbytes / ms / 1e6:.0f} GB/s  {nbytes / ms / 1e6 / 6549.1:.3f} of the roofline")
PY
cat $L
