"""C4 hard voxelization with the fused mean epilogue, a few steps (profiling target):
python tests/native/run_mean.py [steps] [packed]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from detmatch_b200 import _cabi, synth  # noqa: E402
from detmatch_b200._torch_glue import ptr, stream_ptr  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
cfg = synth.CONFIGS["C4"]
F, c, P, V = cfg["frames"], cfg["c"], cfg["max_num_points"], cfg["max_voxels"]
pts = [synth.lidar_frame(cfg["n"], c, synth.seed_for(4, k), cfg["r_max"]).cuda() for k in range(F)]
means = torch.empty((F, V, c), dtype=torch.float32, device="cuda")
coors = torch.empty((F, V, 3), dtype=torch.int32, device="cuda")
num = torch.empty((F, V), dtype=torch.int32, device="cuda")
vnum = torch.empty((F,), dtype=torch.int32, device="cuda")
L, dev = _cabi.lib(), torch.device("cuda:0")
vs, rg = _cabi.f3(cfg["voxel_size"]), _cabi.f6(cfg["point_cloud_range"])
need = L.pcfe_hard_voxelize_workspace_bytes(cfg["n"], F, 0, vs, rg, P, V)
ws = torch.empty(need, dtype=torch.uint8, device="cuda")
frames = (_cabi.Frame * F)()
for i, t in enumerate(pts):
    frames[i] = _cabi.Frame(t.data_ptr(), t.size(0), means[i].data_ptr(), coors[i].data_ptr(), num[i].data_ptr())
for _ in range(steps):
    _cabi.check(L.pcfe_hard_voxelize_mean_batch_f32(frames, F, c, vs, rg, None, P, V, ptr(vnum), ptr(ws), ws.numel(), 0,
                                                    stream_ptr(dev)), "mean")
torch.cuda.synchronize()
print("voxels", int(vnum.sum()))
