"""oracle/vfe_mean.py -- numpy restatement of the mean voxel feature encoder.

TEST INFRASTRUCTURE ONLY (imported by tests/ and __graft_entry__.smoke(), never by the product).

Restates HardSimpleVFE.forward, mmdet3d/models/voxel_encoders/voxel_encoder.py:41-44:
    features[:, :, :num_features].sum(dim=1) / num_points.type_as(features).view(-1, 1)
in float32 with a FIXED association: slot order, (((s0 + s1) + s2) + ...) over all max_points
slots (the zero padding included), then one IEEE division by float32(num_points).

Pinning: ATen does not promise an association for sum(); its CPU kernel picks one from the memory
layout (measured in this container, torch 2.11: identical to the slot-order sum for C = 4 and
P = 5, different in the last bit for ~1 % of the elements for C = 5).  tests/golden/vfe_mean.npz
holds outputs of the reference expression itself (tests/golden/make_golden.py: vfe_cases);
tests/test_oracle.py pins this restatement to them within 2 ulp and reports where it is exact.
The CUDA kernels are compared with THIS restatement bit for bit.
"""
import numpy as np


def hard_simple_vfe(features, num_points, num_features=None):
    features = np.asarray(features, dtype=np.float32)
    n, p, c = features.shape
    nf = c if num_features is None else int(num_features)
    acc = features[:, 0, :nf].copy()
    for s in range(1, p):
        acc = (acc + features[:, s, :nf]).astype(np.float32)  # float32 add, round to nearest even
    with np.errstate(divide="ignore", invalid="ignore"):
        return (acc / np.asarray(num_points).astype(np.float32).reshape(-1, 1)).astype(np.float32)
