"""oracle/scatter.py -- numpy restatement of DynamicScatter's forward and backward.

TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product).

Restates mmdet3d/ops/voxel/src/scatter_points_cuda.cu:183-303:
  * rows with a negative coordinate are dropped (:202), the voxels are the remaining distinct rows
    in lexicographic order (:204-210, at::unique_dim sorted), coors_map / reduce_count as there;
  * forward reduce (:85-106, :243): max over the voxel's points (NaN ignored like fmaxf), sum in
    POINT ORDER in float32 (the reference's atomicAdd order is unspecified), mean = sum / count;
  * backward (:108-181): sum copies, mean divides by the count, max routes the gradient to the
    lowest-index point whose feature equals the maximum.
The reference has no CPU implementation of this op; pinning is against the reference TEST's own
construction (tests/test_models/test_voxel_encoder/test_dynamic_scatter.py:55-84: torch unique +
masked mean / max per voxel, allclose atol 1e-2 rtol 1e-5), evaluated by torch in the build
container and stored in tests/golden/dynamic_scatter.npz.
"""
import numpy as np


def forward(feats, coors, reduce_type):
    feats = np.asarray(feats, dtype=np.float32)
    coors = np.asarray(coors, dtype=np.int32)
    n, c = feats.shape
    keep = (coors >= 0).all(axis=1)
    kept_rows = coors[keep]
    if kept_rows.shape[0] == 0:
        return (np.zeros((0, c), np.float32), np.zeros((0, coors.shape[1]), np.int32), np.full((n,), -1, np.int32),
                np.zeros((0,), np.int32))
    voxel_coors, inv, count = np.unique(kept_rows, axis=0, return_inverse=True, return_counts=True)
    coors_map = np.full((n,), -1, np.int32)
    coors_map[keep] = inv.reshape(-1).astype(np.int32)
    m = voxel_coors.shape[0]
    if reduce_type == "max":
        out = np.full((m, c), -np.inf, np.float32)
        f = np.where(np.isnan(feats), -np.inf, feats).astype(np.float32)
        np.maximum.at(out, coors_map[keep], f[keep])
    else:
        out = np.zeros((m, c), np.float32)
        np.add.at(out, coors_map[keep], feats[keep])  # unbuffered, in point order, float32
        if reduce_type == "mean":
            out = (out / count.astype(np.float32).reshape(-1, 1)).astype(np.float32)
    return out, voxel_coors.astype(np.int32), coors_map, count.astype(np.int32)


def backward(grad_voxel, feats, voxel_feats, coors_map, count, reduce_type):
    feats = np.asarray(feats, dtype=np.float32)
    n, c = feats.shape
    grad = np.zeros((n, c), np.float32)
    keep = coors_map >= 0
    if reduce_type in ("sum", "mean"):
        g = grad_voxel[coors_map[keep]]
        if reduce_type == "mean":
            g = (g / count[coors_map[keep]].astype(np.float32).reshape(-1, 1)).astype(np.float32)
        grad[keep] = g
        return grad
    m = voxel_feats.shape[0]
    src = np.full((m, c), n, np.int64)
    idx = np.nonzero(keep)[0]
    eq = feats[idx] == voxel_feats[coors_map[idx]]
    for j in range(c):
        np.minimum.at(src[:, j], coors_map[idx][eq[:, j]], idx[eq[:, j]])
    for j in range(c):
        ok = src[:, j] < n
        grad[src[ok, j], j] = grad_voxel[ok, j]
    return grad
