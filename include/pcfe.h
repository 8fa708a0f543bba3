/*
 * include/pcfe.h -- C ABI of the B200-native point-cloud front end (libpcfe.so).
 *
 * This is the drop-in boundary for the one hot path of Divadi/DetMatch this repository
 * replaces: point-to-cell assignment (hard / dynamic voxelization) and point-to-box assignment
 * (points_in_boxes_{gpu,batch,cpu}).  Every entry point below names the reference interface it
 * replaces (paths relative to the reference tree).  Plain pointers and sizes only: no torch,
 * pybind or C++ types cross this boundary.  INTEGRATION.md shows the binding a maintainer of
 * the reference would add (ctypes / pybind stubs for voxel_layer and roiaware_pool3d_ext).
 *
 * Conventions
 *   - All data pointers are DEVICE pointers on CUDA device `device` unless marked "host".
 *   - `stream` is a cudaStream_t (passed as void*); NULL means the legacy default stream.
 *     Every call only ENQUEUES work on `stream`; nothing here synchronises the host.  Results
 *     (including voxel counts) stay on the device -- the caller decides when to read them.
 *   - The current CUDA device of the calling thread is preserved.
 *   - Return value: 0 = PCFE_OK, < 0 = argument error (see enum), > 0 = a cudaError_t.
 *     Never exit()s, never throws (the reference's launchers call exit(-1):
 *     mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cuda.cu:120-125,145-149).
 *   - float32 points only (the reference's CUDA path is float-only in practice:
 *     mmdet3d/ops/voxel/src/voxelization_cuda.cu:294).
 *   - Results are bit-identical to the reference's CPU ops (voxelization_cpu.cpp,
 *     points_in_boxes_cpu.cpp), NOT to its CUDA ops where the two differ.
 */
#ifndef PCFE_H_
#define PCFE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCFE_VERSION 100 /* 0.1.0 */

enum {
  PCFE_OK = 0,
  PCFE_ERR_NULL = -1,        /* a required pointer is NULL */
  PCFE_ERR_SHAPE = -2,       /* negative size, c < 3, boxes/points inner dim wrong */
  PCFE_ERR_GRID = -3,        /* voxel grid empty or has >= 2^32-1 cells */
  PCFE_ERR_WORKSPACE = -4,   /* workspace too small (see *_workspace_bytes) */
  PCFE_ERR_ALIGN = -5,       /* pointer not aligned to its element size */
  PCFE_ERR_CAPS = -6,        /* max_points / max_voxels negative (use dynamic_voxelize for -1) */
  PCFE_ERR_TOO_LARGE = -7,   /* n >= 2^31-1 points in one frame, or t too large */
  PCFE_ERR_DEVICE = -8       /* not an sm_100 device / no CUDA device */
};

int pcfe_version(void);
/* Static string for a return code (argument errors and cudaGetErrorString). */
const char* pcfe_error_string(int code);
/* Number of kernel launches + memsets enqueued by this library since load (bench.py's
 * "gpu_launches" is the difference across the timed region). */
uint64_t pcfe_launch_count(void);

/* Optional per-kernel device timing of the hard-voxelize launch sequence (used by bench.py to
 * attribute the step time; off by default, costs two event records per launch when on).
 * pcfe_profile_report synchronises the recorded events, writes "name total_ms launches" lines
 * into buf and clears the records; returns the number of lines. */
int pcfe_profile_enable(int on);
int pcfe_profile_report(char* buf, size_t cap);

/* Test / tuning knobs (not part of the drop-in surface; every setting computes the same results):
 * "hv_path" 0 auto | 1 global-memory path | 2 shared-memory bucket path | 3 persistent pipeline;
 * "hv_force_overflow" 1 = every frame also takes the overflow fallback; "hv_bucket_avg" = target
 * points per bucket; "hv_bucket_variant", "hv_expand_variant", "hv_pdl", ... select kernel
 * variants (see pcfe_debug_set in csrc/voxelize.cu); "pib_grid" 0 = first-hit point-in-box
 * assignment by brute force over all boxes instead of the per-frame box grid. */
int pcfe_debug_set(const char* name, int value);

/* grid[j] = (int)roundf((range[3+j]-range[j])/voxel_size[j]) in float32, x y z order.
 * Replaces: voxelization_cpu.cpp:119-122 / voxelization_cuda.cu:200-206.  Host-only helper. */
int pcfe_grid_size(const float voxel_size[3], const float coors_range[6], int32_t grid[3]);

/* ---------------------------------------------------------------------------------------------
 * dynamic voxelization
 * Replaces: voxel_layer.dynamic_voxelize(points, coors, voxel_size, coors_range, NDim=3)
 *           mmdet3d/ops/voxel/src/voxelization.h:71-83, voxelization_cpu.cpp:144-169,
 *           call site mmdet3d/ops/voxel/voxelize.py:42-44.
 * points (n, c) float32 row-major, c >= 3; coors (n, 3) int32 = (z, y, x), or (-1,-1,-1) for a
 * point outside the range (CPU semantics, voxelization_cpu.cpp:32-37; the reference's CUDA
 * kernel writes partial -1s, voxelization_cuda.cu:38-54 -- not reproduced).
 * voxel_size / coors_range are HOST float32 arrays (pybind narrows the Python floats to
 * std::vector<float> at this same point).  n == 0 is a no-op.
 * ------------------------------------------------------------------------------------------- */
int pcfe_dynamic_voxelize_f32(const float* points, int64_t n, int c,
                              const float voxel_size[3], const float coors_range[6],
                              int32_t* coors, int device, void* stream);

/* Batched: `num_frames` independent frames in one launch.  Host arrays of device pointers. */
int pcfe_dynamic_voxelize_batch_f32(const float* const* points, const int64_t* n, int num_frames,
                                    int c, const float voxel_size[3], const float coors_range[6],
                                    int32_t* const* coors, int device, void* stream);

/* ---------------------------------------------------------------------------------------------
 * hard voxelization
 * Replaces: voxel_layer.hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size,
 *                                     coors_range, max_points, max_voxels, NDim=3) -> voxel_num
 *           mmdet3d/ops/voxel/src/voxelization.h:51-69, voxelization_cpu.cpp:43-142,
 *           voxelization_cuda.cu:184-326; call site mmdet3d/ops/voxel/voxelize.py:46-58.
 *
 * Semantics (voxelization_cpu.cpp:68-96): distinct in-range voxels are numbered in order of
 * their first point; the first `max_voxels` of them are kept; inside a voxel the first
 * `max_points` points (by index) fill slots 0.. ; all other slots are +0.0f.
 *   voxels      (max_voxels, max_points, c) float32   rows [0, voxel_num) are written COMPLETELY
 *   coors       (max_voxels, 3) int32 (z, y, x)       (data and zero padding); rows >= voxel_num
 *   num_points  (max_voxels,) int32                   are not touched.
 *   voxel_num   device int32[1] = min(#distinct voxels, max_voxels).  The reference returns
 *               this as a host int (forcing a device sync, voxelization_cuda.cu:322); here the
 *               caller reads it when it needs it.
 * max_points >= 0 and max_voxels >= 0 (the reference's -1 "unbounded" is routed to
 * dynamic_voxelize by its Python wrapper, voxelize.py:41).
 * workspace: >= pcfe_hard_voxelize_workspace_bytes(...) bytes, 256-byte aligned, device memory;
 * contents are scratch.  n == 0 sets voxel_num = 0.
 * ------------------------------------------------------------------------------------------- */
typedef struct pcfe_frame {
  const float* points; /* (n, c) */
  int64_t n;
  float* voxels;       /* (max_voxels, max_points, c) */
  int32_t* coors;      /* (max_voxels, 3) */
  int32_t* num_points; /* (max_voxels,) */
} pcfe_frame_t;

/* Minimum workspace for a batch whose largest frame has n_max points.  `frames_in_flight`
 * (1..num_frames) is how many frames are processed per wave; larger values need more memory
 * but fewer launches.  Pass 0 to let the library pick (sized to keep scratch L2-resident). */
size_t pcfe_hard_voxelize_workspace_bytes(int64_t n_max, int num_frames, int frames_in_flight,
                                          const float voxel_size[3], const float coors_range[6],
                                          int max_points, int max_voxels);

int pcfe_hard_voxelize_f32(const float* points, int64_t n, int c,
                           const float voxel_size[3], const float coors_range[6],
                           int max_points, int max_voxels,
                           float* voxels, int32_t* coors, int32_t* num_points,
                           int32_t* voxel_num, void* workspace, size_t workspace_bytes,
                           int device, void* stream);

/* Batched: what the detectors' per-frame Python loop does (mmdet3d/models/detectors/
 * openpcdet.py:59-76, voxelnet.py:50-67), as one launch sequence.  `frames` is a HOST array;
 * voxel_num is device int32[num_frames]. */
int pcfe_hard_voxelize_batch_f32(const pcfe_frame_t* frames, int num_frames, int c,
                                 const float voxel_size[3], const float coors_range[6],
                                 int max_points, int max_voxels, int32_t* voxel_num,
                                 void* workspace, size_t workspace_bytes, int device, void* stream);

/* The same with PointsRangeFilter fused in front (mmdet3d/datasets/pipelines/transforms_3d.py
 * PointsRangeFilter -> BasePoints.in_range_3d, mmdet3d/core/points/base_points.py:223-228): only
 * points with filter_range[j] < p[j] < filter_range[3 + j] on x, y, z (strict, float32) take part.
 * Results equal "drop the other rows (order kept), then hard_voxelize" -- voxels, coors,
 * num_points, voxel_num bit for bit -- without the compaction pass or the filtered copy. */
int pcfe_hard_voxelize_batch_filtered_f32(const pcfe_frame_t* frames, int num_frames, int c,
                                          const float voxel_size[3], const float coors_range[6],
                                          const float filter_range[6], int max_points, int max_voxels,
                                          int32_t* voxel_num, void* workspace, size_t workspace_bytes,
                                          int device, void* stream);

/* ---------------------------------------------------------------------------------------------
 * mean voxel feature encoder (the consumer of hard voxelization in every SECOND / PV-RCNN config)
 * Replaces: HardSimpleVFE.forward, mmdet3d/models/voxel_encoders/voxel_encoder.py:27-44:
 *           features[:, :, :num_features].sum(dim=1) / num_points.type_as(features).view(-1, 1)
 * Arithmetic: float32, slot-order sum ((((s0 + s1) + s2) + ...) over all max_points slots, absent
 * slots being +0) followed by one IEEE divide by (float)num_points.  (ATen's CPU sum picks its
 * association from the memory layout -- it equals the slot-order sum for some shapes and differs
 * in the last bit for others; oracle/vfe_mean.py is the restatement the tests pin bit for bit, and
 * they bound the distance to ATen's result by 2 ulp.)
 *
 * pcfe_voxel_mean_f32: voxels (m, max_points, c) zero padded, num_points (m,) -> out (m, c).
 * voxel_num (device int32, may be NULL) limits the rows to min(*voxel_num, m) so that the call can
 * follow a batched voxelization without a host synchronisation.
 *
 * pcfe_hard_voxelize_mean_batch_f32: hard voxelization with the encoder fused into the expansion:
 * frames[i].voxels is a (max_voxels, c) buffer that receives the means; the (max_voxels,
 * max_points, c) tensor is never written (its 100 B per voxel are 3/4 of the step's compulsory
 * bytes).  coors / num_points / voxel_num as pcfe_hard_voxelize_batch_f32.  filter_range may be
 * NULL (no fused PointsRangeFilter).  Only max_points == 5 with c == 4 or 5 and 16-byte aligned
 * frames (the record path); PCFE_ERR_SHAPE otherwise -- callers then run
 * pcfe_hard_voxelize_batch_f32 + pcfe_voxel_mean_f32 (detmatch_b200.ops.voxel does).
 * ------------------------------------------------------------------------------------------- */
int pcfe_voxel_mean_f32(const float* voxels, const int32_t* num_points, const int32_t* voxel_num,
                        int64_t m, int max_points, int c, float* out, int device, void* stream);

int pcfe_hard_voxelize_mean_batch_f32(const pcfe_frame_t* frames, int num_frames, int c,
                                      const float voxel_size[3], const float coors_range[6],
                                      const float* filter_range, int max_points, int max_voxels,
                                      int32_t* voxel_num, void* workspace, size_t workspace_bytes,
                                      int device, void* stream);

/* ---------------------------------------------------------------------------------------------
 * batched hard voxelization with concatenated ("packed") outputs
 * Replaces: the detectors' voxelize(): per-frame voxel_layer + F.pad(coors, (1, 0), value=i) +
 *           torch.cat of voxels / num_points / coors over the batch,
 *           mmdet3d/models/detectors/openpcdet.py:59-76, voxelnet.py:50-67 (and, with mean != 0,
 *           the HardSimpleVFE call that follows, voxel_encoder.py:27-44).
 * points / n: HOST arrays of num_frames device pointers / row counts.  With M_f = voxel_num[f]
 * and S_f = M_0 + ... + M_(f-1), frame f's results are rows [S_f, S_f + M_f) of
 *   voxels_cat      (cap_rows, max_points, c)   mean == 0
 *                   (cap_rows, c)               mean != 0 (per-voxel means, see above)
 *   coors_batch     (cap_rows, 4) int32 = (f, z, y, x); 16-byte aligned
 *   num_points_cat  (cap_rows,)   int32
 * exactly the tensors the reference's torch.cat produces; rows >= S_F are not written.  No host
 * synchronisation: the offsets are formed on the device from voxel_num (device int32[num_frames],
 * also an output).  cap_rows >= sum_f min(n[f], max_voxels), else PCFE_ERR_WORKSPACE.
 * filter_range may be NULL.  Record path only: max_points == 5, c == 4 or 5, 16-byte aligned
 * frames and voxels_cat; PCFE_ERR_SHAPE otherwise (callers then concatenate the outputs of
 * pcfe_hard_voxelize_batch_f32).  Workspace as pcfe_hard_voxelize_batch_f32.
 * ------------------------------------------------------------------------------------------- */
int pcfe_hard_voxelize_packed_batch_f32(const float* const* points, const int64_t* n, int num_frames, int c,
                                        const float voxel_size[3], const float coors_range[6],
                                        const float* filter_range, int max_points, int max_voxels, int mean,
                                        float* voxels_cat, int32_t* coors_batch, int32_t* num_points_cat,
                                        int64_t cap_rows, int32_t* voxel_num, void* workspace,
                                        size_t workspace_bytes, int device, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DynamicScatter: per-voxel max / sum / mean of point features (consumer of dynamic voxelization)
 * Replaces: voxel_layer.dynamic_point_to_voxel_forward / _backward
 *           mmdet3d/ops/voxel/src/scatter_points_cuda.cu:183-303 (kernels :85-181),
 *           voxelization.h:96-123; call sites mmdet3d/ops/voxel/scatter_points.py:28,43.
 * coors (n, ndim) int32, ndim 1..4; a row with any coordinate outside [0, dims[j]) is dropped
 * (the reference drops rows with a negative coordinate, :202; callers pass dims = column maxima
 * + 1).  Voxels are the distinct kept rows in lexicographic order -- what at::unique_dim(sorted)
 * returns after the reference strips its (-1, ...) row -- found with a two-level occupancy bitmap
 * over the dims[0] x ... box instead of a sort (product of dims <= 2^38; the workspace grows with
 * min(n, cells / 256), 64 bytes each, plus cells / 1024 bytes).
 *
 * Two calls, because the number of voxels sizes the outputs (the reference synchronises inside
 * unique_dim for the same reason):
 *   _map_i32:    coors_map (n,) = voxel id of every point or -1; *num_voxels (device int32) = M.
 *   _reduce_f32: voxel_feats (m, c), voxel_coors (m, ndim), point_count (m,) for m = M read back
 *                by the caller.  reduce: PCFE_REDUCE_MAX is order independent (bit-exact,
 *                NaN inputs ignored like fmaxf, a voxel of only NaNs stays -inf);
 *                PCFE_REDUCE_SUM / _MEAN accumulate with float atomicAdd like the reference
 *                (:99) -- association unspecified there and here; _MEAN divides (IEEE) by the
 *                count (:243).
 *   _backward_f32: grad_feats (n, c), every element written (:108-181): SUM copies the voxel's
 *                gradient, MEAN divides it by the count, MAX routes it to the lowest-index point
 *                whose feature equals the maximum (workspace >= m * c * 4 bytes for MAX, else
 *                unused).
 * ------------------------------------------------------------------------------------------- */
enum { PCFE_REDUCE_SUM = 0, PCFE_REDUCE_MEAN = 1, PCFE_REDUCE_MAX = 2 }; /* scatter_points_cuda.cu:7 */

size_t pcfe_dynamic_scatter_workspace_bytes(const int32_t* dims, int ndim, int64_t n);
int pcfe_dynamic_scatter_map_i32(const int32_t* coors, int64_t n, int ndim, const int32_t* dims,
                                 int32_t* coors_map, int32_t* num_voxels, void* workspace,
                                 size_t workspace_bytes, int device, void* stream);
int pcfe_dynamic_scatter_reduce_f32(const float* feats, const int32_t* coors, const int32_t* coors_map,
                                    int64_t n, int c, int ndim, int reduce, int64_t m,
                                    float* voxel_feats, int32_t* voxel_coors, int32_t* point_count,
                                    int device, void* stream);
int pcfe_dynamic_scatter_backward_f32(const float* grad_voxel_feats, const float* feats,
                                      const float* voxel_feats, const int32_t* coors_map,
                                      const int32_t* point_count, int64_t n, int64_t m, int c,
                                      int reduce, float* grad_feats, void* workspace,
                                      size_t workspace_bytes, int device, void* stream);

/* ---------------------------------------------------------------------------------------------
 * points in boxes
 * Replaces: roiaware_pool3d_ext.points_in_boxes_{gpu,batch,cpu}(boxes, points, out)
 *           mmdet3d/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:40-47,126-136,
 *           points_in_boxes_cuda.cu:51-203, points_in_boxes_cpu.cpp:16-69;
 *           call sites mmdet3d/ops/roiaware_pool3d/points_in_boxes.py:46-48,78-80,119-121.
 * boxes (b, t, 7) float32 = (cx, cy, cz_bottom, w, l, h, rz); points (b, m, 3) float32.
 * The inside test is the CPU one (points_in_boxes_cpu.cpp:25-40) bit for bit: glibc-exact
 * cosf/sinf of (float)(rz + pi/2), un-fused rotation, inclusive z slab, strict x/y faces.
 *   _part   -> out (b, m)    int32: lowest box index containing the point, else -1   ("gpu")
 *   _all    -> out (b, m, t) int32 0/1, point-major                                  ("batch")
 *   _boxmajor: boxes (t,7), points (n,3) -> out (t, n) int32 0/1                     ("cpu" layout)
 * Every element of `out` is written (no pre-fill needed).  b, m or t == 0 is a no-op.
 * workspace: >= pcfe_points_in_boxes_workspace_bytes(b, t) bytes, 256-byte aligned (prepared boxes,
 * reject records and, for _part_, a per-frame 64 x 64 grid of box lists: 48 + 128 bytes per box,
 * 16 KB per frame).
 * ------------------------------------------------------------------------------------------- */
size_t pcfe_points_in_boxes_workspace_bytes(int b, int t);

int pcfe_points_in_boxes_part_f32(const float* boxes, const float* points, int b, int t, int64_t m,
                                  int32_t* out, void* workspace, size_t workspace_bytes,
                                  int device, void* stream);
int pcfe_points_in_boxes_all_f32(const float* boxes, const float* points, int b, int t, int64_t m,
                                 int32_t* out, void* workspace, size_t workspace_bytes,
                                 int device, void* stream);
int pcfe_points_in_boxes_boxmajor_f32(const float* boxes, const float* points, int t, int64_t n,
                                      int32_t* out, void* workspace, size_t workspace_bytes,
                                      int device, void* stream);

/* ---------------------------------------------------------------------------------------------
 * RoI-aware point pooling (PartA2's RoI head)
 * Replaces: roiaware_pool3d_ext.forward / backward
 *           mmdet3d/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:49-123 (bindings :127-128),
 *           roiaware_pool3d_kernel.cu:44-361; call sites mmdet3d/ops/roiaware_pool3d/roiaware_pool3d.py:80-82,105-106.
 * rois (boxes_num, 7) = (cx, cy, cz_bottom, w, l, h, rz), pts (pts_num, 3), pts_feature (pts_num, channels), float32.
 * _forward_: pts_idx_of_voxels (boxes_num, out_x, out_y, out_z, max_pts_each_voxel) int32: entry 0 of a
 *   voxel = number of points kept (at most max_pts_each_voxel - 1), entries 1.. = their indices in
 *   ascending point order, the rest UNSPECIFIED (the reference zero-fills; nothing reads them);
 *   pooled_features (boxes_num, out_x, out_y, out_z, channels): max (pool_method 0: first point of the
 *   list with the strictly largest feature; 0 for an empty voxel) or average in list order
 *   (pool_method 1); argmax (same shape, int32, pool_method 0 only): that point's index or -1.
 *   Every element of pooled_features / argmax and every counter is written: no pre-zeroed tensors.
 *   Point -> RoI membership is the CPU inside test of points_in_boxes (above) bit for bit; the voxel
 *   inside the RoI follows roiaware_pool3d_kernel.cu:60-78 in float32.  out_* <= 255 (:72-73).
 * _backward_: grad_in (pts_num, channels) is zeroed and receives grad_out routed to argmax (max) or
 *   grad_out / max(count, 1) to every listed point (average) with float atomicAdd, as in :264-341
 *   (summation order unspecified there and here).
 * Returns 0 / <0 / >0 as everywhere; no workspace, no allocation, no host synchronisation.
 * ------------------------------------------------------------------------------------------- */
int pcfe_roiaware_pool3d_forward_f32(const float* rois, const float* pts, const float* pts_feature,
                                     int boxes_num, int64_t pts_num, int channels, int max_pts_each_voxel,
                                     int out_x, int out_y, int out_z, int pool_method, int32_t* argmax,
                                     int32_t* pts_idx_of_voxels, float* pooled_features, int device,
                                     void* stream);
int pcfe_roiaware_pool3d_backward_f32(const int32_t* pts_idx_of_voxels, const int32_t* argmax,
                                      const float* grad_out, int boxes_num, int out_x, int out_y, int out_z,
                                      int channels, int max_pts_each_voxel, int pool_method, int64_t pts_num,
                                      float* grad_in, int device, void* stream);

/* OpenPCDet variant -- what PV-RCNN's point head calls (pcdet/models/dense_heads/
 * point_head_template.py:82-89).
 * Replaces: roiaware_pool3d_cuda.points_in_boxes_gpu / points_in_boxes_cpu,
 *           thirdparty/Spconv-OpenPCDet/pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:98-118,
 *           143-168 (bindings :175-176), roiaware_pool3d_kernel.cu:16-37,313-336;
 *           call sites pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py:23,39.
 * boxes = (x, y, z_CENTRE, dx, dy, dz, heading); inside <=> |z - cz| <= dz/2 and, after rotating
 * (x - cx, y - cy) by -heading, |lx| < dx/2 + MARGIN and |ly| < dy/2 + MARGIN.  Arithmetic of
 * check_pt_in_box3d_cpu (roiaware_pool3d.cpp:121-140) bit for bit: float32, glibc cosf / sinf, no
 * contraction, right-hand sides in double.  MARGIN is each entry's own: 1e-5 for _gpu_ (the CUDA
 * kernel's, :27 of the .cu), 1e-2 for _cpu_ (:131 of the .cpp).
 * _gpu_: out (b, m) = lowest containing box index or -1, every element written.
 * _cpu_: out (t, n) box-major 0/1.   Workspace: pcfe_points_in_boxes_workspace_bytes(b, t). */
int pcfe_pcdet_points_in_boxes_gpu_f32(const float* boxes, const float* points, int b, int t,
                                       int64_t m, int32_t* out, void* workspace,
                                       size_t workspace_bytes, int device, void* stream);
int pcfe_pcdet_points_in_boxes_cpu_f32(const float* boxes, const float* points, int t, int64_t n,
                                       int32_t* out, void* workspace, size_t workspace_bytes,
                                       int device, void* stream);

/* Test hook: device evaluation of the glibc-exact sinf/cosf used for the boxes.
 * x, s, c are device float arrays of length n. */
int pcfe_debug_sincosf(const float* x, int64_t n, float* s, float* c, int device, void* stream);

/* Test hook: the bin kernel's hoisted-reciprocal cell computation against the plain IEEE divide
 * (voxelization_cpu.cpp:23-29) for EVERY float32 bit pattern of one coordinate, on the grid
 * [lo, hi) with cell size vs on all axes.  out = device uint64[2]: {mismatches, first bad bits}. */
int pcfe_debug_axis_sweep(float lo, float vs, float hi, uint64_t* out, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCFE_H_ */
