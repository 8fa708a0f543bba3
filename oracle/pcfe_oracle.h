/*
 * oracle/pcfe_oracle.h -- CPU restatement of the reference's point-to-cell and
 * point-to-box CPU ops.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product path (detmatch_b200/,
 * include/pcfe.h) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   - the reference's own known-answer tests (the .npz files under tests/golden, see
 *     tests/golden/make_golden.py for how they were produced), and
 *   - the reference's unmodified voxelization_cpu.cpp / points_in_boxes_cpu.cpp
 *     compiled in place into oracle/_ref/ (oracle/build_ref.py).
 *
 * Reference files restated (paths relative to /root/reference):
 *   mmdet3d/ops/voxel/src/voxelization_cpu.cpp:7-41    dynamic_voxelize_kernel
 *   mmdet3d/ops/voxel/src/voxelization_cpu.cpp:43-99   hard_voxelize_kernel
 *   mmdet3d/ops/voxel/src/voxelization_cpu.cpp:105-169 entry points (grid size)
 *   mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cpu.cpp:16-69
 * Third-party arithmetic: host glibc libm sinf/cosf (the reference calls them,
 * so does the oracle); pcfe_oracle_sincosf() restates glibc's algorithm so the
 * device implementation can be validated against it and against the host libm.
 */
#ifndef PCFE_ORACLE_H_
#define PCFE_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* grid[j] = (int)roundf((range[3+j]-range[j])/vs[j]) in float32
 * (voxelization_cpu.cpp:119-122, :155-158). */
void pcfe_oracle_grid_size(const float vs[3], const float range[6], int grid[3]);

/* voxelization_cpu.cpp:7-41 + :144-169.  coors is (n,3) int32, (z,y,x) or
 * (-1,-1,-1).  points is (n,c) row-major float32, c >= 3. */
int pcfe_oracle_dynamic_voxelize_f32(const float* points, int64_t n, int c,
                                     const float vs[3], const float range[6],
                                     int32_t* coors);

/* float64 input: arithmetic is carried out in double because float range /
 * voxel_size promote (voxelization_cpu.cpp:23 with T=double). */
int pcfe_oracle_dynamic_voxelize_f64(const double* points, int64_t n, int c,
                                     const float vs[3], const float range[6],
                                     int32_t* coors);

/* voxelization_cpu.cpp:43-99 + :105-142.  voxels (max_voxels,max_points,c),
 * coors (max_voxels,3), num (max_voxels) must be zero-filled by the caller
 * exactly like voxelize.py:46-50 does.  Returns voxel_num (>= 0) or < 0 on
 * argument error.  max_points == -1 / max_voxels == -1 mean unbounded, as in
 * the reference (:78, :90); the caller must then size the buffers itself. */
int pcfe_oracle_hard_voxelize_f32(const float* points, int64_t n, int c,
                                  const float vs[3], const float range[6],
                                  int max_points, int max_voxels,
                                  float* voxels, int32_t* coors, int32_t* num);

int pcfe_oracle_hard_voxelize_f64(const double* points, int64_t n, int c,
                                  const float vs[3], const float range[6],
                                  int max_points, int max_voxels,
                                  double* voxels, int32_t* coors, int32_t* num);

/* points_in_boxes_cpu.cpp:42-69.  boxes (t,7), points (n,3), out (t,n) int32
 * 0/1, box-major.  Uses the HOST libm cosf/sinf like the reference. */
int pcfe_oracle_points_in_boxes_cpu(const float* boxes, int t,
                                    const float* points, int64_t n,
                                    int32_t* out);

/* Same test but with pcfe_oracle_sincosf() instead of the host libm: what the
 * device computes.  Equal to the function above wherever the restated trig
 * equals the host's (checked exhaustively in tests). */
/* OpenPCDet variant (thirdparty/Spconv-OpenPCDet/pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:
 * 121-168; margin 1e-2 = the CPU op, 1e-5 = the CUDA kernel roiaware_pool3d_kernel.cu:16-37).
 * boxes (t,7) = (x, y, z_centre, dx, dy, dz, heading), out (t,n) 0/1. */
int pcfe_oracle_pcdet_points_in_boxes(const float* boxes, int t, const float* points, int64_t n,
                                      float margin, int32_t* out);

int pcfe_oracle_points_in_boxes_restated(const float* boxes, int t,
                                         const float* points, int64_t n,
                                         int32_t* out);

/* RoI-aware point pooling, roiaware_pool3d_kernel.cu:44-361 restated with the CPU inside test
 * (see pcfe_oracle.c).  pts_idx_of_voxels (N,ox,oy,oz,mp), argmax / pooled (N,ox,oy,oz,C). */
int pcfe_oracle_roiaware_pool3d_forward(const float* rois, int boxes_num, const float* pts, int64_t pts_num,
                                        const float* pts_feature, int channels, int mp, int out_x, int out_y,
                                        int out_z, int pool_method, int32_t* argmax, int32_t* pts_idx_of_voxels,
                                        float* pooled);
int pcfe_oracle_roiaware_pool3d_backward(const int32_t* pts_idx_of_voxels, const int32_t* argmax,
                                         const float* grad_out, int boxes_num, int out_x, int out_y, int out_z,
                                         int channels, int mp, int pool_method, int64_t pts_num, float* grad_in);

/* glibc >= 2.28 sinf/cosf (ARM optimized-routines algorithm), restated.
 * sysdeps/ieee754/flt-32/{s_sinf.c,s_cosf.c,sincosf.h,s_sincosf_data.c}. */
void pcfe_oracle_sincosf(float x, float* sinp, float* cosp);

/* Host libm pass-through, so Python can compare without ctypes-ing libm. */
void pcfe_oracle_host_sincosf(float x, float* sinp, float* cosp);

/* Array forms (host libm / restated) for comparing against the device implementation. */
void pcfe_oracle_host_sincosf_array(const float* x, int64_t n, float* sinp, float* cosp);
void pcfe_oracle_sincosf_array(const float* x, int64_t n, float* sinp, float* cosp);

/* Exhaustive sweep helper: compares restated vs host sinf/cosf for every
 * float whose bit pattern is in [lo_bits, hi_bits) (both signs are the caller's
 * business).  Returns the number of mismatches (sin or cos); first mismatching
 * bit pattern in *first_bad (0 if none). */
int64_t pcfe_oracle_sincosf_sweep(uint32_t lo_bits, uint32_t hi_bits,
                                  uint32_t stride, uint32_t* first_bad);

#ifdef __cplusplus
}
#endif
#endif /* PCFE_ORACLE_H_ */
