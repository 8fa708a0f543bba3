// detmatch_b200/csrc/roiaware_pool3d.cu -- RoI-aware point pooling (second half).
//
// Replaces roiaware_pool3d_ext.forward / backward:
//   mmdet3d/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:49-123 (bindings),
//   mmdet3d/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:44-361 (mask -> collect -> pool, backward).
// Semantics reproduced: a point belongs to a RoI if the CPU op's inside test holds
// (points_in_boxes_cpu.cpp:16-40 -- the convention of every point-in-box entry of this library; the
// reference's CUDA kernel evaluates the same expressions with device cos / sin and FMA contraction, so
// the two differ only for points within an ulp of a face or of a voxel boundary); its voxel inside
// the RoI is (int((lx + l/2) / (l/out_x)), int((ly + w/2) / (w/out_y)), int((z - z_bottom) / (h/out_z)))
// clamped to the grid (:65-78); a voxel keeps the first max_pts_each_voxel - 1 points in point order
// with their count in entry 0 (:96-118); max pooling takes the first point of the list whose feature is
// the strict maximum, average pooling sums in list order (:121-215).
//
// The reference materialises an (N, npoints) mask with a cudaMalloc per call and collects with ONE
// THREAD PER BOX walking all points.  Here one CTA per RoI streams the points once (conservative xy
// reject first), compacts a chunk's hits in thread order through shared memory and lets one warp append
// them -- __match_any_sync on the voxel gives every hit its rank among the chunk's hits of the same
// voxel -- so the lists come out in point order without any mask array.  Pooling is a warp per
// (RoI, voxel) with lanes over the channels (coalesced feature rows).  Every output element is written:
// no zero-filled tensors are needed.
#include "pcfe_common.cuh"

namespace pcfe {
namespace {

#include "pib_dev.cuh"

constexpr int kRapThreads = 256;

__global__ void __launch_bounds__(kRapThreads)
rap_collect_kernel(const float* __restrict__ rois, const float* __restrict__ pts, const int pts_num,
                   const int out_x, const int out_y, const int out_z, const int mp /* max_pts_each_voxel */,
                   int32_t* __restrict__ pts_idx_of_voxels) {
  __shared__ PBox s_box;
  __shared__ RBox s_rej;
  __shared__ float s_raw[7];
  __shared__ uint32_t s_wcount[kRapThreads / 32];
  __shared__ uint2 s_hits[kRapThreads];  // {voxel, point} of a chunk's hits in point order
  const int box = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nvox = out_x * out_y * out_z;
  volatile int32_t* pv = pts_idx_of_voxels + (size_t)box * nvox * mp;
  if (tid == 0) {
    const PBox q = make_pbox_mmdet(rois + (size_t)box * 7);
    s_box = q;
    s_rej = make_rbox(q);
  }
  if (tid < 7) s_raw[tid] = rois[(size_t)box * 7 + tid];
  for (int v = tid; v < nvox; v += kRapThreads) pv[(size_t)v * mp] = 0;  // entry 0 is the counter
  __syncthreads();
  const PBox b = s_box;
  const RBox rj = s_rej;
  const float zb = s_raw[2], w = s_raw[3], l = s_raw[4], h = s_raw[5];
  // roiaware_pool3d_kernel.cu:65-67: resolutions, float divisions by the (int) grid size
  const float x_res = __fdiv_rn(l, (float)out_x), y_res = __fdiv_rn(w, (float)out_y), z_res = __fdiv_rn(h, (float)out_z);
  const float half_l = __fmul_rn(l, 0.5f), half_w = __fmul_rn(w, 0.5f);  // l / 2, w / 2: exact

  for (int c0 = 0; c0 < pts_num; c0 += kRapThreads) {
    const int i = c0 + tid;
    bool hit = false;
    uint32_t vox = 0;
    if (i < pts_num) {
      const float x = __ldg(pts + (size_t)i * 3), y = __ldg(pts + (size_t)i * 3 + 1), z = __ldg(pts + (size_t)i * 3 + 2);
      if (!xy_reject(x, y, rj) && in_box(x, y, z, b)) {
        hit = true;
        const float sx = __fsub_rn(x, b.cx), sy = __fsub_rn(y, b.cy);
        const float lx = __fadd_rn(__fmul_rn(sx, b.cosa), __fmul_rn(sy, -b.sina));
        const float ly = __fadd_rn(__fmul_rn(sx, b.sina), __fmul_rn(sy, b.cosa));
        const float lz = __fsub_rn(z, zb);
        // :69-75  unsigned idx = int(q); min(max(idx, 0), out - 1) on unsigned values
        const uint32_t xi = (uint32_t)__float2int_rz(__fdiv_rn(__fadd_rn(lx, half_l), x_res));
        const uint32_t yi = (uint32_t)__float2int_rz(__fdiv_rn(__fadd_rn(ly, half_w), y_res));
        const uint32_t zi = (uint32_t)__float2int_rz(__fdiv_rn(lz, z_res));
        const uint32_t xc = min(xi, (uint32_t)(out_x - 1)), yc = min(yi, (uint32_t)(out_y - 1)), zc = min(zi, (uint32_t)(out_z - 1));
        vox = (xc * (uint32_t)out_y + yc) * (uint32_t)out_z + zc;
      }
    }
    const uint32_t hits = __ballot_sync(0xFFFFFFFFu, hit);
    if (lane == 0) s_wcount[warp] = (uint32_t)__popc(hits);
    __syncthreads();
    uint32_t base = 0, total = 0;
#pragma unroll
    for (int k = 0; k < kRapThreads / 32; ++k) {
      const uint32_t cnt = s_wcount[k];
      base += k < warp ? cnt : 0u;
      total += cnt;
    }
    if (total == 0) {  // block-uniform
      __syncthreads();  // s_wcount is rewritten by the next chunk
      continue;
    }
    if (hit) s_hits[base + __popc(hits & ((1u << lane) - 1u))] = make_uint2(vox, (uint32_t)i);
    __syncthreads();
    if (warp == 0) {  // ordered append: hit j before hit j + 1
      for (uint32_t j0 = 0; j0 < total; j0 += 32) {
        const bool act = j0 + lane < total;
        const uint32_t amask = __ballot_sync(0xFFFFFFFFu, act);
        if (act) {
          const uint2 e = s_hits[j0 + lane];
          const uint32_t peers = __match_any_sync(amask, e.x);
          const int rank = __popc(peers & ((1u << lane) - 1u));
          volatile int32_t* vp = pv + (size_t)e.x * mp;
          const int cnt = vp[0];
          __syncwarp(amask);  // every peer has read the count before the leader updates it
          const int pos = cnt + rank;
          if (pos < mp - 1) vp[1 + pos] = (int32_t)e.y;  // :112-115: at most mp - 1 points per voxel
          if (rank == 0) vp[0] = min(cnt + __popc(peers), mp - 1);
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
}

// warp per (RoI, voxel), lanes over channels.  pool_method 0: max (+ argmax), 1: average.
__global__ void __launch_bounds__(kRapThreads)
rap_pool_kernel(const float* __restrict__ pts_feature, const int32_t* __restrict__ pts_idx_of_voxels,
                const long long nvox_total, const int channels, const int mp, const int pool_method,
                float* __restrict__ pooled, int32_t* __restrict__ argmax) {
  const long long v = (long long)blockIdx.x * (kRapThreads / 32) + (threadIdx.x >> 5);
  if (v >= nvox_total) return;
  const int lane = threadIdx.x & 31;
  const int32_t* __restrict__ lst = pts_idx_of_voxels + (size_t)v * mp;
  const int total = __ldg(lst);
  float* __restrict__ po = pooled + (size_t)v * channels;
  int32_t* __restrict__ ao = argmax ? argmax + (size_t)v * channels : nullptr;
  for (int c = lane; c < channels; c += 32) {
    if (pool_method == 0) {
      int arg = -1;
      float mx = __int_as_float(0xFF800000);  // (float)-1e50 = -inf (:152)
      for (int k = 1; k <= total; ++k) {
        const int idx = __ldg(lst + k);
        const float f = __ldg(pts_feature + (size_t)idx * channels + c);
        if (f > mx) {
          mx = f;
          arg = idx;
        }
      }
      po[c] = arg != -1 ? mx : 0.0f;  // the reference leaves its zero-initialised output untouched (:163-165)
      ao[c] = arg;
    } else {
      float sum = 0.0f;
      for (int k = 1; k <= total; ++k)
        sum = __fadd_rn(sum, __ldg(pts_feature + (size_t)__ldg(lst + k) * channels + c));
      po[c] = total > 0 ? __fdiv_rn(sum, (float)total) : 0.0f;  // :207-214
    }
  }
}

// thread per (RoI, voxel, channel): roiaware_pool3d_kernel.cu:264-341
__global__ void __launch_bounds__(kRapThreads)
rap_backward_kernel(const int32_t* __restrict__ pts_idx_of_voxels, const int32_t* __restrict__ argmax,
                    const float* __restrict__ grad_out, const long long elems, const int channels, const int mp,
                    const int pool_method, float* __restrict__ grad_in) {
  const long long e = (long long)blockIdx.x * kRapThreads + threadIdx.x;
  if (e >= elems) return;
  int c;
  const long long v = elem_row(e, channels, c);
  const float g = __ldg(grad_out + e);
  if (pool_method == 0) {
    const int a = __ldg(argmax + e);
    if (a != -1) atomicAdd(grad_in + (size_t)a * channels + c, g);
  } else {
    const int32_t* __restrict__ lst = pts_idx_of_voxels + (size_t)v * mp;
    const int total = __ldg(lst);
    const float cur = __fdiv_rn(1.0f, fmaxf((float)total, 1.0f));
    const float add = __fmul_rn(g, cur);
    for (int k = 1; k <= total; ++k) atomicAdd(grad_in + (size_t)__ldg(lst + k) * channels + c, add);
  }
}

__global__ void rap_fill_kernel(float* p, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = 0.0f;
}

}  // namespace
}  // namespace pcfe

using namespace pcfe;

extern "C" int pcfe_roiaware_pool3d_forward_f32(const float* rois, const float* pts, const float* pts_feature,
                                                int boxes_num, int64_t pts_num, int channels, int max_pts_each_voxel,
                                                int out_x, int out_y, int out_z, int pool_method, int32_t* argmax,
                                                int32_t* pts_idx_of_voxels, float* pooled_features, int device,
                                                void* stream) {
  if (boxes_num < 0 || pts_num < 0 || channels < 0 || max_pts_each_voxel < 1) return PCFE_ERR_SHAPE;
  if (out_x < 1 || out_y < 1 || out_z < 1 || out_x > 255 || out_y > 255 || out_z > 255) return PCFE_ERR_SHAPE;  // :72-73
  if (pool_method != 0 && pool_method != 1) return PCFE_ERR_SHAPE;
  if (pts_num >= 0x7FFFFFFF) return PCFE_ERR_TOO_LARGE;
  if (boxes_num == 0) return PCFE_OK;
  if (!rois || !pts_idx_of_voxels || (pts_num > 0 && !pts)) return PCFE_ERR_NULL;
  if (channels > 0 && (!pooled_features || (pts_num > 0 && !pts_feature) || (pool_method == 0 && !argmax))) return PCFE_ERR_NULL;
  DeviceGuard guard(device);
  PCFE_CUDA_TRY(guard.err);
  cudaStream_t st = (cudaStream_t)stream;
  rap_collect_kernel<<<(unsigned)boxes_num, kRapThreads, 0, st>>>(rois, pts, (int)pts_num, out_x, out_y, out_z,
                                                                  max_pts_each_voxel, pts_idx_of_voxels);
  PCFE_LAUNCH_CHECK();
  if (channels > 0) {
    const long long nv = (long long)boxes_num * out_x * out_y * out_z;
    rap_pool_kernel<<<(unsigned)((nv + kRapThreads / 32 - 1) / (kRapThreads / 32)), kRapThreads, 0, st>>>(
        pts_feature, pts_idx_of_voxels, nv, channels, max_pts_each_voxel, pool_method, pooled_features, argmax);
    PCFE_LAUNCH_CHECK();
  }
  return PCFE_OK;
}

extern "C" int pcfe_roiaware_pool3d_backward_f32(const int32_t* pts_idx_of_voxels, const int32_t* argmax,
                                                 const float* grad_out, int boxes_num, int out_x, int out_y, int out_z,
                                                 int channels, int max_pts_each_voxel, int pool_method, int64_t pts_num,
                                                 float* grad_in, int device, void* stream) {
  if (boxes_num < 0 || channels < 0 || max_pts_each_voxel < 1 || pts_num < 0) return PCFE_ERR_SHAPE;
  if (out_x < 1 || out_y < 1 || out_z < 1) return PCFE_ERR_SHAPE;
  if (pool_method != 0 && pool_method != 1) return PCFE_ERR_SHAPE;
  const long long n_in = (long long)pts_num * channels;
  if (n_in > 0 && !grad_in) return PCFE_ERR_NULL;
  DeviceGuard guard(device);
  PCFE_CUDA_TRY(guard.err);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_in > 0) {  // grad_in = 0 (the reference's wrapper allocates it with new_zeros)
    rap_fill_kernel<<<(unsigned)std::min<long long>((n_in + 255) / 256, 1184), 256, 0, st>>>(grad_in, n_in);
    PCFE_LAUNCH_CHECK();
  }
  const long long elems = (long long)boxes_num * out_x * out_y * out_z * channels;
  if (elems == 0 || n_in == 0) return PCFE_OK;
  if (!grad_out || (pool_method == 0 ? !argmax : !pts_idx_of_voxels)) return PCFE_ERR_NULL;
  rap_backward_kernel<<<(unsigned)((elems + kRapThreads - 1) / kRapThreads), kRapThreads, 0, st>>>(
      pts_idx_of_voxels, argmax, grad_out, elems, channels, max_pts_each_voxel, pool_method, grad_in);
  PCFE_LAUNCH_CHECK();
  return PCFE_OK;
}
