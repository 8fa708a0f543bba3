// detmatch_b200/csrc/scatter.cu -- DynamicScatter (points -> per-voxel max / sum / mean) for sm_100a.
//
// Replaces dynamic_point_to_voxel_forward / dynamic_point_to_voxel_backward
// (mmdet3d/ops/voxel/src/scatter_points_cuda.cu:183-303 with the kernels at :85-181; Python
// side mmdet3d/ops/voxel/scatter_points.py:9-99).  See include/pcfe.h for the contract.
//
// The reference finds the voxels with at::unique_dim over the (N, ndim) coordinate rows -- a
// lexicographic sort of N rows -- and then reduces with float atomics.  Here the voxel ids come
// from an occupancy bitmap over the coordinate box instead: one bit per cell in row-major
// (lexicographic) order, set by the points; the rank of a cell's bit among the set bits IS its
// position in the sorted unique list, so a popcount prefix over the bitmap replaces the sort
// (the same device the hard voxelizer uses over point indices).  Traffic: the bitmap (1 bit per
// cell: 11 MB for a 1408 x 1600 x 40 KITTI grid) is written once and read twice; the points are
// read twice (mark, map) plus once for the reduction.
//
// max is order independent and therefore deterministic and bit-exact; sum / mean use float
// atomicAdd exactly like the reference (:99), whose result depends on the arrival order.
#include <algorithm>
#include <cfloat>

#include "hv_common.cuh"
#include "pcfe_common.cuh"

namespace pcfe {
namespace {

constexpr int kDsThreads = 256;
constexpr int kDsScanThreads = 1024;  // words per scan block
constexpr int kMaxDim = 4;

struct DsDims {
  int ndim;
  int d[kMaxDim];
};

// row-major cell index of a coordinate row, or -1 when any coordinate is outside [0, d[j])
// (scatter_points_cuda.cu:202: rows with a negative coordinate are dropped)
__device__ __forceinline__ long long ds_key(const int32_t* __restrict__ row, const DsDims& dm) {
  long long key = 0;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < kMaxDim; ++j) {
    if (j < dm.ndim) {
      const int v = __ldg(row + j);
      ok &= (v >= 0) & (v < dm.d[j]);
      key = key * dm.d[j] + v;
    }
  }
  return ok ? key : -1ll;
}

__global__ void __launch_bounds__(kDsThreads)
ds_mark_kernel(const int32_t* __restrict__ coors, const long long n, const DsDims dm, uint32_t* __restrict__ bitmap) {
  const long long i = (long long)blockIdx.x * kDsThreads + threadIdx.x;
  if (i >= n) return;
  const long long key = ds_key(coors + i * dm.ndim, dm);
  if (key >= 0) atomicOr(&bitmap[key >> 5], 1u << (key & 31));
}

// popcount of every 1024-word block
__global__ void __launch_bounds__(kDsScanThreads)
ds_blocksum_kernel(const uint32_t* __restrict__ bitmap, uint32_t* __restrict__ blocksum) {
  __shared__ uint32_t warp_sums[33];
  const size_t w = (size_t)blockIdx.x * kDsScanThreads + threadIdx.x;
  uint32_t total;
  block_exscan((uint32_t)__popc(bitmap[w]), warp_sums, &total);
  if (threadIdx.x == 0) blocksum[blockIdx.x] = total;
}

// exclusive scan of the block sums in place (one CTA), total -> *num_voxels
__global__ void __launch_bounds__(kDsScanThreads)
ds_scan_blocks_kernel(uint32_t* __restrict__ blocksum, const int nblk, int32_t* __restrict__ num_voxels) {
  __shared__ uint32_t warp_sums[33];
  uint32_t carry = 0;
  for (int b0 = 0; b0 < nblk; b0 += kDsScanThreads) {
    const int b = b0 + threadIdx.x;
    const uint32_t v = b < nblk ? blocksum[b] : 0u;
    uint32_t total;
    const uint32_t ex = block_exscan(v, warp_sums, &total);
    if (b < nblk) blocksum[b] = carry + ex;
    carry += total;
    __syncthreads();  // warp_sums is reused by the next trip
  }
  if (threadIdx.x == 0) *num_voxels = (int32_t)carry;
}

// prefix[w] = number of set bits in the words before w
__global__ void __launch_bounds__(kDsScanThreads)
ds_prefix_kernel(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ blocksum,
                 uint32_t* __restrict__ prefix) {
  __shared__ uint32_t warp_sums[33];
  const size_t w = (size_t)blockIdx.x * kDsScanThreads + threadIdx.x;
  uint32_t total;
  const uint32_t ex = block_exscan((uint32_t)__popc(bitmap[w]), warp_sums, &total);
  prefix[w] = blocksum[blockIdx.x] + ex;
}

__global__ void __launch_bounds__(kDsThreads)
ds_map_kernel(const int32_t* __restrict__ coors, const long long n, const DsDims dm,
              const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ prefix,
              int32_t* __restrict__ coors_map) {
  const long long i = (long long)blockIdx.x * kDsThreads + threadIdx.x;
  if (i >= n) return;
  const long long key = ds_key(coors + i * dm.ndim, dm);
  int32_t vid = -1;
  if (key >= 0) {
    const uint32_t bits = __ldg(&bitmap[key >> 5]);
    vid = (int32_t)(__ldg(&prefix[key >> 5]) + (uint32_t)__popc(bits & ((1u << (key & 31)) - 1u)));
  }
  coors_map[i] = vid;
}

__global__ void __launch_bounds__(kDsThreads)
ds_init_kernel(float* __restrict__ voxel_feats, const long long mc, const float v, int32_t* __restrict__ count,
               const long long m) {
  const long long i = (long long)blockIdx.x * kDsThreads + threadIdx.x;
  if (i < mc) voxel_feats[i] = v;
  if (i < m) count[i] = 0;
}

// fmaxf semantics of the reference's CAS loop (:22-30): NaN inputs are ignored, -0 < +0 is not
// distinguished by fmaxf but the bit patterns order them, which only ever replaces -0 by +0.
__device__ __forceinline__ void atomic_max_float(float* addr, float val) {
  if (val != val) return;
  const uint32_t bits = __float_as_uint(val);
  if (!(bits >> 31)) atomicMax(reinterpret_cast<int*>(addr), (int)bits);
  else atomicMin(reinterpret_cast<unsigned int*>(addr), bits);
}

// thread = one (point, feature) element: coalesced reads of feats, atomics on the voxel row
template <int REDUCE>
__global__ void __launch_bounds__(kDsThreads)
ds_reduce_kernel(const float* __restrict__ feats, const int32_t* __restrict__ coors,
                 const int32_t* __restrict__ coors_map, const long long n, const int c, const int ndim,
                 float* __restrict__ voxel_feats, int32_t* __restrict__ voxel_coors, int32_t* __restrict__ count) {
  const long long e = (long long)blockIdx.x * kDsThreads + threadIdx.x;
  if (e >= n * c) return;
  const long long i = e / c;
  const int j = (int)(e - i * c);
  const int32_t vid = __ldg(coors_map + i);
  if (vid < 0) return;
  const float x = __ldg(feats + e);
  if (REDUCE == PCFE_REDUCE_MAX) atomic_max_float(voxel_feats + (size_t)vid * c + j, x);
  else atomicAdd(voxel_feats + (size_t)vid * c + j, x);
  if (j == 0) {
    atomicAdd(count + vid, 1);
    // every point of a voxel writes the same coordinates
    for (int k = 0; k < ndim; ++k) voxel_coors[(size_t)vid * ndim + k] = __ldg(coors + i * ndim + k);
  }
}

__global__ void __launch_bounds__(kDsThreads)
ds_divide_kernel(float* __restrict__ voxel_feats, const int32_t* __restrict__ count, const long long m, const int c) {
  const long long e = (long long)blockIdx.x * kDsThreads + threadIdx.x;
  if (e >= m * c) return;
  voxel_feats[e] = __fdiv_rn(voxel_feats[e], (float)count[e / c]);  // :243 reduced_feats /= reduce_count
}

// scatter_points_cuda.cu:108-135: every element of grad_feats is written (dropped points get 0)
template <int REDUCE>
__global__ void __launch_bounds__(kDsThreads)
ds_backward_add_kernel(const float* __restrict__ grad_voxel, const int32_t* __restrict__ coors_map,
                       const int32_t* __restrict__ count, const long long n, const int c,
                       float* __restrict__ grad_feats) {
  const long long e = (long long)blockIdx.x * kDsThreads + threadIdx.x;
  if (e >= n * c) return;
  const long long i = e / c;
  const int j = (int)(e - i * c);
  const int32_t vid = __ldg(coors_map + i);
  float g = 0.0f;
  if (vid >= 0) {
    g = __ldg(grad_voxel + (size_t)vid * c + j);
    if (REDUCE == PCFE_REDUCE_MEAN) g = __fdiv_rn(g, (float)__ldg(count + vid));
  }
  grad_feats[e] = g;
}

__global__ void __launch_bounds__(kDsThreads)
ds_fill_i32_kernel(int32_t* __restrict__ p, const long long cnt, const int32_t v) {
  const long long i = (long long)blockIdx.x * kDsThreads + threadIdx.x;
  if (i < cnt) p[i] = v;
}

// :137-163: the lowest point index whose feature equals the voxel's maximum
__global__ void __launch_bounds__(kDsThreads)
ds_backward_max_from_kernel(const float* __restrict__ feats, const float* __restrict__ voxel_feats,
                            const int32_t* __restrict__ coors_map, const long long n, const int c,
                            int32_t* __restrict__ reduce_from, float* __restrict__ grad_feats) {
  const long long e = (long long)blockIdx.x * kDsThreads + threadIdx.x;
  if (e >= n * c) return;
  grad_feats[e] = 0.0f;  // :270 grad_feats.fill_(0)
  const long long i = e / c;
  const int j = (int)(e - i * c);
  const int32_t vid = __ldg(coors_map + i);
  if (vid < 0) return;
  if (__ldg(feats + e) == __ldg(voxel_feats + (size_t)vid * c + j)) atomicMin(reduce_from + (size_t)vid * c + j, (int32_t)i);
}

// :165-181 (a voxel whose maximum matches no point -- every input NaN -- scatters nothing; the
// reference would write one row past the end of grad_feats)
__global__ void __launch_bounds__(kDsThreads)
ds_backward_max_scatter_kernel(const float* __restrict__ grad_voxel, const int32_t* __restrict__ reduce_from,
                               const long long m, const int c, const long long n, float* __restrict__ grad_feats) {
  const long long e = (long long)blockIdx.x * kDsThreads + threadIdx.x;
  if (e >= m * c) return;
  const int j = (int)(e % c);
  const long long src = reduce_from[e];
  if (src < n) grad_feats[src * c + j] = __ldg(grad_voxel + e);
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct DsPlan {
  DsDims dm;
  long long cells;
  size_t words;  // multiple of kDsScanThreads
  int nblk;
  size_t bitmap_b, prefix_b, blocksum_b;
};

int ds_make_plan(const int32_t* dims, int ndim, DsPlan* p) {
  if (!dims) return PCFE_ERR_NULL;
  if (ndim < 1 || ndim > kMaxDim) return PCFE_ERR_SHAPE;
  long long cells = 1;
  p->dm.ndim = ndim;
  for (int j = 0; j < kMaxDim; ++j) p->dm.d[j] = 1;
  for (int j = 0; j < ndim; ++j) {
    if (dims[j] < 0) return PCFE_ERR_SHAPE;
    p->dm.d[j] = dims[j];
    cells *= dims[j];
    if (cells > (1ll << 34)) return PCFE_ERR_TOO_LARGE;  // 2 GiB of bitmap
  }
  p->cells = cells;
  const size_t words = (size_t)((cells + 31) / 32);
  p->words = std::max<size_t>((words + kDsScanThreads - 1) / kDsScanThreads, 1) * kDsScanThreads;
  p->nblk = (int)(p->words / kDsScanThreads);
  p->bitmap_b = align256(p->words * 4);
  p->prefix_b = align256(p->words * 4);
  p->blocksum_b = align256((size_t)p->nblk * 4);
  return PCFE_OK;
}

inline unsigned blocks_for(long long cnt) { return (unsigned)((cnt + kDsThreads - 1) / kDsThreads); }

}  // namespace
}  // namespace pcfe

using namespace pcfe;

extern "C" size_t pcfe_dynamic_scatter_workspace_bytes(const int32_t* dims, int ndim) {
  DsPlan p;
  if (ds_make_plan(dims, ndim, &p) != PCFE_OK) return 0;
  return p.bitmap_b + p.prefix_b + p.blocksum_b;
}

extern "C" int pcfe_dynamic_scatter_map_i32(const int32_t* coors, int64_t n, int ndim, const int32_t* dims,
                                            int32_t* coors_map, int32_t* num_voxels, void* workspace,
                                            size_t workspace_bytes, int device, void* stream) {
  DsPlan p;
  int rc = ds_make_plan(dims, ndim, &p);
  if (rc != PCFE_OK) return rc;
  if (n < 0) return PCFE_ERR_SHAPE;
  if (n >= (1ll << 31)) return PCFE_ERR_TOO_LARGE;
  if (!num_voxels) return PCFE_ERR_NULL;
  if (n > 0 && (!coors || !coors_map)) return PCFE_ERR_NULL;
  if (((uintptr_t)coors & 3) || ((uintptr_t)coors_map & 3) || ((uintptr_t)num_voxels & 3)) return PCFE_ERR_ALIGN;
  DeviceGuard guard(device);
  PCFE_CUDA_TRY(guard.err);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0 || p.cells == 0) {  // nothing can be kept
    PCFE_CUDA_TRY(cudaMemsetAsync(num_voxels, 0, 4, st));
    if (n > 0) {
      ds_fill_i32_kernel<<<blocks_for(n), kDsThreads, 0, st>>>(coors_map, n, -1);
      PCFE_LAUNCH_CHECK();
    }
    return PCFE_OK;
  }
  if (!workspace) return PCFE_ERR_NULL;
  if ((uintptr_t)workspace & 255) return PCFE_ERR_ALIGN;
  if (workspace_bytes < p.bitmap_b + p.prefix_b + p.blocksum_b) return PCFE_ERR_WORKSPACE;
  uint32_t* bitmap = (uint32_t*)workspace;
  uint32_t* prefix = (uint32_t*)((char*)workspace + p.bitmap_b);
  uint32_t* blocksum = (uint32_t*)((char*)workspace + p.bitmap_b + p.prefix_b);
  ProfScope ps("dynamic_scatter_map", st);
  PCFE_CUDA_TRY(cudaMemsetAsync(bitmap, 0, p.words * 4, st));
  ds_mark_kernel<<<blocks_for(n), kDsThreads, 0, st>>>(coors, n, p.dm, bitmap);
  PCFE_LAUNCH_CHECK();
  ds_blocksum_kernel<<<p.nblk, kDsScanThreads, 0, st>>>(bitmap, blocksum);
  PCFE_LAUNCH_CHECK();
  ds_scan_blocks_kernel<<<1, kDsScanThreads, 0, st>>>(blocksum, p.nblk, num_voxels);
  PCFE_LAUNCH_CHECK();
  ds_prefix_kernel<<<p.nblk, kDsScanThreads, 0, st>>>(bitmap, blocksum, prefix);
  PCFE_LAUNCH_CHECK();
  ds_map_kernel<<<blocks_for(n), kDsThreads, 0, st>>>(coors, n, p.dm, bitmap, prefix, coors_map);
  PCFE_LAUNCH_CHECK();
  return PCFE_OK;
}

extern "C" int pcfe_dynamic_scatter_reduce_f32(const float* feats, const int32_t* coors, const int32_t* coors_map,
                                               int64_t n, int c, int ndim, int reduce, int64_t m,
                                               float* voxel_feats, int32_t* voxel_coors, int32_t* point_count,
                                               int device, void* stream) {
  if (n < 0 || m < 0 || c < 1 || ndim < 1 || ndim > kMaxDim) return PCFE_ERR_SHAPE;
  if (reduce != PCFE_REDUCE_SUM && reduce != PCFE_REDUCE_MEAN && reduce != PCFE_REDUCE_MAX) return PCFE_ERR_SHAPE;
  if (n * (int64_t)c >= (1ll << 40)) return PCFE_ERR_TOO_LARGE;
  if (m == 0) return PCFE_OK;
  if (!feats || !coors || !coors_map || !voxel_feats || !voxel_coors || !point_count) return PCFE_ERR_NULL;
  DeviceGuard guard(device);
  PCFE_CUDA_TRY(guard.err);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps("dynamic_scatter_reduce", st);
  ds_init_kernel<<<blocks_for(m * c), kDsThreads, 0, st>>>(voxel_feats, m * c, reduce == PCFE_REDUCE_MAX ? -INFINITY : 0.0f,
                                                          point_count, m);
  PCFE_LAUNCH_CHECK();
  if (n > 0) {
    if (reduce == PCFE_REDUCE_MAX)
      ds_reduce_kernel<PCFE_REDUCE_MAX><<<blocks_for(n * c), kDsThreads, 0, st>>>(feats, coors, coors_map, n, c, ndim,
                                                                                 voxel_feats, voxel_coors, point_count);
    else
      ds_reduce_kernel<PCFE_REDUCE_SUM><<<blocks_for(n * c), kDsThreads, 0, st>>>(feats, coors, coors_map, n, c, ndim,
                                                                                 voxel_feats, voxel_coors, point_count);
    PCFE_LAUNCH_CHECK();
  }
  if (reduce == PCFE_REDUCE_MEAN) {
    ds_divide_kernel<<<blocks_for(m * c), kDsThreads, 0, st>>>(voxel_feats, point_count, m, c);
    PCFE_LAUNCH_CHECK();
  }
  return PCFE_OK;
}

extern "C" int pcfe_dynamic_scatter_backward_f32(const float* grad_voxel_feats, const float* feats,
                                                 const float* voxel_feats, const int32_t* coors_map,
                                                 const int32_t* point_count, int64_t n, int64_t m, int c,
                                                 int reduce, float* grad_feats, void* workspace,
                                                 size_t workspace_bytes, int device, void* stream) {
  if (n < 0 || m < 0 || c < 1) return PCFE_ERR_SHAPE;
  if (reduce != PCFE_REDUCE_SUM && reduce != PCFE_REDUCE_MEAN && reduce != PCFE_REDUCE_MAX) return PCFE_ERR_SHAPE;
  if (n == 0) return PCFE_OK;
  if (!grad_feats || !coors_map) return PCFE_ERR_NULL;
  DeviceGuard guard(device);
  PCFE_CUDA_TRY(guard.err);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps("dynamic_scatter_backward", st);
  if (m == 0) {  // :272 no voxel: the gradient is zero
    PCFE_CUDA_TRY(cudaMemsetAsync(grad_feats, 0, (size_t)n * c * 4, st));
    return PCFE_OK;
  }
  if (!grad_voxel_feats || !point_count) return PCFE_ERR_NULL;
  if (reduce != PCFE_REDUCE_MAX) {
    if (reduce == PCFE_REDUCE_MEAN)
      ds_backward_add_kernel<PCFE_REDUCE_MEAN><<<blocks_for(n * c), kDsThreads, 0, st>>>(grad_voxel_feats, coors_map,
                                                                                        point_count, n, c, grad_feats);
    else
      ds_backward_add_kernel<PCFE_REDUCE_SUM><<<blocks_for(n * c), kDsThreads, 0, st>>>(grad_voxel_feats, coors_map,
                                                                                       point_count, n, c, grad_feats);
    PCFE_LAUNCH_CHECK();
    return PCFE_OK;
  }
  if (!feats || !voxel_feats || !workspace) return PCFE_ERR_NULL;
  if (workspace_bytes < (size_t)m * c * 4) return PCFE_ERR_WORKSPACE;
  int32_t* reduce_from = (int32_t*)workspace;
  ds_fill_i32_kernel<<<blocks_for(m * c), kDsThreads, 0, st>>>(reduce_from, m * c, (int32_t)n);
  PCFE_LAUNCH_CHECK();
  ds_backward_max_from_kernel<<<blocks_for(n * c), kDsThreads, 0, st>>>(feats, voxel_feats, coors_map, n, c, reduce_from,
                                                                      grad_feats);
  PCFE_LAUNCH_CHECK();
  ds_backward_max_scatter_kernel<<<blocks_for(m * c), kDsThreads, 0, st>>>(grad_voxel_feats, reduce_from, m, c, n,
                                                                         grad_feats);
  PCFE_LAUNCH_CHECK();
  return PCFE_OK;
}
