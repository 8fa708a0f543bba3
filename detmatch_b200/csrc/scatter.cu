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
// (the same device the hard voxelizer uses over point indices).  The bitmap has two levels so that
// every pass scales with the points and the occupied blocks, not with the volume of the box: the
// coordinates are read three times (mark level 1, mark level 2, map) plus once for the reduction.
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

// ---- two-level occupancy bitmap -----------------------------------------------------------------
// Level 1: one bit per coarse block of kFine = 256 consecutive cells (cells / 256 bits: 0.7 MB for a
// 16-frame batch of 1408 x 1600 x 40 grids).  Level 2: 256 bits (8 words) for every OCCUPIED coarse
// block only, in coarse order -- the rank of a coarse block among the set level-1 bits is its slot.
// Row-major order of the cells = (coarse order, order inside the block), so the rank of a cell's bit
// in the compact level-2 bitmap is its position in the sorted unique list.  Every pass is
// proportional to the points and the occupied blocks, not to the volume of the coordinate box.
constexpr int kFineLog2 = 8, kFine = 1 << kFineLog2, kFineWords = kFine / 32;

__global__ void __launch_bounds__(kDsThreads)
ds_mark1_kernel(const int32_t* __restrict__ coors, const long long n, const DsDims dm, uint32_t* __restrict__ l1) {
  const long long i = (long long)blockIdx.x * kDsThreads + threadIdx.x;
  if (i >= n) return;
  const long long key = ds_key(coors + i * dm.ndim, dm);
  if (key < 0) return;
  const long long ck = key >> kFineLog2;
  const uint32_t bit = 1u << (ck & 31);
  if (!(__ldcg(&l1[ck >> 5]) & bit)) atomicOr(&l1[ck >> 5], bit);  // most points find their block marked
}

// words the scans of a level cover: host constant, or (level 2) the occupied blocks' words rounded
// up to whole scan blocks, known on the device only
__device__ __forceinline__ size_t ds_limit_words(const size_t words_host, const int32_t* __restrict__ nocc) {
  if (!nocc) return words_host;
  const size_t w = (size_t)(*nocc) * kFineWords;
  return (w + kDsScanThreads - 1) / kDsScanThreads * kDsScanThreads;
}

// popcount of every 1024-word block
__global__ void __launch_bounds__(kDsScanThreads)
ds_blocksum_kernel(const uint32_t* __restrict__ bitmap, uint32_t* __restrict__ blocksum, const size_t words_host,
                   const int32_t* __restrict__ nocc) {
  __shared__ uint32_t warp_sums[33];
  const size_t w = (size_t)blockIdx.x * kDsScanThreads + threadIdx.x;
  if ((size_t)blockIdx.x * kDsScanThreads >= ds_limit_words(words_host, nocc)) return;  // block-uniform
  uint32_t total;
  block_exscan((uint32_t)__popc(bitmap[w]), warp_sums, &total);
  if (threadIdx.x == 0) blocksum[blockIdx.x] = total;
}

// exclusive scan of the block sums in place (one CTA), total -> *total_out
__global__ void __launch_bounds__(kDsScanThreads)
ds_scan_blocks_kernel(uint32_t* __restrict__ blocksum, const size_t words_host, const int32_t* __restrict__ nocc,
                      int32_t* __restrict__ total_out) {
  __shared__ uint32_t warp_sums[33];
  const int nblk = (int)(ds_limit_words(words_host, nocc) / kDsScanThreads);
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
  if (threadIdx.x == 0) *total_out = (int32_t)carry;
}

// prefix[w] = number of set bits in the words before w
__global__ void __launch_bounds__(kDsScanThreads)
ds_prefix_kernel(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ blocksum,
                 uint32_t* __restrict__ prefix, const size_t words_host, const int32_t* __restrict__ nocc) {
  __shared__ uint32_t warp_sums[33];
  const size_t w = (size_t)blockIdx.x * kDsScanThreads + threadIdx.x;
  if ((size_t)blockIdx.x * kDsScanThreads >= ds_limit_words(words_host, nocc)) return;
  uint32_t total;
  const uint32_t ex = block_exscan((uint32_t)__popc(bitmap[w]), warp_sums, &total);
  prefix[w] = blocksum[blockIdx.x] + ex;
}

__global__ void __launch_bounds__(kDsThreads)
ds_zero2_kernel(uint32_t* __restrict__ l2, const int32_t* __restrict__ nocc) {
  const size_t lim = ds_limit_words(0, nocc);
  for (size_t w = (size_t)blockIdx.x * kDsThreads + threadIdx.x; w < lim; w += (size_t)gridDim.x * kDsThreads) l2[w] = 0u;
}

// level-2 word and bit of a cell: slot of its coarse block (rank among the level-1 bits) * 8 words
__device__ __forceinline__ size_t ds_fine_word(const long long key, const uint32_t* __restrict__ l1,
                                               const uint32_t* __restrict__ prefix1) {
  const long long ck = key >> kFineLog2;
  const uint32_t bits = __ldg(&l1[ck >> 5]);
  const uint32_t slot = __ldg(&prefix1[ck >> 5]) + (uint32_t)__popc(bits & ((1u << (ck & 31)) - 1u));
  return (size_t)slot * kFineWords + (size_t)((key & (kFine - 1)) >> 5);
}

__global__ void __launch_bounds__(kDsThreads)
ds_mark2_kernel(const int32_t* __restrict__ coors, const long long n, const DsDims dm, const uint32_t* __restrict__ l1,
                const uint32_t* __restrict__ prefix1, uint32_t* __restrict__ l2) {
  const long long i = (long long)blockIdx.x * kDsThreads + threadIdx.x;
  if (i >= n) return;
  const long long key = ds_key(coors + i * dm.ndim, dm);
  if (key < 0) return;
  const size_t fw = ds_fine_word(key, l1, prefix1);
  const uint32_t bit = 1u << (key & 31);
  if (!(__ldcg(&l2[fw]) & bit)) atomicOr(&l2[fw], bit);
}

__global__ void __launch_bounds__(kDsThreads)
ds_map_kernel(const int32_t* __restrict__ coors, const long long n, const DsDims dm, const uint32_t* __restrict__ l1,
              const uint32_t* __restrict__ prefix1, const uint32_t* __restrict__ l2,
              const uint32_t* __restrict__ prefix2, int32_t* __restrict__ coors_map) {
  const long long i = (long long)blockIdx.x * kDsThreads + threadIdx.x;
  if (i >= n) return;
  const long long key = ds_key(coors + i * dm.ndim, dm);
  int32_t vid = -1;
  if (key >= 0) {
    const size_t fw = ds_fine_word(key, l1, prefix1);
    vid = (int32_t)(__ldg(&prefix2[fw]) + (uint32_t)__popc(__ldg(&l2[fw]) & ((1u << (key & 31)) - 1u)));
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
  int j;
  const long long i = elem_row(e, c, j);
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
  int j;
  voxel_feats[e] = __fdiv_rn(voxel_feats[e], (float)count[elem_row(e, c, j)]);  // :243 reduced_feats /= reduce_count
}

// scatter_points_cuda.cu:108-135: every element of grad_feats is written (dropped points get 0)
template <int REDUCE>
__global__ void __launch_bounds__(kDsThreads)
ds_backward_add_kernel(const float* __restrict__ grad_voxel, const int32_t* __restrict__ coors_map,
                       const int32_t* __restrict__ count, const long long n, const int c,
                       float* __restrict__ grad_feats) {
  const long long e = (long long)blockIdx.x * kDsThreads + threadIdx.x;
  if (e >= n * c) return;
  int j;
  const long long i = elem_row(e, c, j);
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
  int j;
  const long long i = elem_row(e, c, j);
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
  int j;
  elem_row(e, c, j);
  const long long src = reduce_from[e];
  if (src < n) grad_feats[src * c + j] = __ldg(grad_voxel + e);
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct DsPlan {
  DsDims dm;
  long long cells;
  size_t words1;     // level-1 words, multiple of kDsScanThreads
  long long cap_occ; // occupied coarse blocks at most: min(n, coarse blocks)
  size_t words2;     // level-2 words at most, multiple of kDsScanThreads
  size_t l1_b, p1_b, s1_b, l2_b, p2_b, s2_b, ctl_b;
  size_t total() const { return l1_b + p1_b + s1_b + l2_b + p2_b + s2_b + ctl_b; }
};

inline size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

int ds_make_plan(const int32_t* dims, int ndim, long long n, DsPlan* p) {
  if (!dims) return PCFE_ERR_NULL;
  if (ndim < 1 || ndim > kMaxDim || n < 0) return PCFE_ERR_SHAPE;
  long long cells = 1;
  p->dm.ndim = ndim;
  for (int j = 0; j < kMaxDim; ++j) p->dm.d[j] = 1;
  for (int j = 0; j < ndim; ++j) {
    if (dims[j] < 0) return PCFE_ERR_SHAPE;
    p->dm.d[j] = dims[j];
    cells *= dims[j];
    if (cells > (1ll << 38)) return PCFE_ERR_TOO_LARGE;  // 128 MiB of level-1 bitmap
  }
  p->cells = cells;
  const long long ncoarse = (cells + kFine - 1) / kFine;
  p->words1 = round_up(std::max<size_t>((size_t)((ncoarse + 31) / 32), 1), kDsScanThreads);
  p->cap_occ = std::max<long long>(std::min<long long>(n, ncoarse), 1);
  p->words2 = round_up((size_t)p->cap_occ * kFineWords, kDsScanThreads);
  p->l1_b = align256(p->words1 * 4);
  p->p1_b = align256(p->words1 * 4);
  p->s1_b = align256(p->words1 / kDsScanThreads * 4);
  p->l2_b = align256(p->words2 * 4);
  p->p2_b = align256(p->words2 * 4);
  p->s2_b = align256(p->words2 / kDsScanThreads * 4);
  p->ctl_b = 256;
  return PCFE_OK;
}

inline unsigned blocks_for(long long cnt) { return (unsigned)((cnt + kDsThreads - 1) / kDsThreads); }

}  // namespace
}  // namespace pcfe

using namespace pcfe;

extern "C" size_t pcfe_dynamic_scatter_workspace_bytes(const int32_t* dims, int ndim, int64_t n) {
  DsPlan p;
  if (ds_make_plan(dims, ndim, n, &p) != PCFE_OK) return 0;
  return p.total();
}

extern "C" int pcfe_dynamic_scatter_map_i32(const int32_t* coors, int64_t n, int ndim, const int32_t* dims,
                                            int32_t* coors_map, int32_t* num_voxels, void* workspace,
                                            size_t workspace_bytes, int device, void* stream) {
  DsPlan p;
  int rc = ds_make_plan(dims, ndim, n, &p);
  if (rc != PCFE_OK) return rc;
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
  if (workspace_bytes < p.total()) return PCFE_ERR_WORKSPACE;
  char* base = (char*)workspace;
  uint32_t* l1 = (uint32_t*)base;
  uint32_t* prefix1 = (uint32_t*)(base + p.l1_b);
  uint32_t* sum1 = (uint32_t*)(base + p.l1_b + p.p1_b);
  uint32_t* l2 = (uint32_t*)(base + p.l1_b + p.p1_b + p.s1_b);
  uint32_t* prefix2 = (uint32_t*)(base + p.l1_b + p.p1_b + p.s1_b + p.l2_b);
  uint32_t* sum2 = (uint32_t*)(base + p.l1_b + p.p1_b + p.s1_b + p.l2_b + p.p2_b);
  int32_t* nocc = (int32_t*)(base + p.l1_b + p.p1_b + p.s1_b + p.l2_b + p.p2_b + p.s2_b);
  const unsigned nblk1 = (unsigned)(p.words1 / kDsScanThreads), nblk2 = (unsigned)(p.words2 / kDsScanThreads);
  ProfScope ps("dynamic_scatter_map", st);
  // level 1: occupied coarse blocks and their slots
  PCFE_CUDA_TRY(cudaMemsetAsync(l1, 0, p.words1 * 4, st));
  ds_mark1_kernel<<<blocks_for(n), kDsThreads, 0, st>>>(coors, n, p.dm, l1);
  PCFE_LAUNCH_CHECK();
  ds_blocksum_kernel<<<nblk1, kDsScanThreads, 0, st>>>(l1, sum1, p.words1, nullptr);
  PCFE_LAUNCH_CHECK();
  ds_scan_blocks_kernel<<<1, kDsScanThreads, 0, st>>>(sum1, p.words1, nullptr, nocc);
  PCFE_LAUNCH_CHECK();
  ds_prefix_kernel<<<nblk1, kDsScanThreads, 0, st>>>(l1, sum1, prefix1, p.words1, nullptr);
  PCFE_LAUNCH_CHECK();
  // level 2: the cells of the occupied blocks (sizes known on the device only: the grids cover the
  // worst case and the blocks beyond *nocc return at once)
  ds_zero2_kernel<<<std::min<unsigned>((unsigned)((p.words2 + kDsThreads - 1) / kDsThreads), 148u * 16u), kDsThreads, 0, st>>>(l2, nocc);
  PCFE_LAUNCH_CHECK();
  ds_mark2_kernel<<<blocks_for(n), kDsThreads, 0, st>>>(coors, n, p.dm, l1, prefix1, l2);
  PCFE_LAUNCH_CHECK();
  ds_blocksum_kernel<<<nblk2, kDsScanThreads, 0, st>>>(l2, sum2, 0, nocc);
  PCFE_LAUNCH_CHECK();
  ds_scan_blocks_kernel<<<1, kDsScanThreads, 0, st>>>(sum2, 0, nocc, num_voxels);
  PCFE_LAUNCH_CHECK();
  ds_prefix_kernel<<<nblk2, kDsScanThreads, 0, st>>>(l2, sum2, prefix2, 0, nocc);
  PCFE_LAUNCH_CHECK();
  ds_map_kernel<<<blocks_for(n), kDsThreads, 0, st>>>(coors, n, p.dm, l1, prefix1, l2, prefix2, coors_map);
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
