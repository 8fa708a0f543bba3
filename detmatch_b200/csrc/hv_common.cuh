// detmatch_b200/csrc/hv_common.cuh -- structures and device helpers shared by the hard
// voxelization paths (hv_bucket.cu: the fast shared-memory path; hv_global.cu: the general
// global-memory path that also serves as the fallback for overflowing frames).
#pragma once

#include "pcfe_common.cuh"

namespace pcfe {

constexpr int kMaxWave = 64;  // frames per launch sequence (size of the kernel-parameter table)

struct HvFrame {
  const float* pts;  // (n, c)
  float* voxels;     // (max_voxels, max_points, c)
  int32_t* coors;    // (max_voxels, 3)
  int32_t* num;      // (max_voxels,)
  int n;
  int pad_;
};

struct HvBatch {
  HvFrame f[kMaxWave];
};

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

#ifdef __CUDACC__
__device__ __forceinline__ void load_xyz(const float* __restrict__ pts, int i, int c, float& x,
                                         float& y, float& z) {
  const float* p = pts + (size_t)i * c;
  x = __ldg(p);
  y = __ldg(p + 1);
  z = __ldg(p + 2);
}

constexpr uint32_t kGold = 0x9E3779B1u;  // multiplicative hash; top bits pick bucket, then slot

// key -> (z, y, x)
__device__ __forceinline__ void decode_key(uint32_t key, const GridParams& g, int32_t* o) {
  const uint32_t plane = (uint32_t)g.gx * (uint32_t)g.gy;
  const uint32_t cz = key / plane;
  const uint32_t rem = key - cz * plane;
  const uint32_t cy = rem / (uint32_t)g.gx;
  const uint32_t cx = rem - cy * (uint32_t)g.gx;
  o[0] = (int32_t)cz;
  o[1] = (int32_t)cy;
  o[2] = (int32_t)cx;
}

// rank of point index m among the "first point of a voxel" flags = voxel id
__device__ __forceinline__ uint32_t first_rank(const uint32_t* __restrict__ bitmask,
                                               const uint32_t* __restrict__ wordprefix, uint32_t m) {
  return wordprefix[m >> 5] + __popc(bitmask[m >> 5] & ((1u << (m & 31)) - 1u));
}

// Sorted P-entry list maintained with atomicMin only.  The list converges to the P smallest
// inserted values in ascending order whatever the interleaving: every entry only ever
// decreases, a value moves on to entry s+1 exactly when a smaller one holds entry s, and a
// displaced value is carried forward by the thread that displaced it.  Works on global memory
// (pass VOLATILE loads through L2) and on shared memory alike.
template <bool GLOBAL>
__device__ __forceinline__ uint32_t list_peek(const uint32_t* p) {
  if (GLOBAL) return __ldcg(p);
  return *reinterpret_cast<const volatile uint32_t*>(p);
}

template <bool GLOBAL>
__device__ __forceinline__ void sorted_insert(uint32_t* lst, const int p, uint32_t v) {
  if (list_peek<GLOBAL>(&lst[p - 1]) < v) return;  // list already full of smaller indices
  int s = 0;
  if (p > 8) {
    // The list is ascending at every instant (an atomicMin at entry s is only issued by a thread
    // that has seen smaller values in all entries before s), so the first entry that can take v
    // is found by bisection; entries skipped hold values < v for good.
    int lo = 0, hi = p - 1;  // lst[hi] >= v was just observed
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (list_peek<GLOBAL>(&lst[mid]) < v) lo = mid + 1;
      else hi = mid;
    }
    s = lo;
  }
  for (; s < p; ++s) {
    if (list_peek<GLOBAL>(&lst[s]) < v) continue;  // monotone: a stale read is only conservative
    const uint32_t old = atomicMin(&lst[s], v);
    if (old == kEmpty) return;
    if (old > v) v = old;
  }
}

// Block-wide exclusive scan of one value per thread (blockDim.x <= 1024, multiple of 32).
// `warp_sums` is a shared array of >= 33 words.  Returns the exclusive prefix; *total gets the
// block total.  Contains two __syncthreads().
__device__ __forceinline__ uint32_t block_exscan(uint32_t v, uint32_t* warp_sums, uint32_t* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    uint32_t w = lane < nw ? warp_sums[lane] : 0u;
    uint32_t wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, d);
      if (lane >= d) wi += t;
    }
    warp_sums[lane] = wi - w;
    if (lane == 31) warp_sums[32] = wi;
  }
  __syncthreads();
  *total = warp_sums[32];
  return warp_sums[wid] + (incl - v);
}
#endif

// ---- global-memory path (hv_global.cu) -------------------------------------------------------
struct HvGlobalPlan {
  GridParams g;
  uint64_t cells;
  int npad, words;
  uint32_t slots;
  int log2_slots, direct;
  size_t table_b, pslot_b, word_b, list_b, per_frame;
};
int hvg_make_plan(int64_t n_max, const float vs[3], const float rg[6], int max_points,
                  int max_voxels, HvGlobalPlan* p);
// Runs frames [0, num_frames) of `frames` through the global-memory path, `wave` frames at a
// time, using `workspace` (>= wave * p.per_frame bytes).
int hvg_run(const pcfe_frame_t* frames, int num_frames, int c, const HvGlobalPlan& p,
            int max_points, int max_voxels, int32_t* voxel_num, void* workspace, int wave,
            cudaStream_t st);

// ---- bucket path (hv_bucket.cu) --------------------------------------------------------------
struct HvBucketPlan {
  GridParams g;
  int npad, words;
  int nb, log2_nb;        // buckets per frame
  int cap;                // entries per bucket region
  int slots, log2_slots;  // shared-memory hash slots per bucket
  int exp_vt;             // voxels per warp tile of the expansion kernel
  // per-frame byte sizes (256-aligned)
  size_t ent_b, lst_b, cells_b, vcell_b, word_b, cnt_b, region_b, per_frame;
  size_t rec_b, firsts_b;  // record-at-first-point variant (P == 5): 32-byte records, voxel id -> first point
  size_t smem_bucket;     // dynamic shared memory of the bucket kernel
  // the overflow fallback (single CTA per frame) reuses the frame's own scratch region
  HvGlobalPlan slow;
};
int hvb_make_plan(int64_t n_max, int c, const float vs[3], const float rg[6], int max_points,
                  int max_voxels, HvBucketPlan* p);
// `nbuf` (1 or 2) wave buffers of wave * p.per_frame bytes each are available in `workspace`;
// with 2 buffers and >= 2 waves consecutive waves overlap on two internal streams.
// mode & kHvMean: frames[i].voxels receives (voxels, c) per-voxel means of the kept points instead
// of (voxels, max_points, c) rows.  mode & kHvPack: every frame carries the SAME output pointers;
// frame f writes at row offset sum(voxel_num[0 .. f)) and its coordinates as (f, z, y, x).
// Both on the record path only (PCFE_ERR_SHAPE otherwise).
constexpr int kHvMean = 1, kHvPack = 2;
int hvb_run(const pcfe_frame_t* frames, int num_frames, int c, const HvBucketPlan& p,
            int max_points, int max_voxels, int32_t* voxel_num, void* workspace, int wave,
            int nbuf, int device, cudaStream_t st, int mode = 0);

extern Knob g_opt_hv_path;         // 0 auto, 1 force global path, 2 force bucket path
extern Knob g_opt_hv_wave;         // frames per wave (0 = automatic)
extern Knob g_opt_force_overflow;  // 1: bucket path treats every frame as overflowed (tests)

}  // namespace pcfe
