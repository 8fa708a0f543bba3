// detmatch_b200/csrc/hv_global.cu -- hard voxelization, general global-memory path.
//
// Works for any (N, C, P, V, grid): the per-frame cell table lives in global memory (L2) and is
// updated with L2 atomics.  It is the path of last resort: the bucket path (hv_bucket.cu) is
// ~10x faster because shared-memory atomics are (measured: 1500 vs 135 Gops/s on B200), and
// falls back to the single-CTA variant at the bottom of this file for frames whose buckets
// overflow.  Behaviour reproduced: mmdet3d/ops/voxel/src/voxelization_cpu.cpp:43-99.
//
//   K1 hvg_key_hash   point -> cell key -> table slot (direct-mapped for small grids, open
//                     addressing otherwise); atomicMin(first point index) per cell.
//   K2 hvg_first_flag a point is "first" iff it is its cell's minimum index; one ballot per warp
//                     gives a bitmask over point indices.
//   K3 hv_scan_flags  per frame: exclusive popcount prefix of the bitmask -> voxel id of every
//                     first point (first-occurrence order); voxel_num = min(#first, V).
//   K4 hvg_assign     points of kept voxels insert their index into the voxel's sorted P-entry
//                     list (atomicMin chain); the first point writes coors.
//   K5 hvg_scatter    voxel-id-ordered write of complete rows (data + zero padding) and counts.
#include <algorithm>

#include "hv_common.cuh"

namespace pcfe {
namespace {

constexpr int kThreads = 256;
constexpr int kPtsPerThread = 4;
constexpr int kTilePts = kThreads * kPtsPerThread;

struct HvgWork {           // per-wave scratch, frame-major with the strides below (elements)
  uint2* table;            // [W][S]      {key, min point index}
  int32_t* pslot;          // [W][npad]   slot of every point, -1 when out of range
  uint32_t* bitmask;       // [W][words]  bit i set <=> point i is the first of its voxel
  uint32_t* wordprefix;    // [W][words]  exclusive popcount prefix of bitmask
  uint32_t* idxlist;       // [W][V*P]    per voxel id: ascending point indices, kEmpty padded
  size_t table_stride, pslot_stride, word_stride, list_stride;
  uint32_t slots;          // S
  int log2_slots;          // hash mode: S == 1 << log2_slots
  int direct;              // 1: slot == key (S == number of cells)
};

__device__ __forceinline__ uint32_t table_find_or_claim(uint2* __restrict__ table, uint32_t key,
                                                        uint32_t slots, int log2_slots, int direct) {
  if (direct) return key;
  uint32_t s = (key * kGold) >> (32 - log2_slots);
  const uint32_t mask = slots - 1u;
  while (true) {
    uint32_t cur = __ldcg(&table[s].x);
    if (cur == key) return s;
    if (cur == kEmpty) {
      cur = atomicCAS(&table[s].x, kEmpty, key);
      if (cur == kEmpty || cur == key) return s;
    }
    s = (s + 1u) & mask;
  }
}

__global__ void __launch_bounds__(kThreads)
hvg_key_hash_kernel(const __grid_constant__ HvBatch batch, const HvgWork w, const GridParams g,
                    const int c) {
  const int f = blockIdx.y;
  const HvFrame& fr = batch.f[f];
  uint2* __restrict__ table = w.table + (size_t)f * w.table_stride;
  int32_t* __restrict__ pslot = w.pslot + (size_t)f * w.pslot_stride;
  const int base = blockIdx.x * kTilePts + threadIdx.x;

  uint32_t key[kPtsPerThread];
#pragma unroll
  for (int k = 0; k < kPtsPerThread; ++k) {
    const int i = base + k * kThreads;
    key[k] = kEmpty;
    if (i < fr.n) {
      float x, y, z;
      load_xyz(fr.pts, i, c, x, y, z);
      int cx, cy, cz;
      key[k] = point_key(x, y, z, g, cx, cy, cz);
    }
  }
#pragma unroll
  for (int k = 0; k < kPtsPerThread; ++k) {
    const int i = base + k * kThreads;
    if (i >= fr.n) break;
    int32_t slot = -1;
    if (key[k] != kEmpty) {
      const uint32_t s = table_find_or_claim(table, key[k], w.slots, w.log2_slots, w.direct);
      // first point index of the cell; a (possibly stale) smaller value means i cannot win
      if (__ldcg(&table[s].y) > (uint32_t)i) atomicMin(&table[s].y, (uint32_t)i);
      slot = (int32_t)s;
    }
    pslot[i] = slot;
  }
}

__global__ void __launch_bounds__(kThreads)
hvg_first_flag_kernel(const __grid_constant__ HvBatch batch, const HvgWork w, const int npad) {
  const int f = blockIdx.y;
  const int n = batch.f[f].n;
  const uint2* __restrict__ table = w.table + (size_t)f * w.table_stride;
  const int32_t* __restrict__ pslot = w.pslot + (size_t)f * w.pslot_stride;
  uint32_t* __restrict__ bitmask = w.bitmask + (size_t)f * w.word_stride;
  const int base = blockIdx.x * kTilePts + threadIdx.x;
#pragma unroll
  for (int k = 0; k < kPtsPerThread; ++k) {
    const int i = base + k * kThreads;  // i < npad is warp-uniform (npad % 32 == 0)
    if (i >= npad) break;
    bool first = false;
    if (i < n) {
      const int32_t s = pslot[i];
      if (s >= 0) first = (__ldcg(&table[s].y) == (uint32_t)i);
    }
    const uint32_t word = __ballot_sync(0xFFFFFFFFu, first);
    if ((threadIdx.x & 31) == 0) bitmask[i >> 5] = word;
  }
}

__global__ void __launch_bounds__(kThreads)
hvg_assign_kernel(const __grid_constant__ HvBatch batch, const HvgWork w, const GridParams g,
                  const int max_points, const int max_voxels) {
  const int f = blockIdx.y;
  const HvFrame& fr = batch.f[f];
  const uint2* __restrict__ table = w.table + (size_t)f * w.table_stride;
  const int32_t* __restrict__ pslot = w.pslot + (size_t)f * w.pslot_stride;
  const uint32_t* __restrict__ bitmask = w.bitmask + (size_t)f * w.word_stride;
  const uint32_t* __restrict__ wordprefix = w.wordprefix + (size_t)f * w.word_stride;
  uint32_t* __restrict__ idxlist = w.idxlist + (size_t)f * w.list_stride;
  const int base = blockIdx.x * kTilePts + threadIdx.x;
#pragma unroll
  for (int k = 0; k < kPtsPerThread; ++k) {
    const int i = base + k * kThreads;
    if (i >= fr.n) break;
    const int32_t s = pslot[i];
    if (s < 0) continue;
    const uint2 e = __ldcg(&table[s]);
    const uint32_t m = e.y;  // first point of this voxel
    const uint32_t vid = first_rank(bitmask, wordprefix, m);
    if (vid >= (uint32_t)max_voxels) continue;  // voxelization_cpu.cpp:78
    if (m == (uint32_t)i) decode_key(w.direct ? (uint32_t)s : e.x, g, fr.coors + (size_t)vid * 3);
    if (max_points > 0)
      sorted_insert<true>(idxlist + (size_t)vid * max_points, max_points, (uint32_t)i);
  }
}

template <int C>
__global__ void __launch_bounds__(kThreads)
hvg_scatter_kernel(const __grid_constant__ HvBatch batch, const HvgWork w, const int c_rt,
                   const int max_points, const int32_t* __restrict__ voxel_num) {
  const int f = blockIdx.y;
  const HvFrame& fr = batch.f[f];
  const int c = C > 0 ? C : c_rt;
  const uint32_t* __restrict__ idxlist = w.idxlist + (size_t)f * w.list_stride;
  const int m = voxel_num[f];
  const long long rows = (long long)m * max_points;
  const long long r0 = (long long)blockIdx.x * kTilePts + threadIdx.x;
#pragma unroll
  for (int k = 0; k < kPtsPerThread; ++k) {
    const long long r = r0 + (long long)k * kThreads;
    if (r >= rows) break;
    const uint32_t idx = __ldcg(&idxlist[r]);
    float* __restrict__ dst = fr.voxels + (size_t)r * c;
    if (C == 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx != kEmpty) v = __ldg(reinterpret_cast<const float4*>(fr.pts + (size_t)idx * 4));
      *reinterpret_cast<float4*>(dst) = v;
    } else {
      const float* __restrict__ src = fr.pts + (size_t)idx * c;
      for (int j = 0; j < c; ++j) dst[j] = (idx != kEmpty) ? __ldg(src + j) : 0.0f;
    }
  }
  // num_points_per_voxel: number of occupied list entries (ascending, kEmpty last)
#pragma unroll
  for (int k = 0; k < kPtsPerThread; ++k) {
    const long long v = r0 + (long long)k * kThreads;
    if (v >= m) break;
    const uint32_t* lst = idxlist + (size_t)v * max_points;
    int cnt = 0;
    for (int s = 0; s < max_points; ++s) cnt += (__ldcg(&lst[s]) != kEmpty) ? 1 : 0;
    fr.num[v] = cnt;
  }
}

template <int C>
void launch_scatter(dim3 grid, cudaStream_t st, const HvBatch& b, const HvgWork& w, int c, int p,
                    const int32_t* vn) {
  hvg_scatter_kernel<C><<<grid, kThreads, 0, st>>>(b, w, c, p, vn);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// K3 (shared with the bucket path): per-frame exclusive popcount prefix over the bitmask words
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads)
hv_scan_flags_kernel(const uint32_t* __restrict__ bitmask_base, const size_t bitmask_stride,
                     uint32_t* __restrict__ prefix_base, const size_t prefix_stride,
                     const int words, const int max_voxels, int32_t* __restrict__ voxel_num,
                     const int paired /* 1: prefix array holds {bitmask word, prefix} pairs */) {
  const int f = blockIdx.x;
  const uint32_t* __restrict__ bitmask = bitmask_base + (size_t)f * bitmask_stride;
  uint32_t* __restrict__ wordprefix = prefix_base + (size_t)f * prefix_stride;
  __shared__ uint32_t warp_sums[33];
  // words handled per thread are consecutive, so this is a plain blocked scan
  const int per = (words + kScanThreads - 1) / kScanThreads;
  const int w0 = threadIdx.x * per;
  uint32_t local = 0;
  for (int j = 0; j < per; ++j) {
    const int idx = w0 + j;
    if (idx < words) local += __popc(bitmask[idx]);
  }
  uint32_t total;
  uint32_t run = block_exscan(local, warp_sums, &total);
  for (int j = 0; j < per; ++j) {
    const int idx = w0 + j;
    if (idx < words) {
      const uint32_t bits = bitmask[idx];
      if (paired) reinterpret_cast<uint2*>(wordprefix)[idx] = make_uint2(bits, run);
      else wordprefix[idx] = run;
      run += __popc(bits);
    }
  }
  if (threadIdx.x == 0) voxel_num[f] = (int32_t)min(total, (uint32_t)max_voxels);
}

int hv_launch_scan(const uint32_t* bitmask, size_t bitmask_stride, uint32_t* prefix,
                   size_t prefix_stride, int words, int max_voxels, int32_t* voxel_num, int frames,
                   int paired, cudaStream_t st) {
  ProfScope ps("hv_scan_flags", st);
  hv_scan_flags_kernel<<<frames, kScanThreads, 0, st>>>(bitmask, bitmask_stride, prefix, prefix_stride,
                                                       words, max_voxels, voxel_num, paired);
  PCFE_LAUNCH_CHECK();
  return PCFE_OK;
}

// ------------------------------------------------------------------------------------------
// Fallback for frames the bucket path flags as overflowed: ONE CTA voxelizes a whole frame with
// the K1..K5 logic above, phases separated by __syncthreads().  Slow (one SM per frame) but
// exact for any input; unflagged frames exit immediately, so the launch costs ~2 us per wave.
// ------------------------------------------------------------------------------------------
constexpr int kSlowThreads = 1024;

// stage 0: everything; stage 1: up to the per-voxel point lists and voxel_num; stage 2: the outputs
// only (scratch, mask and prefix of stage 1 are still in place: one scratch region per frame).  The
// split lets a packed batch (vn_all != nullptr: rows at offset sum(voxel_num of the batch's earlier
// frames), coordinates as (batch index, z, y, x)) learn every frame's voxel count before any
// output row is placed.
__global__ void __launch_bounds__(kSlowThreads)
hvg_slow_frame_kernel(const __grid_constant__ HvBatch batch, uint32_t* __restrict__ overflow,
                      const size_t overflow_stride, const int force, char* __restrict__ scratch_base,
                      const size_t scratch_stride, const HvGlobalPlan p,
                      uint32_t* __restrict__ bitmask_base, const size_t bitmask_stride,
                      uint32_t* __restrict__ prefix_base, const size_t prefix_stride, const int c,
                      const int max_points, const int max_voxels, int32_t* __restrict__ voxel_num,
                      const int frames, const int ring, const int mean, const int stage,
                      const int32_t* __restrict__ vn_all, const int f_first) {
  // CTA r serves frames r, r + ring, ... one after the other in scratch region r (ring == frames:
  // one frame per CTA)
  __shared__ uint32_t warp_sums[33];
  __shared__ uint32_t carry;
  pdl_wait();  // launched as a programmatic dependent of the kernel before it
  for (int f = blockIdx.x; f < frames; f += ring) {
  if (!force && overflow[(size_t)f * overflow_stride] == 0) continue;
  const HvFrame& fr = batch.f[f];
  const int n = fr.n;
  const int tid = threadIdx.x;
  char* scratch = scratch_base + (size_t)blockIdx.x * scratch_stride;
  uint2* table = reinterpret_cast<uint2*>(scratch);
  uint32_t* idxlist = reinterpret_cast<uint32_t*>(scratch + p.table_b);
  int32_t* pslot = reinterpret_cast<int32_t*>(scratch + p.table_b + p.list_b);
  uint32_t* bitmask = bitmask_base + (size_t)f * bitmask_stride;
  uint32_t* wordprefix = prefix_base + (size_t)f * prefix_stride;
  int m;
  if (stage != 2) {
  // phase 0: scratch init (table + lists are contiguous)
  {
    uint32_t* w = reinterpret_cast<uint32_t*>(scratch);
    const size_t nw = (p.table_b + p.list_b) / 4;
    for (size_t i = tid; i < nw; i += kSlowThreads) w[i] = kEmpty;
  }
  __syncthreads();
  // phase 1: keys + table
  for (int i = tid; i < n; i += kSlowThreads) {
    float x, y, z;
    load_xyz(fr.pts, i, c, x, y, z);
    int cx, cy, cz;
    const uint32_t key = point_key(x, y, z, p.g, cx, cy, cz);
    int32_t slot = -1;
    if (key != kEmpty) {
      const uint32_t s = table_find_or_claim(table, key, p.slots, p.log2_slots, p.direct);
      atomicMin(&table[s].y, (uint32_t)i);
      slot = (int32_t)s;
    }
    pslot[i] = slot;
  }
  __syncthreads();
  // phase 2: first flags
  const int npad = (n + 31) & ~31;
  for (int i = tid; i < npad; i += kSlowThreads) {
    bool first = false;
    if (i < n) {
      const int32_t s = pslot[i];
      if (s >= 0) first = (__ldcg(&table[s].y) == (uint32_t)i);
    }
    const uint32_t word = __ballot_sync(0xFFFFFFFFu, first);
    if ((tid & 31) == 0) bitmask[i >> 5] = word;
  }
  __syncthreads();
  // phase 3: popcount prefix, chunk by chunk
  const int words = npad >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int w0 = 0; w0 < words; w0 += kSlowThreads) {
    const int idx = w0 + tid;
    const uint32_t v = idx < words ? (uint32_t)__popc(__ldcg(&bitmask[idx])) : 0u;
    uint32_t total;
    const uint32_t ex = block_exscan(v, warp_sums, &total);
    const uint32_t base = carry;
    if (idx < words) wordprefix[idx] = base + ex;
    __syncthreads();
    if (tid == 0) carry = base + total;
    __syncthreads();
  }
  m = (int)min(carry, (uint32_t)max_voxels);
  if (tid == 0) {
    voxel_num[f] = m;
    // a forced frame (test knob) is flagged like a real overflow, so that the expansion kernel
    // that runs between the two stages leaves it alone
    if (stage == 1 && force) overflow[(size_t)f * overflow_stride] = 1u;
  }
  // phase 4: lists
  if (max_points > 0)
    for (int i = tid; i < n; i += kSlowThreads) {
      const int32_t s = pslot[i];
      if (s < 0) continue;
      const uint32_t vid = first_rank(bitmask, wordprefix, __ldcg(&table[s].y));
      if (vid >= (uint32_t)max_voxels) continue;
      sorted_insert<true>(idxlist + (size_t)vid * max_points, max_points, (uint32_t)i);
    }
  __syncthreads();  // (also: the next frame reuses `carry`)
  if (stage == 1) continue;
  } else {
    m = voxel_num[f];
  }
  // phase 5: coordinates, rows (mean: the per-voxel mean of the kept points instead -- slot-order
  // sum over all max_points slots, absent ones being +0, then an IEEE divide) and counts
  size_t off = 0;
  if (vn_all)
    for (int k = 0; k < f_first + f; ++k) off += (size_t)vn_all[k];
  for (int i = tid; i < n; i += kSlowThreads) {
    const int32_t s = pslot[i];
    if (s < 0) continue;
    const uint2 e = __ldcg(&table[s]);
    if (e.y != (uint32_t)i) continue;
    const uint32_t vid = first_rank(bitmask, wordprefix, e.y);
    if (vid >= (uint32_t)max_voxels) continue;
    int32_t zyx[3];
    decode_key(p.direct ? (uint32_t)s : e.x, p.g, zyx);
    if (vn_all) {
      int32_t* o = fr.coors + (off + vid) * 4;
      o[0] = f_first + f; o[1] = zyx[0]; o[2] = zyx[1]; o[3] = zyx[2];
    } else {
      int32_t* o = fr.coors + (size_t)vid * 3;
      o[0] = zyx[0]; o[1] = zyx[1]; o[2] = zyx[2];
    }
  }
  const long long rows = (long long)m * max_points;
  if (mean) {
    float* out = fr.voxels + off * c;
    for (long long e = tid; e < (long long)m * c; e += kSlowThreads) {
      int j;
      const long long v = elem_row(e, c, j);
      const uint32_t* lst = idxlist + (size_t)v * max_points;
      float a = 0.0f;
      int cnt = 0;
      for (int sl = 0; sl < max_points; ++sl) {
        const uint32_t idx = __ldcg(&lst[sl]);
        const float x = (idx != kEmpty) ? __ldg(fr.pts + (size_t)idx * c + j) : 0.0f;
        a = sl == 0 ? x : __fadd_rn(a, x);
        cnt += (idx != kEmpty) ? 1 : 0;
      }
      out[e] = __fdiv_rn(a, (float)cnt);
    }
  } else {
    float* out = fr.voxels + off * (size_t)max_points * c;
    for (long long r = tid; r < rows; r += kSlowThreads) {
      const uint32_t idx = __ldcg(&idxlist[r]);
      float* dst = out + (size_t)r * c;
      const float* src = fr.pts + (size_t)idx * c;
      for (int j = 0; j < c; ++j) dst[j] = (idx != kEmpty) ? __ldg(src + j) : 0.0f;
    }
  }
  for (int v = tid; v < m; v += kSlowThreads) {
    const uint32_t* lst = idxlist + (size_t)v * max_points;
    int cnt = 0;
    for (int s = 0; s < max_points; ++s) cnt += (__ldcg(&lst[s]) != kEmpty) ? 1 : 0;
    fr.num[off + v] = cnt;
  }
  __syncthreads();  // the next frame reuses the scratch region and `carry`
  }
}

// mode: bit 0 = mean epilogue; stage / vn_all / f_first: see the kernel
int hvg_launch_slow(const HvBatch& b, int frames, uint32_t* overflow, size_t overflow_stride,
                    int force, char* scratch_base, size_t scratch_stride, const HvGlobalPlan& p,
                    uint32_t* bitmask, size_t bitmask_stride, uint32_t* prefix, size_t prefix_stride,
                    int c, int max_points, int max_voxels, int32_t* voxel_num, cudaStream_t st, int mean,
                    int stage, const int32_t* vn_all, int f_first) {
  ProfScope ps("hv_slow_fallback", st);
  PCFE_CUDA_TRY(launch_pdl(hvg_slow_frame_kernel, dim3((unsigned)frames), dim3(kSlowThreads), 0, st, true, b, overflow,
                           overflow_stride, force, scratch_base, scratch_stride, p, bitmask, bitmask_stride,
                           prefix, prefix_stride, c, max_points, max_voxels, voxel_num, frames, frames, mean,
                           stage, vn_all, f_first));
  PCFE_LAUNCH_CHECK();
  return PCFE_OK;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int hvg_make_plan(int64_t n_max, const float vs[3], const float rg[6], int max_points,
                  int max_voxels, HvGlobalPlan* p) {
  if (!vs || !rg) return PCFE_ERR_NULL;
  if (n_max < 0) return PCFE_ERR_SHAPE;
  if (n_max >= 0x7FFFFFFF - 4096) return PCFE_ERR_TOO_LARGE;
  if (max_points < 0 || max_voxels < 0) return PCFE_ERR_CAPS;
  make_grid_params(vs, rg, &p->g);
  if (p->g.gx <= 0 || p->g.gy <= 0 || p->g.gz <= 0) return PCFE_ERR_GRID;
  p->cells = (uint64_t)p->g.gx * (uint64_t)p->g.gy * (uint64_t)p->g.gz;
  if (p->cells >= 0xFFFFFFFFull) return PCFE_ERR_GRID;
  p->npad = (int)((n_max + 31) / 32 * 32);
  if (p->npad == 0) p->npad = 32;
  p->words = p->npad / 32;
  // Small grids (pillars) are direct-mapped; large ones hashed at load factor <= 0.75.
  if (p->cells <= 4ull * (uint64_t)p->npad) {
    p->direct = 1;
    p->slots = (uint32_t)p->cells;
    p->log2_slots = 0;
  } else {
    p->direct = 0;
    const uint64_t want = std::max<uint64_t>(1024, ((uint64_t)p->npad * 4 + 2) / 3);
    int lg = 10;
    while ((1ull << lg) < want) ++lg;
    p->log2_slots = lg;
    p->slots = 1u << lg;
  }
  // a frame cannot produce more voxels than it has points
  const size_t vmax = (size_t)std::min<int64_t>(max_voxels, std::max<int64_t>(n_max, 1));
  p->table_b = align256((size_t)p->slots * sizeof(uint2));
  p->list_b = align256(std::max<size_t>(vmax * (size_t)max_points, 1) * sizeof(uint32_t));
  p->pslot_b = align256((size_t)p->npad * sizeof(int32_t));
  p->word_b = align256((size_t)p->words * sizeof(uint32_t));
  p->per_frame = p->table_b + p->list_b + p->pslot_b + 2 * p->word_b;
  return PCFE_OK;
}

int hvg_run(const pcfe_frame_t* frames, int num_frames, int c, const HvGlobalPlan& p,
            int max_points, int max_voxels, int32_t* voxel_num, void* workspace, int wave,
            cudaStream_t st) {
  // scratch layout: [tables | lists] (both memset to 0xFF) | pslot | bitmask | wordprefix
  char* base = (char*)workspace;
  HvgWork w;
  w.table = (uint2*)base;
  w.idxlist = (uint32_t*)(base + (size_t)wave * p.table_b);
  w.pslot = (int32_t*)(base + (size_t)wave * (p.table_b + p.list_b));
  w.bitmask = (uint32_t*)(base + (size_t)wave * (p.table_b + p.list_b + p.pslot_b));
  w.wordprefix = (uint32_t*)(base + (size_t)wave * (p.table_b + p.list_b + p.pslot_b + p.word_b));
  w.table_stride = p.table_b / sizeof(uint2);
  w.list_stride = p.list_b / sizeof(uint32_t);
  w.pslot_stride = p.pslot_b / sizeof(int32_t);
  w.word_stride = p.word_b / sizeof(uint32_t);
  w.slots = p.slots;
  w.log2_slots = p.log2_slots;
  w.direct = p.direct;

  // the fast C==4 scatter needs 16-byte aligned rows
  bool vec4_ok = (c == 4);
  for (int k = 0; k < num_frames && vec4_ok; ++k)
    vec4_ok = !(((uintptr_t)frames[k].points & 15) || ((uintptr_t)frames[k].voxels & 15));

  for (int f0 = 0; f0 < num_frames; f0 += wave) {
    const int wv = std::min(wave, num_frames - f0);
    HvBatch b;
    int64_t wn_max = 0;
    for (int k = 0; k < wv; ++k) {
      const pcfe_frame_t& fr = frames[f0 + k];
      b.f[k] = HvFrame{fr.points, fr.voxels, fr.coors, fr.num_points, (int)fr.n, 0};
      wn_max = std::max(wn_max, fr.n);
    }
    {
      ProfScope ps("memset_scratch", st);
      PCFE_CUDA_TRY(cudaMemsetAsync(base, 0xFF, (size_t)wave * (p.table_b + p.list_b), st));
      count_launch();
    }
    const int wnpad = std::max((int)((wn_max + 31) / 32 * 32), 32);
    const int wwords = wnpad / 32;
    const dim3 pgrid((unsigned)((wnpad + kTilePts - 1) / kTilePts), (unsigned)wv);
    {
      ProfScope ps("hvg_key_hash", st);
      hvg_key_hash_kernel<<<pgrid, kThreads, 0, st>>>(b, w, p.g, c);
      PCFE_LAUNCH_CHECK();
    }
    {
      ProfScope ps("hvg_first_flag", st);
      hvg_first_flag_kernel<<<pgrid, kThreads, 0, st>>>(b, w, wnpad);
      PCFE_LAUNCH_CHECK();
    }
    int rc = hv_launch_scan(w.bitmask, w.word_stride, w.wordprefix, w.word_stride, wwords,
                            max_voxels, voxel_num + f0, wv, 0, st);
    if (rc != PCFE_OK) return rc;
    {
      ProfScope ps("hvg_assign", st);
      hvg_assign_kernel<<<pgrid, kThreads, 0, st>>>(b, w, p.g, max_points, max_voxels);
      PCFE_LAUNCH_CHECK();
    }
    // rows are bounded by both caps and by the points that exist
    const int64_t vmax = std::min<int64_t>(max_voxels, wn_max);
    const int64_t rows = std::max<int64_t>(vmax * std::max(max_points, 1), 1);
    const dim3 sgrid((unsigned)((rows + kTilePts - 1) / kTilePts), (unsigned)wv);
    {
      ProfScope ps("hvg_scatter", st);
      if (vec4_ok) launch_scatter<4>(sgrid, st, b, w, c, max_points, voxel_num + f0);
      else if (c == 5) launch_scatter<5>(sgrid, st, b, w, c, max_points, voxel_num + f0);
      else launch_scatter<0>(sgrid, st, b, w, c, max_points, voxel_num + f0);
      PCFE_LAUNCH_CHECK();
    }
  }
  return PCFE_OK;
}

}  // namespace pcfe
