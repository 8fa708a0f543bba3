// detmatch_b200/csrc/hv_bucket.cu -- hard voxelization, shared-memory bucket path (the fast path).
//
// Behaviour reproduced bit for bit: mmdet3d/ops/voxel/src/voxelization_cpu.cpp:43-99.
//
// Measured on B200 (tests/native/microbench.cu): random L2 atomics 135 Gops/s, shared-memory
// atomics 1500 Gops/s chip-wide.  So points are first partitioned by a hash of their cell key into
// NB buckets per frame (streaming, coalesced), and all grouping work -- cell de-duplication,
// per-cell point counts, per-cell sorted lists of the first P point indices -- happens in the
// shared memory of the one CTA that owns a bucket.  Only two things are global per frame: the
// bit-per-point "is first point of its voxel" mask whose popcount prefix gives voxel ids in
// first-occurrence order, and the voxel-id -> record map the final expansion pass reads.
//
//   A  hvb_bin      rows -> cell key -> bucket = top hash bits; a 4096-point tile is counting-
//                   sorted by bucket in shared memory and written as contiguous runs of
//                   (key, point index) entries into the per-bucket regions (one L2 atomicAdd
//                   per (tile, bucket)).
//   B  hvb_bucket   one CTA per (frame, bucket): smem open-addressing table of the bucket's
//                   cells; pass 1 counts points per cell; a block scan sizes per-cell lists
//                   (min(count, P) entries); pass 2 inserts point indices into the sorted lists
//                   (smem atomicMin chains); then each cell is emitted as a variable-length record
//                   {key, len, idx[len]} plus a directory entry, and its first point is flagged
//                   in the bitmask.
//   C  hv_scan_flags popcount prefix of the bitmask per frame; voxel_num = min(#cells, V).
//   D  hvb_order    directory entry -> voxel id = rank of the cell's first point;
//                   order[vid] = record offset.
//   E  hvb_expand   voxel-id order: a warp resolves the source points of 32 consecutive output rows
//                   and moves them with transposed, fully coalesced stores (data + zero padding);
//                   plus coors and counts.
//   F  fallback     (hv_global.cu, one CTA per frame) for frames whose bucket regions overflowed
//                   (heavy duplication / adversarial keys); a no-op launch otherwise.
#include <algorithm>

#include "hv_common.cuh"

namespace pcfe {

int hv_launch_scan(const uint32_t* bitmask, size_t bitmask_stride, uint32_t* prefix,
                   size_t prefix_stride, int words, int max_voxels, int32_t* voxel_num, int frames,
                   int paired, cudaStream_t st);
int hvg_launch_slow(const HvBatch& b, int frames, const uint32_t* overflow, size_t overflow_stride,
                    int force, char* scratch_base, size_t scratch_stride, const HvGlobalPlan& p,
                    uint32_t* bitmask, size_t bitmask_stride, uint32_t* prefix, size_t prefix_stride,
                    int c, int max_points, int max_voxels, int32_t* voxel_num, cudaStream_t st);

int g_opt_bucket_avg = 1024;  // target points per bucket (tunable through pcfe_debug_set)

namespace {

struct HvbWork {
  char* region;          // [W] per-frame scratch: ent | rec | dir | order
  size_t region_stride;  // bytes
  size_t rec_off, dir_off, order_off;  // byte offsets inside a frame region
  uint32_t* zero;        // [W] per-frame zeroed block: bitmask[words] | ctl
  size_t zero_stride;    // words
  size_t ctl_off;        // words: ctl = bucket_cnt[nb] | rec_cursor | dir_cursor | overflow
  uint32_t* wordprefix;  // [W][2 * words]  {bitmask word, exclusive popcount prefix} pairs
  size_t word_stride;    // words
  int nb, log2_nb, cap, slots, log2_slots;
  uint32_t rec_words, dir_cap;

  __device__ __forceinline__ uint2* ent(int f) const { return reinterpret_cast<uint2*>(region + (size_t)f * region_stride); }
  __device__ __forceinline__ uint32_t* rec(int f) const { return reinterpret_cast<uint32_t*>(region + (size_t)f * region_stride + rec_off); }
  __device__ __forceinline__ uint2* dir(int f) const { return reinterpret_cast<uint2*>(region + (size_t)f * region_stride + dir_off); }
  __device__ __forceinline__ uint32_t* order(int f) const { return reinterpret_cast<uint32_t*>(region + (size_t)f * region_stride + order_off); }
  __device__ __forceinline__ uint32_t* bitmask(int f) const { return zero + (size_t)f * zero_stride; }
  __device__ __forceinline__ uint32_t* ctl(int f) const { return zero + (size_t)f * zero_stride + ctl_off; }
  __device__ __forceinline__ uint32_t* prefix(int f) const { return wordprefix + (size_t)f * word_stride; }
};
// ctl word indices after the nb bucket counters
constexpr int kCtlRec = 0, kCtlDir = 1, kCtlOverflow = 2;

// ------------------------------------------------------------------------------------------
// A: partition points into hash buckets
// ------------------------------------------------------------------------------------------
constexpr int kBinThreads = 512;
constexpr int kBinPerThread = 8;
constexpr int kBinTile = kBinThreads * kBinPerThread;  // 4096 points
constexpr int kMaxBuckets = 1024;

__global__ void __launch_bounds__(kBinThreads)
hvb_bin_kernel(const __grid_constant__ HvBatch batch, const HvbWork w, const GridParams g,
               const int c) {
  __shared__ uint32_t hist[kMaxBuckets];   // entries of this tile per bucket
  __shared__ uint32_t soff[kMaxBuckets];   // exclusive prefix of hist (staging offsets)
  __shared__ uint32_t gbase[kMaxBuckets];  // position of this tile's run inside the bucket
  __shared__ uint2 stage[kBinTile];
  __shared__ uint32_t warp_sums[33];

  const int f = blockIdx.y;
  const HvFrame& fr = batch.f[f];
  const int tid = threadIdx.x;
  const int tile0 = blockIdx.x * kBinTile;
  if (tile0 >= fr.n) return;
  for (int b = tid; b < w.nb; b += kBinThreads) hist[b] = 0;
  __syncthreads();

  uint32_t key[kBinPerThread];
  uint32_t rank[kBinPerThread];
  const int shift = 32 - w.log2_nb;
#pragma unroll
  for (int k = 0; k < kBinPerThread; ++k) {
    const int i = tile0 + k * kBinThreads + tid;
    key[k] = kEmpty;
    if (i < fr.n) {
      float x, y, z;
      load_xyz(fr.pts, i, c, x, y, z);
      int cx, cy, cz;
      key[k] = point_key(x, y, z, g, cx, cy, cz);
    }
  }
#pragma unroll
  for (int k = 0; k < kBinPerThread; ++k) {
    if (key[k] != kEmpty) {
      const uint32_t b = w.log2_nb ? (key[k] * kGold) >> shift : 0u;
      rank[k] = atomicAdd(&hist[b], 1u);
    }
  }
  __syncthreads();
  // reserve a run in every bucket this tile contributes to; scan hist for the staging offsets
  uint32_t* ctl = w.ctl(f);
  uint32_t total = 0;
  for (int b0 = 0; b0 < w.nb; b0 += kBinThreads) {
    const int b = b0 + tid;
    const uint32_t h = b < w.nb ? hist[b] : 0u;
    uint32_t tot;
    const uint32_t ex = block_exscan(h, warp_sums, &tot);
    if (b < w.nb) {
      soff[b] = total + ex;
      uint32_t gb = 0;
      if (h) {
        gb = atomicAdd(&ctl[b], h);
        if (gb + h > (uint32_t)w.cap) ctl[w.nb + kCtlOverflow] = 1u;  // frame takes the fallback
      }
      gbase[b] = gb;
    }
    total += tot;
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < kBinPerThread; ++k) {
    if (key[k] != kEmpty) {
      const uint32_t b = w.log2_nb ? (key[k] * kGold) >> shift : 0u;
      stage[soff[b] + rank[k]] = make_uint2(key[k], (uint32_t)(tile0 + k * kBinThreads + tid));
    }
  }
  __syncthreads();
  uint2* __restrict__ ent = w.ent(f);
  for (uint32_t j = tid; j < total; j += kBinThreads) {
    const uint2 e = stage[j];
    const uint32_t b = w.log2_nb ? (e.x * kGold) >> shift : 0u;
    const uint32_t dst = gbase[b] + (j - soff[b]);
    if (dst < (uint32_t)w.cap) ent[(size_t)b * w.cap + dst] = e;
  }
}

// ------------------------------------------------------------------------------------------
// B: one CTA per (frame, bucket), everything in shared memory
// ------------------------------------------------------------------------------------------
constexpr int kBucketThreads = 256;
constexpr int kMaxCap = 2048;  // entries per bucket; kMaxCap / kBucketThreads record offsets in registers

// dynamic shared memory (words): hkey[S] | hval[S] | eidx[cap] | lists[cap] | eslot[cap] (u16) |
// slotlist[cap] (u16)
__global__ void __launch_bounds__(kBucketThreads)
hvb_bucket_kernel(const HvbWork w, const int pe /* max(max_points, 1) */) {
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ uint32_t warp_sums[33];
  __shared__ uint32_t s_nclaimed, s_rec_base, s_dir_base;

  const int f = blockIdx.y, b = blockIdx.x, tid = threadIdx.x;
  uint32_t* ctl = w.ctl(f);
  if (ctl[w.nb + kCtlOverflow]) return;  // set by the bin kernel: the frame takes the fallback
  const int S = w.slots, cap = w.cap;
  uint32_t* hkey = smem;
  uint32_t* hval = hkey + S;   // pass 1: point count; after the scan: (list offset << 16) | list len
  uint32_t* eidx = hval + S;
  uint32_t* lists = eidx + cap;
  uint16_t* eslot = reinterpret_cast<uint16_t*>(lists + cap);
  uint16_t* slotlist = eslot + cap;

  const int ne = (int)min(ctl[b], (uint32_t)cap);
  if (ne == 0) return;
  {  // hkey = kEmpty, hval = 0 (S is a multiple of 4; the two arrays are contiguous)
    uint4* k4 = reinterpret_cast<uint4*>(hkey);
    uint4* v4 = reinterpret_cast<uint4*>(hval);
    for (int s = tid; s < S / 4; s += kBucketThreads) {
      k4[s] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
      v4[s] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (tid == 0) s_nclaimed = 0u;
  __syncthreads();

  // pass 1: find-or-claim the cell's slot, count its points
  const uint2* __restrict__ ent = w.ent(f) + (size_t)b * cap;
  const uint32_t smask = (uint32_t)S - 1u;
  const int sshift = 32 - w.log2_nb - w.log2_slots;  // hash bits right below the bucket bits
  for (int e = tid; e < ne; e += kBucketThreads) {
    const uint2 en = __ldcs(&ent[e]);
    uint32_t s = ((en.x * kGold) >> sshift) & smask;
    while (true) {
      uint32_t cur = *reinterpret_cast<volatile uint32_t*>(&hkey[s]);
      if (cur == en.x) break;
      if (cur == kEmpty) {
        cur = atomicCAS(&hkey[s], kEmpty, en.x);
        if (cur == kEmpty) {
          slotlist[atomicAdd(&s_nclaimed, 1u)] = (uint16_t)s;
          break;
        }
        if (cur == en.x) break;
      }
      s = (s + 1u) & smask;
    }
    atomicAdd(&hval[s], 1u);
    eslot[e] = (uint16_t)s;
    eidx[e] = en.y;
  }
  __syncthreads();

  // size the per-cell lists (min(count, P) entries each, in claim order) and the records
  // (2 + len words each) with ONE block scan of the packed pair (record words << 16 | list len):
  // both totals stay below 2^16 because cap <= 2048.  Cell j of chunk q is handled by thread
  // j - q * kBucketThreads, which also initialises the cell's list.
  const int nv = (int)s_nclaimed;
  uint32_t run = 0;  // packed running totals
  uint32_t my_rec_off[kMaxCap / kBucketThreads];
#pragma unroll
  for (int q = 0; q < kMaxCap / kBucketThreads; ++q) {
    const int j = q * kBucketThreads + tid;
    if (q * kBucketThreads >= nv) break;  // block-uniform
    uint32_t len = 0;
    int s = 0;
    if (j < nv) {
      s = slotlist[j];
      len = min(hval[s], (uint32_t)pe);
    }
    uint32_t tot;
    const uint32_t ex = run + block_exscan(j < nv ? ((len + 2u) << 16) | len : 0u, warp_sums, &tot);
    if (j < nv) {
      const uint32_t off = ex & 0xFFFFu;
      hval[s] = (off << 16) | len;
      for (uint32_t t = 0; t < len; ++t) lists[off + t] = kEmpty;
    }
    my_rec_off[q] = ex >> 16;
    run += tot;
    __syncthreads();  // warp_sums is reused by the next chunk
  }
  const uint32_t run_rec = run >> 16;
  if (tid == 0) {
    s_rec_base = atomicAdd(&ctl[w.nb + kCtlRec], run_rec);
    s_dir_base = atomicAdd(&ctl[w.nb + kCtlDir], (uint32_t)nv);
  }
  __syncthreads();

  // pass 2: sorted lists of the first P point indices of every cell
  for (int e = tid; e < ne; e += kBucketThreads) {
    const uint32_t hv = hval[eslot[e]];
    sorted_insert<false>(lists + (hv >> 16), (int)(hv & 0xFFFFu), eidx[e]);
  }
  __syncthreads();

  // emit records + directory entries and flag first points
  const uint32_t rec_base = s_rec_base, dir_base = s_dir_base;
  if (rec_base + run_rec > w.rec_words || dir_base + (uint32_t)nv > w.dir_cap) {
    if (tid == 0) ctl[w.nb + kCtlOverflow] = 1u;  // cannot happen: arenas are sized for n points
    return;
  }
  uint32_t* __restrict__ rec = w.rec(f);
  uint2* __restrict__ dir = w.dir(f);
  uint32_t* __restrict__ bitmask = w.bitmask(f);
#pragma unroll
  for (int q = 0; q < kMaxCap / kBucketThreads; ++q) {
    const int j = q * kBucketThreads + tid;
    if (j >= nv) break;
    const int s = slotlist[j];
    const uint32_t hv = hval[s];
    const uint32_t len = hv & 0xFFFFu;  // >= 1: every claimed cell has a point and pe >= 1
    const uint32_t* lst = lists + (hv >> 16);
    const uint32_t off = rec_base + my_rec_off[q];
    uint32_t* r = rec + off;
    r[0] = hkey[s];
    r[1] = len;
    for (uint32_t t = 0; t < len; ++t) r[2 + t] = lst[t];
    const uint32_t first = lst[0];  // lists are ascending: entry 0 is the cell's first point
    dir[dir_base + j] = make_uint2(off, first);
    atomicOr(&bitmask[first >> 5], 1u << (first & 31));
  }
}

// ------------------------------------------------------------------------------------------
// D: voxel id of every cell; order[vid] = record offset
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
hvb_order_kernel(const HvbWork w, const int max_voxels) {
  const int f = blockIdx.y;
  const uint32_t* ctl = w.ctl(f);
  if (ctl[w.nb + kCtlOverflow]) return;
  const uint32_t ncell = min(ctl[w.nb + kCtlDir], w.dir_cap);
  const uint32_t t = blockIdx.x * 256u + threadIdx.x;
  if (t >= ncell) return;
  const uint2 d = w.dir(f)[t];
  // wordprefix holds {bitmask word, exclusive prefix} pairs: one 8-byte random read per cell
  const uint2 bp = reinterpret_cast<const uint2*>(w.prefix(f))[d.y >> 5];
  const uint32_t vid = bp.y + __popc(bp.x & ((1u << (d.y & 31)) - 1u));
  if (vid < (uint32_t)max_voxels) w.order(f)[vid] = d.x;  // voxelization_cpu.cpp:78
}

// ------------------------------------------------------------------------------------------
// E: expansion in voxel-id order
// ------------------------------------------------------------------------------------------
// A warp owns 32 consecutive output rows (32 * C contiguous floats of the voxels buffer).  Lane l
// first resolves the source point of row l (order -> record -> point index), then the warp moves
// the 32 * C words in C fully coalesced store instructions: in iteration t lane l handles word
// t * 32 + l of the chunk, i.e. column (word % C) of row (word / C), whose source index comes
// from the owning lane by shuffle.  Gathers touch ~32 / C + 1 distinct point rows per instruction
// instead of 32 (a per-thread row copy costs one L1 tag lookup per lane and column).
constexpr int kExpThreads = 256;
constexpr int kExpRowsPerWarp = 32;
constexpr int kExpIters = 4;                                          // row groups per warp
constexpr int kExpTile = (kExpThreads / 32) * kExpRowsPerWarp * kExpIters;  // rows per CTA

template <int C>
__global__ void __launch_bounds__(kExpThreads)
hvb_expand_kernel(const __grid_constant__ HvBatch batch, const HvbWork w, const GridParams g,
                  const int c_rt, const int max_points, const int32_t* __restrict__ voxel_num) {
  const int f = blockIdx.y;
  if (w.ctl(f)[w.nb + kCtlOverflow]) return;
  const HvFrame& fr = batch.f[f];
  const int c = C > 0 ? C : c_rt;
  const int m = voxel_num[f];
  const long long rows = (long long)m * max_points;
  const long long r_base = (long long)blockIdx.x * kExpTile;
  const uint32_t* __restrict__ rec = w.rec(f);
  const uint32_t* __restrict__ order = w.order(f);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  // coors + counts: one thread per voxel id
#pragma unroll
  for (int k = 0; k < kExpTile / kExpThreads; ++k) {
    const long long v = r_base + k * kExpThreads + tid;
    if (v < m) {
      const uint32_t q = order[v];
      decode_key(rec[q], g, fr.coors + (size_t)v * 3);
      fr.num[v] = (int32_t)min(rec[q + 1], (uint32_t)max_points);
    }
  }
  if (r_base >= rows) return;
  const float* __restrict__ pts = fr.pts;
  // Three phases with all kExpIters row groups of the warp in flight in each, so that the
  // dependent chain order -> record -> point row costs three memory latencies per warp, not
  // three per row group.
  uint32_t q[kExpIters], s[kExpIters], src[kExpIters];
  long long wr0[kExpIters];
#pragma unroll
  for (int it = 0; it < kExpIters; ++it) {
    wr0[it] = r_base + ((long long)(it * (kExpThreads / 32) + wid)) * kExpRowsPerWarp;
    const long long r = wr0[it] + lane;
    q[it] = kEmpty;
    s[it] = 0;
    if (r < rows) {
      const uint32_t vid = (uint32_t)r / (uint32_t)max_points;  // rows < 2^31 (checked by the plan)
      s[it] = (uint32_t)r - vid * (uint32_t)max_points;
      q[it] = order[vid];
    }
  }
#pragma unroll
  for (int it = 0; it < kExpIters; ++it) {
    src[it] = kEmpty;  // point index feeding the row, kEmpty = zero padding
    if (q[it] != kEmpty) {
      // both loads are issued together; the arena has max_points words of slack behind the
      // last record, so reading the list entry before knowing the length stays in bounds
      const uint32_t len = rec[q[it] + 1];
      const uint32_t idx = rec[q[it] + 2 + s[it]];
      if (s[it] < len) src[it] = idx;
    }
  }
  float v[kExpIters][C > 0 ? C : 1];
  if (C > 0) {
#pragma unroll
    for (int it = 0; it < kExpIters; ++it) {
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const int t = k * 32 + lane;
        const int row = t / C;
        const int col = t - row * C;
        const uint32_t sidx = __shfl_sync(0xFFFFFFFFu, src[it], row);
        v[it][k] = (sidx != kEmpty) ? __ldg(pts + (size_t)sidx * C + col) : 0.0f;
      }
    }
#pragma unroll
    for (int it = 0; it < kExpIters; ++it) {
      if (wr0[it] >= rows) break;  // warp-uniform
      const int nwords = (int)min((long long)kExpRowsPerWarp, rows - wr0[it]) * C;
      float* __restrict__ dst = fr.voxels + (size_t)wr0[it] * C;
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const int t = k * 32 + lane;
        if (t < nwords) __stcs(dst + t, v[it][k]);
      }
    }
  } else {
    for (int it = 0; it < kExpIters; ++it) {
      if (wr0[it] >= rows) break;  // warp-uniform
      const int nwords = (int)min((long long)kExpRowsPerWarp, rows - wr0[it]) * c;
      float* __restrict__ dst = fr.voxels + (size_t)wr0[it] * c;
      for (int t = lane; t < kExpRowsPerWarp * c; t += 32) {  // warp-uniform trip count
        const int row = t / c;
        const int col = t - row * c;
        const uint32_t sidx = __shfl_sync(0xFFFFFFFFu, src[it], row & 31);
        if (t < nwords) __stcs(dst + t, (sidx != kEmpty) ? __ldg(pts + (size_t)sidx * c + col) : 0.0f);
      }
    }
  }
}

template <int C>
int launch_expand(dim3 grid, cudaStream_t st, const HvBatch& b, const HvbWork& w,
                  const GridParams& g, int c, int p, const int32_t* vn) {
  hvb_expand_kernel<C><<<grid, kExpThreads, 0, st>>>(b, w, g, c, p, vn);
  PCFE_LAUNCH_CHECK();
  return PCFE_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int hvb_make_plan(int64_t n_max, int c, const float vs[3], const float rg[6], int max_points,
                  int max_voxels, HvBucketPlan* p) {
  (void)c;
  int rc = hvg_make_plan(n_max, vs, rg, max_points, max_voxels, &p->slow);
  if (rc != PCFE_OK) return rc;
  if (max_points > 0xFFFF) return PCFE_ERR_CAPS;  // list length is packed into 16 bits
  if ((int64_t)max_voxels * (int64_t)std::max(max_points, 1) >= (1ll << 31)) return PCFE_ERR_TOO_LARGE;
  p->g = p->slow.g;
  p->npad = p->slow.npad;
  p->words = p->slow.words;
  const int target = std::max(64, std::min(g_opt_bucket_avg, 1400));
  int lg = 0;
  while ((1 << lg) < kMaxBuckets && ((int64_t)target << lg) < n_max) ++lg;
  p->log2_nb = lg;
  p->nb = 1 << lg;
  const int64_t avg = (n_max + p->nb - 1) / p->nb;
  if (avg + avg / 4 > kMaxCap) return PCFE_ERR_TOO_LARGE;  // caller uses the global path
  // head-room for hash imbalance and for heavily populated cells (pillars): exceeding it is not
  // an error, the frame just takes the fallback
  p->cap = (int)std::min<int64_t>(kMaxCap, avg + std::max<int64_t>(avg / 2, 768) +
                                               std::min<int64_t>(8 * (int64_t)max_points, 512));
  p->cap = (p->cap + 7) & ~7;
  int ls = 6;
  while ((1 << ls) < p->cap + p->cap / 4) ++ls;
  p->log2_slots = ls;
  p->slots = 1 << ls;
  if (p->log2_nb + p->log2_slots > 32) return PCFE_ERR_TOO_LARGE;
  // arenas: every cell needs 2 + len words, sum(len) <= n, cells <= n  ->  3 * npad words
  p->rec_words = (size_t)3 * (size_t)p->npad;
  p->ent_b = align256((size_t)p->nb * (size_t)p->cap * sizeof(uint2));
  p->rec_b = align256((p->rec_words + (size_t)std::max(max_points, 1) + 8) * sizeof(uint32_t));  // + read slack
  p->dir_b = align256((size_t)p->npad * sizeof(uint2));
  const size_t vmax = (size_t)std::min<int64_t>(max_voxels, std::max<int64_t>(n_max, 1));
  p->order_b = align256(std::max<size_t>(vmax, 1) * sizeof(uint32_t));
  p->word_b = align256((size_t)p->words * sizeof(uint32_t));
  p->cnt_b = align256((size_t)(p->nb + 4) * sizeof(uint32_t));
  const size_t fast = p->ent_b + p->rec_b + p->dir_b + p->order_b;
  // the fallback reuses the frame's own region as table | lists | pslot
  const size_t slow = p->slow.table_b + p->slow.list_b + p->slow.pslot_b;
  p->region_b = std::max(fast, slow);
  p->per_frame = p->region_b + 3 * p->word_b + p->cnt_b;  // bitmask + {bits, prefix} pairs
  p->smem_bucket = (size_t)(2 * p->slots + 2 * p->cap) * 4 + (size_t)(2 * p->cap) * 2;
  return PCFE_OK;
}

int hvb_run(const pcfe_frame_t* frames, int num_frames, int c, const HvBucketPlan& p,
            int max_points, int max_voxels, int32_t* voxel_num, void* workspace, int wave,
            cudaStream_t st) {
  // scratch layout per wave: [frame regions] x wave | [bitmask | ctl] x wave (zeroed per wave) |
  // [wordprefix] x wave
  char* base = (char*)workspace;
  HvbWork w;
  w.region = base;
  w.region_stride = p.region_b;
  w.rec_off = p.ent_b;
  w.dir_off = p.ent_b + p.rec_b;
  w.order_off = p.ent_b + p.rec_b + p.dir_b;
  char* zero_base = base + (size_t)wave * p.region_b;
  const size_t zero_per = p.word_b + p.cnt_b;
  w.zero = (uint32_t*)zero_base;
  w.zero_stride = zero_per / sizeof(uint32_t);
  w.ctl_off = p.word_b / sizeof(uint32_t);
  w.wordprefix = (uint32_t*)(zero_base + (size_t)wave * zero_per);
  w.word_stride = 2 * p.word_b / sizeof(uint32_t);
  w.nb = p.nb; w.log2_nb = p.log2_nb; w.cap = p.cap; w.slots = p.slots; w.log2_slots = p.log2_slots;
  w.rec_words = (uint32_t)p.rec_words;
  w.dir_cap = (uint32_t)p.npad;

  PCFE_CUDA_TRY(cudaFuncSetAttribute(hvb_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)p.smem_bucket));
  const int pe = std::max(max_points, 1);

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
      ProfScope ps("memset_ctl", st);
      PCFE_CUDA_TRY(cudaMemsetAsync(zero_base, 0, (size_t)wv * zero_per, st));
      count_launch();
    }
    const int wnpad = std::max((int)((wn_max + 31) / 32 * 32), 32);
    {
      ProfScope ps("hvb_bin", st);
      const dim3 grid((unsigned)((wn_max + kBinTile - 1) / kBinTile), (unsigned)wv);
      hvb_bin_kernel<<<grid, kBinThreads, 0, st>>>(b, w, p.g, c);
      PCFE_LAUNCH_CHECK();
    }
    {
      ProfScope ps("hvb_bucket", st);
      const dim3 grid((unsigned)p.nb, (unsigned)wv);
      hvb_bucket_kernel<<<grid, kBucketThreads, p.smem_bucket, st>>>(w, pe);
      PCFE_LAUNCH_CHECK();
    }
    int rc = hv_launch_scan(w.zero, w.zero_stride, w.wordprefix, w.word_stride, wnpad / 32,
                            max_voxels, voxel_num + f0, wv, 1, st);
    if (rc != PCFE_OK) return rc;
    {
      ProfScope ps("hvb_order", st);
      const dim3 grid((unsigned)((wnpad + 255) / 256), (unsigned)wv);
      hvb_order_kernel<<<grid, 256, 0, st>>>(w, max_voxels);
      PCFE_LAUNCH_CHECK();
    }
    {
      ProfScope ps("hvb_expand", st);
      const int64_t vmax = std::min<int64_t>(max_voxels, wn_max);
      const int64_t rows = std::max<int64_t>(vmax * std::max(max_points, 1), 1);
      const dim3 grid((unsigned)((rows + kExpTile - 1) / kExpTile), (unsigned)wv);
      const int32_t* vn = voxel_num + f0;
      if (c == 4) rc = launch_expand<4>(grid, st, b, w, p.g, c, max_points, vn);
      else if (c == 5) rc = launch_expand<5>(grid, st, b, w, p.g, c, max_points, vn);
      else rc = launch_expand<0>(grid, st, b, w, p.g, c, max_points, vn);
      if (rc != PCFE_OK) return rc;
    }
    rc = hvg_launch_slow(b, wv, w.zero + w.ctl_off + p.nb + kCtlOverflow, w.zero_stride,
                         g_opt_force_overflow, w.region, w.region_stride, p.slow, w.zero,
                         w.zero_stride, w.wordprefix, w.word_stride, c, max_points, max_voxels,
                         voxel_num + f0, st);
    if (rc != PCFE_OK) return rc;
  }
  return PCFE_OK;
}

}  // namespace pcfe
