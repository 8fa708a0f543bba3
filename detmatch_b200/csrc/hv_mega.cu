// detmatch_b200/csrc/hv_mega.cu -- hard voxelization as ONE persistent kernel (the frame pipeline).
//
// Behaviour reproduced bit for bit: mmdet3d/ops/voxel/src/voxelization_cpu.cpp:43-99.
//
// Same algorithm as hv_bucket.cu (partition into hash buckets -> group in shared memory -> voxel
// ids from a popcount prefix of the first-point bitmask -> records permuted into voxel-id order ->
// expansion), but the five launches per wave become four STAGES of one persistent kernel:
//
//   bin(f, tile) -> bucket(f, b) [last one scans the frame's bitmask] -> order(f, chunk) -> expand(f, item)
//
// Work items are handed out by a single global ticket counter in a fixed order: step s carries
// the bin items of frame s, the bucket items of frame s - d1, the order items of frame s - d2 and
// the expand items of frame s - d3, interleaved.  An item waits (one thread spins on a per-frame
// counter with ld.acquire) until the previous stage of its frame is complete.  Every dependency
// points to items with SMALLER tickets, and a ticket is only held by a resident CTA, so the item
// with the smallest incomplete ticket can always run: no deadlock, whatever the CTA count.
//
// Why: (1) only ~d3 + 1 frames are in flight, so the partition entries, the cell records, the
// lists and the voxel-ordered records live and die in the 126 MB L2 (a small RING of scratch
// regions reused every `ring` frames) -- DRAM sees the rows once and the outputs once; (2) the
// DRAM-bound stages (bin, expand) and the shared-memory / issue-bound stage (bucket) run on the
// same SMs at the same time; (3) no launch gaps or tails between phases.
#include <algorithm>
#include <cstdio>

#include "hv_common.cuh"

namespace pcfe {

int hvg_launch_slow_ring(const HvBatch& b, int frames, int ring, uint32_t* overflow,
                         size_t overflow_stride, int force, char* scratch_base, size_t scratch_stride,
                         const HvGlobalPlan& p, uint32_t* bitmask, size_t bitmask_stride,
                         uint32_t* prefix, size_t prefix_stride, int c, int max_points,
                         int max_voxels, int32_t* voxel_num, cudaStream_t st);

int g_opt_mega_d1 = 2, g_opt_mega_d2 = 3, g_opt_mega_d3 = 4;  // stage offsets in steps
int g_opt_mega_ring = 8;                                      // scratch regions
int g_opt_mega_ctas = 0;                                      // CTAs per SM (0 = occupancy)
int g_opt_mega_stats = 0;                                     // debug: print per-stage cycle counts (synchronises)

namespace {

constexpr int kMT = 256;                       // threads per CTA, every stage
constexpr int kMBinPer = 8;                    // points per thread of a bin item
constexpr int kMBinTile = kMT * kMBinPer;      // 2048 points
constexpr int kMOrderPer = 8;                  // cells per thread of an order item
constexpr int kMOrderChunk = kMT * kMOrderPer; // 2048 cells
constexpr int kMWarps = kMT / 32;
constexpr int kMExpTiles = 4;                  // 32-voxel tiles per warp of an expand item
constexpr int kMExpVox = kMWarps * kMExpTiles * 32;  // 1024 voxels
constexpr int kMExpStageWords = 1024;          // per warp (32 voxels x P x C words)
constexpr int kMScanPer = 23;                  // bitmask words per thread and scan chunk (odd: no bank conflicts)
constexpr int kMScanChunk = kMT * kMScanPer;   // 5888 words
constexpr int kMMaxItems = 2048;               // items per step (table of u16)

enum { kSBin = 0, kSBucket = 1, kSOrder = 2, kSExpand = 3, kStages = 4 };
// ctl words of a frame after its nb bucket counters
enum { cList = 0, cCell, cOverflow, cBinDone, cBucketDone, cScanDone, cOrderDone, cExpandDone, cWords };

struct __align__(16) Cell {  // one occupied voxel of a frame
  uint32_t key, len, list_off, first;
};

struct KeyDecode {
  uint32_t plane, gx;      // gx * gy, gx
  uint32_t m_plane, m_gx;  // floor(2^32 / plane), floor(2^32 / gx)
};

struct MegaWork {
  char* region;          // [ring] scratch regions: ent | lists | cells | vcell
  size_t region_stride;  // bytes
  size_t lst_off, cells_off, vcell_off;
  uint32_t* bitmask;     // [frames][words]
  uint32_t* pairs;       // [frames][2 * words]  {bitmask word, exclusive popcount prefix}
  uint32_t* ctl;         // [frames][nb + cWords]
  uint32_t* ticket;      // [1]
  size_t word_stride;    // words (bitmask); pairs use twice that
  size_t ctl_stride;     // words
  int ring, frames, words;
  int nb, log2_nb, cap, slots, log2_slots;
  uint32_t arena_cap;
  int cnt[kStages];      // items per step and stage
  int off[kStages];      // step offset of the stage (off[0] = 0)
  int q;                 // items per step = sum cnt
  int steps;             // frames + off[kSExpand]

  __device__ __forceinline__ char* reg(int f) const { return region + (size_t)(f % ring) * region_stride; }
  __device__ __forceinline__ uint2* ent(int f) const { return reinterpret_cast<uint2*>(reg(f)); }
  __device__ __forceinline__ uint32_t* lst(int f) const { return reinterpret_cast<uint32_t*>(reg(f) + lst_off); }
  __device__ __forceinline__ Cell* cells(int f) const { return reinterpret_cast<Cell*>(reg(f) + cells_off); }
  __device__ __forceinline__ Cell* vcell(int f) const { return reinterpret_cast<Cell*>(reg(f) + vcell_off); }
  __device__ __forceinline__ uint32_t* bm(int f) const { return bitmask + (size_t)f * word_stride; }
  __device__ __forceinline__ uint2* pr(int f) const { return reinterpret_cast<uint2*>(pairs + (size_t)f * 2 * word_stride); }
  __device__ __forceinline__ uint32_t* cl(int f) const { return ctl + (size_t)f * ctl_stride; }
};

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// thread 0 of the CTA: spin until *p >= target
__device__ __forceinline__ void wait_ge(const uint32_t* p, uint32_t target, long long* waited = nullptr) {
  long long t = 0;
  if (waited) t = clock64();
  while (ld_acquire(p) < target) __nanosleep(32);
  if (waited) *waited += clock64() - t;
}
// every thread's writes of the item are ordered before the counter increment: __syncthreads()
// (all writes happen-before thread 0's fence), fence, relaxed atomic
__device__ __forceinline__ void signal_done(uint32_t* p) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(p, 1u);
  }
}

__device__ __forceinline__ uint32_t div_small_err(uint32_t n, uint32_t d, uint32_t m) {
  uint32_t q = __umulhi(n, m);  // true quotient - 2 <= q <= true quotient
  uint32_t r = n - q * d;
  if (r >= d) { ++q; r -= d; }
  if (r >= d) { ++q; }
  return q;
}

// ------------------------------------------------------------------------------------------
// stage A: partition a 2048-point tile of frame f into the frame's hash buckets
// shared memory (words): hist[nb] | soff[nb] | gbase[nb] | stage[2 * 2048]
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_bin(const HvFrame& fr, const MegaWork& w, const GridParams& g,
                                          const int c, const int f, const int tile, uint32_t* smem,
                                          uint32_t* warp_sums) {
  const int tid = threadIdx.x;
  uint32_t* hist = smem;
  uint32_t* soff = hist + w.nb;
  uint32_t* gbase = soff + w.nb;
  uint2* stage = reinterpret_cast<uint2*>(smem + ((3 * w.nb + 3) & ~3));
  // this tile's share of the first-point bitmask is zeroed here (64 words per 2048 points; tiles
  // past the frame's end clear the rest of the row)
  if (tid < kMBinTile / 32) {
    const int wd = tile * (kMBinTile / 32) + tid;
    if (wd < w.words) w.bm(f)[wd] = 0u;
  }
  const int tile0 = tile * kMBinTile;
  if (tile0 >= fr.n) return;
  for (int b = tid; b < w.nb; b += kMT) hist[b] = 0;
  __syncthreads();

  uint32_t key[kMBinPer];
  uint32_t rank[kMBinPer];
  const int shift = 32 - w.log2_nb;
#pragma unroll
  for (int k = 0; k < kMBinPer; ++k) {
    const int i = tile0 + k * kMT + tid;
    key[k] = kEmpty;
    if (i < fr.n) {
      float x, y, z;
      load_xyz(fr.pts, i, c, x, y, z);
      int cx, cy, cz;
      key[k] = point_key(x, y, z, g, cx, cy, cz);
    }
  }
#pragma unroll
  for (int k = 0; k < kMBinPer; ++k) {
    if (key[k] != kEmpty) {
      const uint32_t b = w.log2_nb ? (key[k] * kGold) >> shift : 0u;
      rank[k] = atomicAdd(&hist[b], 1u);
    }
  }
  __syncthreads();
  uint32_t* ctl = w.cl(f);
  uint32_t total = 0;
  for (int b0 = 0; b0 < w.nb; b0 += kMT) {
    const int b = b0 + tid;
    const uint32_t h = b < w.nb ? hist[b] : 0u;
    uint32_t tot;
    const uint32_t ex = block_exscan(h, warp_sums, &tot);
    if (b < w.nb) {
      soff[b] = total + ex;
      uint32_t gb = 0;
      if (h) {
        gb = atomicAdd(&ctl[b], h);
        if (gb + h > (uint32_t)w.cap) ctl[w.nb + cOverflow] = 1u;  // frame takes the fallback
      }
      gbase[b] = gb;
    }
    total += tot;
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < kMBinPer; ++k) {
    if (key[k] != kEmpty) {
      const uint32_t b = w.log2_nb ? (key[k] * kGold) >> shift : 0u;
      stage[soff[b] + rank[k]] = make_uint2(key[k], (uint32_t)(tile0 + k * kMT + tid));
    }
  }
  __syncthreads();
  uint2* ent = w.ent(f);
  for (uint32_t j = tid; j < total; j += kMT) {
    const uint2 e = stage[j];
    const uint32_t b = w.log2_nb ? (e.x * kGold) >> shift : 0u;
    const uint32_t dst = gbase[b] + (j - soff[b]);
    if (dst < (uint32_t)w.cap) ent[(size_t)b * w.cap + dst] = e;
  }
}

// ------------------------------------------------------------------------------------------
// stage B: group bucket b of frame f in shared memory (chains + register sort, P <= PT)
// shared memory (words): hkey[S] | head[S] | eidx[cap] | enext[cap] (u16) | slotlist[cap] (u16)
// ------------------------------------------------------------------------------------------
template <int PT>
__device__ __forceinline__ void stage_bucket(const MegaWork& w, const int pe, const int f, const int b,
                                             uint32_t* smem, uint32_t* warp_sums, uint32_t* s_misc) {
  constexpr uint32_t kNil = 0xFFFFu;
  const int tid = threadIdx.x;
  uint32_t* ctl = w.cl(f);
  if (__ldcg(&ctl[w.nb + cOverflow])) return;
  const int S = w.slots, cap = w.cap;
  uint32_t* hkey = smem;
  uint32_t* head = hkey + S;
  uint32_t* eidx = head + S;
  uint16_t* enext = reinterpret_cast<uint16_t*>(eidx + cap);
  uint16_t* slotlist = enext + cap;

  const int ne = (int)min(__ldcg(&ctl[b]), (uint32_t)cap);
  if (ne == 0) return;
  {
    uint4* k4 = reinterpret_cast<uint4*>(hkey);  // hkey and head are contiguous: 2 * S words
    for (int s = tid; s < S / 2; s += kMT) k4[s] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
  }
  if (tid == 0) s_misc[0] = 0u;  // cells claimed
  __syncthreads();

  const uint2* ent = w.ent(f) + (size_t)b * cap;
  const uint32_t smask = (uint32_t)S - 1u;
  const int sshift = 32 - w.log2_nb - w.log2_slots;
  for (int e = tid; e < ne; e += kMT) {
    const uint2 en = __ldcg(&ent[e]);
    uint32_t s = ((en.x * kGold) >> sshift) & smask;
    while (true) {
      uint32_t cur = *reinterpret_cast<volatile uint32_t*>(&hkey[s]);
      if (cur == en.x) break;
      if (cur == kEmpty) {
        cur = atomicCAS(&hkey[s], kEmpty, en.x);
        if (cur == kEmpty) {
          slotlist[atomicAdd(&s_misc[0], 1u)] = (uint16_t)s;
          break;
        }
        if (cur == en.x) break;
      }
      s = (s + 1u) & smask;
    }
    const uint32_t prev = atomicExch(&head[s], (uint32_t)e);
    enext[e] = (uint16_t)(prev == kEmpty ? kNil : prev);
    eidx[e] = en.y;
  }
  __syncthreads();

  const int nv = (int)s_misc[0];
  uint32_t* glst = w.lst(f);
  Cell* cells = w.cells(f);
  uint32_t* bitmask = w.bm(f);
  constexpr int kCellsPerThread = 2;
#pragma unroll 1
  for (int j0 = 0; j0 < nv; j0 += kCellsPerThread * kMT) {
    uint32_t sorted[kCellsPerThread][PT];
    uint32_t key[kCellsPerThread], len[kCellsPerThread];
    uint32_t mine = 0;
#pragma unroll
    for (int u = 0; u < kCellsPerThread; ++u) {
      const int j = j0 + u * kMT + tid;
#pragma unroll
      for (int t = 0; t < PT; ++t) sorted[u][t] = kEmpty;
      uint32_t cnt = 0;
      key[u] = 0;
      if (j < nv) {
        const int s = slotlist[j];
        key[u] = hkey[s];
        uint32_t e = head[s];
        while (e != kNil) {  // chain walk; the P smallest indices stay in registers, ascending
          uint32_t v = eidx[e];
          e = enext[e];
          ++cnt;
#pragma unroll
          for (int t = 0; t < PT; ++t) {
            const uint32_t lo = min(sorted[u][t], v);
            v = max(sorted[u][t], v);
            sorted[u][t] = lo;
          }
        }
      }
      len[u] = min(cnt, (uint32_t)pe);
      mine += len[u];
    }
    uint32_t tot;
    uint32_t off = block_exscan(mine, warp_sums, &tot);
    const int ncell = min(kCellsPerThread * kMT, nv - j0);
    if (tid == 0) {
      s_misc[1] = atomicAdd(&ctl[w.nb + cList], tot);
      s_misc[2] = atomicAdd(&ctl[w.nb + cCell], (uint32_t)ncell);
    }
    __syncthreads();
    const uint32_t list_base = s_misc[1], cell_base = s_misc[2];
    if (list_base + tot > w.arena_cap || cell_base + (uint32_t)ncell > w.arena_cap) {
      if (tid == 0) ctl[w.nb + cOverflow] = 1u;  // cannot happen: arenas hold one entry per point
      return;
    }
#pragma unroll
    for (int u = 0; u < kCellsPerThread; ++u) {
      const int jl = u * kMT + tid;
      if (j0 + jl < nv) {
        const uint32_t lo = list_base + off;
#pragma unroll
        for (int t = 0; t < PT; ++t)
          if ((uint32_t)t < len[u]) glst[lo + t] = sorted[u][t];
        Cell cl;
        cl.key = key[u];
        cl.len = len[u];
        cl.list_off = lo;
        cl.first = sorted[u][0];
        cells[cell_base + jl] = cl;
        atomicOr(&bitmask[sorted[u][0] >> 5], 1u << (sorted[u][0] & 31));
        off += len[u];
      }
    }
    __syncthreads();  // s_misc / warp_sums are reused by the next round
  }
}

// popcount prefix of frame f's bitmask -> {word, prefix} pairs; voxel_num = min(#cells, V).
// Run by the LAST bucket CTA of the frame.  shared memory: kMScanChunk words.
__device__ __forceinline__ void frame_scan(const MegaWork& w, const int f, const int max_voxels,
                                           int32_t* voxel_num, uint32_t* smem, uint32_t* warp_sums) {
  const int tid = threadIdx.x;
  const uint32_t* bm = w.bm(f);
  uint2* pr = w.pr(f);
  uint32_t carry = 0;
  for (int c0 = 0; c0 < w.words; c0 += kMScanChunk) {
    const int nw = min(kMScanChunk, w.words - c0);
    for (int i = tid; i < nw; i += kMT) smem[i] = __ldcg(&bm[c0 + i]);  // coalesced, independent loads
    __syncthreads();
    const int lo = tid * kMScanPer, hi = min(lo + kMScanPer, nw);
    uint32_t sum = 0;
    for (int i = lo; i < hi; ++i) sum += __popc(smem[i]);
    uint32_t tot;
    uint32_t run = carry + block_exscan(sum, warp_sums, &tot);
    for (int i = lo; i < hi; ++i) {
      const uint32_t x = smem[i];
      pr[c0 + i] = make_uint2(x, run);
      run += __popc(x);
    }
    carry += tot;
    __syncthreads();  // smem / warp_sums reused
  }
  if (tid == 0) voxel_num[f] = (int32_t)min(carry, (uint32_t)max_voxels);  // voxelization_cpu.cpp:78
}

// ------------------------------------------------------------------------------------------
// stage C: cells [chunk * 2048, +2048) of frame f -> vcell[voxel id]
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_order(const MegaWork& w, const int f, const int chunk,
                                            const uint32_t ncell, const int max_voxels) {
  const uint4* cells = reinterpret_cast<const uint4*>(w.cells(f));
  uint4* vcell = reinterpret_cast<uint4*>(w.vcell(f));
  const uint2* pr = w.pr(f);
  uint4 raw[kMOrderPer];
#pragma unroll
  for (int k = 0; k < kMOrderPer; ++k) {
    const uint32_t t = (uint32_t)chunk * kMOrderChunk + k * kMT + threadIdx.x;
    raw[k] = make_uint4(0u, 0u, 0u, 0u);
    if (t < ncell) raw[k] = __ldcg(cells + t);
  }
  uint2 bp[kMOrderPer];
#pragma unroll
  for (int k = 0; k < kMOrderPer; ++k) {
    const uint32_t t = (uint32_t)chunk * kMOrderChunk + k * kMT + threadIdx.x;
    bp[k] = make_uint2(0u, 0u);
    if (t < ncell) bp[k] = __ldcg(pr + (raw[k].w >> 5));
  }
#pragma unroll
  for (int k = 0; k < kMOrderPer; ++k) {
    const uint32_t t = (uint32_t)chunk * kMOrderChunk + k * kMT + threadIdx.x;
    if (t < ncell) {
      const uint32_t vid = bp[k].y + __popc(bp[k].x & ((1u << (raw[k].w & 31)) - 1u));
      if (vid < (uint32_t)max_voxels) vcell[vid] = raw[k];  // voxelization_cpu.cpp:78
    }
  }
}

// ------------------------------------------------------------------------------------------
// stage D: expansion of voxels [item * 1024, +1024) of frame f, in voxel-id order
// ------------------------------------------------------------------------------------------
// P, C fixed: lane = voxel; its cell record, list entries and rows are loaded back to back
// (independent loads in flight), staged in shared memory, and the 32-voxel tile leaves as a
// float4 stream.  shared memory: 8 warps x 32 x P x C words.
template <int C, int PT>
__device__ __forceinline__ void stage_expand_fixed(const HvFrame& fr, const MegaWork& w, const KeyDecode& kd,
                                                   const int f, const int item, const int m, float* smem) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* stage = smem + wid * (32 * PT * C);
  const uint4* vcell = reinterpret_cast<const uint4*>(w.vcell(f));
  const uint32_t* lst = w.lst(f);
  const float* __restrict__ pts = fr.pts;
  const uint32_t n = (uint32_t)fr.n;
  float* st = stage + lane * (PT * C);
#pragma unroll 1
  for (int it = 0; it < kMExpTiles; ++it) {
    const int v0 = item * kMExpVox + (wid * kMExpTiles + it) * 32;
    if (v0 >= m) break;  // warp-uniform
    const int nvox = min(32, m - v0);
    uint4 cl = make_uint4(0u, 0u, 0u, 0u);  // key, len, list_off, first
    if (lane < nvox) cl = __ldcg(vcell + v0 + lane);
    const uint32_t len = min(cl.y, (uint32_t)PT);
    uint32_t idx[PT];
#pragma unroll
    for (int j = 0; j < PT; ++j) idx[j] = (uint32_t)j < len ? __ldcg(lst + cl.z + j) : kEmpty;
    if (lane < nvox) {
      const uint32_t cz = div_small_err(cl.x, kd.plane, kd.m_plane);
      const uint32_t rem = cl.x - cz * kd.plane;
      const uint32_t cy = div_small_err(rem, kd.gx, kd.m_gx);
      int32_t* co = fr.coors + (uint32_t)(v0 + lane) * 3u;
      co[0] = (int32_t)cz;
      co[1] = (int32_t)cy;
      co[2] = (int32_t)(rem - cy * kd.gx);
      fr.num[v0 + lane] = (int32_t)len;
    }
    float r[PT][C];
#pragma unroll
    for (int j = 0; j < PT; ++j) {
#pragma unroll
      for (int k = 0; k < C; ++k) r[j][k] = 0.0f;
      if (idx[j] != kEmpty) {
        if (C == 4) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(pts) + idx[j]);
          r[j][0] = a.x; r[j][1] = a.y; r[j][2] = a.z; r[j][3] = a.w;
        } else if (C == 5 && idx[j] + 1u < n) {
          // words [5 idx, 5 idx + 5) lie inside the two aligned 16-byte chunks starting at word
          // (5 idx) & ~3; idx + 1 < n keeps the second chunk inside the buffer
          const uint32_t w0 = idx[j] * 5u;
          const float4* p4 = reinterpret_cast<const float4*>(pts) + (w0 >> 2);
          const float4 a = __ldg(p4), b = __ldg(p4 + 1);
          const uint32_t o = w0 & 3u;
          const bool o1 = o & 1u, o2 = o & 2u;
          const float t0 = o1 ? a.y : a.x, t1 = o1 ? a.z : a.y, t2 = o1 ? a.w : a.z, t3 = o1 ? b.x : a.w;
          const float t4 = o1 ? b.y : b.x, t5 = o1 ? b.z : b.y, t6 = o1 ? b.w : b.z;
          r[j][0] = o2 ? t2 : t0; r[j][1] = o2 ? t3 : t1; r[j][2] = o2 ? t4 : t2;
          r[j][3 % C] = o2 ? t5 : t3; r[j][4 % C] = o2 ? t6 : t4;
        } else {
          const float* __restrict__ src = pts + (size_t)idx[j] * C;
#pragma unroll
          for (int k = 0; k < C; ++k) r[j][k] = __ldg(src + k);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < PT; ++j)
#pragma unroll
      for (int k = 0; k < C; ++k) st[j * C + k] = r[j][k];
    __syncwarp();
    const uint32_t w0 = (uint32_t)v0 * (PT * C);
    float* __restrict__ dst = fr.voxels + w0;
    const int nwords = nvox * (PT * C);
    const int n4 = nwords >> 2;  // w0 % 4 == 0 because v0 % 32 == 0; buffers are 16-byte aligned
    for (int i = lane; i < n4; i += 32)
      __stcs(reinterpret_cast<float4*>(dst) + i, reinterpret_cast<const float4*>(stage)[i]);
    for (int i = (n4 << 2) + lane; i < nwords; i += 32) dst[i] = stage[i];
    __syncwarp();
  }
}

// any (P, C) with 32 * P * C <= 1024 words per warp tile
__device__ __forceinline__ void stage_expand_rt(const HvFrame& fr, const MegaWork& w, const GridParams& g,
                                                const int c, const int p, const int f, const int item,
                                                const int m, float* smem) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* stage = smem + wid * kMExpStageWords;
  const uint4* vcell = reinterpret_cast<const uint4*>(w.vcell(f));
  const uint32_t* lst = w.lst(f);
  const float* __restrict__ pts = fr.pts;
  const int tile_words = 32 * p * c;
#pragma unroll 1
  for (int it = 0; it < kMExpTiles; ++it) {
    const int v0 = item * kMExpVox + (wid * kMExpTiles + it) * 32;
    if (v0 >= m) break;
    const int nvox = min(32, m - v0);
    uint4 cl = make_uint4(0u, 0u, 0u, 0u);
    if (lane < nvox) cl = __ldcg(vcell + v0 + lane);
    for (int i = lane; i < (tile_words + 3) / 4; i += 32)
      reinterpret_cast<float4*>(stage)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    const uint32_t len = min(cl.y, (uint32_t)p);
    if (lane < nvox) {
      decode_key(cl.x, g, fr.coors + (size_t)(v0 + lane) * 3);
      fr.num[v0 + lane] = (int32_t)len;
    }
    float* vstage = stage + (size_t)lane * p * c;
    for (uint32_t s = 0; s < len; ++s) {
      const uint32_t idx = __ldcg(lst + cl.z + s);
      const float* __restrict__ src = pts + (size_t)idx * c;
      for (int j = 0; j < c; ++j) vstage[s * c + j] = __ldg(src + j);
    }
    __syncwarp();
    float* __restrict__ dst = fr.voxels + (size_t)v0 * p * c;
    const int nwords = nvox * p * c;
    for (int i = lane; i < nwords; i += 32) dst[i] = stage[i];
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// the persistent kernel
// ------------------------------------------------------------------------------------------
#ifndef PCFE_MEGA_MINB
#define PCFE_MEGA_MINB 4
#endif
template <int C, int P, int PT>
__global__ void __launch_bounds__(kMT, PCFE_MEGA_MINB)
hvm_kernel(const __grid_constant__ HvBatch batch, const __grid_constant__ MegaWork w, const GridParams g, const KeyDecode kd,
           const int c_rt, const int p_rt, const int max_voxels, int32_t* __restrict__ voxel_num,
           unsigned long long* __restrict__ stats /* debug: per stage {items, wait, work} cycles */) {
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ uint16_t s_table[kMMaxItems];  // merged item order of one step
  __shared__ uint32_t warp_sums[33];
  __shared__ uint32_t s_misc[4];
  __shared__ uint32_t s_ticket;
  __shared__ int s_arg;
  long long waited = 0;
  const int tid = threadIdx.x;
  const int c = C > 0 ? C : c_rt;
  const int p = P > 0 ? P : p_rt;

  // Item order inside a step: item i of a stage with n items sits at position (2i + 1) / (2n);
  // stages are merged by position (ties: lower stage first), which interleaves them evenly.
  for (int r = tid; r < w.q; r += kMT) {
    int s = 0, i = r;
    while (i >= w.cnt[s]) { i -= w.cnt[s]; ++s; }
    const int ns = w.cnt[s];
    int rank = i;
    for (int s2 = 0; s2 < kStages; ++s2) {
      if (s2 == s || w.cnt[s2] == 0) continue;
      // items j of s2 before (s, i): (2j + 1) * ns < (2i + 1) * n2, or <= when s2 < s
      const int a = (2 * i + 1) * w.cnt[s2];
      const int mm = s2 < s ? a / ns : (a - 1) / ns;
      rank += min((mm + 1) >> 1, w.cnt[s2]);
    }
    s_table[rank] = (uint16_t)((s << 12) | i);
  }
  if (tid == 0) s_ticket = atomicAdd(w.ticket, 1u);
  __syncthreads();

  const uint32_t total = (uint32_t)w.steps * (uint32_t)w.q;
  const long long t_begin = stats ? clock64() : 0;
  long long* wp = stats ? &waited : nullptr;
  while (true) {
    waited = 0;
    const uint32_t t = s_ticket;
    __syncthreads();
    if (t >= total) break;
    uint32_t next = 0;
    if (tid == 0) next = atomicAdd(w.ticket, 1u);  // lands while this item runs
    const int step = (int)(t / (uint32_t)w.q);
    const uint32_t code = s_table[t - (uint32_t)step * (uint32_t)w.q];
    const int stage = (int)(code >> 12), idx = (int)(code & 0xFFFu);
    const int f = step - w.off[stage];
    long long t0 = 0, t1 = 0;
    if (stats && tid == 0) t0 = clock64();
    if (f >= 0 && f < w.frames) {
      uint32_t* ctl = w.cl(f) + w.nb;
      const HvFrame& fr = batch.f[f];
      if (stage == kSBin) {
        if (tid == 0 && f >= w.ring) wait_ge(w.cl(f - w.ring) + w.nb + cBucketDone, (uint32_t)w.nb, wp);
        __syncthreads();
        stage_bin(fr, w, g, c, f, idx, smem, warp_sums);
        signal_done(&ctl[cBinDone]);
      } else if (stage == kSBucket) {
        if (tid == 0) {
          wait_ge(&ctl[cBinDone], (uint32_t)w.cnt[kSBin], wp);
          if (f >= w.ring) wait_ge(w.cl(f - w.ring) + w.nb + cExpandDone, (uint32_t)w.cnt[kSExpand], wp);
        }
        __syncthreads();
        stage_bucket<PT>(w, p, f, idx, smem, warp_sums, s_misc);
        __syncthreads();
        if (tid == 0) {
          __threadfence();
          s_arg = (atomicAdd(&ctl[cBucketDone], 1u) == (uint32_t)w.nb - 1u) ? 1 : 0;
        }
        __syncthreads();
        if (s_arg) {  // last bucket of the frame: every first-point flag is set
          __threadfence();
          if (!__ldcg(&ctl[cOverflow])) frame_scan(w, f, max_voxels, voxel_num, smem, warp_sums);
          __syncthreads();
          if (tid == 0) {
            __threadfence();
            atomicExch(&ctl[cScanDone], 1u);
          }
        }
      } else if (stage == kSOrder) {
        if (tid == 0) {
          wait_ge(&ctl[cScanDone], 1u, wp);
          const uint32_t ncell = __ldcg(&ctl[cOverflow]) ? 0u : min(__ldcg(&ctl[cCell]), w.arena_cap);
          s_arg = (int)ncell;
        }
        __syncthreads();
        const uint32_t ncell = (uint32_t)s_arg;
        if ((uint32_t)idx * kMOrderChunk < ncell) stage_order(w, f, idx, ncell, max_voxels);
        signal_done(&ctl[cOrderDone]);
      } else {
        if (tid == 0) {
          wait_ge(&ctl[cScanDone], 1u, wp);
          int m = __ldcg(&ctl[cOverflow]) ? 0 : __ldcg(&voxel_num[f]);
          if (idx * kMExpVox >= m) m = 0;
          else wait_ge(&ctl[cOrderDone], (uint32_t)w.cnt[kSOrder], wp);
          s_arg = m;
        }
        __syncthreads();
        const int m = s_arg;
        if (m > 0) {
          if (C > 0) stage_expand_fixed<(C > 0 ? C : 4), (P > 0 ? P : 1)>(fr, w, kd, f, idx, m, reinterpret_cast<float*>(smem));
          else stage_expand_rt(fr, w, g, c, p, f, idx, m, reinterpret_cast<float*>(smem));
        }
        signal_done(&ctl[cExpandDone]);
      }
    }
    __syncthreads();
    if (stats && tid == 0 && f >= 0 && f < w.frames) {
      t1 = clock64();
      atomicAdd(&stats[stage * 4 + 0], 1ull);
      atomicAdd(&stats[stage * 4 + 1], (unsigned long long)waited);
      atomicAdd(&stats[stage * 4 + 2], (unsigned long long)(t1 - t0));
    }
    if (tid == 0) s_ticket = next;
    __syncthreads();
  }
  if (stats && tid == 0) atomicAdd(&stats[16], (unsigned long long)(clock64() - t_begin));
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// Bytes of workspace the pipeline needs for `frames` frames with `ring` scratch regions.
static size_t mega_bytes(const HvBucketPlan& p, int frames, int ring, size_t* small_per_frame) {
  const size_t ctl_b = align256((size_t)(p.nb + cWords) * sizeof(uint32_t));
  const size_t small = 3 * p.word_b + ctl_b;  // bitmask + pairs + ctl
  if (small_per_frame) *small_per_frame = small;
  return (size_t)ring * p.region_b + (size_t)frames * small + 256;
}

bool hvm_eligible(const pcfe_frame_t* frames, int num_frames, int c, const HvBucketPlan& p,
                  int max_points, int max_voxels, size_t workspace_bytes, int* ring_out) {
  if (max_points < 1 || max_points > 8 || 32 * max_points * c > kMExpStageWords) return false;
  if (max_voxels >= (1 << 24) || p.nb > 1024) return false;
  for (int k = 0; k < num_frames; ++k)
    if (((uintptr_t)frames[k].voxels & 15) || ((uintptr_t)frames[k].points & 15)) return false;
  const int chunk = std::min(num_frames, kMaxWave);
  int ring = std::min(std::max(g_opt_mega_ring, 1), chunk);
  while (ring > 1 && mega_bytes(p, chunk, ring, nullptr) > workspace_bytes) --ring;
  if (mega_bytes(p, chunk, ring, nullptr) > workspace_bytes) return false;
  if (ring < chunk && ring <= g_opt_mega_d3) return false;  // ring reuse needs ring > d3 (tickets)
  *ring_out = ring;
  return true;
}

int hvm_run(const pcfe_frame_t* frames, int num_frames, int c, const HvBucketPlan& p, int max_points,
            int max_voxels, int32_t* voxel_num, void* workspace, int ring, int device,
            cudaStream_t st) {
  MegaWork w;
  size_t small = 0;
  const int chunk = std::min(num_frames, kMaxWave);
  mega_bytes(p, chunk, ring, &small);
  const size_t ctl_b = small - 3 * p.word_b;
  char* base = (char*)workspace;
  w.region = base;
  w.region_stride = p.region_b;
  w.lst_off = p.ent_b;
  w.cells_off = p.ent_b + p.lst_b;
  w.vcell_off = p.ent_b + p.lst_b + p.cells_b;
  char* per = base + (size_t)ring * p.region_b;
  w.bitmask = (uint32_t*)per;
  w.pairs = (uint32_t*)(per + (size_t)chunk * p.word_b);
  w.ctl = (uint32_t*)(per + (size_t)chunk * 3 * p.word_b);
  w.ticket = (uint32_t*)(per + (size_t)chunk * small);
  w.word_stride = p.word_b / sizeof(uint32_t);
  w.ctl_stride = ctl_b / sizeof(uint32_t);
  w.ring = ring;
  w.words = p.words;
  w.nb = p.nb; w.log2_nb = p.log2_nb; w.cap = p.cap; w.slots = p.slots; w.log2_slots = p.log2_slots;
  w.arena_cap = (uint32_t)p.npad;

  KeyDecode kd;
  kd.plane = (uint32_t)p.g.gx * (uint32_t)p.g.gy;
  kd.gx = (uint32_t)p.g.gx;
  kd.m_plane = (uint32_t)(0x100000000ull / kd.plane);
  kd.m_gx = (uint32_t)(0x100000000ull / kd.gx);

  // dynamic shared memory: the largest stage
  const size_t smem_bin = (size_t)(((3 * p.nb + 3) & ~3) + 2 * kMBinTile) * 4;
  const size_t smem_bucket = (size_t)(2 * p.slots + p.cap) * 4 + (size_t)(2 * p.cap) * 2;
  const size_t smem_scan = (size_t)kMScanChunk * 4;
  const bool fixed = max_points == 5 && (c == 4 || c == 5);
  const size_t smem_exp = fixed ? (size_t)kMWarps * 32 * max_points * c * 4 : (size_t)kMWarps * kMExpStageWords * 4;
  const size_t smem = std::max(std::max(smem_bin, smem_bucket), std::max(smem_scan, smem_exp));

  const void* fn;
  if (fixed && c == 4) fn = (const void*)hvm_kernel<4, 5, 5>;
  else if (fixed) fn = (const void*)hvm_kernel<5, 5, 5>;
  else if (max_points <= 5) fn = (const void*)hvm_kernel<0, 0, 5>;
  else fn = (const void*)hvm_kernel<0, 0, 8>;
  PCFE_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0, sms = 0;
  PCFE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kMT, smem));
  PCFE_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  if (per_sm < 1) return PCFE_ERR_TOO_LARGE;
  if (g_opt_mega_ctas > 0) per_sm = std::min(per_sm, g_opt_mega_ctas);

  for (int f0 = 0; f0 < num_frames; f0 += kMaxWave) {
    const int wv = std::min(kMaxWave, num_frames - f0);
    HvBatch b;
    int64_t n_max = 0;
    for (int k = 0; k < wv; ++k) {
      const pcfe_frame_t& fr = frames[f0 + k];
      b.f[k] = HvFrame{fr.points, fr.voxels, fr.coors, fr.num_points, (int)fr.n, 0};
      n_max = std::max(n_max, fr.n);
    }
    w.frames = wv;
    const int64_t vmax = std::max<int64_t>(std::min<int64_t>(max_voxels, n_max), 1);
    w.cnt[kSBin] = (int)std::max<int64_t>((std::max<int64_t>(n_max, (int64_t)p.words * 32) + kMBinTile - 1) / kMBinTile, 1);
    w.cnt[kSBucket] = p.nb;
    w.cnt[kSOrder] = (int)((std::max<int64_t>(n_max, 1) + kMOrderChunk - 1) / kMOrderChunk);
    w.cnt[kSExpand] = (int)((vmax + kMExpVox - 1) / kMExpVox);
    w.off[kSBin] = 0;
    w.off[kSBucket] = std::max(g_opt_mega_d1, 1);
    w.off[kSOrder] = std::max(g_opt_mega_d2, w.off[kSBucket]);
    w.off[kSExpand] = std::max(g_opt_mega_d3, w.off[kSOrder]);
    if (wv > ring && ring <= w.off[kSExpand]) return PCFE_ERR_WORKSPACE;
    w.q = 0;
    for (int s = 0; s < kStages; ++s) {
      if (w.cnt[s] >= 4096) return PCFE_ERR_TOO_LARGE;
      w.q += w.cnt[s];
    }
    if (w.q > kMMaxItems) return PCFE_ERR_TOO_LARGE;
    w.steps = wv + w.off[kSExpand];
    {
      ProfScope ps("memset_ctl", st);
      PCFE_CUDA_TRY(cudaMemsetAsync(w.ctl, 0, (size_t)chunk * ctl_b + 256, st));
      count_launch();
    }
    {
      ProfScope ps("hvm_pipeline", st);
      const unsigned grid = (unsigned)std::min<int64_t>((int64_t)per_sm * sms, (int64_t)w.steps * w.q);
      int32_t* vn = voxel_num + f0;
      unsigned long long* stats = nullptr;
      if (g_opt_mega_stats) {
        PCFE_CUDA_TRY(cudaMalloc(&stats, 32 * sizeof(unsigned long long)));
        PCFE_CUDA_TRY(cudaMemsetAsync(stats, 0, 32 * sizeof(unsigned long long), st));
      }
      void* args[] = {(void*)&b, (void*)&w, (void*)&p.g, (void*)&kd, (void*)&c, (void*)&max_points,
                      (void*)&max_voxels, (void*)&vn, (void*)&stats};
      PCFE_CUDA_TRY(cudaLaunchKernel(fn, dim3(grid), dim3(kMT), args, smem, st));
      PCFE_LAUNCH_CHECK();
      if (stats) {
        unsigned long long h[32];
        PCFE_CUDA_TRY(cudaStreamSynchronize(st));
        PCFE_CUDA_TRY(cudaMemcpy(h, stats, sizeof h, cudaMemcpyDeviceToHost));
        cudaFree(stats);
        static const char* nm[] = {"bin", "bucket", "order", "expand"};
        fprintf(stderr, "[hvm] grid %u (%d/SM) q %d steps %d  cta-cycles %.3g\n", grid, per_sm, w.q, w.steps, (double)h[16]);
        for (int s = 0; s < kStages; ++s)
          fprintf(stderr, "[hvm] %-7s items %7llu  wait %6.2f%%  total %6.2f%% of cta-cycles  (%.0f cyc/item, wait %.0f)\n",
                  nm[s], h[s * 4], 100.0 * h[s * 4 + 1] / h[16], 100.0 * h[s * 4 + 2] / h[16],
                  (double)h[s * 4 + 2] / std::max<unsigned long long>(h[s * 4], 1),
                  (double)h[s * 4 + 1] / std::max<unsigned long long>(h[s * 4], 1));
      }
    }
    int rc = hvg_launch_slow_ring(b, wv, ring, w.ctl + p.nb + cOverflow, w.ctl_stride,
                                  g_opt_force_overflow, w.region, w.region_stride, p.slow, w.bitmask,
                                  w.word_stride, w.pairs, 2 * w.word_stride, c, max_points,
                                  max_voxels, voxel_num + f0, st);
    if (rc != PCFE_OK) return rc;
  }
  return PCFE_OK;
}

}  // namespace pcfe
