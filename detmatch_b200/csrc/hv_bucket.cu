// detmatch_b200/csrc/hv_bucket.cu -- hard voxelization, shared-memory bucket path (the fast path).
//
// Behaviour reproduced bit for bit: mmdet3d/ops/voxel/src/voxelization_cpu.cpp:43-99.
//
// Measured on B200 (tests/native/microbench.cu, mb_emit.cu): random L2 atomics 135 Gops/s,
// shared-memory atomics 1500 Gops/s chip-wide, partial-sector scattered L2 writes 3-7x the cost of
// full sectors.  So points are first partitioned by a hash of their cell key into NB buckets per
// frame (streaming, coalesced), all grouping work -- cell de-duplication, the first P point indices
// of every cell in ascending order -- happens in the shared memory of the one CTA that owns a
// bucket, and nothing is ever scattered by voxel id.  The only frame-wide structure is a bit per
// point, "is the first point of its voxel": the popcount rank of that bit IS the voxel id
// (first-occurrence order without sorting).
//
// Record path (max_points == 5, C = 4 or 5, 16-byte aligned buffers: KITTI / Waymo voxels):
//   hvb_zero          counters and mask words of the wave
//   hvb_bin           rows -> cell key (hoisted-reciprocal exact division) -> bucket = top hash
//                     bits; a 4096-point tile is counting-sorted by bucket in shared memory and
//                     written as contiguous runs of (key, point index) entries (one L2 atomicAdd
//                     per (tile, bucket))
//   hvb_bucket_rec    one CTA per (frame, bucket): TMA-staged entries, open-addressing table, one
//                     chain per cell, register sort of the 5 smallest indices; a cell with more
//                     than one point writes ONE 16-byte record at rec[first point]; one 64-bit
//                     atomicOr sets "first point" and "has a record"
//   hvb_scan_firsts   voxel id -> first point index (+ has-record bit): coalesced compaction of the
//                     mask; voxel_num = min(#cells, V)
//   hvb_expand_rec    voxel-id order, pipelined over 32-voxel tiles (firsts -> records -> rows),
//                     lane = output word, 128-byte coalesced streaming stores, coordinates
//                     recomputed from the first point
//   fallback          (hv_global.cu, one CTA per frame) for frames whose bucket regions overflowed
//                     (heavy duplication / adversarial keys); a no-op launch otherwise
// The launches after hvb_zero are programmatic dependents of their predecessor.
//
// General path (any max_points / C / alignment): hvb_bin -> hvb_bucket_small<5|8> (P <= 8: chains +
// register sort) or hvb_bucket_rank (any P: bitonic sort of (cell, point) keys) -> cells + list arena ->
// hv_scan_flags -> hvb_order (cell -> voxel-id order) -> hvb_expand / _fixed / _pipe / _words.
#include <algorithm>
#include <mutex>

#include "hv_common.cuh"

namespace pcfe {

int hv_launch_scan(const uint32_t* bitmask, size_t bitmask_stride, uint32_t* prefix,
                   size_t prefix_stride, int words, int max_voxels, int32_t* voxel_num, int frames,
                   int paired, cudaStream_t st);
int hvg_launch_slow(const HvBatch& b, int frames, uint32_t* overflow, size_t overflow_stride,
                    int force, char* scratch_base, size_t scratch_stride, const HvGlobalPlan& p,
                    uint32_t* bitmask, size_t bitmask_stride, uint32_t* prefix, size_t prefix_stride,
                    int c, int max_points, int max_voxels, int32_t* voxel_num, cudaStream_t st, int mean,
                    int stage, const int32_t* vn_all, int f_first);

Knob g_opt_bucket_avg{1024};  // target points per bucket (tunable through pcfe_debug_set)
Knob g_opt_bucket_variant{0};  // 1: general kernels (cells + list arena) even where the record path applies
Knob g_opt_no_fast_div{0};      // 1: __fdiv_rn for every point (no hoisted reciprocal)
Knob g_opt_expand_variant{0};  // 1: un-pipelined fixed-P expansion kernel
Knob g_opt_expand_prefetch{1};  // frames of L2 prefetch distance in the expansion (0 = off)
Knob g_opt_pdl{1};              // programmatic dependent launch between the record path's kernels
Knob g_opt_expand_ctas{0};      // > 0: persistent expansion with this many CTAs per SM
Knob g_opt_expand_tiles{0};     // > 0: 32-voxel tiles per warp of the record expansion (default 3)
Knob g_opt_warp_dedup{0};       // 1: warp-level key de-duplication (__match_any_sync) in front of the bucket table
Knob g_opt_bin_small{2};        // partition tile: 0 = 4096 points, 1 = 1024 points, 2 = by batch size
Knob g_opt_overlap{1};          // 0: waves of a multi-wave batch run one after the other on the caller's stream
Knob g_opt_expand_map{2};       // record expansion, tiles of a warp: 0 consecutive, 1 round-robin inside the CTA, 2 round-robin over the frame (default: measured best)
Knob g_opt_cluster{0};          // 1: record path with one thread-block cluster per frame (hv_cluster.cuh) -- measured slower, see profiles/r02_cluster_*

namespace {

struct __align__(16) Cell {  // one occupied voxel of a frame
  uint32_t key;       // linear cell index
  uint32_t len;       // min(#points, P): entries in its list
  uint32_t list_off;  // offset of its ascending point-index list in the frame's list arena
  uint32_t first;     // its first point (== list entry 0)
};

struct HvbWork {
  char* region;          // [W] per-frame scratch: ent | lists | cells | vcell
  size_t region_stride;  // bytes
  size_t lst_off, cells_off, vcell_off;  // byte offsets inside a frame region
  size_t rec_off, firsts_off;            // record-at-first-point variant: ent | rec | firsts
  uint32_t* zero;        // [W] per-frame zeroed block: bitmask[2 * words] | ctl  (the record path keeps a
                         // second bit per point, "its cell has more points", in 64-bit words)
  size_t zero_stride;    // words
  size_t ctl_off;        // words: ctl = bucket_cnt[nb] | list_cursor | cell_cursor | overflow
  uint32_t* wordprefix;  // [W][2 * words]  {bitmask word, exclusive popcount prefix} pairs
  size_t word_stride;    // words
  int nb, log2_nb, cap, slots, log2_slots;
  uint32_t arena_cap;    // capacity of the list arena and of the cell array (entries)

  __device__ __forceinline__ uint2* ent(int f) const { return reinterpret_cast<uint2*>(region + (size_t)f * region_stride); }
  __device__ __forceinline__ uint32_t* lst(int f) const { return reinterpret_cast<uint32_t*>(region + (size_t)f * region_stride + lst_off); }
  __device__ __forceinline__ Cell* cells(int f) const { return reinterpret_cast<Cell*>(region + (size_t)f * region_stride + cells_off); }
  __device__ __forceinline__ Cell* vcell(int f) const { return reinterpret_cast<Cell*>(region + (size_t)f * region_stride + vcell_off); }
  __device__ __forceinline__ uint4* rec(int f) const { return reinterpret_cast<uint4*>(region + (size_t)f * region_stride + rec_off); }
  __device__ __forceinline__ uint32_t* firsts(int f) const { return reinterpret_cast<uint32_t*>(region + (size_t)f * region_stride + firsts_off); }
  __device__ __forceinline__ uint32_t* bitmask(int f) const { return zero + (size_t)f * zero_stride; }
  __device__ __forceinline__ uint32_t* ctl(int f) const { return zero + (size_t)f * zero_stride + ctl_off; }
  __device__ __forceinline__ uint32_t* prefix(int f) const { return wordprefix + (size_t)f * word_stride; }
};
// ctl word indices after the nb bucket counters
constexpr int kCtlList = 0, kCtlCell = 1, kCtlOverflow = 2;

// ------------------------------------------------------------------------------------------
// A: partition points into hash buckets
// ------------------------------------------------------------------------------------------
#ifndef PCFE_BIN_THREADS
#define PCFE_BIN_THREADS 256
#endif
#ifndef PCFE_BIN_PER_THREAD
#define PCFE_BIN_PER_THREAD 16
#endif
#ifndef PCFE_BIN_MINB
#define PCFE_BIN_MINB 4  // measured: 256 threads x 16 points beats 512 x 8 (0.0707 vs 0.0743 ms), 1024 x 4 loses
#endif
constexpr int kBinThreads = PCFE_BIN_THREADS;
constexpr int kBinPerThread = PCFE_BIN_PER_THREAD;
constexpr int kBinTile = kBinThreads * kBinPerThread;  // 4096 points
constexpr int kBinPerThreadSmall = 4;
constexpr int kBinTileSmall = kBinThreads * kBinPerThreadSmall;  // 1024 points
constexpr int kMaxBuckets = 1024;

__global__ void hvb_zero_kernel(uint4* __restrict__ p, const size_t n16) {
  // the partition kernel may start right away: its row loads and key computation need nothing from
  // here, and it waits (griddepcontrol.wait) for this grid's completion before its first counter atomic
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_uint4(0u, 0u, 0u, 0u);
}

// CT = features per point at compile time (row loads become base + immediate), 0 = run time.
// PT = points per thread: 16 (4096-point tiles) for batches that fill the SMs several times over, 4
// (1024-point tiles) for small batches, where one CTA's latency chain IS the kernel's duration
// (8 C4 frames: 14.8 us with 352 CTAs of 4096 points, profiles/r02_b8_launches.csv).
template <int CT, int PT>
__global__ void __launch_bounds__(kBinThreads, PT >= 16 ? PCFE_BIN_MINB : 6)
hvb_bin_kernel(const __grid_constant__ HvBatch batch, const HvbWork w, const GridParams g,
               const int c_rt, const int use_fast_div) {
  const int c = CT > 0 ? CT : c_rt;
  __shared__ uint32_t hist[kMaxBuckets];   // entries of this tile per bucket
  __shared__ uint32_t soff[kMaxBuckets];   // exclusive prefix of hist (staging offsets)
  __shared__ uint32_t delta[kMaxBuckets];  // (entry position in the frame's ent array) - (staging position)
  constexpr int kBinPerThread = PT, kBinTile = kBinThreads * PT;
  __shared__ uint2 stage[kBinTile];
  __shared__ uint32_t warp_sums[33];
  __shared__ uint32_t s_overflow;

  const int f = blockIdx.y;
  const int n = batch.f[f].n;
  const int tid = threadIdx.x;
  const int tile0 = blockIdx.x * kBinTile;
  if (tile0 >= n) return;
  // All row loads of the thread are issued before any arithmetic: one exposed DRAM latency per
  // tile.  ax/ay/az end up holding p - range_min.
  float ax[kBinPerThread], ay[kBinPerThread], az[kBinPerThread];
  {
    const float* __restrict__ p = batch.f[f].pts + (size_t)(tile0 + tid) * c;
    const int stride = kBinThreads * c;
    if (tile0 + kBinTile <= n) {
#pragma unroll
      for (int k = 0; k < kBinPerThread; ++k) {
        ax[k] = __ldg(p + k * stride);
        ay[k] = __ldg(p + k * stride + 1);
        az[k] = __ldg(p + k * stride + 2);
      }
    } else {
      const float qnan = __int_as_float(0x7FC00000);  // past the end: NaN fails every range test
#pragma unroll
      for (int k = 0; k < kBinPerThread; ++k) {
        const bool in = tile0 + tid + k * kBinThreads < n;
        ax[k] = in ? __ldg(p + k * stride) : qnan;
        ay[k] = in ? __ldg(p + k * stride + 1) : qnan;
        az[k] = in ? __ldg(p + k * stride + 2) : qnan;
      }
    }
  }
  for (int b = tid; b < w.nb; b += kBinThreads) hist[b] = 0;
  if (tid == 0) s_overflow = 0u;
  __syncthreads();

  // cell keys.  One branch per thread, not per point: when every difference of the thread is
  // inside the guard of the hoisted-reciprocal division (pcfe_common.cuh) the eight keys are
  // computed without any control flow; otherwise all eight take the plain IEEE divide.
  uint32_t key[kBinPerThread];
  bool guard_ok = use_fast_div != 0;
  uint32_t fmask = 0xFFFFFFFFu;  // bit k: point k passes the fused PointsRangeFilter
  if (g.filter) {
    fmask = 0u;
#pragma unroll
    for (int k = 0; k < kBinPerThread; ++k) fmask |= filter_pass(ax[k], ay[k], az[k], g) ? (1u << k) : 0u;
  }
#pragma unroll
  for (int k = 0; k < kBinPerThread; ++k) {
    ax[k] = __fsub_rn(ax[k], g.x0);
    ay[k] = __fsub_rn(ay[k], g.y0);
    az[k] = __fsub_rn(az[k], g.z0);
    guard_ok = guard_ok & fast_div_guard(ax[k]) & fast_div_guard(ay[k]) & fast_div_guard(az[k]);
  }
  if (guard_ok) {
    const FastAxes fa = make_fast_axes(g);
#pragma unroll
    for (int k = 0; k < kBinPerThread; ++k) {
      const float qx = fast_div(ax[k], g.vx, fa.rx), qy = fast_div(ay[k], g.vy, fa.ry), qz = fast_div(az[k], g.vz, fa.rz);
      // 0 <= q < 2^31 <=> bits(q) < bits(2^31) as unsigned; -0 and NaN cannot come out of the guard
      const uint32_t qmax = max(max(__float_as_uint(qx), __float_as_uint(qy)), __float_as_uint(qz));
      const int cx = __float2int_rz(qx), cy = __float2int_rz(qy), cz = __float2int_rz(qz);
      const bool ok = (qmax < 0x4F000000u) & (cx < g.gx) & (cy < g.gy) & (cz < g.gz);
      const uint32_t lin = ((uint32_t)cz * (uint32_t)g.gy + (uint32_t)cy) * (uint32_t)g.gx + (uint32_t)cx;
      key[k] = (ok && ((fmask >> k) & 1u)) ? lin : kEmpty;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kBinPerThread; ++k) {
      // voxelization_cpu.cpp:23-29 on the differences (axis_cell() without its subtract)
      const float qx = __fdiv_rn(ax[k], g.vx), qy = __fdiv_rn(ay[k], g.vy), qz = __fdiv_rn(az[k], g.vz);
      const bool in = (qx >= 0.0f) & (qx < 2147483648.0f) & (qy >= 0.0f) & (qy < 2147483648.0f) &
                      (qz >= 0.0f) & (qz < 2147483648.0f);
      const int cx = in ? __float2int_rz(qx) : -1, cy = in ? __float2int_rz(qy) : -1, cz = in ? __float2int_rz(qz) : -1;
      const bool ok = in & (cx < g.gx) & (cy < g.gy) & (cz < g.gz);
      const uint32_t lin = ((uint32_t)cz * (uint32_t)g.gy + (uint32_t)cy) * (uint32_t)g.gx + (uint32_t)cx;
      key[k] = (ok && ((fmask >> k) & 1u)) ? lin : kEmpty;
    }
  }
  // rank of every entry inside its (tile, bucket) run; rb = bucket << 16 | rank
  uint32_t rb[kBinPerThread];
  const int shift = 32 - w.log2_nb;
#pragma unroll
  for (int k = 0; k < kBinPerThread; ++k) {
    rb[k] = 0;
    if (key[k] != kEmpty) {
      const uint32_t b = w.log2_nb ? (key[k] * kGold) >> shift : 0u;
      rb[k] = (b << 16) | atomicAdd(&hist[b], 1u);
    }
  }
  __syncthreads();
  // reserve a run in every bucket this tile contributes to; scan hist for the staging offsets
  pdl_wait();  // the counters are zeroed by the previous kernel of the stream (rows are caller input)
  uint32_t* ctl = w.ctl(f);
  uint32_t total = 0;
  for (int b0 = 0; b0 < w.nb; b0 += kBinThreads) {
    const int b = b0 + tid;
    const uint32_t h = b < w.nb ? hist[b] : 0u;
    uint32_t gb = 0;
    if (h) gb = atomicAdd(&ctl[b], h);  // in flight during the block scan
    uint32_t tot;
    const uint32_t ex = block_exscan(h, warp_sums, &tot);
    if (b < w.nb) {
      soff[b] = total + ex;
      delta[b] = (uint32_t)b * (uint32_t)w.cap + gb - (total + ex);
      if (gb + h > (uint32_t)w.cap) {  // the frame takes the fallback; its entries are not needed
        ctl[w.nb + kCtlOverflow] = 1u;
        s_overflow = 1u;
      }
    }
    total += tot;
    __syncthreads();
  }
  if (s_overflow) return;
#pragma unroll
  for (int k = 0; k < kBinPerThread; ++k) {
    if (key[k] != kEmpty)
      stage[soff[rb[k] >> 16] + (rb[k] & 0xFFFFu)] = make_uint2(key[k], (uint32_t)(tile0 + k * kBinThreads + tid));
  }
  __syncthreads();
  uint2* __restrict__ ent = w.ent(f);
  for (uint32_t j = tid; j < total; j += kBinThreads) {
    const uint2 e = stage[j];
    const uint32_t b = w.log2_nb ? (e.x * kGold) >> shift : 0u;
    ent[delta[b] + j] = e;
  }
}

// ------------------------------------------------------------------------------------------
// B: one CTA per (frame, bucket), everything in shared memory
// ------------------------------------------------------------------------------------------
#ifndef PCFE_BUCKET_THREADS
#define PCFE_BUCKET_THREADS 256
#endif
constexpr int kBucketThreads = PCFE_BUCKET_THREADS;
#ifndef PCFE_REC_STRIDE
#define PCFE_REC_STRIDE 1  // uint4 units per record slot (2 = one 32-byte sector per slot)
#endif
#ifndef PCFE_TABLE_NUM
#define PCFE_TABLE_NUM 2  // table slots >= 2 x entries of the bucket (measured: 1.25x costs +8 % in probe retries)
#define PCFE_TABLE_DEN 1
#endif
constexpr int kMaxCap = 2048;  // entries per bucket (list offsets are packed into 16 bits)

// ------------------------------------------------------------------------------------------
// B': the same job for P <= PT (PT = 5 or 8): cells keep their points in a linked list built with
// one shared-memory atomicExch per entry; then one thread per cell walks the chain and keeps the
// P smallest point indices in sorted REGISTERS (branch-free min/max insertion), so there is no
// count pass, no second pass over the entries and no list array in shared memory.  Lists go to
// the arena straight from registers: consecutive cells write consecutive ranges.
// dynamic shared memory (words): hkey[S] | head[S] | eidx[cap] | enext[cap] (u16) | slotlist[cap] (u16)
// ------------------------------------------------------------------------------------------
template <int PT>
__global__ void __launch_bounds__(kBucketThreads)
hvb_bucket_small_kernel(const HvbWork w, const int pe /* 1 <= pe <= PT */) {
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ uint32_t warp_sums[33];
  __shared__ uint32_t s_nclaimed, s_list_base, s_cell_base;
  constexpr uint32_t kNil = 0xFFFFu;

  const int f = blockIdx.y, b = blockIdx.x, tid = threadIdx.x;
  uint32_t* ctl = w.ctl(f);
  if (ctl[w.nb + kCtlOverflow]) return;
  const int S = w.slots, cap = w.cap;
  uint32_t* hkey = smem;
  uint32_t* head = hkey + S;  // entry number of the most recently linked point of the cell
  uint32_t* eidx = head + S;
  uint16_t* enext = reinterpret_cast<uint16_t*>(eidx + cap);
  uint16_t* slotlist = enext + cap;

  const int ne = (int)min(ctl[b], (uint32_t)cap);
  if (ne == 0) return;
  {
    uint4* k4 = reinterpret_cast<uint4*>(hkey);  // hkey and head are contiguous: 2 * S words
    for (int s = tid; s < S / 2; s += kBucketThreads) k4[s] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
  }
  if (tid == 0) s_nclaimed = 0u;
  __syncthreads();

  const uint2* __restrict__ ent = w.ent(f) + (size_t)b * cap;
  const uint32_t smask = (uint32_t)S - 1u;
  const int sshift = 32 - w.log2_nb - w.log2_slots;
  const uint32_t lane_lt = (1u << (tid & 31)) - 1u;
  // warp-uniform trip count: every lane takes part in the ballot below
  for (int e0 = 0; e0 < ne; e0 += kBucketThreads) {
    const int e = e0 + tid;
    const bool valid = e < ne;
    uint2 en = make_uint2(0u, 0u);
    if (valid) en = __ldcs(&ent[e]);
    uint32_t s = ((en.x * kGold) >> sshift) & smask;
    bool claimed = false;
    if (valid) {
      // one CAS per probe: it either claims the slot, finds the cell, or reports a collision
      while (true) {
        const uint32_t old = atomicCAS(&hkey[s], kEmpty, en.x);
        claimed = old == kEmpty;
        if (claimed || old == en.x) break;
        s = (s + 1u) & smask;
      }
    }
    // claimed slots are appended to the cell list with one shared-memory atomic per warp
    const uint32_t cm = __ballot_sync(0xFFFFFFFFu, claimed);
    if (cm) {
      uint32_t base = 0;
      if ((tid & 31) == __ffs(cm) - 1) base = atomicAdd(&s_nclaimed, (uint32_t)__popc(cm));
      base = __shfl_sync(0xFFFFFFFFu, base, __ffs(cm) - 1);
      if (claimed) slotlist[base + __popc(cm & lane_lt)] = (uint16_t)s;
    }
    if (valid) {
      const uint32_t prev = atomicExch(&head[s], (uint32_t)e);
      enext[e] = (uint16_t)(prev == kEmpty ? kNil : prev);
      eidx[e] = en.y;
    }
  }
  __syncthreads();

  const int nv = (int)s_nclaimed;
  uint32_t* __restrict__ glst = w.lst(f);
  Cell* __restrict__ cells = w.cells(f);
  uint32_t* __restrict__ bitmask = w.bitmask(f);
  // kCellsPerThread cells per thread and round (cells tid, tid + T, ...): a typical bucket
  // (~370 cells) needs one round, i.e. one block scan and three barriers for the whole tail
  constexpr int kCellsPerThread = 2;
#pragma unroll 1
  for (int j0 = 0; j0 < nv; j0 += kCellsPerThread * kBucketThreads) {
    uint32_t sorted[kCellsPerThread][PT];
    uint32_t key[kCellsPerThread], len[kCellsPerThread];
    uint32_t mine = 0;
#pragma unroll
    for (int u = 0; u < kCellsPerThread; ++u) {
      const int j = j0 + u * kBucketThreads + tid;
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
    const int ncell = min(kCellsPerThread * kBucketThreads, nv - j0);
    if (tid == 0) {
      s_list_base = atomicAdd(&ctl[w.nb + kCtlList], tot);
      s_cell_base = atomicAdd(&ctl[w.nb + kCtlCell], (uint32_t)ncell);
    }
    __syncthreads();
    const uint32_t list_base = s_list_base, cell_base = s_cell_base;
    if (list_base + tot > w.arena_cap || cell_base + (uint32_t)ncell > w.arena_cap) {
      if (tid == 0) ctl[w.nb + kCtlOverflow] = 1u;  // cannot happen: arenas hold one entry per point
      return;
    }
#pragma unroll
    for (int u = 0; u < kCellsPerThread; ++u) {
      const int jl = u * kBucketThreads + tid;  // cell number inside this round
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
    __syncthreads();  // s_list_base / warp_sums are reused by the next round
  }
}

// ------------------------------------------------------------------------------------------
// B'': the record-at-first-point bucket kernel (P == 5).  Same grouping as B' -- open-addressing
// table, chains, register sort -- with three changes measured on B200:
//   * the bucket's entries arrive in shared memory with ONE TMA bulk copy (the per-thread global
//     loads exposed a DRAM/L2 latency per loop trip: 18 % of all stall samples);
//   * one CAS per probe (claims the slot, finds the cell or reports a collision) and a hand-written
//     warp-aggregated append to the cell list;
//   * every cell with more than one point leaves as one record at rec[first point]: no lists,
//     cursors or block scan.
// dynamic shared memory: ents[cap] (uint2) | hkey[S] | head[S] | slotlist[cap] (u16)
// ------------------------------------------------------------------------------------------
// DEDUP: entries of a warp that carry the same cell key are found with __match_any_sync BEFORE the table is
// touched: the group's lowest lane probes / claims the slot for all of them (one CAS loop per group), the
// group links itself into a private chain (lane -> next lower peer, no atomics) and its lowest lane pushes
// the whole chain with ONE atomicExch.  Pays when neighbouring entries share cells -- un-shuffled
// (sweep-ordered) frames: 46 % of the points of a 4096-point tile repeat a cell of the tile, and a tile's
// entries of one bucket sit next to each other -- and costs a few instructions per entry when they do not
// (PointShuffle'd frames: 0.03 % in-warp repeats); `hv_warp_dedup`, measured in profiles/r02_summary.md.
template <bool DEDUP>
__global__ void __launch_bounds__(kBucketThreads)
hvb_bucket_rec_kernel(const HvbWork w, const int pe /* 1 <= pe <= 5 */, const int spec /* entries copied before ne is known */) {
  constexpr int PT = 5;
  constexpr uint32_t kNil = 0xFFFFFFFFu;
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t s_nclaimed;

  // frames in REVERSE launch order: the bin kernel wrote the entries of the last frames most
  // recently (still in L2), and the records this kernel writes last (first frames) are the ones
  // the expansion reads first
  const int f = (int)(gridDim.y - 1u - blockIdx.y), b = blockIdx.x, tid = threadIdx.x;
  const uint32_t* ctl = w.ctl(f);
  const int cap = w.cap;
  uint2* ents = reinterpret_cast<uint2*>(smem);  // {key, point}; .x becomes the chain link once inserted
  uint32_t* hkey = smem + 2 * cap;
  const uint2* gent = w.ent(f) + (size_t)b * cap;
  // the first `spec` entries are requested before the bucket's fill count has arrived (one global
  // latency instead of two in front of the insert loop); the rest, if any, follows
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
  }
  __syncthreads();  // barrier objects initialised before any use (also keeps racecheck quiet)
  pdl_wait();       // everything below reads what the bin kernel wrote
  pdl_trigger();
  if (tid == 0) {
    mbar_expect_tx(&bar[0], (uint32_t)spec * 8u);
    bulk_g2s(ents, gent, (uint32_t)spec * 8u, &bar[0]);
    s_nclaimed = 0u;
  }
  const uint32_t overflow = ctl[w.nb + kCtlOverflow];
  const int ne = (int)min(ctl[b], (uint32_t)cap);
  // table size for THIS bucket: a power of two >= 2 ne, at most the plan's `slots` (>= 1.25 cap)
  const int want = max(PCFE_TABLE_NUM * ne / PCFE_TABLE_DEN, 64);
  const int S = min(1 << (32 - __clz(want - 1)), w.slots);  // next power of two
  uint32_t* head = hkey + S;
  uint16_t* slotlist = reinterpret_cast<uint16_t*>(hkey + 2 * w.slots);
  if (tid == 0 && ne > spec) {
    const uint32_t bytes = (uint32_t)(((ne - spec) + 1) & ~1) * 8u;  // 16-byte granules; cap and spec are even
    mbar_expect_tx(&bar[1], bytes);
    bulk_g2s(ents + spec, gent + spec, bytes, &bar[1]);
  }
  {
    uint4* k4 = reinterpret_cast<uint4*>(hkey);  // hkey and head are contiguous: 2 * S words
    for (int s = tid; s < S / 2; s += kBucketThreads) k4[s] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
  }
  __syncthreads();  // table and barriers initialised
  mbar_wait(&bar[0], 0);  // always: the copy must not outlive the CTA's shared memory
  if (overflow || ne == 0) return;
  if (ne > spec) mbar_wait(&bar[1], 0);

  const uint32_t smask = (uint32_t)S - 1u;
  const int sshift = 32 - w.log2_nb - w.log2_slots;
  const uint32_t lane = (uint32_t)tid & 31u;
  const uint32_t lane_lt = (1u << lane) - 1u;
  for (int e0 = 0; e0 < ne; e0 += kBucketThreads) {  // warp-uniform trip count
    const int e = e0 + tid;
    const bool valid = e < ne;
    const uint32_t key = valid ? ents[e].x : 0u;
    uint32_t s = ((key * kGold) >> sshift) & smask;
    bool claimed = false;
    uint32_t peers = 1u << lane;  // lanes of this warp whose entry has the same key (DEDUP)
    if (DEDUP) {  // (lanes past the end take part with a dummy value and are masked out of every group)
      const uint32_t vmask = __ballot_sync(0xFFFFFFFFu, valid);
      peers = __match_any_sync(0xFFFFFFFFu, valid ? key : (0xFFFFFF00u | lane)) & (valid ? vmask : ~0u);
    }
    const int leader = __ffs(peers) - 1;
    if (valid && (int)lane == leader) {
      while (true) {  // one CAS per probe: claims the slot, finds the cell, or reports a collision
        const uint32_t old = atomicCAS(&hkey[s], kEmpty, key);
        claimed = old == kEmpty;
        if (claimed || old == key) break;
        s = (s + 1u) & smask;
      }
    }
    if (DEDUP) s = __shfl_sync(0xFFFFFFFFu, s, leader);
    const uint32_t cm = __ballot_sync(0xFFFFFFFFu, claimed);
    if (cm) {  // claimed slots join the cell list: one shared-memory atomic per warp, issued by the
               // lane elect.sync picks (ptxas re-aggregates an atomic under an ordinary predicate)
      uint32_t leader, is_leader;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync %0|p, 0xffffffff;\n\tselp.u32 %1, 1, 0, p;\n\t}" : "=r"(leader), "=r"(is_leader));
      uint32_t base = 0;
      if (is_leader)
        asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(base) : "r"(smem_u32(&s_nclaimed)), "r"(__popc(cm)) : "memory");
      base = __shfl_sync(0xFFFFFFFFu, base, leader);
      if (claimed) slotlist[base + __popc(cm & lane_lt)] = (uint16_t)s;
    }
    if (valid) {
      if (!DEDUP) {
        ents[e].x = atomicExch(&head[s], (uint32_t)e);  // link: previous head of the cell, or kNil
      } else if ((int)lane == leader) {  // the group's chain: highest lane -> ... -> this lane -> previous head
        ents[e].x = atomicExch(&head[s], (uint32_t)(e - (int)lane + (31 - __clz(peers))));
      } else {
        ents[e].x = (uint32_t)(e - (int)lane + (31 - __clz(peers & lane_lt)));
      }
    }
  }
  __syncthreads();

  const int nv = (int)s_nclaimed;
  unsigned long long* __restrict__ bm64 = reinterpret_cast<unsigned long long*>(w.bitmask(f));
  uint4* __restrict__ rec = w.rec(f);
#pragma unroll 1
  for (int j = tid; j < nv; j += kBucketThreads) {
    uint32_t sorted[PT];
#pragma unroll
    for (int t = 0; t < PT; ++t) sorted[t] = kEmpty;
    const int s = slotlist[j];
    uint32_t cnt = 0;
    uint32_t e = head[s];
    // chain walk; the 5 smallest point indices stay in registers, ascending.  The first 5 steps are
    // peeled: step k inserts into a sorted prefix of k entries (k compare-exchanges instead of 5).
#pragma unroll
    for (int step = 0; step < PT; ++step) {
      if (e != kNil) {
        const uint2 en = ents[e];
        uint32_t v = en.y;
        e = en.x;
        ++cnt;
#pragma unroll
        for (int t = 0; t < step; ++t) {
          const uint32_t lo = min(sorted[t], v);
          v = max(sorted[t], v);
          sorted[t] = lo;
        }
        sorted[step] = v;
      }
    }
    while (e != kNil) {  // more than 5 points: full insertion, the largest of the six drops out
      const uint2 en = ents[e];
      uint32_t v = en.y;
      e = en.x;
      ++cnt;
#pragma unroll
      for (int t = 0; t < PT; ++t) {
        const uint32_t lo = min(sorted[t], v);
        v = max(sorted[t], v);
        sorted[t] = lo;
      }
    }
    // A cell with one kept point needs no record at all: bit `first` of the low mask word says
    // "first point of a voxel", the same bit of the high word says "the voxel has more points, see
    // rec[first]" -- one 64-bit atomic sets both.  rec[first] = {idx1, idx2, idx3, idx4} (kEmpty =
    // no point; slots >= pe stay empty) owns a whole 32-byte sector: the two stores leave the SM as
    // one full-sector write (a lone 16-byte store is a partial-sector L2 write, measured +12 %).
    // The voxel's coordinates are recomputed from its first point by the expansion.
    const uint32_t first = sorted[0];
    const bool more = cnt > 1u && pe > 1;
    if (more) {
      rec[PCFE_REC_STRIDE * (size_t)first] = make_uint4(sorted[1], (cnt > 2u && pe > 2) ? sorted[2] : kEmpty,
                                          (cnt > 3u && pe > 3) ? sorted[3] : kEmpty,
                                          (cnt > 4u && pe > 4) ? sorted[4] : kEmpty);
      if (PCFE_REC_STRIDE == 2) rec[2 * (size_t)first + 1] = make_uint4(0u, 0u, 0u, 0u);
    }
    atomicOr(&bm64[first >> 5], (1ull << (first & 31)) | (more ? (1ull << (32 + (first & 31))) : 0ull));
  }
}

// Ascending bitonic sort of N = 256 E keys held E per thread (thread t: elements t E .. t E + E - 1).
// Comparator distances below E stay in registers, below 32 E they are warp shuffles; only the
// remaining log2(N / (32 E)) (log2(N / (32 E)) + 1) / 2 stages go through shared memory (K, N words)
// with block barriers -- 6 of the 55 stages for N = 1024.  All comparators point the same way: the
// first stage of every merge pairs i with its mirror image inside the block (i ^ (kk - 1)), the
// following ones pair i with i ^ d.
// Only the in-register parts are unrolled; the merges across threads are ONE rolled loop over
// (kk, d) whose shuffle masks and partners are run-time values.  (Fully unrolled, the kernel was 7800
// instructions = 125 KB of code and 27 % of its stall samples were instruction fetches,
// profiles/r02_c5_ncu_summary.txt.)
template <int E>
__device__ __forceinline__ void bitonic_local_sort(uint32_t (&v)[E]) {  // merges kk = 2 .. E inside the thread
#pragma unroll
  for (int kk = 2; kk <= E; kk <<= 1) {
#pragma unroll
    for (int d = kk >> 1; d > 0; d >>= 1) {
      const bool flip = d == (kk >> 1);
#pragma unroll
      for (int r = 0; r < E; ++r) {
        const int q = flip ? (r ^ (kk - 1)) : (r ^ d);
        if (r < q) {
          const uint32_t a = v[r], b = v[q];
          v[r] = min(a, b);
          v[q] = max(a, b);
        }
      }
    }
  }
}
template <int E>
__device__ __forceinline__ void bitonic_local_merge(uint32_t (&v)[E]) {  // distances E / 2 .. 1 of a merge
#pragma unroll
  for (int d = E >> 1; d > 0; d >>= 1) {
#pragma unroll
    for (int r = 0; r < E; ++r) {
      const int q = r ^ d;
      if (r < q) {
        const uint32_t a = v[r], b = v[q];
        v[r] = min(a, b);
        v[q] = max(a, b);
      }
    }
  }
}
template <int E>
__device__ __forceinline__ void bitonic_sort_256(uint32_t (&v)[E], uint32_t* K, const int tid) {
  constexpr int T = kBucketThreads;
  bitonic_local_sort<E>(v);
#pragma unroll 1
  for (int tk = 2; tk <= T; tk <<= 1) {  // merge of blocks of tk threads (kk = tk E elements)
#pragma unroll 1
    for (int td = tk >> 1; td > 0; td >>= 1) {  // partner thread distance (d = td E)
      const bool flip = td == (tk >> 1);
      const int pmask = flip ? (tk - 1) : td;  // partner thread = tid ^ pmask
      const bool lower = !(tid & td);
      if (pmask < 32) {  // partner in another lane of this warp
        uint32_t o[E];
#pragma unroll
        for (int r = 0; r < E; ++r) {
          o[r] = __shfl_xor_sync(0xFFFFFFFFu, flip ? v[E - 1 - r] : v[r], pmask);
        }
#pragma unroll
        for (int r = 0; r < E; ++r) v[r] = lower ? min(v[r], o[r]) : max(v[r], o[r]);
      } else {  // partner in another warp: through shared memory
#pragma unroll
        for (int r = 0; r < E; ++r) K[tid * E + r] = v[r];
        __syncthreads();
        const int pt = tid ^ pmask;
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const uint32_t o = K[pt * E + (flip ? E - 1 - r : r)];
          v[r] = lower ? min(v[r], o) : max(v[r], o);
        }
        __syncthreads();
      }
    }
    bitonic_local_merge<E>(v);
  }
}

// ------------------------------------------------------------------------------------------
// B3: general max_points (pillars: P = 64).  No chains, no atomicMin lists, nothing whose cost depends
// on how the points are spread over the cells: every entry finds its cell's slot (one CAS per probe)
// and bumps the cell's count; a block scan over the cells (claim order) turns the counts into
// segment offsets; then the bucket's entries are SORTED as 32-bit keys (cell number << 21 | point
// index) by a bitonic network in shared memory.  Sorted position - segment offset is the entry's
// rank inside its cell: ranks < P are kept.  A 260-point pillar costs exactly as much as 260
// one-point cells (per-entry rank counting spent 82 % of its instructions on such pillars, at 12
// active lanes; atomicMin lists serialise on them).
// dynamic shared memory: ents[cap] (uint2: TMA destination; once the entries are in registers the
// same words hold the sort keys, then the output lists) | hkey[S] | hcell[S] (count, then cell
// number) | ccnt | coff | loff | slotlist (u16 x cap each)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBucketThreads)
hvb_bucket_rank_kernel(const HvbWork w, const int pe /* max(max_points, 1) */, const int spec,
                       const int slots /* table allocation: power of two > cap */) {
  constexpr int kPer = kMaxCap / kBucketThreads;  // entries per thread, at most
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t warp_sums[33];
  __shared__ uint32_t s_nclaimed, s_list_base, s_cell_base;

  const int f = (int)(gridDim.y - 1u - blockIdx.y), b = blockIdx.x, tid = threadIdx.x;
  uint32_t* ctl = w.ctl(f);
  const int cap = w.cap;
  uint2* ents = reinterpret_cast<uint2*>(smem);
  uint32_t* K = smem;  // sort keys, then the output lists (valid after pass 1)
  uint32_t* hkey = smem + 2 * cap;
  uint32_t* hcell = hkey + slots;
  uint16_t* ccnt = reinterpret_cast<uint16_t*>(hcell + slots);
  uint16_t* coff = ccnt + cap;
  uint16_t* loff = coff + cap;
  uint16_t* slotlist = loff + cap;
  const uint2* gent = w.ent(f) + (size_t)b * cap;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
  }
  __syncthreads();  // barrier objects initialised before any use
  if (tid == 0) {
    mbar_expect_tx(&bar[0], (uint32_t)spec * 8u);
    bulk_g2s(ents, gent, (uint32_t)spec * 8u, &bar[0]);
    s_nclaimed = 0u;
  }
  const uint32_t overflow = ctl[w.nb + kCtlOverflow];
  const int ne = (int)min(ctl[b], (uint32_t)cap);
  int S = 64;
  while (S < ne + (ne >> 2)) S <<= 1;
  S = min(S, slots);
  if (tid == 0 && ne > spec) {
    const uint32_t bytes = (uint32_t)(((ne - spec) + 1) & ~1) * 8u;
    mbar_expect_tx(&bar[1], bytes);
    bulk_g2s(ents + spec, gent + spec, bytes, &bar[1]);
  }
  for (int s = tid; s < S; s += kBucketThreads) {
    hkey[s] = kEmpty;
    hcell[s] = 0u;
  }
  __syncthreads();
  mbar_wait(&bar[0], 0);  // always: the copy must not outlive the CTA's shared memory
  if (overflow || ne == 0) return;
  if (ne > spec) mbar_wait(&bar[1], 0);

  // pass 1: slot of the entry's cell (one CAS per probe), arrival rank inside the cell
  const uint32_t smask = (uint32_t)S - 1u;
  const int sshift = 32 - w.log2_nb - w.log2_slots;
  const uint32_t lane = (uint32_t)tid & 31u;
  const uint32_t lane_lt = (1u << lane) - 1u;
  uint32_t myidx[kPer], mysr[kPer];
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    myidx[k] = 0u;
    mysr[k] = 0u;
    if (k * kBucketThreads < ne) {  // block-uniform: every lane takes part in the ballot
      const int e = k * kBucketThreads + tid;
      const bool valid = e < ne;
      uint2 en = make_uint2(0u, 0u);
      if (valid) en = ents[e];
      uint32_t s = ((en.x * kGold) >> sshift) & smask;
      bool claimed = false;
      if (valid) {
        while (true) {
          const uint32_t old = atomicCAS(&hkey[s], kEmpty, en.x);
          claimed = old == kEmpty;
          if (claimed || old == en.x) break;
          s = (s + 1u) & smask;
        }
      }
      const uint32_t cm = __ballot_sync(0xFFFFFFFFu, claimed);
      if (cm) {  // see hvb_bucket_rec_kernel
        uint32_t leader, is_leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync %0|p, 0xffffffff;\n\tselp.u32 %1, 1, 0, p;\n\t}" : "=r"(leader), "=r"(is_leader));
        uint32_t base = 0;
        if (is_leader)
          asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(base) : "r"(smem_u32(&s_nclaimed)), "r"(__popc(cm)) : "memory");
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (claimed) slotlist[base + __popc(cm & lane_lt)] = (uint16_t)s;
      }
      if (valid) {
        const uint32_t r = atomicAdd(&hcell[s], 1u);
        myidx[k] = en.y;
        mysr[k] = s | (r << 16);
      }
    }
  }
  __syncthreads();

  // per cell, in claim order: count -> offset of its entries in the regrouped array (low half) and
  // of its output list (high half: min(count, P) entries), both from ONE packed block scan
  const int nv = (int)s_nclaimed;
  uint32_t run = 0;
#pragma unroll 1
  for (int j0 = 0; j0 < nv; j0 += kBucketThreads) {
    const int j = j0 + tid;
    uint32_t cnt = 0;
    int s = 0;
    if (j < nv) {
      s = slotlist[j];
      cnt = hcell[s];
    }
    const uint32_t pack = (min(cnt, (uint32_t)pe) << 16) | cnt;
    uint32_t tot;
    const uint32_t ex = run + block_exscan(pack, warp_sums, &tot);
    if (j < nv) {
      ccnt[j] = (uint16_t)cnt;
      coff[j] = (uint16_t)(ex & 0xFFFFu);
      loff[j] = (uint16_t)(ex >> 16);
      hcell[s] = (uint32_t)j;
    }
    run += tot;
    __syncthreads();  // warp_sums is reused by the next chunk; hcell / coff visible below
  }
  const uint32_t total_list = run >> 16;
  if (tid == 0) {
    s_list_base = atomicAdd(&ctl[w.nb + kCtlList], total_list);
    s_cell_base = atomicAdd(&ctl[w.nb + kCtlCell], (uint32_t)nv);
  }

  // pass 2: sort keys (cell number << 21 | point index), padded to a power of two with kEmpty
  int N = kBucketThreads;
  while (N < ne) N <<= 1;  // <= 2048 = kMaxCap words: fits the entry buffer (2 cap words, cap >= 832)
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int e = k * kBucketThreads + tid;
    if (e < N) K[e] = e < ne ? ((hcell[mysr[k] & 0xFFFFu] << 21) | myidx[k]) : kEmpty;
  }
  __syncthreads();
  const uint32_t list_base = s_list_base, cell_base = s_cell_base;
  if (list_base + total_list > w.arena_cap || cell_base + (uint32_t)nv > w.arena_cap) {
    if (tid == 0) ctl[w.nb + kCtlOverflow] = 1u;  // cannot happen: arenas hold one entry per point
    return;
  }
  // sort: thread t takes elements t E .. t E + E - 1 (N = 256 E), see bitonic_sort_256
  uint32_t kv[kPer];
  const int E = N / kBucketThreads;  // 1, 2, 4 or 8 (block-uniform)
#pragma unroll
  for (int r = 0; r < kPer; ++r) kv[r] = r < E ? K[tid * E + r] : kEmpty;
  __syncthreads();  // the network reuses K
  if (E == 1) {
    uint32_t v[1] = {kv[0]};
    bitonic_sort_256<1>(v, K, tid);
    kv[0] = v[0];
  } else if (E == 2) {
    uint32_t v[2] = {kv[0], kv[1]};
    bitonic_sort_256<2>(v, K, tid);
    kv[0] = v[0]; kv[1] = v[1];
  } else if (E == 4) {
    uint32_t v[4] = {kv[0], kv[1], kv[2], kv[3]};
    bitonic_sort_256<4>(v, K, tid);
    kv[0] = v[0]; kv[1] = v[1]; kv[2] = v[2]; kv[3] = v[3];
  } else {
#if PCFE_BUCKET_THREADS == 256
    bitonic_sort_256<8>(kv, K, tid);
#endif
  }
  // pass 3: sorted position i of cell j is rank i - coff[j]; the first P of every cell are kept
  // (the network's last barrier is behind us: K is free for the lists)
#pragma unroll
  for (int r = 0; r < kPer; ++r) {
    const int i = tid * E + r;
    if (r < E && i < ne) {
      const uint32_t j = kv[r] >> 21;
      const uint32_t rank = (uint32_t)i - coff[j];
      if (rank < (uint32_t)pe) K[loff[j] + rank] = kv[r] & 0x1FFFFFu;
    }
  }
  __syncthreads();

  // emit: the CTA's lists are contiguous in shared memory in claim order, so the list arena gets
  // one coalesced copy; every cell becomes one 16-byte record; first points are flagged
  uint32_t* __restrict__ glst = w.lst(f) + list_base;
  for (uint32_t i = tid; i < total_list; i += kBucketThreads) glst[i] = K[i];
  Cell* __restrict__ cells = w.cells(f) + cell_base;
  uint32_t* __restrict__ bitmask = w.bitmask(f);
  for (int j = tid; j < nv; j += kBucketThreads) {
    Cell cl;
    cl.key = hkey[slotlist[j]];
    cl.len = min((uint32_t)ccnt[j], (uint32_t)pe);  // >= 1: every claimed cell has a point and pe >= 1
    cl.list_off = list_base + loff[j];
    cl.first = K[loff[j]];  // lists are ascending: entry 0 is the cell's first point
    cells[j] = cl;
    atomicOr(&bitmask[cl.first >> 5], 1u << (cl.first & 31));
  }
}

// ------------------------------------------------------------------------------------------
// D: voxel id of every cell = rank of its first point; vcell[vid] = cell
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
hvb_order_kernel(const HvbWork w, const int max_voxels) {
  const int f = blockIdx.y;
  const uint32_t* ctl = w.ctl(f);
  if (ctl[w.nb + kCtlOverflow]) return;
  const uint32_t ncell = min(ctl[w.nb + kCtlCell], w.arena_cap);
  const uint32_t t = blockIdx.x * 256u + threadIdx.x;
  if (t >= ncell) return;
  const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(w.cells(f)) + t);
  // wordprefix holds {bitmask word, exclusive prefix} pairs: one 8-byte random read per cell
  const uint2 bp = reinterpret_cast<const uint2*>(w.prefix(f))[raw.w >> 5];
  const uint32_t vid = bp.y + __popc(bp.x & ((1u << (raw.w & 31)) - 1u));
  if (vid < (uint32_t)max_voxels)  // voxelization_cpu.cpp:78
    reinterpret_cast<uint4*>(w.vcell(f))[vid] = raw;
}

// ------------------------------------------------------------------------------------------
// E: expansion in voxel-id order
// ------------------------------------------------------------------------------------------
// A warp owns tiles of VT consecutive voxels = VT * P * C contiguous output floats, staged in
// shared memory: the stage is zeroed (that is the padding), every REAL row (slot < len) is
// gathered into place -- C == 5 rows with two aligned 16-byte loads, C == 4 with one -- and the
// tile leaves as one float4 stream.  Padding rows (two thirds of the C4 output) cost no loads and
// a quarter of a store instruction per word.
#ifndef PCFE_EXP_WARPS
#define PCFE_EXP_WARPS 4
#endif
constexpr int kExpWarps = PCFE_EXP_WARPS;
constexpr int kExpThreads = kExpWarps * 32;
#ifndef PCFE_EXP_TILES
#define PCFE_EXP_TILES 2
#endif
constexpr int kExpTilesPerWarp = PCFE_EXP_TILES;
constexpr int kExpStageWords = 1024;  // per warp (a 32-voxel C4 tile is 800 words)

template <int C>
__device__ __forceinline__ void stage_row(const float* __restrict__ pts, uint32_t idx, int n, int c,
                                          float* st, bool vec_ok) {
  if (C == 4 && vec_ok) {
    *reinterpret_cast<float4*>(st) = __ldg(reinterpret_cast<const float4*>(pts) + idx);
#ifndef PCFE_EXP_SCALAR_ROWS
  } else if (C == 5 && vec_ok && idx + 1u < (uint32_t)n) {
    // words [5 idx, 5 idx + 5) lie inside the two aligned 16-byte chunks starting at word
    // (5 idx) & ~3; idx + 1 < n keeps the second chunk inside the buffer
    const uint32_t w0 = idx * 5u;
    const float4* p4 = reinterpret_cast<const float4*>(pts) + (w0 >> 2);
    const float4 a = __ldg(p4), b = __ldg(p4 + 1);
    const uint32_t o = w0 & 3u;
    const float r0 = a.x, r1 = a.y, r2 = a.z, r3 = a.w, r4 = b.x, r5 = b.y, r6 = b.z, r7 = b.w;
    st[0] = o == 0 ? r0 : o == 1 ? r1 : o == 2 ? r2 : r3;
    st[1] = o == 0 ? r1 : o == 1 ? r2 : o == 2 ? r3 : r4;
    st[2] = o == 0 ? r2 : o == 1 ? r3 : o == 2 ? r4 : r5;
    st[3] = o == 0 ? r3 : o == 1 ? r4 : o == 2 ? r5 : r6;
    st[4] = o == 0 ? r4 : o == 1 ? r5 : o == 2 ? r6 : r7;
#endif
  } else {
    const float* __restrict__ src = pts + (size_t)idx * c;
    for (int j = 0; j < c; ++j) st[j] = __ldg(src + j);
  }
}

#ifndef PCFE_EXP_CHUNK
#define PCFE_EXP_CHUNK 8
#endif
constexpr int kExpChunk = PCFE_EXP_CHUNK;  // list slots resolved per lane per round (loads in flight)

template <int C>
__global__ void __launch_bounds__(kExpThreads)
hvb_expand_kernel(const __grid_constant__ HvBatch batch, const HvbWork w, const GridParams g,
                  const int c_rt, const int max_points, const int vt,
                  const int32_t* __restrict__ voxel_num, const int vec_ok) {
  extern __shared__ __align__(16) float stage_all[];
  const int f = blockIdx.y;
  if (w.ctl(f)[w.nb + kCtlOverflow]) return;
  const HvFrame& fr = batch.f[f];
  const int c = C > 0 ? C : c_rt;
  const int m = voxel_num[f];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int tile_words = vt * max_points * c;
  float* stage = stage_all + (size_t)wid * kExpStageWords;
  const Cell* __restrict__ vcell = w.vcell(f);
  const uint32_t* __restrict__ lst = w.lst(f);
  const float* __restrict__ pts = fr.pts;
  // lane -> (voxel of the tile, slot group): vt voxels x (32 / vt) lanes each; a lane handles
  // slots sg, sg + 32/vt, ... of its voxel.  vt == 32: one lane per voxel, all its slots.
  const int v = lane & (vt - 1);
  const int sg = lane / vt;
  const int sstep = 32 / vt;

#pragma unroll 1
  for (int it = 0; it < kExpTilesPerWarp; ++it) {
    const int v0 = ((blockIdx.x * kExpWarps + wid) * kExpTilesPerWarp + it) * vt;
    if (v0 >= m) break;  // warp-uniform
    const int nvox = min(vt, m - v0);
    uint4 cl = make_uint4(0u, 0u, 0u, 0u);  // key, len, list_off, first
    if (v < nvox) cl = __ldg(reinterpret_cast<const uint4*>(vcell) + (v0 + v));
    // the zero padding (overlaps the latency of the cell load)
    for (int i = lane; i < (tile_words + 3) / 4; i += 32)
      reinterpret_cast<float4*>(stage)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    const uint32_t len = min(cl.y, (uint32_t)max_points);
    if (v < nvox && sg == 0) {
      decode_key(cl.x, g, fr.coors + (size_t)(v0 + v) * 3);
      fr.num[v0 + v] = (int32_t)len;
    }
    float* vstage = stage + (size_t)v * max_points * c;
    for (uint32_t s0 = sg; s0 < len; s0 += kExpChunk * sstep) {
      uint32_t idx[kExpChunk];
#pragma unroll
      for (int j = 0; j < kExpChunk; ++j) {  // all list loads of the round are issued together
        const uint32_t s = s0 + j * sstep;
        idx[j] = s < len ? __ldg(lst + cl.z + s) : kEmpty;
      }
#pragma unroll
      for (int j = 0; j < kExpChunk; ++j) {
        const uint32_t s = s0 + j * sstep;
        if (idx[j] != kEmpty) stage_row<C>(pts, idx[j], fr.n, c, vstage + (size_t)s * c, vec_ok != 0);
      }
    }
    __syncwarp();
    const size_t w0 = (size_t)v0 * max_points * c;
    float* __restrict__ dst = fr.voxels + w0;
    const int nwords = nvox * max_points * c;
    if (vec_ok && (w0 & 3) == 0) {
      const int n4 = nwords >> 2;
      for (int i = lane; i < n4; i += 32)
        __stcs(reinterpret_cast<float4*>(dst) + i, reinterpret_cast<const float4*>(stage)[i]);
      for (int i = (n4 << 2) + lane; i < nwords; i += 32) dst[i] = stage[i];
    } else {
      for (int i = lane; i < nwords; i += 32) dst[i] = stage[i];
    }
    __syncwarp();
  }
}

// ---- P fixed at compile time (P = 5: KITTI / Waymo voxels): no loops, no zero-fill pass ----------
// Lane = voxel.  The lane loads its cell record, then its PT list entries and the rows behind
// them (all independent: up to 2 * PT 16-byte loads in flight per lane), and writes all PT * C
// words of its voxel -- data or zeros -- into the warp's stage; the stage leaves as a float4
// stream.  Key decoding uses host-computed reciprocals instead of two 32-bit divisions.
struct KeyDecode {
  uint32_t plane, gx;          // gx * gy, gx
  uint32_t m_plane, m_gx;      // floor(2^32 / plane), floor(2^32 / gx)
};

// (a divisor of 1 would need m = 2^32: 2^32 - 1 is within the two correction steps as well)
inline KeyDecode make_key_decode(const GridParams& g) {
  KeyDecode kd;
  kd.plane = (uint32_t)g.gx * (uint32_t)g.gy;
  kd.gx = (uint32_t)g.gx;
  kd.m_plane = (uint32_t)std::min<uint64_t>(0x100000000ull / kd.plane, 0xFFFFFFFFull);
  kd.m_gx = (uint32_t)std::min<uint64_t>(0x100000000ull / kd.gx, 0xFFFFFFFFull);
  return kd;
}

__device__ __forceinline__ uint32_t div_small_err(uint32_t n, uint32_t d, uint32_t m) {
  uint32_t q = __umulhi(n, m);  // true quotient - 2 <= q <= true quotient
  uint32_t r = n - q * d;
  if (r >= d) { ++q; r -= d; }
  if (r >= d) { ++q; }
  return q;
}

template <int C, int PT>
__global__ void __launch_bounds__(kExpThreads)
hvb_expand_fixed_kernel(const __grid_constant__ HvBatch batch, const HvbWork w, const KeyDecode kd,
                        const int32_t* __restrict__ voxel_num, const int vec_ok) {
  __shared__ __align__(16) float stage_all[kExpWarps * 32 * PT * C];
  const int f = blockIdx.y;
  if (w.ctl(f)[w.nb + kCtlOverflow]) return;
  const HvFrame& fr = batch.f[f];
  const int m = voxel_num[f];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* stage = stage_all + wid * (32 * PT * C);
  const uint4* __restrict__ vcell = reinterpret_cast<const uint4*>(w.vcell(f));
  const uint32_t* __restrict__ lst = w.lst(f);
  const float* __restrict__ pts = fr.pts;
  const uint32_t n = (uint32_t)fr.n;
  float* st = stage + lane * (PT * C);

#pragma unroll 1
  for (int it = 0; it < kExpTilesPerWarp; ++it) {
    const int v0 = ((blockIdx.x * kExpWarps + wid) * kExpTilesPerWarp + it) * 32;
    if (v0 >= m) break;  // warp-uniform
    const int nvox = min(32, m - v0);
    uint4 cl = make_uint4(0u, 0u, 0u, 0u);  // key, len, list_off, first
    if (lane < nvox) cl = __ldg(vcell + v0 + lane);
    const uint32_t len = min(cl.y, (uint32_t)PT);
    uint32_t idx[PT];
#pragma unroll
    for (int j = 0; j < PT; ++j) idx[j] = (uint32_t)j < len ? __ldg(lst + cl.z + j) : kEmpty;
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
        if (C == 4 && vec_ok) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(pts) + idx[j]);
          r[j][0] = a.x; r[j][1] = a.y; r[j][2] = a.z; r[j][3] = a.w;
        } else if (C == 5 && vec_ok && idx[j] + 1u < n) {
          const uint32_t w0 = idx[j] * 5u;
          const float4* p4 = reinterpret_cast<const float4*>(pts) + (w0 >> 2);
          const float4 a = __ldg(p4), b = __ldg(p4 + 1);
          const uint32_t o = w0 & 3u;
          // shift the 8 loaded words left by o in two conditional steps (by 1, then by 2)
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
    if (vec_ok) {  // w0 % 4 == 0 because v0 % 32 == 0
      const int n4 = nwords >> 2;
      for (int i = lane; i < n4; i += 32)
        __stcs(reinterpret_cast<float4*>(dst) + i, reinterpret_cast<const float4*>(stage)[i]);
      for (int i = (n4 << 2) + lane; i < nwords; i += 32) dst[i] = stage[i];
    } else {
      for (int i = lane; i < nwords; i += 32) dst[i] = stage[i];
    }
    __syncwarp();
  }
}

// ---- the same tile, software-pipelined over the tiles of a warp ------------------------------------
// A tile needs three dependent round trips (cell record -> list entries -> rows).  Here a warp owns
// kPipeTiles consecutive tiles and keeps one round trip of each kind in flight at any time: while the
// rows of tile t are on their way it has already issued the list loads of tile t + 1 and the record
// load of tile t + 2, so the loop pays ONE memory latency per tile instead of three.
#ifndef PCFE_EXP_PIPE_TILES
#define PCFE_EXP_PIPE_TILES 4
#endif
constexpr int kPipeTiles = PCFE_EXP_PIPE_TILES;

template <int C, int PT>
__global__ void __launch_bounds__(kExpThreads)
hvb_expand_pipe_kernel(const __grid_constant__ HvBatch batch, const HvbWork w, const KeyDecode kd,
                       const int32_t* __restrict__ voxel_num, const int frames, const int pf_dist) {
  __shared__ __align__(16) float stage_all[kExpWarps * 32 * PT * C];
  const int f = blockIdx.y;
  // The row gathers below are random 32-byte sector reads; served from DRAM they waste most of
  // every burst.  So the CTAs expanding frame f pull frame f + pf_dist into L2 with sequential
  // bulk prefetches (one slice per CTA), and the gathers of that frame hit L2 later.
  if (pf_dist > 0 && f + pf_dist < frames && threadIdx.x < 32) {
    const HvFrame& nf = batch.f[f + pf_dist];
    const size_t total = ((size_t)nf.n * C * 4) & ~(size_t)15;
    const size_t slice = ((total + gridDim.x - 1) / gridDim.x + 511) & ~(size_t)511;  // 32 lanes x 16 B
    const size_t lo = (size_t)blockIdx.x * slice + (size_t)threadIdx.x * (slice / 32);
    if (lo < total) {
      const uint32_t bytes = (uint32_t)min(slice / 32, total - lo);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(nf.pts) + lo), "r"(bytes) : "memory");
    }
  }
  if (w.ctl(f)[w.nb + kCtlOverflow]) return;
  const HvFrame& fr = batch.f[f];
  const int m = voxel_num[f];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* stage = stage_all + wid * (32 * PT * C);
  const uint4* __restrict__ vcell = reinterpret_cast<const uint4*>(w.vcell(f));
  const uint32_t* __restrict__ lst = w.lst(f);
  const float* __restrict__ pts = fr.pts;
  float* st = stage + lane * (PT * C);
  const int vbase = (blockIdx.x * kExpWarps + wid) * (kPipeTiles * 32);
  if (vbase >= m) return;  // warp-uniform

  auto load_cell = [&](int v0) {
    uint4 cl = make_uint4(0u, 0u, 0u, 0u);  // key, len, list_off, first
    if (v0 + lane < m) cl = __ldg(vcell + v0 + lane);
    return cl;
  };
  uint4 cl_cur = load_cell(vbase);
  uint4 cl_nxt = load_cell(vbase + 32);
  uint32_t idx_cur[PT];
  {
    const uint32_t len = min(cl_cur.y, (uint32_t)PT);
#pragma unroll
    for (int j = 0; j < PT; ++j) idx_cur[j] = (uint32_t)j < len ? __ldg(lst + cl_cur.z + j) : kEmpty;
  }

#pragma unroll 1
  for (int it = 0; it < kPipeTiles; ++it) {
    const int v0 = vbase + it * 32;
    if (v0 >= m) break;  // warp-uniform
    const int nvox = min(32, m - v0);
    // rows of this tile
    float4 ra[PT], rb[PT];
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      ra[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      rb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx_cur[j] != kEmpty) {
        if (C == 4) {
          ra[j] = __ldg(reinterpret_cast<const float4*>(pts) + idx_cur[j]);
        } else {
          // words [5 idx, 5 idx + 5) lie inside the two aligned 16-byte chunks starting at word
          // (5 idx) & ~3.  Both chunks hold at least one word of the row, so both lie inside the
          // (16-byte aligned) buffer's allocation even for the last row.
          const uint32_t w0 = idx_cur[j] * 5u;
          const float4* p4 = reinterpret_cast<const float4*>(pts) + (w0 >> 2);
          ra[j] = __ldg(p4);
          rb[j] = __ldg(p4 + 1);
        }
      }
    }
    // list entries of the next tile, cell records of the one after
    uint32_t idx_nxt[PT];
    {
      const uint32_t len = min(cl_nxt.y, (uint32_t)PT);
#pragma unroll
      for (int j = 0; j < PT; ++j) idx_nxt[j] = (uint32_t)j < len ? __ldg(lst + cl_nxt.z + j) : kEmpty;
    }
    const uint4 cl_nn = load_cell(v0 + 64);
    // coordinates and count of this tile's voxels
    const uint32_t len = min(cl_cur.y, (uint32_t)PT);
    if (lane < nvox) {
      const uint32_t cz = div_small_err(cl_cur.x, kd.plane, kd.m_plane);
      const uint32_t rem = cl_cur.x - cz * kd.plane;
      const uint32_t cy = div_small_err(rem, kd.gx, kd.m_gx);
      int32_t* co = fr.coors + (uint32_t)(v0 + lane) * 3u;
      co[0] = (int32_t)cz;
      co[1] = (int32_t)cy;
      co[2] = (int32_t)(rem - cy * kd.gx);
      fr.num[v0 + lane] = (int32_t)len;
    }
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      if (C == 4) {
        *reinterpret_cast<float4*>(st + j * 4) = ra[j];
      } else {
        const float4 a = ra[j], b = rb[j];
        const uint32_t o = (idx_cur[j] * 5u) & 3u;  // kEmpty * 5 & 3 = 3: zeros either way
        const bool o1 = o & 1u, o2 = o & 2u;
        const float t0 = o1 ? a.y : a.x, t1 = o1 ? a.z : a.y, t2 = o1 ? a.w : a.z, t3 = o1 ? b.x : a.w;
        const float t4 = o1 ? b.y : b.x, t5 = o1 ? b.z : b.y, t6 = o1 ? b.w : b.z;
        st[j * C + 0] = o2 ? t2 : t0;
        st[j * C + 1] = o2 ? t3 : t1;
        st[j * C + 2] = o2 ? t4 : t2;
        st[j * C + 3] = o2 ? t5 : t3;
        st[j * C + 4 % C] = o2 ? t6 : t4;
      }
    }
    __syncwarp();
    const uint32_t w0 = (uint32_t)v0 * (PT * C);  // % 4 == 0 because v0 % 32 == 0
    float* __restrict__ dst = fr.voxels + w0;
    const int nwords = nvox * (PT * C);
    const int n4 = nwords >> 2;
    for (int i = lane; i < n4; i += 32)
      __stcs(reinterpret_cast<float4*>(dst) + i, reinterpret_cast<const float4*>(stage)[i]);
    for (int i = (n4 << 2) + lane; i < nwords; i += 32) dst[i] = stage[i];
    __syncwarp();
    cl_cur = cl_nxt;
    cl_nxt = cl_nn;
#pragma unroll
    for (int j = 0; j < PT; ++j) idx_cur[j] = idx_nxt[j];
  }
}

// ---- record-at-first-point variant: voxel ids -> first points, then a pipelined expansion -----------
// firsts[v] = index of the first point of voxel v = position of the v-th set bit of the frame's
// bitmask.  CTA (slice, frame) re-counts the bits before its slice (<= 22 KB of L2-resident words)
// instead of waiting for a prefix, then writes the positions of its own set bits.
constexpr int kFirstsThreads = 256;

// CTA (slice, frame) owns 256 consecutive bitmask words (one per thread).  It re-counts the bits
// before its slice (<= 22 KB of L2-resident words) instead of waiting for a prefix, expands its
// own set bits into shared memory and writes them out as one contiguous, coalesced run (scattered
// 4-byte stores would be partial-sector L2 writes: measured 7x slower).
__global__ void __launch_bounds__(kFirstsThreads)
hvb_scan_firsts_kernel(const HvbWork w, const int words, const int max_voxels,
                       int32_t* __restrict__ voxel_num) {
  __shared__ uint32_t warp_sums[33];
  __shared__ uint32_t stage[kFirstsThreads * 32];
  const int f = blockIdx.y, tid = threadIdx.x;
  pdl_wait();
  pdl_trigger();
  if (w.ctl(f)[w.nb + kCtlOverflow]) return;
  // 64-bit mask words: low half = "first point of a voxel", high half = "that voxel has more points"
  const uint2* __restrict__ bm = reinterpret_cast<const uint2*>(w.bitmask(f));
  uint32_t* __restrict__ firsts = w.firsts(f);
  const int lo = blockIdx.x * kFirstsThreads;
  const int wd = lo + tid;
  const uint2 mine = wd < words ? bm[wd] : make_uint2(0u, 0u);
  uint32_t bits = mine.x;
  uint32_t sum = 0;
  // every load of the re-count is issued before the first popcount (one L2 round trip, not lo / 256)
  for (int i0 = tid; i0 < lo; i0 += 8 * kFirstsThreads) {
    uint32_t t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = i0 + k * kFirstsThreads < lo ? __ldg(&bm[i0 + k * kFirstsThreads].x) : 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += __popc(t[k]);
  }
  uint32_t before;
  block_exscan(sum, warp_sums, &before);
  __syncthreads();  // warp_sums is reused below
  uint32_t total;
  uint32_t pos = block_exscan((uint32_t)__popc(bits), warp_sums, &total);
  while (bits) {
    const int bit = __ffs(bits) - 1;
    bits &= bits - 1u;
    // firsts[v] = first point of voxel v | "has a record" << 31
    stage[pos++] = ((uint32_t)wd * 32u + (uint32_t)bit) | (((mine.y >> bit) & 1u) << 31);
  }
  __syncthreads();
  for (uint32_t i = tid; i < total; i += kFirstsThreads)
    if (before + i < (uint32_t)max_voxels) firsts[before + i] = stage[i];  // voxelization_cpu.cpp:78
  if (blockIdx.x == gridDim.x - 1 && tid == 0)
    voxel_num[f] = (int32_t)min(before + total, (uint32_t)max_voxels);
}

// Expansion.  Three dependent round trips per tile of 32 voxels (firsts -> record -> rows), one of
// each kind in flight: rows of tile t, records of tile t + 1, first-point indices of tile t + 2.
// Records are fetched with lane = voxel; the rows are fetched with lane = OUTPUT WORD: word
// w = lane + 32 k of the tile belongs to (voxel, slot) q = w / C and feature w % C, so the C
// consecutive lanes of a row read one 4 C-byte piece of memory (one L1 wavefront per row instead of
// two 16-byte gathers per lane) and the loaded word goes straight from its register to its final
// place in a 128-byte coalesced store -- no shared-memory transposition of the data.  The kernel is
// bound by L1 wavefronts (random gathers), not by DRAM: the same frames resident in L2 run no faster.
#ifndef PCFE_EXP_REC_MINB
#define PCFE_EXP_REC_MINB 8
#endif
#ifndef PCFE_EXP_SPLIT
#define PCFE_EXP_SPLIT 1  // > 1: a tile's output words are fetched and stored in that many groups
#endif
#ifndef PCFE_EXP_MEAN_MINB  // the mean epilogue keeps C words per lane instead of P * C
#define PCFE_EXP_MEAN_MINB 8
#endif
// MEAN: instead of the (P, C) rows of a voxel, the mean of its points is written (fr.voxels is a
// (max_voxels, C) buffer): sum over the P slots in slot order (absent slots are +0, as in the
// zero-padded tensor HardSimpleVFE sums, voxel_encoder.py:27-44), IEEE divide by the count.
// PACK: the frames' outputs are concatenated (the detectors' torch.cat of the per-frame results,
// openpcdet.py:69-76 / voxelnet.py:60-67): every frame's buffers are the SAME arrays, frame f writes
// its rows at offset sum(voxel_num of the batch's earlier frames) and its coordinates as
// (batch index, z, y, x) rows of 16 bytes.
#ifndef PCFE_EXP_POLICY
#define PCFE_EXP_POLICY 2  // (measured: 2 = -0.45 % of the C4 step, 1 and 4 slower) bit 0: row gathers L2 evict_last, bit 1: row prefetch evict_last, bit 2: firsts / records evict_first
#endif
__device__ __forceinline__ float ldg_hint(const float* p, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint32_t ldg_hint(const uint32_t* p, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint4 ldg_hint(const uint4* p, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
  return v;
}
template <int C, bool MEAN, bool PACK>
__global__ void __launch_bounds__(kExpThreads, MEAN ? PCFE_EXP_MEAN_MINB : PCFE_EXP_REC_MINB)
hvb_expand_rec_kernel(const __grid_constant__ HvBatch batch, const HvbWork w, const GridParams g,
                      const int use_fast_div,
                      const int32_t* __restrict__ voxel_num, const int frames, const int pf_dist,
                      const int coors_vec /* every coors buffer is 16-byte aligned */,
                      const int tiles_x /* CTA-tiles per frame; the 1-D grid strides over tiles_x * frames */,
                      const int32_t* __restrict__ vn_all /* PACK: voxel_num of the batch's frame 0 */,
                      const int f_first /* PACK: batch index of this launch's frame 0 */,
                      const int pipe_tiles /* 32-voxel tiles per warp */,
                      const int map /* which tiles a warp takes, see below */) {
  constexpr int PT = 5;
  constexpr int W = PT * C;  // output words per voxel
  // source word (point index * C + feature) of every output word of a warp's tile, kEmpty = zero
  __shared__ uint32_t eff_all[MEAN ? 1 : kExpWarps * 32 * W];
  __shared__ __align__(16) int32_t coor_all[kExpWarps * 96];  // (z, y, x) of a tile: one coalesced store
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t* eff = eff_all + (MEAN ? 0 : wid * (32 * W));
  int32_t* cstage = coor_all + wid * 96;
  uint64_t pol_last = 0, pol_first = 0;
  if (PCFE_EXP_POLICY) {
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
  }
#pragma unroll 1
  for (int wi = blockIdx.x; wi < tiles_x * frames; wi += gridDim.x) {
  const int f = wi / tiles_x, bx = wi - f * tiles_x;
  if (pf_dist > 0 && f + pf_dist < frames && threadIdx.x < 32) {  // see hvb_expand_pipe_kernel
    const HvFrame& nf = batch.f[f + pf_dist];
    const size_t total = ((size_t)nf.n * C * 4) & ~(size_t)15;
    const size_t slice = ((total + tiles_x - 1) / tiles_x + 511) & ~(size_t)511;
    const size_t lo = (size_t)bx * slice + (size_t)threadIdx.x * (slice / 32);
    if (lo < total) {
      const uint32_t bytes = (uint32_t)min(slice / 32, total - lo);
      if (PCFE_EXP_POLICY & 2)
        asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(reinterpret_cast<const char*>(nf.pts) + lo), "r"(bytes), "l"(pol_last) : "memory");
      else
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(nf.pts) + lo), "r"(bytes) : "memory");
    }
  }
  pdl_wait();  // (the prefetch above only touches the caller's input rows)
  pdl_trigger();
  if (w.ctl(f)[w.nb + kCtlOverflow]) continue;
  const HvFrame& fr = batch.f[f];
  const int m = voxel_num[f];
  const uint32_t* __restrict__ firsts = w.firsts(f);
  const uint4* __restrict__ rec = w.rec(f);
  const float* __restrict__ pts = fr.pts;
  // tile of the warp's step `it`: base + it * tstride.  map 0: a warp's tiles are consecutive; 1: the CTA's
  // tiles are dealt to its warps round-robin (the CTA writes one contiguous run per step); 2: the frame's
  // tiles are dealt to all its warps round-robin (the frame's CTAs sweep it together: what they gather at any moment
  // comes from one narrow band of the frame)
  int tbase = (bx * kExpWarps + wid) * pipe_tiles, tstride = 1;
  if (map == 1) {
    tbase = bx * kExpWarps * pipe_tiles + wid;
    tstride = kExpWarps;
  } else if (map == 2) {
    const int nct = (m + 32 * kExpWarps * pipe_tiles - 1) / (32 * kExpWarps * pipe_tiles);  // CTAs the frame needs
    if (bx >= nct) continue;
    tbase = bx * kExpWarps + wid;
    tstride = kExpWarps * nct;
  }
  const int vbase = tbase * 32, vstride = tstride * 32;
  if (vbase >= m) continue;  // warp-uniform
  size_t off = 0;  // PACK: rows of the earlier frames
  if (PACK) {
    int part = 0;
    for (int k = lane; k < f_first + f; k += 32) part += __ldg(vn_all + k);
    off = (size_t)__reduce_add_sync(0xFFFFFFFFu, part);
  }

  // firsts[v] = first point of voxel v | (the voxel has more points: see rec[first]) << 31
  auto load_first = [&](int v0) { return v0 + lane < m ? ((PCFE_EXP_POLICY & 4) ? ldg_hint(firsts + v0 + lane, pol_first) : __ldg(firsts + v0 + lane)) : kEmpty; };
  auto load_rec = [&](uint32_t fi) {
    uint4 r = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);  // no points besides the first
    if (fi != kEmpty && (fi >> 31))
      r = (PCFE_EXP_POLICY & 4) ? ldg_hint(rec + PCFE_REC_STRIDE * (size_t)(fi & 0x7FFFFFFFu), pol_first) : __ldg(rec + PCFE_REC_STRIDE * (size_t)(fi & 0x7FFFFFFFu));
    return r;
  };
  const FastAxes fa = make_fast_axes(g);
  uint32_t fi_cur = load_first(vbase);
  uint32_t fi_nxt = load_first(vbase + vstride);
  uint4 ra_cur = load_rec(fi_cur);

#pragma unroll 1
  for (int it = 0; it < pipe_tiles; ++it) {
    const int v0 = vbase + it * vstride;
    if (v0 >= m) break;  // warp-uniform
    const int nvox = min(32, m - v0);
    const bool have = fi_cur != kEmpty;  // false for lanes past the end
    const uint32_t first = fi_cur & 0x7FFFFFFFu;
    const uint32_t len = (have ? 1u : 0u) + (ra_cur.x != kEmpty) + (ra_cur.y != kEmpty) + (ra_cur.z != kEmpty) + (ra_cur.w != kEmpty);
    if (!MEAN) {  // lane = voxel: the 25 (20) source words of its output, slot by slot
      const uint32_t idx5[PT] = {have ? first : kEmpty, ra_cur.x, ra_cur.y, ra_cur.z, ra_cur.w};
#pragma unroll
      for (int j = 0; j < PT; ++j) {
        const uint32_t b0 = idx5[j] * (uint32_t)C;
#pragma unroll
        for (int q = 0; q < C; ++q) eff[lane * W + j * C + q] = idx5[j] != kEmpty ? b0 + (uint32_t)q : kEmpty;
      }
    }
    // the voxel's coordinates are the cell of its first point (recomputed: no key is stored)
    float px = 0.f, py = 0.f, pz = 0.f;
    if (have) {
      const float* __restrict__ fp = pts + (size_t)first * C;
      if (PCFE_EXP_POLICY & 1) { px = ldg_hint(fp, pol_last); py = ldg_hint(fp + 1, pol_last); pz = ldg_hint(fp + 2, pol_last); }
      else { px = __ldg(fp); py = __ldg(fp + 1); pz = __ldg(fp + 2); }
      fr.num[off + v0 + lane] = (int32_t)len;
    }
    __syncwarp();
    // rows of this tile, lane = output word
    constexpr int HW = (W + PCFE_EXP_SPLIT - 1) / PCFE_EXP_SPLIT;  // output words per lane fetched together
    float val[MEAN ? C : HW];
    if (MEAN) {
      // lane = output word of the (32, C) mean tile: word o = lane + 32 r is feature o % C of voxel
      // o / C, whose point indices sit in that lane's registers (shuffles, no shared memory); the
      // C lanes of a voxel read the C consecutive words of each of its rows
      const uint32_t f0 = have ? first : kEmpty;
#pragma unroll
      for (int r = 0; r < C; ++r) {
        const int o = lane + 32 * r;
        const int vl = o / C, q = o - vl * C;
        const uint32_t i0 = __shfl_sync(0xFFFFFFFFu, f0, vl), i1 = __shfl_sync(0xFFFFFFFFu, ra_cur.x, vl),
                       i2 = __shfl_sync(0xFFFFFFFFu, ra_cur.y, vl), i3 = __shfl_sync(0xFFFFFFFFu, ra_cur.z, vl),
                       i4 = __shfl_sync(0xFFFFFFFFu, ra_cur.w, vl);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
        if (i0 != kEmpty) s0 = __ldg(pts + (size_t)i0 * C + q);
        if (i1 != kEmpty) s1 = __ldg(pts + (size_t)i1 * C + q);
        if (i2 != kEmpty) s2 = __ldg(pts + (size_t)i2 * C + q);
        if (i3 != kEmpty) s3 = __ldg(pts + (size_t)i3 * C + q);
        if (i4 != kEmpty) s4 = __ldg(pts + (size_t)i4 * C + q);
        const int cnt = (i0 != kEmpty) + (i1 != kEmpty) + (i2 != kEmpty) + (i3 != kEmpty) + (i4 != kEmpty);
        // slot-order sum over all five slots (absent = +0), IEEE divide by the count
        const float a = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s0, s1), s2), s3), s4);
        val[r] = cnt ? __fdiv_rn(a, (float)cnt) : 0.0f;
      }
    } else {
#pragma unroll
    for (int k = 0; k < HW; ++k) {
      const uint32_t src = eff[lane + 32 * k];  // word lane + 32 k of the tile
      val[k] = 0.0f;
      if (src != kEmpty) val[k] = (PCFE_EXP_POLICY & 1) ? ldg_hint(pts + src, pol_last) : __ldg(pts + src);
    }
    }
    // records of the next tile, first-point indices of the one after
    const uint4 ra_nxt = load_rec(fi_nxt);
    const uint32_t fi_nn = load_first(v0 + 2 * vstride);
    if (have) {  // voxelization_cpu.cpp:23-29 (the point is in range: it produced a cell key)
      const float ax = __fsub_rn(px, g.x0), ay = __fsub_rn(py, g.y0), az = __fsub_rn(pz, g.z0);
      float qx, qy, qz;
      if (use_fast_div && fast_div_guard(ax) && fast_div_guard(ay) && fast_div_guard(az)) {
        qx = fast_div(ax, g.vx, fa.rx); qy = fast_div(ay, g.vy, fa.ry); qz = fast_div(az, g.vz, fa.rz);
      } else {
        qx = __fdiv_rn(ax, g.vx); qy = __fdiv_rn(ay, g.vy); qz = __fdiv_rn(az, g.vz);
      }
      if (PACK) {  // one 16-byte row per lane: a 512-byte coalesced store
        reinterpret_cast<int4*>(fr.coors)[off + v0 + lane] =
            make_int4(f_first + f, __float2int_rz(qz), __float2int_rz(qy), __float2int_rz(qx));
      } else {
        cstage[lane * 3 + 0] = __float2int_rz(qz);
        cstage[lane * 3 + 1] = __float2int_rz(qy);
        cstage[lane * 3 + 2] = __float2int_rz(qx);
      }
    }
    __syncwarp();
    if (MEAN) {
      float* __restrict__ mdst = fr.voxels + (off + v0) * C;
#pragma unroll
      for (int r = 0; r < C; ++r)
        if (lane + 32 * r < nvox * C) mdst[lane + 32 * r] = val[r];
    } else {
    float* __restrict__ dst = fr.voxels + (off + v0) * W;
#pragma unroll
    for (int h0 = 0; h0 < W; h0 += HW) {
      if (h0 > 0) {  // PCFE_EXP_SPLIT > 1: the next group of words (fewer live registers per warp)
#pragma unroll
        for (int k = 0; k < HW; ++k) {
          val[k] = 0.0f;
          if (h0 + k < W) {
            const uint32_t src = eff[lane + 32 * (h0 + k)];
            if (src != kEmpty) val[k] = __ldg(pts + src);
          }
        }
      }
      if (nvox == 32) {
#pragma unroll
        for (int k = 0; k < HW; ++k)
          if (h0 + k < W) __stcs(dst + lane + 32 * (h0 + k), val[k]);
      } else {
#pragma unroll
        for (int k = 0; k < HW; ++k)
          if (h0 + k < W && lane + 32 * (h0 + k) < nvox * W) __stcs(dst + lane + 32 * (h0 + k), val[k]);
      }
    }
    }
    if (!PACK) {  // coordinates: 3 * nvox words, contiguous; v0 % 32 == 0 keeps the run 16-byte aligned
      int32_t* __restrict__ cdst = fr.coors + (size_t)v0 * 3;
      const int cw = nvox * 3;
      if (coors_vec) {
        if (lane < (cw >> 2)) reinterpret_cast<int4*>(cdst)[lane] = reinterpret_cast<const int4*>(cstage)[lane];
        if (lane < (cw & 3)) cdst[(cw & ~3) + lane] = cstage[(cw & ~3) + lane];
      } else {
        for (int i = lane; i < cw; i += 32) cdst[i] = cstage[i];
      }
    }
    __syncwarp();
    fi_cur = fi_nxt;
    fi_nxt = fi_nn;
    ra_cur = ra_nxt;
  }
  __syncwarp();
  }
}

// ---- expansion for long voxels (pillars: P * C = 320 words, a multiple of 4) -----------------------
// A warp owns kWordsVox consecutive voxels, which are contiguous in memory: ONE run of nvox * W / 4 float4s.
// Lane = voxel for the cell records (one coalesced load), the key decoding (host-computed reciprocals)
// and the coordinate / count stores; then
//   phase 1  every float4 behind its voxel's data is a zero store (depends on the cell records only);
//   phase 2  the data float4s of all voxels, numbered through a prefix over the voxels' lengths: list
//            entries and row words are fetched by the lane that stores them -- two dependent loads, all
//            lanes busy whatever the lengths are.
// The zero stores are in flight while phase 2 waits for its lists and rows.  No shared memory.  C == 0:
// features per point at run time.  Needs 16-byte aligned voxel buffers.
// History (profiles/r02_summary.md): a first version decoded one voxel at a time (211 warp-instructions
// per voxel, 0.191 ms per 16-frame C5 step); the second walked the voxels one after the other with lane =
// output word and the lists two voxels ahead (123 per voxel, 0.183 ms, issue-active 64 %, DRAM 57 %); this
// one needs ~50 per voxel and is no faster at 8 voxels per warp (0.187 ms): the kernel is bound by how its
// writes reach DRAM -- a variant with the data phase cut out still takes 0.154 ms for 700 MB of zeros where a
// grid-stride fill of the same bytes runs at 7.1 TB/s.  What helps is a SMALLER contiguous region per CTA
// (the chip then writes a narrower window at any moment): 4 voxels per warp 0.176 ms, 2: 0.179, 1: 0.220, 16: 0.200.
#ifndef PCFE_WORDS_VOX
#define PCFE_WORDS_VOX 4
#endif
constexpr int kWordsVox = PCFE_WORDS_VOX;
template <int C>
__global__ void __launch_bounds__(kExpThreads)
hvb_expand_words_kernel(const __grid_constant__ HvBatch batch, const HvbWork w, const KeyDecode kd,
                             const int c_rt, const int max_points, const int32_t* __restrict__ voxel_num,
                             const uint32_t m_w4 /* ceil(2^16 / (W / 4)) when i / (W / 4) = (i * m_w4) >> 16 holds for all i < kWordsVox * W / 4, else 0 */) {
  const int f = blockIdx.y;
  pdl_wait();
  if (w.ctl(f)[w.nb + kCtlOverflow]) return;
  const HvFrame& fr = batch.f[f];
  const int c = C > 0 ? C : c_rt;
  const int m = voxel_num[f];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int W = max_points * c, W4 = W >> 2;  // W % 4 == 0
  const uint4* __restrict__ vcell = reinterpret_cast<const uint4*>(w.vcell(f));
  const uint32_t* __restrict__ lst = w.lst(f);
  const float* __restrict__ pts = fr.pts;
  const int v0 = (blockIdx.x * kExpWarps + wid) * kWordsVox;
  if (v0 >= m) return;
  const int nvox = min(kWordsVox, m - v0);
  uint4 cl = make_uint4(0u, 0u, 0u, 0u);  // key, len, list_off, first
  if (lane < nvox) cl = __ldg(vcell + v0 + lane);
  const uint32_t len = min(cl.y, (uint32_t)max_points);
  const uint32_t nreal = len * (uint32_t)c;          // 0 for lanes >= nvox
  const uint32_t nreal4 = (nreal + 3u) >> 2;          // data float4s of the lane's voxel (the last one zero-padded)
  float4* __restrict__ dst4 = reinterpret_cast<float4*>(fr.voxels + (size_t)v0 * W);
  // phase 1: zero padding
  const int total4 = nvox * W4;
#pragma unroll 2
  for (int i4 = lane; i4 < ((total4 + 31) & ~31); i4 += 32) {
    const uint32_t v = m_w4 ? ((uint32_t)i4 * m_w4) >> 16 : (uint32_t)i4 / (uint32_t)W4;
    const uint32_t r4 = (uint32_t)i4 - v * (uint32_t)W4;
    const uint32_t nr4 = __shfl_sync(0xFFFFFFFFu, nreal4, (int)(v & 31u));
    if (i4 < total4 && r4 >= nr4) __stcs(dst4 + i4, make_float4(0.f, 0.f, 0.f, 0.f));
  }
  // coordinates and counts
  if (lane < nvox) {
    const uint32_t cz = div_small_err(cl.x, kd.plane, kd.m_plane);
    const uint32_t rem = cl.x - cz * kd.plane;
    const uint32_t cy = div_small_err(rem, kd.gx, kd.m_gx);
    int32_t* co = fr.coors + (size_t)(v0 + lane) * 3;
    co[0] = (int32_t)cz;
    co[1] = (int32_t)cy;
    co[2] = (int32_t)(rem - cy * kd.gx);
    fr.num[v0 + lane] = (int32_t)len;
  }
  // phase 2: data float4s, numbered d = 0 .. D - 1 through the voxels
  uint32_t incl = nreal4;
#pragma unroll
  for (int d = 1; d < kWordsVox; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += t;
  }
  const uint32_t D = __shfl_sync(0xFFFFFFFFu, incl, kWordsVox - 1);
  uint32_t bound[kWordsVox];  // inclusive prefix of every voxel, in every lane
#pragma unroll
  for (int j = 0; j < kWordsVox; ++j) bound[j] = __shfl_sync(0xFFFFFFFFu, incl, j);
  for (uint32_t d0 = 0; d0 < D; d0 += 32) {  // warp-uniform trip count
    const uint32_t d = d0 + (uint32_t)lane;
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < kWordsVox - 1; ++j) v += d >= bound[j] ? 1u : 0u;
    const uint32_t inc_v = __shfl_sync(0xFFFFFFFFu, incl, (int)v), n4_v = __shfl_sync(0xFFFFFFFFu, nreal4, (int)v);
    const uint32_t nr_v = __shfl_sync(0xFFFFFFFFu, nreal, (int)v), off_v = __shfl_sync(0xFFFFFFFFu, cl.z, (int)v);
    if (d < D) {
      const uint32_t r4 = d - (inc_v - n4_v);
      const uint32_t w0 = 4u * r4;
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t wd = w0 + (uint32_t)k;
        o[k] = 0.0f;
        if (wd < nr_v) {
          const uint32_t sl = wd / (uint32_t)c;
          const uint32_t idx = __ldg(lst + off_v + sl);
          o[k] = __ldg(pts + (size_t)idx * c + (wd - sl * (uint32_t)c));
        }
      }
      __stcs(dst4 + (size_t)v * W4 + r4, make_float4(o[0], o[1], o[2], o[3]));
    }
  }
}

template <int C>
int launch_expand(dim3 grid, cudaStream_t st, const HvBatch& b, const HvbWork& w,
                  const GridParams& g, int c, int p, int vt, const int32_t* vn, int vec_ok) {
  const size_t smem = (size_t)kExpWarps * kExpStageWords * sizeof(float);
  hvb_expand_kernel<C><<<grid, kExpThreads, smem, st>>>(b, w, g, c, p, vt, vn, vec_ok);
  PCFE_LAUNCH_CHECK();
  return PCFE_OK;
}

#include "hv_cluster.cuh"

// One cluster of kCS CTAs per frame, as many clusters as the device keeps resident (each loops over
// frames cluster, cluster + n, ...).  Returns PCFE_ERR_SHAPE when the device cannot run the cluster
// shape at all: the caller then takes the multi-launch sequence.
struct HvcDevice {
  int tried = 0, max_clusters = 0;
};
HvcDevice g_hvc_dev[64][3];

template <int CT>
int hvc_launch(int device, cudaStream_t st, bool pdl, const HvBatch& b, const HvbWork& w, const GridParams& g,
               int c, int fdiv, int max_voxels, int32_t* voxel_num, int frames) {
  auto kern = hvc_group_kernel<CT>;
  HvcDevice& d = g_hvc_dev[device][CT == 4 ? 0 : CT == 5 ? 1 : 2];
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kCT);
  cfg.dynamicSmemBytes = sizeof(HvcSmem);
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  if (!d.tried) {
    d.tried = 1;
    d.max_clusters = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HvcSmem)) == cudaSuccess &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      cfg.gridDim = dim3(kCS * 8);
      cfg.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) == cudaSuccess) d.max_clusters = nc;
    }
    (void)cudaGetLastError();
  }
  if (d.max_clusters < 1) return PCFE_ERR_SHAPE;
  const int ncl = std::min(d.max_clusters, frames);
  cfg.gridDim = dim3((unsigned)(kCS * ncl));
  cfg.numAttrs = pdl ? 2 : 1;
  PCFE_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, b, w, g, c, fdiv, max_voxels, voxel_num, frames));
  PCFE_LAUNCH_CHECK();
  return PCFE_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int hvb_make_plan(int64_t n_max, int c, const float vs[3], const float rg[6], int max_points,
                  int max_voxels, HvBucketPlan* p) {
  int rc = hvg_make_plan(n_max, vs, rg, max_points, max_voxels, &p->slow);
  if (rc != PCFE_OK) return rc;
  if (max_points > 0xFFFF) return PCFE_ERR_CAPS;  // list length is packed into 16 bits
  if ((int64_t)max_voxels * (int64_t)std::max(max_points, 1) >= (1ll << 31)) return PCFE_ERR_TOO_LARGE;
  p->g = p->slow.g;
  p->npad = p->slow.npad;
  p->words = p->slow.words;
  const int target = std::max(64, std::min((int)g_opt_bucket_avg, 1400));
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
  // arenas: one list entry per point at most, one cell per point at most
  p->ent_b = align256((size_t)p->nb * (size_t)p->cap * sizeof(uint2));
  p->lst_b = align256((size_t)p->npad * sizeof(uint32_t));
  p->cells_b = align256((size_t)p->npad * sizeof(Cell));
  const size_t vmax = (size_t)std::min<int64_t>(max_voxels, std::max<int64_t>(n_max, 1));
  p->vcell_b = align256(std::max<size_t>(vmax, 1) * sizeof(Cell));
  p->word_b = align256((size_t)p->words * sizeof(uint32_t));
  p->cnt_b = align256((size_t)(p->nb + 16) * sizeof(uint32_t)) + 256;
  p->rec_b = align256((size_t)p->npad * 16 * PCFE_REC_STRIDE);
  p->firsts_b = align256(std::max<size_t>(vmax, 1) * sizeof(uint32_t));
  const size_t fast = std::max(p->ent_b + p->lst_b + p->cells_b + p->vcell_b, p->ent_b + p->rec_b + p->firsts_b);
  // the fallback reuses the frame's own region as table | lists | pslot
  const size_t slow = p->slow.table_b + p->slow.list_b + p->slow.pslot_b;
  p->region_b = std::max(fast, slow);
  p->per_frame = p->region_b + 4 * p->word_b + p->cnt_b;  // bitmask (2 words per 32 points) + {bits, prefix} pairs
  // expansion tile: voxels per warp such that a tile fits the per-warp stage
  if ((int64_t)std::max(max_points, 1) * c > kExpStageWords) return PCFE_ERR_TOO_LARGE;
  p->exp_vt = 32;
  while (p->exp_vt > 1 && (int64_t)p->exp_vt * std::max(max_points, 1) * c > kExpStageWords) p->exp_vt >>= 1;
  p->smem_bucket = (size_t)(2 * p->slots + 2 * p->cap) * 4 + (size_t)(2 * p->cap) * 2;
  return PCFE_OK;
}

// Two auxiliary streams per device so that consecutive waves overlap: the bucket kernel is bound
// by shared-memory atomics, bin/expand by DRAM, so wave k+1's grouping hides under wave k's
// expansion.  Forked from / joined to the caller's stream with events; the caller sees plain
// stream-ordered semantics.
struct AuxStreams {
  cudaStream_t s[2] = {nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
};
static AuxStreams g_aux[64];
static std::mutex g_aux_mu;

static int get_aux(int device, AuxStreams** out) {
  if (device < 0 || device >= 64) return PCFE_ERR_DEVICE;
  AuxStreams& a = g_aux[device];
  if (!a.fork) {
    for (int k = 0; k < 2; ++k) {
      PCFE_CUDA_TRY(cudaStreamCreateWithFlags(&a.s[k], cudaStreamNonBlocking));
      PCFE_CUDA_TRY(cudaEventCreateWithFlags(&a.join[k], cudaEventDisableTiming));
    }
    PCFE_CUDA_TRY(cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming));
  }
  *out = &a;
  return PCFE_OK;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) and size, not per call
static int ensure_dyn_smem(const void* kernel, int slot, int device, size_t bytes) {
  static size_t g_set[8][64] = {};
  if (device < 0 || device >= 64 || slot < 0 || slot >= 8) return PCFE_ERR_DEVICE;
  if (g_set[slot][device] >= bytes) return PCFE_OK;
  PCFE_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  g_set[slot][device] = bytes;
  return PCFE_OK;
}

static int hvb_run_waves(const pcfe_frame_t* frames, int num_frames, int c, const HvBucketPlan& p,
                         int max_points, int max_voxels, int32_t* voxel_num, void* workspace, int wave,
                         int device, cudaStream_t user_st, int mode, bool overlap, AuxStreams* aux);

int hvb_run(const pcfe_frame_t* frames, int num_frames, int c, const HvBucketPlan& p,
            int max_points, int max_voxels, int32_t* voxel_num, void* workspace, int wave,
            int nbuf, int device, cudaStream_t user_st, int mode) {
  std::lock_guard<std::mutex> lk(g_aux_mu);
  const int pack = (mode & kHvPack) ? 1 : 0;
  const int nwaves = (num_frames + wave - 1) / wave;
  // (packed output: a wave places its rows behind those of the waves before it -- one stream)
  const bool overlap = nbuf >= 2 && nwaves >= 2 && !pack && g_opt_overlap != 0;
  AuxStreams* aux = nullptr;
  if (overlap) {
    int rc = get_aux(device, &aux);
    if (rc != PCFE_OK) return rc;
    PCFE_CUDA_TRY(cudaEventRecord(aux->fork, user_st));
    for (int k = 0; k < 2; ++k) PCFE_CUDA_TRY(cudaStreamWaitEvent(aux->s[k], aux->fork, 0));
  }
  const int rc = hvb_run_waves(frames, num_frames, c, p, max_points, max_voxels, voxel_num, workspace, wave, device,
                               user_st, mode, overlap, aux);
  if (overlap) {
    // always joined, also after an error inside the wave loop: work already enqueued on the auxiliary
    // streams stays ordered before whatever the caller enqueues next on its stream
    for (int k = 0; k < 2; ++k) {
      const cudaError_t e1 = cudaEventRecord(aux->join[k], aux->s[k]);
      const cudaError_t e2 = e1 == cudaSuccess ? cudaStreamWaitEvent(user_st, aux->join[k], 0) : e1;
      if (e2 != cudaSuccess && rc == PCFE_OK) return (int)e2;
    }
  }
  return rc;
}

static int hvb_run_waves(const pcfe_frame_t* frames, int num_frames, int c, const HvBucketPlan& p,
                         int max_points, int max_voxels, int32_t* voxel_num, void* workspace, int wave,
                         int device, cudaStream_t user_st, int mode, bool overlap, AuxStreams* aux) {
  const int mean = mode & kHvMean, pack = (mode & kHvPack) ? 1 : 0;
  // scratch layout per wave buffer: [frame regions] x wave | [bitmask | ctl] x wave (zeroed per
  // wave) | [wordprefix pairs] x wave
  const size_t buf_bytes = (size_t)wave * p.per_frame;
  HvbWork w;
  w.region_stride = p.region_b;
  w.lst_off = p.ent_b;
  w.cells_off = p.ent_b + p.lst_b;
  w.vcell_off = p.ent_b + p.lst_b + p.cells_b;
  w.rec_off = p.ent_b;
  w.firsts_off = p.ent_b + p.rec_b;
  const size_t zero_per = 2 * p.word_b + p.cnt_b;
  w.zero_stride = zero_per / sizeof(uint32_t);
  w.ctl_off = 2 * p.word_b / sizeof(uint32_t);
  w.word_stride = 2 * p.word_b / sizeof(uint32_t);
  w.nb = p.nb; w.log2_nb = p.log2_nb; w.cap = p.cap; w.slots = p.slots; w.log2_slots = p.log2_slots;
  w.arena_cap = (uint32_t)p.npad;

  {
    int rc = ensure_dyn_smem((const void*)hvb_bucket_small_kernel<5>, 0, device, p.smem_bucket);
    if (rc == PCFE_OK) rc = ensure_dyn_smem((const void*)hvb_bucket_small_kernel<8>, 1, device, p.smem_bucket);
    if (rc == PCFE_OK)
      rc = ensure_dyn_smem((const void*)hvb_bucket_rec_kernel<false>, 2, device,
                           (size_t)(2 * p.slots + 2 * p.cap) * 4 + (size_t)(2 * p.cap) * 2);
    if (rc == PCFE_OK)
      rc = ensure_dyn_smem((const void*)hvb_bucket_rec_kernel<true>, 4, device,
                           (size_t)(2 * p.slots + 2 * p.cap) * 4 + (size_t)(2 * p.cap) * 2);
    if (rc != PCFE_OK) return rc;
  }
  const int pe = std::max(max_points, 1);
  int vec_ok = 1;  // float4 tile stream / vector row loads need 16-byte aligned buffers
  for (int k = 0; k < num_frames && vec_ok; ++k)
    vec_ok = !((!mean && ((uintptr_t)frames[k].voxels & 15)) || ((uintptr_t)frames[k].points & 15));
  int coors_vec = 1;
  for (int k = 0; k < num_frames && coors_vec; ++k) coors_vec = !((uintptr_t)frames[k].coors & 15);

  int wave_idx = 0;
  for (int f0 = 0; f0 < num_frames; f0 += wave, ++wave_idx) {
    const int wv = std::min(wave, num_frames - f0);
    cudaStream_t st = overlap ? aux->s[wave_idx & 1] : user_st;
    char* base = (char*)workspace + (overlap ? (size_t)(wave_idx & 1) * buf_bytes : 0);
    char* zero_base = base + (size_t)wave * p.region_b;
    w.region = base;
    w.zero = (uint32_t*)zero_base;
    w.wordprefix = (uint32_t*)(zero_base + (size_t)wave * zero_per);
    HvBatch b;
    int64_t wn_max = 0;
    for (int k = 0; k < wv; ++k) {
      const pcfe_frame_t& fr = frames[f0 + k];
      b.f[k] = HvFrame{fr.points, fr.voxels, fr.coors, fr.num_points, (int)fr.n, 0};
      wn_max = std::max(wn_max, fr.n);
    }
    // the mean epilogue and the packed output exist on the record path only (P == 5, C = 4 / 5,
    // aligned buffers)
    if ((mean || pack) && !(max_points == 5 && (c == 4 || c == 5) && vec_ok && max_voxels < (1 << 24) && wn_max < 0xFFFFFF &&
                  g_opt_bucket_variant != 1))
      return PCFE_ERR_SHAPE;
    // P == 5 with 16-byte aligned rows: the record-at-first-point variant (no order pass)
    const bool use_rec = max_points == 5 && (c == 4 || c == 5) && vec_ok && max_voxels < (1 << 24) && wn_max < 0xFFFFFF &&
                         g_opt_bucket_variant != 1;
    // ... whose partition + grouping + voxel numbering is ONE cluster kernel when the frames fit
    bool clustered = false;
    if (use_rec && g_opt_cluster && wn_max <= kCMaxPoints && device >= 0 && device < 64) {
      ProfScope ps("hvc_group", st);
      const int fdiv = fast_div_sizes_ok(p.g) && !g_opt_no_fast_div ? 1 : 0;
      int rcc = c == 4 ? hvc_launch<4>(device, st, g_opt_pdl != 0, b, w, p.g, c, fdiv, max_voxels, voxel_num + f0, wv)
                       : hvc_launch<5>(device, st, g_opt_pdl != 0, b, w, p.g, c, fdiv, max_voxels, voxel_num + f0, wv);
      if (rcc == PCFE_OK) clustered = true;
      else if (rcc != PCFE_ERR_SHAPE) return rcc;
    }
    if (!clustered) {
      ProfScope ps("memset_ctl", st);
      if (g_opt_pdl) {  // a kernel instead of a memset node, so that the bin kernel can be its
                        // programmatic dependent (zero_per is a multiple of 256 bytes)
        const size_t n16 = (size_t)wv * zero_per / 16;
        hvb_zero_kernel<<<(unsigned)std::min<size_t>((n16 + 255) / 256, 1184), 256, 0, st>>>(reinterpret_cast<uint4*>(zero_base), n16);
        PCFE_LAUNCH_CHECK();
      } else {
        PCFE_CUDA_TRY(cudaMemsetAsync(zero_base, 0, (size_t)wv * zero_per, st));
        count_launch();
      }
    }
    const int wnpad = std::max((int)((wn_max + 31) / 32 * 32), 32);
    if (!clustered) {
      ProfScope ps("hvb_bin", st);
      // (a wave of empty frames still runs the sequence: one idle tile, voxel_num = 0 from the scan)
      const int64_t big_tiles = std::max<int64_t>((wn_max + kBinTile - 1) / kBinTile, 1);
      // small batches: 1024-point tiles, four times the CTAs with a quarter of the latency chain each
      const int bin_small = g_opt_bin_small;
      const bool small = bin_small == 1 || (bin_small != 0 && big_tiles * wv < 2 * 148 * PCFE_BIN_MINB);
      const int tile = small ? kBinTileSmall : kBinTile;
      const dim3 grid((unsigned)std::max<int64_t>((wn_max + tile - 1) / tile, 1), (unsigned)wv);
      const int fdiv = fast_div_sizes_ok(p.g) && !g_opt_no_fast_div ? 1 : 0;
      const bool pdl = g_opt_pdl != 0;
#define PCFE_LAUNCH_BIN(CC)                                                                                              \
  do {                                                                                                                   \
    if (small) PCFE_CUDA_TRY(launch_pdl(hvb_bin_kernel<CC, kBinPerThreadSmall>, grid, dim3(kBinThreads), 0, st, pdl, b, w, p.g, c, fdiv)); \
    else PCFE_CUDA_TRY(launch_pdl(hvb_bin_kernel<CC, kBinPerThread>, grid, dim3(kBinThreads), 0, st, pdl, b, w, p.g, c, fdiv));            \
  } while (0)
      if (c == 4) PCFE_LAUNCH_BIN(4);
      else if (c == 5) PCFE_LAUNCH_BIN(5);
      else PCFE_LAUNCH_BIN(0);
#undef PCFE_LAUNCH_BIN
      PCFE_LAUNCH_CHECK();
    }
    int rc = PCFE_OK;
    if (use_rec) {
      if (!clustered) {
        ProfScope ps("hvb_bucket", st);
        const dim3 grid((unsigned)p.nb, (unsigned)wv);
        const size_t smem_rec = (size_t)(2 * p.slots + 2 * p.cap) * 4 + (size_t)p.cap * 2;
        // speculative first copy: the average bucket fill of the largest frame, rounded up to 64
        // entries, never more than the region
        const int spec = (int)std::min<int64_t>(p.cap, std::max<int64_t>((((wn_max + p.nb - 1) / p.nb) + 63) / 64 * 64, 64));
        if (g_opt_warp_dedup)
          PCFE_CUDA_TRY(launch_pdl(hvb_bucket_rec_kernel<true>, grid, dim3(kBucketThreads), smem_rec, st, g_opt_pdl != 0, w, pe, spec));
        else
          PCFE_CUDA_TRY(launch_pdl(hvb_bucket_rec_kernel<false>, grid, dim3(kBucketThreads), smem_rec, st, g_opt_pdl != 0, w, pe, spec));
        PCFE_LAUNCH_CHECK();
      }
      if (!clustered) {
        ProfScope ps("hvb_scan_firsts", st);
        PCFE_CUDA_TRY(launch_pdl(hvb_scan_firsts_kernel, dim3((unsigned)((wnpad / 32 + kFirstsThreads - 1) / kFirstsThreads), (unsigned)wv),
                                 dim3(kFirstsThreads), 0, st, g_opt_pdl != 0, w, wnpad / 32, max_voxels, voxel_num + f0));
        PCFE_LAUNCH_CHECK();
      }
      if (pack) {  // every frame's voxel count has to be known before the first row is placed
        rc = hvg_launch_slow(b, wv, w.zero + w.ctl_off + p.nb + kCtlOverflow, w.zero_stride, g_opt_force_overflow,
                             w.region, w.region_stride, p.slow, w.zero, w.zero_stride, w.wordprefix, w.word_stride,
                             c, max_points, max_voxels, voxel_num + f0, st, mean, 1, voxel_num, f0);
        if (rc != PCFE_OK) return rc;
      }
      {
        ProfScope ps("hvb_expand", st);
        const int fdiv = fast_div_sizes_ok(p.g) && !g_opt_no_fast_div ? 1 : 0;
        const int64_t vmax = std::max<int64_t>(std::min<int64_t>(max_voxels, wn_max), 1);
        // tiles per warp, dealt round-robin over the frame's warps (hv_expand_map = 2: the frame's CTAs sweep it
        // together, so the rows / records gathered at any moment come from one narrow band of the frame instead of
        // one 38 KB region per CTA -- the read side's locality, not the stores', profiles/r02_summary.md): 4 for batches of 16 frames and more (64 frames: 3 / 4 / 5 tiles = 0.3508 /
        // 0.3452 / 0.3481 ms, consecutive tiles: 0.3553; 32 frames 0.1868 / 0.1850 / 0.1867), 3 below (4 frames:
        // 0.0350 / 0.0388; C1, 16 frames of 16 000 voxels: 3); `hv_expand_tiles` overrides
        // (the rule: 4 when the voxel capacity of the wave fills the GPU's 1184 CTA slots twice over with 4-tile warps)
        const int pipe_tiles = g_opt_expand_tiles > 0 ? (int)g_opt_expand_tiles : (vmax * wv >= 2ll * 1184 * kExpWarps * 4 * 32 ? 4 : 3);
        const int pper = kExpWarps * pipe_tiles * 32;
        const dim3 pgrid((unsigned)((vmax + pper - 1) / pper), (unsigned)wv);
        const int32_t* vn = voxel_num + f0;
        const int tiles_x = (int)pgrid.x;
        unsigned egrid = pgrid.x * pgrid.y;
        if (g_opt_expand_ctas > 0) egrid = std::min<unsigned>(egrid, (unsigned)(g_opt_expand_ctas * 148));
        const bool pdl = g_opt_pdl != 0;
#define PCFE_LAUNCH_EXPAND_REC(CC, MM, PP)                                                                      \
  PCFE_CUDA_TRY(launch_pdl(hvb_expand_rec_kernel<CC, MM, PP>, dim3(egrid), dim3(kExpThreads), 0, st, pdl, b, w, p.g, \
                           fdiv, vn, wv, (int)g_opt_expand_prefetch, coors_vec, tiles_x, (const int32_t*)voxel_num, f0, pipe_tiles, (int)g_opt_expand_map))
#define PCFE_LAUNCH_EXPAND_REC_C(CC)                                \
  do {                                                              \
    if (mean && pack) PCFE_LAUNCH_EXPAND_REC(CC, true, true);       \
    else if (mean) PCFE_LAUNCH_EXPAND_REC(CC, true, false);         \
    else if (pack) PCFE_LAUNCH_EXPAND_REC(CC, false, true);         \
    else PCFE_LAUNCH_EXPAND_REC(CC, false, false);                  \
  } while (0)
        if (c == 4) PCFE_LAUNCH_EXPAND_REC_C(4);
        else PCFE_LAUNCH_EXPAND_REC_C(5);
#undef PCFE_LAUNCH_EXPAND_REC_C
#undef PCFE_LAUNCH_EXPAND_REC
        PCFE_LAUNCH_CHECK();
      }
    } else {
    {
      ProfScope ps("hvb_bucket", st);
      const dim3 grid((unsigned)p.nb, (unsigned)wv);
      // chains are indexed with 16 bits (0xFFFF = end): cap <= kMaxCap = 2048 always fits
      const size_t smem_small = (size_t)(2 * p.slots + p.cap) * 4 + (size_t)(2 * p.cap) * 2;
      if (pe <= 5)
        hvb_bucket_small_kernel<5><<<grid, kBucketThreads, smem_small, st>>>(w, pe);
      else if (pe <= 8)
        hvb_bucket_small_kernel<8><<<grid, kBucketThreads, smem_small, st>>>(w, pe);
      else {  // (sort keys hold 21-bit point indices: the plan admits at most 1024 x 2048 points)
        // table: the smallest power of two above cap (the plan's `slots` keeps the load under 0.8
        // even for a full bucket of distinct cells; here a full bucket is allowed to probe longer
        // in exchange for one more resident CTA)
        int rslots = 64;
        while (rslots <= p.cap) rslots <<= 1;
        rslots = std::min(rslots, p.slots);
        const size_t smem_rank = (size_t)p.cap * 16 + (size_t)rslots * 8;
        const int spec = (int)std::min<int64_t>(p.cap, std::max<int64_t>((((wn_max + p.nb - 1) / p.nb) + 63) / 64 * 64, 64));
        {
          const int rca = ensure_dyn_smem((const void*)hvb_bucket_rank_kernel, 3, device, smem_rank);
          if (rca != PCFE_OK) return rca;
        }
        hvb_bucket_rank_kernel<<<grid, kBucketThreads, smem_rank, st>>>(w, pe, spec, rslots);
      }
      PCFE_LAUNCH_CHECK();
    }
    rc = hv_launch_scan(w.zero, w.zero_stride, w.wordprefix, w.word_stride, wnpad / 32,
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
      const int64_t vmax = std::max<int64_t>(std::min<int64_t>(max_voxels, wn_max), 1);
      const int per_cta = kExpWarps * kExpTilesPerWarp * p.exp_vt;
      const dim3 grid((unsigned)((vmax + per_cta - 1) / per_cta), (unsigned)wv);
      const int32_t* vn = voxel_num + f0;
      if (max_points == 5 && (c == 4 || c == 5) && max_voxels < (1 << 24)) {
        const KeyDecode kd = make_key_decode(p.g);
        const int per = kExpWarps * kExpTilesPerWarp * 32;
        const dim3 fgrid((unsigned)((vmax + per - 1) / per), (unsigned)wv);
        const int pper = kExpWarps * kPipeTiles * 32;
        const dim3 pgrid((unsigned)((vmax + pper - 1) / pper), (unsigned)wv);
        if (vec_ok && g_opt_expand_variant == 0) {
          if (c == 4) hvb_expand_pipe_kernel<4, 5><<<pgrid, kExpThreads, 0, st>>>(b, w, kd, vn, wv, g_opt_expand_prefetch);
          else hvb_expand_pipe_kernel<5, 5><<<pgrid, kExpThreads, 0, st>>>(b, w, kd, vn, wv, g_opt_expand_prefetch);
        } else if (c == 4) hvb_expand_fixed_kernel<4, 5><<<fgrid, kExpThreads, 0, st>>>(b, w, kd, vn, vec_ok);
        else hvb_expand_fixed_kernel<5, 5><<<fgrid, kExpThreads, 0, st>>>(b, w, kd, vn, vec_ok);
        PCFE_LAUNCH_CHECK();
      } else if (max_points * c >= 64 && (max_points * c) % 4 == 0 && vec_ok && g_opt_expand_variant == 0) {
        // long voxels: lane = output word, 32 voxels per warp, one after the other
        const KeyDecode kd = make_key_decode(p.g);
        const dim3 wgrid((unsigned)((vmax + kExpWarps * kWordsVox - 1) / (kExpWarps * kWordsVox)), (unsigned)wv);
        const int w4 = max_points * c / 4;
        uint32_t m_w4 = (int64_t)kWordsVox * w4 < 65536 ? (uint32_t)((65536 + w4 - 1) / w4) : 0u;
        for (int i = 0; m_w4 && i < kWordsVox * w4 + 32; ++i)
          if ((int)(((uint32_t)i * m_w4) >> 16) != i / w4) m_w4 = 0u;  // (the reciprocal must be exact on the range used)
        if (c == 4) hvb_expand_words_kernel<4><<<wgrid, kExpThreads, 0, st>>>(b, w, kd, c, max_points, vn, m_w4);
        else if (c == 5) hvb_expand_words_kernel<5><<<wgrid, kExpThreads, 0, st>>>(b, w, kd, c, max_points, vn, m_w4);
        else hvb_expand_words_kernel<0><<<wgrid, kExpThreads, 0, st>>>(b, w, kd, c, max_points, vn, m_w4);
        PCFE_LAUNCH_CHECK();
      } else if (c == 4) rc = launch_expand<4>(grid, st, b, w, p.g, c, max_points, p.exp_vt, vn, vec_ok);
      else if (c == 5) rc = launch_expand<5>(grid, st, b, w, p.g, c, max_points, p.exp_vt, vn, vec_ok);
      else rc = launch_expand<0>(grid, st, b, w, p.g, c, max_points, p.exp_vt, vn, vec_ok);
      if (rc != PCFE_OK) return rc;
    }
    }
    rc = hvg_launch_slow(b, wv, w.zero + w.ctl_off + p.nb + kCtlOverflow, w.zero_stride,
                         g_opt_force_overflow, w.region, w.region_stride, p.slow, w.zero,
                         w.zero_stride, w.wordprefix, w.word_stride, c, max_points, max_voxels,
                         voxel_num + f0, st, mean, pack ? 2 : 0, pack ? voxel_num : nullptr, f0);
    if (rc != PCFE_OK) return rc;
  }
  return PCFE_OK;
}

}  // namespace pcfe
