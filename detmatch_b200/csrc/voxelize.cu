// detmatch_b200/csrc/voxelize.cu -- hard and dynamic voxelization for sm_100a.
//
// Replaces the reference's voxel_layer.{hard_voxelize,dynamic_voxelize}
// (mmdet3d/ops/voxel/src/voxelization_cpu.cpp:7-169 is the behaviour reproduced bit for bit;
//  mmdet3d/ops/voxel/src/voxelization_cuda.cu:184-371 is the O(N^2) + <<<1,1>>> GPU path this
//  supersedes).  See include/pcfe.h for the contract and DESIGN.md for the data layout.
//
// Hard voxelization, per wave of frames (all kernels are batched over frames with gridDim.y):
//   K1 hv_key_hash   point -> linear cell key -> slot of a per-frame table (direct-mapped when
//                    the grid is small, open-addressing hash otherwise); atomicMin(first point
//                    index) per cell.  Remembers the slot of every point.
//   K2 hv_first_flag a point is "first" iff it is its cell's minimum index; one ballot per warp
//                    gives a bitmask over point indices.
//   K3 hv_scan_flags per frame: exclusive popcount-prefix of the bitmask -> rank of every first
//                    point = voxel id in first-occurrence order; voxel_num = min(#first, V).
//   K4 hv_assign     every point of a kept voxel inserts its index into that voxel's sorted
//                    P-entry list (atomicMin chain: the list converges to the P smallest
//                    indices in ascending order, independent of arrival order); the first
//                    point writes coors.
//   K5 hv_scatter    voxel-id-ordered, coalesced write of complete voxel rows (data + zero
//                    padding) and num_points_per_voxel; point rows are gathered through L2.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "pcfe_common.cuh"

namespace pcfe {

std::atomic<uint64_t> g_launches{0};

// ---- optional per-kernel timing ------------------------------------------------------------
bool g_prof_on = false;
namespace {
struct ProfRec { const char* name; cudaEvent_t a, b; };
std::vector<ProfRec> g_prof_recs;
std::mutex g_prof_mu;
}  // namespace
void prof_begin(const char* name, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{name, nullptr, nullptr};
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  g_prof_recs.push_back(r);
}
void prof_end(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_recs.empty()) cudaEventRecord(g_prof_recs.back().b, st);
}

namespace {

constexpr int kMaxWave = 64;       // frames per launch sequence (kernel-parameter table size)
constexpr int kThreads = 256;
constexpr int kPtsPerThread = 4;
constexpr int kTilePts = kThreads * kPtsPerThread;

struct HvFrame {
  const float* pts;
  float* voxels;
  int32_t* coors;
  int32_t* num;
  int n;
  int pad_;
};

struct HvBatch {
  HvFrame f[kMaxWave];
};

struct HvWork {            // per-wave scratch, frame-major with the strides below (elements)
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

__device__ __forceinline__ void load_xyz(const float* __restrict__ pts, int i, int c, float& x,
                                         float& y, float& z) {
  const float* p = pts + (size_t)i * c;
  x = __ldg(p);
  y = __ldg(p + 1);
  z = __ldg(p + 2);
}

// ------------------------------------------------------------------------------------------
// dynamic voxelization: one thread per point
// ------------------------------------------------------------------------------------------
struct DynFrame {
  const float* pts;
  int32_t* coors;
  int n;
  int pad_;
};
struct DynBatch {
  DynFrame f[kMaxWave];
};

__global__ void __launch_bounds__(kThreads)
dyn_voxelize_kernel(const __grid_constant__ DynBatch batch, const GridParams g, const int c) {
  const DynFrame& fr = batch.f[blockIdx.y];
  const int base = blockIdx.x * kTilePts + threadIdx.x;
#pragma unroll
  for (int k = 0; k < kPtsPerThread; ++k) {
    const int i = base + k * kThreads;
    if (i >= fr.n) break;
    float x, y, z;
    load_xyz(fr.pts, i, c, x, y, z);
    int cx, cy, cz;
    const uint32_t key = point_key(x, y, z, g, cx, cy, cz);
    int32_t* o = fr.coors + (size_t)i * 3;
    const bool ok = key != kEmpty;  // voxelization_cpu.cpp:32-37: all three -1 on failure
    o[0] = ok ? cz : -1;
    o[1] = ok ? cy : -1;
    o[2] = ok ? cx : -1;
  }
}

// ------------------------------------------------------------------------------------------
// K1: key + table insert
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_slot(uint32_t key, int log2_slots) {
  return (key * 0x9E3779B1u) >> (32 - log2_slots);
}

__global__ void __launch_bounds__(kThreads)
hv_key_hash_kernel(const __grid_constant__ HvBatch batch, const HvWork w, const GridParams g,
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
      uint32_t s;
      if (w.direct) {
        s = key[k];
      } else {
        s = hash_slot(key[k], w.log2_slots);
        const uint32_t mask = w.slots - 1u;
        while (true) {
          uint32_t cur = __ldcg(&table[s].x);
          if (cur == key[k]) break;
          if (cur == kEmpty) {
            cur = atomicCAS(&table[s].x, kEmpty, key[k]);
            if (cur == kEmpty || cur == key[k]) break;
          }
          s = (s + 1u) & mask;
        }
      }
      // first point index of the cell; a (possibly stale) smaller value means i cannot win
      if (__ldcg(&table[s].y) > (uint32_t)i) atomicMin(&table[s].y, (uint32_t)i);
      slot = (int32_t)s;
    }
    pslot[i] = slot;
  }
}

// ------------------------------------------------------------------------------------------
// K2: first-occurrence flags (bit per point index)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
hv_first_flag_kernel(const __grid_constant__ HvBatch batch, const HvWork w, const int npad) {
  const int f = blockIdx.y;
  const int n = batch.f[f].n;
  const uint2* __restrict__ table = w.table + (size_t)f * w.table_stride;
  const int32_t* __restrict__ pslot = w.pslot + (size_t)f * w.pslot_stride;
  uint32_t* __restrict__ bitmask = w.bitmask + (size_t)f * w.word_stride;
  const int base = blockIdx.x * kTilePts + threadIdx.x;
#pragma unroll
  for (int k = 0; k < kPtsPerThread; ++k) {
    const int i = base + k * kThreads;  // warp-uniform validity of i < npad (npad % 32 == 0)
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

// ------------------------------------------------------------------------------------------
// K3: per-frame exclusive popcount prefix over the bitmask words
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads)
hv_scan_flags_kernel(const __grid_constant__ HvBatch batch, const HvWork w, const int words,
                     const int max_voxels, int32_t* __restrict__ voxel_num) {
  const int f = blockIdx.x;
  const uint32_t* __restrict__ bitmask = w.bitmask + (size_t)f * w.word_stride;
  uint32_t* __restrict__ wordprefix = w.wordprefix + (size_t)f * w.word_stride;
  __shared__ uint32_t warp_sums[kScanThreads / 32];
  __shared__ uint32_t carry_s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  // words handled per thread, consecutive, so the scan is a plain blocked scan
  const int per = (words + kScanThreads - 1) / kScanThreads;
  const int w0 = threadIdx.x * per;
  uint32_t local = 0;
  for (int j = 0; j < per; ++j) {
    const int idx = w0 + j;
    if (idx < words) local += __popc(bitmask[idx]);
  }
  // inclusive warp scan of `local`
  uint32_t incl = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t v = warp_sums[lane];
    uint32_t vi = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, vi, d);
      if (lane >= d) vi += t;
    }
    warp_sums[lane] = vi - v;  // exclusive
    if (lane == 31) carry_s = vi;
  }
  __syncthreads();
  uint32_t run = warp_sums[wid] + (incl - local);
  for (int j = 0; j < per; ++j) {
    const int idx = w0 + j;
    if (idx < words) {
      wordprefix[idx] = run;
      run += __popc(bitmask[idx]);
    }
  }
  if (threadIdx.x == 0) {
    const uint32_t total = carry_s;
    voxel_num[f] = (int32_t)min(total, (uint32_t)max_voxels);
  }
}

// ------------------------------------------------------------------------------------------
// K4: voxel ids, coors, sorted per-voxel point lists
// ------------------------------------------------------------------------------------------
// The list converges to the P smallest inserted values in ascending order whatever the
// interleaving: every slot only ever decreases, a value moves on to slot s+1 exactly when a
// smaller one holds slot s, and displaced values are carried forward by the displacing thread.
__device__ __forceinline__ void sorted_insert(uint32_t* __restrict__ lst, const int p, uint32_t v) {
  if (__ldcg(&lst[p - 1]) < v) return;  // list already full of smaller indices
  for (int s = 0; s < p; ++s) {
    if (__ldcg(&lst[s]) < v) continue;  // monotone: a stale read is only ever conservative
    const uint32_t old = atomicMin(&lst[s], v);
    if (old == kEmpty) return;
    if (old > v) v = old;
  }
}

__global__ void __launch_bounds__(kThreads)
hv_assign_kernel(const __grid_constant__ HvBatch batch, const HvWork w, const GridParams g,
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
    const uint32_t vid = wordprefix[m >> 5] + __popc(bitmask[m >> 5] & ((1u << (m & 31)) - 1u));
    if (vid >= (uint32_t)max_voxels) continue;  // voxelization_cpu.cpp:78
    if (m == (uint32_t)i) {
      const uint32_t key = w.direct ? (uint32_t)s : e.x;
      const uint32_t plane = (uint32_t)g.gx * (uint32_t)g.gy;
      const uint32_t cz = key / plane;
      const uint32_t rem = key - cz * plane;
      const uint32_t cy = rem / (uint32_t)g.gx;
      const uint32_t cx = rem - cy * (uint32_t)g.gx;
      int32_t* o = fr.coors + (size_t)vid * 3;
      o[0] = (int32_t)cz;
      o[1] = (int32_t)cy;
      o[2] = (int32_t)cx;
    }
    if (max_points > 0) sorted_insert(idxlist + (size_t)vid * max_points, max_points, (uint32_t)i);
  }
}

// ------------------------------------------------------------------------------------------
// K5: voxel-ordered scatter of complete rows
// ------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kThreads)
hv_scatter_kernel(const __grid_constant__ HvBatch batch, const HvWork w, const int c_rt,
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
  // num_points_per_voxel: number of occupied list entries (the list is ascending, kEmpty last)
  const long long v0 = (long long)blockIdx.x * kTilePts + threadIdx.x;
#pragma unroll
  for (int k = 0; k < kPtsPerThread; ++k) {
    const long long v = v0 + (long long)k * kThreads;
    if (v >= m) break;
    const uint32_t* lst = idxlist + (size_t)v * max_points;
    int cnt = 0;
    for (int s = 0; s < max_points; ++s) cnt += (__ldcg(&lst[s]) != kEmpty) ? 1 : 0;
    fr.num[v] = cnt;
  }
}

__global__ void hv_zero_counts_kernel(int32_t* voxel_num, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) voxel_num[i] = 0;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct HvPlan {
  GridParams g;
  uint64_t cells;
  int npad, words;
  uint32_t slots;
  int log2_slots, direct;
  size_t table_b, pslot_b, word_b, list_b, per_frame;
};

int make_plan(int64_t n_max, const float vs[3], const float rg[6], int max_points, int max_voxels,
              HvPlan* p) {
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
    uint64_t want = std::max<uint64_t>(1024, ((uint64_t)p->npad * 4 + 2) / 3);
    int lg = 10;
    while ((1ull << lg) < want) ++lg;
    p->log2_slots = lg;
    p->slots = 1u << lg;
  }
  p->table_b = align256((size_t)p->slots * sizeof(uint2));
  p->list_b = align256(std::max<size_t>((size_t)max_voxels * (size_t)max_points, 1) * sizeof(uint32_t));
  p->pslot_b = align256((size_t)p->npad * sizeof(int32_t));
  p->word_b = align256((size_t)p->words * sizeof(uint32_t));
  p->per_frame = p->table_b + p->list_b + p->pslot_b + 2 * p->word_b;
  return PCFE_OK;
}

constexpr size_t kL2ScratchBudget = 40ull << 20;  // keep a wave's scratch resident in L2

int auto_wave(const HvPlan& p, int num_frames) {
  size_t w = kL2ScratchBudget / p.per_frame;
  w = std::max<size_t>(1, std::min<size_t>(w, (size_t)kMaxWave));
  return (int)std::min<size_t>(w, (size_t)std::max(num_frames, 1));
}

template <int C>
void launch_scatter(dim3 grid, cudaStream_t st, const HvBatch& b, const HvWork& w, int c, int p,
                    const int32_t* vn) {
  hv_scatter_kernel<C><<<grid, kThreads, 0, st>>>(b, w, c, p, vn);
}

}  // namespace
}  // namespace pcfe

using namespace pcfe;

extern "C" int pcfe_grid_size(const float vs[3], const float rg[6], int32_t grid[3]) {
  if (!vs || !rg || !grid) return PCFE_ERR_NULL;
  for (int j = 0; j < 3; ++j) {
    volatile float span = rg[3 + j] - rg[j];  // float32 subtract, then float32 divide
    volatile float q = span / vs[j];
    grid[j] = (int32_t)roundf(q);
  }
  return PCFE_OK;
}

extern "C" int pcfe_dynamic_voxelize_batch_f32(const float* const* points, const int64_t* n,
                                               int num_frames, int c, const float vs[3],
                                               const float rg[6], int32_t* const* coors,
                                               int device, void* stream) {
  if (num_frames < 0 || c < 3) return PCFE_ERR_SHAPE;
  if (num_frames == 0) return PCFE_OK;
  if (!points || !n || !coors || !vs || !rg) return PCFE_ERR_NULL;
  GridParams g;
  make_grid_params(vs, rg, &g);
  if (g.gx <= 0 || g.gy <= 0 || g.gz <= 0) return PCFE_ERR_GRID;
  for (int k = 0; k < num_frames; ++k) {  // validate everything before touching the device
    if (n[k] < 0) return PCFE_ERR_SHAPE;
    if (n[k] >= 0x7FFFFFFF - 4096) return PCFE_ERR_TOO_LARGE;
    if (n[k] > 0 && (!points[k] || !coors[k])) return PCFE_ERR_NULL;
    if (((uintptr_t)points[k] & 3) || ((uintptr_t)coors[k] & 3)) return PCFE_ERR_ALIGN;
  }
  DeviceGuard guard(device);
  PCFE_CUDA_TRY(guard.err);
  cudaStream_t st = (cudaStream_t)stream;
  for (int f0 = 0; f0 < num_frames; f0 += kMaxWave) {
    const int wv = std::min(kMaxWave, num_frames - f0);
    DynBatch b;
    int64_t n_max = 0;
    for (int k = 0; k < wv; ++k) {
      b.f[k] = DynFrame{points[f0 + k], coors[f0 + k], (int)n[f0 + k], 0};
      n_max = std::max(n_max, n[f0 + k]);
    }
    if (n_max == 0) continue;
    dim3 grid((unsigned)((n_max + kTilePts - 1) / kTilePts), (unsigned)wv);
    dyn_voxelize_kernel<<<grid, kThreads, 0, st>>>(b, g, c);
    PCFE_LAUNCH_CHECK();
  }
  return PCFE_OK;
}

extern "C" int pcfe_dynamic_voxelize_f32(const float* points, int64_t n, int c, const float vs[3],
                                         const float rg[6], int32_t* coors, int device,
                                         void* stream) {
  return pcfe_dynamic_voxelize_batch_f32(&points, &n, 1, c, vs, rg, &coors, device, stream);
}

extern "C" size_t pcfe_hard_voxelize_workspace_bytes(int64_t n_max, int num_frames,
                                                     int frames_in_flight, const float vs[3],
                                                     const float rg[6], int max_points,
                                                     int max_voxels) {
  HvPlan p;
  if (make_plan(n_max, vs, rg, max_points, max_voxels, &p) != PCFE_OK) return 0;
  int w = frames_in_flight > 0 ? std::min(frames_in_flight, kMaxWave) : auto_wave(p, num_frames);
  w = std::min(w, std::max(num_frames, 1));
  return (size_t)w * p.per_frame;
}

extern "C" int pcfe_hard_voxelize_batch_f32(const pcfe_frame_t* frames, int num_frames, int c,
                                            const float vs[3], const float rg[6], int max_points,
                                            int max_voxels, int32_t* voxel_num, void* workspace,
                                            size_t workspace_bytes, int device, void* stream) {
  if (num_frames < 0 || c < 3) return PCFE_ERR_SHAPE;
  if (num_frames == 0) return PCFE_OK;
  if (!frames || !voxel_num) return PCFE_ERR_NULL;
  int64_t n_max = 0;
  for (int k = 0; k < num_frames; ++k) {
    const pcfe_frame_t& fr = frames[k];
    if (fr.n < 0) return PCFE_ERR_SHAPE;
    if (fr.n > 0 && !fr.points) return PCFE_ERR_NULL;
    if (max_voxels > 0 && (!fr.coors || !fr.num_points || (max_points > 0 && !fr.voxels)))
      return PCFE_ERR_NULL;
    if (((uintptr_t)fr.points & 3) || ((uintptr_t)fr.voxels & 3) || ((uintptr_t)fr.coors & 3) ||
        ((uintptr_t)fr.num_points & 3))
      return PCFE_ERR_ALIGN;
    n_max = std::max(n_max, fr.n);
  }
  HvPlan p;
  int rc = make_plan(n_max, vs, rg, max_points, max_voxels, &p);
  if (rc != PCFE_OK) return rc;

  DeviceGuard guard(device);
  PCFE_CUDA_TRY(guard.err);
  cudaStream_t st = (cudaStream_t)stream;

  if (n_max == 0 || max_voxels == 0) {  // nothing can be recorded: voxel_num = 0 for every frame
    hv_zero_counts_kernel<<<(num_frames + 255) / 256, 256, 0, st>>>(voxel_num, num_frames);
    PCFE_LAUNCH_CHECK();
    return PCFE_OK;
  }
  if (!workspace) return PCFE_ERR_NULL;
  if ((uintptr_t)workspace & 255) return PCFE_ERR_ALIGN;
  const int wave = (int)std::min<size_t>(std::min(kMaxWave, num_frames), workspace_bytes / p.per_frame);
  if (wave < 1) return PCFE_ERR_WORKSPACE;

  // scratch layout: [tables | lists] (both memset to 0xFF) | pslot | bitmask | wordprefix
  char* base = (char*)workspace;
  HvWork w;
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
    const int wnpad = (int)((wn_max + 31) / 32 * 32);
    const int wwords = std::max(wnpad / 32, 1);
    const dim3 pgrid((unsigned)std::max<int64_t>((wnpad + kTilePts - 1) / kTilePts, 1), (unsigned)wv);
    {
      ProfScope ps("hv_key_hash", st);
      hv_key_hash_kernel<<<pgrid, kThreads, 0, st>>>(b, w, p.g, c);
      PCFE_LAUNCH_CHECK();
    }
    {
      ProfScope ps("hv_first_flag", st);
      hv_first_flag_kernel<<<pgrid, kThreads, 0, st>>>(b, w, std::max(wnpad, 32));
      PCFE_LAUNCH_CHECK();
    }
    {
      ProfScope ps("hv_scan_flags", st);
      hv_scan_flags_kernel<<<wv, kScanThreads, 0, st>>>(b, w, wwords, max_voxels, voxel_num + f0);
      PCFE_LAUNCH_CHECK();
    }
    {
      ProfScope ps("hv_assign", st);
      hv_assign_kernel<<<pgrid, kThreads, 0, st>>>(b, w, p.g, max_points, max_voxels);
      PCFE_LAUNCH_CHECK();
    }
    // rows are bounded by both caps and by the points that exist
    const int64_t vmax = std::min<int64_t>(max_voxels, wn_max);
    const int64_t rows = std::max<int64_t>(vmax * std::max(max_points, 1), 1);
    const dim3 sgrid((unsigned)((rows + kTilePts - 1) / kTilePts), (unsigned)wv);
    {
      ProfScope ps("hv_scatter", st);
      if (vec4_ok) launch_scatter<4>(sgrid, st, b, w, c, max_points, voxel_num + f0);
      else if (c == 5) launch_scatter<5>(sgrid, st, b, w, c, max_points, voxel_num + f0);
      else launch_scatter<0>(sgrid, st, b, w, c, max_points, voxel_num + f0);
      PCFE_LAUNCH_CHECK();
    }
  }
  return PCFE_OK;
}

extern "C" int pcfe_hard_voxelize_f32(const float* points, int64_t n, int c, const float vs[3],
                                      const float rg[6], int max_points, int max_voxels,
                                      float* voxels, int32_t* coors, int32_t* num_points,
                                      int32_t* voxel_num, void* workspace, size_t workspace_bytes,
                                      int device, void* stream) {
  pcfe_frame_t fr{points, n, voxels, coors, num_points};
  return pcfe_hard_voxelize_batch_f32(&fr, 1, c, vs, rg, max_points, max_voxels, voxel_num,
                                      workspace, workspace_bytes, device, stream);
}

extern "C" int pcfe_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  return PCFE_OK;
}

// Synchronises the recorded events, writes one "name total_ms launches" line per kernel into
// buf (NUL terminated, truncated to cap) and clears the records.  Returns the number of lines.
extern "C" int pcfe_profile_report(char* buf, size_t cap) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::map<std::string, std::pair<double, int>> acc;
  std::vector<std::string> order;
  for (auto& r : g_prof_recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess) cudaEventElapsedTime(&ms, r.a, r.b);
    if (!acc.count(r.name)) order.push_back(r.name);
    acc[r.name].first += ms;
    acc[r.name].second += 1;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof_recs.clear();
  std::string out;
  for (auto& n : order) {
    char line[160];
    snprintf(line, sizeof(line), "%s %.6f %d\n", n.c_str(), acc[n].first, acc[n].second);
    out += line;
  }
  if (buf && cap > 0) {
    const size_t k = std::min(cap - 1, out.size());
    memcpy(buf, out.data(), k);
    buf[k] = 0;
  }
  return (int)order.size();
}
