// detmatch_b200/csrc/voxelize.cu -- C-ABI entry points for hard and dynamic voxelization
// (include/pcfe.h), the dynamic-voxelize kernel, path selection and the optional profiler.
//
// Replaces the reference's voxel_layer.{hard_voxelize,dynamic_voxelize}
// (mmdet3d/ops/voxel/src/voxelization_cpu.cpp:7-169 is the behaviour reproduced bit for bit;
//  mmdet3d/ops/voxel/src/voxelization_cuda.cu:184-371 is the O(N^2) + <<<1,1>>> GPU path this
//  supersedes).  Hard voxelization has two implementations behind the same entry point:
//  hv_bucket.cu (shared-memory bucket path, the default) and hv_global.cu (global-memory path,
//  used for shapes the bucket path does not cover and as its overflow fallback).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "hv_common.cuh"

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


Knob g_opt_hv_path{0};
Knob g_opt_force_overflow{0};
Knob g_opt_hv_wave{0};
extern Knob g_opt_bucket_avg;
extern Knob g_opt_bucket_variant;
extern Knob g_opt_expand_variant;
extern Knob g_opt_expand_ctas;
extern Knob g_opt_cluster;
extern Knob g_opt_overlap;
extern Knob g_opt_bin_small;
extern Knob g_opt_warp_dedup;
extern Knob g_opt_expand_tiles;
extern Knob g_opt_pdl;
extern Knob g_opt_pib_grid;
extern Knob g_opt_no_fast_div;
extern Knob g_opt_expand_prefetch;
extern Knob g_opt_expand_map;

namespace {

constexpr int kThreads = 256;
constexpr int kPtsPerThread = 4;
constexpr int kTilePts = kThreads * kPtsPerThread;

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

// test hook: point_key_fast() against point_key() for every float32 bit pattern of x
__global__ void axis_sweep_kernel(const GridParams g, const int fast, unsigned long long* out) {
  const FastAxes fa = make_fast_axes(g);
  const float y = __fadd_rn(g.y0, g.vy), z = __fadd_rn(g.z0, g.vz);  // an in-range cell on the other axes
  unsigned long long bad = 0, first = ~0ull;
  const uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) * 256u;
  for (uint32_t k = 0; k < 256u; ++k) {
    const float x = __uint_as_float(base + k);
    int cx, cy, cz;
    const uint32_t want = point_key(x, y, z, g, cx, cy, cz);
    const uint32_t got = point_key_fast(x, y, z, g, fa, fast != 0);
    if (want != got) {
      ++bad;
      first = min(first, (unsigned long long)(base + k));
    }
  }
  if (bad) {
    atomicAdd(&out[0], bad);
    atomicMin(&out[1], first);
  }
}

__global__ void hv_zero_counts_kernel(int32_t* voxel_num, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) voxel_num[i] = 0;
}

constexpr size_t kL2ScratchBudget = 1ull << 30;  // measured: fewer, larger waves beat L2-resident scratch (launch gaps dominate)

struct HvChoice {
  bool bucket;
  HvBucketPlan bp;
  HvGlobalPlan gp;
  size_t per_frame;
};

int choose_path(int64_t n_max, int c, const float vs[3], const float rg[6], int max_points,
                int max_voxels, HvChoice* ch) {
  int rc = hvg_make_plan(n_max, vs, rg, max_points, max_voxels, &ch->gp);
  if (rc != PCFE_OK) return rc;
  ch->bucket = false;
  if (g_opt_hv_path != 1) {
    rc = hvb_make_plan(n_max, c, vs, rg, max_points, max_voxels, &ch->bp);
    if (rc == PCFE_OK) ch->bucket = true;
    else if (g_opt_hv_path == 2) return rc;
  }
  ch->per_frame = ch->bucket ? ch->bp.per_frame : ch->gp.per_frame;
  return PCFE_OK;
}

// Frames per wave.  Measured on B200 (C4): one 64-frame wave 0.66 ms; launch gaps make many
// small waves slower (wave 8: 1.05 ms).  Overlapping two 32-frame waves on two streams gains
// only ~1 % (every kernel already fills the SMs), so overlap is used only when a batch needs
// more than one wave anyway.
int auto_wave(size_t per_frame, int num_frames, bool bucket) {
  const int knob_wave = g_opt_hv_wave;
  if (knob_wave > 0) return std::min(std::min(knob_wave, kMaxWave), std::max(num_frames, 1));
  size_t w = kL2ScratchBudget / per_frame;
  w = std::max<size_t>(1, std::min<size_t>(w, (size_t)kMaxWave));
  (void)bucket;
  return (int)std::min<size_t>(w, (size_t)std::max(num_frames, 1));
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
  HvChoice ch;
  // c = 3: the most permissive row length (the bucket plan's eligibility shrinks with c, its scratch
  // size does not depend on it), so the size covers whichever plan the call with the real c selects
  if (choose_path(n_max, 3, vs, rg, max_points, max_voxels, &ch) != PCFE_OK) return 0;
  // sized for whichever path needs more, so a later pcfe_debug_set cannot invalidate it
  const size_t per = std::max(ch.per_frame, ch.gp.per_frame);
  int w = frames_in_flight > 0 ? std::min(frames_in_flight, kMaxWave) : auto_wave(per, num_frames, ch.bucket);
  w = std::min(w, std::max(num_frames, 1));
  const int nbuf = (ch.bucket && num_frames > w) ? 2 : 1;  // double-buffered waves overlap
  return (size_t)w * per * nbuf;
}

static int hv_batch_impl(const pcfe_frame_t* frames, int num_frames, int c, const float vs[3],
                         const float rg[6], const float* filter /* 6 floats or NULL */, int max_points,
                         int max_voxels, int32_t* voxel_num, void* workspace, size_t workspace_bytes,
                         int device, void* stream, int mode = 0) {
  if (num_frames < 0 || c < 3) return PCFE_ERR_SHAPE;
  if (num_frames == 0) return PCFE_OK;
  if (!frames || !voxel_num) return PCFE_ERR_NULL;
  int64_t n_max = 0;
  for (int k = 0; k < num_frames; ++k) {
    const pcfe_frame_t& fr = frames[k];
    if (fr.n < 0) return PCFE_ERR_SHAPE;
    if (fr.n > 0 && !fr.points) return PCFE_ERR_NULL;
    // (an empty frame produces no rows: its output pointers are never dereferenced)
    if (max_voxels > 0 && fr.n > 0 && (!fr.coors || !fr.num_points || (max_points > 0 && !fr.voxels)))
      return PCFE_ERR_NULL;
    if (((uintptr_t)fr.points & 3) || ((uintptr_t)fr.voxels & 3) || ((uintptr_t)fr.coors & 3) ||
        ((uintptr_t)fr.num_points & 3))
      return PCFE_ERR_ALIGN;
    n_max = std::max(n_max, fr.n);
  }
  HvChoice ch;
  int rc = choose_path(n_max, c, vs, rg, max_points, max_voxels, &ch);
  if (rc != PCFE_OK) return rc;
  if (filter) {  // fused PointsRangeFilter: every path reads it from its GridParams
    for (GridParams* g : {&ch.bp.g, &ch.bp.slow.g, &ch.gp.g}) {
      g->filter = 1;
      for (int j = 0; j < 3; ++j) {
        g->flo[j] = filter[j];
        g->fhi[j] = filter[3 + j];
      }
    }
  }

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
  const size_t fit = workspace_bytes / ch.per_frame;  // frames whose scratch fits
  if (fit < 1) return PCFE_ERR_WORKSPACE;
  int wave = auto_wave(ch.per_frame, num_frames, ch.bucket);
  int nbuf = (ch.bucket && num_frames > wave && fit >= 2 * (size_t)wave) ? 2 : 1;
  if ((size_t)wave * nbuf > fit) {  // smaller workspace than recommended: single buffer, smaller waves
    nbuf = 1;
    wave = (int)std::min<size_t>(wave, fit);
  }
  if (mode) {  // mean epilogue / packed output: record path of the bucket launch sequence only
    if (!ch.bucket) return PCFE_ERR_SHAPE;
    return hvb_run(frames, num_frames, c, ch.bp, max_points, max_voxels, voxel_num, workspace, wave, nbuf,
                   device, st, mode);
  }
  if (ch.bucket)
    return hvb_run(frames, num_frames, c, ch.bp, max_points, max_voxels, voxel_num, workspace, wave,
                   nbuf, device, st);
  return hvg_run(frames, num_frames, c, ch.gp, max_points, max_voxels, voxel_num, workspace, wave, st);
}

extern "C" int pcfe_hard_voxelize_batch_f32(const pcfe_frame_t* frames, int num_frames, int c,
                                            const float vs[3], const float rg[6], int max_points,
                                            int max_voxels, int32_t* voxel_num, void* workspace,
                                            size_t workspace_bytes, int device, void* stream) {
  return hv_batch_impl(frames, num_frames, c, vs, rg, nullptr, max_points, max_voxels, voxel_num, workspace,
                       workspace_bytes, device, stream);
}

extern "C" int pcfe_hard_voxelize_batch_filtered_f32(const pcfe_frame_t* frames, int num_frames, int c,
                                                     const float vs[3], const float rg[6],
                                                     const float filter_range[6], int max_points,
                                                     int max_voxels, int32_t* voxel_num, void* workspace,
                                                     size_t workspace_bytes, int device, void* stream) {
  if (!filter_range) return PCFE_ERR_NULL;
  return hv_batch_impl(frames, num_frames, c, vs, rg, filter_range, max_points, max_voxels, voxel_num,
                       workspace, workspace_bytes, device, stream);
}

extern "C" int pcfe_hard_voxelize_mean_batch_f32(const pcfe_frame_t* frames, int num_frames, int c,
                                                 const float vs[3], const float rg[6],
                                                 const float* filter_range, int max_points, int max_voxels,
                                                 int32_t* voxel_num, void* workspace, size_t workspace_bytes,
                                                 int device, void* stream) {
  if (max_points != 5 || (c != 4 && c != 5)) return PCFE_ERR_SHAPE;
  return hv_batch_impl(frames, num_frames, c, vs, rg, filter_range, max_points, max_voxels, voxel_num,
                       workspace, workspace_bytes, device, stream, kHvMean);
}

extern "C" int pcfe_hard_voxelize_packed_batch_f32(const float* const* points, const int64_t* n, int num_frames,
                                                   int c, const float vs[3], const float rg[6],
                                                   const float* filter_range, int max_points, int max_voxels,
                                                   int mean, float* voxels_cat, int32_t* coors_batch,
                                                   int32_t* num_points_cat, int64_t cap_rows, int32_t* voxel_num,
                                                   void* workspace, size_t workspace_bytes, int device,
                                                   void* stream) {
  if (max_points != 5 || (c != 4 && c != 5)) return PCFE_ERR_SHAPE;
  if (num_frames < 0 || cap_rows < 0) return PCFE_ERR_SHAPE;
  if (num_frames == 0) return PCFE_OK;
  if (!points || !n) return PCFE_ERR_NULL;
  if ((uintptr_t)coors_batch & 15) return PCFE_ERR_ALIGN;
  std::vector<pcfe_frame_t> frames((size_t)num_frames);
  int64_t worst = 0;  // rows the batch can produce at most
  for (int k = 0; k < num_frames; ++k) {
    if (n[k] < 0) return PCFE_ERR_SHAPE;
    frames[k] = pcfe_frame_t{points[k], n[k], voxels_cat, coors_batch, num_points_cat};
    worst += std::min<int64_t>(n[k], std::max(max_voxels, 0));
  }
  if (cap_rows < worst) return PCFE_ERR_WORKSPACE;
  return hv_batch_impl(frames.data(), num_frames, c, vs, rg, filter_range, max_points, max_voxels, voxel_num,
                       workspace, workspace_bytes, device, stream, kHvPack | (mean ? kHvMean : 0));
}

// One thread per output element; a warp's loads of slot s cover 32 consecutive words of at most
// ceil(32 / c) + 1 voxels, the remaining words of those sectors are used by the next slots (L1).
__global__ void __launch_bounds__(256)
voxel_mean_kernel(const float* __restrict__ voxels, const int32_t* __restrict__ num_points,
                  const int32_t* __restrict__ voxel_num, const long long m_cap, const int max_points,
                  const int c, float* __restrict__ out) {
  const long long m = voxel_num ? min((long long)__ldg(voxel_num), m_cap) : m_cap;
  const long long total = m * c;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int q;
    const long long v = elem_row(e, c, q);
    const float* src = voxels + (size_t)v * max_points * c + q;
    float a = __ldg(src);
#pragma unroll 5
    for (int sl = 1; sl < max_points; ++sl) a = __fadd_rn(a, __ldg(src + (size_t)sl * c));
    out[e] = __fdiv_rn(a, (float)__ldg(num_points + v));
  }
}

// Short voxels (max_points * c <= 64 words): a warp stages 32 voxels -- one contiguous run of
// 32 * W words, read as 128-byte coalesced loads -- in shared memory, then lane = output word adds
// the slots of its (voxel, feature) in slot order and the (32, nf) tile leaves as coalesced stores.
constexpr int kVfeWarps = 8, kVfeMaxW = 64;

__global__ void __launch_bounds__(kVfeWarps * 32)
voxel_mean_tile_kernel(const float* __restrict__ voxels, const int32_t* __restrict__ num_points,
                       const int32_t* __restrict__ voxel_num, const long long m_cap, const int max_points,
                       const int c, const int nf /* row stride of voxels is c, the first nf features are reduced */,
                       float* __restrict__ out) {
  extern __shared__ float vfe_smem[];
  const long long m = voxel_num ? min((long long)__ldg(voxel_num), m_cap) : m_cap;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int W = max_points * c;
  float* buf = vfe_smem + (size_t)wid * 32 * W;
  const long long tiles = (m + 31) / 32;
  for (long long tile = (long long)blockIdx.x * kVfeWarps + wid; tile < tiles; tile += (long long)gridDim.x * kVfeWarps) {
    const long long v0 = tile * 32;
    const int nvox = (int)min(32ll, m - v0);
    const float* __restrict__ src = voxels + (size_t)v0 * W;
    const int words = nvox * W;
    for (int i = lane; i < words; i += 32) buf[i] = __ldg(src + i);
    __syncwarp();
    const int outs = nvox * nf;
    float* __restrict__ dst = out + (size_t)v0 * nf;
    for (int o = lane; o < outs; o += 32) {
      const int v = o / nf, q = o - v * nf;
      const float* row = buf + v * W + q;
      float a = row[0];
      for (int sl = 1; sl < max_points; ++sl) a = __fadd_rn(a, row[sl * c]);
      dst[o] = __fdiv_rn(a, (float)__ldg(num_points + v0 + v));
    }
    __syncwarp();  // the tile buffer is reused
  }
}

extern "C" int pcfe_voxel_mean_f32(const float* voxels, const int32_t* num_points, const int32_t* voxel_num,
                                   int64_t m, int max_points, int c, float* out, int device, void* stream) {
  if (m < 0 || max_points < 1 || c < 1) return PCFE_ERR_SHAPE;
  if (m == 0) return PCFE_OK;
  if (!voxels || !num_points || !out) return PCFE_ERR_NULL;
  if (((uintptr_t)voxels & 3) || ((uintptr_t)num_points & 3) || ((uintptr_t)out & 3) || ((uintptr_t)voxel_num & 3))
    return PCFE_ERR_ALIGN;
  DeviceGuard guard(device);
  PCFE_CUDA_TRY(guard.err);
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)m * c;
  const long long want = (total + 255) / 256;
  const int grid = (int)std::min<long long>(want, 148ll * 8 * 16);
  ProfScope ps("voxel_mean", st);
  if ((long long)max_points * c <= kVfeMaxW) {
    const size_t smem = (size_t)kVfeWarps * 32 * max_points * c * sizeof(float);  // <= 64 KB
    PCFE_CUDA_TRY(cudaFuncSetAttribute(voxel_mean_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tiles = (m + 31) / 32;
    const int tgrid = (int)std::min<long long>((tiles + kVfeWarps - 1) / kVfeWarps, 148ll * 8 * 8);
    voxel_mean_tile_kernel<<<tgrid, kVfeWarps * 32, smem, st>>>(voxels, num_points, voxel_num, m, max_points, c, c, out);
    PCFE_LAUNCH_CHECK();
    return PCFE_OK;
  }
  voxel_mean_kernel<<<grid, 256, 0, st>>>(voxels, num_points, voxel_num, m, max_points, c, out);
  PCFE_LAUNCH_CHECK();
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

extern "C" int pcfe_debug_axis_sweep(float lo, float vs, float hi, uint64_t* out, int device, void* stream) {
  if (!out) return PCFE_ERR_NULL;
  const float v3[3] = {vs, vs, vs};
  const float r6[6] = {lo, lo, lo, hi, hi, hi};
  GridParams g;
  make_grid_params(v3, r6, &g);
  if (g.gx <= 0) return PCFE_ERR_GRID;
  DeviceGuard guard(device);
  PCFE_CUDA_TRY(guard.err);
  cudaStream_t st = (cudaStream_t)stream;
  const uint64_t init[2] = {0ull, ~0ull};
  PCFE_CUDA_TRY(cudaMemcpyAsync(out, init, sizeof init, cudaMemcpyHostToDevice, st));
  axis_sweep_kernel<<<(1u << 24) / 256u, 256, 0, st>>>(g, fast_div_sizes_ok(g) ? 1 : 0, (unsigned long long*)out);
  PCFE_LAUNCH_CHECK();
  return PCFE_OK;
}

// Test / tuning knobs (every setting computes the same, bit-exact results; they select code paths
// and launch parameters): "hv_path" (0 auto, 1 global-memory path, 2 bucket path),
// "hv_cluster" (1: one cluster per frame instead of the record path's launch sequence), "hv_overlap" (0: the
// waves of a multi-wave batch run one after the other instead of on two internal streams), "hv_bin_small"
// (partition tile: 0 = 4096 points, 1 = 1024, 2 = by batch size), "hv_warp_dedup" (1: __match_any_sync key
// de-duplication in front of the bucket table), "hv_force_overflow" (1: every frame also runs the overflow fallback), "hv_bucket_avg"
// (target points per bucket), "hv_wave" (frames per launch sequence), "hv_bucket_variant" (1: general
// kernels instead of the record path), "hv_expand_variant" (1: un-pipelined expansion kernels),
// "hv_expand_prefetch" (frames of L2 prefetch distance), "hv_no_fast_div", "hv_pdl",
// "hv_expand_ctas" (persistent expansion), "hv_expand_map" (record expansion, tiles of a warp: 0 consecutive, 1 round-robin
// inside the CTA, 2 round-robin over the frame = default), "pib_grid" (0: brute-force first-hit point-in-box assignment).  Returns
// PCFE_ERR_SHAPE for an unknown name.
extern "C" int pcfe_debug_set(const char* name, int value) {
  if (!name) return PCFE_ERR_NULL;
#ifdef PCFE_NO_DEBUG_KNOBS
  (void)value;
  return PCFE_ERR_SHAPE;  // production build: the knobs are not reachable
#endif
  if (!strcmp(name, "hv_path")) g_opt_hv_path = value;
  else if (!strcmp(name, "hv_force_overflow")) g_opt_force_overflow = value;
  else if (!strcmp(name, "hv_bucket_avg")) g_opt_bucket_avg = value;
  else if (!strcmp(name, "hv_wave")) g_opt_hv_wave = value;
  else if (!strcmp(name, "hv_bucket_variant")) g_opt_bucket_variant = value;
  else if (!strcmp(name, "hv_expand_variant")) g_opt_expand_variant = value;
  else if (!strcmp(name, "hv_expand_prefetch")) g_opt_expand_prefetch = value;
  else if (!strcmp(name, "hv_no_fast_div")) g_opt_no_fast_div = value;
  else if (!strcmp(name, "hv_pdl")) g_opt_pdl = value;
  else if (!strcmp(name, "pib_grid")) g_opt_pib_grid = value;
  else if (!strcmp(name, "hv_expand_ctas")) g_opt_expand_ctas = value;
  else if (!strcmp(name, "hv_cluster")) g_opt_cluster = value;
  else if (!strcmp(name, "hv_overlap")) g_opt_overlap = value;
  else if (!strcmp(name, "hv_bin_small")) g_opt_bin_small = value;
  else if (!strcmp(name, "hv_warp_dedup")) g_opt_warp_dedup = value;
  else if (!strcmp(name, "hv_expand_tiles")) g_opt_expand_tiles = value;
  else if (!strcmp(name, "hv_expand_map")) g_opt_expand_map = value;
  else return PCFE_ERR_SHAPE;
  return PCFE_OK;
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
