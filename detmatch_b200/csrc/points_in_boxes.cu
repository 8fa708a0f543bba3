// detmatch_b200/csrc/points_in_boxes.cu -- point-in-rotated-box assignment for sm_100a.
//
// Replaces roiaware_pool3d_ext.points_in_boxes_{gpu,batch,cpu}
// (mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cuda.cu:51-203 are the kernels superseded;
//  mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cpu.cpp:16-40 is the arithmetic reproduced
//  bit for bit).  See include/pcfe.h for the contract.
//
// Structure: a prologue kernel turns every box into 8 floats {cx, cy, cz_centre, h/2, cos, sin,
// l/2, w/2} -- the trigonometry is glibc's sinf/cosf algorithm evaluated in double with the same
// fused operations the x86-64 FMA build of glibc uses, so cosa/sina equal the host's bit for
// bit -- and the per-pair kernels stage the prepared boxes of one frame in shared memory with a
// TMA bulk copy (cp.async.bulk + mbarrier) and evaluate the test with explicitly un-fused
// float32 operations.
#include <algorithm>

#include "hv_common.cuh"
#include "pcfe_common.cuh"

namespace pcfe {
Knob g_opt_pib_grid{1};  // 0: brute-force first-hit assignment for every frame (test knob)
namespace {

#include "pib_dev.cuh"

__global__ void debug_sincosf_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ s,
                                     float* __restrict__ c) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    s[i] = glibc_sin_or_cos(x[i], 0);
    c[i] = glibc_sin_or_cos(x[i], 1);
  }
}

// uniform xy grid over a frame's boxes (first-hit assignment, see pib_grid_build_kernel)
constexpr int kGridN = 64, kGridCells = kGridN * kGridN, kGridMaxSpan = 8, kGridMaxBoxes = 4096;
constexpr int kGridMaxList = 128;  // boxes per cell: longer lists (an outlier box stretching the grid) -> brute force
constexpr int kGridThreads = 1024;

struct __align__(16) GridHdr {
  float x0, y0, sx, sy;  // origin, cells per unit length
  int32_t ok, pad0, pad1, pad2;
};
static_assert(sizeof(GridHdr) == 32, "GridHdr must be 32 bytes");


// pcdet == 0: mmdet3d boxes (cx, cy, cz_bottom, w, l, h, rz), points_in_boxes_cpu.cpp:16-40.
// pcdet != 0: OpenPCDet boxes (cx, cy, cz_CENTRE, dx, dy, dz, heading) and test,
//   thirdparty/Spconv-OpenPCDet/pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:121-140
//   (the CUDA twin roiaware_pool3d_kernel.cu:16-37 differs in MARGIN only): rotation by -heading,
//   |z - cz| > dz / 2.0 rejects, inside <=> |lx| < dx / 2.0 + MARGIN and |ly| < dy / 2.0 + MARGIN
//   with the right-hand sides formed in double from the float MARGIN.
__global__ void pib_prepare_kernel(const float* __restrict__ boxes, int64_t nboxes,
                                   PBox* __restrict__ out, RBox* __restrict__ rout, const int pcdet,
                                   const float margin) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nboxes) return;
  const float* b = boxes + i * 7;
  const float cx = b[0], cy = b[1], cz = b[2], w = b[3], l = b[4], h = b[5], rz = b[6];
  PBox p;
  p.cx = cx;
  p.cy = cy;
  if (pcdet) {
    p.czc = cz;
    // float a, double T:  a > T <=> a > rd(T);  a < T <=> a < ru(T)  (no float lies strictly
    // between T and its directed roundings; NaN / Inf propagate)
    p.hh = __double2float_rd(__dmul_rn((double)h, 0.5));
    p.hl = __double2float_ru(__dadd_rn(__dmul_rn((double)w, 0.5), (double)margin));  // local x <-> dx = b[3]
    p.hw = __double2float_ru(__dadd_rn(__dmul_rn((double)l, 0.5), (double)margin));  // local y <-> dy = b[4]
    const float rot = -rz;
    p.cosa = glibc_sin_or_cos(rot, 1);
    p.sina = glibc_sin_or_cos(rot, 0);
    out[i] = p;
    RBox r;
    r.cx = cx;
    r.cy = cy;
    const double hl = (double)p.hl, hw = (double)p.hw;
    r.r = __double2float_ru(__dmul_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(hl, hl), __dmul_rn(hw, hw))), 1.0001));
    r.pad = 0.0f;
    rout[i] = r;
    return;
  }
  const PBox q = make_pbox_mmdet(b);
  out[i] = q;
  rout[i] = make_rbox(q);
}

// ------------------------------------------------------------------------------------------
// TMA bulk copy of the prepared boxes of one frame into shared memory
// ------------------------------------------------------------------------------------------

// Loads `count` prepared boxes (32 B each) and their reject records (16 B each) into shared
// memory; 16-byte aligned sources.  All threads call.
__device__ __forceinline__ void stage_boxes(PBox* sboxes, RBox* srej, const PBox* gboxes,
                                            const RBox* grej, int count, uint64_t* bar, uint32_t parity) {
  if (threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)count * (uint32_t)sizeof(PBox);
    const uint32_t rbytes = (uint32_t)count * (uint32_t)sizeof(RBox);
    mbar_expect_tx(bar, bytes + rbytes);
    bulk_g2s(sboxes, gboxes, bytes, bar);
    bulk_g2s(srej, grej, rbytes, bar);
  }
  mbar_wait(bar, parity);
}

constexpr int kBoxChunk = 1024;   // boxes per frame the point-major kernel keeps in shared memory
constexpr int kPointChunk = 512;  // boxes staged per pass by the thread-per-point kernels (24 KB static)
constexpr int kPibThreads = 256;

// ------------------------------------------------------------------------------------------
// points_in_boxes_batch: out (b, m, t), point-major, t <= kBoxChunk.
// Thread = one point against all t boxes (boxes broadcast from shared memory, staged by TMA): the
// conservative xy reject keeps the exact test off 99.5 % of the pairs, the flags are collected as
// bits (one word per 32 boxes) and staged in shared memory; then every warp writes the t int32
// flags of each of its 32 points as contiguous 128-byte (or 512-byte, VEC4) streaming stores.
// dynamic shared memory: PBox[t] | RBox[t] | masks[ceil(t / 32)][256]
// ------------------------------------------------------------------------------------------
template <bool VEC4>
__global__ void __launch_bounds__(kPibThreads)
pib_all_kernel(const PBox* __restrict__ pboxes, const RBox* __restrict__ rboxes,
               const float* __restrict__ points, const int t, const long long m,
               int32_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  PBox* sboxes = reinterpret_cast<PBox*>(smem_raw);
  RBox* srej = reinterpret_cast<RBox*>(smem_raw + (size_t)t * sizeof(PBox));
  uint32_t* masks = reinterpret_cast<uint32_t*>(smem_raw + (size_t)t * (sizeof(PBox) + sizeof(RBox)));

  const int b = blockIdx.y, tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * kPibThreads;
  const long long p = m0 + tid;
  float x = 0.f, y = 0.f, z = 0.f;
  if (p < m) {
    const float* gp = points + ((size_t)b * m + p) * 3;
    x = __ldg(gp); y = __ldg(gp + 1); z = __ldg(gp + 2);
  }
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  stage_boxes(sboxes, srej, pboxes + (size_t)b * t, rboxes + (size_t)b * t, t, &bar, 0);

  const int nw = (t + 31) >> 5;
  for (int j = 0; j < nw; ++j) {
    uint32_t word = 0;
    const int kend = min(32, t - j * 32);
#pragma unroll 8
    for (int bit = 0; bit < kend; ++bit) {
      const int k = j * 32 + bit;
      if (!xy_reject(x, y, srej[k])) word |= (uint32_t)in_box(x, y, z, sboxes[k]) << bit;
    }
    masks[j * kPibThreads + tid] = word;
  }
  __syncwarp();  // a warp writes the points it tested itself

  const int lane = tid & 31, wbase = tid & ~31;
  for (int pl = 0; pl < 32; ++pl) {
    const long long pp = m0 + wbase + pl;
    if (pp >= m) break;  // warp-uniform
    int32_t* __restrict__ row = out + ((size_t)b * m + pp) * t;
    const uint32_t* mrow = masks + wbase + pl;
    if (VEC4) {  // t % 4 == 0 and out is 16-byte aligned: lane writes flags [4 i, 4 i + 4)
      for (int i = lane; i * 4 < t; i += 32) {
        const uint32_t wv = mrow[(i >> 3) * kPibThreads] >> ((i & 7) * 4);
        __stcs(reinterpret_cast<int4*>(row) + i, make_int4(wv & 1u, (wv >> 1) & 1u, (wv >> 2) & 1u, (wv >> 3) & 1u));
      }
    } else {
      for (int k0 = 0; k0 < t; k0 += 32) {
        const uint32_t wv = mrow[(k0 >> 5) * kPibThreads];
        if (k0 + lane < t) __stcs(row + k0 + lane, (int32_t)((wv >> lane) & 1u));
      }
    }
  }
}

// Generic fallback for very large T (boxes do not fit one staging pass): thread per point,
// boxes streamed through shared memory in chunks, strided stores.
__global__ void __launch_bounds__(256)
pib_all_generic_kernel(const PBox* __restrict__ pboxes, const float* __restrict__ points, int t,
                       long long m, int32_t* __restrict__ out) {
  __shared__ __align__(16) PBox sboxes[128];
  const int b = blockIdx.y;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float x = 0.f, y = 0.f, z = 0.f;
  if (p < m) {
    const float* gp = points + ((size_t)b * m + p) * 3;
    x = __ldg(gp); y = __ldg(gp + 1); z = __ldg(gp + 2);
  }
  constexpr int chunk = 128;
  for (int t0 = 0; t0 < t; t0 += chunk) {
    const int cnt = min(chunk, t - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) sboxes[i] = pboxes[(size_t)b * t + t0 + i];
    __syncthreads();
    if (p < m) {
      int32_t* o = out + ((size_t)b * m + p) * t + t0;
      for (int k = 0; k < cnt; ++k) o[k] = in_box(x, y, z, sboxes[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// points_in_boxes_gpu: out (b, m) = lowest containing box index or -1
// points_in_boxes_cpu layout: out (t, n) box-major 0/1   (BOXMAJOR = true, b == 1)
// Thread = one point; boxes broadcast from shared memory, staged by TMA in chunks; conservative
// xy reject in front of the exact test.
// ------------------------------------------------------------------------------------------
template <bool BOXMAJOR>
__global__ void __launch_bounds__(256)
pib_point_kernel(const PBox* __restrict__ pboxes, const RBox* __restrict__ rboxes,
                 const float* __restrict__ points, int t, long long m, int32_t* __restrict__ out) {
  __shared__ __align__(16) PBox sboxes[kPointChunk];
  __shared__ __align__(16) RBox srej[kPointChunk];
  __shared__ __align__(8) uint64_t bar;
  const int b = blockIdx.y;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float x = 0.f, y = 0.f, z = 0.f;
  if (p < m) {
    const float* gp = points + ((size_t)b * m + p) * 3;
    x = __ldg(gp); y = __ldg(gp + 1); z = __ldg(gp + 2);
  }
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  int first = -1;
  uint32_t parity = 0;
  // box-major layout: gridDim.z slices of the boxes (a single frame of points alone does not fill
  // the GPU); the first-hit layout needs all boxes in order and runs with gridDim.z == 1
  const int per_z = BOXMAJOR ? (t + (int)gridDim.z - 1) / (int)gridDim.z : t;
  const int t_begin = BOXMAJOR ? (int)blockIdx.z * per_z : 0, t_end = min(t, t_begin + per_z);
  for (int t0 = t_begin; t0 < t_end; t0 += kPointChunk) {
    const int cnt = min(kPointChunk, t_end - t0);
    if (t0 > t_begin) __syncthreads();  // everyone is done with the previous chunk
    stage_boxes(sboxes, srej, pboxes + (size_t)b * t + t0, rboxes + (size_t)b * t + t0, cnt, &bar, parity);
    parity ^= 1u;
    if (p < m) {
      if (BOXMAJOR) {
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
          int32_t r = 0;
          if (!xy_reject(x, y, srej[k])) r = in_box(x, y, z, sboxes[k]);
          __stcs(out + (size_t)(t0 + k) * m + p, r);
        }
      } else if (first < 0) {
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
          if (!xy_reject(x, y, srej[k]) && in_box(x, y, z, sboxes[k])) {
            first = t0 + k;  // points_in_boxes_cuda.cu:71-75: first hit wins
            break;
          }
        }
      }
    }
  }
  if (!BOXMAJOR && p < m) out[(size_t)b * m + p] = first;
}

// ------------------------------------------------------------------------------------------
// First-hit assignment through a uniform xy grid over the frame's boxes.
// The brute-force kernel above spends 6 instructions on each of the m * t pairs just to reject
// them.  Here every frame gets a 64 x 64 grid over the bounding square of its boxes' reject
// circles; a cell lists, in ascending box index, the boxes whose circle (plus a slack that
// dwarfs float rounding) touches it, and a point runs the exact test only on the list of its own
// cell.  cell(x) = clamp(floor((x - x0) * sx)) is monotone, a box is listed in every cell of
// [cell(cx - r - slack), cell(cx + r + slack)]^2, and a point inside a box has |x - cx| <= r and
// |y - cy| <= r (RBox), so the list of its cell contains every box it can be inside of: the result
// -- lowest containing box index -- is the brute-force one.  Frames the grid cannot represent
// (non-finite boxes, a box spanning more than kGridMaxSpan cells per axis, a cell listing more than
// kGridMaxList boxes, degenerate extent, t > kGridMaxBoxes) are flagged and taken by the brute-force kernel.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int grid_cell(float v, float v0, float s) {
  const float q = floorf(__fmul_rn(__fsub_rn(v, v0), s));
  // NaN compares false twice and lands in cell 0 (the exact test rejects a NaN point anyway)
  return q >= (float)(kGridN - 1) ? kGridN - 1 : (q > 0.0f ? (int)q : 0);
}

__device__ __forceinline__ float grid_slack(float c, float r) {
  return __fadd_rn(1e-2f, __fmul_rn(1e-6f, __fadd_rn(fabsf(c), r)));
}

// one CTA per frame
__global__ void __launch_bounds__(kGridThreads)
pib_grid_build_kernel(const RBox* __restrict__ rboxes, const int t, GridHdr* __restrict__ hdrs,
                      uint32_t* __restrict__ starts_base, uint16_t* __restrict__ entries_base) {
  __shared__ uint32_t cnt[kGridCells];
  __shared__ uint32_t warp_sums[33];
  __shared__ float red[4][kGridThreads / 32];
  __shared__ int s_bad;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const RBox* rb = rboxes + (size_t)b * t;
  uint32_t* starts = starts_base + (size_t)b * (kGridCells + 1);
  uint16_t* entries = entries_base + (size_t)b * t * (kGridMaxSpan * kGridMaxSpan);
  GridHdr* hdr = hdrs + b;
  if (tid == 0) s_bad = t > kGridMaxBoxes ? 1 : 0;
  for (int i = tid; i < kGridCells; i += kGridThreads) cnt[i] = 0u;
  // bounds of the reject circles
  float xlo = INFINITY, ylo = INFINITY, xhi = -INFINITY, yhi = -INFINITY;
  bool bad = false;
  for (int k = tid; k < t; k += kGridThreads) {
    const RBox r = rb[k];
    const float sl = grid_slack(fmaxf(fabsf(r.cx), fabsf(r.cy)), r.r);
    const float e = __fadd_rn(r.r, sl);
    bad |= !(isfinite(r.cx) && isfinite(r.cy) && isfinite(e));
    xlo = fminf(xlo, __fsub_rn(r.cx, e)); xhi = fmaxf(xhi, __fadd_rn(r.cx, e));
    ylo = fminf(ylo, __fsub_rn(r.cy, e)); yhi = fmaxf(yhi, __fadd_rn(r.cy, e));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    xlo = fminf(xlo, __shfl_xor_sync(0xFFFFFFFFu, xlo, d)); xhi = fmaxf(xhi, __shfl_xor_sync(0xFFFFFFFFu, xhi, d));
    ylo = fminf(ylo, __shfl_xor_sync(0xFFFFFFFFu, ylo, d)); yhi = fmaxf(yhi, __shfl_xor_sync(0xFFFFFFFFu, yhi, d));
  }
  if (lane == 0) { red[0][wid] = xlo; red[1][wid] = xhi; red[2][wid] = ylo; red[3][wid] = yhi; }
  __syncthreads();
  if (bad) atomicOr(&s_bad, 1);
  xlo = red[0][0]; xhi = red[1][0]; ylo = red[2][0]; yhi = red[3][0];
  for (int k = 1; k < kGridThreads / 32; ++k) {
    xlo = fminf(xlo, red[0][k]); xhi = fmaxf(xhi, red[1][k]); ylo = fminf(ylo, red[2][k]); yhi = fmaxf(yhi, red[3][k]);
  }
  const float sx = __fdiv_rn((float)kGridN, __fsub_rn(xhi, xlo)), sy = __fdiv_rn((float)kGridN, __fsub_rn(yhi, ylo));
  __syncthreads();
  if (tid == 0 && !(isfinite(sx) && isfinite(sy) && sx > 0.0f && sy > 0.0f)) s_bad = 1;
  __syncthreads();
  if (s_bad) {
    if (tid == 0) *hdr = GridHdr{0.f, 0.f, 0.f, 0.f, 0, 0, 0, 0};
    return;
  }
  // pass 1: cells per box, counts per cell
  for (int k = tid; k < t; k += kGridThreads) {
    const RBox r = rb[k];
    const float e = __fadd_rn(r.r, grid_slack(fmaxf(fabsf(r.cx), fabsf(r.cy)), r.r));
    const int cx0 = grid_cell(__fsub_rn(r.cx, e), xlo, sx), cx1 = grid_cell(__fadd_rn(r.cx, e), xlo, sx);
    const int cy0 = grid_cell(__fsub_rn(r.cy, e), ylo, sy), cy1 = grid_cell(__fadd_rn(r.cy, e), ylo, sy);
    if (cx1 - cx0 >= kGridMaxSpan || cy1 - cy0 >= kGridMaxSpan) { atomicOr(&s_bad, 1); continue; }
    for (int cy = cy0; cy <= cy1; ++cy)
      for (int cx = cx0; cx <= cx1; ++cx) atomicAdd(&cnt[cy * kGridN + cx], 1u);
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0) *hdr = GridHdr{0.f, 0.f, 0.f, 0.f, 0, 0, 0, 0};
    return;
  }
  // exclusive scan of the counts: thread i owns cells [16 i, 16 i + 16)
  constexpr int kPerThread = kGridCells / kGridThreads;
  uint32_t local[kPerThread], sum = 0;
  uint32_t longest = 0;
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {
    local[j] = cnt[tid * kPerThread + j];
    sum += local[j];
    longest = max(longest, local[j]);
  }
  if (longest > (uint32_t)kGridMaxList) atomicOr(&s_bad, 1);
  uint32_t total;
  uint32_t run = block_exscan(sum, warp_sums, &total);  // (contains the barriers that publish s_bad)
  __syncthreads();
  if (s_bad) {
    if (tid == 0) *hdr = GridHdr{0.f, 0.f, 0.f, 0.f, 0, 0, 0, 0};
    return;
  }
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {
    starts[tid * kPerThread + j] = run;
    cnt[tid * kPerThread + j] = run;  // becomes the fill cursor
    run += local[j];
  }
  if (tid == kGridThreads - 1) starts[kGridCells] = total;
  __syncthreads();
  // pass 2: fill (unordered inside a cell)
  for (int k = tid; k < t; k += kGridThreads) {
    const RBox r = rb[k];
    const float e = __fadd_rn(r.r, grid_slack(fmaxf(fabsf(r.cx), fabsf(r.cy)), r.r));
    const int cx0 = grid_cell(__fsub_rn(r.cx, e), xlo, sx), cx1 = grid_cell(__fadd_rn(r.cx, e), xlo, sx);
    const int cy0 = grid_cell(__fsub_rn(r.cy, e), ylo, sy), cy1 = grid_cell(__fadd_rn(r.cy, e), ylo, sy);
    for (int cy = cy0; cy <= cy1; ++cy)
      for (int cx = cx0; cx <= cx1; ++cx) entries[atomicAdd(&cnt[cy * kGridN + cx], 1u)] = (uint16_t)k;
  }
  __syncthreads();
  // pass 3: ascending box index inside every cell (first hit = lowest index); the owner of a cell
  // sorts the list it finds in global memory (written by this CTA: visible after the barrier)
#pragma unroll 1
  for (int j = 0; j < kPerThread; ++j) {
    const int cell = tid * kPerThread + j;
    const uint32_t lo = starts[cell], hi = cnt[cell];
    for (uint32_t i = lo + 1; i < hi; ++i) {
      const uint16_t v = entries[i];
      uint32_t q = i;
      while (q > lo && entries[q - 1] > v) { entries[q] = entries[q - 1]; --q; }
      entries[q] = v;
    }
  }
  if (tid == 0) *hdr = GridHdr{xlo, ylo, sx, sy, 1, 0, 0, 0};
}

// thread = two points (256 apart) of a frame: their grid look-ups are issued together, which doubles
// the loads in flight of a kernel that is all dependent loads (cell -> list -> box).  A frame without
// a grid is tested against all boxes straight from global memory (every lane reads the same box:
// uniform, L1-resident loads).
constexpr int kGridPointsPerThread = 2;

__global__ void __launch_bounds__(256)
pib_point_grid_kernel(const PBox* __restrict__ pboxes, const RBox* __restrict__ rboxes,
                      const GridHdr* __restrict__ hdrs, const uint32_t* __restrict__ starts_base,
                      const uint16_t* __restrict__ entries_base, const float* __restrict__ points, const int t,
                      const long long m, int32_t* __restrict__ out) {
  const int b = blockIdx.y;
  const GridHdr h = hdrs[b];
  const long long p0 = (long long)blockIdx.x * (256 * kGridPointsPerThread) + threadIdx.x;
  const float4* pb4 = reinterpret_cast<const float4*>(pboxes + (size_t)b * t);
  float x[kGridPointsPerThread], y[kGridPointsPerThread], z[kGridPointsPerThread];
#pragma unroll
  for (int j = 0; j < kGridPointsPerThread; ++j) {
    const long long p = p0 + 256 * j;
    x[j] = y[j] = z[j] = 0.f;
    if (p < m) {
      const float* gp = points + ((size_t)b * m + p) * 3;
      x[j] = __ldg(gp); y[j] = __ldg(gp + 1); z[j] = __ldg(gp + 2);
    }
  }
  if (!h.ok) {  // brute force, first hit wins
    const RBox* rb = rboxes + (size_t)b * t;
#pragma unroll
    for (int j = 0; j < kGridPointsPerThread; ++j) {
      const long long p = p0 + 256 * j;
      if (p >= m) continue;
      int first = -1;
      for (int k = 0; k < t; ++k) {
        const float4 rr = __ldg(reinterpret_cast<const float4*>(rb + k));
        if (xy_reject(x[j], y[j], RBox{rr.x, rr.y, rr.z, rr.w})) continue;
        const float4 a = __ldg(pb4 + 2 * k), c = __ldg(pb4 + 2 * k + 1);
        const PBox bx{a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
        if (in_box(x[j], y[j], z[j], bx)) {
          first = k;
          break;
        }
      }
      out[(size_t)b * m + p] = first;
    }
    return;
  }
  const uint32_t* starts = starts_base + (size_t)b * (kGridCells + 1);
  const uint16_t* entries = entries_base + (size_t)b * t * (kGridMaxSpan * kGridMaxSpan);
  uint32_t lo[kGridPointsPerThread], hi[kGridPointsPerThread];
#pragma unroll
  for (int j = 0; j < kGridPointsPerThread; ++j) {
    const int cell = grid_cell(y[j], h.y0, h.sy) * kGridN + grid_cell(x[j], h.x0, h.sx);
    lo[j] = __ldg(starts + cell);
    hi[j] = __ldg(starts + cell + 1);
  }
#pragma unroll
  for (int j = 0; j < kGridPointsPerThread; ++j) {
    const long long p = p0 + 256 * j;
    if (p >= m) continue;
    int first = -1;
    for (uint32_t i = lo[j]; i < hi[j]; ++i) {
      const int k = (int)__ldg(entries + i);
      const float4 a = __ldg(pb4 + 2 * k), c = __ldg(pb4 + 2 * k + 1);
      const PBox bx{a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
      if (in_box(x[j], y[j], z[j], bx)) {
        first = k;
        break;
      }
    }
    out[(size_t)b * m + p] = first;
  }
}

__global__ void fill_kernel(int32_t* out, long long n, int32_t v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Returns PCFE_OK to proceed, 1 when there is nothing to compute, < 0 on argument errors.
// `per_point_out` is the number of output elements per point (t for the mask layouts, 1 for the
// first-hit layout): a call whose output is empty is a no-op whatever the pointers are.
int check_common(const float* boxes, const float* points, const int32_t* out, int b, int t,
                 int64_t m, int64_t per_point_out, void* ws, size_t ws_bytes) {
  if (b < 0 || t < 0 || m < 0) return PCFE_ERR_SHAPE;
  if (b > 65535) return PCFE_ERR_TOO_LARGE;
  if ((int64_t)b * (int64_t)t >= (1ll << 31) || m >= (1ll << 40)) return PCFE_ERR_TOO_LARGE;
  if (b == 0 || m == 0 || per_point_out == 0) return 1;  // nothing to write
  if (!points || !out) return PCFE_ERR_NULL;
  if (t > 0 && (!boxes || !ws)) return PCFE_ERR_NULL;
  if (((uintptr_t)boxes & 3) || ((uintptr_t)points & 3) || ((uintptr_t)out & 3)) return PCFE_ERR_ALIGN;
  if (t > 0 && ((uintptr_t)ws & 255)) return PCFE_ERR_ALIGN;
  if (t > 0 && ws_bytes < pcfe_points_in_boxes_workspace_bytes(b, t)) return PCFE_ERR_WORKSPACE;
  return PCFE_OK;
}

// workspace layout: PBox[nboxes] | (256-byte aligned) RBox[nboxes] | GridHdr[b] | starts[b][4097] |
// entries[b][t * 64]   (the grid part is used by the first-hit entries only)
inline RBox* rbox_base(void* ws, int64_t nboxes) {
  return reinterpret_cast<RBox*>((char*)ws + align256((size_t)nboxes * sizeof(PBox)));
}
inline size_t grid_off(int b, int t) {
  return align256((size_t)b * t * sizeof(PBox)) + align256((size_t)b * t * sizeof(RBox));
}
inline size_t grid_hdr_bytes(int b) { return align256((size_t)b * sizeof(GridHdr)); }
inline size_t grid_starts_bytes(int b) { return align256((size_t)b * (kGridCells + 1) * sizeof(uint32_t)); }
inline size_t grid_entries_bytes(int b, int t) {
  if (t > kGridMaxBoxes) return 0;  // such frames never get a grid
  return align256((size_t)b * t * (kGridMaxSpan * kGridMaxSpan) * sizeof(uint16_t));
}

int prepare(const float* boxes, int64_t nboxes, void* ws, cudaStream_t st, int pcdet = 0, float margin = 0.0f) {
  pib_prepare_kernel<<<(unsigned)((nboxes + 127) / 128), 128, 0, st>>>(boxes, nboxes, (PBox*)ws, rbox_base(ws, nboxes),
                                                                       pcdet, margin);
  PCFE_LAUNCH_CHECK();
  return PCFE_OK;
}

}  // namespace
}  // namespace pcfe

using namespace pcfe;

extern "C" size_t pcfe_points_in_boxes_workspace_bytes(int b, int t) {
  if (b <= 0 || t <= 0) return 256;
  return grid_off(b, t) + grid_hdr_bytes(b) + grid_starts_bytes(b) + grid_entries_bytes(b, t);
}

static int part_impl(const float* boxes, const float* points, int b, int t, int64_t m, int32_t* out, void* ws,
                     size_t ws_bytes, int device, void* stream, int pcdet, float margin) {
  int rc = check_common(boxes, points, out, b, t, m, 1, ws, ws_bytes);
  if (rc != PCFE_OK) return rc > 0 ? PCFE_OK : rc;
  DeviceGuard guard(device);
  PCFE_CUDA_TRY(guard.err);
  cudaStream_t st = (cudaStream_t)stream;
  if (t == 0) {  // no boxes: everything is background
    const long long total = (long long)b * m;
    fill_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(out, total, -1);
    PCFE_LAUNCH_CHECK();
    return PCFE_OK;
  }
  PBox* pb = (PBox*)ws;
  if ((rc = prepare(boxes, (int64_t)b * t, ws, st, pcdet, margin)) != PCFE_OK) return rc;
  const RBox* rb = rbox_base(ws, (int64_t)b * t);
  dim3 grid((unsigned)((m + 255) / 256), (unsigned)b);
  if (g_opt_pib_grid) {
    // frames whose boxes fit a 64 x 64 grid are assigned through it, the others by brute force
    char* gbase = (char*)ws + grid_off(b, t);
    GridHdr* hdrs = (GridHdr*)gbase;
    uint32_t* starts = (uint32_t*)(gbase + grid_hdr_bytes(b));
    uint16_t* entries = (uint16_t*)(gbase + grid_hdr_bytes(b) + grid_starts_bytes(b));
    pib_grid_build_kernel<<<b, kGridThreads, 0, st>>>(rb, t, hdrs, starts, entries);
    PCFE_LAUNCH_CHECK();
    const int per_cta = 256 * kGridPointsPerThread;
    dim3 ggrid((unsigned)((m + per_cta - 1) / per_cta), (unsigned)b);
    pib_point_grid_kernel<<<ggrid, 256, 0, st>>>(pb, rb, hdrs, starts, entries, points, t, (long long)m, out);
    PCFE_LAUNCH_CHECK();
    return PCFE_OK;
  }
  pib_point_kernel<false><<<grid, 256, 0, st>>>(pb, rb, points, t, (long long)m, out);
  PCFE_LAUNCH_CHECK();
  return PCFE_OK;
}

extern "C" int pcfe_points_in_boxes_part_f32(const float* boxes, const float* points, int b, int t,
                                             int64_t m, int32_t* out, void* ws, size_t ws_bytes,
                                             int device, void* stream) {
  return part_impl(boxes, points, b, t, m, out, ws, ws_bytes, device, stream, 0, 0.0f);
}

extern "C" int pcfe_pcdet_points_in_boxes_gpu_f32(const float* boxes, const float* points, int b, int t,
                                                  int64_t m, int32_t* out, void* ws, size_t ws_bytes,
                                                  int device, void* stream) {
  return part_impl(boxes, points, b, t, m, out, ws, ws_bytes, device, stream, 1, 1e-5f);  // roiaware_pool3d_kernel.cu:27
}

static int boxmajor_impl(const float* boxes, const float* points, int t, int64_t n, int32_t* out, void* ws,
                         size_t ws_bytes, int device, void* stream, int pcdet, float margin) {
  int rc = check_common(boxes, points, out, 1, t, n, t, ws, ws_bytes);
  if (rc != PCFE_OK) return rc > 0 ? PCFE_OK : rc;
  if (t == 0) return PCFE_OK;
  DeviceGuard guard(device);
  PCFE_CUDA_TRY(guard.err);
  cudaStream_t st = (cudaStream_t)stream;
  PBox* pb = (PBox*)ws;
  if ((rc = prepare(boxes, t, ws, st, pcdet, margin)) != PCFE_OK) return rc;
  // enough CTAs for ~8 per SM: slices of at least 32 boxes
  const long long tiles = (n + 255) / 256;
  const int zs = (int)std::max<long long>(1, std::min<long long>((t + 31) / 32, (148 * 8 + tiles - 1) / tiles));
  dim3 grid((unsigned)tiles, 1, (unsigned)zs);
  pib_point_kernel<true><<<grid, 256, 0, st>>>(pb, rbox_base(ws, t), points, t, (long long)n, out);
  PCFE_LAUNCH_CHECK();
  return PCFE_OK;
}

extern "C" int pcfe_points_in_boxes_boxmajor_f32(const float* boxes, const float* points, int t,
                                                 int64_t n, int32_t* out, void* ws, size_t ws_bytes,
                                                 int device, void* stream) {
  return boxmajor_impl(boxes, points, t, n, out, ws, ws_bytes, device, stream, 0, 0.0f);
}

extern "C" int pcfe_pcdet_points_in_boxes_cpu_f32(const float* boxes, const float* points, int t,
                                                  int64_t n, int32_t* out, void* ws, size_t ws_bytes,
                                                  int device, void* stream) {
  return boxmajor_impl(boxes, points, t, n, out, ws, ws_bytes, device, stream, 1, 1e-2f);  // roiaware_pool3d.cpp:131
}

extern "C" int pcfe_points_in_boxes_all_f32(const float* boxes, const float* points, int b, int t,
                                            int64_t m, int32_t* out, void* ws, size_t ws_bytes,
                                            int device, void* stream) {
  int rc = check_common(boxes, points, out, b, t, m, t, ws, ws_bytes);
  if (rc != PCFE_OK) return rc > 0 ? PCFE_OK : rc;
  DeviceGuard guard(device);
  PCFE_CUDA_TRY(guard.err);
  cudaStream_t st = (cudaStream_t)stream;
  PBox* pb = (PBox*)ws;
  if ((rc = prepare(boxes, (int64_t)b * t, ws, st)) != PCFE_OK) return rc;

  if (t <= kBoxChunk) {  // all boxes of a frame in shared memory
    const bool vec4 = (t % 4 == 0) && (((uintptr_t)out & 15) == 0);
    const size_t smem = (size_t)t * (sizeof(PBox) + sizeof(RBox)) + (size_t)((t + 31) / 32) * kPibThreads * sizeof(uint32_t);
    dim3 grid((unsigned)((m + kPibThreads - 1) / kPibThreads), (unsigned)b);
    const RBox* rb = rbox_base(ws, (int64_t)b * t);
    if (vec4) {
      PCFE_CUDA_TRY(cudaFuncSetAttribute(pib_all_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      pib_all_kernel<true><<<grid, kPibThreads, smem, st>>>(pb, rb, points, t, (long long)m, out);
    } else {
      PCFE_CUDA_TRY(cudaFuncSetAttribute(pib_all_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      pib_all_kernel<false><<<grid, kPibThreads, smem, st>>>(pb, rb, points, t, (long long)m, out);
    }
    PCFE_LAUNCH_CHECK();
  } else {
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)b);
    pib_all_generic_kernel<<<grid, 256, 0, st>>>(pb, points, t, (long long)m, out);
    PCFE_LAUNCH_CHECK();
  }
  return PCFE_OK;
}

extern "C" int pcfe_debug_sincosf(const float* x, int64_t n, float* s, float* c, int device,
                                  void* stream) {
  if (n < 0) return PCFE_ERR_SHAPE;
  if (n == 0) return PCFE_OK;
  if (!x || !s || !c) return PCFE_ERR_NULL;
  DeviceGuard guard(device);
  PCFE_CUDA_TRY(guard.err);
  debug_sincosf_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, s, c);
  PCFE_LAUNCH_CHECK();
  return PCFE_OK;
}

extern "C" int pcfe_version(void) { return PCFE_VERSION; }

extern "C" uint64_t pcfe_launch_count(void) { return g_launches.load(); }

extern "C" const char* pcfe_error_string(int code) {
  switch (code) {
    case PCFE_OK: return "ok";
    case PCFE_ERR_NULL: return "pcfe: required pointer is NULL";
    case PCFE_ERR_SHAPE: return "pcfe: bad shape (negative size, c < 3, or wrong inner dimension)";
    case PCFE_ERR_GRID: return "pcfe: voxel grid is empty or has >= 2^32-1 cells";
    case PCFE_ERR_WORKSPACE: return "pcfe: workspace too small";
    case PCFE_ERR_ALIGN: return "pcfe: pointer is not sufficiently aligned";
    case PCFE_ERR_CAPS: return "pcfe: max_points / max_voxels must be >= 0";
    case PCFE_ERR_TOO_LARGE: return "pcfe: input too large";
    case PCFE_ERR_DEVICE: return "pcfe: no usable CUDA device";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "pcfe: unknown error";
}
