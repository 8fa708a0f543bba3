// detmatch_b200/csrc/hv_cluster.cuh -- hard voxelization, record path: partition + grouping + voxel
// numbering of a frame by ONE thread-block cluster (included by hv_bucket.cu inside its namespaces).
//
// Behaviour reproduced bit for bit: mmdet3d/ops/voxel/src/voxelization_cpu.cpp:43-99.
//
// Replaces hvb_zero + hvb_bin + hvb_bucket_rec + hvb_scan_firsts for frames of up to 196 608 points
// (P == 5, C = 4 / 5).  A cluster of 16 CTAs (1024 threads each) owns a frame:
//   P1  every CTA reads 1/16 of the rows, computes the cell keys and routes (key, point) entries to
//       the CTA that owns the key's hash class -- 32 classes, two per CTA -- straight into that CTA's
//       shared memory (st.shared::cluster).  The exact position of every entry is known before the
//       first store: ranks inside (warp, class) come from warp-private shared-memory counters, the
//       per-class totals of all 16 CTAs are all-gathered through DSMEM (16 x 32 words), so the
//       inboxes are densely packed and nothing is reserved with remote atomics.
//   P2  per class: open-addressing table in shared memory (one CAS per probe, one atomicAdd for the
//       arrival rank inside the cell), ONE block scan over the table slots that yields the cell list
//       and the segment offsets, a scatter of the point indices into cell-contiguous order (into the
//       dead key words of the inbox) and a dense thread-per-cell pass: <= 5 entries -> 9-comparator
//       network, more -> insertion of the rest.  Output as in hvb_bucket_rec: rec[first] for cells
//       with more than one point, one 64-bit atomicOr on the frame's {first, has-more} mask.
//   P3  the mask is compacted into firsts[voxel id] (slice per CTA, totals exchanged through DSMEM)
//       and voxel_num is written: no separate zero / scan launches.
// The partition entries never leave the chip: no `ent` array, no L2 atomics per (tile, bucket).
// Frames whose classes do not fit (heavy duplication of one hash class: > 6912 entries in a class or
// > 12288 for a CTA) are flagged for the single-CTA fallback exactly like bucket overflows.
#pragma once

constexpr int kCS = 16;                         // CTAs per cluster (non-portable size: opt-in attribute)
constexpr int kCT = 1024;                       // threads per CTA
constexpr int kCWarps = kCT / 32;               // 32
constexpr int kCPP = 12;                        // points per thread, at most
constexpr int kCHalf = 6;                       // points whose row loads are in flight together
constexpr int kCV = 2;                          // hash classes per CTA
constexpr int kCClasses = kCS * kCV;            // 32 == warps per CTA == lanes (used below)
constexpr int kCCap = 12288;                    // inbox entries per CTA
constexpr int kCLog2Slots = 13;
constexpr int kCSlots = 1 << kCLog2Slots;       // table slots per class
constexpr int kCClassMax = 6912;                // entries per class (table load <= 0.85)
constexpr int kCIns = (kCClassMax + kCT - 1) / kCT;  // 7 entries per thread and class
constexpr int kCMaxPoints = kCS * kCT * kCPP;   // 196 608
constexpr int kCPad = 33;                       // row pitch of the (warp, class) matrices: conflict-free both ways
static_assert(kCClasses == 32 && kCWarps == 32, "warp w handles class w; lane l handles warp l");

struct HvcSmem {
  uint2 inbox[kCCap];              // {key, point}; .x is reused for the cell-ordered point indices
  uint32_t hkey[kCSlots];          // table keys    (P3: stage of the compacted first points, with hv)
  uint32_t hv[kCSlots];            // count, then (segment offset << 16) | count
  uint16_t celllist[kCClassMax];   // occupied slots in slot order
  uint32_t wcnt[kCWarps * kCPad];  // [warp][class] counts, then exclusive prefix over the warps
  uint32_t comb[kCWarps * kCPad];  // [warp][class] remote shared address of the (warp, class) run
  uint32_t cntmat[kCS * kCClasses];  // [source CTA][class] entries (all-gathered)
  uint32_t tot[kCS];               // P3: first points per mask slice (all-gathered)
  uint32_t warp_sums[33];
  uint32_t ne[kCV], coff[kCV];     // this CTA's classes: entries, offset in the inbox
  uint32_t over;                   // the frame takes the fallback
};

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_count() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ascending compare-exchange
#define PCFE_CE(a, b)                 \
  do {                                \
    const uint32_t _lo = min(a, b);   \
    b = max(a, b);                    \
    a = _lo;                          \
  } while (0)

template <int CT>
__global__ void __launch_bounds__(kCT, 1)
hvc_group_kernel(const __grid_constant__ HvBatch batch, const HvbWork w, const GridParams g, const int c_rt,
                 const int use_fast_div, const int max_voxels, int32_t* __restrict__ voxel_num,
                 const int frames) {
  extern __shared__ __align__(16) unsigned char hvc_smem_raw[];
  HvcSmem& sm = *reinterpret_cast<HvcSmem*>(hvc_smem_raw);
  const int c = CT > 0 ? CT : c_rt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_rank();
  const uint32_t inbox_saddr = smem_u32(sm.inbox);
  const uint32_t cntmat_saddr = smem_u32(sm.cntmat);
  const uint32_t tot_saddr = smem_u32(sm.tot);
  pdl_wait();  // the scratch written below may still be read by the previous launch sequence
  // every CTA of the cluster has started (its shared memory exists) before the first remote store
  cluster_arrive();
  cluster_wait();

#pragma unroll 1
  for (int f = (int)cluster_id(); f < frames; f += (int)cluster_count()) {
    const HvFrame& fr = batch.f[f];
    const int n = fr.n;
    const int npt = (n + kCS * kCT - 1) / (kCS * kCT);  // points per thread (host: <= kCPP)
    const int cta_base = (int)rank * npt * kCT;
    const int words = (n + 31) >> 5;                 // 64-bit mask words of the frame
    const int slice = (words + kCS - 1) / kCS;       // mask words per CTA (<= 384)
    unsigned long long* __restrict__ bm64 = reinterpret_cast<unsigned long long*>(w.bitmask(f));
    uint32_t* __restrict__ ctl = w.ctl(f);

    // ---- frame set-up: counters, this CTA's slice of the mask, L2 prefetch of the next frame ----
    for (int i = tid; i < kCWarps * kCPad; i += kCT) sm.wcnt[i] = 0u;
    if (tid < slice && (int)rank * slice + tid < words) bm64[(int)rank * slice + tid] = 0ull;
    if (tid == 0) sm.over = 0u;
    {
      const int fn = f + (int)cluster_count();
      if (fn < frames) {
        const HvFrame& nf = batch.f[fn];
        const int nnpt = (nf.n + kCS * kCT - 1) / (kCS * kCT);
        const size_t total = ((size_t)nf.n * c * 4) & ~(size_t)15;
        const size_t per = ((size_t)nnpt * c * 4 + 15) & ~(size_t)15;  // bytes per thread: the CTA's slice / 1024
        const size_t lo = ((size_t)rank * kCT + tid) * per;
        if (lo < total && ((uintptr_t)nf.pts & 15) == 0) {
          const uint32_t bytes = (uint32_t)min(per, total - lo);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(nf.pts) + lo), "r"(bytes) : "memory");
        }
      }
    }
    __syncthreads();

    // ---- P1a: rows -> cell keys; rank of every entry inside its (warp, class) run -----------------
    uint32_t key[kCPP];
    uint32_t rbp[kCPP / 2];  // two 16-bit ranks per word
#pragma unroll
    for (int k = 0; k < kCPP / 2; ++k) rbp[k] = 0u;
    const FastAxes fa = make_fast_axes(g);
#pragma unroll
    for (int h = 0; h < kCPP; h += kCHalf) {
      float ax[kCHalf], ay[kCHalf], az[kCHalf];
      const float qnan = 1e30f;  // past the end: fails every range test, but stays inside the fast division's guard
#pragma unroll
      for (int k = 0; k < kCHalf; ++k) {
        const int i = cta_base + (h + k) * kCT + tid;
        const bool in = (h + k) < npt && i < n;
        const float* __restrict__ p = fr.pts + (size_t)i * c;
        ax[k] = in ? __ldg(p) : qnan;
        ay[k] = in ? __ldg(p + 1) : qnan;
        az[k] = in ? __ldg(p + 2) : qnan;
      }
      bool guard_ok = use_fast_div != 0;
      uint32_t fmask = 0xFFFFFFFFu;  // bit k: point k passes the fused PointsRangeFilter
      if (g.filter) {
        fmask = 0u;
#pragma unroll
        for (int k = 0; k < kCHalf; ++k) fmask |= filter_pass(ax[k], ay[k], az[k], g) ? (1u << k) : 0u;
      }
#pragma unroll
      for (int k = 0; k < kCHalf; ++k) {
        ax[k] = __fsub_rn(ax[k], g.x0);
        ay[k] = __fsub_rn(ay[k], g.y0);
        az[k] = __fsub_rn(az[k], g.z0);
        guard_ok = guard_ok & fast_div_guard(ax[k]) & fast_div_guard(ay[k]) & fast_div_guard(az[k]);
      }
      if (guard_ok) {
#pragma unroll
        for (int k = 0; k < kCHalf; ++k) {
          const float qx = fast_div(ax[k], g.vx, fa.rx), qy = fast_div(ay[k], g.vy, fa.ry), qz = fast_div(az[k], g.vz, fa.rz);
          // 0 <= q < 2^31 <=> bits(q) < bits(2^31) as unsigned; -0 and NaN cannot come out of the guard
          const uint32_t qmax = max(max(__float_as_uint(qx), __float_as_uint(qy)), __float_as_uint(qz));
          const int cx = __float2int_rz(qx), cy = __float2int_rz(qy), cz = __float2int_rz(qz);
          const bool ok = (qmax < 0x4F000000u) & (cx < g.gx) & (cy < g.gy) & (cz < g.gz);
          const uint32_t lin = ((uint32_t)cz * (uint32_t)g.gy + (uint32_t)cy) * (uint32_t)g.gx + (uint32_t)cx;
          key[h + k] = (ok && ((fmask >> k) & 1u)) ? lin : kEmpty;
        }
      } else {
#pragma unroll
        for (int k = 0; k < kCHalf; ++k) {
          // voxelization_cpu.cpp:23-29 on the differences (axis_cell() without its subtract)
          const float qx = __fdiv_rn(ax[k], g.vx), qy = __fdiv_rn(ay[k], g.vy), qz = __fdiv_rn(az[k], g.vz);
          const bool in = (qx >= 0.0f) & (qx < 2147483648.0f) & (qy >= 0.0f) & (qy < 2147483648.0f) &
                          (qz >= 0.0f) & (qz < 2147483648.0f);
          const int cx = in ? __float2int_rz(qx) : -1, cy = in ? __float2int_rz(qy) : -1, cz = in ? __float2int_rz(qz) : -1;
          const bool ok = in & (cx < g.gx) & (cy < g.gy) & (cz < g.gz);
          const uint32_t lin = ((uint32_t)cz * (uint32_t)g.gy + (uint32_t)cy) * (uint32_t)g.gx + (uint32_t)cx;
          key[h + k] = (ok && ((fmask >> k) & 1u)) ? lin : kEmpty;
        }
      }
#pragma unroll
      for (int k = 0; k < kCHalf; ++k) {
        if (key[h + k] != kEmpty) {
          const uint32_t cls = (key[h + k] * kGold) >> 27;
          const uint32_t r = atomicAdd(&sm.wcnt[warp * kCPad + cls], 1u);  // < 12 * 32
          rbp[(h + k) >> 1] |= r << (((h + k) & 1) * 16);
        }
      }
    }
    __syncthreads();

    // ---- P1b: warp w = class w, lane l = warp l: prefix over the warps, class total to every CTA ----
    {
      const uint32_t v = sm.wcnt[lane * kCPad + warp];
      uint32_t incl = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
      }
      sm.wcnt[lane * kCPad + warp] = incl - v;
      const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
      if (lane < kCS) st_cluster_u32(mapa_u32(cntmat_saddr, (uint32_t)lane) + (rank * kCClasses + warp) * 4u, total);
    }
    cluster_arrive();
    cluster_wait();  // B1: cntmat complete everywhere

    // ---- P1c: where every (warp, class) run starts in its owner's inbox; deliver the entries -------
    {
      const uint32_t c_me = sm.cntmat[(lane & 15) * kCClasses + warp];
      const uint32_t c_sib = sm.cntmat[(lane & 15) * kCClasses + (warp ^ 1)];
      const uint32_t before = __reduce_add_sync(0xFFFFFFFFu, (lane < (int)rank && lane < kCS) ? c_me : 0u);
      const uint32_t tot_me = __reduce_add_sync(0xFFFFFFFFu, lane < kCS ? c_me : 0u);
      const uint32_t tot_sib = __reduce_add_sync(0xFFFFFFFFu, lane < kCS ? c_sib : 0u);
      const uint32_t owner = (uint32_t)warp >> 1;
      const uint32_t coff = (warp & 1) ? tot_sib : 0u;
      if ((tot_me > (uint32_t)kCClassMax || tot_me + tot_sib > (uint32_t)kCCap) && lane == 0) sm.over = 1u;
      const uint32_t run0 = mapa_u32(inbox_saddr, owner) + 8u * (coff + before);
      sm.comb[lane * kCPad + warp] = run0 + 8u * sm.wcnt[lane * kCPad + warp];
      if (owner == rank && lane == 0) {
        sm.ne[warp & 1] = tot_me;
        sm.coff[warp & 1] = coff;
      }
    }
    __syncthreads();
    const bool over = sm.over != 0u;  // the same decision in all 16 CTAs (same cntmat)
    if (!over) {
#pragma unroll
      for (int k = 0; k < kCPP; ++k) {
        if (key[k] != kEmpty) {
          const uint32_t cls = (key[k] * kGold) >> 27;
          const uint32_t r = (rbp[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
          st_cluster_v2(sm.comb[warp * kCPad + cls] + 8u * r, key[k], (uint32_t)(cta_base + k * kCT + tid));
        }
      }
    }
    cluster_arrive();
    cluster_wait();  // B2: inboxes complete

    // ---- P2: grouping, one hash class after the other ---------------------------------------------
    uint4* __restrict__ rec = w.rec(f);
    if (!over) {
#pragma unroll 1
      for (int v = 0; v < kCV; ++v) {
        const int ne = (int)sm.ne[v];
        uint2* inb = sm.inbox + sm.coff[v];
        {  // hkey and hv are contiguous
          uint4* k4 = reinterpret_cast<uint4*>(sm.hkey);
          uint4* v4 = reinterpret_cast<uint4*>(sm.hv);
          for (int s = tid; s < kCSlots / 4; s += kCT) {
            k4[s] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
            v4[s] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        __syncthreads();
        // insert: slot of the entry's cell (one CAS per probe), arrival rank inside the cell
        uint32_t sr[kCIns];
#pragma unroll
        for (int j = 0; j < kCIns; ++j) {
          const int e = tid + j * kCT;
          sr[j] = 0u;
          if (e < ne) {
            const uint32_t k = inb[e].x;
            uint32_t s = ((k * kGold) >> (27 - kCLog2Slots)) & (uint32_t)(kCSlots - 1);
            while (true) {
              const uint32_t old = atomicCAS(&sm.hkey[s], kEmpty, k);
              if (old == kEmpty || old == k) break;
              s = (s + 1u) & (uint32_t)(kCSlots - 1);
            }
            const uint32_t r = atomicAdd(&sm.hv[s], 1u);
            sr[j] = s | (r << kCLog2Slots);
          }
        }
        __syncthreads();
        // one scan over the slots (8 per thread): cell numbers (low half) and segment offsets (high half)
        uint32_t ncell;
        {
          uint4* v4 = reinterpret_cast<uint4*>(sm.hv);
          const uint4 a = v4[2 * tid], b = v4[2 * tid + 1];
          const uint32_t cn[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
          uint32_t pre[8];
          uint32_t run = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            pre[i] = run;
            run += (cn[i] << 16) | (cn[i] ? 1u : 0u);
          }
          uint32_t total;
          const uint32_t ex = block_exscan(run, sm.warp_sums, &total);
          ncell = total & 0xFFFFu;
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t p = ex + pre[i];
            o[i] = (p & 0xFFFF0000u) | cn[i];
            if (cn[i]) sm.celllist[p & 0xFFFFu] = (uint16_t)(8 * tid + i);
          }
          v4[2 * tid] = make_uint4(o[0], o[1], o[2], o[3]);
          v4[2 * tid + 1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
        __syncthreads();
        // point indices into cell-contiguous order (the key words of the inbox are dead by now)
#pragma unroll
        for (int j = 0; j < kCIns; ++j) {
          const int e = tid + j * kCT;
          if (e < ne) {
            const uint32_t s = sr[j] & (uint32_t)(kCSlots - 1), r = sr[j] >> kCLog2Slots;
            inb[(sm.hv[s] >> 16) + r].x = inb[e].y;
          }
        }
        __syncthreads();
        // one thread per cell: the 5 smallest point indices, ascending
#pragma unroll 1
        for (int j = tid; j < (int)ncell; j += kCT) {
          const uint32_t hvv = sm.hv[sm.celllist[j]];
          const uint32_t off = hvv >> 16, cnt = hvv & 0xFFFFu;
          uint32_t s0 = inb[off].x;
          uint32_t s1 = cnt > 1u ? inb[off + 1].x : kEmpty;
          uint32_t s2 = cnt > 2u ? inb[off + 2].x : kEmpty;
          uint32_t s3 = cnt > 3u ? inb[off + 3].x : kEmpty;
          uint32_t s4 = cnt > 4u ? inb[off + 4].x : kEmpty;
          if (cnt > 1u) {
            PCFE_CE(s0, s1); PCFE_CE(s3, s4); PCFE_CE(s2, s4); PCFE_CE(s2, s3); PCFE_CE(s1, s4);
            PCFE_CE(s0, s3); PCFE_CE(s0, s2); PCFE_CE(s1, s3); PCFE_CE(s1, s2);
            for (uint32_t t = 5; t < cnt; ++t) {  // more than 5 points: the largest of the six drops out
              uint32_t x = inb[off + t].x;
              PCFE_CE(s0, x); PCFE_CE(s1, x); PCFE_CE(s2, x); PCFE_CE(s3, x); PCFE_CE(s4, x);
            }
            // rec[first] = {idx1 .. idx4}, kEmpty = no point (see hvb_bucket_rec_kernel)
            rec[PCFE_REC_STRIDE * (size_t)s0] = make_uint4(s1, s2, s3, s4);
            if (PCFE_REC_STRIDE == 2) rec[2 * (size_t)s0 + 1] = make_uint4(0u, 0u, 0u, 0u);
          }
          atomicOr(&bm64[s0 >> 5], (1ull << (s0 & 31)) | (cnt > 1u ? (1ull << (32 + (s0 & 31))) : 0ull));
        }
        __syncthreads();  // the next class re-initialises the table
      }
    }
    cluster_arrive();
    cluster_wait();  // B3: the frame's mask is complete

    // ---- P3: firsts[voxel id] = position of the v-th set bit (| has-more << 31), voxel_num --------
    uint32_t* stage = sm.hkey;  // hkey | hv: 16384 words >= 384 * 32
    const int wd = (int)rank * slice + tid;
    uint2 mine = make_uint2(0u, 0u);
    if (!over && tid < slice && wd < words) {
      const unsigned long long m64 = __ldcg(bm64 + wd);
      mine = make_uint2((uint32_t)m64, (uint32_t)(m64 >> 32));
    }
    uint32_t total;
    uint32_t pos = block_exscan((uint32_t)__popc(mine.x), sm.warp_sums, &total);
    if (tid < kCS) st_cluster_u32(mapa_u32(tot_saddr, (uint32_t)tid) + rank * 4u, total);
    {
      uint32_t bits = mine.x;
      while (bits) {
        const int bit = __ffs(bits) - 1;
        bits &= bits - 1u;
        stage[pos++] = ((uint32_t)wd * 32u + (uint32_t)bit) | (((mine.y >> bit) & 1u) << 31);
      }
    }
    cluster_arrive();
    cluster_wait();  // B4: slice totals everywhere (also a CTA barrier: the stage is complete)
    if (!over) {
      uint32_t before = 0, all = 0;
#pragma unroll
      for (int r = 0; r < kCS; ++r) {
        const uint32_t t = sm.tot[r];
        before += r < (int)rank ? t : 0u;
        all += t;
      }
      uint32_t* __restrict__ firsts = w.firsts(f);
      for (uint32_t i = tid; i < total; i += kCT)
        if (before + i < (uint32_t)max_voxels) firsts[before + i] = stage[i];  // voxelization_cpu.cpp:78
      if (rank == 0 && tid == 0) {
        voxel_num[f] = (int32_t)min(all, (uint32_t)max_voxels);
        ctl[w.nb + kCtlOverflow] = 0u;
      }
    } else if (rank == 0 && tid == 0) {
      ctl[w.nb + kCtlOverflow] = 1u;  // hvg_slow_frame_kernel voxelizes the frame
    }
    __syncthreads();  // stage / over / ne are rewritten by the next frame
  }
}
