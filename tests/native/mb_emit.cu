// tests/native/mb_emit.cu -- probe: how fast can 100-byte voxel records be written at SCATTERED
// voxel ids (partial 32-byte sectors merged in L2) compared with the in-order float4 stream the
// expansion kernel uses?  Decides whether a bucket CTA may emit its own voxels (no order pass).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_emit mb_emit.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <numeric>
#include <random>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int kW = 25;  // words per voxel (P = 5, C = 5)

// in-order: warp tile = 32 voxels = 800 words = 200 float4
__global__ void emit_stream(float* __restrict__ out, int m, size_t frame_stride) {
  const int f = blockIdx.y;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int v0 = warp * 32;
  if (v0 >= m) return;
  const int nw = min(32, m - v0) * kW;
  float4* dst = reinterpret_cast<float4*>(out + f * frame_stride + (size_t)v0 * kW);
  for (int i = lane; i < nw / 4; i += 32) __stcs(dst + i, make_float4(1.f, 2.f, 3.f, 4.f));
}

// scattered: CTA (f, b) owns voxels perm[f][b * per .. +per); warp takes 32 at a time, stages
// their words in shared memory, then stores word by word: lane -> (voxel w/25, word w%25)
template <int CS>
__global__ void emit_scatter(float* __restrict__ out, const int* __restrict__ perm, int m, int per,
                             size_t frame_stride, int32_t* __restrict__ coors, int32_t* __restrict__ num,
                             int with_coors) {
  __shared__ float stage[8][32 * kW];
  __shared__ int svid[8][32];
  const int f = blockIdx.y, b = blockIdx.x;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lo = b * per, hi = min(m, lo + per);
  float* fo = out + f * frame_stride;
  for (int j0 = lo + wid * 32; j0 < hi; j0 += 8 * 32) {
    const int j = j0 + lane;
    int vid = -1;
    if (j < hi) vid = perm[(size_t)f * m + j];
    svid[wid][lane] = vid;
    for (int k = 0; k < kW; ++k) stage[wid][lane * kW + k] = (float)(vid + k);
    if (with_coors && vid >= 0) {
      int32_t* co = coors + ((size_t)f * m + vid) * 3;
      co[0] = vid; co[1] = j; co[2] = b;
      num[(size_t)f * m + vid] = 5;
    }
    __syncwarp();
    const int nv = min(32, hi - j0);
    for (int w = lane; w < nv * kW; w += 32) {
      const int v = (w * 1311) >> 15;  // w / 25 for w < 800 (1311/32768 = 1/24.995)
      const int k = w - v * kW;
      float* p = fo + (size_t)svid[wid][v] * kW + k;
      if (CS) __stcs(p, stage[wid][w]); else *p = stage[wid][w];
    }
    __syncwarp();
  }
}

int main(int argc, char** argv) {
  const int F = 64, M = 95410, NB = 256;
  const int per = (M + NB - 1) / NB;
  const size_t frame_stride = (size_t)150000 * kW;
  float* out; int* perm; int32_t *coors, *num;
  CK(cudaMalloc(&out, F * frame_stride * 4));
  CK(cudaMalloc(&perm, (size_t)F * M * 4));
  CK(cudaMalloc(&coors, (size_t)F * M * 12));
  CK(cudaMalloc(&num, (size_t)F * M * 4));
  std::vector<int> h((size_t)F * M);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double payload = (double)F * M * kW * 4;
  auto run = [&](const char* name, auto launch) {
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("%-44s %8.4f ms  %7.1f GB/s payload\n", name, ms, payload / ms * 1e-6);
    return 0;
  };
  run("stream float4 stcs (in order)", [&] {
    emit_stream<<<dim3((M + 127) / 128, F), 128>>>(out, M, frame_stride); });
  for (int mode = 0; mode < 4; ++mode) {
    std::mt19937 g(123);
    for (int f = 0; f < F; ++f) {
      int* p = h.data() + (size_t)f * M;
      std::iota(p, p + M, 0);
      if (mode >= 1) std::shuffle(p, p + M, g);
      if (mode == 2) for (int b = 0; b < NB; ++b) std::sort(p + std::min(M, b * per), p + std::min(M, (b + 1) * per));
      if (mode == 3) {  // strided: CTA b owns vids b, b + NB, ...
        int k = 0;
        for (int b = 0; b < NB; ++b) for (int v = b; v < M && k < M; v += NB) p[k++] = v;
      }
    }
    CK(cudaMemcpy(perm, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    const char* mn[] = {"identity", "random perm", "random perm, sorted per CTA", "strided by NB"};
    char name[128];
    for (int cs = 0; cs < 2; ++cs) for (int wc = 0; wc < 2; ++wc) {
      snprintf(name, sizeof name, "scatter %-28s cs=%d coors=%d", mn[mode], cs, wc);
      run(name, [&] {
        if (cs) emit_scatter<1><<<dim3(NB, F), 256>>>(out, perm, M, per, frame_stride, coors, num, wc);
        else emit_scatter<0><<<dim3(NB, F), 256>>>(out, perm, M, per, frame_stride, coors, num, wc); });
    }
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
