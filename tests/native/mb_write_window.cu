// How DRAM takes a store stream as a function of WHERE the resident CTAs write at one moment
// (round 2, second session: the reason the record expansion deals a frame's tiles round-robin).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tests/native/mb_write_window.cu -o tests/native/mb_write_window
// Every variant writes the same `bytes` of zeros with 16-byte streaming stores, 128 threads per CTA:
//   grid-stride : thread t of the whole grid writes float4 t, t + T, t + 2T, ... (the chip inside one moving window)
//   region R    : CTA b owns the contiguous R bytes [b R, (b + 1) R) and writes them front to back
//   region R, 4 warps apart : the same region, every warp writing its own contiguous quarter (what a
//                 "warp = consecutive tiles" assignment does)
#include <cstdio>
#include <cuda_runtime.h>

__global__ void w_gridstride(float4* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) __stcs(&out[i], make_float4(0.f, 0.f, 0.f, 0.f));
}
__global__ void w_region(float4* __restrict__ out, size_t n, int r4 /* float4s per CTA */) {
  const size_t base = (size_t)blockIdx.x * r4;
  for (int i = threadIdx.x; i < r4 && base + i < n; i += blockDim.x) __stcs(&out[base + i], make_float4(0.f, 0.f, 0.f, 0.f));
}
__global__ void w_region_warps(float4* __restrict__ out, size_t n, int r4) {
  const int q4 = r4 / 4, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t base = (size_t)blockIdx.x * r4 + (size_t)wid * q4;
  for (int i = lane; i < q4 && base + i < n; i += 32) __stcs(&out[base + i], make_float4(0.f, 0.f, 0.f, 0.f));
}
template <typename F>
float time_ms(F f, int reps = 7) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  return best;
}
int main() {
  const size_t bytes = 700ull << 20, n = bytes / 16;
  float4* buf; cudaMalloc(&buf, bytes);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float ms = time_ms([&] { w_gridstride<<<sms * 16, 128>>>(buf, n); });
  printf("grid-stride, %d CTAs            : %.4f ms  %.0f GB/s\n", sms * 16, ms, bytes / ms / 1e6);
  for (int kb : {2, 5, 10, 20, 40, 80, 160, 640}) {
    const int r4 = kb * 1024 / 16;
    const unsigned grid = (unsigned)((n + r4 - 1) / r4);
    ms = time_ms([&] { w_region<<<grid, 128>>>(buf, n, r4); });
    float ms2 = time_ms([&] { w_region_warps<<<grid, 128>>>(buf, n, r4); });
    printf("region %3d KB per CTA (%7u CTAs): %.4f ms  %.0f GB/s   warps apart: %.4f ms  %.0f GB/s\n", kb, grid, ms, bytes / ms / 1e6, ms2, bytes / ms2 / 1e6);
  }
  return 0;
}
