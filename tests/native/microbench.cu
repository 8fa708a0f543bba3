// tests/native/microbench.cu -- throughput probes that ground the kernel design:
// shared-memory atomics, L2 atomics and random L2 sector reads on the B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int OP>  // 0 add, 1 min, 2 cas, 3 plain store, 4 plain load
__global__ void smem_atomics(int iters, int slots, uint32_t* out) {
  extern __shared__ uint32_t s[];
  for (int i = threadIdx.x; i < slots; i += blockDim.x) s[i] = 0xFFFFFFFFu;
  __syncthreads();
  uint32_t st = mix(blockIdx.x * 1024 + threadIdx.x + 1);
  uint32_t acc = 0;
  for (int i = 0; i < iters; ++i) {
    uint32_t a = mix(st + i) % slots;
    if (OP == 0) atomicAdd(&s[a], 1u);
    if (OP == 1) atomicMin(&s[a], st + i);
    if (OP == 2) acc += atomicCAS(&s[a], 0xFFFFFFFFu, st);
    if (OP == 3) s[a] = st;
    if (OP == 4) acc += s[a];
  }
  __syncthreads();
  if (acc == 12345) out[0] = acc + s[0];
}

template <int OP>  // 0 atomicMin, 1 atomicAdd, 2 load 4B (ldcg), 3 CAS, 4 red (no return min), 5 store 4B
__global__ void l2_random(uint32_t* __restrict__ buf, uint32_t mask, int iters, uint32_t* out) {
  uint32_t st = mix((blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 7);
  uint32_t acc = 0;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    uint32_t a = mix(st + i * 0x9E3779B9u) & mask;
    if (OP == 0) acc += atomicMin(&buf[a], st + i);
    if (OP == 1) acc += atomicAdd(&buf[a], 1u);
    if (OP == 2) acc += __ldcg(&buf[a]);
    if (OP == 3) acc += atomicCAS(&buf[a], 0xFFFFFFFFu, st);
    if (OP == 4) atomicMin(&buf[a], st + i);
    if (OP == 5) buf[a] = st;
  }
  if (acc == 12345) out[0] = acc;
}

__global__ void stream_copy(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i];
}
__global__ void stream_write(float4* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) __stcs(&out[i], make_float4(0.f, 0.f, 0.f, 0.f));
}

template <typename F>
float time_ms(F f, int reps = 5) {
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
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s SMs %d L2 %d MB smem/SM %zu KB clock %d MHz\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20, p.sharedMemPerMultiprocessor >> 10, p.clockRate / 1000);
  uint32_t* out; CK(cudaMalloc(&out, 4));
  const int sms = p.multiProcessorCount;
  // ---- shared memory atomics -------------------------------------------------------------
  {
    const int iters = 4096, slots = 8192, threads = 512;
    const int grid = sms * 2;
    const double ops = (double)grid * threads * iters;
    const char* names[5] = {"atomicAdd", "atomicMin", "atomicCAS", "store", "load"};
    float ms[5];
    ms[0] = time_ms([&] { smem_atomics<0><<<grid, threads, slots * 4>>>(iters, slots, out); });
    ms[1] = time_ms([&] { smem_atomics<1><<<grid, threads, slots * 4>>>(iters, slots, out); });
    ms[2] = time_ms([&] { smem_atomics<2><<<grid, threads, slots * 4>>>(iters, slots, out); });
    ms[3] = time_ms([&] { smem_atomics<3><<<grid, threads, slots * 4>>>(iters, slots, out); });
    ms[4] = time_ms([&] { smem_atomics<4><<<grid, threads, slots * 4>>>(iters, slots, out); });
    for (int k = 0; k < 5; ++k)
      printf("smem random %-10s: %8.1f Gops/s chip  (%.2f ops/clk/SM at 1.9 GHz)\n", names[k], ops / ms[k] / 1e6, ops / ms[k] / 1e6 / sms / 1.9);
  }
  // ---- L2 random ops -----------------------------------------------------------------------
  for (size_t mb : {2, 8, 32, 128, 1024}) {
    size_t words = (mb << 20) / 4;
    uint32_t* buf; CK(cudaMalloc(&buf, words * 4)); CK(cudaMemset(buf, 0xFF, words * 4));
    const int iters = 64, threads = 256, grid = sms * 8 * 4;
    const double ops = (double)grid * threads * iters;
    const char* names[6] = {"atomicMin(ret)", "atomicAdd(ret)", "load4B", "atomicCAS", "red.min", "store4B"};
    float ms[6];
    ms[0] = time_ms([&] { l2_random<0><<<grid, threads>>>(buf, (uint32_t)words - 1, iters, out); });
    ms[1] = time_ms([&] { l2_random<1><<<grid, threads>>>(buf, (uint32_t)words - 1, iters, out); });
    ms[2] = time_ms([&] { l2_random<2><<<grid, threads>>>(buf, (uint32_t)words - 1, iters, out); });
    ms[3] = time_ms([&] { l2_random<3><<<grid, threads>>>(buf, (uint32_t)words - 1, iters, out); });
    ms[4] = time_ms([&] { l2_random<4><<<grid, threads>>>(buf, (uint32_t)words - 1, iters, out); });
    ms[5] = time_ms([&] { l2_random<5><<<grid, threads>>>(buf, (uint32_t)words - 1, iters, out); });
    for (int k = 0; k < 6; ++k) printf("global random %-14s over %5zu MB: %8.1f Gops/s\n", names[k], mb, ops / ms[k] / 1e6);
    cudaFree(buf);
  }
  // ---- streaming ---------------------------------------------------------------------------
  {
    size_t bytes = 2ull << 30; float4 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes));
    float ms = time_ms([&] { stream_copy<<<sms * 16, 512>>>(a, b, bytes / 16); });
    printf("stream copy   : %.1f GB/s (r+w)\n", 2.0 * bytes / ms / 1e6);
    ms = time_ms([&] { stream_write<<<sms * 16, 512>>>(b, bytes / 16); });
    printf("stream write  : %.1f GB/s\n", 1.0 * bytes / ms / 1e6);
    ms = time_ms([&] { cudaMemsetAsync(b, 0, bytes); });
    printf("cudaMemset    : %.1f GB/s\n", 1.0 * bytes / ms / 1e6);
  }
  return 0;
}
