// Microbenchmark: HBM read bandwidth for (A) the GEMV fragment pattern on row-major [N,K] weights
// (per warp instruction: 8 rows x 64 B, rows 2K bytes apart) vs (B) fully contiguous 512 B per warp
// instruction (tile-major packed weights). Both: 256 threads, 16 x 16B loads in flight per thread.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// A: CTA owns 16 rows x K; warp w owns k-slice; thread (g,t): rows g, g+8; per 64-chunk 4 loads
__global__ void __launch_bounds__(256, 2) pat_rows(const uint16_t* W, int N, int K, uint32_t* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * 16;
  const uint16_t* wa = W + (size_t)(row0 + g) * K;
  const uint16_t* wb = W + (size_t)(row0 + g + 8) * K;
  const int chunks = K / 64, per = chunks / 8, c0 = warp * per, c1 = c0 + per;
  uint32_t acc = 0;
  for (int c = c0; c < c1; c += 4) {
    uint4 v[16];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k0 = (c + u) * 64 + 8 * t;
      v[4 * u] = ldg_stream(wa + k0); v[4 * u + 1] = ldg_stream(wa + k0 + 32);
      v[4 * u + 2] = ldg_stream(wb + k0); v[4 * u + 3] = ldg_stream(wb + k0 + 32);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
  }
  if (acc == 0x12345678u) out[blockIdx.x] = acc;
}

// B: same work split, but the CTA's 16 x K tile is stored contiguously in access order
__global__ void __launch_bounds__(256, 2) pat_packed(const uint16_t* W, int N, int K, uint32_t* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = K / 64, per = chunks / 8, c0 = warp * per, c1 = c0 + per;
  const uint16_t* base = W + (size_t)blockIdx.x * 16 * K;  // tile = [chunk][instr 4][lane 32][8 elems]
  uint32_t acc = 0;
  for (int c = c0; c < c1; c += 4) {
    uint4 v[16];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint16_t* p = base + (size_t)(c + u) * 1024 + lane * 8;
#pragma unroll
      for (int i = 0; i < 4; ++i) v[4 * u + i] = ldg_stream(p + i * 256);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
  }
  if (acc == 0x12345678u) out[blockIdx.x] = acc;
}

int main() {
  const int shapes[4][2] = {{12288, 4096}, {4096, 4096}, {22016, 4096}, {4096, 11008}};
  uint32_t* out; cudaMalloc(&out, 1 << 20);
  for (int s = 0; s < 4; ++s) {
    const int N = shapes[s][0], K = shapes[s][1];
    const size_t bytes = (size_t)N * K * 2;
    const int NB = 6;
    uint16_t* bufs[NB];
    for (int i = 0; i < NB; ++i) { cudaMalloc(&bufs[i], bytes); cudaMemset(bufs[i], i + 1, bytes); }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
      for (int it = 0; it < 3; ++it) (mode ? pat_packed : pat_rows)<<<N / 16, 256>>>(bufs[it % NB], N, K, out);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      const int iters = 30;
      for (int it = 0; it < iters; ++it) (mode ? pat_packed : pat_rows)<<<N / 16, 256>>>(bufs[it % NB], N, K, out);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("N=%d K=%d %s: %.2f us  %.1f GB/s\n", N, K, mode ? "packed" : "rows  ", ms * 1000 / iters, bytes / (ms / iters * 1e-3) / 1e9);
    }
    for (int i = 0; i < NB; ++i) cudaFree(bufs[i]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
