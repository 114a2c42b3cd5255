// Store-pattern microbenchmark for the GEMM epilogue: 148 persistent CTAs x 8 warps write a [M, N] bf16 matrix tile by
// tile (128 x 256) the way an epilogue does, with different lane -> address mappings. No MMA, no loads: the achievable
// global-store rate of each pattern.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/store_pattern store_pattern.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void st_v8(void* p, uint32_t a) {
  asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(p), "r"(a) : "memory");
}
__device__ __forceinline__ void st_v4(void* p, uint32_t a) {
  asm volatile("st.global.v4.b32 [%0], {%1, %1, %1, %1};" ::"l"(p), "r"(a) : "memory");
}

// mode 0: thread = row, 32-column chunk = 2 x 32-byte stores per lane (the v2 epilogue)
// mode 1: thread = row, 64-column pair = 4 x 32-byte stores per lane (one full 128-byte line per lane)
// mode 2: coalesced: 8 lanes x 16 bytes cover one row's 128-byte line, a warp instruction writes 4 full lines
// mode 3: coalesced 32 bytes per lane: 4 lanes per row line, 8 rows per instruction
template <int MODE>
__global__ void __launch_bounds__(256, 1) store_kernel(uint16_t* out, int M, int N, int ld) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, half = warp >> 2;
  const int mb = (M + 127) / 128, nb = (N + 255) / 256;
  for (int t = blockIdx.x; t < mb * nb; t += gridDim.x) {
    const int m0 = (t % mb) * 128 + q * 32, n0 = (t / mb) * 256;
    if (MODE == 0) {
      for (int c = half; c < 8; c += 2) {
        const int col = n0 + c * 32;
        if (col + 32 <= N && m0 + lane < M) {
          uint16_t* d = out + (size_t)(m0 + lane) * ld + col;
          st_v8(d, lane); st_v8(d + 16, lane);
        }
      }
    } else if (MODE == 1) {
      for (int c = half; c < 4; c += 2) {
        const int col = n0 + c * 64;
        if (col + 64 <= N && m0 + lane < M) {
          uint16_t* d = out + (size_t)(m0 + lane) * ld + col;
          st_v8(d, lane); st_v8(d + 16, lane); st_v8(d + 32, lane); st_v8(d + 48, lane);
        }
      }
    } else if (MODE == 2) {
      for (int c = half; c < 4; c += 2) {
        const int col = n0 + c * 64;
        if (col + 64 <= N)
          for (int r = 0; r < 32; r += 4) {
            const int row = m0 + r + (lane >> 3);
            if (row < M) st_v4(out + (size_t)row * ld + col + (lane & 7) * 8, lane);
          }
      }
    } else {
      for (int c = half; c < 4; c += 2) {
        const int col = n0 + c * 64;
        if (col + 64 <= N)
          for (int r = 0; r < 32; r += 8) {
            const int row = m0 + r + (lane >> 2);
            if (row < M) st_v8(out + (size_t)row * ld + col + (lane & 3) * 16, lane);
          }
      }
    }
  }
}

template <int MODE>
float run(uint16_t* out, int M, int N, int ld, int grid) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) store_kernel<MODE><<<grid, 256>>>(out, M, N, ld);
  cudaEventRecord(a);
  for (int i = 0; i < 20; ++i) store_kernel<MODE><<<grid, 256>>>(out, M, N, ld);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / 20 * 1e3f;
}

int main() {
  const int M = 40960;
  uint16_t* out;
  cudaMalloc(&out, (size_t)M * 2560 * 2);
  for (int N : {320, 960, 1280, 2560}) {
    for (int grid : {148, 296}) {
      float t0 = run<0>(out, M, N, N, grid), t1 = run<1>(out, M, N, N, grid), t2 = run<2>(out, M, N, N, grid), t3 = run<3>(out, M, N, N, grid);
      const double mb = (double)M * (N / 64 * 64) * 2 / 1e6;
      printf("{\"M\": %d, \"N\": %d, \"grid\": %d, \"MB\": %.1f, \"row32B_us\": %.1f, \"row128B_us\": %.1f, \"coal16B_us\": %.1f, \"coal32B_us\": %.1f, "
             "\"row32B_TBs\": %.2f, \"coal16B_TBs\": %.2f}\n", M, N, grid, mb, t0, t1, t2, t3, mb / t0, mb / t2);
    }
  }
  return 0;
}
