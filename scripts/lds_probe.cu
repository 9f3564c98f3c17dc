// Shared-memory data-pipe cost of the load patterns the projection kernels use, measured as
// SM cycles per warp-wide load instruction with the pipe saturated (16 warps per SM, independent
// loads):  nvcc -arch=sm_100a -O3 -o lds_probe lds_probe.cu
//
// pattern: which 16-byte (LDS.128) or 8-byte (LDS.64) element of a row lane l reads
//   0  l            every lane its own element (no sharing)
//   1  l % 8        eight distinct elements, repeated per quarter warp (lanes 0-7 distinct)
//   2  l / 4        eight distinct elements, four ADJACENT lanes share one
//   3  l / 8        four distinct elements, a quarter warp shares one
//   4  0            one element for the whole warp
//   5  l % 16       sixteen distinct, repeated per half warp
//   6  l / 2        sixteen distinct, two adjacent lanes share one
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int ROWS = 64;

__device__ __forceinline__ int elem_of(int pattern, int l) {
  switch (pattern) {
    case 0: return l;
    case 1: return l % 8;
    case 2: return l / 4;
    case 3: return l / 8;
    case 4: return 0;
    case 5: return l % 16;
    default: return l / 2;
  }
}

template <int WIDTH>  // 16: LDS.128, 8: LDS.64
__global__ void __launch_bounds__(512) probe(double* out, long long* cycles, int pattern) {
  __shared__ __align__(16) double sh[ROWS * 64 + 64];  // rows of 64 doubles
  for (int i = threadIdx.x; i < ROWS * 64 + 64; i += blockDim.x) sh[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = elem_of(pattern, lane);
  unsigned acc0 = 0;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sh);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < ITERS; ++i) {
    const int row = (i + warp) & (ROWS - 1);
    // (volatile asm: the load keeps its full width although one word of it is consumed -- one
    // integer instruction per load, so that the load pipe and nothing else is the limit)
    if (WIDTH == 16) {
      unsigned a, b, c, d;
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                   : "r"(base + (unsigned)(row * 64 + 2 * e) * 8u));
      acc0 ^= a;
    } else {
      unsigned a, b;
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(base + (unsigned)(row * 64 + e) * 8u));
      acc0 ^= a;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = (double)acc0;
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * sizeof(double));
  cudaMalloc(&cyc, sizeof(long long));
  const char* names[7] = {"l (all distinct)", "l % 8", "l / 4", "l / 8", "0 (one element)", "l % 16", "l / 2"};
  printf("SM cycles per warp-wide load instruction, 16 warps per SM issuing back to back\n");
  printf("%-20s %10s %10s\n", "lane -> element", "LDS.128", "LDS.64");
  for (int p = 0; p < 7; ++p) {
    double r[2];
    for (int w = 0; w < 2; ++w) {
      for (int rep = 0; rep < 2; ++rep) {
        if (w == 0) probe<16><<<148, 512>>>(out, cyc, p);
        else probe<8><<<148, 512>>>(out, cyc, p);
        cudaDeviceSynchronize();
      }
      long long hc;
      cudaMemcpy(&hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
      r[w] = (double)hc / ((double)ITERS * 16);  // 16 warps per block, one block per SM
    }
    printf("%-20s %10.2f %10.2f\n", names[p], r[0], r[1]);
  }
  return cudaGetLastError() != cudaSuccess;
}
