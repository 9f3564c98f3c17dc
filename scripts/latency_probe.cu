// Dependent-chain latencies of the instructions the projection kernel is made of, measured on
// the device with clock64() by one warp:  nvcc -arch=sm_100a -O3 -o latency_probe latency_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int N = 4096;

template <int OP>
__global__ void probe(double* out, long long* cycles, const int* chase_g, double seed, int iseed) {
  __shared__ int chase_s[1024];
  __shared__ double dsh[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    chase_s[i] = (i * 33 + 1) & 1023;
    dsh[i] = 1.0 + 1e-9 * i;
  }
  __syncthreads();
  double x = seed, y = seed * 0.5 + 1e-3;
  int k = iseed + (threadIdx.x & 31);
  float f = (float)seed;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, y, 1e-9);                       // DFMA
    if (OP == 1) x = x * y;                                 // DMUL
    if (OP == 2) x = x + y;                                 // DADD
    if (OP == 3) k = chase_s[k & 1023];                     // LDS (32-bit) pointer chase
    if (OP == 4) k = __shfl_sync(0xffffffffu, k, (k + 1) & 31);  // SHFL
    if (OP == 5) k = __ldg(chase_g + (k & 1023));           // LDG, L1 hit, pointer chase
    if (OP == 6) x = (double)((float)x) + 1e-9;             // F2F.F32.F64 + F2F.F64.F32 + DADD
    if (OP == 7) x = (double)((int)x) + 1.5;                // F2I.F64 + I2F.F64 + DADD
    if (OP == 8) k = __popc(__ballot_sync(0xffffffffu, k & 1)) + k;  // VOTE + POPC + IADD
    if (OP == 9) f = fmaf(f, 1.0001f, 1e-9f);               // FFMA
    if (OP == 10) x = dsh[((int)__double2loint(x)) & 1023]; // LDS.64 with the address from the value
    if (OP == 11) k = __ffs(k | 0x100) + (k << 1);          // BREV + FLO + shift/add
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[OP] = t1 - t0;
  out[OP * 32 + (threadIdx.x & 31)] = x + k + f;
}

int main() {
  double* out;
  long long* cyc;
  int* chase;
  cudaMalloc(&out, 32 * 16 * sizeof(double));
  cudaMalloc(&cyc, 16 * sizeof(long long));
  cudaMalloc(&chase, 1024 * sizeof(int));
  int h[1024];
  for (int i = 0; i < 1024; ++i) h[i] = (i * 33 + 1) & 1023;
  cudaMemcpy(chase, h, sizeof(h), cudaMemcpyHostToDevice);
  const char* names[12] = {"DFMA", "DMUL", "DADD", "LDS.32 chase", "SHFL", "LDG L1-hit chase",
                           "F2F f64->f32->f64 + DADD", "F2I.F64 + I2F.F64 + DADD", "VOTE + POPC + IADD",
                           "FFMA", "LDS.64 (addr from value)", "BREV + FLO + SHL/IADD"};
  for (int rep = 0; rep < 2; ++rep) {
    probe<0><<<1, 32>>>(out, cyc, chase, 1.0000001, 3);
    probe<1><<<1, 32>>>(out, cyc, chase, 1.0000001, 3);
    probe<2><<<1, 32>>>(out, cyc, chase, 1.0000001, 3);
    probe<3><<<1, 32>>>(out, cyc, chase, 1.0000001, 3);
    probe<4><<<1, 32>>>(out, cyc, chase, 1.0000001, 3);
    probe<5><<<1, 32>>>(out, cyc, chase, 1.0000001, 3);
    probe<6><<<1, 32>>>(out, cyc, chase, 1.0000001, 3);
    probe<7><<<1, 32>>>(out, cyc, chase, 1.0000001, 3);
    probe<8><<<1, 32>>>(out, cyc, chase, 1.0000001, 3);
    probe<9><<<1, 32>>>(out, cyc, chase, 1.0000001, 3);
    probe<10><<<1, 32>>>(out, cyc, chase, 1.0000001, 3);
    probe<11><<<1, 32>>>(out, cyc, chase, 1.0000001, 3);
    cudaDeviceSynchronize();
  }
  long long hc[16];
  cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  printf("dependent-chain latency, cycles per step (one warp, %d steps)\n", N);
  for (int i = 0; i < 12; ++i) printf("%-28s %7.1f\n", names[i], (double)hc[i] / N);
  // throughput view: 4 and 8 independent DFMA chains per thread are covered by mtn_fp64_peak
  return cudaGetLastError() != cudaSuccess;
}
