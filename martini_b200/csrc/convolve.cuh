// Beam convolution (row f2 of SURVEY section 8): the step right after the projection.
//
// Replaces the per-channel scipy.signal.fftconvolve(slice, beam.kernel, mode="same") loop of
// Martini.convolve_beam (martini/martini.py:863-901) by one direct convolution kernel over the
// (nx, ny, C) cube, channel fastest:
//     out[x, y, c] = scale * sum_{a, b} in[x + ka/2 - a, y + kb/2 - b, c] * K[a, b]
// (zero outside the cube; ka, kb odd).  Lane = channel, so every tap is one coalesced load;
// each thread produces CONV_TY consecutive y outputs so a loaded input value feeds up to
// CONV_TY FMAs; the beam image sits in shared memory (broadcast reads).  FP64-pipe bound.
#pragma once

#include "common.cuh"

namespace mtn {

constexpr int CONV_TY = 8;        // outputs along y per thread
constexpr int CONV_MAX_TAPS = 96 * 96;

__global__ void __launch_bounds__(128) convolve_beam_kernel(
    const double* __restrict__ in, double* __restrict__ out, int nx, int ny, int nc,
    const double* __restrict__ K, int ka, int kb, double scale) {
  MTN_DYN_SMEM(double, sK);
  for (int i = threadIdx.x; i < ka * kb; i += blockDim.x) sK[i] = K[i];
  __syncthreads();
  // block = 32 channels x 4 y-strips; grid = (channel blocks, y strips of 4*CONV_TY, x)
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y0 = (blockIdx.y * 4 + (threadIdx.x >> 5)) * CONV_TY;
  const int x = blockIdx.z;
  if (c >= nc || y0 >= ny) return;
  const int ha = ka / 2, hb = kb / 2;
  double acc[CONV_TY];
#pragma unroll
  for (int t = 0; t < CONV_TY; ++t) acc[t] = 0.0;
  for (int a = 0; a < ka; ++a) {
    const int xi = x + ha - a;
    if (xi < 0 || xi >= nx) continue;
    const double* row = in + (size_t)xi * ny * nc + c;
    const double* Ka = sK + a * kb;
    // input y' contributes to output y = y' - hb + b, i.e. tap b = y - y' + hb
    for (int yi = max(0, y0 - hb); yi < min(ny, y0 + CONV_TY + hb); ++yi) {
      const double v = row[(size_t)yi * nc];
#pragma unroll
      for (int t = 0; t < CONV_TY; ++t) {
        const int b = y0 + t - yi + hb;
        if (b >= 0 && b < kb) acc[t] = fma(v, Ka[b], acc[t]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < CONV_TY; ++t)
    if (y0 + t < ny) out[((size_t)x * ny + y0 + t) * nc + c] = acc[t] * scale;
}

}  // namespace mtn
