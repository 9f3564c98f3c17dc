// Beam convolution (row f2 of SURVEY section 8): the step right after the projection.
//
// Replaces the per-channel scipy.signal.fftconvolve(slice, beam.kernel, mode="same") loop of
// Martini.convolve_beam (martini/martini.py:863-901) by one direct convolution kernel over the
// (nx, ny, C) cube, channel fastest:
//     out[x, y, c] = scale * sum_{a, b} in[x + ka/2 - a, y + kb/2 - b, c] * K[a, b]
// (zero outside the cube; ka, kb odd).
//
// Lane = channel, so every input load is one coalesced 256-byte row of 32 channels.  A thread
// owns a register tile of CONV_TX x CONV_TY outputs of its channel (32 float64 accumulators):
// for every input row it walks the taps b with a sliding window of CONV_TY input values held in
// registers (the b loop is unrolled by CONV_TY, so the window rotates through the registers
// without a move), and every new input value feeds up to CONV_TX * CONV_TY FMAs -- one global
// load and CONV_TX broadcast reads of the beam image (shared memory) per 32 FMAs.  Algorithmic
// work per voxel: ka * kb FMAs; FP64-pipe bound (the round-1 kernel re-read every input value
// ka times from L2: 6.4 FMAs per load, L2-bound).
#pragma once

#include "common.cuh"

namespace mtn {

constexpr int CONV_TX = 4;        // outputs along x per thread
constexpr int CONV_TY = 8;        // outputs along y per thread
constexpr int CONV_WARPS = 4;     // y strips per block
constexpr int CONV_MAX_TAPS = 28000;  // beam image in shared memory: 8 * 28000 B = 219 KB (e.g. 167 x 167)

__global__ void __launch_bounds__(CONV_WARPS * 32) convolve_beam_kernel(
    const double* __restrict__ in, double* __restrict__ out, int nx, int ny, int nc,
    const double* __restrict__ K, int ka, int kb, double scale) {
  MTN_DYN_SMEM(double, sK);
  for (int i = threadIdx.x; i < ka * kb; i += blockDim.x) sK[i] = K[i];
  __syncthreads();
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y0 = (blockIdx.y * CONV_WARPS + (threadIdx.x >> 5)) * CONV_TY;
  const int x0 = blockIdx.z * CONV_TX;
  if (c >= nc || y0 >= ny) return;
  const int ha = ka / 2, hb = kb / 2;
  double acc[CONV_TX][CONV_TY];
#pragma unroll
  for (int i = 0; i < CONV_TX; ++i)
#pragma unroll
    for (int t = 0; t < CONV_TY; ++t) acc[i][t] = 0.0;

  // input row xi feeds output row x0 + i through tap a_i = x0 + i + ha - xi
  for (int xi = max(0, x0 - ha); xi <= min(nx - 1, x0 + CONV_TX - 1 + ha); ++xi) {
    const double* row = in + (size_t)xi * ny * nc + c;
    auto load = [&](int yi) { return (yi >= 0 && yi < ny) ? row[(size_t)yi * nc] : 0.0; };
    const double* Ka[CONV_TX];
    bool on[CONV_TX];
#pragma unroll
    for (int i = 0; i < CONV_TX; ++i) {
      const int a = x0 + i + ha - xi;
      on[i] = a >= 0 && a < ka && x0 + i < nx;
      Ka[i] = sK + (on[i] ? a : 0) * kb;
    }
    // window: at tap b the outputs y0 + t need the inputs s + t, s = y0 + hb - b; the input
    // s + t lives in w[(t - b) & 7], so one step of b replaces exactly one register
    double w[CONV_TY];
#pragma unroll
    for (int t = 1; t < CONV_TY; ++t) w[t] = load(y0 + hb + t);
    for (int b0 = 0; b0 < kb; b0 += CONV_TY) {
#pragma unroll
      for (int r = 0; r < CONV_TY; ++r) {
        const int b = b0 + r;
        if (b < kb) {  // (warp-uniform)
          w[(CONV_TY - r) & (CONV_TY - 1)] = load(y0 + hb - b);
#pragma unroll
          for (int i = 0; i < CONV_TX; ++i) {
            if (on[i]) {  // (block-uniform)
              const double k = Ka[i][b];
#pragma unroll
              for (int t = 0; t < CONV_TY; ++t)
                acc[i][t] = fma(w[(t - r + CONV_TY) & (CONV_TY - 1)], k, acc[i][t]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < CONV_TX; ++i)
#pragma unroll
    for (int t = 0; t < CONV_TY; ++t)
      if (x0 + i < nx && y0 + t < ny) out[((size_t)(x0 + i) * ny + y0 + t) * nc + c] = acc[i][t] * scale;
}

}  // namespace mtn
