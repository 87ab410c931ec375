// Implicit-GEMM A-operand addressing shared by the SIMT and tcgen05 conv kernels.
// GEMM view of a convolution:  D[m, n] = sum_k A[m, k] * Wt[n, k]
//   m = output pixel (b, oy, ox) flattened,  n = output channel,
//   k = reduction index cut into 32-float chunks; chunk j of row m is 32 CONTIGUOUS
//       floats of the NHWC input (that is what makes the loads 128-byte coalesced):
//   CONV_1x1  : chunk j = channels [32j, 32j+32) of the pixel itself
//   CONV_3x3  : chunk j = tap (j / (Cin/32)) -> pixel (oy+dy, ox+dx), channels 32*(j % (Cin/32))..
//               whole chunk zero when the tap falls in the padding
//   CONV_STEM7: 7x7 stride-2 pad-3 stem: one kernel ROW (7 taps x Cin floats) is contiguous in
//               NHWC memory, so chunk j = floats [32c, 32c+32) of row ky = j / chunks_per_row,
//               c = j % chunks_per_row, starting at pixel (2oy+ky-3, 2ox-3); validity is per
//               4-float group (a group never straddles pixels because Cin % 4 == 0)
#pragma once
#include "common.cuh"

struct PixelCoord { int b, oy, ox; };

__device__ __forceinline__ PixelCoord decode_pixel(int m, int Ho, int Wo) {
  PixelCoord c;
  c.ox = m % Wo;
  const int t = m / Wo;
  c.oy = t % Ho;
  c.b = t / Ho;
  return c;
}

// Returns the address of the chunk (may be out of bounds where vmask bit is 0) and an
// 8-bit validity mask, one bit per float4 group.
__device__ __forceinline__ const float* chunk_ptr(const ConvParams& p, const PixelCoord& c, int j, uint32_t& vmask) {
  if (p.mode == CONV_1x1) {
    vmask = 0xFFu;
    return p.in + ((size_t)(c.b * p.H + c.oy) * p.W + c.ox) * p.Cin + 32 * j;
  } else if (p.mode == CONV_3x3) {
    const int cpc = p.Cin >> 5;
    const int tap = j / cpc, cc = j - tap * cpc;
    const int iy = c.oy + tap / 3 - 1, ix = c.ox + tap % 3 - 1;
    const bool ok = (iy >= 0) & (iy < p.H) & (ix >= 0) & (ix < p.W);
    vmask = ok ? 0xFFu : 0u;
    return p.in + ((ptrdiff_t)(c.b * p.H + iy) * p.W + ix) * p.Cin + 32 * cc;
  } else {
    const int ky = j / p.chunks_per_row, cc = j - ky * p.chunks_per_row;
    const int iy = 2 * c.oy + ky - 3, ix0 = 2 * c.ox - 3;
    const bool rowok = (iy >= 0) & (iy < p.H);
    uint32_t msk = 0;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int pix = (32 * cc + 4 * g) / p.Cin;
      const int ix = ix0 + pix;
      if (rowok & (pix < 7) & (ix >= 0) & (ix < p.W)) msk |= 1u << g;
    }
    vmask = msk;
    return p.in + ((ptrdiff_t)(c.b * p.H + iy) * p.W + ix0) * p.Cin + 32 * cc;
  }
}
