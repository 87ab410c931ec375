// FP32 SIMT implicit-GEMM convolution: the exact-arithmetic engine (conv backend 0).
// Same fused interface as the tcgen05 kernel (pre-activation affine+ReLU prologue, bias,
// ReLU, residual add, NHWC or NCHW store) so that every layer of the network can run on
// either engine; used to validate the tensor-core path layer by layer on the GPU and as
// the FP32 fall-back for shapes the tensor-core kernel does not take.
// Replaces torch.nn.Conv2d + BatchNorm2d + ReLU chains of reference
// lib/models/layers/Residual.py:20-35 and lib/models/hg.py:95-117.
#include "conv_gather.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 32, THREADS = 256;

__global__ void __launch_bounds__(THREADS)
conv_simt_kernel(const ConvParams p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int M = p.B * p.Ho * p.Wo;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // each thread stages 2 float4 of A and 2 float4 of B per chunk
  int a_row[2], a_g[2];
  PixelCoord a_pc[2];
  bool a_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int f = tid * 2 + i;
    a_row[i] = f >> 3; a_g[i] = f & 7;
    const int m = m0 + a_row[i];
    a_ok[i] = m < M;
    a_pc[i] = decode_pixel(a_ok[i] ? m : 0, p.Ho, p.Wo);
  }
  const int nchunks = p.K / BK;
  for (int j = 0; j < nchunks; ++j) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint32_t vm;
      const float* ptr = chunk_ptr(p, a_pc[i], j, vm);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_ok[i] && ((vm >> a_g[i]) & 1u)) {
        v = *reinterpret_cast<const float4*>(ptr + 4 * a_g[i]);
        if (p.pre_scale) {
          const int k = 32 * j + 4 * a_g[i];
          const float4 sc = *reinterpret_cast<const float4*>(p.pre_scale + k);
          const float4 sh = *reinterpret_cast<const float4*>(p.pre_shift + k);
          v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f); v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
          v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f); v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
        }
      }
      As[4 * a_g[i] + 0][a_row[i]] = v.x; As[4 * a_g[i] + 1][a_row[i]] = v.y;
      As[4 * a_g[i] + 2][a_row[i]] = v.z; As[4 * a_g[i] + 3][a_row[i]] = v.w;
      const int n = n0 + a_row[i];
      float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < p.Cout_pad) wv = *reinterpret_cast<const float4*>(p.w + (size_t)n * p.K + 32 * j + 4 * a_g[i]);
      Bs[4 * a_g[i] + 0][a_row[i]] = wv.x; Bs[4 * a_g[i] + 1][a_row[i]] = wv.y;
      Bs[4 * a_g[i] + 2][a_row[i]] = wv.z; Bs[4 * a_g[i] + 3][a_row[i]] = wv.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(av[i], bv[jj], acc[i][jj]);
    }
    __syncthreads();
  }
  // epilogue
  const int nb = n0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int n = nb + jj;
      if (n >= p.Cout) continue;
      float v = acc[i][jj] + p.bias[n];
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.residual) v += p.residual[(size_t)m * p.out_c + n];
      if (p.out_nchw) {
        const int hw = p.Ho * p.Wo;
        const int b = m / hw, r = m - b * hw;
        p.out[((size_t)b * p.Cout + n) * hw + r] = v;
      } else {
        p.out[(size_t)m * p.out_c + n] = v;
      }
    }
  }
}

}  // namespace

int launch_conv_simt(suo_ctx* ctx, const ConvParams& p, cudaStream_t s) {
  const int M = p.B * p.Ho * p.Wo;
  if (p.K % 32 || p.Cin % 4) { ctx->set_error("conv_simt: K % 32 / Cin % 4", __FILE__, __LINE__); return SUO_E_INVALID; }
  dim3 grid((M + BM - 1) / BM, (p.Cout_pad + BN - 1) / BN);
  conv_simt_kernel<<<grid, THREADS, 0, s>>>(p);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}
