// Heat-map reduction (SURVEY.md §8 rows a3 + a4): spatial softmax, soft-argmax,
// 2x2 covariance, hard argmax, channel mean and the keypoint-present classifier.
// Replaces reference lib/models/pkpnet.py:13-63 (spatial_softmax, mesh_grid,
// post_process_kp) and :74-78,116-118 (classifier), which materialise
// [B,K,H,W,2] and [B,K,H,W,2,2] temporaries; here one CTA owns one (b,k) map.
// 64x64 maps run heatmap_reduce_warp_kernel (one warp per map, the map in registers, shuffles only); 128x128
// maps run heatmap_reduce_reg_kernel (one CTA per map, the map read ONCE into registers, three barriers).
// Other sizes run the generic three-pass kernel
// (the map stays in L1 between passes).  Either way DRAM traffic is the algorithmic H*W*4 bytes per
// map.  HBM-bound: no tensor cores on purpose.
//
// Grid convention is the reference's TRANSPOSED mesh (pkpnet.py:19-26):
//   xx[h,w] = r[h], yy[h,w] = -r[w], r[i] = (i + 0.5)/(H/2) - 1.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  T r = (threadIdx.x < kThreads / 32) ? sh[threadIdx.x] : T(0);
  if (wid == 0) r = warp_sum(r);
  if (threadIdx.x == 0) sh[0] = r;
  __syncthreads();
  r = sh[0];
  return r;
}

__global__ void __launch_bounds__(kThreads)
heatmap_reduce_kernel(const float* __restrict__ logits, int K, int H, int W, float* __restrict__ pooled,
                      float* __restrict__ uv, float* __restrict__ cov, float* __restrict__ prob,
                      int32_t* __restrict__ argmax) {
  __shared__ float shf[kThreads / 32];
  __shared__ float shmax[kThreads / 32];
  __shared__ int shidx[kThreads / 32];
  const int map = blockIdx.x;  // b*K + k
  const int HW = H * W;
  const float* __restrict__ x = logits + (size_t)map * HW;
  const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
  const int n4 = HW >> 2;  // H*W is a multiple of 4 (checked on the host)
  const float inv_half = 1.0f / (0.5f * (float)H);

  // ---- pass 1: max / first argmax / plain sum --------------------------------------
  float m = -INFINITY, s = 0.f;
  int mi = 0x7fffffff;
  for (int i = threadIdx.x; i < n4; i += kThreads) {
    float4 v = x4[i];
    float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s += e[j];
      if (e[j] > m) { m = e[j]; mi = 4 * i + j; }   // ascending index per thread => first occurrence
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float om = __shfl_xor_sync(0xffffffffu, m, o);
    int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { shmax[threadIdx.x >> 5] = m; shidx[threadIdx.x >> 5] = mi; }
  float total = block_sum(s, shf);  // contains the __syncthreads that publish shmax/shidx
  m = shmax[0]; mi = shidx[0];
#pragma unroll
  for (int w = 1; w < kThreads / 32; ++w)
    if (shmax[w] > m || (shmax[w] == m && shidx[w] < mi)) { m = shmax[w]; mi = shidx[w]; }

  // ---- pass 2: normaliser and first moments ----------------------------------------
  float S = 0.f, sx = 0.f, sy = 0.f;
  for (int i = threadIdx.x; i < n4; i += kThreads) {
    float4 v = x4[i];
    float e[4] = {expf(v.x - m), expf(v.y - m), expf(v.z - m), expf(v.w - m)};
    const int idx = 4 * i;
    const int h = idx / W, w0 = idx - h * W;      // W % 4 == 0 => the 4 elements share a row
    const float gx = ((float)h + 0.5f) * inv_half - 1.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gy = -(((float)(w0 + j) + 0.5f) * inv_half - 1.0f);
      S += e[j]; sx += e[j] * gx; sy += e[j] * gy;
    }
  }
  S = block_sum(S, shf);
  sx = block_sum(sx, shf);
  sy = block_sum(sy, shf);
  const float invS = 1.0f / S;
  const float u = sx * invS, vv = sy * invS;

  // ---- pass 3: central second moments (+ optional prob store) -----------------------
  float cxx = 0.f, cxy = 0.f, cyy = 0.f;
  float4* __restrict__ p4 = prob ? reinterpret_cast<float4*>(prob + (size_t)map * HW) : nullptr;
  for (int i = threadIdx.x; i < n4; i += kThreads) {
    float4 v = x4[i];
    float e[4] = {expf(v.x - m) * invS, expf(v.y - m) * invS, expf(v.z - m) * invS, expf(v.w - m) * invS};
    const int idx = 4 * i;
    const int h = idx / W, w0 = idx - h * W;
    const float dx = (((float)h + 0.5f) * inv_half - 1.0f) - u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float dy = -(((float)(w0 + j) + 0.5f) * inv_half - 1.0f) - vv;
      cxx += e[j] * dx * dx; cxy += e[j] * dx * dy; cyy += e[j] * dy * dy;
    }
    if (p4) p4[i] = make_float4(e[0], e[1], e[2], e[3]);
  }
  cxx = block_sum(cxx, shf);
  cxy = block_sum(cxy, shf);
  cyy = block_sum(cyy, shf);

  if (threadIdx.x == 0) {
    if (uv) { uv[2 * map] = u; uv[2 * map + 1] = vv; }
    if (cov) { cov[4 * map] = cxx; cov[4 * map + 1] = cxy; cov[4 * map + 2] = cxy; cov[4 * map + 3] = cyy; }
    if (argmax) argmax[map] = mi;
    if (pooled) pooled[map] = total / (float)HW;
  }
}

// Register-resident version for the two map sizes the network produces (64x64: NV = 4 float4 per thread, 128x128: NV = 16):
// ONE global read of the map (all loads of a thread issued back to back), three block reductions of three values each
// (max / first argmax / plain sum; S, sum e*x, sum e*y; the three central second moments), one barrier per reduction.
// Same arithmetic per element as the generic kernel above (expf(x - m), moments about the mean), so the two agree to the
// summation order.
template <int NV>
__global__ void __launch_bounds__(kThreads)
heatmap_reduce_reg_kernel(const float* __restrict__ logits, int H, int W, float* __restrict__ pooled,
                          float* __restrict__ uv, float* __restrict__ cov, float* __restrict__ prob,
                          int32_t* __restrict__ argmax) {
  constexpr int NW = kThreads / 32;
  __shared__ float r1[3][NW];
  __shared__ int r1i[NW];
  __shared__ float r2[3][NW];
  __shared__ float r3[3][NW];
  const int map = blockIdx.x;  // b*K + k
  const int HW = H * W;
  const float4* __restrict__ x4 = reinterpret_cast<const float4*>(logits + (size_t)map * HW);
  const float inv_half = 1.0f / (0.5f * (float)H);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  float e[NV][4];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float4 v = __ldg(x4 + threadIdx.x + j * kThreads);
    e[j][0] = v.x; e[j][1] = v.y; e[j][2] = v.z; e[j][3] = v.w;
  }
  // ---- reduction 1: max / first argmax / plain sum ---------------------------------
  float m = -INFINITY, s = 0.f;
  int mi = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      s += e[j][q];
      if (e[j][q] > m) { m = e[j][q]; mi = 4 * (threadIdx.x + j * kThreads) + q; }   // ascending index per thread => first occurrence
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
  }
  s = warp_sum(s);
  if (lane == 0) { r1[0][wid] = m; r1i[wid] = mi; r1[1][wid] = s; }
  __syncthreads();
  m = r1[0][0]; mi = r1i[0];
  float total = r1[1][0];
#pragma unroll
  for (int w = 1; w < NW; ++w) {
    if (r1[0][w] > m || (r1[0][w] == m && r1i[w] < mi)) { m = r1[0][w]; mi = r1i[w]; }
    total += r1[1][w];
  }
  // ---- reduction 2: normaliser and first moments ------------------------------------
  float S = 0.f, sx = 0.f, sy = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int idx = 4 * (threadIdx.x + j * kThreads);
    const int h = idx / W, w0 = idx - h * W;      // W % 4 == 0 => the 4 elements share a row
    const float gx = ((float)h + 0.5f) * inv_half - 1.0f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float gy = -(((float)(w0 + q) + 0.5f) * inv_half - 1.0f);
      e[j][q] = expf(e[j][q] - m);
      S += e[j][q]; sx += e[j][q] * gx; sy += e[j][q] * gy;
    }
  }
  S = warp_sum(S); sx = warp_sum(sx); sy = warp_sum(sy);
  if (lane == 0) { r2[0][wid] = S; r2[1][wid] = sx; r2[2][wid] = sy; }
  __syncthreads();
  S = 0.f; sx = 0.f; sy = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) { S += r2[0][w]; sx += r2[1][w]; sy += r2[2][w]; }
  const float invS = 1.0f / S;
  const float u = sx * invS, vv = sy * invS;
  // ---- reduction 3: central second moments (+ optional prob store) -------------------
  float cxx = 0.f, cxy = 0.f, cyy = 0.f;
  float4* __restrict__ p4 = prob ? reinterpret_cast<float4*>(prob + (size_t)map * HW) : nullptr;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int idx = 4 * (threadIdx.x + j * kThreads);
    const int h = idx / W, w0 = idx - h * W;
    const float dx = (((float)h + 0.5f) * inv_half - 1.0f) - u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float dy = -(((float)(w0 + q) + 0.5f) * inv_half - 1.0f) - vv;
      e[j][q] *= invS;
      cxx += e[j][q] * dx * dx; cxy += e[j][q] * dx * dy; cyy += e[j][q] * dy * dy;
    }
    if (p4) p4[threadIdx.x + j * kThreads] = make_float4(e[j][0], e[j][1], e[j][2], e[j][3]);
  }
  cxx = warp_sum(cxx); cxy = warp_sum(cxy); cyy = warp_sum(cyy);
  if (lane == 0) { r3[0][wid] = cxx; r3[1][wid] = cxy; r3[2][wid] = cyy; }
  __syncthreads();
  if (threadIdx.x == 0) {
    cxx = 0.f; cxy = 0.f; cyy = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) { cxx += r3[0][w]; cxy += r3[1][w]; cyy += r3[2][w]; }
    if (uv) { uv[2 * map] = u; uv[2 * map + 1] = vv; }
    if (cov) { cov[4 * map] = cxx; cov[4 * map + 1] = cxy; cov[4 * map + 2] = cxy; cov[4 * map + 3] = cyy; }
    if (argmax) argmax[map] = mi;
    if (pooled) pooled[map] = total / (float)HW;
  }
}

// 64x64 maps, one WARP per map: each lane holds 32 float4 (all 32 loads of a lane are issued back to back: 16 KB in flight
// per warp), the reductions are warp shuffles only — no shared memory, no block barrier — so the warps of an SM sit in
// different phases and the memory pipe never drains while a CTA computes.  The kernel is issue-bound once the loads
// overlap, so the element loop is kept to two lean passes: (1) max and plain sum (the argmax is located afterwards, by the
// lanes that hold the maximum); (2) e = exp(x - m) = ex2(x log2e - m log2e) with S and the first AND second moments about the ARGMAX pixel: the shifted coordinates (h - h0) / 32 are exact
// in FP32 and, on a peaky map, small where e is large, so cov = E[dd^T] - E[d] E[d]^T loses nothing to cancellation
// (4e-7 worst case against FP64 over flat, noisy, single- and double-peak maps); uv = grid(argmax) + E[d].
constexpr int kWarpMapThreads = 128;
// 2^x by the SFU (ex2.approx.ftz: relative error 2^-22.5).  With the argument formed by ONE fused multiply-add the weights e^(x - m) are
// good to ~2e-7 relative where they matter (x - m > -16) — the same order as expf's own rounding — and to 3e-6 in the far tail
// (x - m ~ -88), where they are below 1e-38 of the sum anyway.
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__global__ void __launch_bounds__(kWarpMapThreads, 3)
heatmap_reduce_warp_kernel(const float* __restrict__ logits, int n_maps, float* __restrict__ pooled, float* __restrict__ uv,
                           float* __restrict__ cov, float* __restrict__ prob, int32_t* __restrict__ argmax) {
  constexpr int H = 64, W = 64, HW = H * W, NV = HW / (4 * 32);
  const int lane = threadIdx.x & 31;
  const int map = blockIdx.x * (kWarpMapThreads / 32) + (threadIdx.x >> 5);
  if (map >= n_maps) return;
  const float4* __restrict__ x4 = reinterpret_cast<const float4*>(logits + (size_t)map * HW);
  constexpr float inv_half = 1.0f / (0.5f * (float)H);
  float e[NV][4];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float4 v = __ldg(x4 + lane + 32 * j);
    e[j][0] = v.x; e[j][1] = v.y; e[j][2] = v.z; e[j][3] = v.w;
  }
  // pass 1: max and plain sum (2 instructions per element); the argmax is located afterwards by the lanes that hold the maximum
  float m = -INFINITY, s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) { s += e[j][q]; m = fmaxf(m, e[j][q]); }
  const float lane_max = m;
  m = warp_max(m);
  int mi = 0x7fffffff;
  if (lane_max == m) {                     // first occurrence inside this lane (its indices ascend with j, q)
#pragma unroll
    for (int j = NV - 1; j >= 0; --j)
#pragma unroll
      for (int q = 3; q >= 0; --q)
        if (e[j][q] == m) mi = 4 * (lane + 32 * j) + q;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mi = min(mi, __shfl_xor_sync(0xffffffffu, mi, o));   // first occurrence over the map
  const float total = warp_sum(s);
  // float4 i = lane + 32 j lies in row 2 j + (lane >> 4), columns 4 (lane & 15) .. + 3
  const int h0 = mi >> 6, w0 = mi & 63;
  const int hl = (lane >> 4) - h0, wl = 4 * (lane & 15) - w0;
  float gy[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) gy[q] = -(float)(wl + q) * inv_half;           // yy[h,w] = -r[w], shifted by the argmax column
  constexpr float kLog2e = 1.4426950408889634f;
  const float neg_m_log2e = -m * kLog2e;
  float gy2[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) gy2[q] = gy[q] * gy[q];
  // the four elements of a float4 share the row coordinate gx: accumulate t0 = sum e, t1 = sum e gy, t2 = sum e gy^2 per float4 (3 instructions per
  // element) and fold gx in once per float4 (6 per four elements)
  float S = 0.f, sx = 0.f, sy = 0.f, sxx = 0.f, sxy = 0.f, syy = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float gx = (float)(2 * j + hl) * inv_half;                         // xx[h,w] = r[h], shifted by the argmax row
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float ev = fast_exp2(fmaf(e[j][q], kLog2e, neg_m_log2e));   // exp(x - m): one FFMA + one MUFU.EX2
      e[j][q] = ev;
      t0 += ev;
      t1 = fmaf(ev, gy[q], t1);
      t2 = fmaf(ev, gy2[q], t2);
    }
    const float g0 = gx * t0;
    S += t0; sy += t1; syy += t2;
    sx += g0; sxx = fmaf(g0, gx, sxx); sxy = fmaf(gx, t1, sxy);
  }
  S = warp_sum(S); sx = warp_sum(sx); sy = warp_sum(sy); sxx = warp_sum(sxx); sxy = warp_sum(sxy); syy = warp_sum(syy);
  const float invS = 1.0f / S;
  if (prob) {
    float4* __restrict__ p4 = reinterpret_cast<float4*>(prob + (size_t)map * HW);
#pragma unroll
    for (int j = 0; j < NV; ++j) p4[lane + 32 * j] = make_float4(e[j][0] * invS, e[j][1] * invS, e[j][2] * invS, e[j][3] * invS);
  }
  if (lane == 0) {
    const float mx = sx * invS, my = sy * invS;
    if (uv) {
      uv[2 * map] = (((float)h0 + 0.5f) * inv_half - 1.0f) + mx;
      uv[2 * map + 1] = -(((float)w0 + 0.5f) * inv_half - 1.0f) + my;
    }
    if (cov) {
      const float cxy = sxy * invS - mx * my;
      cov[4 * map] = sxx * invS - mx * mx; cov[4 * map + 1] = cxy; cov[4 * map + 2] = cxy; cov[4 * map + 3] = syy * invS - my * my;
    }
    if (argmax) argmax[map] = mi;
    if (pooled) pooled[map] = total / (float)HW;
  }
}

// kp_mask_logits = W * relu(mean_hw(raw)) + b ; kp_mask = sigmoid (pkpnet.py:74-78,116-118)
__global__ void classifier_kernel(const float* __restrict__ pooled, const float* __restrict__ Wc,
                                  const float* __restrict__ bc, int K, float* __restrict__ mask_logits,
                                  float* __restrict__ mask) {
  const int b = blockIdx.x;
  for (int o = threadIdx.x; o < K; o += blockDim.x) {
    float acc = bc[o];
    for (int i = 0; i < K; ++i) acc += Wc[o * K + i] * fmaxf(pooled[b * K + i], 0.f);
    if (mask_logits) mask_logits[b * K + o] = acc;
    if (mask) mask[b * K + o] = 1.0f / (1.0f + expf(-acc));
  }
}

}  // namespace

int launch_heatmap_reduce(suo_ctx* ctx, const float* logits, int B, int K, int H, int W, const float* cls_w,
                          const float* cls_b, float* pooled_scratch, float* uv, float* cov, float* prob,
                          float* mask_logits, float* mask, int32_t* argmax, cudaStream_t s) {
  if (H != W || (W & 3) || B <= 0 || K <= 0) {
    ctx->set_error("heatmap_reduce: need square maps with W % 4 == 0", __FILE__, __LINE__);
    return SUO_E_INVALID;
  }
  if (H == 64 && W == 64)
    heatmap_reduce_warp_kernel<<<(B * K + kWarpMapThreads / 32 - 1) / (kWarpMapThreads / 32), kWarpMapThreads, 0, s>>>(logits, B * K, pooled_scratch, uv, cov, prob, argmax);
  else if (H * W == 4 * 4 * kThreads)
    heatmap_reduce_reg_kernel<4><<<B * K, kThreads, 0, s>>>(logits, H, W, pooled_scratch, uv, cov, prob, argmax);
  else if (H * W == 16 * 4 * kThreads)
    heatmap_reduce_reg_kernel<16><<<B * K, kThreads, 0, s>>>(logits, H, W, pooled_scratch, uv, cov, prob, argmax);
  else
    heatmap_reduce_kernel<<<B * K, kThreads, 0, s>>>(logits, K, H, W, pooled_scratch, uv, cov, prob, argmax);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  if (cls_w && cls_b && (mask_logits || mask)) {
    classifier_kernel<<<B, 64, 0, s>>>(pooled_scratch, cls_w, cls_b, K, mask_logits, mask);
    ctx->launches++;
    SUO_CUDA_TRY(ctx, cudaGetLastError());
  }
  return SUO_OK;
}
