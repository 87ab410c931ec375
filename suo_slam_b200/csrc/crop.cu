// Crop + prior concat and the two elementwise layers of the
// hourglass (2x2 max-pool, nearest x2 up-sample + add).
//
// crop_concat replaces torchvision.ops.roi_align(images, boxes, output_size=R)
// (defaults spatial_scale=1, sampling_ratio=-1, aligned=False) followed by
// torch.cat([crops, prior_kp], 1) at reference lib/models/pkpnet.py:91-101, and writes
// the NHWC tensor the conv engine consumes.  roi_align semantics follow the public
// torchvision op: adaptive ceil(roi/R) samples per bin, bilinear taps with clamp,
// samples outside [-1, size] contribute 0, mean over the samples.
#include <algorithm>
#include "common.cuh"

namespace {

// Image taps: FP32 NCHW planes (what the reference's model(...) receives) or the camera's u8 HWC frame, converted
// per tap exactly like the reference converts the whole frame on the host, float32(img) / 255
// (lib/object_slam.py:1092) — same value, a quarter of the bytes over PCIe.
template <bool U8>
__device__ __forceinline__ float tap(const void* __restrict__ img, int H, int W, int c, int y, int x) {
  if (U8) return __fdiv_rn((float)static_cast<const unsigned char*>(img)[((size_t)y * W + x) * 3 + c], 255.f);
  return static_cast<const float*>(img)[((size_t)c * H + y) * W + x];
}

template <bool U8>
__device__ __forceinline__ float bilinear(const void* __restrict__ img, int c, int H, int W, float y, float x) {
  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) return 0.f;
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int y_low = (int)y, x_low = (int)x, y_high, x_high;
  if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else { y_high = y_low + 1; }
  if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else { x_high = x_low + 1; }
  const float ly = y - (float)y_low, lx = x - (float)x_low, hy = 1.f - ly, hx = 1.f - lx;
  const float v1 = tap<U8>(img, H, W, c, y_low, x_low), v2 = tap<U8>(img, H, W, c, y_low, x_high);
  const float v3 = tap<U8>(img, H, W, c, y_high, x_low), v4 = tap<U8>(img, H, W, c, y_high, x_high);
  const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
  return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
}

// One thread per output pixel: 3 roi-aligned colour values (+ channel 3 = prior plane 0 or 0).
// out_c == 4 : write one float4 per pixel.  out_c == 48: write channels 0..2 only; the prior
// planes are transposed in by prior_to_nhwc_kernel.
template <bool U8>
__global__ void __launch_bounds__(256)
roi_align_kernel(const void* __restrict__ images, int H, int W, const float* __restrict__ boxes,
                 const int32_t* __restrict__ box_img, int R, float* __restrict__ out, int out_c,
                 float* __restrict__ out_pad, int pad_w, int pad_h) {
  const int crop = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= R * R) return;
  const int ph = pix / R, pw = pix - ph * R;
  const float* __restrict__ bx = boxes + 4 * crop;
  const float roi_start_w = bx[0], roi_start_h = bx[1], roi_end_w = bx[2], roi_end_h = bx[3];
  const float roi_w = fmaxf(roi_end_w - roi_start_w, 1.0f), roi_h = fmaxf(roi_end_h - roi_start_h, 1.0f);
  const float bin_h = roi_h / (float)R, bin_w = roi_w / (float)R;
  const int grid_h = (int)ceilf(roi_h / (float)R), grid_w = (int)ceilf(roi_w / (float)R);
  const float count = fmaxf((float)(grid_h * grid_w), 1.0f);
  const void* __restrict__ img = U8 ? static_cast<const void*>(static_cast<const unsigned char*>(images) + (size_t)box_img[crop] * 3 * H * W)
                                    : static_cast<const void*>(static_cast<const float*>(images) + (size_t)box_img[crop] * 3 * H * W);
  float acc[3] = {0.f, 0.f, 0.f};
  for (int iy = 0; iy < grid_h; ++iy) {
    const float y = roi_start_h + (float)ph * bin_h + ((float)iy + 0.5f) * bin_h / (float)grid_h;
    for (int ix = 0; ix < grid_w; ++ix) {
      const float x = roi_start_w + (float)pw * bin_w + ((float)ix + 0.5f) * bin_w / (float)grid_w;
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] += bilinear<U8>(img, c, H, W, y, x);
    }
  }
  float* __restrict__ o = out + ((size_t)crop * R * R + pix) * out_c;
  if (out_c == 4) {
    const float4 v = make_float4(acc[0] / count, acc[1] / count, acc[2] / count, 0.f);
    *reinterpret_cast<float4*>(o) = v;
    // second copy with a 3-pixel zero border (rows of pad_w pixels, pad_h rows per crop): what the TMA-fed stem reads (conv_tc.cu)
    if (out_pad) *reinterpret_cast<float4*>(out_pad + (((size_t)crop * pad_h + ph + 3) * pad_w + pw + 3) * 4) = v;
  } else {
    o[0] = acc[0] / count; o[1] = acc[1] / count; o[2] = acc[2] / count;
  }
}

// priors [L,K,R,R] (NCHW planes) -> channels 3..3+K-1 of the NHWC tensor, zero padding above.
// Block = 64 consecutive pixels of one crop; smem transpose so both sides are coalesced.
__global__ void __launch_bounds__(256)
prior_to_nhwc_kernel(const float* __restrict__ priors, int K, int RR, float* __restrict__ out, int out_c) {
  __shared__ float tile[64][49];
  const int crop = blockIdx.y, pix0 = blockIdx.x * 64;
  for (int i = threadIdx.x; i < 64 * 45; i += 256) {   // channels 3..47
    const int c = 3 + i / 64, px = i % 64;
    float v = 0.f;
    if (c - 3 < K && priors) v = priors[((size_t)crop * K + (c - 3)) * RR + pix0 + px];
    tile[px][c] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 45; i += 256) {
    const int px = i / 45, c = 3 + i % 45;
    out[((size_t)crop * RR + pix0 + px) * out_c + c] = tile[px][c];
  }
}

__global__ void __launch_bounds__(256)
maxpool2_kernel(const float4* __restrict__ in, int Ho, int Wo, int C4, size_t n_out4, float4* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out4; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    size_t r = i / C4;
    const int x = (int)(r % Wo); r /= Wo;
    const int y = (int)(r % Ho);
    const size_t b = r / Ho;
    const size_t Wi = 2 * (size_t)Wo;
    const size_t base = ((b * 2 * Ho + 2 * y) * Wi + 2 * x) * C4 + c;
    const float4 a = in[base], bq = in[base + C4], cq = in[base + Wi * C4], d = in[base + Wi * C4 + C4];
    out[i] = make_float4(fmaxf(fmaxf(a.x, bq.x), fmaxf(cq.x, d.x)), fmaxf(fmaxf(a.y, bq.y), fmaxf(cq.y, d.y)),
                         fmaxf(fmaxf(a.z, bq.z), fmaxf(cq.z, d.z)), fmaxf(fmaxf(a.w, bq.w), fmaxf(cq.w, d.w)));
  }
}

// out[b,y,x,:] = up1[b,y,x,:] + low[b,y/2,x/2,:]   (F.interpolate nearest x2 + add, hg.py:56-58)
__global__ void __launch_bounds__(256)
upsample_add_kernel(const float4* __restrict__ up1, const float4* __restrict__ low, int H, int W, int C4,
                    size_t n4, float4* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    size_t r = i / C4;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const size_t b = r / H;
    const float4 a = up1[i];
    const float4 l = low[((b * (H / 2) + (y >> 1)) * (size_t)(W / 2) + (x >> 1)) * C4 + c];
    out[i] = make_float4(a.x + l.x, a.y + l.y, a.z + l.z, a.w + l.w);
  }
}

}  // namespace

int launch_crop_concat(suo_ctx* ctx, const void* images, int n_img, int H, int W, const float* boxes,
                       const int32_t* box_img, int L, const float* priors, int num_kp, int R, float* out, int out_c,
                       cudaStream_t s, int images_u8, float* out_pad, int pad_w, int pad_h) {
  (void)n_img;
  if (!(out_c == 4 || out_c == 48) || (out_c == 4 && priors) || num_kp > 45 || (R * R) % 64) {
    ctx->set_error("crop_concat: out_c must be 4 (no priors) or 48", __FILE__, __LINE__);
    return SUO_E_INVALID;
  }
  dim3 g((R * R + 255) / 256, L);
  if (images_u8) roi_align_kernel<true><<<g, 256, 0, s>>>(images, H, W, boxes, box_img, R, out, out_c, out_pad, pad_w, pad_h);
  else roi_align_kernel<false><<<g, 256, 0, s>>>(images, H, W, boxes, box_img, R, out, out_c, out_pad, pad_w, pad_h);
  ctx->launches++;
  if (out_c == 48 && num_kp >= 0) {   // num_kp < 0: RGB only, the caller renders the prior channels itself (prior.cu)
    dim3 g2(R * R / 64, L);
    prior_to_nhwc_kernel<<<g2, 256, 0, s>>>(priors, num_kp, R * R, out, out_c);
    ctx->launches++;
  }
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

int launch_maxpool2(suo_ctx* ctx, const float* in, int B, int H, int W, int C, float* out, cudaStream_t s) {
  const size_t n4 = (size_t)B * (H / 2) * (W / 2) * (C / 4);
  const int blocks = (int)std::min<size_t>((n4 + 255) / 256, 148 * 16);
  maxpool2_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(in), H / 2, W / 2, C / 4, n4,
                                         reinterpret_cast<float4*>(out));
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

int launch_upsample_add(suo_ctx* ctx, const float* up1, const float* low, int B, int H, int W, int C, float* out,
                        cudaStream_t s) {
  const size_t n4 = (size_t)B * H * W * (C / 4);
  const int blocks = (int)std::min<size_t>((n4 + 255) / 256, 148 * 16);
  upsample_add_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(up1), reinterpret_cast<const float4*>(low),
                                             H, W, C / 4, n4, reinterpret_cast<float4*>(out));
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}
