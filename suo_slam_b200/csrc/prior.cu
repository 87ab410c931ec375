// Prior-detection heat maps rendered on the device.
//
// Replaces utils.make_prior_kp_input / draw_gaussian_2d / gaussian_2d (reference lib/utils/utils.py:356-411),
// which ObjectSLAM calls per symmetric object (lib/object_slam.py:513-514) to build the [41,256,256] prior planes
// on the CPU (8 ms / object) before copying 10.75 MB / crop to the GPU.  Here the planes are a pure function of
// (prior_uv[K,2], mask[K]): each valid keypoint stamps the reference's fixed 91x91 Gaussian (sigma 15 ->
// tmpSize 45; cv2.GaussianBlur(delta 91x91, ksize 91, sigma 0) / max, prior_gauss_table.inc) with its centre at
// (round(u_px), round(v_px)); the pasted window is 90x90 because the slice end is exclusive (utils.py:369-384).
//   * render_priors_planes_kernel writes the reference layout [L,K,R,R] (drop-in for make_prior_kp_input);
//   * render_priors_nhwc_kernel writes channels 3..47 of the network's NHWC input directly, so a forward with
//     keypoint priors never materialises or transfers the planes.
// Pixel maths is done with explicit FP32 round-to-nearest operations in the reference's order (the reference gets
// FP32 there because prior_uv is a float32 array, object_slam.py:510), and rounding is half-to-even like Python's
// round(): the stamp position is bit-exact.
#include "common.cuh"

namespace {

__device__ const unsigned int kPriorQ[46 * 46] = {
#include "prior_gauss_table.inc"
};

constexpr int HALF = 45;    // tmpSize = ceil(3 * sigma), sigma = 15
constexpr int WIN = 90;     // pasted window (exclusive slice end)

// Upper-left corner of keypoint k's window, or false if nothing is drawn.
__device__ __forceinline__ bool stamp_origin(float u, float v, unsigned char m, int vh, int vw, int ndc, int& ulx, int& uly) {
  if (!m || !isfinite(u) || !isfinite(v)) return false;                       // utils.py:402
  if (ndc) {                                                                  // utils.py:404-406
    const float cu = fminf(fmaxf(u, -1.f), 1.f), cv = fminf(fmaxf(v, -1.f), 1.f);
    u = __fsub_rn(__fadd_rn(__fdiv_rn(__fmul_rn(cu, (float)vw), 2.f), (float)vw / 2.f), 0.5f);
    v = __fsub_rn((float)vh - 0.5f, __fadd_rn(__fdiv_rn(__fmul_rn(cv, (float)vh), 2.f), (float)vh / 2.f));
  }
  const int px = (int)rintf(u), py = (int)rintf(v);                           // int(round(.)), half to even
  ulx = px - HALF; uly = py - HALF;
  const int brx = px + HALF, bry = py + HALF;
  if (ulx > vw || uly > vh || brx < 1 || bry < 1) return false;               // utils.py:372-373
  return true;
}

__device__ __forceinline__ float stamp_value(int ulx, int uly, int y, int x) {
  const int i = y - uly, j = x - ulx;
  if (i < 0 || j < 0 || i >= WIN || j >= WIN) return 0.f;
  const int qi = min(i, 2 * HALF - i), qj = min(j, 2 * HALF - j);
  return __uint_as_float(kPriorQ[qi * 46 + qj]);
}

// grid (ceil(R*R/256), K, L): one thread per pixel of one plane
__global__ void __launch_bounds__(256)
render_priors_planes_kernel(const float* __restrict__ uv, const unsigned char* __restrict__ mask, int K, int vh, int vw,
                            int ndc, float* __restrict__ out) {
  const int k = blockIdx.y, crop = blockIdx.z;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= vh * vw) return;
  int ulx, uly;
  const size_t lk = (size_t)crop * K + k;
  float val = 0.f;
  if (stamp_origin(uv[2 * lk], uv[2 * lk + 1], mask[lk], vh, vw, ndc, ulx, uly)) val = stamp_value(ulx, uly, pix / vw, pix % vw);
  out[lk * vh * vw + pix] = val;
}

// grid (R*R*12/256, L): one thread per (pixel, 4 channels) of the 48-channel NHWC input; channels 0..2 (RGB) are
// left to roi_align_kernel, 3..3+K-1 are the prior planes, the rest is zero padding.
__global__ void __launch_bounds__(256)
render_priors_nhwc_kernel(const float* __restrict__ uv, const unsigned char* __restrict__ mask, int K, int R,
                          float* __restrict__ out) {
  __shared__ int s_ulx[48], s_uly[48];
  __shared__ unsigned char s_ok[48];
  const int crop = blockIdx.y;
  if (threadIdx.x < 48) {
    const int k = threadIdx.x;
    int ulx = 0, uly = 0;
    bool ok = false;
    if (k < K) { const size_t lk = (size_t)crop * K + k; ok = stamp_origin(uv[2 * lk], uv[2 * lk + 1], mask[lk], R, R, 1, ulx, uly); }
    s_ulx[k] = ulx; s_uly[k] = uly; s_ok[k] = ok;
  }
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R * R * 12) return;
  const int pix = t / 12, q = t % 12, y = pix / R, x = pix % R;
  float v[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int k = 4 * q + c - 3;
    v[c] = (k >= 0 && k < K && s_ok[k]) ? stamp_value(s_ulx[k], s_uly[k], y, x) : 0.f;
  }
  float* o = out + ((size_t)crop * R * R + pix) * 48 + 4 * q;
  if (q == 0) o[3] = v[3];
  else *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
}

}  // namespace

int launch_render_priors_planes(suo_ctx* ctx, const float* uv, const uint8_t* mask, int L, int K, int vh, int vw, int ndc,
                                float* out, cudaStream_t s) {
  if (L <= 0 || K <= 0 || K > 65535 || L > 65535 || vh <= 0 || vw <= 0) { ctx->set_error("render_priors: bad shape", __FILE__, __LINE__); return SUO_E_INVALID; }
  dim3 g((vh * vw + 255) / 256, K, L);
  render_priors_planes_kernel<<<g, 256, 0, s>>>(uv, mask, K, vh, vw, ndc, out);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

int launch_render_priors_nhwc(suo_ctx* ctx, const float* uv, const uint8_t* mask, int L, int K, int R, float* out, int out_c,
                              cudaStream_t s) {
  if (out_c != 48 || K > 45 || L <= 0 || L > 65535) { ctx->set_error("render_priors_nhwc: the prior input layout has 48 channels", __FILE__, __LINE__); return SUO_E_INVALID; }
  dim3 g((R * R * 12 + 255) / 256, L);
  render_priors_nhwc_kernel<<<g, 256, 0, s>>>(uv, mask, K, R, out);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}
