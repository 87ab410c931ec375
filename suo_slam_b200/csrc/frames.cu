// Device-resident glue between the three hot-path calls of a single-view frame
// (SURVEY.md §8 rows a4', a8-graph-build, f1): keypoint gating + compaction, PnP input
// normalisation, PnP acceptance test, BA graph assembly and result scatter — so that a frame
// batch goes image -> poses without the device->host->device round trips the reference makes at
// lib/object_slam.py:1100-1115 (.cpu().numpy()), :1123-1165 (per-object pnp) and :745-837
// (per-edge pybind graph construction).
#include "common.cuh"
#include "chi2.cuh"

namespace {

// ---- a4': gating (object_slam.py:1100-1115) + ordered compaction (boolean-mask indexing :1128,:1138) ----
// One warp per crop.  ys are normalised by K_bbox^-1 exactly like pnp() (object_slam.py:35-36).
__global__ void gate_compact_kernel(const float* __restrict__ uv, const float* __restrict__ cov,
                                    const float* __restrict__ kp_mask, const uint8_t* __restrict__ model_mask,
                                    const double* __restrict__ model_kps, const double* __restrict__ K_bbox, int K,
                                    float kp_var_thresh, float bbox_thresh, double* __restrict__ xs,
                                    double* __restrict__ ys, int32_t* __restrict__ counts,
                                    int32_t* __restrict__ kp_index, uint8_t* __restrict__ kp_used) {
  const int crop = blockIdx.x, lane = threadIdx.x;
  const double* Kb = K_bbox + 9 * crop;
  // inverse of a general 3x3 (adjugate), as np.linalg.inv would give up to rounding
  const double a = Kb[0], b = Kb[1], c = Kb[2], d = Kb[3], e = Kb[4], f = Kb[5], g = Kb[6], h = Kb[7], i = Kb[8];
  const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  const double i00 = (e * i - f * h) / det, i01 = (c * h - b * i) / det, i02 = (b * f - c * e) / det;
  const double i10 = (f * g - d * i) / det, i11 = (a * i - c * g) / det, i12 = (c * d - a * f) / det;
  int base = 0;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int k = k0 + lane;
    bool use = false;
    float u = 0.f, v = 0.f;
    if (k < K) {
      const size_t m = (size_t)crop * K + k;
      u = uv[2 * m]; v = uv[2 * m + 1];
      use = (kp_mask[m] > 0.3f) && model_mask[m];
      use = use && (fminf(u, v) > -bbox_thresh) && (fmaxf(u, v) < bbox_thresh);
      // std = sqrt(diag(cov)) < 2 * kp_var_thresh on both axes; NaN (negative variance) fails like numpy's comparison
      use = use && (sqrtf(cov[4 * m]) < 2.f * kp_var_thresh) && (sqrtf(cov[4 * m + 3]) < 2.f * kp_var_thresh);
      kp_used[m] = use ? 1 : 0;
    }
    const unsigned ball = __ballot_sync(0xffffffffu, use);
    if (use) {
      const int j = base + __popc(ball & ((1u << lane) - 1u));
      const size_t o = (size_t)crop * K + j;
      const size_t m = (size_t)crop * K + k;
      xs[3 * o] = model_kps[3 * m]; xs[3 * o + 1] = model_kps[3 * m + 1]; xs[3 * o + 2] = model_kps[3 * m + 2];
      const double ud = (double)u, vd = (double)v;
      // points_2d @ KinvT[:2,:2] + KinvT[2:3,:2]  ==  Kinv[:2,:2] @ p + Kinv[:2,2]
      ys[2 * o] = ud * i00 + vd * i01 + i02;
      ys[2 * o + 1] = ud * i10 + vd * i11 + i12;
      kp_index[o] = k;
    }
    base += __popc(ball);
  }
  if (lane == 0) counts[crop] = base;
}

// frame_start[f] = first crop with box_img >= f (box_img sorted ascending); frame_start[n_img] = L
__global__ void frame_ranges_kernel(const int32_t* __restrict__ box_img, int L, int n_img, int32_t* __restrict__ frame_start) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f > n_img) return;
  int lo = 0, hi = L;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (box_img[mid] < f) lo = mid + 1; else hi = mid; }
  frame_start[f] = lo;
}

// ---- BA graph assembly for single-view frames (object_slam.py:745-837 with camera fixed) ----
// One thread block per frame (thread 0 builds the ordered lists; a frame has <= ~20 objects x 41 edges).
// Vertex ids: crop c of frame f -> c + f ; camera of frame f -> frame_end + f.
// Edge slots: frame f owns [frame_start[f]*K, ...), count in edge_cnt[f].
__global__ void ba_assemble_kernel(const int32_t* __restrict__ frame_start, const int32_t* __restrict__ counts,
                                   const int32_t* __restrict__ kp_index, const double* __restrict__ xs,
                                   const float* __restrict__ uv, const float* __restrict__ cov,
                                   const double* __restrict__ K_bbox, const double* __restrict__ diameter,
                                   const double* __restrict__ T_pnp, int K,
                                   double* __restrict__ poses, uint8_t* __restrict__ fixed,
                                   int32_t* __restrict__ prob_vert, int32_t* __restrict__ vert_cnt,
                                   int32_t* __restrict__ prob_edge, int32_t* __restrict__ edge_cnt,
                                   int32_t* __restrict__ e_obj, int32_t* __restrict__ e_cam, double* __restrict__ cam_k,
                                   double* __restrict__ p, double* __restrict__ uvd, double* __restrict__ info,
                                   uint8_t* __restrict__ inliers, int32_t* __restrict__ edge_src,
                                   uint8_t* __restrict__ accepted) {
  const int f = blockIdx.x;
  const int c0 = frame_start[f], c1 = frame_start[f + 1];
  const int vbase = c0 + f, vcam = c1 + f;
  if (threadIdx.x == 0) {
    prob_vert[f] = vbase; vert_cnt[f] = c1 - c0 + 1;
    prob_edge[f] = c0 * K;
    double* Tc = poses + 12 * (size_t)vcam;
    for (int q = 0; q < 12; ++q) Tc[q] = (q == 0 || q == 5 || q == 10) ? 1.0 : 0.0;
    fixed[vcam] = 1;
  }
  __shared__ int s_ne;
  if (threadIdx.x == 0) s_ne = 0;
  __syncthreads();
  // acceptance of the PnP pose (object_slam.py:1145-1151): not identity, T[2,3] > 0.5 diameter, >= 4 points
  for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
    const double* T = T_pnp + 16 * (size_t)c;
    bool ident = true;   // np.allclose(res, eye(4)): |a-b| <= 1e-8 + 1e-5 |b|
    for (int q = 0; q < 16; ++q) {
      const double bref = (q % 5 == 0) ? 1.0 : 0.0;
      if (!(fabs(T[q] - bref) <= 1e-8 + 1e-5 * fabs(bref))) ident = false;
    }
    const bool ok = !ident && counts[c] >= 4 && T[11] > 0.5 * diameter[c];
    accepted[c] = ok ? 1 : 0;
    double* Tv = poses + 12 * (size_t)(c + f);
    for (int r = 0; r < 3; ++r) for (int q = 0; q < 4; ++q) Tv[4 * r + q] = ok ? T[4 * r + q] : ((r == q) ? 1.0 : 0.0);
    fixed[c + f] = ok ? 0 : 1;
  }
  __syncthreads();
  // edges in (crop, gated keypoint) order: thread 0 scans the accepted crops' counts (chunks of kScan crops), then one
  // thread per (crop, slot) writes its edge
  constexpr int kScan = 256;
  __shared__ int s_off[kScan + 1];
  const size_t ebase = (size_t)c0 * K;
  for (int cb = c0; cb < c1; cb += kScan) {
    const int nc = min(kScan, c1 - cb);
    if (threadIdx.x == 0) {
      int ne = s_ne;
      for (int q = 0; q < nc; ++q) { s_off[q] = ne; ne += accepted[cb + q] ? counts[cb + q] : 0; }
      s_off[nc] = ne;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nc * K; idx += blockDim.x) {
      const int q = idx / K, j = idx - q * K, c = cb + q;
      if (!accepted[c] || j >= counts[c]) continue;
      const double* Kb = K_bbox + 9 * c;
      const size_t src = (size_t)c * K + j;
      const int k = kp_index[src];
      const size_t m = (size_t)c * K + k;
      const size_t e = ebase + s_off[q] + j;
      e_obj[e] = c + f; e_cam[e] = vcam;
      cam_k[4 * e] = Kb[0]; cam_k[4 * e + 1] = Kb[4]; cam_k[4 * e + 2] = Kb[2]; cam_k[4 * e + 3] = Kb[5];   // object_slam.py:799
      p[3 * e] = xs[3 * src]; p[3 * e + 1] = xs[3 * src + 1]; p[3 * e + 2] = xs[3 * src + 2];
      uvd[2 * e] = (double)uv[2 * m]; uvd[2 * e + 1] = (double)uv[2 * m + 1];
      // information = inv(cov) (object_slam.py:825-828, no clamping), cov is the fp32 network output
      const double s00 = cov[4 * m], s01 = cov[4 * m + 1], s10 = cov[4 * m + 2], s11 = cov[4 * m + 3];
      const double det = s00 * s11 - s01 * s10;
      info[4 * e] = s11 / det; info[4 * e + 1] = -s01 / det; info[4 * e + 2] = -s10 / det; info[4 * e + 3] = s00 / det;
      inliers[e] = 1;
      edge_src[e] = (int32_t)m;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_ne = s_off[nc];
    __syncthreads();
  }
  if (threadIdx.x == 0) edge_cnt[f] = s_ne;
}

// ---- result records for the all-gather ----
// One fixed-size record per crop: everything a rank that did not process the crop needs before a step over ALL objects
// (camera-pose voting lib/object_slam.py:975-1072, joint graph :736-837): both poses, the acceptance flag, the gated
// keypoints with their covariances and the per-keypoint flags.  Layout (include/suo_b200.h, suo_record_bytes):
//   f64 T_pnp[12], f64 T_ba[12], i32 crop_id, i32 accepted, i32 n_used, i32 n_ba_inliers, f32 uv[K][2], f32 cov[K][4],
//   u8 flags[K] (bit 0 = gated in, bit 1 = BA inlier), zero padding to a multiple of 8 bytes.
__global__ void pack_records_kernel(const int32_t* __restrict__ crop_ids, int id_base, const double* __restrict__ T_pnp,
                                    const double* __restrict__ T_ba, const uint8_t* __restrict__ kp_used,
                                    const uint8_t* __restrict__ ba_inliers, const float* __restrict__ uv,
                                    const float* __restrict__ cov, int K, int rec_bytes, uint8_t* __restrict__ out) {
  const int c = blockIdx.x;
  uint8_t* r = out + (size_t)c * rec_bytes;
  double* rd = reinterpret_cast<double*>(r);
  int32_t* ri = reinterpret_cast<int32_t*>(r + 192);
  float* ruv = reinterpret_cast<float*>(r + 208);
  float* rcov = ruv + 2 * K;
  uint8_t* rfl = reinterpret_cast<uint8_t*>(rcov + 4 * K);
  const double* Tp = T_pnp + 16 * (size_t)c;
  int n_used = 0, n_inl = 0;
  for (int k = threadIdx.x; k < K; k += 32) {
    const size_t m = (size_t)c * K + k;
    const int u = kp_used ? (kp_used[m] != 0) : 0, b = ba_inliers ? (ba_inliers[m] != 0) : 0;
    n_used += u; n_inl += b;
    rfl[k] = (uint8_t)(u | (b << 1));
    ruv[2 * k] = uv ? uv[2 * m] : 0.f; ruv[2 * k + 1] = uv ? uv[2 * m + 1] : 0.f;
    for (int q = 0; q < 4; ++q) rcov[4 * k + q] = cov ? cov[4 * m + q] : 0.f;
  }
  for (int k = K + threadIdx.x; k < rec_bytes - 208 - 24 * K; k += 32) rfl[k] = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { n_used += __shfl_xor_sync(0xffffffffu, n_used, o); n_inl += __shfl_xor_sync(0xffffffffu, n_inl, o); }
  if (threadIdx.x < 12) {
    rd[threadIdx.x] = Tp[threadIdx.x];
    rd[12 + threadIdx.x] = T_ba ? T_ba[12 * (size_t)c + threadIdx.x] : ((threadIdx.x % 5 == 0) ? 1.0 : 0.0);
  }
  if (threadIdx.x == 0) {
    bool ident = true;   // identity == PnP failure (lambdatwist convention, lib/object_slam.py:38-39)
    for (int q = 0; q < 12; ++q) if (fabs(Tp[q] - ((q % 5 == 0) ? 1.0 : 0.0)) > 1e-8 + ((q % 5 == 0) ? 1e-5 : 0.0)) ident = false;
    ri[0] = crop_ids ? crop_ids[c] : id_base + c;
    ri[1] = ident ? 0 : 1; ri[2] = n_used; ri[3] = n_inl;
  }
}

// scatter BA results back to per-crop / per-keypoint arrays
__global__ void ba_scatter_kernel(const int32_t* __restrict__ frame_start, const int32_t* __restrict__ edge_cnt,
                                  const int32_t* __restrict__ edge_src, const uint8_t* __restrict__ inliers,
                                  const double* __restrict__ poses, const uint8_t* __restrict__ accepted, int K,
                                  double* __restrict__ T_ba, uint8_t* __restrict__ ba_inliers) {
  const int f = blockIdx.x;
  const int c0 = frame_start[f], c1 = frame_start[f + 1];
  if (ba_inliers) {
    for (int i = threadIdx.x; i < (c1 - c0) * K; i += blockDim.x) ba_inliers[(size_t)c0 * K + i] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < edge_cnt[f]; e += blockDim.x) {
      const size_t ge = (size_t)c0 * K + e;
      ba_inliers[edge_src[ge]] = inliers[ge];
    }
  }
  if (T_ba) {
    for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
      const double* Tv = poses + 12 * (size_t)(c + f);
      for (int q = 0; q < 12; ++q) T_ba[12 * (size_t)c + q] = accepted[c] ? Tv[q] : ((q == 0 || q == 5 || q == 10) ? 1.0 : 0.0);
    }
  }
}

// ---- f1: chi2 inlier counting of keypoint detections under pose hypotheses ----
// The inner loop shared by ObjectSLAM.__estimate_camera_pose (lib/object_slam.py:1030-1066: every camera-pose
// hypothesis is scored by the keypoints it explains over ALL objects) and __maybe_reinit_objects (:645-680: PnP pose
// vs current estimate over the last 15 views).  One warp per (pose, detection) pair:
//   p_C = R p_O + t (utils.transform_pts), uvw = K p_C, keep w > 0, res = uv - uvw.xy / w,
//   cov diagonal floored at 1e-4 (:1054, :669) then inverted — or 1 / manual_kp_std^2 when there is no network
//   covariance (:1059-1061) — chi2 = res^T inf res, count chi2 <= gate.
// The reference inverts the float32 covariance in float32 (np.linalg.inv); here the float32 inputs are widened and
// the 2x2 inverse is closed-form FP64: identical counts unless a chi2 sits within ~1e-6 relative of the gate.
__global__ void chi2_count_kernel(const double* __restrict__ T, const int32_t* __restrict__ pair_det,
                                  const int32_t* __restrict__ det_off, const double* __restrict__ model_kp,
                                  const double* __restrict__ Kmat, const float* __restrict__ uv, const float* __restrict__ cov,
                                  const uint8_t* __restrict__ use, double inv_manual_var, double gate, int n_pairs,
                                  int32_t* __restrict__ counts) {
  const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (pair >= n_pairs) return;
  const double* P = T + 12 * (size_t)pair;
  const int d = pair_det[pair];
  const double* Kd = Kmat + 9 * (size_t)d;
  int cnt = 0;
  for (int r = det_off[d] + lane; r < det_off[d + 1]; r += 32) {
    if (use && !use[r]) continue;
    cnt += chi2_inlier(P, Kd, model_kp + 3 * (size_t)r, uv[2 * r], uv[2 * r + 1], cov ? cov + 4 * (size_t)r : nullptr, inv_manual_var, gate) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) counts[pair] = cnt;
}

}  // namespace

int launch_gate_compact(suo_ctx* ctx, const float* uv, const float* cov, const float* kp_mask, const uint8_t* model_mask,
                        const double* model_kps, const double* K_bbox, int L, int K, float kp_var_thresh, float bbox_thresh,
                        double* xs, double* ys, int32_t* counts, int32_t* kp_index, uint8_t* kp_used, cudaStream_t s) {
  gate_compact_kernel<<<L, 32, 0, s>>>(uv, cov, kp_mask, model_mask, model_kps, K_bbox, K, kp_var_thresh, bbox_thresh, xs, ys,
                                       counts, kp_index, kp_used);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

int launch_frame_ranges(suo_ctx* ctx, const int32_t* box_img, int L, int n_img, int32_t* frame_start, cudaStream_t s) {
  frame_ranges_kernel<<<(n_img + 1 + 127) / 128, 128, 0, s>>>(box_img, L, n_img, frame_start);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

int launch_ba_assemble(suo_ctx* ctx, int n_img, const int32_t* frame_start, const int32_t* counts, const int32_t* kp_index,
                       const double* xs, const float* uv, const float* cov, const double* K_bbox, const double* diameter,
                       const double* T_pnp, int K, double* poses, uint8_t* fixed, int32_t* prob_vert, int32_t* vert_cnt,
                       int32_t* prob_edge, int32_t* edge_cnt, int32_t* e_obj, int32_t* e_cam, double* cam_k, double* p,
                       double* uvd, double* info, uint8_t* inliers, int32_t* edge_src, uint8_t* accepted, cudaStream_t s) {
  ba_assemble_kernel<<<n_img, 128, 0, s>>>(frame_start, counts, kp_index, xs, uv, cov, K_bbox, diameter, T_pnp, K, poses, fixed,
                                          prob_vert, vert_cnt, prob_edge, edge_cnt, e_obj, e_cam, cam_k, p, uvd, info, inliers,
                                          edge_src, accepted);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

int launch_ba_scatter(suo_ctx* ctx, int n_img, const int32_t* frame_start, const int32_t* edge_cnt, const int32_t* edge_src,
                      const uint8_t* inliers, const double* poses, const uint8_t* accepted, int K, double* T_ba,
                      uint8_t* ba_inliers, cudaStream_t s) {
  ba_scatter_kernel<<<n_img, 64, 0, s>>>(frame_start, edge_cnt, edge_src, inliers, poses, accepted, K, T_ba, ba_inliers);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

int launch_chi2_counts(suo_ctx* ctx, int n_pairs, const double* T, const int32_t* pair_det, const int32_t* det_off,
                       const double* model_kp, const double* K, const float* uv, const float* cov, const uint8_t* use,
                       double manual_kp_std, double gate, int32_t* counts, cudaStream_t s) {
  if (n_pairs <= 0) return SUO_OK;
  chi2_count_kernel<<<(n_pairs + 7) / 8, 256, 0, s>>>(T, pair_det, det_off, model_kp, K, uv, cov, use,
                                                      1.0 / (manual_kp_std * manual_kp_std), gate, n_pairs, counts);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

int launch_pack_records(suo_ctx* ctx, const int32_t* crop_ids, int id_base, const double* T_pnp, const double* T_ba,
                        const uint8_t* kp_used, const uint8_t* ba_inliers, const float* uv, const float* cov, int L, int K,
                        int rec_bytes, uint8_t* out, cudaStream_t s) {
  if (L <= 0) return SUO_OK;
  pack_records_kernel<<<L, 32, 0, s>>>(crop_ids, id_base, T_pnp, T_ba, kp_used, ba_inliers, uv, cov, K, rec_bytes, out);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}
