// Device-resident SLAM-mode frame: everything ObjectSLAM.process_view does between its three downward calls
// when single_view_mode is off (reference lib/object_slam.py:393-421) — per-crop NDC camera matrices (utils.fix_K_for_bbox_ndc,
// lib/utils/utils.py:416-429), the camera-pose vote over the PnP poses of the non-symmetric objects (__estimate_camera_pose,
// :975-1072), the prior keypoints of the symmetric objects projected from the map (:486-514), initialisation of unmapped objects
// (:577-592), the re-initialisation test over the last views (__maybe_reinit_objects, :595-697) and the graph of the curr_only
// camera solve (optimize(curr_only=True), :716-837) — so that one frame goes image -> camera pose without leaving the GPU:
//   forward(non-symmetric crops) -> gate -> PnP -> VOTE -> PRIOR UV -> forward(symmetric crops, priors stamped on the device)
//   -> gate -> PnP -> INIT -> RE-INIT -> curr_only LM (ba.cu)
// The reference does all of this in NumPy on the host, with float32 staging of the map poses (:1005-1008, :619-633) that is
// reproduced here.  Every kernel below is a single small block: the stage is latency-bound glue, the point is the missing
// device -> host -> device round trips.
#include "common.cuh"
#include "chi2.cuh"

namespace {

__device__ __forceinline__ void mat34_to44_mul(const double* A, const double* B, double* C) {   // C = A B for [R|t] 3x4 blocks of SE(3) 4x4
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      C[4 * r + c] = A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c] + A[4 * r + 2] * B[8 + c] + (c == 3 ? A[4 * r + 3] : 0.0);
  }
}
__device__ __forceinline__ void se3_inv34(const double* T, double* O) {          // utils.invert_SE3: [R^T | -R^T t]
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) O[4 * r + c] = T[4 * c + r];
    O[4 * r + 3] = -(T[r] * T[3] + T[4 + r] * T[7] + T[8 + r] * T[11]);
  }
}
__device__ __forceinline__ void round_f32(const double* T, double* O) {           // np.float32 staging of a pose
#pragma unroll
  for (int q = 0; q < 12; ++q) O[q] = (double)(float)T[q];
}
// acceptance of a PnP pose (object_slam.py:1145-1151): not the identity "failure" pose, T[2,3] > 0.5 diameter, >= 4 points
__device__ __forceinline__ bool pnp_accepted(const double* T16, int count, double diameter) {
  bool ident = true;
  for (int q = 0; q < 16; ++q) {
    const double bref = (q % 5 == 0) ? 1.0 : 0.0;
    if (!(fabs(T16[q] - bref) <= 1e-8 + 1e-5 * fabs(bref))) ident = false;
  }
  return !ident && count >= 4 && T16[11] > 0.5 * diameter;
}

// K_bbox = S T K per crop; `raw` is the FP64 product (what the prior projection uses, :500), `f32` the same rounded through float32
// (what __run_kp_model keeps in K_bbox_np and hands to pnp() and the edges, :1082-1086,1140)
__global__ void slam_kbbox_kernel(const double* __restrict__ Kc, const float* __restrict__ boxes, int L, double* __restrict__ raw,
                                  double* __restrict__ f32) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= L) return;
  // the boxes are float32 and so is the reference's scalar arithmetic on them (utils.fix_K_for_bbox_ndc on a float32 bbox, NumPy >= 2:
  // w = x2 - x1 and 2.0 / w in float32, IEEE-rounded — __fsub_rn / __fdiv_rn keep the compiler from contracting or approximating them)
  const double x1 = boxes[4 * c], y1 = boxes[4 * c + 1];
  const float w = __fsub_rn(boxes[4 * c + 2], boxes[4 * c]), h = __fsub_rn(boxes[4 * c + 3], boxes[4 * c + 1]);
  const double sx = (double)__fdiv_rn(2.0f, w), sy = (double)__fdiv_rn(-2.0f, h);
  // S T = [[sx, 0, sx * (-x1) - 1], [0, sy, sy * (-y1) + 1], [0, 0, 1]]
  const double M[9] = {sx, 0.0, sx * (-x1) + (-1.0), 0.0, sy, sy * (-y1) + 1.0, 0.0, 0.0, 1.0};
  for (int r = 0; r < 3; ++r)
    for (int q = 0; q < 3; ++q) {
      const double v = M[3 * r] * Kc[q] + M[3 * r + 1] * Kc[3 + q] + M[3 * r + 2] * Kc[6 + q];
      raw[9 * c + 3 * r + q] = v;
      f32[9 * c + 3 * r + q] = (double)(float)v;
    }
}

constexpr int SLAM_MAX_CROPS = 128;

struct VoteArgs {
  int n1, K, first_view;
  const double* T_pnp; const int32_t* counts; const int32_t* kp_index; const double* xs; const float* uv; const float* cov;
  const double* Kb; const double* diameter; const uint8_t* map_valid; const double* T_OtoG;
  double inv_manual_var, gate; int min_inliers;
  double* T_GtoC; int32_t* status;      // status[0] = camera pose known, [1] = votes of the winner, [2] = number of hypotheses
};

// __estimate_camera_pose (:975-1072): every non-symmetric object with a PnP pose and a map pose proposes T_GtoC = T_OtoC_pnp inv(T_OtoG);
// a hypothesis is scored by the keypoints it explains (chi2 <= 5.991) over ALL voting objects; the first best one with >= 4 wins.
__global__ void __launch_bounds__(256) slam_vote_kernel(const VoteArgs a) {
  __shared__ double hyp[SLAM_MAX_CROPS][12];
  __shared__ int voter[SLAM_MAX_CROPS];
  __shared__ int cnt[SLAM_MAX_CROPS];
  __shared__ int nv;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (a.first_view) {         // the first view defines the world frame (:408-411)
    if (tid < 12) a.T_GtoC[tid] = (tid % 5 == 0) ? 1.0 : 0.0;
    if (tid == 0) { a.status[0] = 1; a.status[1] = 0; a.status[2] = 0; }
    return;
  }
  if (tid == 0) {
    int n = 0;
    for (int c = 0; c < a.n1 && n < SLAM_MAX_CROPS; ++c)
      if (a.map_valid[c] && pnp_accepted(a.T_pnp + 16 * (size_t)c, a.counts[c], a.diameter[c])) voter[n++] = c;
    nv = n;
  }
  __syncthreads();
  for (int i = tid; i < nv; i += blockDim.x) {
    double inv[12];
    se3_inv34(a.T_OtoG + 12 * (size_t)voter[i], inv);
    mat34_to44_mul(a.T_pnp + 16 * (size_t)voter[i], inv, hyp[i]);
    cnt[i] = 0;
  }
  __syncthreads();
  for (int pair = warp; pair < nv * nv; pair += blockDim.x / 32) {
    const int i = pair / nv, j = pair - i * nv, cj = voter[j];
    double Tj[12], P[12];
    round_f32(a.T_OtoG + 12 * (size_t)cj, Tj);
    mat34_to44_mul(hyp[i], Tj, P);
    int n = 0;
    for (int t = lane; t < a.counts[cj]; t += 32) {
      const size_t row = (size_t)cj * a.K + t, m = (size_t)cj * a.K + a.kp_index[row];
      n += chi2_inlier(P, a.Kb + 9 * (size_t)cj, a.xs + 3 * row, a.uv[2 * m], a.uv[2 * m + 1], a.cov ? a.cov + 4 * m : nullptr, a.inv_manual_var, a.gate) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) atomicAdd(&cnt[i], n);
  }
  __syncthreads();
  if (tid == 0) {
    int best = -1, best_n = -1;
    for (int i = 0; i < nv; ++i)
      if (cnt[i] >= a.min_inliers && cnt[i] > best_n) { best = i; best_n = cnt[i]; }      // :1068-1070
    a.status[0] = best >= 0; a.status[1] = best_n; a.status[2] = nv;
    for (int q = 0; q < 12; ++q) a.T_GtoC[q] = best >= 0 ? hyp[best][q] : ((q % 5 == 0) ? 1.0 : 0.0);
  }
}

// prior keypoints of the symmetric crops (:486-514): project the masked model keypoints with T_GtoC T_OtoG through the crop's
// (unrounded) NDC camera matrix; all depths must be positive, else the crop gets no prior (zero planes)
__global__ void slam_prior_uv_kernel(int n1, int K, const int32_t* __restrict__ status, const double* __restrict__ T_GtoC,
                                     const uint8_t* __restrict__ map_valid, const double* __restrict__ T_OtoG,
                                     const double* __restrict__ model_kps, const uint8_t* __restrict__ model_mask,
                                     const double* __restrict__ Kb_raw, float* __restrict__ prior_uv, uint8_t* __restrict__ prior_mask) {
  const int c = n1 + blockIdx.x;
  const bool on = status[0] && map_valid[c];
  double P[12];
  if (on) mat34_to44_mul(T_GtoC, T_OtoG + 12 * (size_t)c, P);
  const double* Kd = Kb_raw + 9 * (size_t)c;
  int bad = 0;
  float u[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  bool mk[2] = {false, false};
  for (int it = 0, k = threadIdx.x; it < 2; ++it, k += blockDim.x) {      // K <= 2 * blockDim.x (checked by the launcher)
    if (!on || k >= K) continue;
    const size_t m = (size_t)c * K + k;
    mk[it] = model_mask[m] != 0;
    if (!mk[it]) continue;
    const double x = model_kps[3 * m], y = model_kps[3 * m + 1], z = model_kps[3 * m + 2];
    const double pc0 = x * P[0] + y * P[1] + z * P[2] + P[3];            // utils.transform_pts: pts @ R^T + t
    const double pc1 = x * P[4] + y * P[5] + z * P[6] + P[7];
    const double pc2 = x * P[8] + y * P[9] + z * P[10] + P[11];
    const double d = pc0 * Kd[6] + pc1 * Kd[7] + pc2 * Kd[8];
    if (!(d > 0)) bad = 1;
    u[it][0] = (float)((pc0 * Kd[0] + pc1 * Kd[1] + pc2 * Kd[2]) / d);
    u[it][1] = (float)((pc0 * Kd[3] + pc1 * Kd[4] + pc2 * Kd[5]) / d);
  }
  const int any_bad = __syncthreads_or(bad);
  for (int it = 0, k = threadIdx.x; it < 2; ++it, k += blockDim.x) {
    if (k >= K) continue;
    const size_t m = (size_t)c * K + k;
    const bool w = on && !any_bad && mk[it];
    prior_uv[2 * m] = w ? u[it][0] : 0.f; prior_uv[2 * m + 1] = w ? u[it][1] : 0.f;
    prior_mask[m] = w ? 1 : 0;
  }
}

// symmetric crops are only processed when the camera pose could be recovered (:413-418): otherwise they leave no detection
__global__ void slam_drop_group_kernel(int n1, int L, int K, const int32_t* __restrict__ status, int32_t* __restrict__ counts,
                                       uint8_t* __restrict__ kp_used, double* __restrict__ T_pnp) {
  if (status[0]) return;
  const int c = n1 + blockIdx.x;
  if (c >= L) return;
  for (int k = threadIdx.x; k < K; k += blockDim.x) kp_used[(size_t)c * K + k] = 0;
  if (threadIdx.x < 16) T_pnp[16 * (size_t)c + threadIdx.x] = (threadIdx.x % 5 == 0) ? 1.0 : 0.0;
  if (threadIdx.x == 0) counts[c] = 0;
}

struct MapArgs {
  int L, K, n_views, n_hist;
  const int32_t* status; const double* T_GtoC;
  const double* T_pnp; const int32_t* counts; const int32_t* kp_index; const double* xs; const float* uv; const float* cov;
  const double* Kb; const double* diameter;
  const uint8_t* map_valid_in; const double* T_OtoG_in;
  uint8_t* map_valid; double* T_OtoG;            // outputs (updated map)
  const int32_t* hist_crop; const double* hist_T_GtoC; const double* hist_K; const int32_t* hist_off;
  const double* hist_model_kp; const float* hist_uv; const float* hist_cov;
  double inv_manual_var, gate;
  int32_t* rcounts;      // [L,2]: keypoints explained by the PnP pose / by the map pose over the checked views
  uint8_t* reinit;       // [L]
  int init_from;         // first crop whose object may be initialised in this view (n_nonsym after a failed vote, :566-575 returns before :577-592)
};

// copy the map, then initialise unmapped objects that got a PnP pose in this view: T_OtoG = inv(T_GtoC) T_OtoC (:541-556, :577-592)
__global__ void slam_init_objects_kernel(const MapArgs a) {
  for (int c = threadIdx.x; c < a.L; c += blockDim.x) {
    bool valid = a.map_valid_in[c] != 0;
    double T[12];
    for (int q = 0; q < 12; ++q) T[q] = valid ? a.T_OtoG_in[12 * (size_t)c + q] : ((q % 5 == 0) ? 1.0 : 0.0);
    if (!valid && c >= a.init_from && a.status[0] && pnp_accepted(a.T_pnp + 16 * (size_t)c, a.counts[c], a.diameter[c])) {
      double inv[12];
      se3_inv34(a.T_GtoC, inv);
      mat34_to44_mul(inv, a.T_pnp + 16 * (size_t)c, T);
      valid = true;
    }
    for (int q = 0; q < 12; ++q) a.T_OtoG[12 * (size_t)c + q] = T[q];
    a.map_valid[c] = valid ? 1 : 0;
    a.rcounts[2 * c] = 0; a.rcounts[2 * c + 1] = 0; a.reinit[c] = 0;
  }
}

// __maybe_reinit_objects (:595-697), counting: one warp per (object, view) — history detections first, then this view's own
__global__ void slam_reinit_count_kernel(const MapArgs a) {
  if (a.n_views < 2 || !a.status[0]) return;
  const int b = blockIdx.x, lane = threadIdx.x;
  const bool hist = b < a.n_hist;
  const int c = hist ? a.hist_crop[b] : b - a.n_hist;
  if (c < 0 || c >= a.L || !a.map_valid[c] || !pnp_accepted(a.T_pnp + 16 * (size_t)c, a.counts[c], a.diameter[c])) return;
  double inv[12], TpG[12], Test[12], G[12], Ppnp[12], Pest[12];
  se3_inv34(a.T_GtoC, inv);
  mat34_to44_mul(inv, a.T_pnp + 16 * (size_t)c, TpG);                     // Ts_OtoG_pnp (:627)
  round_f32(a.T_OtoG + 12 * (size_t)c, Test);                             // Ts_OtoG_estim, float32 (:619-622)
  round_f32(hist ? a.hist_T_GtoC + 12 * (size_t)b : a.T_GtoC, G);         // Ts_GtoCi, float32 (:630-633)
  mat34_to44_mul(G, TpG, Ppnp);
  mat34_to44_mul(G, Test, Pest);
  int np = 0, ne = 0;
  if (hist) {
    const double* Kd = a.hist_K + 9 * (size_t)b;
    for (int r = a.hist_off[b] + lane; r < a.hist_off[b + 1]; r += 32) {
      const float* cv = a.hist_cov ? a.hist_cov + 4 * (size_t)r : nullptr;
      np += chi2_inlier(Ppnp, Kd, a.hist_model_kp + 3 * (size_t)r, a.hist_uv[2 * r], a.hist_uv[2 * r + 1], cv, a.inv_manual_var, a.gate) ? 1 : 0;
      ne += chi2_inlier(Pest, Kd, a.hist_model_kp + 3 * (size_t)r, a.hist_uv[2 * r], a.hist_uv[2 * r + 1], cv, a.inv_manual_var, a.gate) ? 1 : 0;
    }
  } else {
    const double* Kd = a.Kb + 9 * (size_t)c;
    for (int t = lane; t < a.counts[c]; t += 32) {
      const size_t row = (size_t)c * a.K + t, m = (size_t)c * a.K + a.kp_index[row];
      const float* cv = a.cov ? a.cov + 4 * m : nullptr;
      np += chi2_inlier(Ppnp, Kd, a.xs + 3 * row, a.uv[2 * m], a.uv[2 * m + 1], cv, a.inv_manual_var, a.gate) ? 1 : 0;
      ne += chi2_inlier(Pest, Kd, a.xs + 3 * row, a.uv[2 * m], a.uv[2 * m + 1], cv, a.inv_manual_var, a.gate) ? 1 : 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { np += __shfl_xor_sync(0xffffffffu, np, o); ne += __shfl_xor_sync(0xffffffffu, ne, o); }
  if (lane == 0) { atomicAdd(&a.rcounts[2 * c], np); atomicAdd(&a.rcounts[2 * c + 1], ne); }
}

// ... and the decision: re-initialise when the PnP pose explains >= 3 keypoints and more than three times what the map pose explains (:683-687)
__global__ void slam_reinit_apply_kernel(const MapArgs a) {
  if (a.n_views < 2 || !a.status[0]) return;
  for (int c = threadIdx.x; c < a.L; c += blockDim.x) {
    const int np = a.rcounts[2 * c], ne = a.rcounts[2 * c + 1];
    if (np >= 3 && np > 3 * ne) {
      double inv[12], TpG[12];
      se3_inv34(a.T_GtoC, inv);
      mat34_to44_mul(inv, a.T_pnp + 16 * (size_t)c, TpG);
      for (int q = 0; q < 12; ++q) a.T_OtoG[12 * (size_t)c + q] = TpG[q];
      a.reinit[c] = 1;
    }
  }
}

struct CurrBaArgs {
  int L, K;
  const int32_t* status; const double* T_GtoC; const uint8_t* map_valid; const double* T_OtoG;
  const int32_t* counts; const int32_t* kp_index; const double* xs; const float* uv; const float* cov; const double* Kb;
  double* poses; uint8_t* fixed; int32_t* prob_vert; int32_t* vert_cnt; int32_t* prob_edge; int32_t* edge_cnt;
  int32_t* e_obj; int32_t* e_cam; double* cam_k; double* p; double* uvd; double* info; uint8_t* inliers; int32_t* edge_src;
};

// graph of optimize(curr_only=True) (:716-837): one free camera vertex; one EdgeSE3ProjectFromFixedObject per gated keypoint of every
// mapped object of the view (p_inG = T_OtoG p_O, types_object_slam.h:66-79), in detection order; fewer than 3 measurements: skipped
__global__ void slam_ba_assemble_kernel(const CurrBaArgs a) {
  __shared__ int off[SLAM_MAX_CROPS + 1];
  if (threadIdx.x == 0) {
    int ne = 0;
    for (int c = 0; c < a.L; ++c) { off[c] = ne; ne += (a.status[0] && a.map_valid[c]) ? a.counts[c] : 0; }
    off[a.L] = ne;
    a.prob_vert[0] = 0; a.vert_cnt[0] = 1; a.prob_edge[0] = 0;
    a.edge_cnt[0] = ne < 3 ? 0 : ne;
    a.fixed[0] = 0;
    for (int q = 0; q < 12; ++q) a.poses[q] = a.T_GtoC[q];
  }
  __syncthreads();
  if (off[a.L] < 3) return;
  for (int idx = threadIdx.x; idx < a.L * a.K; idx += blockDim.x) {
    const int c = idx / a.K, j = idx - c * a.K;
    if (!(a.status[0] && a.map_valid[c]) || j >= a.counts[c]) continue;
    const size_t src = (size_t)c * a.K + j, m = (size_t)c * a.K + a.kp_index[src], e = off[c] + j;
    const double* T = a.T_OtoG + 12 * (size_t)c;
    const double* Kd = a.Kb + 9 * (size_t)c;
    const double x = a.xs[3 * src], y = a.xs[3 * src + 1], z = a.xs[3 * src + 2];
    a.e_obj[e] = -1; a.e_cam[e] = 0;
    a.cam_k[4 * e] = Kd[0]; a.cam_k[4 * e + 1] = Kd[4]; a.cam_k[4 * e + 2] = Kd[2]; a.cam_k[4 * e + 3] = Kd[5];
    a.p[3 * e] = T[0] * x + T[1] * y + T[2] * z + T[3];
    a.p[3 * e + 1] = T[4] * x + T[5] * y + T[6] * z + T[7];
    a.p[3 * e + 2] = T[8] * x + T[9] * y + T[10] * z + T[11];
    a.uvd[2 * e] = (double)a.uv[2 * m]; a.uvd[2 * e + 1] = (double)a.uv[2 * m + 1];
    const double s00 = a.cov[4 * m], s01 = a.cov[4 * m + 1], s10 = a.cov[4 * m + 2], s11 = a.cov[4 * m + 3];
    const double det = s00 * s11 - s01 * s10;
    a.info[4 * e] = s11 / det; a.info[4 * e + 1] = -s01 / det; a.info[4 * e + 2] = -s10 / det; a.info[4 * e + 3] = s00 / det;
    a.inliers[e] = 1;
    a.edge_src[e] = (int32_t)m;
  }
}

__global__ void slam_ba_scatter_kernel(int L, int K, const int32_t* __restrict__ edge_cnt, const int32_t* __restrict__ edge_src,
                                       const uint8_t* __restrict__ inliers, const double* __restrict__ poses, const int32_t* __restrict__ ba_stats,
                                       double* __restrict__ T_GtoC, uint8_t* __restrict__ ba_inliers, int32_t* __restrict__ status) {
  for (int i = threadIdx.x; i < L * K; i += blockDim.x) ba_inliers[i] = 0;
  __syncthreads();
  const int ne = edge_cnt[0];
  int n_in = 0;
  for (int e = threadIdx.x; e < ne; e += blockDim.x) { ba_inliers[edge_src[e]] = inliers[e]; n_in += inliers[e]; }
  if (threadIdx.x < 12 && ne > 0) T_GtoC[threadIdx.x] = poses[threadIdx.x];
  atomicAdd(&status[5], n_in);
  if (threadIdx.x == 0) { status[3] = ne; status[4] = ne > 0 ? ba_stats[2] : 0; }
}

}  // namespace

int launch_slam_kbbox(suo_ctx* ctx, const double* Kc, const float* boxes, int L, double* raw, double* f32, cudaStream_t s) {
  slam_kbbox_kernel<<<(L + 127) / 128, 128, 0, s>>>(Kc, boxes, L, raw, f32);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

int launch_slam_vote(suo_ctx* ctx, int n1, int K, int first_view, const double* T_pnp, const int32_t* counts, const int32_t* kp_index,
                     const double* xs, const float* uv, const float* cov, const double* Kb, const double* diameter, const uint8_t* map_valid,
                     const double* T_OtoG, double manual_kp_std, double gate, double* T_GtoC, int32_t* status, cudaStream_t s) {
  VoteArgs a{n1, K, first_view, T_pnp, counts, kp_index, xs, uv, cov, Kb, diameter, map_valid, T_OtoG, 1.0 / (manual_kp_std * manual_kp_std), gate, 4, T_GtoC, status};
  slam_vote_kernel<<<1, 256, 0, s>>>(a);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

int launch_slam_prior_uv(suo_ctx* ctx, int n1, int L, int K, const int32_t* status, const double* T_GtoC, const uint8_t* map_valid, const double* T_OtoG,
                         const double* model_kps, const uint8_t* model_mask, const double* Kb_raw, float* prior_uv, uint8_t* prior_mask, cudaStream_t s) {
  if (L <= n1) return SUO_OK;
  if (K > 128) { ctx->set_error("suo_slam_frame: more than 128 keypoints", __FILE__, __LINE__); return SUO_E_INVALID; }
  slam_prior_uv_kernel<<<L - n1, 64, 0, s>>>(n1, K, status, T_GtoC, map_valid, T_OtoG, model_kps, model_mask, Kb_raw, prior_uv, prior_mask);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

int launch_slam_drop_group(suo_ctx* ctx, int n1, int L, int K, const int32_t* status, int32_t* counts, uint8_t* kp_used, double* T_pnp, cudaStream_t s) {
  if (L <= n1) return SUO_OK;
  slam_drop_group_kernel<<<L - n1, 64, 0, s>>>(n1, L, K, status, counts, kp_used, T_pnp);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

int launch_slam_map_update(suo_ctx* ctx, int L, int K, int n_views, int n_hist, const int32_t* status, const double* T_GtoC, const double* T_pnp,
                           const int32_t* counts, const int32_t* kp_index, const double* xs, const float* uv, const float* cov, const double* Kb,
                           const double* diameter, const uint8_t* map_valid_in, const double* T_OtoG_in, uint8_t* map_valid, double* T_OtoG,
                           const int32_t* hist_crop, const double* hist_T_GtoC, const double* hist_K, const int32_t* hist_off,
                           const double* hist_model_kp, const float* hist_uv, const float* hist_cov, double manual_kp_std, double gate,
                           int32_t* rcounts, uint8_t* reinit, cudaStream_t s, int init_from) {
  MapArgs a{L, K, n_views, n_hist, status, T_GtoC, T_pnp, counts, kp_index, xs, uv, cov, Kb, diameter, map_valid_in, T_OtoG_in, map_valid, T_OtoG,
            hist_crop, hist_T_GtoC, hist_K, hist_off, hist_model_kp, hist_uv, hist_cov, 1.0 / (manual_kp_std * manual_kp_std), gate, rcounts, reinit,
            init_from};
  slam_init_objects_kernel<<<1, 128, 0, s>>>(a);
  slam_reinit_count_kernel<<<n_hist + L, 32, 0, s>>>(a);
  slam_reinit_apply_kernel<<<1, 128, 0, s>>>(a);
  ctx->launches += 3;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

int launch_slam_ba_assemble(suo_ctx* ctx, int L, int K, const int32_t* status, const double* T_GtoC, const uint8_t* map_valid, const double* T_OtoG,
                            const int32_t* counts, const int32_t* kp_index, const double* xs, const float* uv, const float* cov, const double* Kb,
                            double* poses, uint8_t* fixed, int32_t* prob_vert, int32_t* vert_cnt, int32_t* prob_edge, int32_t* edge_cnt, int32_t* e_obj,
                            int32_t* e_cam, double* cam_k, double* p, double* uvd, double* info, uint8_t* inliers, int32_t* edge_src, cudaStream_t s) {
  CurrBaArgs a{L, K, status, T_GtoC, map_valid, T_OtoG, counts, kp_index, xs, uv, cov, Kb, poses, fixed, prob_vert, vert_cnt, prob_edge, edge_cnt,
               e_obj, e_cam, cam_k, p, uvd, info, inliers, edge_src};
  slam_ba_assemble_kernel<<<1, 128, 0, s>>>(a);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

int launch_slam_ba_scatter(suo_ctx* ctx, int L, int K, const int32_t* edge_cnt, const int32_t* edge_src, const uint8_t* inliers, const double* poses,
                           const int32_t* ba_stats, double* T_GtoC, uint8_t* ba_inliers, int32_t* status, cudaStream_t s) {
  slam_ba_scatter_kernel<<<1, 128, 0, s>>>(L, K, edge_cnt, edge_src, inliers, poses, ba_stats, T_GtoC, ba_inliers, status);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}
