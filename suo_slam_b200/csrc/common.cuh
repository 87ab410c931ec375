// Shared host/device helpers for libsuo_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/suo_b200.h"

#define SUO_CUDA_TRY(ctx, expr)                                                        \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SUO_E_CUDA;                                                               \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- one conv layer as the engine sees it -----------------------------------------
enum ConvMode : int { CONV_1x1 = 0, CONV_3x3 = 1, CONV_STEM7 = 2 };

struct ConvParams {
  const float* in;        // NHWC [B,H,W,Cin]  (Cin = floats per pixel as stored)
  const float* w;         // SIMT: [Cout_pad][K] row-major, K in gather order
  const float* w_packed;  // tcgen05: per (n_tile, k_chunk) SW128 images, hi then lo
  const float* bias;      // [Cout_pad] (zeros where absent)
  const float* pre_scale; // [Cin] or nullptr
  const float* pre_shift; // [Cin] or nullptr
  const float* residual;  // NHWC [B,Ho,Wo,Cout_store] or nullptr
  float* out;             // NHWC [B,Ho,Wo,out_c] or NCHW [B,Cout,Ho,Wo] (out_nchw)
  int B, H, W, Cin;       // input geometry
  int Ho, Wo;             // output geometry
  int Cout;               // real output channels written
  int Cout_pad;           // padded to the MMA N granularity
  int out_c;              // channel stride of the NHWC output / residual
  int K;                  // GEMM K (multiple of 32), in gather order
  int mode;               // ConvMode
  int chunks_per_row;     // CONV_STEM7: 32-float chunks per kernel row (ceil(7*Cin/32))
  int relu;               // ReLU after bias
  int out_nchw;           // store NCHW planes instead of NHWC
  long long* dbg;         // optional [5][512] clock64 timeline of CTA 0 (tools only), else nullptr
  const uint16_t* w_packed16;  // tcgen05 fp16x3: per (n_tile, 64-chunk) hi / lo' FP16 images
  int math;               // 0 = TF32 (1 or 3 passes), 1 = FP16x3 split
  int* range_flag;        // fp16x3: set to 1 when an operand exceeds the FP16 range
  int in_split;           // input is two FP16 planes (hi, lo') -> A operand by TMA (tmap_hi / tmap_lo)
  int out_split;          // write the output as two FP16 planes (out_plane = halfs between them)
  size_t out_plane;
  alignas(64) unsigned char tmap_hi[128];   // CUtensorMap of the hi plane [B,H,W,C] fp16, box = one 128-pixel tile x 64 ch
  alignas(64) unsigned char tmap_lo[128];
  int epi_tma;            // tmap_out (/ tmap_out_lo, tmap_res) are valid: the epilogue may store (and fetch the skip tensor) by TMA
  int mma_merge;          // issue A_hi x [B_hi | B_lo'] as one N = 2*BN MMA (fp16x3 / tf32x3)
  alignas(64) unsigned char tmap_out[128];     // FP32 output [rows, out_c], box 32 rows x 32 floats — or the hi plane [rows, C] FP16, box 32 x 64
  alignas(64) unsigned char tmap_out_lo[128];  // lo' plane (out_split)
  alignas(64) unsigned char tmap_res[128];     // skip tensor, FP32 [rows, out_c], box 32 x 32
  unsigned long long* trace;  // developer tool (SUO_TRACE): [0] = launch counter, then {globaltimer start, end, grid, smid} per persistent conv launch
  int raw_tma;            // tmap_raw is valid: a register-fed 1x1 conv may fetch its FP32 input by TMA (plan 3)
  alignas(64) unsigned char tmap_raw[128];     // FP32 input [rows, Cin], box 128 rows x 32 floats
  int stem_raw;           // CONV_STEM7 on the RGB-only layout: tmap_raw is a 4-D map over the zero-bordered input copy whose rows are the
                          // OVERLAPPING 32-float windows of one kernel row (stride 2 pixels): the stem runs on the raw-TMA plan (3)
  int halo;               // tmap_hhi / tmap_hlo are valid and the A-halo kernel (conv_halo.cu) may run this 3x3 layer
  alignas(64) unsigned char tmap_hhi[128];     // FP16 planes [B,H,W,C] with box {64 ch, W, 128 / W + 2 rows, 1}: one column-shifted variant of a tile
  alignas(64) unsigned char tmap_hlo[128];
  int pair;               // tmap_w is valid and the CTA-pair kernel (conv_pair.cu) may run this layer
  alignas(64) unsigned char tmap_w[128];       // FP16x3 weight images as rows of 64 halfs (128 B), box 64 rows, no swizzle (the images are pre-swizzled)
};

struct suo_ctx;
// cudaFuncSetAttribute is per device: one flag per (kernel instance, device) so that several contexts on different GPUs of one
// process configure each kernel on each of them.  `flags` is the launcher's own static array.
inline bool first_use_on_device(bool (&flags)[64], int* num_sms) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (num_sms) cudaDeviceGetAttribute(num_sms, cudaDevAttrMultiProcessorCount, dev);
  dev &= 63;
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}
// 3x3 conv as a CTA pair (tcgen05.mma.cta_group::2, conv_pair.cu): eligibility test and launch
bool conv_pair_eligible(const ConvParams& p, int passes);
int launch_conv_pair(suo_ctx* ctx, const ConvParams& p, cudaStream_t s);
// 3x3 conv as a CTA pair that fetches the activations once per column shift (conv_halo.cu)
bool conv_halo_eligible(const ConvParams& p, int passes);
int launch_conv_halo(suo_ctx* ctx, const ConvParams& p, cudaStream_t s);
int launch_conv_simt(suo_ctx* ctx, const ConvParams& p, cudaStream_t s);
int launch_conv_tc(suo_ctx* ctx, const ConvParams& p, int tf32_passes, cudaStream_t s);
// host-side packing of canonical [Cout_pad][K] weights into the tcgen05 smem images
size_t conv_tc_packed_floats(int Cout_pad, int K);
void conv_tc_pack_weights(const float* w, int Cout_pad, int K, float* dst);
int conv_tc_block_n(int Cout_pad);
size_t conv_tc_packed16_halfs(int Cout_pad, int K);
void conv_tc_pack_weights_f16(const float* w, int Cout_pad, int K, uint16_t* dst);
void conv_tc_host_split_f16(const float* x, size_t n, uint16_t* hi, uint16_t* lo);
void conv_tc_host_join_f16(const uint16_t* hi, const uint16_t* lo, size_t n, float* x);

int launch_heatmap_reduce(suo_ctx* ctx, const float* logits, int B, int K, int H, int W, const float* cls_w,
                          const float* cls_b, float* pooled_scratch, float* uv, float* cov, float* prob,
                          float* mask_logits, float* mask, int32_t* argmax, cudaStream_t s);
int launch_crop_concat(suo_ctx* ctx, const void* images, int n_img, int H, int W, const float* boxes,
                       const int32_t* box_img, int L, const float* priors, int num_kp, int R, float* out, int out_c,
                       cudaStream_t s, int images_u8 = 0, float* out_pad = nullptr, int pad_w = 0, int pad_h = 0);
int launch_render_priors_planes(suo_ctx* ctx, const float* uv, const uint8_t* mask, int L, int K, int vh, int vw, int ndc,
                                float* out, cudaStream_t s);
int launch_render_priors_nhwc(suo_ctx* ctx, const float* uv, const uint8_t* mask, int L, int K, int R, float* out, int out_c,
                              cudaStream_t s);
int launch_maxpool2(suo_ctx* ctx, const float* in, int B, int H, int W, int C, float* out, cudaStream_t s);
int launch_upsample_add(suo_ctx* ctx, const float* up1, const float* low, int B, int H, int W, int C, float* out,
                        cudaStream_t s);
int launch_pnp_batch(suo_ctx* ctx, const double* xs, const double* ys, const int32_t* offsets, int n_obj,
                     double threshold, uint64_t seed, const uint64_t* obj_keys, double* T_out, int32_t* stats,
                     cudaStream_t s, int max_pts);
int launch_pnp_batch_counts(suo_ctx* ctx, const double* xs, const double* ys, const int32_t* offsets,
                            const int32_t* counts, int n_obj, double threshold, uint64_t seed, const uint64_t* obj_keys,
                            double* T_out, int32_t* stats, cudaStream_t s, int max_pts);
int launch_ba_batch_scratch(suo_ctx* ctx, int n_prob, const int32_t* prob_vert, const int32_t* prob_edge, double* poses,
                            const uint8_t* fixed, const int32_t* e_obj, const int32_t* e_cam, const double* cam_k,
                            const double* p, const double* uv, const double* info, uint8_t* inliers,
                            const int32_t* its, int n_rounds, double huber_delta, double chi2_gate,
                            int init_with_outliers, int32_t* stats, double* err_scratch, uint8_t* level_scratch,
                            int8_t* fv_scratch, cudaStream_t s, const int32_t* vert_cnt = nullptr,
                            const int32_t* edge_cnt = nullptr, int single_vertex = 0);

int launch_gate_compact(suo_ctx* ctx, const float* uv, const float* cov, const float* kp_mask, const uint8_t* model_mask,
                        const double* model_kps, const double* K_bbox, int L, int K, float kp_var_thresh, float bbox_thresh,
                        double* xs, double* ys, int32_t* counts, int32_t* kp_index, uint8_t* kp_used, cudaStream_t s);
int launch_frame_ranges(suo_ctx* ctx, const int32_t* box_img, int L, int n_img, int32_t* frame_start, cudaStream_t s);
int launch_ba_assemble(suo_ctx* ctx, int n_img, const int32_t* frame_start, const int32_t* counts, const int32_t* kp_index,
                       const double* xs, const float* uv, const float* cov, const double* K_bbox, const double* diameter,
                       const double* T_pnp, int K, double* poses, uint8_t* fixed, int32_t* prob_vert, int32_t* vert_cnt,
                       int32_t* prob_edge, int32_t* edge_cnt, int32_t* e_obj, int32_t* e_cam, double* cam_k, double* p,
                       double* uvd, double* info, uint8_t* inliers, int32_t* edge_src, uint8_t* accepted, cudaStream_t s);
int launch_ba_scatter(suo_ctx* ctx, int n_img, const int32_t* frame_start, const int32_t* edge_cnt, const int32_t* edge_src,
                      const uint8_t* inliers, const double* poses, const uint8_t* accepted, int K, double* T_ba,
                      uint8_t* ba_inliers, cudaStream_t s);
int launch_pack_records(suo_ctx* ctx, const int32_t* crop_ids, int id_base, const double* T_pnp, const double* T_ba,
                        const uint8_t* kp_used, const uint8_t* ba_inliers, const float* uv, const float* cov, int L, int K,
                        int rec_bytes, uint8_t* out, cudaStream_t s);
int launch_chi2_counts(suo_ctx* ctx, int n_pairs, const double* T, const int32_t* pair_det, const int32_t* det_off,
                       const double* model_kp, const double* K, const float* uv, const float* cov, const uint8_t* use,
                       double manual_kp_std, double gate, int32_t* counts, cudaStream_t s);

// SLAM-mode frame glue (slam.cu)
int launch_slam_kbbox(suo_ctx* ctx, const double* Kc, const float* boxes, int L, double* raw, double* f32, cudaStream_t s);
int launch_slam_vote(suo_ctx* ctx, int n1, int K, int first_view, const double* T_pnp, const int32_t* counts, const int32_t* kp_index,
                     const double* xs, const float* uv, const float* cov, const double* Kb, const double* diameter, const uint8_t* map_valid,
                     const double* T_OtoG, double manual_kp_std, double gate, double* T_GtoC, int32_t* status, cudaStream_t s);
int launch_slam_prior_uv(suo_ctx* ctx, int n1, int L, int K, const int32_t* status, const double* T_GtoC, const uint8_t* map_valid, const double* T_OtoG,
                         const double* model_kps, const uint8_t* model_mask, const double* Kb_raw, float* prior_uv, uint8_t* prior_mask, cudaStream_t s);
int launch_slam_drop_group(suo_ctx* ctx, int n1, int L, int K, const int32_t* status, int32_t* counts, uint8_t* kp_used, double* T_pnp, cudaStream_t s);
int launch_slam_map_update(suo_ctx* ctx, int L, int K, int n_views, int n_hist, const int32_t* status, const double* T_GtoC, const double* T_pnp,
                           const int32_t* counts, const int32_t* kp_index, const double* xs, const float* uv, const float* cov, const double* Kb,
                           const double* diameter, const uint8_t* map_valid_in, const double* T_OtoG_in, uint8_t* map_valid, double* T_OtoG,
                           const int32_t* hist_crop, const double* hist_T_GtoC, const double* hist_K, const int32_t* hist_off,
                           const double* hist_model_kp, const float* hist_uv, const float* hist_cov, double manual_kp_std, double gate,
                           int32_t* rcounts, uint8_t* reinit, cudaStream_t s, int init_from = 0);
int launch_slam_ba_assemble(suo_ctx* ctx, int L, int K, const int32_t* status, const double* T_GtoC, const uint8_t* map_valid, const double* T_OtoG,
                            const int32_t* counts, const int32_t* kp_index, const double* xs, const float* uv, const float* cov, const double* Kb,
                            double* poses, uint8_t* fixed, int32_t* prob_vert, int32_t* vert_cnt, int32_t* prob_edge, int32_t* edge_cnt, int32_t* e_obj,
                            int32_t* e_cam, double* cam_k, double* p, double* uvd, double* info, uint8_t* inliers, int32_t* edge_src, cudaStream_t s);
int launch_slam_ba_scatter(suo_ctx* ctx, int L, int K, const int32_t* edge_cnt, const int32_t* edge_src, const uint8_t* inliers, const double* poses,
                           const int32_t* ba_stats, double* T_GtoC, uint8_t* ba_inliers, int32_t* status, cudaStream_t s);

struct suo_ctx {
  int device = 0;
  int max_crops = 0, crop_res = 0, num_kp = 0;
  std::string err;
  long long launches = 0;
  int opt_backend = 1, opt_passes = 3, opt_graph = 1, opt_persistent = 1, opt_multistream = 0, opt_math = 1;
  int opt_pair = 1;     // 1 = 3x3 convs on FP16-plane tensors run as CTA pairs (conv_pair.cu); SUO_PAIR=0 / SUO_OPT_CONV_PAIR turns it off
  int opt_halo = 1;     // 1 = 3x3 convs at 64x64 / 32x32 / 16x16 run on the A-halo kernel (conv_halo.cu); SUO_HALO=0 / SUO_OPT_CONV_HALO turn it off
  int opt_stem_tma = 1; // 1 = the RGB-only stem fetches its operand by TMA from a zero-bordered input copy (SUO_STEM_TMA=0: register gathers)
  int opt_pdl = 1;      // 1 = the persistent conv kernels are launched with programmatic stream serialization (SUO_PDL / SUO_OPT_PDL)
  int opt_epi_tma = 1, opt_mma_merge = 1, opt_raw_tma = 1;
  int opt_pnp_max_pts = 64;  // SUO_OPT_PNP_MAX_POINTS: shared-memory point capacity per object of DEVICE-pointer suo_pnp_batch calls
  int opt_act_reuse = 1;     // activation buffers with disjoint live ranges share one allocation (SUO_ACT_REUSE=0: one allocation per buffer, needed by SUO_OPT_MULTISTREAM)
  int opt_ba_blockdiag = 0;  // SUO_OPT_BA_BLOCK_DIAGONAL: device-pointer suo_ba_batch calls skip the host-side structure check
  int opt_slam_sfm = 0;      // SUO_OPT_SLAM_SFM: its = [10, 10, 40, 40] for suo_slam_frame's curr_only solve
  unsigned long long* trace = nullptr;       // SUO_TRACE: device launch trace of the persistent conv kernels (dumped by suo_destroy)
  int opt_grid_cap = 0;                      // > 0: persistent conv kernels use at most this many CTAs (SUO_GRID_CAP; concurrent-stream experiments)   // developer switches (SUO_EPI_TMA / SUO_MMA_MERGE): TMA-store epilogue, merged hi|lo' weight MMA
  void* net = nullptr;  // NetState (net_exec.cu)
  void* scratch = nullptr; size_t scratch_bytes = 0;        // device scratch for host-pointer calls
  void* pinned = nullptr; size_t pinned_bytes = 0;          // pinned staging
  void set_error(const std::string& m, const char* f, int l) {
    err = m + " (" + f + ":" + std::to_string(l) + ")";
  }
};
