/* libsuo_b200 — C ABI of the B200-native SUO-SLAM per-frame hot path.
 *
 * Plain C: pointers + sizes, no torch / C++ types.  Every entry point returns 0 on
 * success and a negative SUO_E_* code on failure (suo_last_error() gives the text);
 * nothing throws across the boundary.  One suo_ctx per GPU; a ctx is not thread
 * safe, different ctxs are.  `stream` is a cudaStream_t passed as void* (NULL = the
 * legacy default stream).  `on_device` != 0 means every data pointer of the call is
 * a device pointer on the ctx's GPU and the call only enqueues work on `stream`;
 * `on_device` == 0 means host pointers: the call stages them through pinned memory,
 * runs, copies results back and synchronises the stream before returning (this is
 * the reference-facing form: the reference's pybind entry points take/return numpy).
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference checkout, rpng/suo_slam @ 5de01433).
 */
#ifndef SUO_B200_H_
#define SUO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct suo_ctx suo_ctx;

enum {
  SUO_OK = 0,
  SUO_E_INVALID = -1,   /* bad argument / shape */
  SUO_E_CUDA = -2,      /* CUDA runtime error (text in suo_last_error) */
  SUO_E_STATE = -3,     /* e.g. forward before load_weights */
  SUO_E_NOMEM = -4,
  SUO_E_RANGE = -5      /* fp16x3 conv math saw an activation outside the FP16 range; switch to tf32x3 */
};

/* Options for suo_set_option */
enum {
  SUO_OPT_CONV_BACKEND = 1, /* 0 = FP32 SIMT implicit GEMM, 1 = tcgen05 tensor cores (default; math per SUO_OPT_CONV_MATH) */
  SUO_OPT_TF32_PASSES = 2,  /* 3 = 3xTF32 split (FP32-equivalent, default), 1 = single-pass TF32 */
  SUO_OPT_USE_GRAPH = 3,    /* 1 = replay the forward as a CUDA graph (default), 0 = eager launches */
  SUO_OPT_CONV_PERSISTENT = 4, /* 1 = persistent tcgen05 conv kernel with overlapped epilogue (default), 0 = one tile per CTA */
  SUO_OPT_MULTISTREAM = 5,     /* 1 = run the hourglass resolution levels on concurrent streams (graph branches); only with SUO_ACT_REUSE=0 in the
                                  environment at suo_create (shared activation allocations assume one stream) */
  SUO_OPT_CONV_MATH = 6,       /* 1 = FP16x3 split (default): x = hi + 2^-11 lo in two FP16 numbers, range-guarded;
                                  0 = TF32 split (SUO_OPT_TF32_PASSES) */
  SUO_OPT_CONV_FUSE = 7,       /* retired (round 1's fused conv2 + conv3 kernels measured slower than the two kernels and were removed): only 0 is accepted */
  SUO_OPT_CONV_PAIR = 8,       /* 1 (default) = 3x3 convs on FP16-plane tensors run as CTA pairs (tcgen05.mma.cta_group::2: each CTA of
                                   a 2-CTA cluster loads half of the weight rows); 0 = one CTA per tile. Results are identical. */
  SUO_OPT_PDL = 9,             /* 1 (default) = the persistent conv kernels use programmatic dependent launch (the next kernel's CTAs are
                                   scheduled and run their prologue while the previous kernel drains; griddepcontrol.wait before any
                                   activation is touched); 0 = plain stream order */
  SUO_OPT_CONV_HALO = 10,      /* 1 (default) = the 3x3 convs at 64x64 / 32x32 / 16x16 fetch their activations once per column shift (A-halo
                                   CTA-pair kernel, conv_halo.cu); its accumulation order differs from the other 3x3 kernels: equal to FP32
                                   rounding, not bit for bit.  0 = CTA-pair kernel of SUO_OPT_CONV_PAIR */
  SUO_OPT_PNP_MAX_POINTS = 12,  /* points per object the PnP kernel reserves shared memory for on DEVICE-pointer suo_pnp_batch calls (default 64,
                                   4..4096).  Host-pointer calls size it from `offsets`.  An object with more points is NOT truncated: it
                                   returns the identity (failure) with stats[.,0] = -1 */
  SUO_OPT_BA_BLOCK_DIAGONAL = 11 /* 1 = the caller vouches that every graph passed to suo_ba_batch with DEVICE pointers has exactly one free
                                   vertex per edge and at most 64 vertices (single-view / curr_only graphs): the call then only enqueues the
                                   shared-memory kernel instead of copying the index arrays to the host to choose a kernel (a violation is
                                   reported per problem through stats[.,0] = -1 / -2 / -3).  2 = additionally every problem has ONE free vertex and at
                                   most one fixed one (BASELINE config 3, the curr_only camera solve): one warp per problem, state in registers.
                                   0 (default) = inspect the structure (host-pointer calls always do, and pick the same kernels) */
  , SUO_OPT_SLAM_SFM = 13       /* 1 = suo_slam_frame's curr_only solve runs its = [10, 10, 40, 40] as ObjectSLAM does in sfm_mode
                                   (lib/object_slam.py:843-846); 0 (default) = [10] * 4 (SLAM mode) */
};

/* BA vertex/edge conventions (see suo_ba_batch) */
enum { SUO_BA_EDGE_UNARY = -1 };

/* ---- lifetime ------------------------------------------------------------------ */
/* max_crops: largest crop batch one forward will see; crop_res: network input side
 * (256 -> 64x64 heat-maps, lib/models/pkpnet.py:67 input_res); num_kp: 41
 * (lib/labeling/kp_config.py:93-94). */
int suo_create(int device, int max_crops, int crop_res, int num_kp, suo_ctx** out);
void suo_destroy(suo_ctx* ctx);
const char* suo_last_error(const suo_ctx* ctx);
int suo_set_option(suo_ctx* ctx, int option, int value);
/* Number of this library's kernels launched since ctx creation (bench.py gpu_launches). */
long long suo_kernel_launches(const suo_ctx* ctx);
/* Bytes of HBM held by the network's activation tensors (after suo_load_weights): tensors whose live ranges in the layer program do not
 * overlap share one allocation.  *unshared (may be NULL) = what one allocation per tensor would take. */
size_t suo_activation_bytes(const suo_ctx* ctx, size_t* unshared);

/* ---- weights ------------------------------------------------------------------- */
/* Replaces PkpNet.load_state_dict(checkpoint['model']) (lib/object_slam.py:92-97).
 * `blob` is the packed, BN-folded image produced by suo_slam_b200.weights.pack_state_dict
 * (host pointer, copied). */
int suo_load_weights(suo_ctx* ctx, const void* blob, size_t nbytes);

/* Synchronises and reports (then clears) the FP16-range flag of the fp16x3 conv math mode: SUO_OK or
 * SUO_E_RANGE.  Host-pointer calls check it themselves; device-pointer (asynchronous) callers poll it. */
int suo_check_range(suo_ctx* ctx);

/* ---- network forward ----------------------------------------------------------- */
/* Replaces PkpNet.forward(images, boxes, prior_kp) (lib/models/pkpnet.py:80-119):
 * roi_align crop + prior concat -> 2-stack hourglass -> spatial softmax ->
 * soft-argmax / 2x2 covariance -> keypoint-present classifier.
 *   images   [n_img,3,H,W] f32 in [0,1] (NCHW, as the reference passes them)
 *   boxes    [L,4] xyxy f32, box_img[L] = image index of each box
 *   priors   [L,num_kp,R,R] f32 or NULL (NULL == all-zero prior planes)
 * Outputs (any may be NULL to skip): uv[L,K,2], cov[L,K,2,2], logits[L,K,R/4,R/4],
 * prob (same shape), mask_logits[L,K], mask[L,K], argmax[L,K] (flat h*W+w of the max
 * logit, first occurrence). */
int suo_forward(suo_ctx* ctx, const float* images, int n_img, int H, int W,
                const float* boxes, const int32_t* box_img, int L, const float* priors,
                float* uv, float* cov, float* logits, float* prob,
                float* mask_logits, float* mask, int32_t* argmax,
                int on_device, void* stream);

/* suo_forward with the priors given as KEYPOINTS instead of dense planes: prior_uv [L,K,2] f32
 * (NDC, as ObjectSLAM builds prior_uv_full, lib/object_slam.py:510-512) and prior_mask [L,K] u8.  The planes that
 * utils.make_prior_kp_input (lib/utils/utils.py:398-411) would have produced are stamped straight into the network
 * input on the device: identical result, no 10.75 MB / crop of prior planes built on the CPU and copied over. */
int suo_forward_kp_priors(suo_ctx* ctx, const float* images, int n_img, int H, int W,
                          const float* boxes, const int32_t* box_img, int L,
                          const float* prior_uv, const uint8_t* prior_mask,
                          float* uv, float* cov, float* logits, float* prob,
                          float* mask_logits, float* mask, int32_t* argmax,
                          int on_device, void* stream);

/* Replaces utils.make_prior_kp_input(kp_uv, kp_uv_mask, img_shape, ndc) (lib/utils/utils.py:398-411 ->
 * draw_gaussian_2d :364-385 -> gaussian_2d :356-361) for a batch: prior_uv [L,K,2] f32, prior_mask [L,K] u8 ->
 * out [L,K,height,width] f32, bit-identical to the reference (same 91x91 cv2.GaussianBlur stamp, same rounding). */
int suo_render_priors(suo_ctx* ctx, const float* prior_uv, const uint8_t* prior_mask, int L, int K,
                      int height, int width, int ndc, float* out, int on_device, void* stream);

/* Stand-alone heat-map reduction (spatial_softmax + post_process_kp + classifier,
 * lib/models/pkpnet.py:13-63,74-78,106-118).  logits [B,K,H,W] f32; cls_w [K,K],
 * cls_b [K] may be NULL (then mask outputs are skipped). */
int suo_heatmap_reduce(suo_ctx* ctx, const float* logits, int B, int K, int H, int W,
                       const float* cls_w, const float* cls_b,
                       float* uv, float* cov, float* prob, float* mask_logits, float* mask,
                       int32_t* argmax, int on_device, void* stream);

/* roi_align crop + prior concat alone (lib/models/pkpnet.py:91-101); out is the NHWC
 * tensor the network consumes: [L,R,R,out_c] with out_c = 4 (priors == NULL: RGB + one
 * zero plane) or 48 (RGB, num_kp prior planes, zero padding). */
int suo_crop_concat(suo_ctx* ctx, const float* images, int n_img, int H, int W,
                    const float* boxes, const int32_t* box_img, int L, const float* priors,
                    int R, float* out, int out_c, int on_device, void* stream);

/* One convolution layer through the library's conv engine — test/bench hook for the
 * dense-contraction kernels (no reference counterpart; the reference calls cuDNN/MKL-DNN
 * through torch.nn.Conv2d, lib/models/layers/Residual.py:9-18).
 *   in  [B,H,W,Cin] NHWC f32;  w [Cout,kh,kw,Cin] f32;  bias [Cout] or NULL
 *   pre_scale/pre_shift [Cin] or NULL: input prologue relu(x*s+t) (pre-activation BN)
 *   residual [B,Ho,Wo,Cout] or NULL; relu: apply ReLU after bias (before residual is
 *   never needed by the network; residual layers have relu == 0)
 *   ksize in {1,3,7}; stride 1 (ksize 1,3; pad (ksize-1)/2) or 2 (ksize 7, pad 3)
 *   backend: 0 SIMT FP32, 1 tcgen05 TF32 (tf32_passes as SUO_OPT_TF32_PASSES), 2 tcgen05 FP16x3;
 *            3 = FP16x3 with the A operand fed by TMA from pre-split FP16 planes (the call splits `in` on the host),
 *            4 = FP16x3 writing the output as FP16 planes (re-joined on the host), 5 = both. */
int suo_conv2d(suo_ctx* ctx, const float* in, int B, int H, int W, int Cin,
               const float* w, const float* bias, int Cout, int ksize, int stride,
               const float* pre_scale, const float* pre_shift, const float* residual, int relu,
               float* out, int backend, int tf32_passes, int on_device, void* stream);

/* ---- PnP ----------------------------------------------------------------------- */
/* Replaces lambdatwist.pnp(xs, ys, threshold) (thirdparty/lambdatwist/pnp_python_binding.cpp:32-62
 * -> PNP::compute / refine, pnp_ransac.cpp:188-326) for a BATCH of objects.
 *   xs [N,3] f64 model points, ys [N,2] f64 pinhole-normalised image points,
 *   offsets[n_obj+1] row ranges per object (object o owns rows offsets[o]..offsets[o+1])
 *   seed / obj_keys[n_obj] (NULL -> key = object index): counter-based RANSAC sampling
 *   T_out [n_obj,16] row-major 4x4; identity == failure, exactly like the reference
 *   stats [n_obj,5] i32 or NULL: best_inliers, best_iter, total_iters, refine1_its, refine2_its
 * Any number of points per object up to 4876 (the points of one object live in shared memory; the reference's own
 * benchmark uses 250, thirdparty/lambdatwist/test_pnp.cpp:68-147). */
int suo_pnp_batch(suo_ctx* ctx, const double* xs, const double* ys, const int32_t* offsets, int n_obj,
                  double threshold, uint64_t seed, const uint64_t* obj_keys,
                  double* T_out, int32_t* stats, int on_device, void* stream);

/* ---- bundle adjustment --------------------------------------------------------- */
/* Replaces the g2o solve inside ObjectSLAM.optimize() (lib/object_slam.py:842-896 driving
 * SparseOptimizer::optimize / OptimizationAlgorithmLevenberg::solve with
 * EdgeSE3ProjectFromObject / EdgeSE3ProjectFromFixedObject,
 * thirdparty/g2opy/g2o/types/object_slam/types_object_slam.cpp:45-201) for a BATCH of
 * independent problems (one LM lambda / accept test per problem, as g2o has per graph).
 *   poses     [n_vert,12] f64 row-major [R|t], in/out (all problems' vertices concatenated)
 *   fixed     [n_vert] u8
 *   prob_vert [n_prob+1], prob_edge[n_prob+1]: vertex / edge ranges of each problem
 *   e_obj     [n_edges] object vertex (global index) or SUO_BA_EDGE_UNARY (p is then p_inG)
 *   e_cam     [n_edges] camera vertex (global index)
 *   cam_k [n_edges,4] fx fy cx cy; p [n_edges,3]; uv [n_edges,2]; info [n_edges,4]
 *   inliers   [n_edges] u8 in/out
 *   its[n_rounds] LM iterations per round; huber_delta; chi2_gate; init_with_outliers
 *   stats     [n_prob,3] i32 or NULL: rounds run, outer iterations, LM trials
 * Graphs in which every edge touches exactly one non-fixed vertex (single-view mode: camera fixed; curr_only mode:
 * objects folded into p_inG) with at most 64 vertices run on the shared-memory kernel (csrc/ba.cu); coupled
 * camera+object graphs (global BA, lib/object_slam.py:736-778 -> LinearSolverCholmod) and larger graphs run on the
 * Schur-complement kernel (csrc/ba_global.cu).  The call routes by graph structure. */
int suo_ba_batch(suo_ctx* ctx, int n_prob, const int32_t* prob_vert, const int32_t* prob_edge,
                 double* poses, const uint8_t* fixed, int n_vert,
                 const int32_t* e_obj, const int32_t* e_cam, const double* cam_k, const double* p,
                 const double* uv, const double* info, uint8_t* inliers, int n_edges,
                 const int32_t* its, int n_rounds, double huber_delta, double chi2_gate,
                 int init_with_outliers, int32_t* stats, int on_device, void* stream);

/* The 2-vector error of every edge as the most recent suo_ba_batch call left it — what g2o keeps inside each edge
 * after SparseOptimizer::optimize and ObjectSLAM.optimize() reads back through e.chi2() without recomputing
 * (lib/object_slam.py:881-883; after a rejected LM trial that is the REJECTED state's error,
 * optimization_algorithm_levenberg.cpp:120-141).  err [n_edges,2] f64; n_edges must match that call. */
int suo_ba_last_errors(suo_ctx* ctx, double* err, int n_edges, int on_device, void* stream);

/* computeError + linearizeOplus of a batch of independent edges (thirdparty/g2opy/g2o/types/object_slam/
 * types_object_slam.cpp:45-60,70-123 EdgeSE3ProjectFromObject; :156-169,177-201 EdgeSE3ProjectFromFixedObject when
 * T_obj == NULL): the device functions the LM kernels use, exposed so that the analytic Jacobians can be checked
 * against central differences (the recipe the reference left commented out at :108-122).
 *   T_obj [n,12] or NULL, T_cam [n,12] row-major [R|t]; cam_k [n,4]; p [n,3]; uv [n,2]
 *   err [n,2], J_obj [n,2,6], J_cam [n,2,6] (tangent order omega, upsilon; update T <- exp(dx) T); any may be NULL */
int suo_edge_linearize(suo_ctx* ctx, int n_edges, const double* T_obj, const double* T_cam, const double* cam_k,
                       const double* p, const double* uv, double* err, double* J_obj, double* J_cam,
                       int on_device, void* stream);

/* ---- keypoints -> poses ---------------------------------------------------------- */
/* Everything ObjectSLAM does with the network output of single-view frames, device resident:
 * keypoint gating (lib/object_slam.py:1100-1115), per-object PnP (:1123-1165 -> lambdatwist.pnp)
 * and the single-view optimize() (:703-930, camera fixed at identity, its=[10]*4).
 *   uv [L,K,2] f32, cov [L,K,4] f32, kp_mask [L,K] f32 (sigmoid output); box_img [L] sorted:
 *   crops of one image form one BA graph (one LM lambda per image, as g2o has per optimizer).
 * Other arguments / outputs as suo_frames. */
int suo_solve_keypoints(suo_ctx* ctx, const float* uv, const float* cov, const float* kp_mask, const int32_t* box_img,
                        int n_img, int L, const double* model_kps, const uint8_t* model_mask, const double* K_bbox,
                        const double* diameter, double kp_var_thresh, double bbox_thresh, uint64_t seed, int run_ba,
                        double* T_pnp, double* T_ba, uint8_t* kp_used, uint8_t* ba_inliers, int on_device, void* stream);

/* ---- hypothesis scoring -------------------------------------- */
/* The chi2 inlier count that ObjectSLAM.__estimate_camera_pose (lib/object_slam.py:1030-1066, camera-pose voting) and
 * __maybe_reinit_objects (:645-680) evaluate in nested Python loops, for a batch of (pose, detection) pairs:
 *   T_pairs [n_pairs,12] f64 row-major [R|t] = T_OtoC hypothesis of each pair, pair_det[n_pairs] its detection
 *   det_off[n_det+1] keypoint rows of each detection; model_kp [N,3] f64; K [n_det,9] f64 (the detection's bbox-NDC
 *   camera matrix); uv [N,2] f32; cov [N,4] f32 or NULL (then inf = I / manual_kp_std^2, :1059-1061);
 *   use [N] u8 or NULL (the detection's "inliers" mask, :1035)
 *   counts[n_pairs]: keypoints in front of the camera whose chi2 (cov diagonal floored at 1e-4) is <= chi2_gate. */
int suo_chi2_inlier_counts(suo_ctx* ctx, int n_pairs, const double* T_pairs, const int32_t* pair_det, int n_det,
                           const int32_t* det_off, const double* model_kp, const double* K, const float* uv,
                           const float* cov, const uint8_t* use, double manual_kp_std, double chi2_gate,
                           int32_t* counts, int on_device, void* stream);

/* ---- fused per-frame pipeline -------------------------------------------------- */
/* One call for a batch of single-view frames: forward -> keypoint gating
 * (lib/object_slam.py:1100-1115) -> per-object PnP (:1123-1165) -> single-view BA
 * (optimize(), :703-930, camera fixed at identity) with everything resident on the GPU.
 *   model_kps [L,K,3] f64, model_mask [L,K] u8, K_bbox [L,9] f64 (fix_K_for_bbox_ndc),
 *   diameter [L] f64; thresholds as ObjectSLAM.__init__ (kp_var_thresh, bbox_thresh)
 * Outputs: T_pnp [L,16] f64 (identity = rejected), T_ba [L,12] f64, kp_used [L,K] u8,
 * ba_inliers [L,K] u8, uv [L,K,2] f32, cov [L,K,4] f32 (any may be NULL). */
int suo_frames(suo_ctx* ctx, const float* images, int n_img, int H, int W,
               const float* boxes, const int32_t* box_img, int L, const float* priors,
               const double* model_kps, const uint8_t* model_mask, const double* K_bbox,
               const double* diameter, double kp_var_thresh, double bbox_thresh,
               uint64_t seed, int run_ba,
               double* T_pnp, double* T_ba, uint8_t* kp_used, uint8_t* ba_inliers,
               float* uv, float* cov, int on_device, void* stream);

/* suo_frames on the camera's own frames: images_hwc [n_img,H,W,3] u8, as ObjectSLAM.process_view receives them
 * (lib/object_slam.py:327-328).  The reference converts the frame on the host, float32(img) / 255 (:1092), and copies
 * 4 bytes per value to the GPU; here the same division is applied per bilinear tap inside the crop kernel (identical
 * crop values), so a frame costs a quarter of the PCIe bytes.  Everything else as suo_frames. */
int suo_frames_u8(suo_ctx* ctx, const uint8_t* images_hwc, int n_img, int H, int W,
                  const float* boxes, const int32_t* box_img, int L, const float* priors,
                  const double* model_kps, const uint8_t* model_mask, const double* K_bbox,
                  const double* diameter, double kp_var_thresh, double bbox_thresh,
                  uint64_t seed, int run_ba,
                  double* T_pnp, double* T_ba, uint8_t* kp_used, uint8_t* ba_inliers,
                  float* uv, float* cov, int on_device, void* stream);

/* ---- SLAM-mode frame ------------------------------------------ */
/* One view of ObjectSLAM.process_view with single_view_mode off (lib/object_slam.py:393-421), device resident from the camera
 * frame to the refined camera pose:
 *   forward on the non-symmetric crops -> gating -> per-object PnP                      (__process_objects(False), :394-397)
 *   camera-pose vote over those PnP poses against the map                                (__estimate_camera_pose, :975-1072)
 *   prior keypoints of the symmetric crops projected from the map, stamped on the device (:486-514, utils.make_prior_kp_input)
 *   forward on the symmetric crops with those priors -> gating -> PnP                    (__process_objects(True), :413-418)
 *   initialisation of unmapped objects (:577-592), re-initialisation test over the last views (__maybe_reinit_objects, :595-697)
 *   optimize(curr_only=True): LM on the camera vertex with one unary edge per gated keypoint, its = [10] * 4 (:703-930)
 * with no host work between the stages.  Crops [0, n_nonsym) are the non-symmetric objects, [n_nonsym, L) the symmetric ones
 * (mesh_db[obj]["is_symmetric"], :343); L <= 128.
 *   image_hwc [H,W,3] u8; K_cam [9] f64; boxes [L,4] xyxy f32 (already inflated, :390-391); model_kps [L,K,3] f64;
 *   model_mask [L,K] u8; diameter [L] f64; map_valid [L] u8 and T_OtoG [L,12] f64: the map pose of each crop's object, if any
 *   n_views: views processed so far INCLUDING this one (1 = first view: camera = identity, the PnP poses define the map)
 *   history for the re-initialisation test: the objects' detections in up to 14 earlier views — n_hist detections, hist_crop[h] =
 *     the crop (object) it belongs to, hist_T_GtoC [n_hist,12] the camera pose of that view, hist_K [n_hist,9] its bbox-NDC camera
 *     matrix, hist_off [n_hist+1] its keypoint rows in hist_model_kp [N,3] f64 / hist_uv [N,2] f32 / hist_cov [N,4] f32 (or NULL)
 *   manual_kp_std: used instead of the network covariance when a cov pointer is NULL (:1059-1061); init_with_outliers: :848-851
 * Outputs (any may be NULL): T_GtoC [12] f64 (after the curr_only solve); status [8] i32 = {camera pose known, votes of the winning
 * hypothesis, number of hypotheses, curr_only edges, LM trials, curr_only inlier edges, 0, 0}; T_pnp [L,16]; kp_used, ba_inliers
 * [L,K] u8; uv [L,K,2], cov [L,K,4] f32; prior_uv [L,K,2] f32 + prior_mask [L,K] u8 (what the symmetric crops were given);
 * K_bbox [L,9] f64; T_OtoG_out [L,12] + map_valid_out [L] (the updated map); reinit [L] u8, reinit_counts [L,2] i32 (pnp, estim).
 * Camera pose known to the caller — T_GtoC_init [12] f64 (NULL with cam_init_mode 0 = the vote above): the vote is skipped, status =
 * {1, 0, 0, ...} and the view proceeds from that pose (priors, object initialisation, re-initialisation test, curr_only solve).
 *   cam_init_mode 1: known BEFORE the view — external odometry (process_view's cam_pose, :349-353, pass n_nonsym = 0 as the reference treats
 *                    every object as symmetric then) or __backup_estimate_camera_pose called because no non-symmetric object is in view (:372-391);
 *   cam_init_mode 2: __backup_estimate_camera_pose after a FAILED vote (:404-411; a first call with mode 0 returned status[0] = 0): as mode 1,
 *                    but the unmapped objects of the non-symmetric group are not initialised (their pass returned at :566-575).
 * The backup pose itself (bbox-centroid PnP = one suo_pnp_batch object, or the constant-velocity guess, :933-973) is the caller's:
 * slam.SlamTracker does it.  Not covered (bookkeeping the caller keeps; slam.SlamTracker mirrors both): object culling
 * (:913-930), the periodic global optimize() (one coupled suo_ba_batch problem). */
int suo_slam_frame(suo_ctx* ctx, const uint8_t* image_hwc, int H, int W, const double* K_cam,
                   const float* boxes, int L, int n_nonsym,
                   const double* model_kps, const uint8_t* model_mask, const double* diameter,
                   const uint8_t* map_valid, const double* T_OtoG, int n_views,
                   int n_hist, const int32_t* hist_crop, const double* hist_T_GtoC, const double* hist_K,
                   const int32_t* hist_off, const double* hist_model_kp, const float* hist_uv, const float* hist_cov,
                   double kp_var_thresh, double bbox_thresh, double manual_kp_std, int init_with_outliers, uint64_t seed,
                   double* T_GtoC, int32_t* status, double* T_pnp, uint8_t* kp_used, uint8_t* ba_inliers,
                   float* uv, float* cov, float* prior_uv, uint8_t* prior_mask, double* K_bbox,
                   double* T_OtoG_out, uint8_t* map_valid_out, uint8_t* reinit, int32_t* reinit_counts,
                   const double* T_GtoC_init, int cam_init_mode, int on_device, void* stream);

/* Asynchronous, double-buffered form of suo_frames_u8 for a STREAM of frame batches (host pointers, priors == NULL):
 * submit enqueues the host->device copies of the batch on the library's copy stream and the frame path on `stream`
 * and returns at once; suo_frames_wait blocks until that slot's results are in the host output buffers (and reports
 * the FP16-range flag like the synchronous call).  With two slots the copies of batch i+1 overlap the kernels of
 * batch i.  Input and output buffers must stay valid (and should be pinned) until the wait returns; a slot must be
 * waited for before it is submitted again, and both slots must use the SAME `stream` (they share the executor; a
 * different stream while the other slot is pending is SUO_E_INVALID).  slot in {0, 1}.  records_dev (DEVICE pointer, L records, or NULL): the
 * batch's result records (suo_pack_records layout, crop_id = record_id_base + index) are also packed there on `stream`,
 * ready for suo_allgather_results — the multi-GPU exchange then needs no host round trip. */
int suo_frames_u8_submit(suo_ctx* ctx, int slot, const uint8_t* images_hwc, int n_img, int H, int W,
                         const float* boxes, const int32_t* box_img, int L,
                         const double* model_kps, const uint8_t* model_mask, const double* K_bbox,
                         const double* diameter, double kp_var_thresh, double bbox_thresh,
                         uint64_t seed, int run_ba,
                         double* T_pnp, double* T_ba, uint8_t* kp_used, uint8_t* ba_inliers,
                         float* uv, float* cov, void* records_dev, int record_id_base, void* stream);
int suo_frames_wait(suo_ctx* ctx, int slot);

/* ---- multi-GPU exchange (BASELINE.json configs[3]) -------------- */
/* The reference has no inference-side distribution.  Crops / frames are sharded over one process per GPU and the
 * per-crop results are exchanged ONCE, as fixed-size records, before any step that needs all objects (camera-pose
 * voting lib/object_slam.py:975-1072, the joint graph :736-837).  Record of one crop (suo_record_bytes(K) bytes,
 * 1240 for K = 41, little endian):
 *   f64 T_pnp[12] | f64 T_ba[12] | i32 crop_id, accepted, n_used, n_ba_inliers | f32 uv[K][2] | f32 cov[K][4] |
 *   u8 flags[K] (bit 0 = keypoint passed the gate, bit 1 = BA inlier) | zero padding to a multiple of 8 bytes. */
size_t suo_record_bytes(int num_kp);
/* Packs the outputs of suo_frames / suo_solve_keypoints into records (one device kernel).  crop_ids [L] or NULL
 * (then crop_id = id_base + index); T_pnp [L,16]; T_ba [L,12], kp_used, ba_inliers, uv, cov may be NULL. */
int suo_pack_records(suo_ctx* ctx, const int32_t* crop_ids, int id_base, const double* T_pnp, const double* T_ba,
                     const uint8_t* kp_used, const uint8_t* ba_inliers, const float* uv, const float* cov, int L,
                     void* records, int on_device, void* stream);
/* ncclAllGather of n_local records per rank on `stream` (device pointers; `nccl_comm` is the caller's ncclComm_t;
 * `out` holds world_size * n_local records, rank-major).  The library calls the NCCL the process has already loaded
 * (dlopen of libnccl.so.2): it does not link NCCL itself.  Every rank passes the same n_local (pad with crop_id -1). */
int suo_allgather_results(suo_ctx* ctx, void* nccl_comm, const void* records, size_t rec_bytes, int n_local,
                          void* out, void* stream);

/* ---- measurement ----------------------------------------------------------------- */
/* Per-op CUDA-event timing of the network program on the current input buffer (eager launches):
 * average ms per forward in the conv kernels and in the pool / up-sample kernels.  bench.py uses it
 * for the roofline block; it has no reference counterpart (the reference times with time.time(),
 * lib/utils/utils.py:20-23). */
int suo_profile_network(suo_ctx* ctx, int L, int with_priors, int iters, float* conv_ms, float* other_ms, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SUO_B200_H_ */
