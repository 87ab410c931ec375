// C-ABI entry points of libsuo_b200 (include/suo_b200.h) and the network executor.
//
// The network topology is not hard-coded here: suo_load_weights() receives a packed blob
// (suo_slam_b200/weights.py) holding a buffer table, an op list ("program") and a float pool
// with BN-folded weights.  The executor owns the activation buffers (NHWC FP32 in HBM, one
// buffer per op output: 180 GB of HBM make liveness-based reuse unnecessary at these sizes),
// packs the weights for the tensor-core kernel once, and replays the op list — as a CUDA
// graph per (crop count, variant) — on the caller's stream.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <array>
#include <map>
#include <memory>
#include <tuple>

#include <cuda.h>
#include <dlfcn.h>

#include "common.cuh"
#include "ba_math.cuh"

namespace {

constexpr int32_t kMagic = 0x574F5553;  // 'SUOW'

// CUtensorMap of one FP16 plane [B,H,W,C] (NHWC) whose box is one 128-pixel output tile x 64 channels (128 B,
// SWIZZLE_128B, zero fill outside the tensor = the 3x3 conv's padding).  Driver entry point fetched at run time so the
// library does not link libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;
int make_plane_tmap(suo_ctx* ctx, void* out128, const void* base, int C, int W, int H, int B) {
  EncodeTiledFn& fn = g_encode_tiled;
  if (!fn) {
    cudaDriverEntryPointQueryResult qres;
    void* f = nullptr;
    SUO_CUDA_TRY(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres));
    if (!f || qres != cudaDriverEntryPointSuccess) { ctx->set_error("cuTensorMapEncodeTiled not available", __FILE__, __LINE__); return SUO_E_CUDA; }
    fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  if (!out128) return SUO_OK;       // entry-point resolution only
  int bw, bh, bb;
  if (W >= 128) { bw = 128; bh = 1; bb = 1; }
  else { bw = W; bh = std::min(H, 128 / W); bb = 128 / (W * bh); }
  if (bw * bh * bb != 128 || C % 64) { ctx->set_error("make_plane_tmap: tile does not cover whole rows", __FILE__, __LINE__); return SUO_E_INVALID; }
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { ctx->set_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r), __FILE__, __LINE__); return SUO_E_CUDA; }
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
  memcpy(out128, &m, 128);
  return SUO_OK;
}
// The same planes with the box of the A-halo kernel (conv_halo.cu): {64 ch, W px, 128 / W + 2 rows, 1 image} — one column-shifted variant
// of a 128-pixel tile including the row above and the row below it; zero fill outside the image = the 3x3 conv's padding in x and y.
int make_plane_tmap_halo(suo_ctx* ctx, void* out128, const void* base, int C, int W, int H, int B) {
  if (!(W == 16 || W == 32 || W == 64) || C % 64 || (H * W) % 128) return SUO_E_INVALID;
  int rc = make_plane_tmap(ctx, nullptr, nullptr, 0, 0, 0, 0);   // resolves the driver entry point
  if (rc) return rc;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)W, (cuuint32_t)(128 / W + 2), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = g_encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { ctx->set_error("cuTensorMapEncodeTiled (halo) failed: " + std::to_string((int)r), __FILE__, __LINE__); return SUO_E_CUDA; }
  memcpy(out128, &m, 128);
  return SUO_OK;
}
// CUtensorMap of a row-major 2-D view [rows, cols] (the NHWC tensor with pixels flattened) whose box is 32 rows x 128 bytes
// (32 floats or 64 halfs), SWIZZLE_128B: what one epilogue warp stores (or fetches, for the skip tensor) per TMA operation.
int make_rows_tmap(suo_ctx* ctx, void* out128, const void* base, bool fp16, size_t rows, int cols, int box_rows = 32) {
  CUtensorMap m;
  const int es = fp16 ? 2 : 4;
  if (((size_t)cols * es) % 16 || (reinterpret_cast<uintptr_t>(base) & 15)) { ctx->set_error("make_rows_tmap: row pitch / base not 16-byte aligned", __FILE__, __LINE__); return SUO_E_INVALID; }
  int rc = make_plane_tmap(ctx, nullptr, nullptr, 0, 0, 0, 0);   // resolves the driver entry point
  (void)rc;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * es};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(&m, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims,
                              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { ctx->set_error("cuTensorMapEncodeTiled (rows) failed: " + std::to_string((int)r), __FILE__, __LINE__); return SUO_E_CUDA; }
  memcpy(out128, &m, 128);
  return SUO_OK;
}
// CUtensorMap over the FP16x3 weight images of a layer viewed as rows of 64 halfs (128 B): box = 64 rows, NO swizzle —
// the images are stored pre-swizzled, the copy must be verbatim.  Used by the CTA-pair kernel, where each CTA fetches
// half of the rows of every image and the bytes complete on the leader CTA's mbarrier (conv_pair.cu).
int make_weight_tmap(suo_ctx* ctx, void* out128, const uint16_t* base, size_t halfs) {
  if (halfs % 64 || (reinterpret_cast<uintptr_t>(base) & 127)) { ctx->set_error("make_weight_tmap: images not 128-byte aligned", __FILE__, __LINE__); return SUO_E_INVALID; }
  int rc = make_plane_tmap(ctx, nullptr, nullptr, 0, 0, 0, 0);   // resolves the driver entry point
  if (rc) return rc;
  CUtensorMap m;
  const cuuint64_t dims[2] = {64, (cuuint64_t)(halfs / 64)};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {64, 64};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { ctx->set_error("cuTensorMapEncodeTiled (weights) failed: " + std::to_string((int)r), __FILE__, __LINE__); return SUO_E_CUDA; }
  memcpy(out128, &m, 128);
  return SUO_OK;
}
enum OpType : int32_t { OP_CONV = 0, OP_MAXPOOL = 1, OP_UPADD = 2 };

struct BlobHeader {
  int32_t magic, version, num_kp, n_bufs, n_ops, n_floats;
  int32_t in_buf_noprior, in_buf_prior, logits_buf, cls_w_off, cls_b_off, heat_div;
  int32_t reserved[4];
};
struct BufDesc { int32_t div, C, kind; };   // kind 1: may be stored as two FP16 planes when the fp16x3 tensor-core path runs
struct OpDesc {
  int32_t type, variant, in, out, res, mode, Cin, Cout, Cout_pad, K, cpr, relu, out_nchw, w_off, b_off, pre_off;
};

struct NetState {
  BlobHeader h{};
  std::vector<BufDesc> bufs;
  std::vector<OpDesc> ops;
  float* pool = nullptr;                     // device float pool (canonical weights, biases, prologues)
  std::vector<float*> packed;                // per op: tcgen05 TF32 weight images (device) or nullptr
  std::vector<uint16_t*> packed16;           // per op: tcgen05 FP16x3 weight images (device) or nullptr
  int* range_flag = nullptr;                 // device int: FP16 operand range exceeded
  bool w_exceeds_fp16 = false;               // some folded weight is outside the FP16 range (or not finite): fp16x3 math is refused
  std::vector<std::array<unsigned char, 1152>> tmaps;   // per op: in hi / in lo / out (FP32 or hi) / out lo / skip / raw input / weight-image CUtensorMaps (split mode)
  std::vector<int> halo_ok;                             // per op: the A-halo maps (bytes 896.. of tmaps) are valid
  std::vector<int> epi_ok, raw_ok, wmap_ok;             // per op: the output / skip maps, the FP32 input map, the weight-image map are valid
  // TMA-fed stem (RGB-only layout): a second copy of the network input with a 3-pixel zero border, rows of stem_wp pixels, stem_hp rows
  // per crop, and the 4-D tensor map of its overlapping kernel-row windows
  float* stem_pad = nullptr;
  int stem_wp = 0, stem_hp = 0, stem_ok = 0;
  alignas(64) unsigned char stem_tmap[128];
  std::vector<float*> act;                   // per buffer: device activation tensor (buffers with disjoint live ranges share an allocation)
  std::vector<float*> act_slots;             // the allocations behind `act`
  size_t act_bytes = 0, act_unshared_bytes = 0;   // allocated / what one allocation per buffer would take
  float* pooled = nullptr;                   // [max_crops, K] channel means
  float *d_uv = nullptr, *d_cov = nullptr, *d_mask = nullptr, *d_mask_logits = nullptr;
  int32_t* d_argmax = nullptr;
  struct GraphKey { int L, variant, backend, passes, persistent, multi, math, pair, pdl, halo; bool operator<(const GraphKey& o) const {
    return std::tie(L, variant, backend, passes, persistent, multi, math, pair, pdl, halo) < std::tie(o.L, o.variant, o.backend, o.passes, o.persistent, o.multi, o.math, o.pair, o.pdl, o.halo); } };
  std::map<GraphKey, cudaGraphExec_t> graphs;
  // resolution-level streams: independent branches of the hourglass (up1 at full resolution vs the low-resolution
  // sub-hourglass, hg.py:37-58) run concurrently; cross-stream edges are CUDA events (also inside graph capture)
  cudaStream_t side[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> op_done;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct DevScratch {   // growable device block (+ an optional pinned host mirror of it) for host-pointer calls
  void* d = nullptr; size_t dn = 0;
  void* h = nullptr; size_t hn = 0;
  int grow(suo_ctx* ctx, size_t n) {
    if (n <= dn) return SUO_OK;
    if (d) { cudaDeviceSynchronize(); cudaFree(d); }      // enqueued work (asynchronous submits included) may still use the old block
    dn = 0; d = nullptr;
    SUO_CUDA_TRY(ctx, cudaMalloc(&d, n));
    dn = n;
    return SUO_OK;
  }
  int mirror(suo_ctx* ctx, size_t n) {                    // every user of the mirror synchronises its stream before it returns
    if (n <= hn) return SUO_OK;
    if (h) cudaFreeHost(h);
    hn = 0; h = nullptr;
    const size_t cap = align_up(n + n / 2, 4096);
    SUO_CUDA_TRY(ctx, cudaHostAlloc(&h, cap, cudaHostAllocDefault));
    hn = cap;
    return SUO_OK;
  }
  void release() { if (d) cudaFree(d); if (h) cudaFreeHost(h); d = h = nullptr; dn = hn = 0; }
};

// Host-pointer calls with many SMALL arrays (pnp(), optimize(), a SLAM-mode view): a cudaMemcpyAsync from pageable memory costs 10-20 us
// whatever its size, and these calls make 6-30 of them.  The arrays are packed with memcpy into the pinned mirror of the device block at
// the offsets the device pointers have, and cross PCIe as ONE copy each way.
struct Stage {
  uint8_t* d0; uint8_t* h0;
  size_t in_lo = SIZE_MAX, in_hi = 0, out_lo = SIZE_MAX, out_hi = 0;
  Stage(const DevScratch& b) : d0(static_cast<uint8_t*>(b.d)), h0(static_cast<uint8_t*>(b.h)) {}
  template <typename T> void in(const T* dev, const T* src, size_t n) {
    const size_t o = reinterpret_cast<const uint8_t*>(dev) - d0;
    std::memcpy(h0 + o, src, n * sizeof(T));
    in_lo = std::min(in_lo, o); in_hi = std::max(in_hi, o + n * sizeof(T));
  }
  cudaError_t send(cudaStream_t s) const { return in_hi > in_lo ? cudaMemcpyAsync(d0 + in_lo, h0 + in_lo, in_hi - in_lo, cudaMemcpyHostToDevice, s) : cudaSuccess; }
  template <typename T> void want(const T* dev, size_t n) {
    const size_t o = reinterpret_cast<const uint8_t*>(dev) - d0;
    out_lo = std::min(out_lo, o); out_hi = std::max(out_hi, o + n * sizeof(T));
  }
  cudaError_t fetch(cudaStream_t s) const { return out_hi > out_lo ? cudaMemcpyAsync(h0 + out_lo, d0 + out_lo, out_hi - out_lo, cudaMemcpyDeviceToHost, s) : cudaSuccess; }
  template <typename T> void out(T* dst, const T* dev, size_t n) const { if (dst) std::memcpy(dst, h0 + (reinterpret_cast<const uint8_t*>(dev) - d0), n * sizeof(T)); }
};

struct CtxExtra {
  NetState net;
  DevScratch io;        // staged inputs/outputs of host-pointer calls
  DevScratch ba;        // BA err/level/fv scratch
  DevScratch fr;        // suo_frames workspace
  DevScratch bg;        // coupled-graph BA structure + workspace (ba_global.cu)
  int32_t* pnp_off = nullptr;   // row offsets c * num_kp of the gated keypoint lists (grown on demand)
  int pnp_off_n = 0;
  uint64_t* pnp_keys = nullptr; // RANSAC object keys 0, 1, 2, ... (a frame's crop index, also for the second crop group of a SLAM-mode frame)
  bool loaded = false;
  // double-buffered asynchronous frame batches (suo_frames_u8_submit / suo_frames_wait): per slot the staged inputs and
  // results and two events; one copy stream shared by the slots so that the host->device copy of batch i+1 overlaps batch i
  struct FrameSlot {
    DevScratch in, out;
    cudaEvent_t h2d_done = nullptr, done = nullptr;
    int* flag = nullptr;          // this batch's FP16-range flag: snapshot of the executor's flag taken (and the flag cleared) in stream order after the batch
    bool pending = false;
    cudaStream_t stream = nullptr; // the stream the pending batch runs on
  } slot[2];
  cudaStream_t copy_stream = nullptr, aux_stream = nullptr;
  // NCCL entry point for suo_allgather_results, resolved at first use from the libnccl the process already loaded
  void* nccl_allgather = nullptr;
  const double* last_err = nullptr;   // per-edge errors left by the most recent suo_ba_batch (device, inside `ba`)
  int last_err_n = 0;
};

CtxExtra* X(suo_ctx* c) { return reinterpret_cast<CtxExtra*>(c->net); }


// bump allocator over one device scratch block
struct Bump {
  uint8_t* base; size_t off = 0;
  template <typename T> T* take(size_t n) { off = align_up(off, 256); T* p = reinterpret_cast<T*>(base + off); off += n * sizeof(T); return p; }
};

// One conv op of the program as the engine sees it, for L crops.
void fill_conv_params(suo_ctx* ctx, NetState& N, size_t i, int L, int backend, int passes, ConvParams& p) {
  const OpDesc& o = N.ops[i];
  const BufDesc& bi = N.bufs[o.in];
  const BufDesc& bo = N.bufs[o.out];
  const int R = ctx->crop_res;
  p.in = N.act[o.in];
  p.w = N.pool + o.w_off;
  p.w_packed = N.packed[i];
  p.bias = N.pool + o.b_off;
  p.pre_scale = o.pre_off >= 0 ? N.pool + o.pre_off : nullptr;
  p.pre_shift = o.pre_off >= 0 ? N.pool + o.pre_off + o.Cin : nullptr;
  p.residual = o.res >= 0 ? N.act[o.res] : nullptr;
  p.out = N.act[o.out];
  p.B = L; p.H = R / bi.div; p.W = R / bi.div; p.Cin = o.Cin;
  p.Ho = R / bo.div; p.Wo = R / bo.div;
  p.Cout = o.Cout; p.Cout_pad = o.Cout_pad; p.out_c = bo.C; p.K = o.K; p.mode = o.mode;
  p.chunks_per_row = o.cpr; p.relu = o.relu; p.out_nchw = o.out_nchw;
  p.math = ctx->opt_math; p.w_packed16 = N.packed16[i]; p.range_flag = N.range_flag;
  const bool split_mode = backend == 1 && ctx->opt_math == 1 && ctx->opt_persistent && passes == 3;
  p.in_split = split_mode && bi.kind == 1;
  p.out_split = split_mode && bo.kind == 1;
  p.out_plane = (size_t)ctx->max_crops * p.Ho * p.Wo * bo.C;
  p.mma_merge = ctx->opt_mma_merge;
  p.trace = ctx->trace;
  const unsigned char* tm = N.tmaps[i].data();
  if (p.in_split) { memcpy(p.tmap_hi, tm, 128); memcpy(p.tmap_lo, tm + 128, 128); }
  p.epi_tma = split_mode && ctx->opt_epi_tma && N.epi_ok[i];
  if (p.epi_tma) { memcpy(p.tmap_out, tm + 256, 128); memcpy(p.tmap_out_lo, tm + 384, 128); memcpy(p.tmap_res, tm + 512, 128); }
  p.raw_tma = p.epi_tma && ctx->opt_raw_tma && N.raw_ok[i] && !p.in_split;
  if (const char* e = getenv("SUO_RAW_ONLY_OP")) { if (atoi(e) >= 0 && atoi(e) != (int)i) p.raw_tma = 0; }   // developer bisect switch
  if (p.raw_tma) memcpy(p.tmap_raw, tm + 640, 128);
  p.stem_raw = p.epi_tma && ctx->opt_stem_tma && N.stem_ok && o.mode == CONV_STEM7 && o.Cin == 4 && o.cpr == 1 && p.H == R;
  if (p.stem_raw) memcpy(p.tmap_raw, N.stem_tmap, 128);
  p.pair = p.epi_tma && ctx->opt_pair && N.wmap_ok[i];
  p.halo = p.pair && ctx->opt_halo && N.halo_ok[i];
  if (p.halo) { memcpy(p.tmap_hhi, tm + 896, 128); memcpy(p.tmap_hlo, tm + 1024, 128); }
  if (p.pair) memcpy(p.tmap_w, tm + 768, 128);
}

int run_program(suo_ctx* ctx, int L, int variant, int backend, int passes, cudaStream_t s) {
  NetState& N = X(ctx)->net;
  const int R = ctx->crop_res;
  const bool multi = ctx->opt_multistream != 0 && !ctx->opt_act_reuse;      // shared activation slots assume program order on ONE stream
  auto level_of = [&](const OpDesc& o) { const int d = N.bufs[o.out].div; return !multi ? 0 : (d <= 4 ? 0 : (d == 8 ? 1 : 2)); };
  cudaStream_t streams[3] = {s, s, s};
  if (multi) {
    for (int k = 0; k < 2; ++k)
      if (!N.side[k]) SUO_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&N.side[k], cudaStreamNonBlocking));
    streams[1] = N.side[0]; streams[2] = N.side[1];
    if (N.op_done.size() != N.ops.size()) {
      N.op_done.resize(N.ops.size());
      for (auto& e : N.op_done) SUO_CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
  }
  std::vector<int> writer(N.bufs.size(), -1);
  for (size_t i = 0; i < N.ops.size(); ++i) {
    const OpDesc& o = N.ops[i];
    if (o.variant != 2 && o.variant != variant) continue;
    const BufDesc& bi = N.bufs[o.in];
    const BufDesc& bo = N.bufs[o.out];
    const int lvl = level_of(o);
    cudaStream_t st = streams[lvl];
    if (multi) {
      const int deps[2] = {o.in, o.res};
      for (int b : deps) {
        if (b < 0 || writer[b] < 0) continue;
        if (level_of(N.ops[writer[b]]) != lvl) SUO_CUDA_TRY(ctx, cudaStreamWaitEvent(st, N.op_done[writer[b]], 0));
      }
    }
    int rc = SUO_OK;
    if (o.type == OP_CONV) {
      ConvParams p{};
      fill_conv_params(ctx, N, i, L, backend, passes, p);
      rc = backend == 1 ? launch_conv_tc(ctx, p, passes, st) : launch_conv_simt(ctx, p, st);
    } else if (o.type == OP_MAXPOOL) {
      rc = launch_maxpool2(ctx, N.act[o.in], L, R / bi.div, R / bi.div, bi.C, N.act[o.out], st);
    } else if (o.type == OP_UPADD) {
      rc = launch_upsample_add(ctx, N.act[o.in], N.act[o.res], L, R / bo.div, R / bo.div, bo.C, N.act[o.out], st);
    } else {
      ctx->set_error("unknown op type in program", __FILE__, __LINE__);
      return SUO_E_INVALID;
    }
    if (rc != SUO_OK) return rc;
    writer[o.out] = (int)i;
    if (multi) {
      // record completion if a later op on another level stream consumes this output
      const int produced = N.ops[i].out;
      bool cross = false;
      for (size_t k = i + 1; k < N.ops.size() && !cross; ++k) {
        const OpDesc& c = N.ops[k];
        if (c.variant != 2 && c.variant != variant) continue;
        if ((c.in == produced || c.res == produced) && level_of(c) != lvl) cross = true;
        if (c.out == produced) break;
      }
      if (cross) SUO_CUDA_TRY(ctx, cudaEventRecord(N.op_done[i], st));
    }
  }
  return SUO_OK;
}

int run_network(suo_ctx* ctx, int L, int variant, cudaStream_t s) {
  NetState& N = X(ctx)->net;
  if (N.w_exceeds_fp16 && ctx->opt_backend == 1 && ctx->opt_math == 1) {
    ctx->set_error("fp16x3 conv math: a BN-folded weight exceeds the FP16 range (|w| > 6e4) or is not finite; use SUO_OPT_CONV_MATH = 0 (tf32x3)", __FILE__, __LINE__);
    return SUO_E_RANGE;
  }
  const int backend = ctx->opt_backend, passes = ctx->opt_passes;
  if (!ctx->opt_graph) return run_program(ctx, L, variant, backend, passes, s);
  NetState::GraphKey key{L, variant, backend, passes, ctx->opt_persistent, ctx->opt_multistream, ctx->opt_math, ctx->opt_pair, ctx->opt_pdl, ctx->opt_halo};
  auto it = N.graphs.find(key);
  if (it == N.graphs.end()) {
    // warm the kernels once outside capture (cudaFuncSetAttribute etc.), then capture
    int rc = run_program(ctx, L, variant, backend, passes, s);
    if (rc != SUO_OK) return rc;
    SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    cudaStream_t cs;
    SUO_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    SUO_CUDA_TRY(ctx, cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    const long long before = ctx->launches;
    rc = run_program(ctx, L, variant, backend, passes, cs);
    ctx->launches = before;   // captured, not launched
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(cs, &g);
    if (rc != SUO_OK || e != cudaSuccess) {
      if (g) cudaGraphDestroy(g);
      cudaStreamDestroy(cs);
      if (rc == SUO_OK) { ctx->set_error(std::string("graph capture: ") + cudaGetErrorString(e), __FILE__, __LINE__); rc = SUO_E_CUDA; }
      return rc;
    }
    cudaGraphExec_t ge;
    SUO_CUDA_TRY(ctx, cudaGraphInstantiate(&ge, g, 0));
    cudaGraphDestroy(g);
    cudaStreamDestroy(cs);
    it = N.graphs.emplace(key, ge).first;
    return SUO_OK;   // the warm-up run above already produced this call's result
  }
  SUO_CUDA_TRY(ctx, cudaGraphLaunch(it->second, s));
  // count the kernels the graph replays
  long long n = 0;
  for (size_t i = 0; i < N.ops.size(); ++i) {
    const OpDesc& o = N.ops[i];
    if (o.variant != 2 && o.variant != variant) continue;
    ++n;
  }
  ctx->launches += n;
  return SUO_OK;
}

int check_ctx(suo_ctx* ctx) {
  if (!ctx) return SUO_E_INVALID;
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) { ctx->set_error(cudaGetErrorString(e), __FILE__, __LINE__); return SUO_E_CUDA; }
  return SUO_OK;
}

}  // namespace

extern "C" {

size_t suo_record_bytes(int num_kp);

int suo_create(int device, int max_crops, int crop_res, int num_kp, suo_ctx** out) {
  if (!out || max_crops <= 0 || crop_res < 64 || (crop_res & (crop_res - 1)) || num_kp <= 0 || num_kp > 45) return SUO_E_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SUO_E_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return SUO_E_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SUO_E_CUDA;
  if (prop.major != 10) return SUO_E_CUDA;   // sm_100a only: there is no fallback path
  suo_ctx* c = new suo_ctx();
  c->device = device; c->max_crops = max_crops; c->crop_res = crop_res; c->num_kp = num_kp;
  c->net = new CtxExtra();
  // developer overrides (tests run both conv kernel variants through the same ABI)
  if (const char* e = getenv("SUO_CONV_PERSISTENT")) c->opt_persistent = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SUO_USE_GRAPH")) c->opt_graph = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SUO_MULTISTREAM")) c->opt_multistream = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SUO_ACT_REUSE")) c->opt_act_reuse = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SUO_CONV_MATH")) c->opt_math = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SUO_EPI_TMA")) c->opt_epi_tma = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SUO_PAIR")) c->opt_pair = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SUO_PDL")) c->opt_pdl = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SUO_STEM_TMA")) c->opt_stem_tma = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SUO_HALO")) c->opt_halo = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SUO_MMA_MERGE")) c->opt_mma_merge = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SUO_RAW_TMA")) c->opt_raw_tma = atoi(e) ? 1 : 0;
  if (const char* e = getenv("SUO_GRID_CAP")) c->opt_grid_cap = atoi(e);
  if (getenv("SUO_TRACE")) {
    if (cudaMalloc(&c->trace, (1 + 4 * 8192) * sizeof(unsigned long long)) == cudaSuccess) cudaMemset(c->trace, 0, (1 + 4 * 8192) * sizeof(unsigned long long));
    else c->trace = nullptr;
  }
  *out = c;
  return SUO_OK;
}

void suo_destroy(suo_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->trace) {
    std::vector<unsigned long long> h(1 + 4 * 8192);
    cudaDeviceSynchronize();
    if (cudaMemcpy(h.data(), ctx->trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
      const std::string path = std::string(getenv("SUO_TRACE") ? getenv("SUO_TRACE") : "trace") + "." + std::to_string((unsigned long long)(uintptr_t)ctx) + ".csv";
      if (FILE* f = fopen(path.c_str(), "w")) {
        fprintf(f, "launch,t_start_ns,t_end_ns,grid,smid\n");
        for (unsigned long long k = 0; k < std::min<unsigned long long>(h[0], 8192); ++k)
          fprintf(f, "%llu,%llu,%llu,%llu,%llu\n", k, h[1 + 4 * k], h[2 + 4 * k], h[3 + 4 * k], h[4 + 4 * k]);
        fclose(f);
      }
    }
    cudaFree(ctx->trace);
  }
  CtxExtra* x = X(ctx);
  if (x) {
    NetState& N = x->net;
    for (auto& kv : N.graphs) cudaGraphExecDestroy(kv.second);
    for (auto& e : N.op_done) cudaEventDestroy(e);
    for (auto& st : N.side) if (st) cudaStreamDestroy(st);
    for (float* p : N.packed) if (p) cudaFree(p);
    for (uint16_t* p : N.packed16) if (p) cudaFree(p);
    if (N.range_flag) cudaFree(N.range_flag);
    if (N.stem_pad) cudaFree(N.stem_pad);
    for (float* p : N.act_slots) if (p) cudaFree(p);
    if (N.pool) cudaFree(N.pool);
    if (N.pooled) cudaFree(N.pooled);
    if (N.d_uv) cudaFree(N.d_uv);
    for (auto& sl : x->slot) {
      sl.in.release(); sl.out.release();
      if (sl.h2d_done) cudaEventDestroy(sl.h2d_done);
      if (sl.done) cudaEventDestroy(sl.done);
      if (sl.flag) cudaFree(sl.flag);
    }
    if (x->copy_stream) cudaStreamDestroy(x->copy_stream);
    if (x->aux_stream) cudaStreamDestroy(x->aux_stream);
    x->io.release(); x->ba.release(); x->fr.release(); x->bg.release();
    if (x->pnp_off) cudaFree(x->pnp_off);
    if (x->pnp_keys) cudaFree(x->pnp_keys);
    delete x;
  }
  delete ctx;
}

const char* suo_last_error(const suo_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
long long suo_kernel_launches(const suo_ctx* ctx) { return ctx ? ctx->launches : 0; }
size_t suo_activation_bytes(const suo_ctx* ctx, size_t* unshared) {
  if (!ctx || !ctx->net) return 0;
  const NetState& N = reinterpret_cast<const CtxExtra*>(ctx->net)->net;
  if (unshared) *unshared = N.act_unshared_bytes;
  return N.act_bytes;
}

int suo_set_option(suo_ctx* ctx, int option, int value) {
  if (!ctx) return SUO_E_INVALID;
  switch (option) {
    case SUO_OPT_CONV_BACKEND: if (value != 0 && value != 1) return SUO_E_INVALID; ctx->opt_backend = value; return SUO_OK;
    case SUO_OPT_TF32_PASSES: if (value != 1 && value != 3) return SUO_E_INVALID; ctx->opt_passes = value; return SUO_OK;
    case SUO_OPT_USE_GRAPH: ctx->opt_graph = value ? 1 : 0; return SUO_OK;
    case SUO_OPT_CONV_PERSISTENT: ctx->opt_persistent = value ? 1 : 0; return SUO_OK;
    case SUO_OPT_MULTISTREAM: ctx->opt_multistream = value ? 1 : 0; return SUO_OK;
    case SUO_OPT_CONV_MATH: if (value != 0 && value != 1) return SUO_E_INVALID; ctx->opt_math = value; return SUO_OK;
    case SUO_OPT_CONV_FUSE: return value == 0 ? SUO_OK : SUO_E_INVALID;      // the fused conv2 + conv3 kernels of round 1 were slower and are gone
    case SUO_OPT_CONV_PAIR: ctx->opt_pair = value ? 1 : 0; return SUO_OK;
    case SUO_OPT_PDL: ctx->opt_pdl = value ? 1 : 0; return SUO_OK;
    case SUO_OPT_CONV_HALO: ctx->opt_halo = value ? 1 : 0; return SUO_OK;
    case SUO_OPT_PNP_MAX_POINTS: if (value < 4 || value > 4096) return SUO_E_INVALID; ctx->opt_pnp_max_pts = value; return SUO_OK;
    case SUO_OPT_BA_BLOCK_DIAGONAL: if (value < 0 || value > 2) return SUO_E_INVALID; ctx->opt_ba_blockdiag = value; return SUO_OK;
    case SUO_OPT_SLAM_SFM: ctx->opt_slam_sfm = value ? 1 : 0; return SUO_OK;
    default: return SUO_E_INVALID;
  }
}

int suo_load_weights(suo_ctx* ctx, const void* blob, size_t nbytes) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  CtxExtra* x = X(ctx);
  if (x->loaded) { ctx->set_error("weights already loaded (create a new ctx)", __FILE__, __LINE__); return SUO_E_STATE; }
  const uint8_t* b = static_cast<const uint8_t*>(blob);
  if (nbytes < sizeof(BlobHeader)) return SUO_E_INVALID;
  NetState& N = x->net;
  memcpy(&N.h, b, sizeof(BlobHeader));
  if (N.h.magic != kMagic || N.h.version != 2 || N.h.num_kp != ctx->num_kp) {
    ctx->set_error("bad weight blob header", __FILE__, __LINE__);
    return SUO_E_INVALID;
  }
  size_t off = sizeof(BlobHeader);
  const size_t need = off + sizeof(BufDesc) * N.h.n_bufs + sizeof(OpDesc) * N.h.n_ops + sizeof(float) * (size_t)N.h.n_floats;
  if (nbytes < need) { ctx->set_error("weight blob truncated", __FILE__, __LINE__); return SUO_E_INVALID; }
  N.bufs.resize(N.h.n_bufs);
  memcpy(N.bufs.data(), b + off, sizeof(BufDesc) * N.h.n_bufs); off += sizeof(BufDesc) * N.h.n_bufs;
  N.ops.resize(N.h.n_ops);
  memcpy(N.ops.data(), b + off, sizeof(OpDesc) * N.h.n_ops); off += sizeof(OpDesc) * N.h.n_ops;
  const float* pool_h = reinterpret_cast<const float*>(b + off);
  {   // a corrupt or stale .suo file must be rejected here, not index out of bounds later
    const long long nf = N.h.n_floats, nb = N.h.n_bufs, Kp = ctx->num_kp;
    auto in_pool = [&](long long o, long long n) { return o >= 0 && n >= 0 && o + n <= nf; };
    auto buf_ok = [&](long long i) { return i >= 0 && i < nb; };
    bool ok = buf_ok(N.h.in_buf_noprior) && buf_ok(N.h.in_buf_prior) && buf_ok(N.h.logits_buf) && in_pool(N.h.cls_w_off, Kp * Kp) &&
              in_pool(N.h.cls_b_off, Kp) && N.h.heat_div > 0 && ctx->crop_res % N.h.heat_div == 0;
    for (const BufDesc& d : N.bufs) ok = ok && d.div > 0 && d.div <= ctx->crop_res && ctx->crop_res % d.div == 0 && d.C > 0 && d.C <= 4096;
    for (const OpDesc& o : N.ops) {
      if (!ok) break;
      ok = buf_ok(o.in) && buf_ok(o.out) && (o.res == -1 || buf_ok(o.res)) && o.in != o.out;
      if (!ok) break;
      if (o.type == OP_CONV) {
        ok = (o.mode == CONV_1x1 || o.mode == CONV_3x3 || o.mode == CONV_STEM7) && o.Cin == N.bufs[o.in].C && o.Cout > 0 && o.Cout <= o.Cout_pad &&
             o.Cout_pad % 64 == 0 && o.Cout <= N.bufs[o.out].C + 63 && o.K > 0 && o.K % 32 == 0 && in_pool(o.w_off, (long long)o.Cout_pad * o.K) &&
             in_pool(o.b_off, o.Cout_pad) && (o.pre_off == -1 || in_pool(o.pre_off, 2LL * o.Cin)) &&
             (o.res == -1 || (N.bufs[o.res].C == N.bufs[o.out].C && N.bufs[o.res].div == N.bufs[o.out].div)) &&
             (o.mode == CONV_1x1 ? o.K == o.Cin : o.mode == CONV_3x3 ? o.K == 9 * o.Cin : (o.cpr > 0 && o.K % (o.cpr * 32) == 0)) &&
             (o.mode == CONV_STEM7 ? N.bufs[o.out].div == 2 * N.bufs[o.in].div : N.bufs[o.out].div == N.bufs[o.in].div);
      } else if (o.type == OP_MAXPOOL) {
        ok = N.bufs[o.out].C == N.bufs[o.in].C && N.bufs[o.out].div == 2 * N.bufs[o.in].div;
      } else if (o.type == OP_UPADD) {
        ok = o.res >= 0 && N.bufs[o.out].C == N.bufs[o.in].C && N.bufs[o.res].C == N.bufs[o.in].C && N.bufs[o.out].div == N.bufs[o.in].div &&
             N.bufs[o.res].div == 2 * N.bufs[o.in].div;
      } else {
        ok = false;
      }
    }
    if (!ok) { N.bufs.clear(); N.ops.clear(); ctx->set_error("weight blob: buffer / op table is inconsistent (corrupt or stale file)", __FILE__, __LINE__); return SUO_E_INVALID; }
  }
  SUO_CUDA_TRY(ctx, cudaMalloc(&N.pool, sizeof(float) * (size_t)N.h.n_floats));
  SUO_CUDA_TRY(ctx, cudaMemcpy(N.pool, pool_h, sizeof(float) * (size_t)N.h.n_floats, cudaMemcpyHostToDevice));
  // tensor-core weight images
  N.packed.assign(N.ops.size(), nullptr);
  N.packed16.assign(N.ops.size(), nullptr);
  std::vector<float> tmp;
  std::vector<uint16_t> tmp16;
  for (size_t i = 0; i < N.ops.size(); ++i) {
    const OpDesc& o = N.ops[i];
    if (o.type != OP_CONV) continue;
    const size_t nf = conv_tc_packed_floats(o.Cout_pad, o.K);
    tmp.resize(nf);
    conv_tc_pack_weights(pool_h + o.w_off, o.Cout_pad, o.K, tmp.data());
    SUO_CUDA_TRY(ctx, cudaMalloc(&N.packed[i], nf * sizeof(float)));
    SUO_CUDA_TRY(ctx, cudaMemcpy(N.packed[i], tmp.data(), nf * sizeof(float), cudaMemcpyHostToDevice));
    {   // fp16x3 splits every weight into two FP16 numbers: a BN-folded weight beyond the FP16 range cannot be represented
      float wmax = 0.f;
      const float* wp = pool_h + o.w_off;
      for (size_t q = 0; q < (size_t)o.Cout_pad * o.K; ++q) wmax = std::max(wmax, std::fabs(wp[q]));
      if (!(wmax <= 6.0e4f)) N.w_exceeds_fp16 = true;
    }
    if (o.K % 64 == 0) {
      const size_t nh = conv_tc_packed16_halfs(o.Cout_pad, o.K);
      tmp16.resize(nh);
      conv_tc_pack_weights_f16(pool_h + o.w_off, o.Cout_pad, o.K, tmp16.data());
      SUO_CUDA_TRY(ctx, cudaMalloc(&N.packed16[i], nh * sizeof(uint16_t)));
      SUO_CUDA_TRY(ctx, cudaMemcpy(N.packed16[i], tmp16.data(), nh * sizeof(uint16_t), cudaMemcpyHostToDevice));
    }
  }
  SUO_CUDA_TRY(ctx, cudaMalloc(&N.range_flag, sizeof(int)));
  SUO_CUDA_TRY(ctx, cudaMemset(N.range_flag, 0, sizeof(int)));
  // activation buffers
  N.act.assign(N.bufs.size(), nullptr);
  const int R = ctx->crop_res;
  {
    // Liveness-based reuse: a buffer lives from its first writer to its last reader in program order (the op list runs in that order on one
    // stream; kernels launched with programmatic dependent launch touch no activation before griddepcontrol.wait, i.e. before every earlier
    // kernel has completed).  Buffers whose live ranges do not overlap share one allocation ("slot"): ~10 slots instead of ~200 buffers.
    // The two network-input buffers (written before the program, partly by stamping) and the logits (read after it) keep their own.
    const int nb = (int)N.bufs.size();
    std::vector<int> first(nb, INT32_MAX), last(nb, -1);
    std::vector<size_t> bytes(nb);
    for (int i = 0; i < nb; ++i) { const size_t side = R / N.bufs[i].div; bytes[i] = (size_t)ctx->max_crops * side * side * N.bufs[i].C * sizeof(float); }
    for (int i = 0; i < (int)N.ops.size(); ++i) {
      const OpDesc& o = N.ops[i];
      last[o.in] = std::max(last[o.in], i);
      if (o.res >= 0) last[o.res] = std::max(last[o.res], i);
      first[o.out] = std::min(first[o.out], i); last[o.out] = std::max(last[o.out], i);
    }
    std::vector<char> own(nb, 0);
    own[N.h.in_buf_noprior] = own[N.h.in_buf_prior] = own[N.h.logits_buf] = 1;
    for (int i = 0; i < nb; ++i) if (first[i] == INT32_MAX || last[i] < 0 || !ctx->opt_act_reuse) own[i] = 1;     // never written / never read by the program
    std::vector<int> order;
    for (int i = 0; i < nb; ++i) if (!own[i]) order.push_back(i);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return first[a] < first[b]; });
    struct Slot { size_t cap; int busy_until; };
    std::vector<Slot> slots;
    std::vector<int> slot_of(nb, -1);
    for (int b : order) {
      int best = -1;
      for (int q = 0; q < (int)slots.size(); ++q) {
        if (slots[q].busy_until >= first[b]) continue;                       // closed intervals: an op's output never aliases its inputs
        if (best < 0) { best = q; continue; }
        const bool fits_q = slots[q].cap >= bytes[b], fits_b = slots[best].cap >= bytes[b];
        if ((fits_q && (!fits_b || slots[q].cap < slots[best].cap)) || (!fits_q && !fits_b && slots[q].cap > slots[best].cap)) best = q;
      }
      if (best < 0) { slots.push_back({bytes[b], last[b]}); best = (int)slots.size() - 1; }
      else { slots[best].cap = std::max(slots[best].cap, bytes[b]); slots[best].busy_until = last[b]; }
      slot_of[b] = best;
    }
    N.act_slots.assign(slots.size(), nullptr);
    N.act_bytes = 0;
    for (size_t q = 0; q < slots.size(); ++q) {
      SUO_CUDA_TRY(ctx, cudaMalloc(&N.act_slots[q], slots[q].cap));
      SUO_CUDA_TRY(ctx, cudaMemset(N.act_slots[q], 0, slots[q].cap));
      N.act_bytes += slots[q].cap;
    }
    for (int i = 0; i < nb; ++i) {
      if (own[i]) {
        float* p = nullptr;
        SUO_CUDA_TRY(ctx, cudaMalloc(&p, bytes[i]));
        SUO_CUDA_TRY(ctx, cudaMemset(p, 0, bytes[i]));
        N.act_slots.push_back(p);
        N.act[i] = p;
        N.act_bytes += bytes[i];
      } else {
        N.act[i] = N.act_slots[slot_of[i]];
      }
    }
    N.act_unshared_bytes = 0;
    for (int i = 0; i < nb; ++i) N.act_unshared_bytes += bytes[i];
  }
  N.tmaps.resize(N.ops.size());
  N.epi_ok.assign(N.ops.size(), 0);
  N.raw_ok.assign(N.ops.size(), 0);
  N.wmap_ok.assign(N.ops.size(), 0);
  N.halo_ok.assign(N.ops.size(), 0);
  for (size_t i = 0; i < N.ops.size(); ++i) {
    const OpDesc& o = N.ops[i];
    if (o.type != OP_CONV) continue;
    if (o.Cout_pad % 128 == 0 && N.packed16[i]) {                        // weight images by TMA for the CTA-pair kernels
      int rcw = make_weight_tmap(ctx, N.tmaps[i].data() + 768, N.packed16[i], conv_tc_packed16_halfs(o.Cout_pad, o.K));
      if (rcw) return rcw;
      N.wmap_ok[i] = 1;
    }
    if (N.bufs[o.in].kind == 1) {
      const int side = R / N.bufs[o.in].div, C = N.bufs[o.in].C;
      const uint16_t* base = reinterpret_cast<const uint16_t*>(N.act[o.in]);
      const size_t plane = (size_t)ctx->max_crops * side * side * C;
      int rc2 = make_plane_tmap(ctx, N.tmaps[i].data(), base, C, side, side, ctx->max_crops);
      if (rc2) return rc2;
      rc2 = make_plane_tmap(ctx, N.tmaps[i].data() + 128, base + plane, C, side, side, ctx->max_crops);
      if (rc2) return rc2;
      if (o.mode == CONV_3x3 && (side == 16 || side == 32 || side == 64) && C % 64 == 0 &&
          !make_plane_tmap_halo(ctx, N.tmaps[i].data() + 896, base, C, side, side, ctx->max_crops) &&
          !make_plane_tmap_halo(ctx, N.tmaps[i].data() + 1024, base + plane, C, side, side, ctx->max_crops))
        N.halo_ok[i] = 1;
    }
    // output (and skip) maps of the TMA-store epilogue: rows = every pixel of every crop slot of the buffer
    const BufDesc& bo = N.bufs[o.out];
    if (o.out_nchw || bo.C % 8) continue;
    const int so = R / bo.div;
    const size_t rows = (size_t)ctx->max_crops * so * so;
    int rc2;
    if (bo.kind == 1) {
      const uint16_t* ob = reinterpret_cast<const uint16_t*>(N.act[o.out]);
      rc2 = make_rows_tmap(ctx, N.tmaps[i].data() + 256, ob, true, rows, bo.C);
      if (!rc2) rc2 = make_rows_tmap(ctx, N.tmaps[i].data() + 384, ob + rows * bo.C, true, rows, bo.C);
    } else {
      rc2 = make_rows_tmap(ctx, N.tmaps[i].data() + 256, N.act[o.out], false, rows, bo.C);
    }
    if (!rc2 && o.res >= 0) rc2 = make_rows_tmap(ctx, N.tmaps[i].data() + 512, N.act[o.res], false, rows, bo.C);
    if (rc2) return rc2;
    N.epi_ok[i] = 1;
    if (o.mode == CONV_1x1 && o.Cin % 64 == 0 && N.bufs[o.in].C == o.Cin) {      // FP32 input [pixels, Cin] fetched by TMA (plan 3)
      const int si = R / N.bufs[o.in].div;
      rc2 = make_rows_tmap(ctx, N.tmaps[i].data() + 640, N.act[o.in], false, (size_t)ctx->max_crops * si * si, o.Cin, 128);
      if (rc2) return rc2;
      N.raw_ok[i] = 1;
    }
  }
  // TMA-fed stem: zero-bordered input copy + tensor map of the overlapping windows (dimension 1 = output column, stride 2 pixels = 32 B,
  // 32 floats = one kernel row of 7 pixels + 1; dimension 2 = bordered input row).  Needs whole output rows per 128-pixel tile.
  N.stem_ok = 0;
  if ((R / 2) % 128 == 0) {
    N.stem_wp = R + 8; N.stem_hp = R + 6;
    const size_t nb = (size_t)ctx->max_crops * N.stem_hp * N.stem_wp * 16;
    SUO_CUDA_TRY(ctx, cudaMalloc(&N.stem_pad, nb));
    SUO_CUDA_TRY(ctx, cudaMemset(N.stem_pad, 0, nb));
    int rcs = make_plane_tmap(ctx, nullptr, nullptr, 0, 0, 0, 0);   // resolves the driver entry point
    if (!rcs) {
      CUtensorMap m;
      const cuuint64_t dims[4] = {32, (cuuint64_t)(R / 2), (cuuint64_t)N.stem_hp, (cuuint64_t)ctx->max_crops};
      const cuuint64_t strides[3] = {32, (cuuint64_t)N.stem_wp * 16, (cuuint64_t)N.stem_hp * N.stem_wp * 16};
      const cuuint32_t box[4] = {32, 128, 1, 1};
      const cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = g_encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, N.stem_pad, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r == CUDA_SUCCESS) { memcpy(N.stem_tmap, &m, 128); N.stem_ok = 1; }     // a driver that rejects overlapping strides: register-gather stem
    }
  }
  const size_t LK = (size_t)ctx->max_crops * ctx->num_kp;
  SUO_CUDA_TRY(ctx, cudaMalloc(&N.pooled, LK * sizeof(float)));
  SUO_CUDA_TRY(ctx, cudaMalloc(&N.d_uv, LK * (2 + 4 + 1 + 1 + 1) * sizeof(float)));
  N.d_cov = N.d_uv + LK * 2; N.d_mask = N.d_cov + LK * 4; N.d_mask_logits = N.d_mask + LK;
  N.d_argmax = reinterpret_cast<int32_t*>(N.d_mask_logits + LK);
  x->loaded = true;
  return SUO_OK;
}

int suo_heatmap_reduce(suo_ctx* ctx, const float* logits, int B, int K, int H, int W, const float* cls_w,
                       const float* cls_b, float* uv, float* cov, float* prob, float* mask_logits, float* mask,
                       int32_t* argmax, int on_device, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (B <= 0 || K <= 0 || !logits) return SUO_E_INVALID;
  CtxExtra* x = X(ctx);
  const size_t n = (size_t)B * K * H * W, BK = (size_t)B * K;
  if (on_device) {
    // pooled scratch
    rc = x->ba.grow(ctx, BK * sizeof(float));
    if (rc) return rc;
    return launch_heatmap_reduce(ctx, logits, B, K, H, W, cls_w, cls_b, static_cast<float*>(x->ba.d), uv, cov, prob,
                                 mask_logits, mask, argmax, s);
  }
  size_t bytes = (n * (prob ? 2 : 1) + BK * 16 + (size_t)K * K + K) * sizeof(float) + 4096;
  rc = x->io.grow(ctx, bytes);
  if (rc) return rc;
  Bump bp{static_cast<uint8_t*>(x->io.d)};
  float* d_log = bp.take<float>(n);
  float* d_prob = prob ? bp.take<float>(n) : nullptr;
  float* d_pooled = bp.take<float>(BK);
  float* d_uv = bp.take<float>(BK * 2);
  float* d_cov = bp.take<float>(BK * 4);
  float* d_ml = bp.take<float>(BK);
  float* d_m = bp.take<float>(BK);
  int32_t* d_am = bp.take<int32_t>(BK);
  float* d_w = cls_w ? bp.take<float>((size_t)K * K) : nullptr;
  float* d_b = cls_b ? bp.take<float>(K) : nullptr;
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_log, logits, n * sizeof(float), cudaMemcpyHostToDevice, s));
  if (d_w) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_w, cls_w, (size_t)K * K * sizeof(float), cudaMemcpyHostToDevice, s));
  if (d_b) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_b, cls_b, K * sizeof(float), cudaMemcpyHostToDevice, s));
  rc = launch_heatmap_reduce(ctx, d_log, B, K, H, W, d_w, d_b, d_pooled, d_uv, d_cov, d_prob, d_ml, d_m, d_am, s);
  if (rc) return rc;
  if (uv) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(uv, d_uv, BK * 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (cov) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(cov, d_cov, BK * 4 * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (prob) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(prob, d_prob, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (mask_logits && d_w) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(mask_logits, d_ml, BK * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (mask && d_w) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(mask, d_m, BK * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (argmax) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(argmax, d_am, BK * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return SUO_OK;
}

int suo_crop_concat(suo_ctx* ctx, const float* images, int n_img, int H, int W, const float* boxes,
                    const int32_t* box_img, int L, const float* priors, int R, float* out, int out_c, int on_device,
                    void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (L <= 0 || n_img <= 0 || !images || !boxes || !box_img || !out) return SUO_E_INVALID;
  if (on_device) return launch_crop_concat(ctx, images, n_img, H, W, boxes, box_img, L, priors, ctx->num_kp, R, out, out_c, s);
  for (int c = 0; c < L; ++c) if (box_img[c] < 0 || box_img[c] >= n_img) { ctx->set_error("suo_crop_concat: box_img out of range", __FILE__, __LINE__); return SUO_E_INVALID; }
  CtxExtra* x = X(ctx);
  const size_t n_im = (size_t)n_img * 3 * H * W, n_out = (size_t)L * R * R * out_c, n_pr = priors ? (size_t)L * ctx->num_kp * R * R : 0;
  rc = x->io.grow(ctx, (n_im + n_out + n_pr + 8 * (size_t)L) * sizeof(float) + 4096);
  if (rc) return rc;
  Bump bp{static_cast<uint8_t*>(x->io.d)};
  float* d_im = bp.take<float>(n_im);
  float* d_box = bp.take<float>(4 * (size_t)L);
  int32_t* d_bi = bp.take<int32_t>(L);
  float* d_pr = priors ? bp.take<float>(n_pr) : nullptr;
  float* d_out = bp.take<float>(n_out);
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_im, images, n_im * sizeof(float), cudaMemcpyHostToDevice, s));
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_box, boxes, 4 * (size_t)L * sizeof(float), cudaMemcpyHostToDevice, s));
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_bi, box_img, L * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  if (d_pr) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_pr, priors, n_pr * sizeof(float), cudaMemcpyHostToDevice, s));
  SUO_CUDA_TRY(ctx, cudaMemsetAsync(d_out, 0, n_out * sizeof(float), s));
  rc = launch_crop_concat(ctx, d_im, n_img, H, W, d_box, d_bi, L, d_pr, ctx->num_kp, R, d_out, out_c, s);
  if (rc) return rc;
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(out, d_out, n_out * sizeof(float), cudaMemcpyDeviceToHost, s));
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return SUO_OK;
}

int suo_conv2d(suo_ctx* ctx, const float* in, int B, int H, int W, int Cin, const float* w, const float* bias,
               int Cout, int ksize, int stride, const float* pre_scale, const float* pre_shift,
               const float* residual, int relu, float* out, int backend, int tf32_passes, int on_device, void* stream) {
  // Test / bench hook: host pointers only (packs weights on the fly).
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if (on_device) { ctx->set_error("suo_conv2d takes host pointers", __FILE__, __LINE__); return SUO_E_INVALID; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ConvParams p{};
  if (ksize == 1 && stride == 1) p.mode = CONV_1x1;
  else if (ksize == 3 && stride == 1) p.mode = CONV_3x3;
  else if (ksize == 7 && stride == 2) p.mode = CONV_STEM7;
  else return SUO_E_INVALID;
  if (Cin % 4 || (p.mode != CONV_STEM7 && Cin % 32)) return SUO_E_INVALID;
  const int Ho = stride == 2 ? H / 2 : H, Wo = stride == 2 ? W / 2 : W;
  const int Cout_pad = (int)align_up(Cout, 64);
  int K, cpr = 0;
  if (p.mode == CONV_1x1) K = Cin;
  else if (p.mode == CONV_3x3) K = 9 * Cin;
  else if (7 * Cin <= 32) { cpr = 1; K = 8 * 32; }              // RGB-only layout: one 32-float chunk per kernel row, an eighth zero row pads K to 256 (weights.py)
  else { cpr = 2 * ((7 * Cin + 63) / 64); K = 7 * cpr * 32; }   // kernel-row stride: multiple of 64 floats (both chunk widths)
  // canonical weights [Cout_pad][K] in gather order from [Cout][kh][kw][Cin]
  std::vector<float> wc((size_t)Cout_pad * K, 0.f), bc(Cout_pad, 0.f);
  for (int co = 0; co < Cout; ++co) {
    if (bias) bc[co] = bias[co];
    for (int ky = 0; ky < ksize; ++ky)
      for (int kx = 0; kx < ksize; ++kx)
        for (int ci = 0; ci < Cin; ++ci) {
          const float v = w[(((size_t)co * ksize + ky) * ksize + kx) * Cin + ci];
          size_t k;
          if (p.mode == CONV_STEM7) k = (size_t)ky * cpr * 32 + (size_t)kx * Cin + ci;
          else k = (size_t)(ky * ksize + kx) * Cin + ci;
          wc[(size_t)co * K + k] = v;
        }
  }
  std::vector<float> wp(conv_tc_packed_floats(Cout_pad, K));
  conv_tc_pack_weights(wc.data(), Cout_pad, K, wp.data());
  std::vector<uint16_t> wp16;
  if (K % 64 == 0) { wp16.resize(conv_tc_packed16_halfs(Cout_pad, K)); conv_tc_pack_weights_f16(wc.data(), Cout_pad, K, wp16.data()); }
  const size_t n_in = (size_t)B * H * W * Cin, n_out = (size_t)B * Ho * Wo * Cout;
  CtxExtra* x = X(ctx);
  rc = x->io.grow(ctx, (n_in + 2 * n_out + wc.size() + wp.size() + wp16.size() / 2 + bc.size() + 2 * (size_t)Cin) * sizeof(float) + 16384);
  if (rc) return rc;
  Bump bp{static_cast<uint8_t*>(x->io.d)};
  float* d_in = bp.take<float>(n_in);
  float* d_out = bp.take<float>(n_out);
  float* d_res = residual ? bp.take<float>(n_out) : nullptr;
  float* d_w = bp.take<float>(wc.size());
  float* d_wp = bp.take<float>(wp.size());
  uint16_t* d_wp16 = wp16.empty() ? nullptr : bp.take<uint16_t>(wp16.size());
  int* d_flag = bp.take<int>(1);
  float* d_b = bp.take<float>(bc.size());
  float* d_ps = pre_scale ? bp.take<float>(Cin) : nullptr;
  float* d_pt = pre_scale ? bp.take<float>(Cin) : nullptr;
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_in, in, n_in * sizeof(float), cudaMemcpyHostToDevice, s));
  if (d_res) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_res, residual, n_out * sizeof(float), cudaMemcpyHostToDevice, s));
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_w, wc.data(), wc.size() * sizeof(float), cudaMemcpyHostToDevice, s));
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_wp, wp.data(), wp.size() * sizeof(float), cudaMemcpyHostToDevice, s));
  if (d_wp16) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_wp16, wp16.data(), wp16.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
  SUO_CUDA_TRY(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), s));
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_b, bc.data(), bc.size() * sizeof(float), cudaMemcpyHostToDevice, s));
  if (d_ps) {
    SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_ps, pre_scale, Cin * sizeof(float), cudaMemcpyHostToDevice, s));
    SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_pt, pre_shift, Cin * sizeof(float), cudaMemcpyHostToDevice, s));
  }
  SUO_CUDA_TRY(ctx, cudaMemsetAsync(d_out, 0, n_out * sizeof(float), s));
  p.in = d_in; p.w = d_w; p.w_packed = d_wp; p.bias = d_b; p.pre_scale = d_ps; p.pre_shift = d_pt; p.residual = d_res;
  p.out = d_out; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout; p.Cout_pad = Cout_pad;
  p.out_c = Cout; p.K = K; p.chunks_per_row = cpr; p.relu = relu; p.out_nchw = 0;
  // backend 2..5: tcgen05 FP16x3; 3/5 feed A by TMA from pre-split FP16 planes, 4/5 write the output as FP16 planes
  // backend 6 = 5 through the CTA-pair kernel (3x3, Cout = 128 only; other shapes run exactly as backend 5)
  const bool want_halo = backend == 7;       // backend 7 = 6 through the A-halo kernel (W in {16, 32, 64})
  const bool want_pair = backend == 6 || backend == 7;
  if (backend == 6 || backend == 7) backend = 5;
  const bool tma_in = backend == 3 || backend == 5, split_out = backend == 4 || backend == 5;
  p.math = backend >= 2 ? 1 : 0; p.w_packed16 = d_wp16; p.range_flag = d_flag;
  if (backend >= 2) backend = 1;
  uint16_t* d_in16 = nullptr;
  std::vector<uint16_t> h_out16;
  if (tma_in) {
    std::vector<uint16_t> h16(2 * n_in);
    conv_tc_host_split_f16(in, n_in, h16.data(), h16.data() + n_in);
    SUO_CUDA_TRY(ctx, cudaMalloc(&d_in16, 2 * n_in * sizeof(uint16_t)));
    SUO_CUDA_TRY(ctx, cudaMemcpy(d_in16, h16.data(), 2 * n_in * sizeof(uint16_t), cudaMemcpyHostToDevice));
    rc = make_plane_tmap(ctx, p.tmap_hi, d_in16, Cin, W, H, B);
    if (!rc) rc = make_plane_tmap(ctx, p.tmap_lo, d_in16 + n_in, Cin, W, H, B);
    if (rc) { cudaFree(d_in16); return rc; }
    p.in_split = 1;
  }
  if (split_out) { p.out_split = 1; p.out_plane = n_out; }   // d_out holds 2 * n_out halfs = n_out floats of storage
  p.mma_merge = ctx->opt_mma_merge;
  if (p.math == 1 && ctx->opt_epi_tma && Cout % 8 == 0) {
    const size_t rows = (size_t)B * Ho * Wo;
    int r2;
    if (split_out) {
      r2 = make_rows_tmap(ctx, p.tmap_out, d_out, true, rows, Cout);
      if (!r2) r2 = make_rows_tmap(ctx, p.tmap_out_lo, reinterpret_cast<uint16_t*>(d_out) + n_out, true, rows, Cout);
    } else {
      r2 = make_rows_tmap(ctx, p.tmap_out, d_out, false, rows, Cout);
    }
    if (!r2 && d_res) r2 = make_rows_tmap(ctx, p.tmap_res, d_res, false, rows, Cout);
    if (!r2 && ctx->opt_raw_tma && !tma_in && p.mode == CONV_1x1 && Cin % 64 == 0) {
      r2 = make_rows_tmap(ctx, p.tmap_raw, d_in, false, (size_t)B * H * W, Cin, 128);
      p.raw_tma = r2 ? 0 : 1;
    }
    if (r2) { if (d_in16) cudaFree(d_in16); return r2; }
    p.epi_tma = 1;
    if (want_pair && p.mode == CONV_3x3 && Cout_pad == 128 && d_wp16) {
      r2 = make_weight_tmap(ctx, p.tmap_w, d_wp16, wp16.size());
      if (r2) { if (d_in16) cudaFree(d_in16); return r2; }
      p.pair = 1;
      if (want_halo && d_in16 && !make_plane_tmap_halo(ctx, p.tmap_hhi, d_in16, Cin, W, H, B) &&
          !make_plane_tmap_halo(ctx, p.tmap_hlo, d_in16 + n_in, Cin, W, H, B))
        p.halo = 1;
    }
  }
  long long* d_dbg = nullptr;
  const char* dbg_path = getenv("SUO_CONV_TIMELINE");
  if (dbg_path && backend >= 1) {
    SUO_CUDA_TRY(ctx, cudaMalloc(&d_dbg, 16 * 512 * sizeof(long long)));
    SUO_CUDA_TRY(ctx, cudaMemsetAsync(d_dbg, 0, 16 * 512 * sizeof(long long), s));
    p.dbg = d_dbg;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, s);
  rc = backend == 1 ? launch_conv_tc(ctx, p, tf32_passes, s) : launch_conv_simt(ctx, p, s);
  cudaEventRecord(e1, s);
  if (rc) return rc;
  if (d_dbg) {
    std::vector<long long> h(16 * 512);
    SUO_CUDA_TRY(ctx, cudaMemcpyAsync(h.data(), d_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, s));
    SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (FILE* f = fopen(dbg_path, "a")) {
      fprintf(f, "# B=%d H=%d W=%d Cin=%d Cout=%d k=%d passes=%d kernel_ms=%.4f\n", B, H, W, Cin, Cout, ksize, tf32_passes, ms);
      fprintf(f, "g,prod_slot_free,prod_arrived,mma_full_a,mma_full_b,w_slot_free,pw0,pw1,pw2,pw3,pw4,pw5,pw6,pw7\n");
      long long t0 = 0;
      for (long long v : h) if (v && (!t0 || v < t0)) t0 = v;
      auto rel = [&](long long v) { return v ? v - t0 : -1; };
      for (int g = 0; g < 512 && (h[g] || h[1024 + g]); ++g) {
        fprintf(f, "%d,%lld,%lld,%lld,%lld,%lld", g, rel(h[g]), rel(h[512 + g]), rel(h[1024 + g]), rel(h[1536 + g]), rel(h[2048 + g]));
        for (int w = 0; w < 8; ++w) fprintf(f, ",%lld", rel(h[(5 + w) * 512 + g]));
        fprintf(f, "\n");
      }
      fprintf(f, "tile,epi_wait_start,epi_acc_full,epi_done\n");
      for (int t = 0; t < 512 && h[13 * 512 + t]; ++t)
        fprintf(f, "%d,%lld,%lld,%lld\n", t, rel(h[13 * 512 + t]), rel(h[14 * 512 + t]), rel(h[15 * 512 + t]));
      fclose(f);
    }
    cudaFree(d_dbg);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(out, d_out, n_out * sizeof(float), cudaMemcpyDeviceToHost, s));
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  if (split_out) {
    h_out16.resize(2 * n_out);
    memcpy(h_out16.data(), out, n_out * sizeof(float));
    conv_tc_host_join_f16(h_out16.data(), h_out16.data() + n_out, n_out, out);
  }
  if (d_in16) cudaFree(d_in16);
  return SUO_OK;
}

int suo_check_range(suo_ctx* ctx) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  CtxExtra* x = X(ctx);
  if (!x->loaded || !x->net.range_flag) return SUO_OK;
  int flag = 0;
  SUO_CUDA_TRY(ctx, cudaMemcpy(&flag, x->net.range_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag) {
    SUO_CUDA_TRY(ctx, cudaMemset(x->net.range_flag, 0, sizeof(int)));
    ctx->set_error("fp16x3 conv math: an activation exceeded the FP16 range (|x| > 6e4); results are invalid, use SUO_OPT_CONV_MATH = 0 (tf32x3)", __FILE__, __LINE__);
    return SUO_E_RANGE;
  }
  return SUO_OK;
}

// priors (dense planes) and prior_uv / prior_mask (keypoint priors rendered on the device) are mutually exclusive.
static int forward_impl(suo_ctx* ctx, const void* images, int n_img, int H, int W, const float* boxes, const int32_t* box_img,
                        int L, const float* priors, const float* prior_uv, const uint8_t* prior_mask, float* uv, float* cov,
                        float* logits, float* prob, float* mask_logits, float* mask, int32_t* argmax, int on_device, void* stream,
                        int images_u8 = 0) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  CtxExtra* x = X(ctx);
  if (!x->loaded) { ctx->set_error("suo_forward before suo_load_weights", __FILE__, __LINE__); return SUO_E_STATE; }
  if (L <= 0 || L > ctx->max_crops || n_img <= 0 || !images || !boxes || !box_img) {
    ctx->set_error("suo_forward: bad crop count / null input", __FILE__, __LINE__);
    return SUO_E_INVALID;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  NetState& N = x->net;
  const int R = ctx->crop_res, K = ctx->num_kp, HM = R / N.h.heat_div;
  const int variant = (priors || prior_uv) ? 1 : 0;
  const int in_buf = variant ? N.h.in_buf_prior : N.h.in_buf_noprior;
  const size_t n_im = (size_t)n_img * 3 * H * W, n_pr = priors ? (size_t)L * K * R * R : 0;
  const size_t n_hm = (size_t)L * K * HM * HM, LK = (size_t)L * K;
  const void* d_im = images;
  const float *d_box = boxes, *d_pr = priors, *d_puv = prior_uv;
  const uint8_t* d_pm = prior_mask;
  const int32_t* d_bi = box_img;
  float* d_prob = prob;
  if (!on_device) {
    for (int c = 0; c < L; ++c) if (box_img[c] < 0 || box_img[c] >= n_img) { ctx->set_error("suo_forward: box_img out of range", __FILE__, __LINE__); return SUO_E_INVALID; }
    rc = x->io.grow(ctx, (n_im + n_pr + 8 * (size_t)L + (prob ? n_hm : 0) + 3 * LK) * sizeof(float) + 8192);   // (u8 images need a quarter of n_im floats)
    if (rc) return rc;
    Bump bp{static_cast<uint8_t*>(x->io.d)};
    float* a = bp.take<float>(n_im);
    float* b = bp.take<float>(4 * (size_t)L);
    int32_t* c = bp.take<int32_t>(L);
    float* d = priors ? bp.take<float>(n_pr) : nullptr;
    d_prob = prob ? bp.take<float>(n_hm) : nullptr;
    if (prior_uv) {
      float* e = bp.take<float>(2 * LK); uint8_t* f = bp.take<uint8_t>(LK);
      SUO_CUDA_TRY(ctx, cudaMemcpyAsync(e, prior_uv, 2 * LK * sizeof(float), cudaMemcpyHostToDevice, s));
      SUO_CUDA_TRY(ctx, cudaMemcpyAsync(f, prior_mask, LK, cudaMemcpyHostToDevice, s));
      d_puv = e; d_pm = f;
    }
    SUO_CUDA_TRY(ctx, cudaMemcpyAsync(a, images, n_im * (images_u8 ? 1 : sizeof(float)), cudaMemcpyHostToDevice, s));
    SUO_CUDA_TRY(ctx, cudaMemcpyAsync(b, boxes, 4 * (size_t)L * sizeof(float), cudaMemcpyHostToDevice, s));
    SUO_CUDA_TRY(ctx, cudaMemcpyAsync(c, box_img, L * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    if (d) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d, priors, n_pr * sizeof(float), cudaMemcpyHostToDevice, s));
    d_im = a; d_box = b; d_bi = c; d_pr = d;
  }
  if (d_puv) {   // keypoint priors: RGB by roi_align, prior channels stamped straight into the NHWC input (prior.cu)
    rc = launch_crop_concat(ctx, d_im, n_img, H, W, d_box, d_bi, L, nullptr, -1, R, N.act[in_buf], N.bufs[in_buf].C, s, images_u8);
    if (rc) return rc;
    rc = launch_render_priors_nhwc(ctx, d_puv, d_pm, L, K, R, N.act[in_buf], N.bufs[in_buf].C, s);
  } else {
    const bool pad = N.stem_ok && N.bufs[in_buf].C == 4;
    rc = launch_crop_concat(ctx, d_im, n_img, H, W, d_box, d_bi, L, d_pr, K, R, N.act[in_buf], N.bufs[in_buf].C, s, images_u8,
                            pad ? N.stem_pad : nullptr, N.stem_wp, N.stem_hp);
  }
  if (rc) return rc;
  rc = run_network(ctx, L, variant, s);
  if (rc) return rc;
  const float* d_logits = N.act[N.h.logits_buf];
  if (on_device) {
    rc = launch_heatmap_reduce(ctx, d_logits, L, K, HM, HM, N.pool + N.h.cls_w_off, N.pool + N.h.cls_b_off, N.pooled, uv, cov,
                               d_prob, mask_logits, mask, argmax, s);
    if (rc) return rc;
    if (logits) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(logits, d_logits, n_hm * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return SUO_OK;
  }
  rc = launch_heatmap_reduce(ctx, d_logits, L, K, HM, HM, N.pool + N.h.cls_w_off, N.pool + N.h.cls_b_off, N.pooled, N.d_uv,
                             N.d_cov, d_prob, N.d_mask_logits, N.d_mask, N.d_argmax, s);
  if (rc) return rc;
  if (uv) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(uv, N.d_uv, LK * 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (cov) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(cov, N.d_cov, LK * 4 * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (mask) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(mask, N.d_mask, LK * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (mask_logits) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(mask_logits, N.d_mask_logits, LK * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (argmax) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(argmax, N.d_argmax, LK * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  if (logits) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(logits, d_logits, n_hm * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (prob) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(prob, d_prob, n_hm * sizeof(float), cudaMemcpyDeviceToHost, s));
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return suo_check_range(ctx);
}

int suo_forward(suo_ctx* ctx, const float* images, int n_img, int H, int W, const float* boxes, const int32_t* box_img,
                int L, const float* priors, float* uv, float* cov, float* logits, float* prob, float* mask_logits,
                float* mask, int32_t* argmax, int on_device, void* stream) {
  return forward_impl(ctx, images, n_img, H, W, boxes, box_img, L, priors, nullptr, nullptr, uv, cov, logits, prob, mask_logits,
                      mask, argmax, on_device, stream);
}

int suo_forward_kp_priors(suo_ctx* ctx, const float* images, int n_img, int H, int W, const float* boxes, const int32_t* box_img,
                          int L, const float* prior_uv, const uint8_t* prior_mask, float* uv, float* cov, float* logits,
                          float* prob, float* mask_logits, float* mask, int32_t* argmax, int on_device, void* stream) {
  if (!prior_uv || !prior_mask) { if (ctx) ctx->set_error("suo_forward_kp_priors: prior_uv / prior_mask are required", __FILE__, __LINE__); return SUO_E_INVALID; }
  return forward_impl(ctx, images, n_img, H, W, boxes, box_img, L, nullptr, prior_uv, prior_mask, uv, cov, logits, prob,
                      mask_logits, mask, argmax, on_device, stream);
}

int suo_render_priors(suo_ctx* ctx, const float* prior_uv, const uint8_t* prior_mask, int L, int K, int height, int width, int ndc,
                      float* out, int on_device, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if (L <= 0 || K <= 0 || height <= 0 || width <= 0 || !prior_uv || !prior_mask || !out) return SUO_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (on_device) return launch_render_priors_planes(ctx, prior_uv, prior_mask, L, K, height, width, ndc, out, s);
  CtxExtra* x = X(ctx);
  const size_t LK = (size_t)L * K, n_out = LK * height * width;
  rc = x->io.grow(ctx, (n_out + 2 * LK) * sizeof(float) + LK + 4096);
  if (rc) return rc;
  Bump bp{static_cast<uint8_t*>(x->io.d)};
  float* a = bp.take<float>(2 * LK); uint8_t* b = bp.take<uint8_t>(LK); float* c = bp.take<float>(n_out);
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(a, prior_uv, 2 * LK * sizeof(float), cudaMemcpyHostToDevice, s));
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(b, prior_mask, LK, cudaMemcpyHostToDevice, s));
  rc = launch_render_priors_planes(ctx, a, b, L, K, height, width, ndc, c, s);
  if (rc) return rc;
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(out, c, n_out * sizeof(float), cudaMemcpyDeviceToHost, s));
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return SUO_OK;
}

int suo_pnp_batch(suo_ctx* ctx, const double* xs, const double* ys, const int32_t* offsets, int n_obj, double threshold,
                  uint64_t seed, const uint64_t* obj_keys, double* T_out, int32_t* stats, int on_device, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if (n_obj <= 0 || !xs || !ys || !offsets || !T_out) return SUO_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (on_device) return launch_pnp_batch(ctx, xs, ys, offsets, n_obj, threshold, seed, obj_keys, T_out, stats, s, ctx->opt_pnp_max_pts);
  const int N = offsets[n_obj];
  int max_pts = 0;
  for (int o = 0; o < n_obj; ++o) {
    if (offsets[o + 1] < offsets[o]) { ctx->set_error("suo_pnp_batch: offsets must be non-decreasing", __FILE__, __LINE__); return SUO_E_INVALID; }
    max_pts = std::max(max_pts, offsets[o + 1] - offsets[o]);
  }
  CtxExtra* x = X(ctx);
  const size_t io_bytes = (size_t)N * 5 * 8 + (size_t)n_obj * (16 * 8 + 5 * 4 + 8 + 4) + 8192;
  rc = x->io.grow(ctx, io_bytes);
  if (!rc) rc = x->io.mirror(ctx, io_bytes);
  if (rc) return rc;
  Bump bp{static_cast<uint8_t*>(x->io.d)};
  double* d_xs = bp.take<double>(3 * (size_t)N);
  double* d_ys = bp.take<double>(2 * (size_t)N);
  int32_t* d_off = bp.take<int32_t>(n_obj + 1);
  uint64_t* d_keys = obj_keys ? bp.take<uint64_t>(n_obj) : nullptr;
  double* d_T = bp.take<double>(16 * (size_t)n_obj);
  int32_t* d_st = bp.take<int32_t>(5 * (size_t)n_obj);
  Stage st(x->io);
  st.in(d_xs, xs, 3 * (size_t)N); st.in(d_ys, ys, 2 * (size_t)N); st.in(d_off, offsets, (size_t)n_obj + 1);
  if (d_keys) st.in(d_keys, obj_keys, (size_t)n_obj);
  SUO_CUDA_TRY(ctx, st.send(s));
  rc = launch_pnp_batch(ctx, d_xs, d_ys, d_off, n_obj, threshold, seed, d_keys, d_T, d_st, s, max_pts);
  if (rc) return rc;
  st.want(d_T, 16 * (size_t)n_obj); st.want(d_st, 5 * (size_t)n_obj);
  SUO_CUDA_TRY(ctx, st.fetch(s));
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  st.out(T_out, d_T, 16 * (size_t)n_obj); st.out(stats, d_st, 5 * (size_t)n_obj);
  return SUO_OK;
}

// ---- coupled camera+object graphs: host-side grouping of the edges, then ba_global_kernel ----------------------
// True when some edge has a free camera AND a free object (global BA, lib/object_slam.py:736-778) or a problem has
// more vertices than the block-diagonal kernel keeps in shared memory.
static bool ba_needs_global(int n_prob, const int32_t* prob_vert, const int32_t* prob_edge, const uint8_t* fixed,
                            const int32_t* e_obj, const int32_t* e_cam) {
  for (int pr = 0; pr < n_prob; ++pr) {
    if (prob_vert[pr + 1] - prob_vert[pr] > 64) return true;
    for (int e = prob_edge[pr]; e < prob_edge[pr + 1]; ++e)
      if (e_obj[e] >= 0 && !fixed[e_obj[e]] && !fixed[e_cam[e]]) return true;
  }
  return false;
}

// True when every problem has one free vertex and at most one fixed one (BASELINE config 3, single-object frames, the curr_only camera solve):
// those run one warp per problem (ba_warp_kernel) instead of one CTA per problem.
static bool ba_single_vertex(int n_prob, const int32_t* prob_vert, const uint8_t* fixed) {
  for (int pr = 0; pr < n_prob; ++pr) {
    const int v0 = prob_vert[pr], nv = prob_vert[pr + 1] - v0;
    if (nv < 1 || nv > 2) return false;
    const int n_free = (fixed[v0] ? 0 : 1) + ((nv == 2 && !fixed[v0 + 1]) ? 1 : 0);
    if (n_free != 1) return false;
  }
  return true;
}

// Host arrays describe the graph structure (indices only); every d_* pointer is the device copy of the packed graph.
static int run_ba_global(suo_ctx* ctx, int n_prob, const int32_t* prob_vert, const int32_t* prob_edge, const uint8_t* fixed,
                         int n_vert, const int32_t* e_obj, const int32_t* e_cam, int n_edges, ba::BaArgs base, cudaStream_t s) {
  CtxExtra* x = X(ctx);
  std::vector<int32_t> perm(n_edges), pair_cam, pair_obj, pair_eoff, prob_pair(n_prob + 1, 0), obj_slot(n_vert, -1),
      slot_vert, prob_slot(n_prob, 0), prob_nobj(n_prob, 0), peer;
  std::vector<long long> prob_peer(n_prob, 0), prob_S(n_prob, 0);
  std::vector<uint8_t> is_obj(n_vert, 0), is_cam(n_vert, 0);
  long long S_total = 0;
  for (int pr = 0; pr < n_prob; ++pr) {
    const int v0 = prob_vert[pr], v1 = prob_vert[pr + 1], e0 = prob_edge[pr], e1 = prob_edge[pr + 1];
    for (int e = e0; e < e1; ++e) {
      if (e_cam[e] < v0 || e_cam[e] >= v1 || (e_obj[e] >= 0 && (e_obj[e] < v0 || e_obj[e] >= v1))) {
        ctx->set_error("suo_ba_batch: edge references a vertex outside its problem", __FILE__, __LINE__); return SUO_E_INVALID;
      }
      is_cam[e_cam[e]] = 1;
      if (e_obj[e] >= 0) is_obj[e_obj[e]] = 1;
      perm[e] = e;
    }
    std::stable_sort(perm.begin() + e0, perm.begin() + e1, [&](int32_t a, int32_t b) {
      return e_cam[a] != e_cam[b] ? e_cam[a] < e_cam[b] : e_obj[a] < e_obj[b];
    });
    prob_pair[pr] = (int32_t)pair_cam.size();
    for (int i = e0; i < e1; ++i) {
      const int e = perm[i];
      if (i == e0 || e_cam[e] != pair_cam.back() || e_obj[e] != pair_obj.back()) { pair_cam.push_back(e_cam[e]); pair_obj.push_back(e_obj[e]); pair_eoff.push_back(i); }
    }
    prob_slot[pr] = (int32_t)slot_vert.size();
    int no = 0;
    for (int v = v0; v < v1; ++v) {
      if (is_obj[v] && is_cam[v] && !fixed[v]) { ctx->set_error("suo_ba_batch: a free vertex is used both as object and as camera", __FILE__, __LINE__); return SUO_E_INVALID; }
      if (is_obj[v] && !fixed[v]) { obj_slot[v] = no++; slot_vert.push_back(v); }
    }
    prob_nobj[pr] = no;
    prob_S[pr] = S_total;
    S_total += (long long)(6 * no) * (6 * no) + 12 * no;
  }
  const int n_pairs = (int)pair_cam.size();
  prob_pair[n_prob] = n_pairs;
  pair_eoff.push_back(prob_edge[n_prob]);   // prob_edge is one offset array: the problems' edge ranges are contiguous
  std::vector<int32_t> cam_poff(n_vert + 1, 0), obj_poff(n_vert + 1, 0), obj_plist(n_pairs);
  for (int gp = 0; gp < n_pairs; ++gp) { cam_poff[pair_cam[gp] + 1]++; if (pair_obj[gp] >= 0) obj_poff[pair_obj[gp] + 1]++; }
  for (int v = 0; v < n_vert; ++v) { cam_poff[v + 1] += cam_poff[v]; obj_poff[v + 1] += obj_poff[v]; }
  {
    std::vector<int32_t> fill(obj_poff.begin(), obj_poff.end() - 1);
    for (int gp = 0; gp < n_pairs; ++gp) if (pair_obj[gp] >= 0) obj_plist[fill[pair_obj[gp]]++] = gp;
  }
  for (int pr = 0; pr < n_prob; ++pr) {
    const int no = prob_nobj[pr], pp0 = prob_pair[pr], pp1 = prob_pair[pr + 1];
    prob_peer[pr] = (long long)peer.size();
    peer.resize(peer.size() + (size_t)(pp1 - pp0) * no, -1);
    int32_t* P = peer.data() + prob_peer[pr];
    for (int a0 = pp0; a0 < pp1;) {
      int a1 = a0;
      while (a1 < pp1 && pair_cam[a1] == pair_cam[a0]) ++a1;
      if (!fixed[pair_cam[a0]])
        for (int q = a0; q < a1; ++q) {
          const int sq = pair_obj[q] >= 0 ? obj_slot[pair_obj[q]] : -1;
          if (sq < 0) continue;
          for (int pp = a0; pp < a1; ++pp) P[(size_t)(pp - pp0) * no + sq] = q;
        }
      a0 = a1;
    }
  }
  // one upload of the structure, one workspace
  struct Seg { const void* src; size_t bytes; size_t off; };
  std::vector<Seg> segs;
  size_t off = 0;
  auto add = [&](const void* src, size_t bytes) { off = align_up(off, 256); segs.push_back({src, bytes, off}); off += bytes; return segs.size() - 1; };
  const size_t i_perm = add(perm.data(), perm.size() * 4), i_pc = add(pair_cam.data(), pair_cam.size() * 4), i_po = add(pair_obj.data(), pair_obj.size() * 4),
               i_pe = add(pair_eoff.data(), pair_eoff.size() * 4), i_pp = add(prob_pair.data(), prob_pair.size() * 4), i_cp = add(cam_poff.data(), cam_poff.size() * 4),
               i_op = add(obj_poff.data(), obj_poff.size() * 4), i_ol = add(obj_plist.data(), obj_plist.size() * 4), i_os = add(obj_slot.data(), obj_slot.size() * 4),
               i_sv = add(slot_vert.data(), slot_vert.size() * 4), i_ps = add(prob_slot.data(), prob_slot.size() * 4), i_peer = add(peer.data(), peer.size() * 4),
               i_ppe = add(prob_peer.data(), prob_peer.size() * 8), i_pn = add(prob_nobj.data(), prob_nobj.size() * 4), i_pS = add(prob_S.data(), prob_S.size() * 8);
  const size_t struct_bytes = align_up(off, 256);
  const size_t work_bytes = ((size_t)S_total + (size_t)n_pairs * 156 + (size_t)n_vert * (36 + 6 + 6 + 36)) * 8 + 2 * (size_t)n_vert * sizeof(ba::SE3q) + n_vert + 8 * 256;
  // the previous launch may still be reading the workspace
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  int rc = x->bg.grow(ctx, struct_bytes + work_bytes);
  if (rc) return rc;
  uint8_t* d = static_cast<uint8_t*>(x->bg.d);
  std::vector<uint8_t> stage(struct_bytes, 0);
  for (const Seg& g : segs) if (g.bytes) memcpy(stage.data() + g.off, g.src, g.bytes);
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d, stage.data(), struct_bytes, cudaMemcpyHostToDevice, s));
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));   // `stage` goes out of scope
  ba::BgArgs a;
  a.g = base;
  auto I = [&](size_t i) { return reinterpret_cast<const int32_t*>(d + segs[i].off); };
  a.perm = I(i_perm); a.pair_cam = I(i_pc); a.pair_obj = I(i_po); a.pair_eoff = I(i_pe); a.prob_pair = I(i_pp); a.cam_poff = I(i_cp);
  a.obj_poff = I(i_op); a.obj_plist = I(i_ol); a.obj_slot = I(i_os); a.slot_vert = I(i_sv); a.prob_slot = I(i_ps); a.peer = I(i_peer);
  a.prob_peer = reinterpret_cast<const long long*>(d + segs[i_ppe].off); a.prob_nobj = I(i_pn);
  a.prob_S = reinterpret_cast<const long long*>(d + segs[i_pS].off);
  Bump wb{d + struct_bytes};
  a.S = wb.take<double>((size_t)S_total); a.pairw = wb.take<double>((size_t)n_pairs * 156);
  a.est = wb.take<ba::SE3q>(n_vert); a.bak = wb.take<ba::SE3q>(n_vert);
  a.Hv = wb.take<double>((size_t)n_vert * 36); a.bv = wb.take<double>((size_t)n_vert * 6); a.xv = wb.take<double>((size_t)n_vert * 6);
  a.Minv = wb.take<double>((size_t)n_vert * 36); a.vact = wb.take<uint8_t>(n_vert);
  return launch_ba_global(ctx, n_prob, a, s);
}

int suo_ba_batch(suo_ctx* ctx, int n_prob, const int32_t* prob_vert, const int32_t* prob_edge, double* poses,
                 const uint8_t* fixed, int n_vert, const int32_t* e_obj, const int32_t* e_cam, const double* cam_k,
                 const double* p, const double* uv, const double* info, uint8_t* inliers, int n_edges,
                 const int32_t* its, int n_rounds, double huber_delta, double chi2_gate, int init_with_outliers,
                 int32_t* stats, int on_device, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if (n_prob <= 0 || n_vert <= 0 || n_edges < 0 || n_rounds <= 0 || n_rounds > 16) return SUO_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CtxExtra* x = X(ctx);
  const size_t scratch = align_up((size_t)n_edges * 16, 256) + align_up(n_edges, 256) * 2 + 1024;
  rc = x->ba.grow(ctx, scratch);
  if (rc) return rc;
  Bump sb{static_cast<uint8_t*>(x->ba.d)};
  double* d_err = sb.take<double>(2 * (size_t)n_edges);
  x->last_err = d_err; x->last_err_n = n_edges;
  uint8_t* d_level = sb.take<uint8_t>(n_edges);
  int8_t* d_fv = sb.take<int8_t>(n_edges);
  auto base_args = [&](const int32_t* pv, const int32_t* pe, double* po, const uint8_t* fx, const int32_t* eo, const int32_t* ec,
                       const double* ck, const double* pp, const double* uu, const double* inf, uint8_t* inl, const int32_t* it, int32_t* st) {
    ba::BaArgs b;
    b.prob_vert = pv; b.prob_edge = pe; b.vert_cnt = nullptr; b.edge_cnt = nullptr; b.poses = po; b.fixed = fx; b.e_obj = eo; b.e_cam = ec;
    b.cam_k = ck; b.p = pp; b.uv = uu; b.info = inf; b.inliers = inl; b.its = it; b.n_rounds = n_rounds; b.huber_delta = huber_delta;
    b.chi2_gate = chi2_gate; b.init_with_outliers = init_with_outliers; b.stats = st; b.err = d_err; b.level = d_level; b.fv_kind = d_fv;
    return b;
  };
  if (on_device && ctx->opt_ba_blockdiag)    // the caller vouches for the structure: no host look at the index arrays, no synchronisation
    return launch_ba_batch_scratch(ctx, n_prob, prob_vert, prob_edge, poses, fixed, e_obj, e_cam, cam_k, p, uv, info, inliers,
                                   its, n_rounds, huber_delta, chi2_gate, init_with_outliers, stats, d_err, d_level, d_fv, s, nullptr, nullptr,
                                   ctx->opt_ba_blockdiag == 2);
  if (on_device) {
    // The graph structure decides the kernel: bring the index arrays to the host (this synchronises `stream`; the
    // device-resident frame path, suo_solve_keypoints / suo_frames, does not come through here).
    std::vector<int32_t> h_pv(n_prob + 1), h_pe(n_prob + 1), h_eo(n_edges), h_ec(n_edges);
    std::vector<uint8_t> h_fx(n_vert);
    SUO_CUDA_TRY(ctx, cudaMemcpyAsync(h_pv.data(), prob_vert, (n_prob + 1) * 4, cudaMemcpyDeviceToHost, s));
    SUO_CUDA_TRY(ctx, cudaMemcpyAsync(h_pe.data(), prob_edge, (n_prob + 1) * 4, cudaMemcpyDeviceToHost, s));
    SUO_CUDA_TRY(ctx, cudaMemcpyAsync(h_eo.data(), e_obj, (size_t)n_edges * 4, cudaMemcpyDeviceToHost, s));
    SUO_CUDA_TRY(ctx, cudaMemcpyAsync(h_ec.data(), e_cam, (size_t)n_edges * 4, cudaMemcpyDeviceToHost, s));
    SUO_CUDA_TRY(ctx, cudaMemcpyAsync(h_fx.data(), fixed, n_vert, cudaMemcpyDeviceToHost, s));
    SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    if (h_pv[n_prob] > n_vert || h_pe[n_prob] > n_edges) { ctx->set_error("suo_ba_batch: offsets exceed n_vert / n_edges", __FILE__, __LINE__); return SUO_E_INVALID; }
    if (ba_needs_global(n_prob, h_pv.data(), h_pe.data(), h_fx.data(), h_eo.data(), h_ec.data()))
      return run_ba_global(ctx, n_prob, h_pv.data(), h_pe.data(), h_fx.data(), n_vert, h_eo.data(), h_ec.data(), n_edges,
                           base_args(prob_vert, prob_edge, poses, fixed, e_obj, e_cam, cam_k, p, uv, info, inliers, its, stats), s);
    return launch_ba_batch_scratch(ctx, n_prob, prob_vert, prob_edge, poses, fixed, e_obj, e_cam, cam_k, p, uv, info, inliers,
                                   its, n_rounds, huber_delta, chi2_gate, init_with_outliers, stats, d_err, d_level, d_fv, s, nullptr, nullptr,
                                   ba_single_vertex(n_prob, h_pv.data(), h_fx.data()));
  }
  // host-side validation (the kernels report the same conditions through stats)
  if (prob_vert[n_prob] > n_vert || prob_edge[n_prob] > n_edges) { ctx->set_error("suo_ba_batch: offsets exceed n_vert / n_edges", __FILE__, __LINE__); return SUO_E_INVALID; }
  for (int pr = 0; pr < n_prob; ++pr)
    for (int e = prob_edge[pr]; e < prob_edge[pr + 1]; ++e)
      if (e_cam[e] < prob_vert[pr] || e_cam[e] >= prob_vert[pr + 1] || (e_obj[e] >= 0 && (e_obj[e] < prob_vert[pr] || e_obj[e] >= prob_vert[pr + 1]))) {
        ctx->set_error("suo_ba_batch: edge references a vertex outside its problem", __FILE__, __LINE__); return SUO_E_INVALID;
      }
  const bool global = ba_needs_global(n_prob, prob_vert, prob_edge, fixed, e_obj, e_cam);
  const size_t io_bytes = (size_t)n_vert * (12 * 8 + 1) + (size_t)n_edges * (4 + 4 + (4 + 3 + 2 + 4) * 8 + 1) + (size_t)n_prob * (8 + 12) + 16 * 4 + 16384;
  rc = x->io.grow(ctx, io_bytes);
  if (!rc) rc = x->io.mirror(ctx, io_bytes);
  if (rc) return rc;
  Bump bp{static_cast<uint8_t*>(x->io.d)};
  int32_t* d_pv = bp.take<int32_t>(n_prob + 1);
  int32_t* d_pe = bp.take<int32_t>(n_prob + 1);
  uint8_t* d_fixed = bp.take<uint8_t>(n_vert);
  int32_t* d_eo = bp.take<int32_t>(n_edges);
  int32_t* d_ec = bp.take<int32_t>(n_edges);
  double* d_k = bp.take<double>(4 * (size_t)n_edges);
  double* d_p = bp.take<double>(3 * (size_t)n_edges);
  double* d_uv = bp.take<double>(2 * (size_t)n_edges);
  double* d_info = bp.take<double>(4 * (size_t)n_edges);
  int32_t* d_its = bp.take<int32_t>(n_rounds);
  double* d_poses = bp.take<double>(12 * (size_t)n_vert);        // the in/out arrays and the statistics last: they come back as one block
  uint8_t* d_inl = bp.take<uint8_t>(n_edges);
  int32_t* d_st = bp.take<int32_t>(3 * (size_t)n_prob);
  Stage st(x->io);
  st.in(d_pv, prob_vert, (size_t)n_prob + 1); st.in(d_pe, prob_edge, (size_t)n_prob + 1);
  st.in(d_fixed, fixed, (size_t)n_vert); st.in(d_eo, e_obj, (size_t)n_edges); st.in(d_ec, e_cam, (size_t)n_edges);
  st.in(d_k, cam_k, 4 * (size_t)n_edges); st.in(d_p, p, 3 * (size_t)n_edges); st.in(d_uv, uv, 2 * (size_t)n_edges);
  st.in(d_info, info, 4 * (size_t)n_edges); st.in(d_its, its, (size_t)n_rounds);
  st.in(d_poses, static_cast<const double*>(poses), 12 * (size_t)n_vert); st.in(d_inl, static_cast<const uint8_t*>(inliers), (size_t)n_edges);
  SUO_CUDA_TRY(ctx, st.send(s));
  if (global)   // coupled camera+object graph (global BA) or a graph too large for the shared-memory kernel
    rc = run_ba_global(ctx, n_prob, prob_vert, prob_edge, fixed, n_vert, e_obj, e_cam, n_edges,
                       base_args(d_pv, d_pe, d_poses, d_fixed, d_eo, d_ec, d_k, d_p, d_uv, d_info, d_inl, d_its, d_st), s);
  else
    rc = launch_ba_batch_scratch(ctx, n_prob, d_pv, d_pe, d_poses, d_fixed, d_eo, d_ec, d_k, d_p, d_uv, d_info, d_inl, d_its,
                                 n_rounds, huber_delta, chi2_gate, init_with_outliers, d_st, d_err, d_level, d_fv, s, nullptr, nullptr,
                                 ba_single_vertex(n_prob, prob_vert, fixed));
  if (rc) return rc;
  st.want(d_poses, 12 * (size_t)n_vert); st.want(d_inl, (size_t)n_edges); st.want(d_st, 3 * (size_t)n_prob);
  SUO_CUDA_TRY(ctx, st.fetch(s));
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  st.out(poses, d_poses, 12 * (size_t)n_vert); st.out(inliers, d_inl, (size_t)n_edges); st.out(stats, d_st, 3 * (size_t)n_prob);
  return SUO_OK;
}

// offsets c * K and keys c of the per-crop keypoint lists: they never change, so they are built once per context and the device path stays
// free of host synchronisation
static int ensure_pnp_tables(suo_ctx* ctx, int L, cudaStream_t s) {
  CtxExtra* xo = X(ctx);
  if (xo->pnp_off_n >= L) return SUO_OK;
  const int n = std::max(L, ctx->max_crops), K = ctx->num_kp;
  std::vector<int32_t> off(n);
  std::vector<uint64_t> keys(n);
  for (int c = 0; c < n; ++c) { off[c] = c * K; keys[c] = (uint64_t)c; }
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));        // (re)allocation only: earlier work may still read the old arrays
  if (xo->pnp_off) cudaFree(xo->pnp_off);
  if (xo->pnp_keys) cudaFree(xo->pnp_keys);
  xo->pnp_off = nullptr; xo->pnp_keys = nullptr; xo->pnp_off_n = 0;
  SUO_CUDA_TRY(ctx, cudaMalloc(&xo->pnp_off, off.size() * sizeof(int32_t)));
  SUO_CUDA_TRY(ctx, cudaMalloc(&xo->pnp_keys, keys.size() * sizeof(uint64_t)));
  SUO_CUDA_TRY(ctx, cudaMemcpy(xo->pnp_off, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  SUO_CUDA_TRY(ctx, cudaMemcpy(xo->pnp_keys, keys.data(), keys.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
  xo->pnp_off_n = n;
  return SUO_OK;
}

// Device-resident part after the network: gating -> PnP -> single-view BA.  All pointers are device
// pointers; d_uv/d_cov/d_mask are the network outputs (or caller-provided keypoints).
static int solve_keypoints_device(suo_ctx* ctx, Bump& bp, const float* d_uv, const float* d_cov, const float* d_mask,
                                  const int32_t* d_bi, int n_img, int L, const double* d_mk, const uint8_t* d_mm,
                                  const double* d_kb, const double* d_diam, double kp_var_thresh, double bbox_thresh,
                                  uint64_t seed, int run_ba, double* d_Tpnp, double* d_Tba, uint8_t* d_used,
                                  uint8_t* d_bain, cudaStream_t s) {
  const int K = ctx->num_kp;
  const size_t LK = (size_t)L * K, NV = (size_t)L + n_img;
  double* d_xs = bp.take<double>(3 * LK); double* d_ys = bp.take<double>(2 * LK);
  int32_t* d_cnt = bp.take<int32_t>(L); int32_t* d_kpi = bp.take<int32_t>(LK);
  int32_t* d_pst = bp.take<int32_t>(5 * (size_t)L);
  int32_t* d_fs = bp.take<int32_t>(n_img + 1);
  int rc = launch_gate_compact(ctx, d_uv, d_cov, d_mask, d_mm, d_mk, d_kb, L, K, (float)kp_var_thresh, (float)bbox_thresh, d_xs,
                               d_ys, d_cnt, d_kpi, d_used, s);
  if (rc) return rc;
  // PnP: object c owns rows [c*K, c*K + count[c])
  rc = ensure_pnp_tables(ctx, L, s);
  if (rc) return rc;
  CtxExtra* xo = X(ctx);
  const int32_t* d_off = xo->pnp_off;
  rc = launch_pnp_batch_counts(ctx, d_xs, d_ys, d_off, d_cnt, L, 0.001, seed, nullptr, d_Tpnp, d_pst, s, K);
  if (rc) return rc;
  if (!run_ba) return SUO_OK;
  rc = launch_frame_ranges(ctx, d_bi, L, n_img, d_fs, s);
  if (rc) return rc;
  double* d_poses = bp.take<double>(12 * NV); uint8_t* d_fixed = bp.take<uint8_t>(NV);
  int32_t* d_pv = bp.take<int32_t>(n_img); int32_t* d_vc = bp.take<int32_t>(n_img);
  int32_t* d_pe = bp.take<int32_t>(n_img); int32_t* d_ecnt = bp.take<int32_t>(n_img);
  int32_t* d_eo = bp.take<int32_t>(LK); int32_t* d_ec = bp.take<int32_t>(LK);
  double* d_ck = bp.take<double>(4 * LK); double* d_p = bp.take<double>(3 * LK); double* d_uvd = bp.take<double>(2 * LK);
  double* d_info = bp.take<double>(4 * LK); uint8_t* d_inl = bp.take<uint8_t>(LK); int32_t* d_esrc = bp.take<int32_t>(LK);
  uint8_t* d_acc = bp.take<uint8_t>(L);
  double* d_err = bp.take<double>(2 * LK); uint8_t* d_lvl = bp.take<uint8_t>(LK); int8_t* d_fv = bp.take<int8_t>(LK);
  int32_t* d_its = bp.take<int32_t>(4); int32_t* d_bst = bp.take<int32_t>(3 * (size_t)n_img);
  SUO_CUDA_TRY(ctx, cudaMemsetAsync(d_fixed, 0, NV, s));
  rc = launch_ba_assemble(ctx, n_img, d_fs, d_cnt, d_kpi, d_xs, d_uv, d_cov, d_kb, d_diam, d_Tpnp, K, d_poses, d_fixed, d_pv,
                          d_vc, d_pe, d_ecnt, d_eo, d_ec, d_ck, d_p, d_uvd, d_info, d_inl, d_esrc, d_acc, s);
  if (rc) return rc;
  static const int32_t its_host[4] = {10, 10, 10, 10};   // single-view mode (object_slam.py:843-846)
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_its, its_host, sizeof(its_host), cudaMemcpyHostToDevice, s));
  rc = launch_ba_batch_scratch(ctx, n_img, d_pv, d_pe, d_poses, d_fixed, d_eo, d_ec, d_ck, d_p, d_uvd, d_info, d_inl, d_its, 4,
                               2.4476519768340177 /* sqrt(5.991) */, 5.991, 0, d_bst, d_err, d_lvl, d_fv, s, d_vc, d_ecnt);
  if (rc) return rc;
  return launch_ba_scatter(ctx, n_img, d_fs, d_ecnt, d_esrc, d_inl, d_poses, d_acc, K, d_Tba, d_bain, s);
}

static size_t solve_workspace_bytes(int L, int K, int n_img) {
  const size_t LK = (size_t)L * K, NV = (size_t)L + n_img;
  return 64 * 256 + LK * (3 + 2 + 4 + 3 + 2 + 4 + 2 + 3) * 8 + LK * (4 + 4 + 4 + 4 + 8) + (size_t)L * (16 * 8 + 5 * 4 + 4 * 3 + 1 + 9 * 8 + 8 + 12 * 8) +
         NV * (12 * 8 + 1) + (size_t)n_img * (4 * 6 + 12) + 4096;
}

int suo_solve_keypoints(suo_ctx* ctx, const float* uv, const float* cov, const float* kp_mask, const int32_t* box_img,
                        int n_img, int L, const double* model_kps, const uint8_t* model_mask, const double* K_bbox,
                        const double* diameter, double kp_var_thresh, double bbox_thresh, uint64_t seed, int run_ba,
                        double* T_pnp, double* T_ba, uint8_t* kp_used, uint8_t* ba_inliers, int on_device, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if (L <= 0 || n_img <= 0 || !uv || !cov || !kp_mask || !box_img || !model_kps || !model_mask || !K_bbox || !diameter) return SUO_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CtxExtra* x = X(ctx);
  const int K = ctx->num_kp;
  const size_t LK = (size_t)L * K;
  rc = x->fr.grow(ctx, solve_workspace_bytes(L, K, n_img) + LK * 64);
  if (rc) return rc;
  Bump bp{static_cast<uint8_t*>(x->fr.d)};
  const float *d_uv = uv, *d_cov = cov, *d_km = kp_mask;
  const int32_t* d_bi = box_img;
  const double *d_mk = model_kps, *d_kb = K_bbox, *d_diam = diameter;
  const uint8_t* d_mm = model_mask;
  if (!on_device) {
    for (int c = 1; c < L; ++c) if (box_img[c] < box_img[c - 1]) { ctx->set_error("box_img must be sorted", __FILE__, __LINE__); return SUO_E_INVALID; }
    float* a = bp.take<float>(2 * LK); float* b = bp.take<float>(4 * LK); float* c = bp.take<float>(LK); int32_t* d = bp.take<int32_t>(L);
    double* e = bp.take<double>(3 * LK); uint8_t* f = bp.take<uint8_t>(LK); double* g = bp.take<double>(9 * (size_t)L); double* h = bp.take<double>(L);
#define H2D(dst, src, n) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, (n), cudaMemcpyHostToDevice, s))
    H2D(a, uv, 8 * LK); H2D(b, cov, 16 * LK); H2D(c, kp_mask, 4 * LK); H2D(d, box_img, 4 * (size_t)L);
    H2D(e, model_kps, 24 * LK); H2D(f, model_mask, LK); H2D(g, K_bbox, 72 * (size_t)L); H2D(h, diameter, 8 * (size_t)L);
#undef H2D
    d_uv = a; d_cov = b; d_km = c; d_bi = d; d_mk = e; d_mm = f; d_kb = g; d_diam = h;
  }
  double* d_Tpnp = (on_device && T_pnp) ? T_pnp : bp.take<double>(16 * (size_t)L);
  double* d_Tba = (on_device && T_ba) ? T_ba : bp.take<double>(12 * (size_t)L);
  uint8_t* d_used = (on_device && kp_used) ? kp_used : bp.take<uint8_t>(LK);
  uint8_t* d_bain = (on_device && ba_inliers) ? ba_inliers : bp.take<uint8_t>(LK);
  rc = solve_keypoints_device(ctx, bp, d_uv, d_cov, d_km, d_bi, n_img, L, d_mk, d_mm, d_kb, d_diam, kp_var_thresh, bbox_thresh, seed,
                              run_ba, d_Tpnp, d_Tba, d_used, d_bain, s);
  if (rc || on_device) return rc;
#define D2H(dst, src, n) if (dst) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, (n), cudaMemcpyDeviceToHost, s))
  D2H(T_pnp, d_Tpnp, 128 * (size_t)L); D2H(kp_used, d_used, LK);
  if (run_ba) { D2H(T_ba, d_Tba, 96 * (size_t)L); D2H(ba_inliers, d_bain, LK); }
#undef D2H
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return SUO_OK;
}

int suo_chi2_inlier_counts(suo_ctx* ctx, int n_pairs, const double* T_pairs, const int32_t* pair_det, int n_det,
                           const int32_t* det_off, const double* model_kp, const double* K, const float* uv, const float* cov,
                           const uint8_t* use, double manual_kp_std, double chi2_gate, int32_t* counts, int on_device, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if (n_pairs <= 0 || n_det <= 0 || !T_pairs || !pair_det || !det_off || !model_kp || !K || !uv || !counts || !(manual_kp_std > 0)) return SUO_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (on_device) return launch_chi2_counts(ctx, n_pairs, T_pairs, pair_det, det_off, model_kp, K, uv, cov, use, manual_kp_std, chi2_gate, counts, s);
  const size_t N = (size_t)det_off[n_det];
  for (int p = 0; p < n_pairs; ++p)
    if (pair_det[p] < 0 || pair_det[p] >= n_det) { ctx->set_error("suo_chi2_inlier_counts: pair_det out of range", __FILE__, __LINE__); return SUO_E_INVALID; }
  CtxExtra* x = X(ctx);
  rc = x->io.grow(ctx, (size_t)n_pairs * (12 * 8 + 4 + 4) + (size_t)(n_det + 1) * 4 + (size_t)n_det * 72 + N * (24 + 8 + 16 + 1) + 16 * 256);
  if (rc) return rc;
  Bump bp{static_cast<uint8_t*>(x->io.d)};
  double* d_T = bp.take<double>(12 * (size_t)n_pairs); int32_t* d_pd = bp.take<int32_t>(n_pairs); int32_t* d_off = bp.take<int32_t>(n_det + 1);
  double* d_mk = bp.take<double>(3 * N); double* d_K = bp.take<double>(9 * (size_t)n_det); float* d_uv = bp.take<float>(2 * N);
  float* d_cov = cov ? bp.take<float>(4 * N) : nullptr; uint8_t* d_use = use ? bp.take<uint8_t>(N) : nullptr;
  int32_t* d_cnt = bp.take<int32_t>(n_pairs);
#define H2D(dst, src, n) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, (n), cudaMemcpyHostToDevice, s))
  H2D(d_T, T_pairs, 96 * (size_t)n_pairs); H2D(d_pd, pair_det, 4 * (size_t)n_pairs); H2D(d_off, det_off, 4 * (size_t)(n_det + 1));
  H2D(d_mk, model_kp, 24 * N); H2D(d_K, K, 72 * (size_t)n_det); H2D(d_uv, uv, 8 * N);
  if (cov) H2D(d_cov, cov, 16 * N);
  if (use) H2D(d_use, use, N);
#undef H2D
  rc = launch_chi2_counts(ctx, n_pairs, d_T, d_pd, d_off, d_mk, d_K, d_uv, d_cov, d_use, manual_kp_std, chi2_gate, d_cnt, s);
  if (rc) return rc;
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(counts, d_cnt, 4 * (size_t)n_pairs, cudaMemcpyDeviceToHost, s));
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return SUO_OK;
}

static int frames_impl(suo_ctx* ctx, const void* images, int images_u8, int n_img, int H, int W, const float* boxes, const int32_t* box_img,
                       int L, const float* priors, const double* model_kps, const uint8_t* model_mask, const double* K_bbox,
                       const double* diameter, double kp_var_thresh, double bbox_thresh, uint64_t seed, int run_ba,
                       double* T_pnp, double* T_ba, uint8_t* kp_used, uint8_t* ba_inliers, float* uv, float* cov,
                       int on_device, void* stream, int slot = -1, void* records = nullptr, int record_id_base = 0) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  CtxExtra* x = X(ctx);
  if (!x->loaded) { ctx->set_error("suo_frames before suo_load_weights", __FILE__, __LINE__); return SUO_E_STATE; }
  if (slot >= 0 && (slot > 1 || on_device)) { ctx->set_error("suo_frames_u8_submit: slot must be 0 or 1", __FILE__, __LINE__); return SUO_E_INVALID; }
  if (L <= 0 || L > ctx->max_crops || n_img <= 0 || !images || !boxes || !box_img || !model_kps || !model_mask || !K_bbox || !diameter) {
    ctx->set_error("suo_frames: bad crop count / null input", __FILE__, __LINE__);
    return SUO_E_INVALID;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  NetState& N = x->net;
  const int K = ctx->num_kp, R = ctx->crop_res;
  const size_t LK = (size_t)L * K, n_im = (size_t)n_img * 3 * H * W, n_pr = priors ? LK * R * R : 0;
  size_t bytes = solve_workspace_bytes(L, K, n_img) + LK * 64;
  const size_t in_bytes = (n_im + n_pr) * sizeof(float) + LK * 25 + (size_t)L * 100 + 16 * 256;
  if (!on_device && slot < 0) bytes += in_bytes;
  rc = x->fr.grow(ctx, bytes);
  if (rc) return rc;
  Bump bp{static_cast<uint8_t*>(x->fr.d)};
  // asynchronous submit: inputs are staged in the slot's own block by the copy stream; `s` waits for them
  CtxExtra::FrameSlot* sl = slot >= 0 ? &x->slot[slot] : nullptr;
  Bump ip = bp;
  cudaStream_t cs = s;
  if (sl) {
    if (sl->pending) { ctx->set_error("suo_frames_u8_submit: slot still pending (call suo_frames_wait first)", __FILE__, __LINE__); return SUO_E_STATE; }
    // both batches run through the ONE executor (activations, keypoint buffers, solver workspace): only stream order keeps them apart
    if (x->slot[slot ^ 1].pending && x->slot[slot ^ 1].stream != s) {
      ctx->set_error("suo_frames_u8_submit: both slots must be submitted on the same stream", __FILE__, __LINE__);
      return SUO_E_INVALID;
    }
    if (!x->copy_stream) SUO_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&x->copy_stream, cudaStreamNonBlocking));
    if (!x->aux_stream) SUO_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&x->aux_stream, cudaStreamNonBlocking));
    if (!sl->flag) SUO_CUDA_TRY(ctx, cudaMalloc(&sl->flag, sizeof(int)));
    if (!sl->h2d_done) {
      SUO_CUDA_TRY(ctx, cudaEventCreateWithFlags(&sl->h2d_done, cudaEventDisableTiming));
      SUO_CUDA_TRY(ctx, cudaEventCreateWithFlags(&sl->done, cudaEventDisableTiming));
    }
    rc = sl->in.grow(ctx, in_bytes);
    if (!rc) rc = sl->out.grow(ctx, (size_t)L * (16 + 12) * 8 + LK * (2 + 24) + 8 * 256);
    if (rc) return rc;
    ip = Bump{static_cast<uint8_t*>(sl->in.d)};
    cs = x->copy_stream;
    // the slot's previous batch (its compute read this staging block) has finished: suo_frames_wait synchronised on `done`
  }
  const void* d_im = images;
  const float *d_box = boxes, *d_pr = priors;
  const int32_t* d_bi = box_img;
  const double *d_mk = model_kps, *d_kb = K_bbox, *d_diam = diameter;
  const uint8_t* d_mm = model_mask;
  if (!on_device) {
    for (int c = 1; c < L; ++c) if (box_img[c] < box_img[c - 1]) { ctx->set_error("suo_frames: box_img must be sorted", __FILE__, __LINE__); return SUO_E_INVALID; }
    if (box_img[0] < 0 || box_img[L - 1] >= n_img) { ctx->set_error("suo_frames: box_img out of range", __FILE__, __LINE__); return SUO_E_INVALID; }
    Bump& q = sl ? ip : bp;
    float* a = q.take<float>(n_im); float* b = q.take<float>(4 * (size_t)L); int32_t* c = q.take<int32_t>(L);
    float* d = priors ? q.take<float>(n_pr) : nullptr;
    double* e = q.take<double>(3 * LK); uint8_t* f = q.take<uint8_t>(LK); double* g = q.take<double>(9 * (size_t)L); double* h = q.take<double>(L);
#define H2D(dst, src, n) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, (n), cudaMemcpyHostToDevice, cs))
    H2D(a, images, n_im * (images_u8 ? 1 : 4)); H2D(b, boxes, 16 * (size_t)L); H2D(c, box_img, 4 * (size_t)L);
    if (d) H2D(d, priors, n_pr * 4);
    H2D(e, model_kps, 24 * LK); H2D(f, model_mask, LK); H2D(g, K_bbox, 72 * (size_t)L); H2D(h, diameter, 8 * (size_t)L);
#undef H2D
    d_im = a; d_box = b; d_bi = c; d_pr = d; d_mk = e; d_mm = f; d_kb = g; d_diam = h;
    if (sl) {
      SUO_CUDA_TRY(ctx, cudaEventRecord(sl->h2d_done, cs));
      SUO_CUDA_TRY(ctx, cudaStreamWaitEvent(s, sl->h2d_done, 0));
    }
  }
  // forward (device path): results land in the executor's own uv / cov / mask buffers
  rc = forward_impl(ctx, d_im, n_img, H, W, d_box, d_bi, L, d_pr, nullptr, nullptr, N.d_uv, N.d_cov, nullptr, nullptr, N.d_mask_logits,
                    N.d_mask, N.d_argmax, 1, stream, images_u8);
  if (rc) return rc;
  Bump op = sl ? Bump{static_cast<uint8_t*>(sl->out.d)} : bp;      // (a slot keeps its results in its own block until suo_frames_wait)
  Bump& ob = sl ? op : bp;
  double* d_Tpnp = (on_device && T_pnp) ? T_pnp : ob.take<double>(16 * (size_t)L);
  double* d_Tba = (on_device && T_ba) ? T_ba : ob.take<double>(12 * (size_t)L);
  uint8_t* d_used = (on_device && kp_used) ? kp_used : ob.take<uint8_t>(LK);
  uint8_t* d_bain = (on_device && ba_inliers) ? ba_inliers : ob.take<uint8_t>(LK);
  float* d_uvo = sl ? ob.take<float>(2 * LK) : N.d_uv;
  float* d_covo = sl ? ob.take<float>(4 * LK) : N.d_cov;
  rc = solve_keypoints_device(ctx, bp, N.d_uv, N.d_cov, N.d_mask, d_bi, n_img, L, d_mk, d_mm, d_kb, d_diam, kp_var_thresh,
                              bbox_thresh, seed, run_ba, d_Tpnp, d_Tba, d_used, d_bain, s);
  if (rc) return rc;
  if (on_device) {
    if (uv) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(uv, N.d_uv, LK * 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (cov) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(cov, N.d_cov, LK * 4 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return SUO_OK;
  }
  if (records) {   // result records for the multi-GPU exchange, packed on the device from this batch's own outputs
    rc = launch_pack_records(ctx, nullptr, record_id_base, d_Tpnp, run_ba ? d_Tba : nullptr, d_used, run_ba ? d_bain : nullptr, N.d_uv, N.d_cov,
                             L, K, (int)suo_record_bytes(K), static_cast<uint8_t*>(records), s);
    if (rc) return rc;
  }
  if (sl) {     // the next batch's forward overwrites the executor's uv / cov: keep this batch's copy in the slot
    if (uv) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_uvo, N.d_uv, LK * 8, cudaMemcpyDeviceToDevice, s));
    if (cov) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(d_covo, N.d_cov, LK * 16, cudaMemcpyDeviceToDevice, s));
  }
#define D2H(dst, src, n) if (dst) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, (n), cudaMemcpyDeviceToHost, s))
  D2H(T_pnp, d_Tpnp, 128 * (size_t)L); D2H(kp_used, d_used, LK);
  D2H(uv, d_uvo, LK * 8); D2H(cov, d_covo, LK * 16);
  if (run_ba) { D2H(T_ba, d_Tba, 96 * (size_t)L); D2H(ba_inliers, d_bain, LK); }
#undef D2H
  if (sl) {
    // this batch's range flag, in stream order: later batches raise (and clear) the executor's flag again, so suo_frames_wait reads the snapshot
    SUO_CUDA_TRY(ctx, cudaMemcpyAsync(sl->flag, N.range_flag, sizeof(int), cudaMemcpyDeviceToDevice, s));
    SUO_CUDA_TRY(ctx, cudaMemsetAsync(N.range_flag, 0, sizeof(int), s));
    SUO_CUDA_TRY(ctx, cudaEventRecord(sl->done, s));
    sl->pending = true;
    sl->stream = s;
    return SUO_OK;
  }
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return suo_check_range(ctx);      // fp16x3: an activation outside the FP16 range invalidates the keypoints and poses
}

int suo_frames(suo_ctx* ctx, const float* images, int n_img, int H, int W, const float* boxes, const int32_t* box_img,
               int L, const float* priors, const double* model_kps, const uint8_t* model_mask, const double* K_bbox,
               const double* diameter, double kp_var_thresh, double bbox_thresh, uint64_t seed, int run_ba,
               double* T_pnp, double* T_ba, uint8_t* kp_used, uint8_t* ba_inliers, float* uv, float* cov,
               int on_device, void* stream) {
  return frames_impl(ctx, images, 0, n_img, H, W, boxes, box_img, L, priors, model_kps, model_mask, K_bbox, diameter, kp_var_thresh,
                     bbox_thresh, seed, run_ba, T_pnp, T_ba, kp_used, ba_inliers, uv, cov, on_device, stream);
}

int suo_frames_u8(suo_ctx* ctx, const uint8_t* images_hwc, int n_img, int H, int W, const float* boxes, const int32_t* box_img,
                  int L, const float* priors, const double* model_kps, const uint8_t* model_mask, const double* K_bbox,
                  const double* diameter, double kp_var_thresh, double bbox_thresh, uint64_t seed, int run_ba,
                  double* T_pnp, double* T_ba, uint8_t* kp_used, uint8_t* ba_inliers, float* uv, float* cov,
                  int on_device, void* stream) {
  return frames_impl(ctx, images_hwc, 1, n_img, H, W, boxes, box_img, L, priors, model_kps, model_mask, K_bbox, diameter, kp_var_thresh,
                     bbox_thresh, seed, run_ba, T_pnp, T_ba, kp_used, ba_inliers, uv, cov, on_device, stream);
}

int suo_ba_last_errors(suo_ctx* ctx, double* err, int n_edges, int on_device, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  CtxExtra* x = X(ctx);
  if (!err || n_edges <= 0 || n_edges != x->last_err_n || !x->last_err) { ctx->set_error("suo_ba_last_errors: no suo_ba_batch result with this edge count", __FILE__, __LINE__); return SUO_E_STATE; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(err, x->last_err, 16 * (size_t)n_edges, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  if (!on_device) SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return SUO_OK;
}

int suo_edge_linearize(suo_ctx* ctx, int n_edges, const double* T_obj, const double* T_cam, const double* cam_k, const double* p,
                       const double* uv, double* err, double* J_obj, double* J_cam, int on_device, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if (n_edges <= 0 || !T_cam || !cam_k || !p || !uv) return SUO_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (on_device) return launch_edge_linearize(ctx, n_edges, T_obj, T_cam, cam_k, p, uv, err, J_obj, J_cam, s);
  CtxExtra* x = X(ctx);
  const size_t n = (size_t)n_edges;
  rc = x->io.grow(ctx, n * 8 * (12 + 12 + 4 + 3 + 2 + 2 + 12 + 12) + 16 * 256);
  if (rc) return rc;
  Bump bp{static_cast<uint8_t*>(x->io.d)};
  double* d_to = T_obj ? bp.take<double>(12 * n) : nullptr; double* d_tc = bp.take<double>(12 * n);
  double* d_k = bp.take<double>(4 * n); double* d_p = bp.take<double>(3 * n); double* d_uv = bp.take<double>(2 * n);
  double* d_e = bp.take<double>(2 * n); double* d_jo = bp.take<double>(12 * n); double* d_jc = bp.take<double>(12 * n);
#define H2D(dst, src, k) if (dst) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, (k), cudaMemcpyHostToDevice, s))
  H2D(d_to, T_obj, 96 * n); H2D(d_tc, T_cam, 96 * n); H2D(d_k, cam_k, 32 * n); H2D(d_p, p, 24 * n); H2D(d_uv, uv, 16 * n);
#undef H2D
  rc = launch_edge_linearize(ctx, n_edges, d_to, d_tc, d_k, d_p, d_uv, d_e, d_jo, d_jc, s);
  if (rc) return rc;
  if (err) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(err, d_e, 16 * n, cudaMemcpyDeviceToHost, s));
  if (J_obj && T_obj) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(J_obj, d_jo, 96 * n, cudaMemcpyDeviceToHost, s));
  if (J_cam) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(J_cam, d_jc, 96 * n, cudaMemcpyDeviceToHost, s));
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return SUO_OK;
}

int suo_slam_frame(suo_ctx* ctx, const uint8_t* image_hwc, int H, int W, const double* K_cam, const float* boxes, int L, int n_nonsym,
                   const double* model_kps, const uint8_t* model_mask, const double* diameter, const uint8_t* map_valid, const double* T_OtoG,
                   int n_views, int n_hist, const int32_t* hist_crop, const double* hist_T_GtoC, const double* hist_K, const int32_t* hist_off,
                   const double* hist_model_kp, const float* hist_uv, const float* hist_cov, double kp_var_thresh, double bbox_thresh,
                   double manual_kp_std, int init_with_outliers, uint64_t seed, double* T_GtoC, int32_t* status, double* T_pnp, uint8_t* kp_used,
                   uint8_t* ba_inliers, float* uv, float* cov, float* prior_uv, uint8_t* prior_mask, double* K_bbox, double* T_OtoG_out,
                   uint8_t* map_valid_out, uint8_t* reinit, int32_t* reinit_counts, const double* T_GtoC_init, int cam_init_mode,
                   int on_device, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  CtxExtra* x = X(ctx);
  if (!x->loaded) { ctx->set_error("suo_slam_frame before suo_load_weights", __FILE__, __LINE__); return SUO_E_STATE; }
  if (cam_init_mode < 0 || cam_init_mode > 2 || (cam_init_mode != 0) != (T_GtoC_init != nullptr)) {
    ctx->set_error("suo_slam_frame: cam_init_mode must be 0 (vote), 1 or 2 (camera pose given in T_GtoC_init)", __FILE__, __LINE__);
    return SUO_E_INVALID;
  }
  if (L <= 0 || L > ctx->max_crops || L > 128 || n_nonsym < 0 || n_nonsym > L || n_views < 1 || n_hist < 0 || !image_hwc || !K_cam || !boxes || !model_kps ||
      !model_mask || !diameter || !map_valid || !T_OtoG || (n_hist > 0 && (!hist_crop || !hist_T_GtoC || !hist_K || !hist_off || !hist_model_kp || !hist_uv)) ||
      !(manual_kp_std > 0)) {
    ctx->set_error("suo_slam_frame: bad crop count / null input", __FILE__, __LINE__);
    return SUO_E_INVALID;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  NetState& N = x->net;
  const int K = ctx->num_kp, n1 = n_nonsym, n2 = L - n_nonsym;
  const size_t LK = (size_t)L * K, n_im = (size_t)3 * H * W;
  size_t NH = 0;
  if (n_hist > 0 && !on_device) {
    for (int h = 0; h < n_hist; ++h) if (hist_off[h + 1] < hist_off[h] || hist_crop[h] < 0 || hist_crop[h] >= L) { ctx->set_error("suo_slam_frame: bad history arrays", __FILE__, __LINE__); return SUO_E_INVALID; }
    NH = (size_t)hist_off[n_hist];
  }
  const size_t in_bytes = n_im + (size_t)L * (16 + 8 + 1 + 96) + LK * 25 + 72 + 96 + (size_t)n_hist * (4 + 96 + 72 + 4) + 4 + NH * (24 + 8 + 16) + 32 * 256;
  const size_t out_bytes = (size_t)L * (128 + 72 + 96 + 1 + 1 + 8 + 4) + LK * (1 + 1 + 8 + 16 + 8 + 1) + 96 + 32 + 32 * 256;
  const size_t work_bytes = (size_t)L * (72 + 4 + 20) + LK * (24 + 16 + 4 + 4 + 4 + 4 + 32 + 24 + 16 + 32 + 1 + 4 + 16 + 1 + 1) + 4096 + 48 * 256;
  rc = x->fr.grow(ctx, in_bytes + out_bytes + work_bytes);
  if (!rc && !on_device) rc = x->fr.mirror(ctx, in_bytes + out_bytes);       // (inputs and outputs come first in the block)
  if (rc) return rc;
  rc = ensure_pnp_tables(ctx, L, s);
  if (rc) return rc;
  Bump bp{static_cast<uint8_t*>(x->fr.d)};
  // ---- inputs ----
  const uint8_t* d_im = image_hwc; const double* d_Kc = K_cam; const float* d_box = boxes; const double* d_mk = model_kps; const uint8_t* d_mm = model_mask;
  const double* d_diam = diameter; const uint8_t* d_mv = map_valid; const double* d_To = T_OtoG;
  const int32_t* d_hc = hist_crop; const double* d_hT = hist_T_GtoC; const double* d_hK = hist_K; const int32_t* d_ho = hist_off;
  const double* d_hm = hist_model_kp; const float* d_hu = hist_uv; const float* d_hcv = hist_cov;
  const double* d_Ti = T_GtoC_init;
  Stage stg(x->fr);
  if (!on_device) {
    // the frame goes straight from the caller's buffer (0.9 MB: one copy either way); the ~15 small arrays through the pinned mirror
    uint8_t* im_ = bp.take<uint8_t>(n_im);
    SUO_CUDA_TRY(ctx, cudaMemcpyAsync(im_, image_hwc, n_im, cudaMemcpyHostToDevice, s));
    d_im = im_;
#define STAGE(T, dst, src, n) { T* dst##_ = bp.take<T>(n); stg.in(dst##_, src, n); dst = dst##_; }
    STAGE(double, d_Kc, K_cam, 9) STAGE(float, d_box, boxes, 4 * (size_t)L) STAGE(double, d_mk, model_kps, 3 * LK)
    if (T_GtoC_init) { STAGE(double, d_Ti, T_GtoC_init, 12) }
    STAGE(uint8_t, d_mm, model_mask, LK) STAGE(double, d_diam, diameter, (size_t)L) STAGE(uint8_t, d_mv, map_valid, (size_t)L) STAGE(double, d_To, T_OtoG, 12 * (size_t)L)
    if (n_hist > 0) {
      STAGE(int32_t, d_hc, hist_crop, (size_t)n_hist) STAGE(double, d_hT, hist_T_GtoC, 12 * (size_t)n_hist) STAGE(double, d_hK, hist_K, 9 * (size_t)n_hist)
      STAGE(int32_t, d_ho, hist_off, (size_t)n_hist + 1) STAGE(double, d_hm, hist_model_kp, 3 * NH) STAGE(float, d_hu, hist_uv, 2 * NH)
      if (hist_cov) { STAGE(float, d_hcv, hist_cov, 4 * NH) }
    }
#undef STAGE
    SUO_CUDA_TRY(ctx, stg.send(s));
  }
  // ---- outputs (device side) ----
#define OUT(T, name, user, n) T* name = (on_device && user) ? user : bp.take<T>(n);
  OUT(double, o_cam, T_GtoC, 12) OUT(int32_t, o_st, status, 8) OUT(double, o_Tpnp, T_pnp, 16 * (size_t)L) OUT(uint8_t, o_used, kp_used, LK)
  OUT(uint8_t, o_bain, ba_inliers, LK) OUT(float, o_uv, uv, 2 * LK) OUT(float, o_cov, cov, 4 * LK) OUT(float, o_puv, prior_uv, 2 * LK)
  OUT(uint8_t, o_pm, prior_mask, LK) OUT(double, o_kb, K_bbox, 9 * (size_t)L) OUT(double, o_To, T_OtoG_out, 12 * (size_t)L)
  OUT(uint8_t, o_mv, map_valid_out, (size_t)L) OUT(uint8_t, o_ri, reinit, (size_t)L) OUT(int32_t, o_rc, reinit_counts, 2 * (size_t)L)
#undef OUT
  // ---- workspace ----
  double* w_kraw = bp.take<double>(9 * (size_t)L); float* w_mask = bp.take<float>(LK); int32_t* w_bi = bp.take<int32_t>(L);
  double* w_xs = bp.take<double>(3 * LK); double* w_ys = bp.take<double>(2 * LK); int32_t* w_cnt = bp.take<int32_t>(L); int32_t* w_kpi = bp.take<int32_t>(LK);
  int32_t* w_pst = bp.take<int32_t>(5 * (size_t)L);
  double* b_poses = bp.take<double>(12); uint8_t* b_fixed = bp.take<uint8_t>(1); int32_t* b_pv = bp.take<int32_t>(2); int32_t* b_vc = bp.take<int32_t>(1);
  int32_t* b_pe = bp.take<int32_t>(2); int32_t* b_ec = bp.take<int32_t>(1); int32_t* b_eo = bp.take<int32_t>(LK); int32_t* b_ecam = bp.take<int32_t>(LK);
  double* b_ck = bp.take<double>(4 * LK); double* b_p = bp.take<double>(3 * LK); double* b_uv = bp.take<double>(2 * LK); double* b_info = bp.take<double>(4 * LK);
  uint8_t* b_inl = bp.take<uint8_t>(LK); int32_t* b_src = bp.take<int32_t>(LK); double* b_err = bp.take<double>(2 * LK); uint8_t* b_lvl = bp.take<uint8_t>(LK);
  int8_t* b_fv = bp.take<int8_t>(LK); int32_t* b_its = bp.take<int32_t>(4); int32_t* b_st = bp.take<int32_t>(3);
  SUO_CUDA_TRY(ctx, cudaMemsetAsync(w_bi, 0, L * sizeof(int32_t), s));            // one image: every crop comes from frame 0
  SUO_CUDA_TRY(ctx, cudaMemsetAsync(o_st, 0, 8 * sizeof(int32_t), s));
  SUO_CUDA_TRY(ctx, cudaMemsetAsync(o_puv, 0, 2 * LK * sizeof(float), s));
  SUO_CUDA_TRY(ctx, cudaMemsetAsync(o_pm, 0, LK, s));
  SUO_CUDA_TRY(ctx, cudaMemsetAsync(w_cnt, 0, L * sizeof(int32_t), s));
  rc = launch_slam_kbbox(ctx, d_Kc, d_box, L, w_kraw, o_kb, s);
  if (rc) return rc;
  auto group = [&](int c0, int n, bool with_priors) -> int {      // forward + gate + PnP of crops [c0, c0 + n)
    const size_t o = (size_t)c0 * K;
    int r = forward_impl(ctx, d_im, 1, H, W, d_box + 4 * (size_t)c0, w_bi, n, nullptr, with_priors ? o_puv + 2 * o : nullptr, with_priors ? o_pm + o : nullptr,
                         o_uv + 2 * o, o_cov + 4 * o, nullptr, nullptr, nullptr, w_mask + o, nullptr, 1, stream, 1);
    if (r) return r;
    r = launch_gate_compact(ctx, o_uv + 2 * o, o_cov + 4 * o, w_mask + o, d_mm + o, d_mk + 3 * o, o_kb + 9 * (size_t)c0, n, K, (float)kp_var_thresh,
                            (float)bbox_thresh, w_xs + 3 * o, w_ys + 2 * o, w_cnt + c0, w_kpi + o, o_used + o, s);
    if (r) return r;
    // rows are addressed from the start of each crop's own slot, so the shifted base pointers keep offsets c * K valid
    return launch_pnp_batch_counts(ctx, w_xs + 3 * o, w_ys + 2 * o, x->pnp_off, w_cnt + c0, n, 0.001, seed, x->pnp_keys + c0, o_Tpnp + 16 * (size_t)c0,
                                   w_pst + 5 * (size_t)c0, s, K);
  };
  if (n1 > 0) { rc = group(0, n1, false); if (rc) return rc; }
  else {
    SUO_CUDA_TRY(ctx, cudaMemsetAsync(o_used, 0, LK, s));
  }
  if (cam_init_mode == 0) {
    rc = launch_slam_vote(ctx, n1, K, n_views == 1, o_Tpnp, w_cnt, w_kpi, w_xs, o_uv, o_cov, o_kb, d_diam, d_mv, d_To, manual_kp_std, 5.991, o_cam, o_st, s);
    if (rc) return rc;
  } else {
    // the caller knows the camera pose (external odometry :349-353, or __backup_estimate_camera_pose :933-973 before the passes / after a
    // failed vote): no vote; status = {1, 0, 0, ...}
    static const int32_t one = 1;
    SUO_CUDA_TRY(ctx, cudaMemcpyAsync(o_cam, d_Ti, 12 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    SUO_CUDA_TRY(ctx, cudaMemcpyAsync(o_st, &one, sizeof(one), cudaMemcpyHostToDevice, s));
  }
  if (n2 > 0) {
    rc = launch_slam_prior_uv(ctx, n1, L, K, o_st, o_cam, d_mv, d_To, d_mk, d_mm, w_kraw, o_puv, o_pm, s);
    if (rc) return rc;
    rc = group(n1, n2, true);
    if (rc) return rc;
    rc = launch_slam_drop_group(ctx, n1, L, K, o_st, w_cnt, o_used, o_Tpnp, s);
    if (rc) return rc;
  }
  rc = launch_slam_map_update(ctx, L, K, n_views, n_hist, o_st, o_cam, o_Tpnp, w_cnt, w_kpi, w_xs, o_uv, o_cov, o_kb, d_diam, d_mv, d_To, o_mv, o_To,
                              d_hc, d_hT, d_hK, d_ho, d_hm, d_hu, d_hcv, manual_kp_std, 5.991, o_rc, o_ri, s, cam_init_mode == 2 ? n1 : 0);
  if (rc) return rc;
  rc = launch_slam_ba_assemble(ctx, L, K, o_st, o_cam, o_mv, o_To, w_cnt, w_kpi, w_xs, o_uv, o_cov, o_kb, b_poses, b_fixed, b_pv, b_vc, b_pe, b_ec, b_eo, b_ecam,
                               b_ck, b_p, b_uv, b_info, b_inl, b_src, s);
  if (rc) return rc;
  static const int32_t its_host[2][4] = {{10, 10, 10, 10}, {10, 10, 40, 40}};   // curr_only in SLAM mode / in sfm_mode (object_slam.py:843-846)
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(b_its, its_host[ctx->opt_slam_sfm], sizeof(its_host[0]), cudaMemcpyHostToDevice, s));
  SUO_CUDA_TRY(ctx, cudaMemsetAsync(b_st, 0, 3 * sizeof(int32_t), s));
  rc = launch_ba_batch_scratch(ctx, 1, b_pv, b_pe, b_poses, b_fixed, b_eo, b_ecam, b_ck, b_p, b_uv, b_info, b_inl, b_its, 4, 2.4476519768340177 /* sqrt(5.991) */,
                               5.991, init_with_outliers, b_st, b_err, b_lvl, b_fv, s, b_vc, b_ec, 1 /* one camera vertex: warp kernel */);
  if (rc) return rc;
  rc = launch_slam_ba_scatter(ctx, L, K, b_ec, b_src, b_inl, b_poses, b_st, o_cam, o_bain, o_st, s);
  if (rc || on_device) return rc;
  // the outputs were taken back to back from the block: one copy into the pinned mirror, then memcpy into the caller's arrays
  stg.want(o_cam, 12); stg.want(o_rc, 2 * (size_t)L);
  SUO_CUDA_TRY(ctx, stg.fetch(s));
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  stg.out(T_GtoC, o_cam, 12); stg.out(status, o_st, 8); stg.out(T_pnp, o_Tpnp, 16 * (size_t)L); stg.out(kp_used, o_used, LK); stg.out(ba_inliers, o_bain, LK);
  stg.out(uv, o_uv, 2 * LK); stg.out(cov, o_cov, 4 * LK); stg.out(prior_uv, o_puv, 2 * LK); stg.out(prior_mask, o_pm, LK); stg.out(K_bbox, o_kb, 9 * (size_t)L);
  stg.out(T_OtoG_out, o_To, 12 * (size_t)L); stg.out(map_valid_out, o_mv, (size_t)L); stg.out(reinit, o_ri, (size_t)L); stg.out(reinit_counts, o_rc, 2 * (size_t)L);
  return suo_check_range(ctx);
}

int suo_frames_u8_submit(suo_ctx* ctx, int slot, const uint8_t* images_hwc, int n_img, int H, int W, const float* boxes,
                         const int32_t* box_img, int L, const double* model_kps, const uint8_t* model_mask, const double* K_bbox,
                         const double* diameter, double kp_var_thresh, double bbox_thresh, uint64_t seed, int run_ba,
                         double* T_pnp, double* T_ba, uint8_t* kp_used, uint8_t* ba_inliers, float* uv, float* cov,
                         void* records_dev, int record_id_base, void* stream) {
  if (slot < 0) return SUO_E_INVALID;
  return frames_impl(ctx, images_hwc, 1, n_img, H, W, boxes, box_img, L, nullptr, model_kps, model_mask, K_bbox, diameter, kp_var_thresh,
                     bbox_thresh, seed, run_ba, T_pnp, T_ba, kp_used, ba_inliers, uv, cov, 0, stream, slot, records_dev, record_id_base);
}

int suo_frames_wait(suo_ctx* ctx, int slot) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if (slot < 0 || slot > 1) return SUO_E_INVALID;
  CtxExtra::FrameSlot& sl = X(ctx)->slot[slot];
  if (!sl.pending) { ctx->set_error("suo_frames_wait: nothing submitted on this slot", __FILE__, __LINE__); return SUO_E_STATE; }
  SUO_CUDA_TRY(ctx, cudaEventSynchronize(sl.done));
  sl.pending = false;
  // (not suo_check_range: its blocking copy on the legacy stream would also wait for the NEXT batch, already queued behind this one, and the
  // copies of the batch after that could no longer overlap it)
  int flag = 0;
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(&flag, sl.flag, sizeof(int), cudaMemcpyDeviceToHost, X(ctx)->aux_stream));
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(X(ctx)->aux_stream));
  if (flag) {
    ctx->set_error("fp16x3 conv math: an activation exceeded the FP16 range (|x| > 6e4) in this batch; results are invalid, use SUO_OPT_CONV_MATH = 0 (tf32x3)", __FILE__, __LINE__);
    return SUO_E_RANGE;
  }
  return SUO_OK;
}

// ---- result records + the single exchange of the multi-GPU path -------------------------------------------------
size_t suo_record_bytes(int num_kp) { return num_kp > 0 ? (size_t)208 + 24 * (size_t)num_kp + (((size_t)num_kp + 7) & ~(size_t)7) : 0; }

int suo_pack_records(suo_ctx* ctx, const int32_t* crop_ids, int id_base, const double* T_pnp, const double* T_ba,
                     const uint8_t* kp_used, const uint8_t* ba_inliers, const float* uv, const float* cov, int L,
                     void* records, int on_device, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if (L <= 0 || !T_pnp || !records) return SUO_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int K = ctx->num_kp;
  const int rb = (int)suo_record_bytes(K);
  if (on_device)
    return launch_pack_records(ctx, crop_ids, id_base, T_pnp, T_ba, kp_used, ba_inliers, uv, cov, L, K, rb, static_cast<uint8_t*>(records), s);
  CtxExtra* x = X(ctx);
  const size_t LK = (size_t)L * K;
  rc = x->io.grow(ctx, (size_t)L * (4 + 28 * 8 + rb) + LK * 26 + 16 * 256);
  if (rc) return rc;
  Bump bp{static_cast<uint8_t*>(x->io.d)};
  int32_t* d_id = crop_ids ? bp.take<int32_t>(L) : nullptr;
  double* d_tp = bp.take<double>(16 * (size_t)L); double* d_tb = T_ba ? bp.take<double>(12 * (size_t)L) : nullptr;
  uint8_t* d_u = kp_used ? bp.take<uint8_t>(LK) : nullptr; uint8_t* d_b = ba_inliers ? bp.take<uint8_t>(LK) : nullptr;
  float* d_uv = uv ? bp.take<float>(2 * LK) : nullptr; float* d_cov = cov ? bp.take<float>(4 * LK) : nullptr;
  uint8_t* d_rec = bp.take<uint8_t>((size_t)L * rb);
#define H2D(dst, src, n) if (dst) SUO_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, (n), cudaMemcpyHostToDevice, s))
  H2D(d_id, crop_ids, 4 * (size_t)L); H2D(d_tp, T_pnp, 128 * (size_t)L); H2D(d_tb, T_ba, 96 * (size_t)L);
  H2D(d_u, kp_used, LK); H2D(d_b, ba_inliers, LK); H2D(d_uv, uv, 8 * LK); H2D(d_cov, cov, 16 * LK);
#undef H2D
  rc = launch_pack_records(ctx, d_id, id_base, d_tp, d_tb, d_u, d_b, d_uv, d_cov, L, K, rb, d_rec, s);
  if (rc) return rc;
  SUO_CUDA_TRY(ctx, cudaMemcpyAsync(records, d_rec, (size_t)L * rb, cudaMemcpyDeviceToHost, s));
  SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  return SUO_OK;
}

int suo_allgather_results(suo_ctx* ctx, void* nccl_comm, const void* records, size_t rec_bytes, int n_local, void* out, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if (!nccl_comm || !records || !out || n_local <= 0 || rec_bytes == 0) return SUO_E_INVALID;
  CtxExtra* x = X(ctx);
  if (!x->nccl_allgather) {
    // the communicator was created by the caller's NCCL: use that same library (already loaded), never a second copy
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW);
    x->nccl_allgather = h ? dlsym(h, "ncclAllGather") : nullptr;
    if (!x->nccl_allgather) { ctx->set_error("suo_allgather_results: libnccl.so.2 / ncclAllGather not found", __FILE__, __LINE__); return SUO_E_STATE; }
  }
  // ncclResult_t ncclAllGather(const void* sendbuff, void* recvbuff, size_t sendcount, ncclDataType_t, ncclComm_t, cudaStream_t); ncclChar = 0
  typedef int (*AllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
  const int r = reinterpret_cast<AllGatherFn>(x->nccl_allgather)(records, out, rec_bytes * (size_t)n_local, 0, nccl_comm, static_cast<cudaStream_t>(stream));
  if (r != 0) { ctx->set_error("ncclAllGather failed: ncclResult_t " + std::to_string(r), __FILE__, __LINE__); return SUO_E_CUDA; }
  ctx->launches++;
  return SUO_OK;
}

// Per-op device timing of the network program (eager launches, one CUDA event pair per op) on
// whatever the input buffer currently holds: average milliseconds per forward spent in conv
// kernels and in the other (pool / up-sample) kernels.  Used by bench.py for the roofline block.
int suo_profile_network(suo_ctx* ctx, int L, int with_priors, int iters, float* conv_ms, float* other_ms, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  CtxExtra* x = X(ctx);
  if (!x->loaded || L <= 0 || L > ctx->max_crops || iters <= 0) return SUO_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  NetState& N = x->net;
  const int variant = with_priors ? 1 : 0;
  std::vector<size_t> idx;
  for (size_t i = 0; i < N.ops.size(); ++i) if (N.ops[i].variant == 2 || N.ops[i].variant == variant) idx.push_back(i);
  std::vector<cudaEvent_t> ev(idx.size() + 1);
  for (auto& e : ev) SUO_CUDA_TRY(ctx, cudaEventCreate(&e));
  double conv = 0, other = 0;
  const int R = ctx->crop_res;
  for (int it = 0; it < iters; ++it) {
    SUO_CUDA_TRY(ctx, cudaEventRecord(ev[0], s));
    for (size_t q = 0; q < idx.size(); ++q) {
      // run exactly one op by temporarily restricting the program
      const OpDesc& o = N.ops[idx[q]];
      const BufDesc& bi = N.bufs[o.in];
      const BufDesc& bo = N.bufs[o.out];
      if (o.type == OP_CONV) {
        ConvParams p{};
        fill_conv_params(ctx, N, idx[q], L, ctx->opt_backend, ctx->opt_passes, p);
        rc = ctx->opt_backend == 1 ? launch_conv_tc(ctx, p, ctx->opt_passes, s) : launch_conv_simt(ctx, p, s);
      } else if (o.type == OP_MAXPOOL) {
        rc = launch_maxpool2(ctx, N.act[o.in], L, R / bi.div, R / bi.div, bi.C, N.act[o.out], s);
      } else {
        rc = launch_upsample_add(ctx, N.act[o.in], N.act[o.res], L, R / bo.div, R / bo.div, bo.C, N.act[o.out], s);
      }
      if (rc) return rc;
      SUO_CUDA_TRY(ctx, cudaEventRecord(ev[q + 1], s));
    }
    SUO_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    FILE* dump = nullptr;
    if (it == iters - 1) { if (const char* path = getenv("SUO_PROFILE_DUMP")) dump = fopen(path, "w"); }
    if (dump) fprintf(dump, "op,type,mode,side_out,Cin,Cout,K,relu,res,pre,ms,gflop\n");
    for (size_t q = 0; q < idx.size(); ++q) {
      float ms = 0;
      SUO_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ev[q], ev[q + 1]));
      const OpDesc& o = N.ops[idx[q]];
      (o.type == OP_CONV ? conv : other) += ms;
      if (dump) {
        const int side = R / N.bufs[o.out].div;
        double gf = o.type == OP_CONV ? 2.0 * L * side * side * (double)o.Cout_pad * o.K * 1e-9 : 0.0;
        fprintf(dump, "%zu,%d,%d,%d,%d,%d,%d,%d,%d,%d,%.4f,%.3f\n", idx[q], o.type, o.mode, side, o.Cin, o.Cout, o.K, o.relu, o.res >= 0, o.pre_off >= 0, ms, gf);
      }
    }
    if (dump) fclose(dump);
  }
  for (auto& e : ev) cudaEventDestroy(e);
  if (conv_ms) *conv_ms = (float)(conv / iters);
  if (other_ms) *other_ms = (float)(other / iters);
  return SUO_OK;
}

}  // extern "C"
