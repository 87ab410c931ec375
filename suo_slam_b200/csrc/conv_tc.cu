// tcgen05 tensor-core implicit-GEMM convolution for sm_100a (conv backend 1).
//
// Replaces the cuDNN/MKL-DNN convolutions the reference reaches through
// torch.nn.Conv2d in lib/models/layers/Residual.py:9-18 and lib/models/hg.py:67,80-86
// (1x1, 3x3 pad 1, 7x7 stride 2 pad 3) together with the BatchNorm/ReLU/residual-add
// glue around them (Residual.py:20-35), which is fused here as a prologue/epilogue.
//
//   D[m, n] = sum_k A[m, k] * Wt[n, k]      m = output pixel, n = output channel
//
// Precision: the reference computes these convs in FP32.  tcgen05 has no FP32 MMA, so
// every FP32 operand is split into two TF32 numbers, x = hi + lo (hi = rna_tf32(x),
// lo = rna_tf32(x - hi)), and three kind::tf32 MMAs accumulate hi*hi + lo*hi + hi*lo into
// the FP32 TMEM accumulator ("3xTF32", relative error ~2^-21 per product, i.e. FP32
// grade).  tf32_passes == 1 issues only hi*hi (plain TF32, 2^-11).
//
// Per CTA: one 128 x BN output tile, K marched in 32-float chunks through a 3-stage
// shared-memory ring.  Warp roles (10 warps):
//   warp 0      : weight producer — one thread issues cp.async.bulk (TMA bulk copy, UBLKCP)
//                 of the pre-swizzled hi/lo weight images of the chunk, completes on an mbarrier
//   warp 1      : TMEM allocator + MMA issuer — one thread issues tcgen05.mma (UTCMMA),
//                 tcgen05.commit releases the smem stage / signals the epilogue
//   warps 2..9  : A producers — coalesced 128-bit global loads of the NHWC activations
//                 (im2col addressing with zero fill for padding), optional pre-activation
//                 BN affine + ReLU, TF32 hi/lo split, stores into the 128B-swizzled K-major
//                 operand tiles, fence.proxy.async + mbarrier arrive;
//                 afterwards the same warps run the epilogue: tcgen05.ld TMEM -> registers,
//                 + bias, ReLU, + residual, 128-bit stores (NHWC) or coalesced plane stores (NCHW).
#include <cuda_fp16.h>
#include <algorithm>
#include <cstring>
#include "conv_gather.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;          // floats per chunk = one 128-byte swizzle row
constexpr int NUM_STAGES = 3;
constexpr int NUM_PRODUCER_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_PRODUCER_WARPS;
constexpr int PREFETCH = 2;          // chunks of A kept in flight in registers
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 4;   // 16 KB


template <int BN>
struct Smem {
  static constexpr int B_TILE_BYTES = BN * BLOCK_K * 4;
  static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
  static constexpr int BAR_OFFSET = NUM_STAGES * STAGE_BYTES;
  static constexpr int STAGING_OFFSET = BAR_OFFSET + 128;           // epilogue transpose buffers: 4 warps x 32 x 36 floats
  static constexpr int STAGING_BYTES = 4 * 32 * 36 * 4;
  static constexpr int TOTAL = STAGING_OFFSET + STAGING_BYTES + 1024;   // + alignment slack
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const ConvParams p, const int passes) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  using S = Smem<BN>;
  const uint32_t bar_base = smem_base + S::BAR_OFFSET;
  auto full_a = [&](int s) { return bar_base + 8 * s; };
  auto full_b = [&](int s) { return bar_base + 8 * (NUM_STAGES + s); };
  auto empty = [&](int s) { return bar_base + 8 * (2 * NUM_STAGES + s); };
  const uint32_t tmem_full = bar_base + 8 * (3 * NUM_STAGES);
  const uint32_t tmem_slot = bar_base + 8 * (3 * NUM_STAGES + 1);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + S::BAR_OFFSET + 8 * (3 * NUM_STAGES + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = p.B * p.Ho * p.Wo;
  const int m0 = blockIdx.x * BLOCK_M;
  const int n_tile = blockIdx.y;
  const int nchunks = p.K / BLOCK_K;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NUM_STAGES; ++s) {
      mbar_init(full_a(s), NUM_PRODUCER_WARPS);
      mbar_init(full_b(s), 1);
      mbar_init(empty(s), 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);   // [0,BN): hi*hi accumulator, [BN,2BN): correction terms
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_gen;

  if (warp == 0) {
    // ===================== weight producer (TMA bulk copies) =====================
    if (lane == 0) {
      const uint32_t bytes_per_part = S::B_TILE_BYTES;
      const float* src = p.w_packed + (size_t)n_tile * nchunks * 2 * (BN * BLOCK_K);
      for (int j = 0; j < nchunks; ++j) {
        const int s = j % NUM_STAGES;
        const uint32_t ph = (j / NUM_STAGES) & 1;
        mbar_wait(empty(s), ph ^ 1);
        const uint32_t dst = smem_base + s * S::STAGE_BYTES + 2 * A_TILE_BYTES;
        const uint32_t nbytes = passes == 3 ? 2 * bytes_per_part : bytes_per_part;
        mbar_arrive_expect_tx(full_b(s), nbytes);
        bulk_g2s(dst, src + (size_t)j * 2 * (BN * BLOCK_K), nbytes, full_b(s));   // hi image (and lo image right behind it)
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_tf32(BLOCK_M, BN);
    for (int j = 0; j < nchunks; ++j) {
      const int s = j % NUM_STAGES;
      const uint32_t ph = (j / NUM_STAGES) & 1;
      mbar_wait(full_a(s), ph);
      mbar_wait(full_b(s), ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_hi = smem_base + s * S::STAGE_BYTES;
        const uint32_t a_lo = a_hi + A_TILE_BYTES;
        const uint32_t b_hi = a_hi + 2 * A_TILE_BYTES;
        const uint32_t b_lo = b_hi + S::B_TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < BLOCK_K / 8; ++kk) {
          const uint64_t dah = make_sw128_desc(a_hi + kk * 32), dbh = make_sw128_desc(b_hi + kk * 32);
          umma_tf32(tmem_acc, dah, dbh, idesc, (j | kk) != 0);
          if (passes == 3) {
            // The tensor core's FP32 accumulate truncates: measured error grows linearly with the number of
            // accumulation steps into one accumulator (~1.4e-8 of max per step).  The small lo*hi + hi*lo
            // terms therefore go to their own accumulator (their truncation error is 2^-11 smaller) and are
            // added in FP32 by the epilogue; the main accumulator sees only K/8 steps.
            const uint64_t dal = make_sw128_desc(a_lo + kk * 32), dbl = make_sw128_desc(b_lo + kk * 32);
            umma_tf32(tmem_acc + BN, dal, dbh, idesc, (j | kk) != 0);
            umma_tf32(tmem_acc + BN, dah, dbl, idesc, 1u);
          }
        }
        umma_commit(empty(s));                       // frees the smem stage once these MMAs retire
        if (j == nchunks - 1) umma_commit(tmem_full);  // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ===================== A producers, then epilogue =====================
    const int pt = threadIdx.x - 64;            // 0..255
    const int g = pt & 7;                       // float4 group inside the 128-byte row
    const int r0 = pt >> 3;                     // rows r0 + 32*i
    PixelCoord pc[4];
    bool rok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + r0 + 32 * i;
      rok[i] = m < M;
      pc[i] = decode_pixel(rok[i] ? m : 0, p.Ho, p.Wo);
    }
    float4 v[PREFETCH][4];
    auto load_chunk = [&](int j, float4 (&dst)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t vm;
        const float* ptr = chunk_ptr(p, pc[i], j, vm);
        dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rok[i] && ((vm >> g) & 1u)) dst[i] = __ldg(reinterpret_cast<const float4*>(ptr + 4 * g));
      }
    };
#pragma unroll
    for (int q = 0; q < PREFETCH; ++q)
      if (q < nchunks) load_chunk(q, v[q]);

    auto produce = [&](int j, float4 (&buf)[4]) {
      const int s = j % NUM_STAGES;
      const uint32_t ph = (j / NUM_STAGES) & 1;
      float4 cur[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) cur[i] = buf[i];
      if (p.pre_scale) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(p.pre_scale + 32 * j + 4 * g));
        const float4 sh = __ldg(reinterpret_cast<const float4*>(p.pre_shift + 32 * j + 4 * g));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (rok[i]) {
            cur[i].x = fmaxf(fmaf(cur[i].x, sc.x, sh.x), 0.f); cur[i].y = fmaxf(fmaf(cur[i].y, sc.y, sh.y), 0.f);
            cur[i].z = fmaxf(fmaf(cur[i].z, sc.z, sh.z), 0.f); cur[i].w = fmaxf(fmaf(cur[i].w, sc.w, sh.w), 0.f);
          }
        }
      }
      if (j + PREFETCH < nchunks) load_chunk(j + PREFETCH, buf);   // keep PREFETCH chunks of loads in flight
      mbar_wait(empty(s), ph ^ 1);
      uint8_t* a_hi = smem_gen + s * S::STAGE_BYTES;
      uint8_t* a_lo = a_hi + A_TILE_BYTES;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + 32 * i;
        const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((g ^ (r & 7)) << 4);
        float4 hi;
        hi.x = tf32_rna(cur[i].x); hi.y = tf32_rna(cur[i].y); hi.z = tf32_rna(cur[i].z); hi.w = tf32_rna(cur[i].w);
        *reinterpret_cast<float4*>(a_hi + off) = hi;
        if (passes == 3) {
          float4 lo;
          lo.x = tf32_rna(cur[i].x - hi.x); lo.y = tf32_rna(cur[i].y - hi.y);
          lo.z = tf32_rna(cur[i].z - hi.z); lo.w = tf32_rna(cur[i].w - hi.w);
          *reinterpret_cast<float4*>(a_lo + off) = lo;
        }
      }
      fence_proxy_async();      // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(full_a(s));
    };
    static_assert(PREFETCH == 2, "producer loop is unrolled for two register buffers");
    for (int j = 0; j < nchunks; j += 2) {
      produce(j, v[0]);
      if (j + 1 < nchunks) produce(j + 1, v[1]);
    }

    // ---- epilogue: TMEM -> registers -> global ----
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;             // which half of the BN columns
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const int HW = p.Ho * p.Wo;
    constexpr int COLS_PER_WARP = BN / 2;
#pragma unroll 1
    for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 32) {
      const int col = half * COLS_PER_WARP + c0;   // column inside the tile
      uint32_t r[32];
      tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)col, r);
      tmem_ld_wait();
      if (passes == 3) {
        uint32_t rc[32];
        tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + col), rc);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) + __uint_as_float(rc[c]));
      }
      const int n_base = n_tile * BN + col;
      if (m < M) {
        if (!p.out_nchw) {
          float* __restrict__ orow = p.out + (size_t)m * p.out_c + n_base;
          const float* __restrict__ rrow = p.residual ? p.residual + (size_t)m * p.out_c + n_base : nullptr;
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            if (n_base + c < p.Cout) {
              const float4 bq = __ldg(reinterpret_cast<const float4*>(p.bias + n_base + c));
              float4 o;
              o.x = __uint_as_float(r[c]) + bq.x; o.y = __uint_as_float(r[c + 1]) + bq.y;
              o.z = __uint_as_float(r[c + 2]) + bq.z; o.w = __uint_as_float(r[c + 3]) + bq.w;
              if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              if (rrow) {
                const float4 rq = *reinterpret_cast<const float4*>(rrow + c);
                o.x += rq.x; o.y += rq.y; o.z += rq.z; o.w += rq.w;
              }
              *reinterpret_cast<float4*>(orow + c) = o;
            }
          }
        } else {
          const int b = m / HW, rem = m - b * HW;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int n = n_base + c;
            if (n < p.Cout) {
              float o = __uint_as_float(r[c]) + __ldg(p.bias + n);
              if (p.relu) o = fmaxf(o, 0.f);
              p.out[((size_t)b * p.Cout + n) * HW + rem] = o;   // lanes = consecutive pixels of plane n: coalesced
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, 2 * BN);
  }
}

// =====================================================================================================
// Persistent variant: grid = #SMs, every CTA loops over output tiles.  The smem ring and the A/weight
// producers run continuously across tile boundaries, the accumulator is double-buffered in TMEM and a
// dedicated epilogue warpgroup drains tile i while the tensor core already works on tile i+1.
//   warp 0      weight producer (TMA bulk copies)          warp 1      TMEM alloc + MMA issuer
//   warps 2..5  epilogue (one TMEM lane quarter each)      warps 6..13 A producers
// TMEM columns: buffer b in {0,1} at b*2*BN: [main | correction].
constexpr int P_NUM_THREADS = 32 * (2 + 4 + NUM_PRODUCER_WARPS);

// Shared-memory plan of the persistent kernel.  EPI selects the epilogue:
//   0  registers -> padded transpose buffer -> st.global (any layout: NCHW logits, unaligned channel counts, TF32 modes)
//   1  TMA-store epilogue: each epilogue warp owns NBUF 4 KB boxes (32 rows x 128 B, 128B swizzle) that it fills from TMEM
//      and hands to cp.async.bulk.tensor stores (and, with a residual, fills first by TMA loads); 3-stage operand ring
//   2  same, tuned for the HBM-bound 1x1 convs that add a skip tensor (K <= 256: two chunk stages are enough):
//      2-stage ring and 5 boxes per warp so that the residual of the next ~3 column groups is always in flight
//   3  plan 1's epilogue for the register-fed 1x1 convs (FP32 input, optional BN+ReLU prologue): the FP32 activations
//      come by TMA into a ring of four 16 KB boxes (128 pixels x 32 floats) and the producer warps only transform
//      shared -> shared (prologue, FP16 hi/lo' split); 2-stage operand ring
constexpr int RAW_BOXES = 4;
constexpr int RAW_BOX_BYTES = BLOCK_M * 32 * 4;
template <int BN, int EPI>
struct PSmem {
  static constexpr int NS = (EPI == 2 || EPI == 3) ? 2 : 3;
  static constexpr int NBUF = (EPI == 2) ? 5 : 2;
  static constexpr int RAW_OFFSET = 2 * (2 * A_TILE_BYTES + 2 * BN * BLOCK_K * 4);      // plan 3: raw ring right behind the 2 operand stages
  static constexpr int RING_BYTES = (EPI == 3) ? RAW_OFFSET + RAW_BOXES * RAW_BOX_BYTES : NS * (2 * A_TILE_BYTES + 2 * BN * BLOCK_K * 4);
  static constexpr int B_TILE_BYTES = BN * BLOCK_K * 4;
  static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
  static constexpr int STAGING_OFFSET = RING_BYTES;                                        // 1024-byte aligned
  static constexpr int STAGING_BYTES = (EPI == 0) ? 4 * 32 * 36 * 4 : 4 * NBUF * 4096;
  static constexpr int BAR_OFFSET = STAGING_OFFSET + STAGING_BYTES;                       // 512 B of mbarriers
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 512;                                    // bias of the layer (<= 256 floats)
  static constexpr int TOTAL = BIAS_OFFSET + 1024 + 1024;                                 // + alignment slack
  static_assert(TOTAL <= 232448, "exceeds the 227 KB a CTA may use");
};
constexpr int P_PREFETCH = 4;         // chunks of A kept in flight in registers (load latency ~2-3k cycles under load)

// ---- A-producer addressing (persistent kernel) -------------------------------------------------------
// Per tile each producer thread decodes its 4 rows once (out of line: it runs once per ~10-40 chunks and
// would otherwise be inlined into every unrolled copy of the chunk loop); per chunk only `off`/`tap` move.
struct LoadCursor {
  int t, j, tap, cc;                        // tile, chunk, 3x3 tap / stem kernel row, 32-float sub-chunk
  ptrdiff_t off;                            // running float offset of this chunk from the row base
  const float* base[4];
  uint32_t mask[4];                         // 1x1: bit0 = row valid; 3x3: bit t = tap t in bounds; stem: bits0-6 ky ok, bits8-14 pixel ok
};

template <int MODE>
__device__ __noinline__ void cursor_set_tile(LoadCursor& c, const ConvParams& p, int M, int num_tiles, int num_n_tiles, int r0) {
  const int m_tile = c.t / num_n_tiles;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m_tile * BLOCK_M + r0 + 32 * i;
    const bool ok = (c.t < num_tiles) && (m < M);
    const PixelCoord pc = decode_pixel(ok ? m : 0, p.Ho, p.Wo);
    uint32_t msk = 0;
    if (MODE == CONV_1x1) {
      c.base[i] = p.in + (size_t)(ok ? m : 0) * p.Cin;
      msk = ok ? 1u : 0u;
    } else if (MODE == CONV_3x3) {
      c.base[i] = p.in + ((size_t)(pc.b * p.H + pc.oy) * p.W + pc.ox) * p.Cin;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int iy = pc.oy + t / 3 - 1, ix = pc.ox + t % 3 - 1;
        if (ok && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) msk |= 1u << t;
      }
    } else {
      const int iy0 = 2 * pc.oy - 3, ix0 = 2 * pc.ox - 3;
      c.base[i] = p.in + ((ptrdiff_t)(pc.b * p.H + iy0) * p.W + ix0) * p.Cin;
#pragma unroll
      for (int t = 0; t < 7; ++t) {
        if (ok && iy0 + t >= 0 && iy0 + t < p.H) msk |= 1u << t;
        if (ok && ix0 + t >= 0 && ix0 + t < p.W) msk |= 1u << (8 + t);
      }
    }
    c.mask[i] = msk;
  }
  c.j = 0; c.tap = 0; c.cc = 0;
  c.off = (MODE == CONV_3x3) ? -(ptrdiff_t)(p.W + 1) * p.Cin : 0;
}

// A_TMA: the A operand comes straight from HBM by TMA (cp.async.bulk.tensor, 128B swizzle, hardware zero fill
// for the 3x3 padding) out of a tensor whose producer layer already wrote it as two FP16 planes (hi, lo'):
// no A-producer warps, no register staging, no proxy fences — the CTA is 6 warps.
template <int BN, int MODE, bool PRE, int MATH, bool A_TMA, int EPI>
// 14 warps -> one SMSP hosts 4 of them -> the 16K-register SMSP file caps every thread at 128 registers
__global__ void __launch_bounds__(A_TMA ? 192 : (EPI == 3 ? P_NUM_THREADS + 32 : P_NUM_THREADS), 1)
conv_tc_persistent_kernel(const __grid_constant__ ConvParams p, const int passes, const int num_m_tiles, const int num_n_tiles) {
  static_assert(!A_TMA || (MATH == MATH_F16 && !PRE && MODE != CONV_STEM7), "TMA-fed A needs pre-split FP16 activations");
  // MATH_TF32: chunk = 32 floats, operands FP32 words read as TF32 (hi = top 19 bits, lo = x - hi).
  // MATH_F16 : chunk = 64 floats, operands FP16: x = hi + 2^-11 lo' with hi = fp16(x with 13 low mantissa bits
  //            cleared) and lo' = fp16((x - hi) * 2^11): 22 mantissa bits in two FP16 numbers; the correction
  //            accumulator is scaled by 2^-11 in the epilogue.  Same 128-byte swizzled rows, twice the K per byte.
  constexpr int KC = (MATH == MATH_F16) ? 64 : 32;       // K elements per chunk
  constexpr int NV = KC / 32;                            // float4 loads per row per thread per chunk
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  using S = PSmem<BN, EPI>;
  constexpr int NS = S::NS;                      // operand ring depth
  static_assert(EPI == 0 || MATH == MATH_F16, "the TMA-store epilogue is built for the fp16x3 path");
  constexpr bool RAW = (EPI == 3);               // A operand: FP32 tensor -> TMA -> raw ring -> producer warps -> operand ring
  static_assert(!RAW || (MODE == CONV_1x1 && !A_TMA), "raw-TMA A path: register-fed 1x1 convs only");
  const uint32_t bar_base = smem_base + S::BAR_OFFSET;
  auto full_a = [&](int s) { return bar_base + 8 * s; };
  auto full_b = [&](int s) { return bar_base + 8 * (NS + s); };
  auto empty = [&](int s) { return bar_base + 8 * (2 * NS + s); };
  auto tmem_full = [&](int b) { return bar_base + 8 * (3 * NS + b); };
  auto tmem_empty = [&](int b) { return bar_base + 8 * (3 * NS + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8 * (3 * NS + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + S::BAR_OFFSET + 8 * (3 * NS + 4));
  auto raw_full = [&](uint32_t b) { return bar_base + 8 * (40 + b); };
  auto raw_empty = [&](uint32_t b) { return bar_base + 8 * (40 + RAW_BOXES + b); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = p.B * p.Ho * p.Wo;
  const int nchunks = p.K / KC;
  const int num_tiles = num_m_tiles * num_n_tiles;

  unsigned long long trace_slot = 0;
  if (p.trace && threadIdx.x == 0 && blockIdx.x == 0) {
    trace_slot = atomicAdd(p.trace, 1ull);
    if (trace_slot < 8192) {
      unsigned long long t; uint32_t smid;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.trace[1 + 4 * trace_slot] = t; p.trace[1 + 4 * trace_slot + 2] = gridDim.x; p.trace[1 + 4 * trace_slot + 3] = smid;
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(full_a(s), A_TMA ? 1 : NUM_PRODUCER_WARPS);
      mbar_init(full_b(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 4); }
    if (EPI != 0)
      for (int b = 0; b < 4 * S::NBUF; ++b) mbar_init(bar_base + 8 * (16 + b), 1);     // residual boxes: [warp][buffer]
    if (RAW)
      for (int b = 0; b < RAW_BOXES; ++b) { mbar_init(raw_full(b), 1); mbar_init(raw_empty(b), NUM_PRODUCER_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 4 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // Programmatic dependent launch (SUO_OPT_PDL): every CTA of this grid is resident (grid <= #SMs, one CTA per SM), so the next
  // conv kernel may be scheduled onto an SM the moment this kernel's CTA leaves it and run its own prologue (barriers, TMEM) there;
  // nothing above touched an activation tensor — from here on the previous kernel must have completed and flushed.
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== weight producer =====================
    if (lane == 0) {
      const uint32_t nbytes = (passes == 3 ? 2u : 1u) * S::B_TILE_BYTES;
      uint32_t g = 0;
      const int HWi = p.H * p.W, cpc64 = p.Cin / 64;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_tile = t / num_n_tiles, n_tile = t - m_tile * num_n_tiles;
        const float* src = p.w_packed + (size_t)n_tile * nchunks * 2 * (BN * BLOCK_K);
        // tile origin in (b, y) — tiles cover whole image rows (128 % W == 0 or W % 128 == 0)
        const int m0 = m_tile * BLOCK_M;
        const int b0 = m0 / HWi, rem = m0 - b0 * HWi;
        const int y0 = rem / p.W, x0 = rem - y0 * p.W;
        for (int j = 0; j < nchunks; ++j, ++g) {
          const int s = g % NS;
          const uint32_t ph = (g / NS) & 1;
          mbar_wait(empty(s), ph ^ 1);
          if (p.dbg && blockIdx.x == 0 && g < 512) p.dbg[4 * 512 + g] = clock64();
          const uint32_t a_dst = smem_base + s * S::STAGE_BYTES;
          const uint32_t dst = a_dst + 2 * A_TILE_BYTES;
          if (A_TMA) {
            const int tap = (MODE == CONV_3x3) ? j / cpc64 : 0, cc = (MODE == CONV_3x3) ? j - tap * cpc64 : j;
            const int dy = (MODE == CONV_3x3) ? tap / 3 - 1 : 0, dx = (MODE == CONV_3x3) ? tap % 3 - 1 : 0;
            mbar_arrive_expect_tx(full_b(s), nbytes + (passes == 3 ? 2u : 1u) * A_TILE_BYTES);
            tma_load_4d(a_dst, p.tmap_hi, 64 * cc, x0 + dx, y0 + dy, b0, full_b(s));
            if (passes == 3) tma_load_4d(a_dst + A_TILE_BYTES, p.tmap_lo, 64 * cc, x0 + dx, y0 + dy, b0, full_b(s));
            mbar_arrive(full_a(s));
          } else {
            mbar_arrive_expect_tx(full_b(s), nbytes);
          }
          bulk_g2s(dst, src + (size_t)j * 2 * (BN * BLOCK_K), nbytes, full_b(s));
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = (MATH == MATH_F16) ? make_idesc_f16(BLOCK_M, BN) : make_idesc_tf32(BLOCK_M, BN);
    // hi and lo' weight images of a chunk are adjacent in shared memory = ONE K-major operand of 2*BN rows: a single
    // N = 2*BN MMA forms A_hi*B_hi (main accumulator) and A_hi*B_lo' (correction accumulator, the next BN TMEM columns)
    // and reads A_hi once instead of twice; A_lo'*B_hi then adds into the correction columns.
    constexpr uint32_t idesc2 = (MATH == MATH_F16) ? make_idesc_f16(BLOCK_M, 2 * BN) : make_idesc_tf32(BLOCK_M, 2 * BN);
    const bool merge = p.mma_merge && passes == 3;
    uint32_t g = 0, i = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++i) {
      const uint32_t b = i & 1;
      mbar_wait(tmem_empty(b), ((i >> 1) & 1) ^ 1);        // epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t acc = tmem_base + b * 2 * BN;
      for (int j = 0; j < nchunks; ++j, ++g) {
        const int s = g % NS;
        const uint32_t ph = (g / NS) & 1;
        mbar_wait(full_a(s), ph);
        if (p.dbg && blockIdx.x == 0 && lane == 0 && g < 512) p.dbg[2 * 512 + g] = clock64();
        mbar_wait(full_b(s), ph);
        if (p.dbg && blockIdx.x == 0 && lane == 0 && g < 512) p.dbg[3 * 512 + g] = clock64();
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_hi = smem_base + s * S::STAGE_BYTES;
          const uint32_t a_lo = a_hi + A_TILE_BYTES;
          const uint32_t b_hi = a_hi + 2 * A_TILE_BYTES;
          const uint32_t b_lo = b_hi + S::B_TILE_BYTES;
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / 8; ++kk) {
            const uint64_t dah = make_sw128_desc(a_hi + kk * 32), dbh = make_sw128_desc(b_hi + kk * 32);
            if (merge) {
              const uint64_t dal = make_sw128_desc(a_lo + kk * 32);
              umma<MATH>(acc, dah, dbh, idesc2, (j | kk) != 0);
              umma<MATH>(acc + BN, dal, dbh, idesc, 1u);
            } else {
              umma<MATH>(acc, dah, dbh, idesc, (j | kk) != 0);
              if (passes == 3) {
                const uint64_t dal = make_sw128_desc(a_lo + kk * 32), dbl = make_sw128_desc(b_lo + kk * 32);
                umma<MATH>(acc + BN, dal, dbh, idesc, (j | kk) != 0);
                umma<MATH>(acc + BN, dah, dbl, idesc, 1u);
              }
            }
          }
          umma_commit(empty(s));
          if (j == nchunks - 1) umma_commit(tmem_full(b));
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue warpgroup =====================
    if constexpr (EPI != 0) {
      // TMEM -> registers -> (bias, ReLU, + residual) -> the warp's own 32-row x 128-byte staging boxes -> TMA store.
      // A lane owns one output row (that is how tcgen05.ld hands the accumulator out); the 16-byte chunk c of row r
      // lives at chunk c ^ (r & 7) — the 128B swizzle of the tensor maps — so the 32 lanes' float4 accesses are
      // bank-conflict free and the TMA engine does the global coalescing.  No global load/store is issued by these
      // warps: the skip tensor arrives by TMA into the same boxes (NBUF - 2 column groups ahead), results leave by TMA.
      constexpr int NBUF = S::NBUF;
      constexpr int D = NBUF - 2;                       // residual prefetch distance in column groups
      const int ew = warp - 2, q = warp & 3;            // staging owner index / TMEM lane quarter
      float* bias_s = reinterpret_cast<float*>(smem_gen + S::BIAS_OFFSET);
      for (int c = threadIdx.x - 64; c < p.Cout_pad; c += 128) bias_s[c] = __ldg(p.bias + c);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const uint32_t stg_base = smem_base + S::STAGING_OFFSET + ew * (NBUF * 4096);
      uint8_t* stg_gen = smem_gen + S::STAGING_OFFSET + ew * (NBUF * 4096);
      auto res_bar = [&](uint32_t bf) { return bar_base + 8 * (16 + ew * NBUF + bf); };
      const bool has_res = p.residual != nullptr && !p.out_split;
      const int swz = lane & 7;
      uint32_t i = 0;
      if (!p.out_split) {
        constexpr int NU = BN / 32;                     // column groups (units) per tile
        auto issue_res_load = [&](uint32_t v) {         // lane 0: residual box of unit v -> buffer v % NBUF
          const uint32_t iv = v / NU, gv = v - iv * NU;
          const long long tv = (long long)blockIdx.x + (long long)iv * gridDim.x;
          if (tv < num_tiles) {
            const int mt = (int)tv / num_n_tiles, nt = (int)tv - mt * num_n_tiles;
            const int n0 = nt * BN + 32 * (int)gv;
            if (n0 < p.Cout && mt * BLOCK_M + q * 32 < M) {
              const uint32_t bar = res_bar(v % NBUF);
              mbar_arrive_expect_tx(bar, 4096);
              tma_load_2d(stg_base + (v % NBUF) * 4096, p.tmap_res, n0, mt * BLOCK_M + q * 32, bar);
            }
          }
        };
        if (has_res && lane == 0)
          for (int v = 0; v < D; ++v) issue_res_load((uint32_t)v);
        uint32_t u = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++i) {
          const int m_tile = t / num_n_tiles, n_tile = t - m_tile * num_n_tiles;
          const uint32_t b = i & 1;
          const uint32_t acc = tmem_base + b * 2 * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
          for (int gi = 0; gi < NU; ++gi, ++u) {
            const uint32_t buf = u % NBUF;
            const int n_base = n_tile * BN + 32 * gi;
            const bool col_ok = n_base < p.Cout && m_tile * BLOCK_M + q * 32 < M;   // box not entirely outside the tensor
            if (lane == 0) {
              bulk_wait_read<1>();                      // the store that last used buffer (u - 2) % NBUF has left shared memory
              if (has_res) issue_res_load(u + D);
            }
            if (gi == 0) {
              if (p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 512) p.dbg[13 * 512 + i] = clock64();
              mbar_wait_backoff<32>(tmem_full(b), (i >> 1) & 1);
              tc_fence_after();
              if (p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 512) p.dbg[14 * 512 + i] = clock64();
            }
            uint32_t r[32];
            if (col_ok) {
              tmem_ld32(acc + (uint32_t)(32 * gi), r);
              tmem_ld_wait();
              if (passes == 3) {
                uint32_t rc[32];
                tmem_ld32(acc + (uint32_t)(BN + 32 * gi), rc);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(fmaf(__uint_as_float(rc[c]), 1.0f / 2048.0f, __uint_as_float(r[c])));
              }
            }
            if (gi == NU - 1) {                         // last TMEM read of this tile: hand the buffer back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(tmem_empty(b));
            }
            __syncwarp();                               // lane 0 has seen the buffer free
            if (col_ok) {
              if (has_res) mbar_wait(res_bar(buf), (u / NBUF) & 1);
              uint8_t* rowp = stg_gen + buf * 4096 + lane * 128;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float4 bq = *reinterpret_cast<const float4*>(bias_s + n_base + 4 * c);
                float4 o = make_float4(__uint_as_float(r[4 * c]) + bq.x, __uint_as_float(r[4 * c + 1]) + bq.y,
                                       __uint_as_float(r[4 * c + 2]) + bq.z, __uint_as_float(r[4 * c + 3]) + bq.w);
                if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                float4* slot = reinterpret_cast<float4*>(rowp + ((c ^ swz) << 4));
                if (has_res) { const float4 rq = *slot; o.x += rq.x; o.y += rq.y; o.z += rq.z; o.w += rq.w; }
                *slot = o;
              }
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(p.tmap_out, stg_base + buf * 4096, n_base, m_tile * BLOCK_M + q * 32);
                bulk_commit();
              }
            }
          }
          if (p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 512) p.dbg[15 * 512 + i] = clock64();
        }
      } else {
        // output as two FP16 planes (hi, lo') for a TMA-fed consumer: 64 columns per unit = one 128-byte row per plane
        constexpr int NU = BN / 64;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++i) {
          const int m_tile = t / num_n_tiles, n_tile = t - m_tile * num_n_tiles;
          const uint32_t b = i & 1;
          const uint32_t acc = tmem_base + b * 2 * BN + ((uint32_t)(q * 32) << 16);
          const bool row_ok = m_tile * BLOCK_M + q * 32 + lane < M;
#pragma unroll 1
          for (int gi = 0; gi < NU; ++gi) {
            const int n_base = n_tile * BN + 64 * gi;
            const bool col_ok = n_base < p.Cout && m_tile * BLOCK_M + q * 32 < M;
            if (gi == 0) {
              if (p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 512) p.dbg[13 * 512 + i] = clock64();
              mbar_wait_backoff<32>(tmem_full(b), (i >> 1) & 1);
              tc_fence_after();
              if (p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 512) p.dbg[14 * 512 + i] = clock64();
            }
            float amax = 0.f;
            uint8_t* rowp = stg_gen + lane * 128;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t r[32], rc[32];
              if (col_ok) {
                tmem_ld32(acc + (uint32_t)(64 * gi + 32 * h), r);
                if (passes == 3) tmem_ld32(acc + (uint32_t)(BN + 64 * gi + 32 * h), rc);
              }
              if (h == 0) {                             // the TMEM loads above are in flight while lane 0 waits for the
                if (lane == 0) bulk_wait_read<0>();     // previous unit's two stores to leave the boxes
                __syncwarp();
              }
              if (col_ok) {
                tmem_ld_wait();
                if (passes == 3) {
#pragma unroll
                  for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(fmaf(__uint_as_float(rc[c]), 1.0f / 2048.0f, __uint_as_float(r[c])));
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {           // 8 outputs -> one 16-byte chunk of each plane
                  uint32_t hp[4], lp[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const int c = 8 * k + 2 * e;
                    float o0 = __uint_as_float(r[c]) + bias_s[n_base + 32 * h + c], o1 = __uint_as_float(r[c + 1]) + bias_s[n_base + 32 * h + c + 1];
                    if (p.relu) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); }
                    amax = fmaxf(amax, fmaxf(fabsf(o0), fabsf(o1)));
                    const float h0 = __uint_as_float(__float_as_uint(o0) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(o1) & 0xFFFFE000u);
                    const __half2 hh = __floats2half2_rn(h0, h1), ll = __floats2half2_rn((o0 - h0) * 2048.f, (o1 - h1) * 2048.f);
                    hp[e] = *reinterpret_cast<const uint32_t*>(&hh); lp[e] = *reinterpret_cast<const uint32_t*>(&ll);
                  }
                  const int ch = ((4 * h + k) ^ swz) << 4;
                  *reinterpret_cast<uint4*>(rowp + ch) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                  *reinterpret_cast<uint4*>(rowp + 4096 + ch) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
                }
              }
            }
            if (gi == NU - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(tmem_empty(b));
            }
            if (col_ok) {
              if (row_ok && amax > 60000.f && p.range_flag) *p.range_flag = 1;
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(p.tmap_out, stg_base, n_base, m_tile * BLOCK_M + q * 32);
                tma_store_2d(p.tmap_out_lo, stg_base + 4096, n_base, m_tile * BLOCK_M + q * 32);
                bulk_commit();
              }
            }
          }
          if (p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 512) p.dbg[15 * 512 + i] = clock64();
        }
      }
      if (lane == 0) bulk_wait_read<0>();               // shared memory must outlive the last stores' reads
    } else {
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int HW = p.Ho * p.Wo;
    uint32_t i = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++i) {
      const int m_tile = t / num_n_tiles, n_tile = t - m_tile * num_n_tiles;
      const uint32_t b = i & 1;
      // TMA-fed kernels have 6 warps -> up to 255 registers per thread: fetch the residual of the WHOLE tile (BN/32
      // column groups x 8 rows per lane) before waiting for the accumulator, so its HBM latency hides behind the main loop
      constexpr int NG = A_TMA ? BN / 32 : 1;
      float4 rqa[NG][8];
      if (A_TMA && p.residual && !p.out_nchw) {
        const int cg = lane & 7, rsub = lane >> 3;
        const int mrow0 = m_tile * BLOCK_M + q * 32 + rsub;
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
          const int n0 = n_tile * BN + 32 * gi + 4 * cg;
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) {
            const int mm = mrow0 + 4 * ii;
            rqa[gi][ii] = (n0 < p.Cout && mm < M) ? *reinterpret_cast<const float4*>(p.residual + (size_t)mm * p.out_c + n0)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      if (p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 512) p.dbg[13 * 512 + i] = clock64();
      mbar_wait_backoff(tmem_full(b), (i >> 1) & 1);
      tc_fence_after();
      if (p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 512) p.dbg[14 * 512 + i] = clock64();
      const uint32_t acc = tmem_base + b * 2 * BN + ((uint32_t)(q * 32) << 16);
      const int m = m_tile * BLOCK_M + q * 32 + lane;
#pragma unroll
      for (int col = 0; col < BN; col += 32) {
        uint32_t r[32];
        tmem_ld32(acc + (uint32_t)col, r);
        tmem_ld_wait();
        if (passes == 3) {
          uint32_t rc[32];
          tmem_ld32(acc + (uint32_t)(BN + col), rc);
          tmem_ld_wait();
          constexpr float kCorrScale = (MATH == MATH_F16) ? (1.0f / 2048.0f) : 1.0f;
#pragma unroll
          for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(fmaf(__uint_as_float(rc[c]), kCorrScale, __uint_as_float(r[c])));
        }
        if (col + 32 >= BN) {                      // last TMEM read of this tile: hand the buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty(b));
        }
        const int n_base = n_tile * BN + col;
        if (!p.out_nchw) {
          // Transpose the 32x32 block through shared memory so that global traffic is coalesced: in TMEM
          // layout a lane owns a whole row (32 lanes -> 32 different 128-byte lines per store instruction);
          // after the transpose 8 lanes cover one row's 128 contiguous bytes (4 full lines per instruction).
          float* stg = reinterpret_cast<float*>(smem_gen + S::STAGING_OFFSET) + (warp - 2) * (32 * 36);
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            *reinterpret_cast<float4*>(stg + lane * 36 + c) =
                make_float4(__uint_as_float(r[c]), __uint_as_float(r[c + 1]), __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3]));
          __syncwarp();
          const int cg = lane & 7, rsub = lane >> 3;
          const int n0 = n_base + 4 * cg;
          const bool ncol_ok = n0 < p.Cout;
          const float4 bq = ncol_ok ? __ldg(reinterpret_cast<const float4*>(p.bias + n0)) : make_float4(0.f, 0.f, 0.f, 0.f);
          const int mrow0 = m_tile * BLOCK_M + q * 32 + rsub;
          float4 rq[8];
          if (A_TMA && p.residual) {
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) rq[ii] = rqa[A_TMA ? col / 32 : 0][ii];
          } else if (p.residual) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int mm = mrow0 + 4 * i;
              rq[i] = (ncol_ok && mm < M) ? *reinterpret_cast<const float4*>(p.residual + (size_t)mm * p.out_c + n0)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int mm = mrow0 + 4 * i;
            float4 o = *reinterpret_cast<const float4*>(stg + (rsub + 4 * i) * 36 + 4 * cg);
            o.x += bq.x; o.y += bq.y; o.z += bq.z; o.w += bq.w;
            if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (p.residual) { o.x += rq[i].x; o.y += rq[i].y; o.z += rq[i].z; o.w += rq[i].w; }
            if (ncol_ok && mm < M) {
              if (p.out_split) {
                // consumer is a TMA-fed conv: write the two FP16 planes it will load (same split as the A producers)
                const float h0 = __uint_as_float(__float_as_uint(o.x) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(o.y) & 0xFFFFE000u);
                const float h2 = __uint_as_float(__float_as_uint(o.z) & 0xFFFFE000u), h3 = __uint_as_float(__float_as_uint(o.w) & 0xFFFFE000u);
                __half2 a = __floats2half2_rn(h0, h1), b2 = __floats2half2_rn(h2, h3);
                __half2 c2 = __floats2half2_rn((o.x - h0) * 2048.f, (o.y - h1) * 2048.f), d2 = __floats2half2_rn((o.z - h2) * 2048.f, (o.w - h3) * 2048.f);
                uint16_t* oh = reinterpret_cast<uint16_t*>(p.out) + (size_t)mm * p.out_c + n0;
                *reinterpret_cast<uint2*>(oh) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b2));
                *reinterpret_cast<uint2*>(oh + p.out_plane) = make_uint2(*reinterpret_cast<uint32_t*>(&c2), *reinterpret_cast<uint32_t*>(&d2));
                if (fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))) > 60000.f && p.range_flag) *p.range_flag = 1;
              } else {
                *reinterpret_cast<float4*>(p.out + (size_t)mm * p.out_c + n0) = o;
              }
            }
          }
          __syncwarp();      // staging is rewritten by the next column group
        } else if (m < M) {
          const int bb = m / HW, rem = m - bb * HW;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int n = n_base + c;
            if (n < p.Cout) {
              float o = __uint_as_float(r[c]) + __ldg(p.bias + n);
              if (p.relu) o = fmaxf(o, 0.f);
              p.out[((size_t)bb * p.Cout + n) * HW + rem] = o;   // lanes = consecutive pixels of plane n: coalesced
            }
          }
        }
      }
      if (p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0 && i < 512) p.dbg[15 * 512 + i] = clock64();
    }
    }
  } else if (RAW && warp == 6 + NUM_PRODUCER_WARPS) {
    // ===================== raw activation loader (plan 3): FP32 [pixels, Cin] -> 16 KB boxes, runs RAW_BOXES ahead =====================
    if (lane == 0) {
      uint32_t x = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_tile = t / num_n_tiles;
        for (int j = 0; j < nchunks; ++j)
          for (int h = 0; h < 2; ++h, ++x) {
            const uint32_t slot = x % RAW_BOXES;
            mbar_wait(raw_empty(slot), ((x / RAW_BOXES) & 1) ^ 1);
            mbar_arrive_expect_tx(raw_full(slot), RAW_BOX_BYTES);
            const uint32_t dst = smem_base + S::RAW_OFFSET + slot * RAW_BOX_BYTES;
            if (p.stem_raw) {
              // 7x7/2 stem: the tile is 128 consecutive pixels of ONE output row (Wo % 128 == 0); box h of chunk j = kernel row
              // ky = 2j + h: for output pixel ox the 32 floats starting at input pixel 2 ox - 3 of input row 2 oy + ky - 3, i.e. at
              // pixel 2 ox of row 2 oy + ky of the zero-bordered copy (dimension 1 of the map strides by 2 pixels = 32 bytes)
              const int m0 = m_tile * BLOCK_M, hw = p.Ho * p.Wo;
              const int b0 = m0 / hw, rem = m0 - b0 * hw, oy = rem / p.Wo, ox0 = rem - oy * p.Wo;
              tma_load_4d(dst, p.tmap_raw, 0, ox0, 2 * oy + 2 * j + h, b0, raw_full(slot));
            } else {
              tma_load_2d(dst, p.tmap_raw, 64 * j + 32 * h, m_tile * BLOCK_M, raw_full(slot));
            }
          }
      }
    }
  } else if (RAW) {
    // ===================== A producers (plan 3): shared -> shared transform =====================
    // Thread (row r, 16-byte operand slot g8) turns K elements [8 g8, 8 g8 + 8) of its four rows — two 16-byte chunks of
    // raw box g8 / 4 — into one 16-byte chunk of the hi plane and one of the lo' plane.  No global loads, no address
    // math per tile; the only waits are the two mbarriers (raw box full, operand stage free).
    const int pt = threadIdx.x - 192;           // 0..255
    const int g8 = pt & 7, r0 = pt >> 3;
    const int hb = g8 >> 2, c2 = 2 * (g8 & 3);
    const uint8_t* raw_gen = smem_gen + S::RAW_OFFSET;
    uint32_t g = 0;
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_base = (t / num_n_tiles) * BLOCK_M;
      for (int j = 0; j < nchunks; ++j, ++g) {
        const uint32_t x0 = 2 * g, slot0 = x0 % RAW_BOXES;           // RAW_BOXES is even: the pair never wraps
        mbar_wait(raw_full(slot0), (x0 / RAW_BOXES) & 1);
        mbar_wait(raw_full(slot0 + 1), (x0 / RAW_BOXES) & 1);
        float4 cur[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = r0 + 32 * i;
          const uint8_t* rowp = raw_gen + (slot0 + hb) * RAW_BOX_BYTES + r * 128;
          cur[2 * i] = *reinterpret_cast<const float4*>(rowp + ((c2 ^ (r & 7)) << 4));
          cur[2 * i + 1] = *reinterpret_cast<const float4*>(rowp + (((c2 + 1) ^ (r & 7)) << 4));
        }
        if (PRE) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.pre_scale + KC * j + 8 * g8) + u);
            const float4 sh = __ldg(reinterpret_cast<const float4*>(p.pre_shift + KC * j + 8 * g8) + u);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float4& x = cur[2 * i + u];
              x.x = fmaxf(fmaf(x.x, sc.x, sh.x), 0.f); x.y = fmaxf(fmaf(x.y, sc.y, sh.y), 0.f);
              x.z = fmaxf(fmaf(x.z, sc.z, sh.z), 0.f); x.w = fmaxf(fmaf(x.w, sc.w, sh.w), 0.f);
            }
          }
        }
        // the boxes are refilled by the async proxy (TMA): order this thread's generic-proxy reads before that write,
        // then release (without the fence the ld.shared of a lane may still be in flight when the next box lands)
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { mbar_arrive(raw_empty(slot0)); mbar_arrive(raw_empty(slot0 + 1)); }
        mbar_wait(empty(stage), phase ^ 1);
        if (p.dbg && blockIdx.x == 0 && pt == 0 && g < 512) p.dbg[0 * 512 + g] = clock64();
        uint8_t* a_hi = smem_gen + stage * S::STAGE_BYTES;
        uint8_t* a_lo = a_hi + A_TILE_BYTES;
        float amax = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = r0 + 32 * i;
          const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((g8 ^ (r & 7)) << 4);
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float4 x = cur[2 * i + u];
            const float h0 = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
            const float h2 = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u), h3 = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
            if (m_base + r < M) amax = fmaxf(amax, fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))));
            __half2 a = __floats2half2_rn(h0, h1), b = __floats2half2_rn(h2, h3);
            __half2 c = __floats2half2_rn((x.x - h0) * 2048.f, (x.y - h1) * 2048.f), d = __floats2half2_rn((x.z - h2) * 2048.f, (x.w - h3) * 2048.f);
            hw[2 * u] = *reinterpret_cast<uint32_t*>(&a); hw[2 * u + 1] = *reinterpret_cast<uint32_t*>(&b);
            lw[2 * u] = *reinterpret_cast<uint32_t*>(&c); lw[2 * u + 1] = *reinterpret_cast<uint32_t*>(&d);
          }
          *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          if (passes == 3) *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        // rows past the last pixel of the batch (M) may hold another tensor's data (activation allocations are shared between tensors with
        // disjoint live ranges) or TMA's zero fill past the end of the buffer: they are kept out of the range check; their products land in
        // accumulator rows that are never stored inside the batch
        if (amax > 60000.f && p.range_flag) *p.range_flag = 1;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_a(stage));
        if (p.dbg && blockIdx.x == 0 && pt == 0 && g < 512) p.dbg[1 * 512 + g] = clock64();
        if (p.dbg && blockIdx.x == 0 && lane == 0 && g < 512) p.dbg[(5 + warp - 6) * 512 + g] = clock64();
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (!A_TMA) {
    // ===================== A producers =====================
    const int pt = threadIdx.x - 192;           // 0..255
    const int g8 = pt & 7;                      // float4 group inside the 128-byte row
    const int r0 = pt >> 3;                     // rows r0 + 32*i
    const int cpc = p.Cin / KC, cpr = p.chunks_per_row * 32 / KC;     // sub-chunks per tap / per stem kernel row
    // RGB-only stem in FP16 chunks: a kernel row is ONE 32-float chunk (chunks_per_row == 1), so a 64-wide chunk is TWO kernel
    // rows: slots g8 0..3 = row 2j, pixels 2 g8 and 2 g8 + 1; slots 4..7 = row 2j + 1 (row 7 does not exist: mask bit 7 is never set)
    const bool stem2 = MODE == CONV_STEM7 && KC == 64 && p.chunks_per_row == 1;
    auto advance = [&](LoadCursor& c) {
      if (++c.j == nchunks) {      // next tile: decode out of line into a scratch copy so that `c` itself stays in registers
        LoadCursor tmp;
        tmp.t = c.t + gridDim.x;
        cursor_set_tile<MODE>(tmp, p, M, num_tiles, num_n_tiles, r0);
        c = tmp;
        return;
      }
      if (MODE == CONV_1x1) { c.off += KC; }
      else if (MODE == CONV_3x3) {
        if (++c.cc == cpc) { c.cc = 0; ++c.tap; c.off = (ptrdiff_t)((c.tap / 3 - 1) * p.W + (c.tap % 3 - 1)) * p.Cin; }
        else c.off += KC;
      } else if (stem2) {
        c.tap += 2; c.off = (ptrdiff_t)c.tap * p.W * p.Cin;
      } else {
        if (++c.cc == cpr) { c.cc = 0; ++c.tap; c.off = (ptrdiff_t)c.tap * p.W * p.Cin; }
        else c.off += KC;
      }
    };
    LoadCursor L;
    {
      LoadCursor tmp;
      tmp.t = blockIdx.x;
      cursor_set_tile<MODE>(tmp, p, M, num_tiles, num_n_tiles, r0);
      L = tmp;
    }
    int st = blockIdx.x, sj = 0;               // store cursor: tile / chunk
    constexpr int PF = (MATH == MATH_F16) ? 2 : P_PREFETCH;       // register prefetch depth (same bytes in flight)
    float4 v[PF][4 * NV];
    uint32_t vbits[PF];
    auto load_chunk = [&](const LoadCursor& c, float4 (&dst)[4 * NV], uint32_t& bits) {
      // this thread's K elements of the chunk: [ (KC/8) * g8, +KC/8 ) -> one 16-byte slot of the 128-byte operand row
      uint32_t pixbit = 0;                      // stem only: which input pixel of the kernel row these elements belong to
      int srow = 0;                             // stem2 only: 0 / 1 = first / second kernel row of the chunk
      ptrdiff_t toff = (KC / 8) * g8;           // this thread's float offset inside the chunk's source row
      if (MODE == CONV_STEM7) {
        int x = KC * c.cc + (KC / 8) * g8;
        if (stem2) { srow = g8 >> 2; x = 8 * (g8 & 3); toff = (ptrdiff_t)srow * p.W * p.Cin + x; }
        const int pix = (p.Cin == 4) ? (x >> 2) : ((x * 1366) >> 16);   // x / Cin for Cin in {4, 48}
        pixbit = pix < 7 ? (1u << (8 + pix)) : 0u;
      }
      bits = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bool ok;
        if (MODE == CONV_1x1) ok = c.mask[i] & 1u;
        else if (MODE == CONV_3x3) ok = (c.mask[i] >> c.tap) & 1u;
        else ok = ((c.mask[i] >> (c.tap + srow)) & 1u) && (c.mask[i] & pixbit);
        const float* __restrict__ q = c.base[i] + c.off + toff;
#pragma unroll
        for (int u = 0; u < NV; ++u) {
          dst[i * NV + u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (MODE == CONV_STEM7 && NV == 2 && u == 1 && p.Cin == 4) {
            // 4-float pixels: the second float4 is the NEXT pixel of the kernel row
            const uint32_t pb2 = (pixbit << 1) & 0x7F00u;
            if (((c.mask[i] >> (c.tap + srow)) & 1u) && (c.mask[i] & pb2)) dst[i * NV + u] = __ldg(reinterpret_cast<const float4*>(q) + u);
          } else if (ok) {
            dst[i * NV + u] = __ldg(reinterpret_cast<const float4*>(q) + u);
          }
        }
        if (ok) bits |= 1u << i;
      }
    };
#pragma unroll
    for (int qq = 0; qq < PF; ++qq) {
      vbits[qq] = 0;
      if (L.t < num_tiles) { load_chunk(L, v[qq], vbits[qq]); advance(L); }
    }
    uint32_t g = 0;
    int stage = 0;
    uint32_t phase = 0;
    auto produce = [&](float4 (&buf)[4 * NV], uint32_t& bits) {
      float4 cur[4 * NV];
#pragma unroll
      for (int i = 0; i < 4 * NV; ++i) cur[i] = buf[i];
      if (PRE) {
#pragma unroll
        for (int u = 0; u < NV; ++u) {
          const float4 sc = __ldg(reinterpret_cast<const float4*>(p.pre_scale + KC * sj + (KC / 8) * g8) + u);
          const float4 sh = __ldg(reinterpret_cast<const float4*>(p.pre_shift + KC * sj + (KC / 8) * g8) + u);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if ((bits >> i) & 1u) {   // (bits of the chunk being stored)
              float4& x = cur[i * NV + u];
              x.x = fmaxf(fmaf(x.x, sc.x, sh.x), 0.f); x.y = fmaxf(fmaf(x.y, sc.y, sh.y), 0.f);
              x.z = fmaxf(fmaf(x.z, sc.z, sh.z), 0.f); x.w = fmaxf(fmaf(x.w, sc.w, sh.w), 0.f);
            }
          }
        }
      }
      const uint32_t bits_now = bits;
      mbar_wait(empty(stage), phase ^ 1);
      if (p.dbg && blockIdx.x == 0 && pt == 0 && g < 512) p.dbg[0 * 512 + g] = clock64();
      uint8_t* a_hi = smem_gen + stage * S::STAGE_BYTES;
      uint8_t* a_lo = a_hi + A_TILE_BYTES;
      float amax = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + 32 * i;
        const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((g8 ^ (r & 7)) << 4);
        if (MATH == MATH_F16) {
          // hi = x with the 13 low mantissa bits cleared (exactly an FP16 number inside FP16's normal range),
          // lo' = (x - hi) * 2^11 (exact in FP32, rounded once to FP16): x = hi + 2^-11 lo' to ~22 bits.
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float4 x = cur[i * NV + (NV == 2 ? u : 0)];
            const float h0 = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
            const float h2 = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u), h3 = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))));
            __half2 a = __floats2half2_rn(h0, h1), b = __floats2half2_rn(h2, h3);
            __half2 c = __floats2half2_rn((x.x - h0) * 2048.f, (x.y - h1) * 2048.f), d = __floats2half2_rn((x.z - h2) * 2048.f, (x.w - h3) * 2048.f);
            hw[2 * u] = *reinterpret_cast<uint32_t*>(&a); hw[2 * u + 1] = *reinterpret_cast<uint32_t*>(&b);
            lw[2 * u] = *reinterpret_cast<uint32_t*>(&c); lw[2 * u + 1] = *reinterpret_cast<uint32_t*>(&d);
          }
          *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          if (passes == 3) *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        } else {
          // TF32 split without conversion instructions: the tensor core reads only the top 19 bits of an FP32
          // word, so hi = x with the low 13 mantissa bits cleared and lo = x - hi (exact) are full-rate ALU ops.
          const float4 x = cur[i * NV];
          float4 hi;
          hi.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); hi.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
          hi.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); hi.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
          *reinterpret_cast<float4*>(a_hi + off) = hi;
          if (passes == 3) {
            float4 lo;
            lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
            *reinterpret_cast<float4*>(a_lo + off) = lo;
          }
        }
      }
      if (MATH == MATH_F16 && amax > 60000.f && p.range_flag) *p.range_flag = 1;   // FP16 operand range exceeded: caller must use tf32x3
      // issue the global loads of a later chunk while this chunk's shared-memory stores drain: the proxy fence
      // below waits for them, and its cost grows with the number of stores still in flight
      (void)bits_now;
      if (L.t < num_tiles) { load_chunk(L, buf, bits); advance(L); }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_a(stage));
      if (p.dbg && blockIdx.x == 0 && pt == 0 && g < 512) p.dbg[1 * 512 + g] = clock64();
      if (p.dbg && blockIdx.x == 0 && lane == 0 && g < 512) p.dbg[(5 + warp - 6) * 512 + g] = clock64();
      ++g;
      if (++stage == NS) { stage = 0; phase ^= 1; }
      if (++sj == nchunks) { sj = 0; st += gridDim.x; }
    };
    static_assert(P_PREFETCH == 4, "producer loop is unrolled for four (TF32) / two (FP16) register buffers");
    while (st < num_tiles) {
      produce(v[0], vbits[0]);
      if (st < num_tiles) produce(v[1], vbits[1]);
      if (PF == 4) {
        if (st < num_tiles) produce(v[PF - 2], vbits[PF - 2]);
        if (st < num_tiles) produce(v[PF - 1], vbits[PF - 1]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.trace && threadIdx.x == 0 && blockIdx.x == 0 && trace_slot < 8192) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[1 + 4 * trace_slot + 1] = t;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 4 * BN);
  }
}

template <int BN, int MODE, bool PRE, int MATH, bool A_TMA, int EPI>
int launch_persistent_inst(suo_ctx* ctx, const ConvParams& p, int passes, cudaStream_t s) {
  static bool configured[64] = {};
  int num_sms = 148;
  if (first_use_on_device(configured, &num_sms)) SUO_CUDA_TRY(ctx, cudaFuncSetAttribute(conv_tc_persistent_kernel<BN, MODE, PRE, MATH, A_TMA, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, PSmem<BN, EPI>::TOTAL));
  const int M = p.B * p.Ho * p.Wo;
  const int mt = (M + BLOCK_M - 1) / BLOCK_M, nt = p.Cout_pad / BN;
  const int grid = std::min(mt * nt, ctx->opt_grid_cap > 0 ? std::min(num_sms, ctx->opt_grid_cap) : num_sms);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(A_TMA ? 192 : (EPI == 3 ? P_NUM_THREADS + 32 : P_NUM_THREADS));
  cfg.dynamicSmemBytes = PSmem<BN, EPI>::TOTAL;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (ctx->opt_pdl && !ctx->opt_multistream) ? 1 : 0;      // level streams join through events: keep those launches plain
  SUO_CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, conv_tc_persistent_kernel<BN, MODE, PRE, MATH, A_TMA, EPI>, p, passes, mt, nt));
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

template <int BN, int MATH, int EPI>
int launch_persistent_epi(suo_ctx* ctx, const ConvParams& p, int passes, cudaStream_t s) {
  if (MATH == MATH_F16 && p.in_split) {
    if (p.pre_scale || p.mode == CONV_STEM7 || p.Cin % 64) { ctx->set_error("conv_tc: TMA-fed A needs a plain 1x1/3x3 conv with Cin % 64 == 0", __FILE__, __LINE__); return SUO_E_INVALID; }
    if (p.mode == CONV_3x3) {
      if (BN == 128 && conv_halo_eligible(p, passes)) return launch_conv_halo(ctx, p, s);
      if (BN == 128 && conv_pair_eligible(p, passes)) return launch_conv_pair(ctx, p, s);
      return launch_persistent_inst<BN, CONV_3x3, false, MATH_F16, true, (EPI >= 2 ? 1 : EPI)>(ctx, p, passes, s);
    }
    return launch_persistent_inst<BN, CONV_1x1, false, MATH_F16, true, (EPI == 3 ? 1 : EPI)>(ctx, p, passes, s);
  }
  if (MATH == MATH_F16 && EPI == 3) {          // FP32 activations by TMA + shared->shared transform (1x1 only)
    if (p.pre_scale) return launch_persistent_inst<BN, CONV_1x1, true, MATH_F16, false, (EPI == 3 ? 3 : 1)>(ctx, p, passes, s);
    return launch_persistent_inst<BN, CONV_1x1, false, MATH_F16, false, (EPI == 3 ? 3 : 1)>(ctx, p, passes, s);
  }
  constexpr int E = (EPI == 2 || EPI == 3) ? 1 : EPI;      // plans 2 and 3 exist for 1x1 kernels only
  if (p.mode == CONV_3x3) return launch_persistent_inst<BN, CONV_3x3, false, MATH, false, E>(ctx, p, passes, s);
  if (p.mode == CONV_STEM7) return launch_persistent_inst<BN, CONV_STEM7, false, MATH, false, E>(ctx, p, passes, s);
  if (p.pre_scale) return launch_persistent_inst<BN, CONV_1x1, true, MATH, false, E>(ctx, p, passes, s);
  return launch_persistent_inst<BN, CONV_1x1, false, MATH, false, E>(ctx, p, passes, s);
}

template <int BN, int MATH>
int launch_persistent(suo_ctx* ctx, const ConvParams& p, int passes, cudaStream_t s) {
  if (MATH == MATH_F16) {
    // TMA-store epilogue: needs the output (and skip) tensor maps, NHWC, <= 256 channels of bias in shared memory
    const bool epi_tma = p.epi_tma && !p.out_nchw && p.Cout_pad <= 256 && !(p.residual && p.out_split);
    if (epi_tma) {
      if (p.residual && p.in_split && p.mode == CONV_1x1 && p.K <= 256) return launch_persistent_epi<BN, MATH_F16, 2>(ctx, p, passes, s);
      if (p.raw_tma && !p.in_split && p.mode == CONV_1x1 && p.Cin % 64 == 0) return launch_persistent_epi<BN, MATH_F16, 3>(ctx, p, passes, s);
      if (p.stem_raw && p.mode == CONV_STEM7 && !p.residual && passes == 3 && p.K == 256 && p.Wo % BLOCK_M == 0)   // TMA-fed stem on plan 3
        return launch_persistent_inst<BN, CONV_1x1, false, MATH_F16, false, 3>(ctx, p, passes, s);
      return launch_persistent_epi<BN, MATH_F16, 1>(ctx, p, passes, s);
    }
  }
  return launch_persistent_epi<BN, MATH, 0>(ctx, p, passes, s);
}

template <int BN>
int launch_bn(suo_ctx* ctx, const ConvParams& p, int passes, cudaStream_t s) {
  static bool configured[64] = {};
  if (first_use_on_device(configured, nullptr)) SUO_CUDA_TRY(ctx, cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN>::TOTAL));
  const int M = p.B * p.Ho * p.Wo;
  dim3 grid((M + BLOCK_M - 1) / BLOCK_M, p.Cout_pad / BN);
  conv_tc_kernel<BN><<<grid, NUM_THREADS, Smem<BN>::TOTAL, s>>>(p, passes);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

inline float host_tf32_rna(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return x;
  u = (u + 0x1000u) & 0xFFFFE000u;     // round-to-nearest, ties away (cvt.rna.tf32.f32)
  float r;
  memcpy(&r, &u, 4);
  return r;
}

}  // namespace

int conv_tc_block_n(int Cout_pad) { return Cout_pad % 128 == 0 ? 128 : 64; }

size_t conv_tc_packed_floats(int Cout_pad, int K) { return (size_t)Cout_pad * K * 2; }

// w: [Cout_pad][K] row-major -> per (n_tile, chunk): hi image then lo image, each BN rows x 128 B
// in the 128B-swizzled K-major layout the UMMA descriptor above describes.
void conv_tc_pack_weights(const float* w, int Cout_pad, int K, float* dst) {
  const int BN = conv_tc_block_n(Cout_pad);
  const int nchunks = K / BLOCK_K;
  for (int nt = 0; nt < Cout_pad / BN; ++nt)
    for (int j = 0; j < nchunks; ++j) {
      float* img = dst + ((size_t)nt * nchunks + j) * 2 * (BN * BLOCK_K);
      for (int r = 0; r < BN; ++r)
        for (int g = 0; g < 8; ++g)
          for (int e = 0; e < 4; ++e) {
            const float x = w[(size_t)(nt * BN + r) * K + 32 * j + 4 * g + e];
            const float hi = host_tf32_rna(x);
            const float lo = host_tf32_rna(x - hi);
            const int off = (r >> 3) * 256 + (r & 7) * 32 + ((g ^ (r & 7)) << 2) + e;
            img[off] = hi;
            img[BN * BLOCK_K + off] = lo;
          }
    }
}

namespace {
inline uint16_t host_f32_to_f16_rn(float f) {       // IEEE round-to-nearest-even, handles subnormals; no NaN inputs expected
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  const int32_t e = (int32_t)((x >> 23) & 0xFF) - 127 + 15;
  uint32_t m = x & 0x7FFFFFu;
  if (e >= 31) return (uint16_t)(sign | 0x7C00u);                    // overflow -> inf
  if (e <= 0) {
    if (e < -10) return (uint16_t)sign;
    m |= 0x800000u;
    const int shift = 14 - e;
    uint32_t h = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1))) ++h;
    return (uint16_t)(sign | h);
  }
  uint32_t h = ((uint32_t)e << 10) | (m >> 13);
  const uint32_t rem = m & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;
  return (uint16_t)(sign | h);
}
inline float host_f16_to_f32(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t e = (h >> 10) & 0x1F, m = h & 0x3FFu, x;
  if (e == 0) {
    if (m == 0) x = sign;
    else { int sh = 0; while (!(m & 0x400u)) { m <<= 1; ++sh; } m &= 0x3FFu; x = sign | ((uint32_t)(127 - 15 - sh + 1) << 23) | (m << 13); }
  } else if (e == 31) x = sign | 0x7F800000u | (m << 13);
  else x = sign | ((e - 15 + 127) << 23) | (m << 13);
  float f;
  memcpy(&f, &x, 4);
  return f;
}
}  // namespace

size_t conv_tc_packed16_halfs(int Cout_pad, int K) { return (size_t)Cout_pad * K * 2; }

// host mirror of the device split: x -> (hi, lo') with x ~= hi + 2^-11 lo'
void conv_tc_host_split_f16(const float* x, size_t n, uint16_t* hi, uint16_t* lo) {
  for (size_t i = 0; i < n; ++i) {
    uint32_t u;
    memcpy(&u, &x[i], 4);
    u &= 0xFFFFE000u;
    float ht;
    memcpy(&ht, &u, 4);
    hi[i] = host_f32_to_f16_rn(ht);
    lo[i] = host_f32_to_f16_rn((x[i] - host_f16_to_f32(hi[i])) * 2048.0f);
  }
}
void conv_tc_host_join_f16(const uint16_t* hi, const uint16_t* lo, size_t n, float* x) {
  for (size_t i = 0; i < n; ++i) x[i] = host_f16_to_f32(hi[i]) + host_f16_to_f32(lo[i]) * (1.0f / 2048.0f);
}

// FP16x3 weight images: per (n_tile, 64-element chunk): hi image then lo' image, BN rows x 128 B, 128B swizzle.
// w = hi + 2^-11 lo'  with hi = fp16(w with the 13 low mantissa bits cleared), lo' = fp16((w - hi) * 2^11).
void conv_tc_pack_weights_f16(const float* w, int Cout_pad, int K, uint16_t* dst) {
  const int BN = conv_tc_block_n(Cout_pad);
  const int nchunks = K / 64;
  for (int nt = 0; nt < Cout_pad / BN; ++nt)
    for (int j = 0; j < nchunks; ++j) {
      uint16_t* img = dst + ((size_t)nt * nchunks + j) * 2 * (BN * 64);
      for (int r = 0; r < BN; ++r)
        for (int g = 0; g < 8; ++g)
          for (int e = 0; e < 8; ++e) {
            const float x = w[(size_t)(nt * BN + r) * K + 64 * j + 8 * g + e];
            uint32_t u;
            memcpy(&u, &x, 4);
            u &= 0xFFFFE000u;
            float ht;
            memcpy(&ht, &u, 4);
            const uint16_t hi = host_f32_to_f16_rn(ht);
            const float hif = host_f16_to_f32(hi);
            const uint16_t lo = host_f32_to_f16_rn((x - hif) * 2048.0f);
            const int off = (r >> 3) * 512 + (r & 7) * 64 + ((g ^ (r & 7)) << 3) + e;     // in halfs
            img[off] = hi;
            img[BN * 64 + off] = lo;
          }
    }
}

int launch_conv_tc(suo_ctx* ctx, const ConvParams& p, int tf32_passes, cudaStream_t s) {
  if (p.K % BLOCK_K || p.Cin % 4 || p.Cout_pad % 64 || (!p.out_nchw && (p.Cout % 4 || p.out_c % 4)) ||
      (p.mode != CONV_STEM7 && p.Cin % 32)) {
    ctx->set_error("conv_tc: unsupported shape", __FILE__, __LINE__);
    return SUO_E_INVALID;
  }
  const int passes = tf32_passes == 1 ? 1 : 3;
  if (ctx->opt_persistent) {
    if (p.math == MATH_F16) {
      if (p.K % 64 || !p.w_packed16) { ctx->set_error("conv_tc fp16x3: K % 64 != 0 or weights not packed", __FILE__, __LINE__); return SUO_E_INVALID; }
      ConvParams q = p;
      q.w_packed = reinterpret_cast<const float*>(p.w_packed16);
      if (conv_tc_block_n(p.Cout_pad) == 128) return launch_persistent<128, MATH_F16>(ctx, q, passes, s);
      return launch_persistent<64, MATH_F16>(ctx, q, passes, s);
    }
    if (conv_tc_block_n(p.Cout_pad) == 128) return launch_persistent<128, MATH_TF32>(ctx, p, passes, s);
    return launch_persistent<64, MATH_TF32>(ctx, p, passes, s);
  }
  if (conv_tc_block_n(p.Cout_pad) == 128) return launch_bn<128>(ctx, p, passes, s);
  return launch_bn<64>(ctx, p, passes, s);
}
