// A-halo 3x3 convolution as a CTA pair for sm_100a: the bottleneck's conv2 (3x3, pad 1, Cin -> 128, folded BN + ReLU; reference
// lib/models/layers/Residual.py:13-15,28-31) with the activations fetched once per COLUMN SHIFT instead of once per tap.
//
// Why: the other 3x3 kernels (conv_tc.cu single CTA, conv_pair.cu) bring a fresh 128-pixel A tile per tap and K chunk — 18 x 2 TMA boxes
// of 128 rows x 128 B per tile — and all of them run at the same ~1240 cycles per chunk whatever their MMA work ("open
// question"): the A delivery is the bound.  Here, per tile and 64-channel chunk, three boxes {64 ch, W px, rows + 2, 1} at x offsets
// dx = -1, 0, +1 (TMA zero-fills the image border in x and y = the conv's padding) land three "variants" of (rows + 2) x W pixels.
// Inside a variant the three dy taps are the SAME shared-memory image at start offsets (dy + 1) * W rows: output pixel m = r W + x reads
// variant row (r + dy + 1) W + x = m + (dy + 1) W (the variant is already shifted in x), the offset is a multiple of 8 rows = 1 KB for
// W >= 8, so the 128B swizzle phase is unchanged and the UMMA descriptor only moves its start address.  A rows per tile: 6 (rows + 2) W
// instead of 18 x 128 (1.5x fewer at W = 64, 2x at 32, 2.4x at 16); each variant feeds 3 x 12 MMAs.
//
// Layout / roles as conv_pair.cu (cluster of 2, tcgen05.mma.cta_group::2 of M = 256, each CTA its own 128 pixels and half of the weight
// rows, every load completes on the leader's barrier, multicast commits, TMA-store epilogue) with two 64 KB variant buffers, a 4-deep
// ring of 16 KB weight halves and separate producer warps for the two (warp 0: variants, warp 6: weights).
// K order: (chunk, dx, dy) instead of (dy, dx, chunk): the same products, accumulated in another order — results agree with the other
// kernels to FP32 rounding, not bit for bit.
#include <cuda_fp16.h>
#include <algorithm>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int HM = 128;                                // pixels per CTA (256 per pair tile)
constexpr int HN = 128;                                // output channels
constexpr int H_VPLANE = 32768;                        // one plane of a variant: up to 256 rows x 128 B
constexpr int H_VBUF = 2 * H_VPLANE;                   // hi | lo'
constexpr int H_BHALF = 8192;                          // this CTA's 64 rows of one weight image
constexpr int H_BSTAGE = 2 * H_BHALF;                  // hi | lo'
constexpr int H_NB = 4;
constexpr int H_B_OFF = 2 * H_VBUF;                    // 128 KB
constexpr int H_STG_OFF = H_B_OFF + H_NB * H_BSTAGE;   // 192 KB
constexpr int H_BAR_OFF = H_STG_OFF + 4 * 2 * 4096;    // 224 KB
constexpr int H_BIAS_OFF = H_BAR_OFF + 512;
constexpr int H_TOTAL = H_BIAS_OFF + HN * 4 + 1024;
constexpr int H_THREADS = 224;
static_assert(H_TOTAL <= 232448, "exceeds the 227 KB a CTA may use");

__global__ void __launch_bounds__(H_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ ConvParams p, const int num_m_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // identical in both CTAs of the pair
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + H_BAR_OFF;
  auto afull = [&](int v) { return bar_base + 8 * v; };                      // leader: both CTAs' variant v has landed
  auto aempty = [&](int v) { return bar_base + 8 * (2 + v); };               // both: the 36 MMAs reading variant buffer v are done
  auto bfull = [&](int s) { return bar_base + 8 * (4 + s); };                // leader: both CTAs' weight halves of stage s have landed
  auto bempty = [&](int s) { return bar_base + 8 * (4 + H_NB + s); };        // both
  auto tmem_full = [&](int b) { return bar_base + 8 * (4 + 2 * H_NB + b); };
  auto tmem_empty = [&](int b) { return bar_base + 8 * (6 + 2 * H_NB + b); };
  const uint32_t tmem_slot = bar_base + 8 * (8 + 2 * H_NB);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + H_BAR_OFF + 8 * (8 + 2 * H_NB));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = (int)blockIdx.x >> 1, num_clusters = (int)gridDim.x >> 1;
  const int num_pair_tiles = (num_m_tiles + 1) >> 1;
  const int M = p.B * p.Ho * p.Wo;
  const int cpc = p.Cin / 64;                           // 64-channel chunks per tap
  const int vrows = (HM / p.W + 2) * p.W;               // rows of a variant
  const uint32_t vbytes = (uint32_t)vrows * 128u;       // bytes of one plane of a variant
  const uint32_t dy_step = (uint32_t)p.W * 128u;        // shared-memory distance of consecutive dy taps

  if (threadIdx.x == 0) {
    for (int v = 0; v < 2; ++v) { mbar_init(afull(v), 1); mbar_init(aempty(v), 1); }
    for (int s = 0; s < H_NB; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== variant producer (both CTAs) =====================
    if (lane == 0) {
      const int HWi = p.H * p.W;
      uint32_t g = 0;
      for (int pt = cluster_id; pt < num_pair_tiles; pt += num_clusters) {
        const int m0 = (2 * pt + (int)rank) * HM;                            // beyond M: coordinates outside the tensor -> zero fill
        const int b0 = m0 / HWi, rem = m0 - b0 * HWi, y0 = rem / p.W;
        for (int cc = 0; cc < cpc; ++cc)
          for (int dx = -1; dx <= 1; ++dx, ++g) {
            const int v = g & 1;
            mbar_wait(aempty(v), ((g >> 1) & 1) ^ 1);
            const uint32_t dst = smem_base + v * H_VBUF;
            const uint32_t lbar = mapa_shared(afull(v), 0);
            if (rank == 0) mbar_arrive_expect_tx(afull(v), 4 * vbytes);      // two planes of this CTA + two of the peer
            tma_load_4d_2sm(dst, p.tmap_hhi, 64 * cc, dx, y0 - 1, b0, lbar);
            tma_load_4d_2sm(dst + H_VPLANE, p.tmap_hlo, 64 * cc, dx, y0 - 1, b0, lbar);
          }
      }
    }
  } else if (warp == 6) {
    // ===================== weight producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t g = 0;
      for (int pt = cluster_id; pt < num_pair_tiles; pt += num_clusters)
        for (int cc = 0; cc < cpc; ++cc)
          for (int dx = 0; dx < 3; ++dx)
            for (int dy = 0; dy < 3; ++dy, ++g) {
              const int s = g % H_NB;
              mbar_wait(bempty(s), ((g / H_NB) & 1) ^ 1);
              const uint32_t dst = smem_base + H_B_OFF + s * H_BSTAGE;
              const uint32_t lbar = mapa_shared(bfull(s), 0);
              const int j = (3 * dy + dx) * cpc + cc;                        // chunk index in the packed (tap, chunk) order
              if (rank == 0) mbar_arrive_expect_tx(bfull(s), 2 * H_BSTAGE);
              tma_load_2d_2sm(dst, p.tmap_w, 0, (2 * j) * HN + 64 * (int)rank, lbar);
              tma_load_2d_2sm(dst + H_BHALF, p.tmap_w, 0, (2 * j + 1) * HN + 64 * (int)rank, lbar);
            }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(2 * HM, HN);
      uint32_t ga = 0, gb = 0, i = 0;
      for (int pt = cluster_id; pt < num_pair_tiles; pt += num_clusters, ++i) {
        const uint32_t b = i & 1;
        mbar_wait_cluster(tmem_empty(b), ((i >> 1) & 1) ^ 1);                // both epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + b * 2 * HN;
        for (int cd = 0; cd < 3 * cpc; ++cd, ++ga) {                         // (chunk, dx)
          const int v = ga & 1;
          mbar_wait(afull(v), (ga >> 1) & 1);
          for (int dy = 0; dy < 3; ++dy, ++gb) {
            const int s = gb % H_NB;
            mbar_wait(bfull(s), (gb / H_NB) & 1);
            tc_fence_after();
            if (lane == 0) {
              const uint32_t a_hi = smem_base + v * H_VBUF + dy * dy_step, a_lo = a_hi + H_VPLANE;
              const uint32_t b_hi = smem_base + H_B_OFF + s * H_BSTAGE, b_lo = b_hi + H_BHALF;
              const uint32_t first = (cd | dy) != 0;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t dah = make_sw128_desc(a_hi + kk * 32), dal = make_sw128_desc(a_lo + kk * 32);
                const uint64_t dbh = make_sw128_desc(b_hi + kk * 32), dbl = make_sw128_desc(b_lo + kk * 32);
                umma2_f16(acc, dah, dbh, idesc, (first | kk) != 0);          // main
                umma2_f16(acc + HN, dah, dbl, idesc, (first | kk) != 0);     // correction: hi x lo'
                umma2_f16(acc + HN, dal, dbh, idesc, 1u);                    //             lo' x hi
              }
              umma2_commit(bempty(s));
              if (dy == 2) umma2_commit(aempty(v));
              if (dy == 2 && cd == 3 * cpc - 1) umma2_commit(tmem_full(b));
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs): TMEM -> bias, ReLU, FP16 hi/lo' split -> staging boxes -> TMA store =====================
    const int ew = warp - 2, q = warp & 3;
    float* bias_s = reinterpret_cast<float*>(smem_gen + H_BIAS_OFF);
    for (int c = threadIdx.x - 64; c < HN; c += 128) bias_s[c] = __ldg(p.bias + c);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const uint32_t stg_base = smem_base + H_STG_OFF + ew * 8192;
    uint8_t* stg_gen = smem_gen + H_STG_OFF + ew * 8192;
    const int swz = lane & 7;
    uint32_t i = 0;
    for (int pt = cluster_id; pt < num_pair_tiles; pt += num_clusters, ++i) {
      const int m_tile = 2 * pt + (int)rank;
      const uint32_t b = i & 1;
      const uint32_t acc = tmem_base + b * 2 * HN + ((uint32_t)(q * 32) << 16);
      const bool row_ok = m_tile * HM + q * 32 + lane < M;
      const bool box_ok = m_tile * HM + q * 32 < M;
#pragma unroll 1
      for (int gi = 0; gi < HN / 64; ++gi) {
        const int n_base = 64 * gi;
        if (gi == 0) {
          mbar_wait_backoff<32>(tmem_full(b), (i >> 1) & 1);
          tc_fence_after();
        }
        float amax = 0.f;
        uint8_t* rowp = stg_gen + lane * 128;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32], rc[32];
          if (box_ok) {
            tmem_ld32(acc + (uint32_t)(64 * gi + 32 * h), r);
            tmem_ld32(acc + (uint32_t)(HN + 64 * gi + 32 * h), rc);
          }
          if (h == 0) {                               // the TMEM loads are in flight while lane 0 waits for the
            if (lane == 0) bulk_wait_read<0>();       // previous unit's two stores to leave the boxes
            __syncwarp();
          }
          if (box_ok) {
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(fmaf(__uint_as_float(rc[c]), 1.0f / 2048.0f, __uint_as_float(r[c])));
#pragma unroll
            for (int k = 0; k < 4; ++k) {             // 8 outputs -> one 16-byte chunk of each plane
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
        if (gi == HN / 64 - 1) {                      // last TMEM read of this tile: hand the accumulator back to the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_shared(tmem_empty(b), 0));
        }
        if (box_ok) {
          if (row_ok && amax > 60000.f && p.range_flag) *p.range_flag = 1;
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(p.tmap_out, stg_base, n_base, m_tile * HM + q * 32);
            tma_store_2d(p.tmap_out_lo, stg_base + 4096, n_base, m_tile * HM + q * 32);
            bulk_commit();
          }
        }
      }
    }
    if (lane == 0) bulk_wait_read<0>();               // shared memory must outlive the last stores' reads
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                 // no CTA leaves while the peer may still signal its barriers / read its operands
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace

bool conv_halo_eligible(const ConvParams& p, int passes) {
  return p.halo && conv_pair_eligible(p, passes) && (p.W == 16 || p.W == 32 || p.W == 64) && p.H * p.W >= HM && (p.H * p.W) % HM == 0;
}

int launch_conv_halo(suo_ctx* ctx, const ConvParams& p, cudaStream_t s) {
  static bool configured[64] = {};
  int num_sms = 148;
  if (first_use_on_device(configured, &num_sms)) SUO_CUDA_TRY(ctx, cudaFuncSetAttribute(conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H_TOTAL));
  const int M = p.B * p.Ho * p.Wo;
  const int mt = (M + HM - 1) / HM, pairs = (mt + 1) / 2;
  int cap = ctx->opt_grid_cap > 0 ? std::min(num_sms, ctx->opt_grid_cap) : num_sms;
  cap = std::max(2, cap & ~1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)std::min(2 * pairs, cap));
  cfg.blockDim = dim3(H_THREADS);
  cfg.dynamicSmemBytes = H_TOTAL;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (ctx->opt_pdl && !ctx->opt_multistream) ? 2 : 1;
  SUO_CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel, p, mt));
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}
