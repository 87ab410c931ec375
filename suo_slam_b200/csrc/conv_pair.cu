// CTA-pair 3x3 convolution for sm_100a: the bottleneck's conv2 (3x3, pad 1, Cin -> 128, folded BN + ReLU; reference
// lib/models/layers/Residual.py:13-15,28-31) as a cluster of two CTAs that run ONE tcgen05.mma.cta_group::2 of M = 256.
//
// Why: per 64-wide K chunk the single-CTA kernel (conv_tc.cu, <128, 3x3, TMA-fed>) brings 32 KB of activations AND 32 KB of weight
// images through L2 into shared memory, and all 148 CTAs pull the same weight lines at the same time.  As a pair, each CTA loads
// its own 128 pixels of A but only HALF of the weight rows (the tensor cores of both SMs read both halves), so a stage is 48 KB,
// the ring is four deep in the same 192 KB, and the L2 -> SM weight traffic halves (ncu: 14 % fewer L2 sectors per layer).
// Measured (profiles/r1_pair_vs_single_probe_ncu.txt): 1.7 % faster on an isolated 64x64 layer, +2.2 % frames/s end to end; both
// kernels keep the tensor pipe busy 64.8 % of the cycles (1236 cycles per chunk against ~800 of MMA) = 0.955 of the sustained
// cuBLAS bf16 rate under the board power cap — what is left is the shared-memory traffic of the operands themselves
// (72-80 KB read by the tensor core + 48-64 KB written by TMA per chunk against 128 B/clk), not load latency.
//
// Pair tile = 256 consecutive output pixels: CTA rank r owns pixels [256 t + 128 r, +128), i.e. accumulator rows
// [128 r, +128) of the M = 256 MMA, which live in its own TMEM.  Per K chunk (64 input channels of one tap):
//   shared memory stage (48 KB, same offsets in both CTAs): A_hi 16K | A_lo' 16K | B_hi rows [64 r, +64) 8K | B_lo' rows [64 r, +64) 8K
//   MMAs (leader only, N = 128):  main += A_hi x B_hi ;  corr += A_hi x B_lo' ;  corr += A_lo' x B_hi
// — the same products accumulated in the same order as the single-CTA kernel's merged N = 256 + N = 128 pair, so the
// two kernels give bit-identical outputs (tests/test_gpu_kernels.py, tests/test_gpu_net.py).
//
// Warps (6 per CTA): 0 producer (TMA: activations by 4-D tensor map with zero fill = the conv's padding, weight halves
// by a 2-D row map over the pre-swizzled images; every load completes on the LEADER's full barrier), 1 TMEM allocator +
// (leader only) MMA issuer, 2..5 epilogue (own TMEM lane quarter -> bias, ReLU, FP16 hi/lo' split -> swizzled staging
// boxes -> TMA store), exactly the split-output epilogue of conv_tc.cu.  tcgen05.commit multicasts "stage free" and
// "accumulator full" to both CTAs; the epilogue warps of both CTAs hand the accumulator back on the leader's barrier.
#include <cuda_fp16.h>
#include <algorithm>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int PM = 128;                                // pixels per CTA (256 per pair tile)
constexpr int PN = 128;                                // output channels
constexpr int P_PLANE = 16384;                         // one 128-row x 128-byte activation image
constexpr int P_BHALF = 8192;                          // this CTA's 64 rows of one weight image
constexpr int P_STAGE = 2 * P_PLANE + 2 * P_BHALF;     // 48 KB
constexpr int P_NS = 4;
constexpr int P_STG_OFF = P_NS * P_STAGE;              // 4 epilogue warps x (hi box + lo' box) x 4 KB
constexpr int P_BAR_OFF = P_STG_OFF + 4 * 2 * 4096;
constexpr int P_BIAS_OFF = P_BAR_OFF + 512;
constexpr int P_TOTAL = P_BIAS_OFF + PN * 4 + 1024;    // + alignment slack
constexpr int P_THREADS = 192;
static_assert(P_TOTAL <= 232448, "exceeds the 227 KB a CTA may use");

__global__ void __launch_bounds__(P_THREADS, 1)
conv3x3_pair_kernel(const __grid_constant__ ConvParams p, const int num_m_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // identical in both CTAs of the pair
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + P_BAR_OFF;
  auto full = [&](int s) { return bar_base + 8 * s; };                       // leader: both CTAs' TMA bytes of stage s have landed
  auto empty = [&](int s) { return bar_base + 8 * (P_NS + s); };             // both: the MMAs reading stage s are done
  auto tmem_full = [&](int b) { return bar_base + 8 * (2 * P_NS + b); };     // both: accumulator b is complete
  auto tmem_empty = [&](int b) { return bar_base + 8 * (2 * P_NS + 2 + b); };// leader: both CTAs' epilogues have drained b
  const uint32_t tmem_slot = bar_base + 8 * (2 * P_NS + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + P_BAR_OFF + 8 * (2 * P_NS + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = (int)blockIdx.x >> 1, num_clusters = (int)gridDim.x >> 1;
  const int num_pair_tiles = (num_m_tiles + 1) >> 1;
  const int M = p.B * p.Ho * p.Wo;
  const int nchunks = p.K / 64;

  unsigned long long trace_slot = 0;          // developer tool (SUO_TRACE): globaltimer at the start / end of CTA 0
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
    for (int s = 0; s < P_NS; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // the peer's barriers exist before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // Programmatic dependent launch (SUO_OPT_PDL): every CTA of this grid is resident (grid <= #SMs, one CTA per SM), so the next
  // conv kernel may be scheduled onto an SM the moment this kernel's CTA leaves it and run its own prologue (barriers, TMEM) there;
  // nothing above touched an activation tensor — from here on the previous kernel must have completed and flushed.
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== producer (both CTAs) =====================
    if (lane == 0) {
      const int HWi = p.H * p.W, cpc64 = p.Cin / 64;
      uint32_t g = 0;
      for (int pt = cluster_id; pt < num_pair_tiles; pt += num_clusters) {
        const int m0 = (2 * pt + (int)rank) * PM;                            // beyond M: coordinates outside the tensor -> zero fill
        const int b0 = m0 / HWi, rem = m0 - b0 * HWi;
        const int y0 = rem / p.W, x0 = rem - y0 * p.W;
        for (int j = 0; j < nchunks; ++j, ++g) {
          const int s = g % P_NS;
          mbar_wait(empty(s), ((g / P_NS) & 1) ^ 1);
          const uint32_t dst = smem_base + s * P_STAGE;
          const uint32_t lbar = mapa_shared(full(s), 0);
          const int tap = j / cpc64, cc = j - tap * cpc64, dy = tap / 3 - 1, dx = tap % 3 - 1;
          if (rank == 0) mbar_arrive_expect_tx(full(s), 2 * P_STAGE);        // this CTA's 48 KB + the peer's
          tma_load_4d_2sm(dst, p.tmap_hi, 64 * cc, x0 + dx, y0 + dy, b0, lbar);
          tma_load_4d_2sm(dst + P_PLANE, p.tmap_lo, 64 * cc, x0 + dx, y0 + dy, b0, lbar);
          // weight images of chunk j: rows [(2j) * 128, +128) = hi, [(2j + 1) * 128, +128) = lo'; this CTA takes 64 of each
          tma_load_2d_2sm(dst + 2 * P_PLANE, p.tmap_w, 0, (2 * j) * PN + 64 * (int)rank, lbar);
          tma_load_2d_2sm(dst + 2 * P_PLANE + P_BHALF, p.tmap_w, 0, (2 * j + 1) * PN + 64 * (int)rank, lbar);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(2 * PM, PN);
      uint32_t g = 0, i = 0;
      for (int pt = cluster_id; pt < num_pair_tiles; pt += num_clusters, ++i) {
        const uint32_t b = i & 1;
        mbar_wait_cluster(tmem_empty(b), ((i >> 1) & 1) ^ 1);                // both epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + b * 2 * PN;
        for (int j = 0; j < nchunks; ++j, ++g) {
          const int s = g % P_NS;
          mbar_wait(full(s), (g / P_NS) & 1);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_hi = smem_base + s * P_STAGE, a_lo = a_hi + P_PLANE, b_hi = a_hi + 2 * P_PLANE, b_lo = b_hi + P_BHALF;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t dah = make_sw128_desc(a_hi + kk * 32), dal = make_sw128_desc(a_lo + kk * 32);
              const uint64_t dbh = make_sw128_desc(b_hi + kk * 32), dbl = make_sw128_desc(b_lo + kk * 32);
              umma2_f16(acc, dah, dbh, idesc, (j | kk) != 0);                // main
              umma2_f16(acc + PN, dah, dbl, idesc, (j | kk) != 0);           // correction: hi x lo'
              umma2_f16(acc + PN, dal, dbh, idesc, 1u);                      //             lo' x hi
            }
            umma2_commit(empty(s));
            if (j == nchunks - 1) umma2_commit(tmem_full(b));
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs): TMEM -> bias, ReLU, FP16 hi/lo' split -> staging boxes -> TMA store =====================
    const int ew = warp - 2, q = warp & 3;
    float* bias_s = reinterpret_cast<float*>(smem_gen + P_BIAS_OFF);
    for (int c = threadIdx.x - 64; c < PN; c += 128) bias_s[c] = __ldg(p.bias + c);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const uint32_t stg_base = smem_base + P_STG_OFF + ew * 8192;
    uint8_t* stg_gen = smem_gen + P_STG_OFF + ew * 8192;
    const int swz = lane & 7;
    uint32_t i = 0;
    for (int pt = cluster_id; pt < num_pair_tiles; pt += num_clusters, ++i) {
      const int m_tile = 2 * pt + (int)rank;
      const uint32_t b = i & 1;
      const uint32_t acc = tmem_base + b * 2 * PN + ((uint32_t)(q * 32) << 16);
      const bool row_ok = m_tile * PM + q * 32 + lane < M;
      const bool box_ok = m_tile * PM + q * 32 < M;
#pragma unroll 1
      for (int gi = 0; gi < PN / 64; ++gi) {
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
            tmem_ld32(acc + (uint32_t)(PN + 64 * gi + 32 * h), rc);
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
        if (gi == PN / 64 - 1) {                      // last TMEM read of this tile: hand the accumulator back to the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_shared(tmem_empty(b), 0));
        }
        if (box_ok) {
          if (row_ok && amax > 60000.f && p.range_flag) *p.range_flag = 1;
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(p.tmap_out, stg_base, n_base, m_tile * PM + q * 32);
            tma_store_2d(p.tmap_out_lo, stg_base + 4096, n_base, m_tile * PM + q * 32);
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
  if (p.trace && threadIdx.x == 0 && blockIdx.x == 0 && trace_slot < 8192) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[1 + 4 * trace_slot + 1] = t;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace

bool conv_pair_eligible(const ConvParams& p, int passes) {
  return p.pair && p.mode == CONV_3x3 && p.math == 1 && passes == 3 && p.in_split && p.out_split && p.epi_tma && !p.residual &&
         !p.pre_scale && !p.out_nchw && p.Cout == PN && p.Cout_pad == PN && p.Cin % 64 == 0 && p.K == 9 * p.Cin && p.H == p.Ho && p.W == p.Wo;
}

int launch_conv_pair(suo_ctx* ctx, const ConvParams& p, cudaStream_t s) {
  static bool configured[64] = {};
  int num_sms = 148;
  if (first_use_on_device(configured, &num_sms)) SUO_CUDA_TRY(ctx, cudaFuncSetAttribute(conv3x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_TOTAL));
  const int M = p.B * p.Ho * p.Wo;
  const int mt = (M + PM - 1) / PM, pairs = (mt + 1) / 2;
  int cap = ctx->opt_grid_cap > 0 ? std::min(num_sms, ctx->opt_grid_cap) : num_sms;
  cap = std::max(2, cap & ~1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)std::min(2 * pairs, cap));
  cfg.blockDim = dim3(P_THREADS);
  cfg.dynamicSmemBytes = P_TOTAL;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (ctx->opt_pdl && !ctx->opt_multistream) ? 2 : 1;
  SUO_CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, conv3x3_pair_kernel, p, mt));
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}
