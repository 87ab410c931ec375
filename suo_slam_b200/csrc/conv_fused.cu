// Fused bottleneck tail for sm_100a: conv2 (3x3, 128 -> 128, folded BN + ReLU) -> conv3 (1x1, 128 -> 256) + bias + skip add
// in ONE persistent tcgen05 kernel (reference lib/models/layers/Residual.py:26-35: out = conv3(relu(bn2(conv2(.)))) + skip).
//
// Why: unfused, conv2 is tensor-pipe bound (~12.4 us per 128-pixel tile) and conv3 is HBM bound (~8.8 us per tile:
// it re-reads conv2's output, reads the skip tensor, writes the trunk) and the two run back to back.  Here conv2's
// accumulator never leaves the SM: the epilogue warps turn it into the swizzled FP16 hi/lo' operand tile of conv3 in shared
// memory, the tensor core runs conv3 from there, and the skip read / trunk write of tile i overlap conv2 of tile i+1.
// DRAM traffic per pixel drops from 3584 B (512 in + 512 out | 512 in + 1024 skip + 1024 out) to 2560 B.
//
// Shared memory (232448 B = all of it): three 64 KB operand buffers [A hi 16K | A lo' 16K | B hi|lo' 32K], 8 output staging
// boxes of 4 KB (one per epilogue warp, 32 rows x 128 B, 128B swizzle, left by TMA store), mbarriers, both biases.
// Buffer 2 doubles as conv3's A operand A3 (2 K-chunks x (hi 16K + lo' 16K)): conv2's main loop rotates over all three
// buffers (a 128 KB ring cannot keep ~107 GB/s per SM of operand traffic in flight over the ~1.6 us L2 latency: the
// first version of this kernel, with a 2-stage ring, ran conv2 at 1.1 us per chunk instead of 0.7), and while A3 is live
// (epilogue fill -> conv3 half 1 done) the few chunks issued in that window alternate between buffers 0 and 1.
// TMEM (512 columns = all of it): conv2 accumulator [0,256) = main | correction, conv3 half-accumulator [256,512).
// conv3 runs as two halves of 128 output channels so that its accumulator fits; both accumulators are single-buffered,
// so the tensor pipe idles while the 8 epilogue warps drain conv2's accumulator (~1 us per ~13 us tile).
//
// Warps: 0 producer (TMA: activations by 4-D tensor map with zero fill for the 3x3 padding, weight images by bulk copy),
// 1 MMA issuer, 2..9 epilogue (TMEM lane quarter = warp & 3, column half = (warp - 2) / 4).
// Producer and MMA warp walk the SAME static schedule of buffer uses per tile i:
//   conv2(i) chunks [i == 0 ? 0 : E, 18) on buffers 0,1,2 | A3 use of buffer 2 | conv3(i) half 0 (2 weight chunks) |
//   conv2(i+1) chunks [0, E) | conv3(i) half 1            (the last four items on buffers 0,1)
// (the E early chunks of the next tile keep the tensor pipe busy while half 0 is being drained).
// Math: fp16x3 exactly as conv_tc.cu (x = hi + 2^-11 lo', merged A_hi x [B_hi | B_lo'] MMA, separate correction columns).
#include <cuda_fp16.h>
#include <algorithm>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int FM = 128;                       // pixels per tile
constexpr int CM = 128;                       // bottleneck width
constexpr int F_NS = 2;                       // dedicated ring stages; buffer 2 is the A3 region
constexpr int F_STAGE = 65536;
constexpr int F_PLANE = 16384;                // one 128-row x 128-byte operand image
constexpr int F_A3_OFF = F_NS * F_STAGE;
constexpr int F_STG_OFF = F_A3_OFF + 4 * F_PLANE;
constexpr int F_BAR_OFF = F_STG_OFF + 8 * 4096;
constexpr int F_BIAS_OFF = F_BAR_OFF + 512;
constexpr int F_TOTAL = F_BIAS_OFF + (CM + 2 * CM) * 4 + 1024;
constexpr int F_THREADS = 320;
constexpr int F_CHUNKS2 = 9 * CM / 64;        // 18 K-chunks of conv2
constexpr int F_EARLY = 2;                    // chunks of the next tile issued between the two conv3 halves
static_assert(F_TOTAL <= 232448, "exceeds the 227 KB a CTA may use");

__global__ void __launch_bounds__(F_THREADS, 1)
conv_fused23_kernel(const __grid_constant__ FusedParams p, const int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + F_BAR_OFF;
  auto full = [&](int b) { return bar_base + 8 * b; };                       // b in {0, 1, 2}: operand buffer filled by TMA
  auto empty = [&](int b) { return bar_base + 8 * (3 + b); };                // ... released by the tensor core
  auto buf_addr = [&](int b) { return b < 2 ? smem_base + b * F_STAGE : smem_base + F_A3_OFF; };
  const uint32_t acc2_full = bar_base + 8 * 6, acc2_empty = bar_base + 8 * 7, a3_full = bar_base + 8 * 8;
  const uint32_t acc3_full = bar_base + 8 * 9, acc3_empty = bar_base + 8 * 10, tmem_slot = bar_base + 8 * 11;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + F_BAR_OFF + 8 * 11);
  // static schedule state, walked identically by the producer and the MMA warp
  struct Sched { uint32_t uses[3], fills[3]; int rr3, rr2; };
  auto pick = [](Sched& sc, bool deep) { int b; if (deep) { b = sc.rr3; sc.rr3 = (sc.rr3 + 1) % 3; } else { b = sc.rr2; sc.rr2 ^= 1; } return b; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = p.B * p.H * p.W;
  const int my_tiles = ((int)blockIdx.x < num_tiles) ? (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto tile_of = [&](int i) { return (int)blockIdx.x + i * (int)gridDim.x; };
  auto stamp = [&](int row, int i) { if (p.dbg && blockIdx.x == 0 && lane == 0 && i < 64) p.dbg[row * 64 + i] = clock64(); };

  if (threadIdx.x == 0) {
    for (int b = 0; b < 3; ++b) { mbar_init(full(b), 1); mbar_init(empty(b), 1); }
    mbar_init(acc2_full, 1); mbar_init(acc2_empty, 8); mbar_init(a3_full, 8);
    mbar_init(acc3_full, 1); mbar_init(acc3_empty, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      Sched sc{{0, 0, 0}, {0, 0, 0}, 0, 0};
      const int HW = p.H * p.W;
      auto acquire = [&](bool deep) {             // next buffer of the schedule, once its previous use has been released
        const int b = pick(sc, deep);
        mbar_wait(empty(b), (sc.uses[b] & 1) ^ 1);
        ++sc.uses[b];
        return b;
      };
      auto put_conv2 = [&](int t, int j, bool deep) {
        const int b = acquire(deep);
        const uint32_t st = buf_addr(b);
        const int m0 = t * FM, b0 = m0 / HW, rem = m0 - b0 * HW, y0 = rem / p.W, x0 = rem - y0 * p.W;
        const int tap = j / (CM / 64), cc = j - tap * (CM / 64), dy = tap / 3 - 1, dx = tap % 3 - 1;
        mbar_arrive_expect_tx(full(b), 4 * F_PLANE);
        tma_load_4d(st, p.tmap_hi, 64 * cc, x0 + dx, y0 + dy, b0, full(b));
        tma_load_4d(st + F_PLANE, p.tmap_lo, 64 * cc, x0 + dx, y0 + dy, b0, full(b));
        bulk_g2s(st + 2 * F_PLANE, p.w2 + (size_t)j * (2 * CM * 64), 2 * F_PLANE, full(b));
      };
      auto put_b3 = [&](int h, int kc) {        // conv3 weight image of (output half h, K-chunk kc): [hi 128 rows | lo' 128 rows]
        const int b = acquire(false);
        mbar_arrive_expect_tx(full(b), 2 * F_PLANE);
        bulk_g2s(buf_addr(b) + 2 * F_PLANE, p.w3 + (size_t)(h * 2 + kc) * (2 * CM * 64), 2 * F_PLANE, full(b));
      };
      for (int i = 0; i < my_tiles; ++i) {
        for (int j = (i == 0 ? 0 : F_EARLY); j < F_CHUNKS2; ++j) put_conv2(tile_of(i), j, true);
        ++sc.uses[2];                             // buffer 2 now serves as A3 (filled by the epilogue warps, released after half 1)
        put_b3(0, 0); put_b3(0, 1);
        if (i + 1 < my_tiles) for (int j = 0; j < F_EARLY; ++j) put_conv2(tile_of(i + 1), j, false);
        put_b3(1, 0); put_b3(1, 1);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc1 = make_idesc_f16(FM, CM), idesc2 = make_idesc_f16(FM, 2 * CM);
    const uint32_t acc2 = tmem_base, acc3 = tmem_base + 2 * CM;
    const uint32_t a3 = smem_base + F_A3_OFF;
    Sched sc{{0, 0, 0}, {0, 0, 0}, 0, 0};
    uint32_t n3 = 0;
    auto next_full = [&](bool deep) {             // next buffer of the schedule, once its TMA bytes have landed
      const int b = pick(sc, deep);
      mbar_wait(full(b), sc.fills[b] & 1);
      ++sc.fills[b];
      tc_fence_after();
      return b;
    };
    auto conv2_chunk = [&](int i, int j, bool deep) {
      if (j == 0) { mbar_wait(acc2_empty, (i & 1) ^ 1); tc_fence_after(); }      // the epilogue has drained conv2's accumulator
      const int b = next_full(deep);
      if (lane == 0) {
        const uint32_t a_hi = buf_addr(b), a_lo = a_hi + F_PLANE, b_hi = a_hi + 2 * F_PLANE;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t dah = make_sw128_desc(a_hi + kk * 32), dal = make_sw128_desc(a_lo + kk * 32), dbh = make_sw128_desc(b_hi + kk * 32);
          umma_f16(acc2, dah, dbh, idesc2, (j | kk) != 0);          // A_hi x [B_hi | B_lo'] -> main | correction
          umma_f16(acc2 + CM, dal, dbh, idesc1, 1u);                // A_lo' x B_hi -> correction
        }
        umma_commit(empty(b));
        if (j == F_CHUNKS2 - 1) umma_commit(acc2_full);
      }
      __syncwarp();
    };
    auto conv3_half = [&](int i, int h) {
      if (h == 0) { mbar_wait(a3_full, i & 1); stamp(2, i); }       // conv3's A operand of this tile is in shared memory
      mbar_wait(acc3_empty, (n3 & 1) ^ 1);                          // the previous half has been drained
      tc_fence_after();
      if (h == 1) stamp(5, i);
      for (int kc = 0; kc < 2; ++kc) {
        const int b = next_full(false);
        if (lane == 0) {
          const uint32_t a_hi = a3 + kc * 2 * F_PLANE, a_lo = a_hi + F_PLANE, b_hi = buf_addr(b) + 2 * F_PLANE;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t dah = make_sw128_desc(a_hi + kk * 32), dal = make_sw128_desc(a_lo + kk * 32), dbh = make_sw128_desc(b_hi + kk * 32);
            umma_f16(acc3, dah, dbh, idesc2, (kc | kk) != 0);
            umma_f16(acc3 + CM, dal, dbh, idesc1, 1u);
          }
          umma_commit(empty(b));
          if (kc == 1) { umma_commit(acc3_full); if (h == 1) umma_commit(empty(2)); }    // half 1 done: buffer 2 (A3) is released
        }
        __syncwarp();
      }
      ++n3;
    };
    for (int i = 0; i < my_tiles; ++i) {
      stamp(0, i);
      for (int j = (i == 0 ? 0 : F_EARLY); j < F_CHUNKS2; ++j) conv2_chunk(i, j, true);
      stamp(1, i);
      conv3_half(i, 0);
      stamp(3, i);
      if (i + 1 < my_tiles) for (int j = 0; j < F_EARLY; ++j) conv2_chunk(i + 1, j, false);
      stamp(4, i);
      conv3_half(i, 1);
      stamp(6, i);
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2, q = warp & 3, cg = ew >> 2;           // staging box / TMEM lane quarter / column half
    float* bias2_s = reinterpret_cast<float*>(smem_gen + F_BIAS_OFF);
    float* bias3_s = bias2_s + CM;
    for (int c = threadIdx.x - 64; c < 3 * CM; c += 256) bias2_s[c] = c < CM ? __ldg(p.bias2 + c) : __ldg(p.bias3 + c - CM);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t box = smem_base + F_STG_OFF + ew * 4096;
    uint8_t* box_gen = smem_gen + F_STG_OFF + ew * 4096;
    uint8_t* a3_gen = smem_gen + F_A3_OFF + cg * 2 * F_PLANE + (q * 32 + lane) * 128;      // this lane's row of K-chunk cg (hi; lo' at + F_PLANE)
    const int swz = lane & 7;
    const int C3 = 2 * CM;
    uint32_t n3 = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int m_tile = tile_of(i);
      const int m = m_tile * FM + q * 32 + lane;
      const bool row_ok = m < M;
      const float* skip_row = p.skip + (size_t)(row_ok ? m : 0) * C3;
      float4 sk[2][8];
      auto load_skip = [&](int h) {                                 // this lane's 2 x 32 skip values of half h
#pragma unroll
        for (int uu = 0; uu < 2; ++uu)
#pragma unroll
          for (int c = 0; c < 8; ++c)
            sk[uu][c] = row_ok ? __ldg(reinterpret_cast<const float4*>(skip_row + CM * h + 32 * (2 * cg + uu)) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      load_skip(0);                                                 // in flight during the wait for conv2's main loop
      // ---- conv2 accumulator -> bias + ReLU -> FP16 hi / lo' -> conv3's A operand ----
      mbar_wait_backoff<32>(acc2_full, i & 1);
      tc_fence_after();
      if (warp == 2) stamp(7, i);
      // (A3 = buffer 2 is free: acc2_full implies every conv2 chunk that used it has been consumed, and the producer
      //  does not refill it before conv3 half 1 releases it)
      float amax = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t r[32], rc[32];
        tmem_ld32(tmem_base + lane_off + (uint32_t)(64 * cg + 32 * h), r);
        tmem_ld32(tmem_base + lane_off + (uint32_t)(CM + 64 * cg + 32 * h), rc);
        tmem_ld_wait();
        if (h == 1) {                                               // last TMEM read of conv2's accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc2_empty);
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(fmaf(__uint_as_float(rc[c]), 1.0f / 2048.0f, __uint_as_float(r[c])));
#pragma unroll
        for (int k = 0; k < 4; ++k) {                               // 8 channels -> one 16-byte chunk of each plane
          uint32_t hp[4], lp[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = 8 * k + 2 * e;
            const float o0 = fmaxf(__uint_as_float(r[c]) + bias2_s[64 * cg + 32 * h + c], 0.f);
            const float o1 = fmaxf(__uint_as_float(r[c + 1]) + bias2_s[64 * cg + 32 * h + c + 1], 0.f);
            amax = fmaxf(amax, fmaxf(o0, o1));
            const float h0 = __uint_as_float(__float_as_uint(o0) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(o1) & 0xFFFFE000u);
            const __half2 hh = __floats2half2_rn(h0, h1), ll = __floats2half2_rn((o0 - h0) * 2048.f, (o1 - h1) * 2048.f);
            hp[e] = *reinterpret_cast<const uint32_t*>(&hh); lp[e] = *reinterpret_cast<const uint32_t*>(&ll);
          }
          const int ch = ((4 * h + k) ^ swz) << 4;
          *reinterpret_cast<uint4*>(a3_gen + ch) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
          *reinterpret_cast<uint4*>(a3_gen + F_PLANE + ch) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
        }
      }
      if (row_ok && amax > 60000.f && p.range_flag) *p.range_flag = 1;
      fence_proxy_async();                                          // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(a3_full);
      if (warp == 2) stamp(8, i);
      // ---- conv3 accumulator halves -> + bias + skip -> staging box -> TMA store ----
#pragma unroll 1
      for (int h = 0; h < 2; ++h, ++n3) {
        if (h == 1) load_skip(1);                                   // in flight while the tensor core runs the early chunks + half 1
        mbar_wait_backoff<32>(acc3_full, n3 & 1);
        tc_fence_after();
        if (warp == 2) stamp(9 + 2 * h, i);
#pragma unroll
        for (int uu = 0; uu < 2; ++uu) {
          const int c0 = 32 * (2 * cg + uu), n_base = CM * h + c0;
          uint32_t r[32], rc[32];
          tmem_ld32(tmem_base + lane_off + (uint32_t)(2 * CM + c0), r);
          tmem_ld32(tmem_base + lane_off + (uint32_t)(3 * CM + c0), rc);
          if (lane == 0) bulk_wait_read<0>();                       // the previous store has left this warp's box
          __syncwarp();
          tmem_ld_wait();
          if (uu == 1) {                                            // last TMEM read of this half
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc3_empty);
          }
          uint8_t* rowp = box_gen + lane * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 bq = *reinterpret_cast<const float4*>(bias3_s + n_base + 4 * c);
            const float4 sq = sk[uu][c];
            float4 o;
            o.x = fmaf(__uint_as_float(rc[4 * c]), 1.0f / 2048.0f, __uint_as_float(r[4 * c])) + bq.x + sq.x;
            o.y = fmaf(__uint_as_float(rc[4 * c + 1]), 1.0f / 2048.0f, __uint_as_float(r[4 * c + 1])) + bq.y + sq.y;
            o.z = fmaf(__uint_as_float(rc[4 * c + 2]), 1.0f / 2048.0f, __uint_as_float(r[4 * c + 2])) + bq.z + sq.z;
            o.w = fmaf(__uint_as_float(rc[4 * c + 3]), 1.0f / 2048.0f, __uint_as_float(r[4 * c + 3])) + bq.w + sq.w;
            *reinterpret_cast<float4*>(rowp + ((c ^ swz) << 4)) = o;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && m_tile * FM + q * 32 < M) {
            tma_store_2d(p.tmap_out, box, n_base, m_tile * FM + q * 32);
            bulk_commit();
          }
        }
        if (warp == 2) stamp(10 + 2 * h, i);
      }
    }
    if (lane == 0) bulk_wait_read<0>();                             // shared memory must outlive the last stores' reads
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int launch_conv_fused23(suo_ctx* ctx, const FusedParams& p, cudaStream_t s) {
  static bool configured[64] = {};
  int num_sms = 148;
  if (first_use_on_device(configured, &num_sms)) SUO_CUDA_TRY(ctx, cudaFuncSetAttribute(conv_fused23_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_TOTAL));
  const int M = p.B * p.H * p.W;
  const int tiles = (M + FM - 1) / FM;
  const int grid = std::min(tiles, ctx->opt_grid_cap > 0 ? std::min(num_sms, ctx->opt_grid_cap) : num_sms);
  conv_fused23_kernel<<<grid, F_THREADS, F_TOTAL, s>>>(p, tiles);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}
