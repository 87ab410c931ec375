// Fused bottleneck tail as a CTA pair for sm_100a: conv2 (3x3, 128 -> 128, folded BN + ReLU) -> conv3 (1x1, 128 -> 256) + bias +
// skip add in ONE persistent kernel whose MMAs are tcgen05.mma.cta_group::2 of M = 256 (reference
// lib/models/layers/Residual.py:26-35: out = conv3(relu(bn2(conv2(.)))) + skip).
//
// Same idea as conv_fused.cu (conv2's accumulator never leaves the SM: the epilogue warps turn it into conv3's swizzled FP16
// hi/lo' operand in shared memory, the skip read / trunk write of tile i overlap conv2 of tile i + 1; DRAM traffic per pixel
// 3584 B -> 2560 B), rebuilt on the CTA-pair scheme of conv_pair.cu because the single-CTA version ran out of shared memory:
// with 64 KB stages the operand ring shrank to two buffers exactly while conv3 needed its weights, and every conv3 weight
// load sat behind a conv2 chunk (clock64 timeline: 14k of 37k cycles per tile serialised on load latency).  As a pair each
// CTA loads half of every weight image, so a conv2 stage is 48 KB and conv3's weights get a buffer of their own:
//
//   shared memory per CTA (232448 B = all of it, identical offsets in both CTAs)
//     S0, S1   2 x 48 KB   conv2 operand stages  [A_hi 16K | A_lo' 16K | B_hi rows [64 r, +64) 8K | B_lo' rows 8K]
//     A3       64 KB       conv3's A operand of this CTA's 128 pixels (2 K-chunks x (hi 16K + lo' 16K)); outside the window
//                          [conv2 accumulator drained -> conv3 half 1 done] its first 48 KB are conv2's third stage S2
//     B3       32 KB       this CTA's 64 rows of conv3's weight images for ONE output half (2 K-chunks x (hi 8K + lo' 8K))
//     staging  8 x 4 KB    one TMA-store box per epilogue warp;   mbarriers, both biases
//   TMEM (512 columns): conv2 accumulator [0,256) = main | correction, conv3 half-accumulator [256,512).
//
// Warps: 0 producer (both CTAs; every load completes on the LEADER's barrier), 1 TMEM allocator + (leader) MMA issuer,
// 2..9 epilogue (TMEM lane quarter = warp & 3, column half = (warp - 2) / 4), both CTAs, each on its own 128 pixels.
// Producers and the MMA warp walk the same static schedule; per pair tile i
//   MMA order      conv2(i) chunks [E, 18) | wait A3(i) | conv3(i) half 0 | conv2(i+1) chunks [0, E) | conv3(i) half 1
//   producer order the same chunk order, with the B3 loads of half 0 / half 1 slipped in where their buffer frees
// (the E early chunks keep the tensor pipe busy while half 0 is drained).  Stage buffers are chosen least-recently-assigned
// among {S0, S1, S2} (main loop) or {S0, S1} (while A3 is live) — a pure function of the schedule, so both sides agree.
// Math: fp16x3 exactly as conv_tc.cu / conv_pair.cu (same products, same accumulation order): bit-identical outputs.
#include <cuda_fp16.h>
#include <algorithm>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int GM = 128;                       // pixels per CTA (256 per pair tile)
constexpr int GC = 128;                       // bottleneck width
constexpr int G_PLANE = 16384;                // 128 rows x 128 B
constexpr int G_BHALF = 8192;                 // 64 rows x 128 B
constexpr int G_STAGE = 2 * G_PLANE + 2 * G_BHALF;      // 48 KB
constexpr int G_A3_OFF = 2 * G_STAGE;                    // 96 KB
constexpr int G_B3_OFF = G_A3_OFF + 4 * G_PLANE;         // 160 KB
constexpr int G_STG_OFF = G_B3_OFF + 4 * G_BHALF;        // 192 KB
constexpr int G_BAR_OFF = G_STG_OFF + 8 * 4096;          // 224 KB
constexpr int G_BIAS_OFF = G_BAR_OFF + 512;
constexpr int G_TOTAL = G_BIAS_OFF + (GC + 2 * GC) * 4 + 1024;
constexpr int G_THREADS = 320;
constexpr int G_CHUNKS2 = 9 * GC / 64;        // 18 K-chunks of conv2
constexpr int G_EARLY = 3;                    // chunks of the next tile issued between the two conv3 halves
static_assert(G_TOTAL <= 232448, "exceeds the 227 KB a CTA may use");

// Static stage-buffer schedule, walked identically by both producers and the MMA warp.
struct Sched {
  uint32_t cnt[3] = {0, 0, 0};                // uses (producer) / fills (MMA warp) per buffer
  int last[3] = {-3, -2, -1};                 // sequence number of the last assignment per buffer
  int seq = 0;
  __device__ int pick(bool deep) {            // least recently assigned buffer among {0,1,2} (deep) or {0,1}
    int b = last[0] <= last[1] ? 0 : 1;
    if (deep && last[2] < last[b]) b = 2;
    last[b] = seq;
    seq += 2;
    return b;
  }
  // buffer 2 becomes A3: it is released (conv3 half 1 done) after the buffers of the G_EARLY early chunks that follow, before any later chunk's
  __device__ void claim_a3() { last[2] = seq + 2 * G_EARLY - 1; }
};

__global__ void __launch_bounds__(G_THREADS, 1)
conv_fused23_pair_kernel(const __grid_constant__ FusedParams p, const int num_m_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + G_BAR_OFF;
  auto full = [&](int b) { return bar_base + 8 * b; };                       // leader: stage b filled by both CTAs' TMA
  auto empty = [&](int b) { return bar_base + 8 * (3 + b); };                // both: stage b released by the tensor core
  auto buf_addr = [&](int b) { return b < 2 ? smem_base + b * G_STAGE : smem_base + G_A3_OFF; };
  const uint32_t acc2_full = bar_base + 8 * 6, acc2_empty = bar_base + 8 * 7, a3_full = bar_base + 8 * 8;
  const uint32_t acc3_full = bar_base + 8 * 9, acc3_empty = bar_base + 8 * 10, b3_full = bar_base + 8 * 11, b3_empty = bar_base + 8 * 12;
  const uint32_t tmem_slot = bar_base + 8 * 13;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + G_BAR_OFF + 8 * 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = (int)blockIdx.x >> 1, num_clusters = (int)gridDim.x >> 1;
  const int num_pair_tiles = (num_m_tiles + 1) >> 1;
  const int M = p.B * p.H * p.W;
  const int my_tiles = cluster_id < num_pair_tiles ? (num_pair_tiles - cluster_id + num_clusters - 1) / num_clusters : 0;
  auto mtile_of = [&](int i) { return 2 * (cluster_id + i * num_clusters) + (int)rank; };
  auto stamp = [&](int row, int i) { if (p.dbg && blockIdx.x == 0 && lane == 0 && i < 64) p.dbg[row * 64 + i] = clock64(); };

  if (threadIdx.x == 0) {
    for (int b = 0; b < 3; ++b) { mbar_init(full(b), 1); mbar_init(empty(b), 1); }
    mbar_init(acc2_full, 1); mbar_init(acc2_empty, 16); mbar_init(a3_full, 16);       // 8 epilogue warps x 2 CTAs arrive on the leader
    mbar_init(acc3_full, 1); mbar_init(acc3_empty, 16);
    mbar_init(b3_full, 1); mbar_init(b3_empty, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===================== producer (both CTAs) =====================
    if (lane == 0) {
      Sched sc;
      uint32_t n3 = 0;                            // conv3 halves whose weights have been requested
      const int HW = p.H * p.W;
      const uint32_t b3_lbar = mapa_shared(b3_full, 0);
      auto put_conv2 = [&](int i, int j, bool deep) {
        const int b = sc.pick(deep);
        mbar_wait(empty(b), (sc.cnt[b] & 1) ^ 1);
        ++sc.cnt[b];
        const uint32_t st = buf_addr(b);
        const uint32_t lbar = mapa_shared(full(b), 0);
        const int m0 = mtile_of(i) * GM, b0 = m0 / HW, rem = m0 - b0 * HW, y0 = rem / p.W, x0 = rem - y0 * p.W;
        const int tap = j / (GC / 64), cc = j - tap * (GC / 64), dy = tap / 3 - 1, dx = tap % 3 - 1;
        if (rank == 0) mbar_arrive_expect_tx(full(b), 2 * G_STAGE);
        tma_load_4d_2sm(st, p.tmap_hi, 64 * cc, x0 + dx, y0 + dy, b0, lbar);
        tma_load_4d_2sm(st + G_PLANE, p.tmap_lo, 64 * cc, x0 + dx, y0 + dy, b0, lbar);
        tma_load_2d_2sm(st + 2 * G_PLANE, p.tmap_w2, 0, (2 * j) * GC + 64 * (int)rank, lbar);
        tma_load_2d_2sm(st + 2 * G_PLANE + G_BHALF, p.tmap_w2, 0, (2 * j + 1) * GC + 64 * (int)rank, lbar);
        // L2 prefetch, paced by the main loop: with chunk j of tile i goes 1/16 of what tile i + 1 of this CTA will read from
        // HBM (128 KB of skip rows, 2 x 32 KB of conv2 input planes: all three are contiguous byte ranges).  Without it the
        // epilogues of all CTAs fetch their skip tiles in the same few microseconds, HBM saturates in bursts and the operand
        // loads of the next tile's first chunks queue behind them (clock64 timeline: 7k-cycle gaps at chunks 3 and 5).
        if (p.prefetch && j < 16 && i + 1 < my_tiles) {
          const size_t m1 = (size_t)mtile_of(i + 1) * GM;
          if (m1 + GM <= (size_t)M) {
            bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(p.skip + m1 * (2 * GC)) + j * 8192, 8192);
            bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(p.in_hi + m1 * GC) + j * 2048, 2048);
            bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(p.in_lo + m1 * GC) + j * 2048, 2048);
          }
        }
      };
      auto put_b3 = [&](int h) {                // this CTA's 64 rows of conv3's weight images of output half h: kc0 hi | kc0 lo' | kc1 hi | kc1 lo'
        mbar_wait(b3_empty, (n3 & 1) ^ 1);
        ++n3;
        if (rank == 0) mbar_arrive_expect_tx(b3_full, 2 * 4 * G_BHALF);
#pragma unroll
        for (int kc = 0; kc < 2; ++kc)
#pragma unroll
          for (int pl = 0; pl < 2; ++pl)
            tma_load_2d_2sm(smem_base + G_B3_OFF + (2 * kc + pl) * G_BHALF, p.tmap_w3, 0, ((h * 2 + kc) * 2 + pl) * GC + 64 * (int)rank, b3_lbar);
      };
      for (int i = 0; i < my_tiles; ++i) {
        const int j0 = i == 0 ? 0 : G_EARLY;
        for (int j = j0; j < G_CHUNKS2; ++j) {
          put_conv2(i, j, true);
          if (j == j0 + 2) put_b3(0);           // behind the first three chunks, so the ring is full while this may wait for half 1 of tile i - 1
        }
        sc.claim_a3();
        ++sc.cnt[2];                            // buffer 2 now serves as A3 (filled by the epilogue warps, released after half 1)
        const int ne = i + 1 < my_tiles ? G_EARLY : 0;
        for (int j = 0; j < ne; ++j) {
          if (j == 2) put_b3(1);                // half 0's MMAs release the B3 buffer long before the third early chunk's stage frees
          put_conv2(i + 1, j, false);
        }
        if (ne <= 2) put_b3(1);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(2 * GM, GC);
      const uint32_t acc2 = tmem_base, acc3 = tmem_base + 2 * GC;
      const uint32_t a3 = smem_base + G_A3_OFF, b3 = smem_base + G_B3_OFF;
      Sched sc;
      uint32_t n3 = 0;
      auto conv2_chunk = [&](int i, int j, bool deep) {
        if (j == 0) { mbar_wait_cluster(acc2_empty, (i & 1) ^ 1); tc_fence_after(); }      // both epilogues have drained conv2's accumulator
        const int b = sc.pick(deep);
        mbar_wait(full(b), sc.cnt[b] & 1);
        ++sc.cnt[b];
        tc_fence_after();
        const int slot = (i == 3 || i == 4) ? (i - 3) * 18 + j : -1;     // timeline: chunks of tiles 3 and 4
        if (slot >= 0) stamp(13, slot);
        if (lane == 0) {
          const uint32_t a_hi = buf_addr(b), a_lo = a_hi + G_PLANE, b_hi = a_hi + 2 * G_PLANE, b_lo = b_hi + G_BHALF;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t dah = make_sw128_desc(a_hi + kk * 32), dal = make_sw128_desc(a_lo + kk * 32);
            const uint64_t dbh = make_sw128_desc(b_hi + kk * 32), dbl = make_sw128_desc(b_lo + kk * 32);
            umma2_f16(acc2, dah, dbh, idesc, (j | kk) != 0);
            umma2_f16(acc2 + GC, dah, dbl, idesc, (j | kk) != 0);
            umma2_f16(acc2 + GC, dal, dbh, idesc, 1u);
          }
          umma2_commit(empty(b));
          if (j == G_CHUNKS2 - 1) umma2_commit(acc2_full);
        }
        if (slot >= 0) stamp(14, slot);
        __syncwarp();
      };
      auto conv3_half = [&](int i, int h) {
        if (h == 0) { mbar_wait_cluster(a3_full, i & 1); stamp(2, i); }     // conv3's A operand of this tile is in both CTAs' shared memory
        mbar_wait_cluster(acc3_empty, (n3 & 1) ^ 1);                        // the previous half has been drained by both CTAs
        if (h == 1) stamp(5, i);
        mbar_wait(b3_full, n3 & 1);
        tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int kc = 0; kc < 2; ++kc) {
            const uint32_t a_hi = a3 + kc * 2 * G_PLANE, a_lo = a_hi + G_PLANE, b_hi = b3 + kc * 2 * G_BHALF, b_lo = b_hi + G_BHALF;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t dah = make_sw128_desc(a_hi + kk * 32), dal = make_sw128_desc(a_lo + kk * 32);
              const uint64_t dbh = make_sw128_desc(b_hi + kk * 32), dbl = make_sw128_desc(b_lo + kk * 32);
              umma2_f16(acc3, dah, dbh, idesc, (kc | kk) != 0);
              umma2_f16(acc3 + GC, dah, dbl, idesc, (kc | kk) != 0);
              umma2_f16(acc3 + GC, dal, dbh, idesc, 1u);
            }
          }
          umma2_commit(b3_empty);
          umma2_commit(acc3_full);
          if (h == 1) umma2_commit(empty(2));                               // half 1 done: buffer 2 (A3) is a conv2 stage again
        }
        __syncwarp();
        ++n3;
      };
      for (int i = 0; i < my_tiles; ++i) {
        stamp(0, i);
        for (int j = (i == 0 ? 0 : G_EARLY); j < G_CHUNKS2; ++j) conv2_chunk(i, j, true);
        stamp(1, i);
        sc.claim_a3();
        conv3_half(i, 0);
        stamp(3, i);
        if (i + 1 < my_tiles) for (int j = 0; j < G_EARLY; ++j) conv2_chunk(i + 1, j, false);
        stamp(4, i);
        conv3_half(i, 1);
        stamp(6, i);
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs) =====================
    const int ew = warp - 2, q = warp & 3, cg = ew >> 2;           // staging box / TMEM lane quarter / column half
    float* bias2_s = reinterpret_cast<float*>(smem_gen + G_BIAS_OFF);
    float* bias3_s = bias2_s + GC;
    for (int c = threadIdx.x - 64; c < 3 * GC; c += 256) bias2_s[c] = c < GC ? __ldg(p.bias2 + c) : __ldg(p.bias3 + c - GC);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t box = smem_base + G_STG_OFF + ew * 4096;
    uint8_t* box_gen = smem_gen + G_STG_OFF + ew * 4096;
    uint8_t* a3_gen = smem_gen + G_A3_OFF + cg * 2 * G_PLANE + (q * 32 + lane) * 128;      // this lane's row of K-chunk cg (hi; lo' at + G_PLANE)
    const uint32_t l_acc2_empty = mapa_shared(acc2_empty, 0), l_a3_full = mapa_shared(a3_full, 0), l_acc3_empty = mapa_shared(acc3_empty, 0);
    const int swz = lane & 7;
    const int C3 = 2 * GC;
    uint32_t n3 = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int m_tile = mtile_of(i);
      const int m = m_tile * GM + q * 32 + lane;
      const bool row_ok = m < M;
      const float* skip_row = p.skip + (size_t)(row_ok ? m : 0) * C3;
      float4 sk[2][8];
      auto load_skip = [&](int h) {                                 // this lane's 2 x 32 skip values of half h
#pragma unroll
        for (int uu = 0; uu < 2; ++uu)
#pragma unroll
          for (int c = 0; c < 8; ++c)
            sk[uu][c] = row_ok ? __ldg(reinterpret_cast<const float4*>(skip_row + GC * h + 32 * (2 * cg + uu)) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      load_skip(0);                                                 // in flight during the wait for conv2's main loop
      // ---- conv2 accumulator -> bias + ReLU -> FP16 hi / lo' -> conv3's A operand ----
      mbar_wait_backoff<32>(acc2_full, i & 1);
      tc_fence_after();
      if (warp == 2) stamp(7, i);
      // (A3 = buffer 2 is free: acc2_full implies every conv2 chunk that used it has been consumed, and the producers
      //  do not refill it before conv3 half 1 releases it)
      float amax = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t r[32], rc[32];
        tmem_ld32(tmem_base + lane_off + (uint32_t)(64 * cg + 32 * h), r);
        tmem_ld32(tmem_base + lane_off + (uint32_t)(GC + 64 * cg + 32 * h), rc);
        tmem_ld_wait();
        if (h == 1) {                                               // last TMEM read of conv2's accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(l_acc2_empty);
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
          *reinterpret_cast<uint4*>(a3_gen + G_PLANE + ch) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
        }
      }
      if (row_ok && amax > 60000.f && p.range_flag) *p.range_flag = 1;
      fence_proxy_async();                                          // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(l_a3_full);
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
          const int c0 = 32 * (2 * cg + uu), n_base = GC * h + c0;
          uint32_t r[32], rc[32];
          tmem_ld32(tmem_base + lane_off + (uint32_t)(2 * GC + c0), r);
          tmem_ld32(tmem_base + lane_off + (uint32_t)(3 * GC + c0), rc);
          if (lane == 0) bulk_wait_read<0>();                       // the previous store has left this warp's box
          __syncwarp();
          tmem_ld_wait();
          if (uu == 1) {                                            // last TMEM read of this half
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(l_acc3_empty);
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
          if (lane == 0 && m_tile * GM + q * 32 < M) {
            tma_store_2d(p.tmap_out, box, n_base, m_tile * GM + q * 32);
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
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace

int launch_conv_fused23_pair(suo_ctx* ctx, const FusedParams& p, cudaStream_t s) {
  static bool configured[64] = {};
  int num_sms = 148;
  if (first_use_on_device(configured, &num_sms)) SUO_CUDA_TRY(ctx, cudaFuncSetAttribute(conv_fused23_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_TOTAL));
  const int M = p.B * p.H * p.W;
  const int mt = (M + GM - 1) / GM, pairs = (mt + 1) / 2;
  int cap = ctx->opt_grid_cap > 0 ? std::min(num_sms, ctx->opt_grid_cap) : num_sms;
  cap = std::max(2, cap & ~1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)std::min(2 * pairs, cap));
  cfg.blockDim = dim3(G_THREADS);
  cfg.dynamicSmemBytes = G_TOTAL;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SUO_CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, conv_fused23_pair_kernel, p, mt));
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}
