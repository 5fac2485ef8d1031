// Kernel template of the fused HashMLP backward (decoder backward on the tensor cores + hash-table scatter); design notes
// in hashdecoder_bwd.cu.  Included by hashdecoder_bwd.cu (headline geometry L = 16, H = 64, compile-time activation, merge
// depth variants) and hashdecoder_bwd_geo.cu (the other F = 2 geometries, activation chosen at run time).
#pragma once
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "hash_device.cuh"
#include "mma_device.cuh"

namespace mri {
namespace {

constexpr int BWD_WARPS = 6;               // 192 threads, 2 blocks per SM (168 registers): 12 warps per SM
constexpr int BWD_THREADS = 32 * BWD_WARPS;

constexpr int ACT_RUNTIME = -1;  // ACT1 template value: GELU / ReLU chosen by the kernel argument `act1`
template <int ACT1>
__device__ __forceinline__ float hidden_act_only(float pre, int act1) {
  if constexpr (ACT1 >= 0) return activate<ACT1>(pre, 1.0f);
  else return act1 == MRI_ACT_GELU ? activate<MRI_ACT_GELU>(pre, 1.0f) : activate<MRI_ACT_RELU>(pre, 1.0f);
}
template <int ACT1>
__device__ __forceinline__ void hidden_act_and_grad(float pre, int act1, float& a, float& g) {
  if constexpr (ACT1 >= 0) {
    act_and_grad<ACT1>(pre, a, g);
  } else {
    if (act1 == MRI_ACT_GELU) act_and_grad<MRI_ACT_GELU>(pre, a, g);
    else act_and_grad<MRI_ACT_RELU>(pre, a, g);
  }
}

// K0 = 2 L real encoding columns; the tensor-core products run on KP = max(K0, 16) columns (zero padding for L = 4)
template <int K0, int H>
constexpr size_t fused_bwd_smem_bytes() {
  constexpr int KP = K0 < 16 ? 16 : K0;
  return 2 * (H * (KP + MMA_PAD) + K0 * (H + MMA_PAD)) * sizeof(__nv_bfloat16)  // W1 and W1^T planes
         + 2 * H * sizeof(float)                                               // b1, w2
         + BWD_WARPS * (2 * 16 * ((H + MMA_PAD) + (KP + MMA_PAD)) * sizeof(__nv_bfloat16)  // per-warp dPre1 / enc planes of one m-tile
                        + H * KP * sizeof(float));                                          // per-warp dW1 accumulator
}

// Decoder backward on the tensor cores with the table scatter fused in.  Per warp and per 16-coordinate m-tile:
//   (1) recompute pre1 = enc W1^T + b1 (mma.sync, as in the forward), form dPre1 = dPre2 w2 act1'(pre1) on the fragments
//   (2) dEnc = dPre1 W1: the accumulator fragments ARE the A fragments of the next mma (no data movement); all four
//       8-column blocks are issued back to back (four independent accumulator chains)
//   (3) dW1 += dPre1^T enc: both operands go through warp-private shared memory and come back transposed with
//       ldmatrix.trans; the 64 x 32 fp32 accumulator lives in warp-private shared memory in fragment order and passes
//       through registers only here - keeping it in registers for the whole loop (round 1) cost 64 of 253 registers and
//       held the kernel at 8 warps per SM, where ncu showed it latency-bound (issue slots 26-32 % busy)
//   (4) scatter: the dEnc fragments (rows g / g+8, level 4*nt2 + t for F = 2) never go to memory - lane pairs (t even /
//       odd) exchange their two levels with one shuffle and act as the lower / upper axis-0 halves of the pair-lane
//       scatter (hash_device.cuh); on the coarse levels duplicates along an axis-0 line are merged first.
// db1 / dw2 / db2 are per-thread column partials reduced once at the end.
//
// STEP = true turns the kernel into the WHOLE training step of HashMLP under the mean-squared-error loss
// (models.py:61-66 + 741-744 and their autograd) in one pass over the batch: the tile's encoding is gathered here (the
// forward kernel's gather, hashdecoder_fwd_impl.cuh) instead of being read back, the prediction and its loss gradient
// 2 (y - target) / n are formed on the spot - the mean's gradient needs no global reduction first - and the backward
// continues on the SAME accumulators: no (n, 32) encoding and no pre-activation ever go to memory, the hidden layer is
// not recomputed, and the gather's L2 loads overlap the scatter's L2 reductions (two different units of the L2) instead
// of running in two kernels back to back.
template <int D, int K0, int H, int ACT1, int MERGE_NT2, bool CONTIGUOUS, bool STEP = false>
__global__ void __launch_bounds__(BWD_THREADS, H == 64 ? 2 : 1) hashdecoder_mma_bwd_kernel(const float* __restrict__ enc, int64_t n,
                                                                           const float* __restrict__ w1, const float* __restrict__ b1,
                                                                           const float* __restrict__ w2, const float* __restrict__ pre2,
                                                                           const float* __restrict__ gy, int act1, int act2,
                                                                           const float* __restrict__ x, const __grid_constant__ LevelTable T,
                                                                           float* __restrict__ grad_tables, float* __restrict__ gw1,
                                                                           float* __restrict__ gb1, float* __restrict__ gw2,
                                                                           float* __restrict__ gb2,
                                                                           // STEP only: tables to gather from, regression targets, output bias,
                                                                           // 1 / n of the mean, loss accumulator, optional predictions
                                                                           const float* __restrict__ tables = nullptr,
                                                                           const float* __restrict__ target = nullptr,
                                                                           const float* __restrict__ b2 = nullptr, float inv_count = 0.0f,
                                                                           float* __restrict__ loss_out = nullptr,
                                                                           float* __restrict__ y_out = nullptr) {
  static_assert((K0 == 8 || K0 == 16 || K0 == 32) && H % 16 == 0, "F = 2, L = 4 / 8 / 16");
  constexpr int KP = K0 < 16 ? 16 : K0;
  constexpr int WS = KP + MMA_PAD;   // row stride of W1 / enc planes (bf16 elements)
  constexpr int TS = H + MMA_PAD;    // row stride of W1^T / dPre1 planes
  constexpr int STAGE = 2 * 16 * (TS + WS);  // bf16 elements of one warp's staging area
  extern __shared__ __align__(16) uint8_t msm[];
  __nv_bfloat16* w_hi = reinterpret_cast<__nv_bfloat16*>(msm);   // [H][WS]
  __nv_bfloat16* w_lo = w_hi + H * WS;
  __nv_bfloat16* wt_hi = w_lo + H * WS;                           // [K0][TS]  (W1 transposed)
  __nv_bfloat16* wt_lo = wt_hi + K0 * TS;
  float* b1s = reinterpret_cast<float*>(wt_lo + K0 * TS);
  float* w2s = b1s + H;
  float* wacc_all = w2s + H;                                      // [BWD_WARPS][H/16][KP/8][32 lanes][4]
  __nv_bfloat16* stage_all = reinterpret_cast<__nv_bfloat16*>(wacc_all + BWD_WARPS * H * KP);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  float4* wacc_s = reinterpret_cast<float4*>(wacc_all + warp * H * KP) + lane;   // + 32 * (jt * KP/8 + nt2)
  __nv_bfloat16* dp_hi = stage_all + warp * STAGE;                 // [16][TS]
  __nv_bfloat16* dp_lo = dp_hi + 16 * TS;
  __nv_bfloat16* e_hi = dp_lo + 16 * TS;                           // [16][WS]
  __nv_bfloat16* e_lo = e_hi + 16 * WS;

  stage_planes<H, K0, KP>(w1, w_hi, w_lo, false);
  stage_planes<H, K0>(w1, wt_hi, wt_lo, true);
  for (int e = threadIdx.x; e < H; e += blockDim.x) {
    b1s[e] = __ldg(b1 + e);
    w2s[e] = __ldg(w2 + e);
  }
  for (int e = threadIdx.x; e < BWD_WARPS * H * KP; e += blockDim.x) wacc_all[e] = 0.0f;
  // the lanes of one instruction work on two different levels: a lane-indexed read of the __grid_constant__ table is a
  // replayed LDC on the long scoreboard (17 % of the stall samples in ncu) - shared memory serves it in one pass
  __shared__ LevelDev lvs[K0 / 2];
  if (threadIdx.x < K0 / 2) lvs[threadIdx.x] = T.lv[threadIdx.x];
  __syncthreads();

  float pb1[H / 8][2], pw2[H / 8][2];
#pragma unroll
  for (int nt = 0; nt < H / 8; ++nt) { pb1[nt][0] = pb1[nt][1] = 0.0f; pw2[nt][0] = pw2[nt][1] = 0.0f; }
  float pb2 = 0.0f;
  float ploss = 0.0f;
  float b2v = 0.0f;
  if constexpr (STEP) b2v = __ldg(b2);

  // inputs of one 16-row m-tile as they come out of global memory; the next tile's are requested before the current
  // tile is processed
  struct TileIn {
    float2 e[KP / 16][2][2];  // [k-tile][8-column half][row g / g+8]
    float gy[2], p2[2];  // STEP: gy holds the regression targets
    float xv[2][D];
  };
  auto fetch = [&](int64_t row0, TileIn& ti) {
    const int64_t r[2] = {row0 + g, row0 + g + 8};
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const bool live = r[rr] < n;
      load_coord<D>(x, live ? r[rr] : 0, ti.xv[rr]);
      if constexpr (STEP) {
        ti.gy[rr] = live ? __ldg(target + r[rr]) : 0.0f;
      } else {
        ti.gy[rr] = live ? __ldg(gy + r[rr]) : 0.0f;
        ti.p2[rr] = live ? __ldg(pre2 + r[rr]) : 0.0f;
#pragma unroll
        for (int kt = 0; kt < KP / 16; ++kt)
#pragma unroll
          for (int half = 0; half < 2; ++half)
            ti.e[kt][half][rr] = (live && 16 * kt + 8 * half < K0)  // columns K0 .. KP-1 are padding (L = 4)
                                     ? __ldg(reinterpret_cast<const float2*>(enc + r[rr] * K0 + 16 * kt + 8 * half + 2 * t))
                                     : make_float2(0.0f, 0.0f);
      }
    }
  };

  // Every warp walks its OWN contiguous range of m-tiles.  With a locality-ordered batch a grid-stride walk would make
  // all resident warps work on neighbouring samples at the same time, i.e. reduce into the same few rows of the coarse
  // levels at once - and the L2 serialises reductions on one address (measured: 2.2x slower when 8192 rows are hot
  // grid-wide).  Warps that are far apart in the batch are far apart in the volume.
  const int64_t tiles = (n + 15) / 16;
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * BWD_WARPS;
  const int64_t wid = static_cast<int64_t>(blockIdx.x) * BWD_WARPS + warp;
  const int64_t tile_end = CONTIGUOUS ? ((wid + 1) * tiles) / n_warps : tiles;
  const int64_t tile_stride = CONTIGUOUS ? 1 : n_warps;
  int64_t tile = CONTIGUOUS ? (wid * tiles) / n_warps : wid;
  TileIn nxt;
  fetch(tile * 16, nxt);
#pragma unroll 1
  for (; tile < tile_end; tile += tile_stride) {
    const int64_t row0 = tile * 16;
    const int64_t r_lo = row0 + g, r_hi = row0 + g + 8;
    const TileIn cur = nxt;
    fetch((tile + tile_stride) * 16, nxt);
    float xv_lo[D], xv_hi[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { xv_lo[d] = cur.xv[0][d]; xv_hi[d] = cur.xv[1][d]; }
    // rows g, g+1 of an 8-row group live and on one axis-0 line (identical coordinates on every other axis)
    uint32_t line_mask[2] = {0u, 0u};
    if constexpr (MERGE_NT2 > 0) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        bool same = g < 7 && (rr == 0 ? r_lo : r_hi) + 1 < n;
#pragma unroll
        for (int d = 1; d < D; ++d) {
          const float mine = rr == 0 ? xv_lo[d] : xv_hi[d];
          const float next = __shfl_down_sync(0xffffffffu, mine, 4);  // every lane takes part: no short-circuit around it
          same = same && next == mine;
        }
        line_mask[rr] = __ballot_sync(0xffffffffu, same);
      }
    }
    float acc[H / 8][4];
    {
      uint32_t a_hi[KP / 16][4], a_lo[KP / 16][4];
      if constexpr (STEP) {
        // the forward kernel's gather: lanes t, t^1 are the axis-0 pair, each walks its half of the corners of the pair's two
        // levels for both rows; one shuffle completes the level sums and leaves level 4q + t of rows g / g+8 in this lane
        if constexpr (K0 < KP) a_hi[0][2] = a_hi[0][3] = a_lo[0][2] = a_lo[0][3] = 0u;
        const int b0g = t & 1;
#pragma unroll
        for (int q = 0; q < K0 / 8; ++q) {
          Feat<2> part[2][2];
#pragma unroll
          for (int which = 0; which < 2; ++which) {
            const LevelDev lv = lvs[4 * q + (t & 2) + which];
            encode_half_level_rows<D>(make_cell<D>(xv_lo, lv), make_cell<D>(xv_hi, lv), b0g, lv, tables + lv.offset, part[which][0],
                                      part[which][1]);
          }
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            float full[2];
#pragma unroll
            for (int f = 0; f < 2; ++f) {
              const float mine = b0g ? part[1][rr].v[f] : part[0][rr].v[f];
              const float send = b0g ? part[0][rr].v[f] : part[1][rr].v[f];
              full[f] = mine + __shfl_xor_sync(0xffffffffu, send, 1);
            }
            split_pair(full[0], full[1], a_hi[q >> 1][2 * (q & 1) + rr], a_lo[q >> 1][2 * (q & 1) + rr]);
          }
        }
      } else {
#pragma unroll
        for (int kt = 0; kt < KP / 16; ++kt)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            split_pair(cur.e[kt][half][0].x, cur.e[kt][half][0].y, a_hi[kt][2 * half + 0], a_lo[kt][2 * half + 0]);
            split_pair(cur.e[kt][half][1].x, cur.e[kt][half][1].y, a_hi[kt][2 * half + 1], a_lo[kt][2 * half + 1]);
          }
      }
#pragma unroll
      for (int kt = 0; kt < KP / 16; ++kt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int col = 16 * kt + 8 * half + 2 * t;
          *reinterpret_cast<uint32_t*>(e_hi + g * WS + col) = a_hi[kt][2 * half];
          *reinterpret_cast<uint32_t*>(e_lo + g * WS + col) = a_lo[kt][2 * half];
          *reinterpret_cast<uint32_t*>(e_hi + (g + 8) * WS + col) = a_hi[kt][2 * half + 1];
          *reinterpret_cast<uint32_t*>(e_lo + (g + 8) * WS + col) = a_lo[kt][2 * half + 1];
        }
      hidden_mma<KP, H>(a_hi, a_lo, w_hi, w_lo, b1s, g, t, acc);
    }
    float dp2_lo = 0.0f, dp2_hi = 0.0f;
    if constexpr (STEP) {
      // prediction of rows g / g+8 (decoder layer 2 on the accumulator fragments, quad reduction), loss and its gradient
      float s_lo = 0.0f, s_hi = 0.0f;
#pragma unroll
      for (int nt = 0; nt < H / 8; ++nt) {
        const float wl = w2s[8 * nt + 2 * t], wh = w2s[8 * nt + 2 * t + 1];
        s_lo = fmaf(hidden_act_only<ACT1>(acc[nt][0], act1), wl, s_lo);
        s_lo = fmaf(hidden_act_only<ACT1>(acc[nt][1], act1), wh, s_lo);
        s_hi = fmaf(hidden_act_only<ACT1>(acc[nt][2], act1), wl, s_hi);
        s_hi = fmaf(hidden_act_only<ACT1>(acc[nt][3], act1), wh, s_hi);
      }
      s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 1); s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 2);
      s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 1); s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 2);
      const float p_lo = s_lo + b2v, p_hi = s_hi + b2v;
      const float y_lo = activate_rt(act2, p_lo, 1.0f), y_hi = activate_rt(act2, p_hi, 1.0f);
      const float e_lo_ = y_lo - cur.gy[0], e_hi_ = y_hi - cur.gy[1];
      if (r_lo < n) {
        dp2_lo = 2.0f * e_lo_ * inv_count * activate_grad_rt(act2, p_lo, 1.0f);
        if (t == 0) { ploss = fmaf(e_lo_, e_lo_, ploss); if (y_out) y_out[r_lo] = y_lo; }
      }
      if (r_hi < n) {
        dp2_hi = 2.0f * e_hi_ * inv_count * activate_grad_rt(act2, p_hi, 1.0f);
        if (t == 0) { ploss = fmaf(e_hi_, e_hi_, ploss); if (y_out) y_out[r_hi] = y_hi; }
      }
    } else {
      if (r_lo < n) dp2_lo = cur.gy[0] * activate_grad_rt(act2, cur.p2[0], 1.0f);
      if (r_hi < n) dp2_hi = cur.gy[1] * activate_grad_rt(act2, cur.p2[1], 1.0f);
    }
    if (t == 0) pb2 += dp2_lo + dp2_hi;
    float dacc[K0 / 8][4];
    {
      uint32_t da_hi[H / 16][4], da_lo[H / 16][4];
#pragma unroll
      for (int nt = 0; nt < H / 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float dp2 = e < 2 ? dp2_lo : dp2_hi;
          float a, gp;
          hidden_act_and_grad<ACT1>(acc[nt][e], act1, a, gp);
          pw2[nt][e & 1] = fmaf(dp2, a, pw2[nt][e & 1]);
          const float dpre = dp2 * w2s[8 * nt + 2 * t + (e & 1)] * gp;
          pb1[nt][e & 1] += dpre;
          acc[nt][e] = dpre;
        }
        uint32_t h0, l0, h1, l1;
        split_pair(acc[nt][0], acc[nt][1], h0, l0);  // row g
        split_pair(acc[nt][2], acc[nt][3], h1, l1);  // row g + 8
        const int col = 8 * nt + 2 * t;
        *reinterpret_cast<uint32_t*>(dp_hi + g * TS + col) = h0;
        *reinterpret_cast<uint32_t*>(dp_lo + g * TS + col) = l0;
        *reinterpret_cast<uint32_t*>(dp_hi + (g + 8) * TS + col) = h1;
        *reinterpret_cast<uint32_t*>(dp_lo + (g + 8) * TS + col) = l1;
        // accumulator fragment -> A fragment of the dEnc product (k-tile nt/2, halves by nt parity)
        da_hi[nt / 2][2 * (nt & 1) + 0] = h0; da_hi[nt / 2][2 * (nt & 1) + 1] = h1;
        da_lo[nt / 2][2 * (nt & 1) + 0] = l0; da_lo[nt / 2][2 * (nt & 1) + 1] = l1;
      }
      // dEnc (16 x K0) = dPre1 (16 x H) . W1 (H x K0); B[k = j][n = kenc] = W1^T planes [kenc][j]
#pragma unroll
      for (int nt2 = 0; nt2 < K0 / 8; ++nt2) {
        dacc[nt2][0] = dacc[nt2][1] = dacc[nt2][2] = dacc[nt2][3] = 0.0f;
#pragma unroll
        for (int kt2 = 0; kt2 < H / 16; ++kt2) {
          const int off = (8 * nt2 + g) * TS + 16 * kt2 + 2 * t;
          const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(wt_hi + off), bh1 = *reinterpret_cast<const uint32_t*>(wt_hi + off + 8);
          const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(wt_lo + off), bl1 = *reinterpret_cast<const uint32_t*>(wt_lo + off + 8);
          mma_bf16_16816(dacc[nt2], da_lo[kt2], bh0, bh1);
          mma_bf16_16816(dacc[nt2], da_hi[kt2], bl0, bl1);
          mma_bf16_16816(dacc[nt2], da_hi[kt2], bh0, bh1);
        }
      }
    }
    __syncwarp();
    // dW1 (H x K0) += dPre1^T (H x 16) . enc (16 x K0), operands transposed on the way out of shared memory
    {
      const int lm = lane >> 3, lr = lane & 7;
      uint32_t bh[KP / 16][4], bl[KP / 16][4];  // per pair of 8-column blocks: {b0, b1} of the first, {b0, b1} of the second
#pragma unroll
      for (int np = 0; np < KP / 16; ++np) {
        const int off = (8 * (lm & 1) + lr) * WS + 8 * (2 * np + (lm >> 1));
        ldmatrix_x4_trans(bh[np], e_hi + off);
        ldmatrix_x4_trans(bl[np], e_lo + off);
      }
#pragma unroll
      for (int jt = 0; jt < H / 16; ++jt) {
        uint32_t ah[4], al[4];
        const int off = (8 * (lm >> 1) + lr) * TS + 16 * jt + 8 * (lm & 1);
        ldmatrix_x4_trans(ah, dp_hi + off);
        ldmatrix_x4_trans(al, dp_lo + off);
#pragma unroll
        for (int nt2 = 0; nt2 < KP / 8; ++nt2) {
          float4 c4 = wacc_s[32 * (jt * (KP / 8) + nt2)];
          float c[4] = {c4.x, c4.y, c4.z, c4.w};
          const uint32_t h0 = bh[nt2 >> 1][2 * (nt2 & 1)], h1 = bh[nt2 >> 1][2 * (nt2 & 1) + 1];
          const uint32_t l0 = bl[nt2 >> 1][2 * (nt2 & 1)], l1 = bl[nt2 >> 1][2 * (nt2 & 1) + 1];
          mma_bf16_16816(c, al, h0, h1);
          mma_bf16_16816(c, ah, l0, l1);
          mma_bf16_16816(c, ah, h0, h1);
          wacc_s[32 * (jt * (KP / 8) + nt2)] = make_float4(c[0], c[1], c[2], c[3]);
        }
      }
    }
    __syncwarp();
    // fused scatter: this lane holds dEnc of level 4*nt2 + t, its pair partner (t ^ 1) the neighbouring level.  The
    // (level of the pair, row of the tile) combinations run as a real loop: fully unrolled the scatter alone was ~60 KB of
    // SASS and the warps stalled on instruction fetch (ncu: stall_no_instruction 1.3 per issue with 12 warps per SM)
    const int b0 = t & 1;
#pragma unroll
    for (int nt2 = 0; nt2 < K0 / 8; ++nt2) {
      float pv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) pv[e] = __shfl_xor_sync(0xffffffffu, dacc[nt2][e], 1);
#pragma unroll 1
      for (int wr = 0; wr < 4; ++wr) {
        const int which = wr >> 1, rr = wr & 1;    // which: the pair's even / odd level; rr: rows g / g + 8 of the m-tile
        const bool mine = (which == b0);
        const LevelDev lv = lvs[4 * nt2 + (t & ~1) + which];
        float* tbl = grad_tables + lv.offset;
        const bool live = (rr ? r_hi : r_lo) < n;
        const float g0 = mine ? (rr ? dacc[nt2][2] : dacc[nt2][0]) : (rr ? pv[2] : pv[0]);
        const float g1 = mine ? (rr ? dacc[nt2][3] : dacc[nt2][1]) : (rr ? pv[3] : pv[1]);
        float xr[D];
#pragma unroll
        for (int d = 0; d < D; ++d) xr[d] = rr ? xv_hi[d] : xv_lo[d];
        const Cell<D> cell = make_cell<D>(xr, lv);
        const float w0 = b0 ? cell.wu[0] : cell.wl[0];
        float v0 = g0 * w0, v1 = g1 * w0;
        bool active = live;
        if (nt2 < MERGE_NT2) {
          // coarse levels of a locality-ordered batch: duplicates along the axis-0 line are summed in registers
          const MergedHalf mh = merge_line_runs(cell.lo[0] + static_cast<uint32_t>(b0), v0, v1, live, rr ? line_mask[1] : line_mask[0], lane);
          v0 = mh.v0; v1 = mh.v1; active = mh.active;
        }
        if (active) scatter_half_level_weighted<D>(cell, b0, lv, tbl, v0, v1);
      }
    }
  }

  // ---- flush ----
  __syncthreads();
  // dW1: sum of the warps' accumulators; element (j, k) sits at fragment position (jt, nt2, lane', c)
  for (int e = threadIdx.x; e < H * K0; e += blockDim.x) {
    const int j = e / K0, k = e - j * K0;
    const int jt = j >> 4, jr = j & 15, nt2 = k >> 3, kc = k & 7;
    const int idx = ((jt * (KP / 8) + nt2) * 32 + 4 * (jr & 7) + (kc >> 1)) * 4 + 2 * (jr >> 3) + (kc & 1);
    float sum = 0.0f;
#pragma unroll
    for (int w = 0; w < BWD_WARPS; ++w) sum += wacc_all[w * H * KP + idx];
    red_add_f32(gw1 + e, sum);
  }
  // column partials: sum over the 8 row groups (lanes with equal t), then one atomic per warp and column
#pragma unroll
  for (int nt = 0; nt < H / 8; ++nt)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float vb = pb1[nt][q], vw = pw2[nt][q];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        vb += __shfl_xor_sync(0xffffffffu, vb, o);
        vw += __shfl_xor_sync(0xffffffffu, vw, o);
      }
      if (g == 0) {
        red_add_f32(gb1 + 8 * nt + 2 * t + q, vb);
        red_add_f32(gw2 + 8 * nt + 2 * t + q, vw);
      }
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pb2 += __shfl_xor_sync(0xffffffffu, pb2, o);
  if (lane == 0) red_add_f32(gb2, pb2);
  if constexpr (STEP) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ploss += __shfl_xor_sync(0xffffffffu, ploss, o);
    if (lane == 0) red_add_f32(loss_out, ploss * inv_count);
  }
}

template <int D, int K0, int H, int ACT1, int MERGE_NT2, bool CONTIGUOUS>
int launch_fused_bwd(const float* enc, int64_t n, const float* w1, const float* b1, const float* w2, const float* pre2,
                     const float* gy, int act1, int act2, const float* x, const LevelTable& T, float* grad_tables, float* gw1, float* gb1,
                     float* gw2, float* gb2, cudaStream_t s) {
  constexpr size_t smem = fused_bwd_smem_bytes<K0, H>();
  auto kernel = hashdecoder_mma_bwd_kernel<D, K0, H, ACT1, MERGE_NT2, CONTIGUOUS>;
  static DeviceCache resident_cache;  // resident blocks of this kernel on the current device (one persistent wave)
  const int dev = DeviceCache::device();
  int resident = resident_cache.slot[dev].load(std::memory_order_acquire);
  if (resident == 0) {
    MRI_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 0;
    MRI_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, BWD_THREADS, smem));
    resident = (per_sm > 0 ? per_sm : 1) * sm_count();
    resident_cache.slot[dev].store(resident, std::memory_order_release);
  }
  int64_t blocks = ((n + 15) / 16 + BWD_WARPS - 1) / BWD_WARPS;
  if (blocks > resident) blocks = resident;
  kernel<<<static_cast<int>(blocks), BWD_THREADS, smem, s>>>(enc, n, w1, b1, w2, pre2, gy, act1, act2, x, T, grad_tables, gw1, gb1,
                                                             gw2, gb2, nullptr, nullptr, nullptr, 0.0f, nullptr, nullptr);
  MRI_LAUNCH_OK("hashdecoder_mma_bwd_kernel");
  return MRI_OK;
}


// the whole HashMLP + MSE training step (STEP = true): same launch shape as the backward
template <int D, int K0, int H, int ACT1, int MERGE_NT2>
int launch_fused_step(const float* x, const float* target, int64_t n, const float* tables, const LevelTable& T, const float* w1,
                      const float* b1, const float* w2, const float* b2, int act1, int act2, float inv_count, float* grad_tables,
                      float* gw1, float* gb1, float* gw2, float* gb2, float* loss_out, float* y_out, cudaStream_t s) {
  constexpr size_t smem = fused_bwd_smem_bytes<K0, H>();
  auto kernel = hashdecoder_mma_bwd_kernel<D, K0, H, ACT1, MERGE_NT2, true, true>;
  static DeviceCache resident_cache;
  const int dev = DeviceCache::device();
  int resident = resident_cache.slot[dev].load(std::memory_order_acquire);
  if (resident == 0) {
    MRI_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 0;
    MRI_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, BWD_THREADS, smem));
    resident = (per_sm > 0 ? per_sm : 1) * sm_count();
    resident_cache.slot[dev].store(resident, std::memory_order_release);
  }
  int64_t blocks = ((n + 15) / 16 + BWD_WARPS - 1) / BWD_WARPS;
  if (blocks > resident) blocks = resident;
  kernel<<<static_cast<int>(blocks), BWD_THREADS, smem, s>>>(nullptr, n, w1, b1, w2, nullptr, nullptr, act1, act2, x, T, grad_tables, gw1,
                                                             gb1, gw2, gb2, tables, target, b2, inv_count, loss_out, y_out);
  MRI_LAUNCH_OK("hashdecoder_mma_bwd_kernel<step>");
  return MRI_OK;
}

}  // namespace
}  // namespace mri
