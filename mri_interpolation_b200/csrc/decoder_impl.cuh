// Kernel templates of the stand-alone fused 2-layer decoder (design notes in decoder.cu).  Instantiated per input width
// K0 in decoder_k16.cu / decoder_k32.cu / decoder_k64.cu (16 kernels each) so that the three files compile in parallel -
// as one translation unit the 48 kernels took 3.2 of the library's 4 build minutes.
#pragma once
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "hash_device.cuh"
#include "mma_device.cuh"

namespace mri {
namespace {

constexpr int DP_PAD = 4;  // dPre1 rows are H + 4 floats apart: conflict-free float4 stores, still 16-byte aligned


// hidden pre-activations of one coordinate: h[j] = b1[j] + sum_k enc[k] W1[j][k]  (W1t = W1 transposed in smem)
// the whole encoding row of a coordinate: K0/4 independent 16-byte loads issued back-to-back (one latency, not K0/4)
template <int K0>
__device__ __forceinline__ void load_row(const float* __restrict__ enc_row, float4 (&e)[K0 / 4]) {
#pragma unroll
  for (int q = 0; q < K0 / 4; ++q) e[q] = __ldg(reinterpret_cast<const float4*>(enc_row) + q);
}

template <int K0, int H>
__device__ __forceinline__ void hidden_pre(const float4 (&row)[K0 / 4], const float* __restrict__ w1t,
                                           const float* __restrict__ b1s, float (&h)[H]) {
#pragma unroll
  for (int j = 0; j < H; ++j) h[j] = b1s[j];
#pragma unroll
  for (int k4 = 0; k4 < K0 / 4; ++k4) {
    const float e[4] = {row[k4].x, row[k4].y, row[k4].z, row[k4].w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4* row = reinterpret_cast<const float4*>(w1t + (4 * k4 + u) * H);
#pragma unroll
      for (int q = 0; q < H / 4; ++q) {
        const float4 w = row[q];
        h[4 * q + 0] = fmaf(e[u], w.x, h[4 * q + 0]);
        h[4 * q + 1] = fmaf(e[u], w.y, h[4 * q + 1]);
        h[4 * q + 2] = fmaf(e[u], w.z, h[4 * q + 2]);
        h[4 * q + 3] = fmaf(e[u], w.w, h[4 * q + 3]);
      }
    }
  }
}

template <int K0, int H, int ACT1>
__global__ void __launch_bounds__(DEC_THREADS, 3) decoder2_fwd_kernel(const float* __restrict__ enc, int64_t n,
                                                                    const float* __restrict__ w1, const float* __restrict__ b1,
                                                                    const float* __restrict__ w2, const float* __restrict__ b2,
                                                                    int act2, float* __restrict__ y, float* __restrict__ pre2_out) {
  __shared__ __align__(16) float w1t[K0 * H];
  __shared__ __align__(16) float b1s[H];
  __shared__ __align__(16) float w2s[H];
  for (int e = threadIdx.x; e < K0 * H; e += DEC_THREADS) {
    const int j = e / K0, k = e - j * K0;
    w1t[k * H + j] = __ldg(w1 + e);
  }
  for (int e = threadIdx.x; e < H; e += DEC_THREADS) {
    b1s[e] = __ldg(b1 + e);
    w2s[e] = __ldg(w2 + e);
  }
  const float b2v = __ldg(b2);
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * DEC_THREADS;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * DEC_THREADS + threadIdx.x; i < n; i += stride) {
    float4 row[K0 / 4];
    load_row<K0>(enc + i * K0, row);
    float h[H];
    hidden_pre<K0, H>(row, w1t, b1s, h);
    float pre2 = b2v;
#pragma unroll
    for (int j = 0; j < H; ++j) pre2 = fmaf(activate<ACT1>(h[j], 1.0f), w2s[j], pre2);
    y[i] = activate_rt(act2, pre2, 1.0f);
    if (pre2_out) pre2_out[i] = pre2;
  }
}

template <int K0, int H, int ACT1>
__global__ void __launch_bounds__(DEC_THREADS, 2) decoder2_bwd_kernel(const float* __restrict__ enc, int64_t n,
                                                                    const float* __restrict__ w1, const float* __restrict__ b1,
                                                                    const float* __restrict__ w2, const float* __restrict__ pre2,
                                                                    const float* __restrict__ gy, int act2, float* __restrict__ denc,
                                                                    float* __restrict__ gw1, float* __restrict__ gb1,
                                                                    float* __restrict__ gw2, float* __restrict__ gb2) {
  constexpr int JG = DEC_THREADS / K0;  // threads along the hidden axis in the dW1 phase
  constexpr int JPT = H / JG;           // hidden units per thread in the dW1 phase
  constexpr int DPS = H + DP_PAD;       // row stride of the per-coordinate staging arrays
  static_assert(DEC_THREADS % K0 == 0 && H % JG == 0 && JPT % 4 == 0 && H <= DEC_THREADS, "unsupported decoder shape");
  extern __shared__ __align__(16) float dsm[];
  float* w1t = dsm;                        // [K0][H]    for the hidden-layer recompute
  float* w1s = w1t + K0 * H;               // [H][K0]    for dEnc = W1^T dPre1
  float* b1s = w1s + K0 * H;               // [H]
  float* w2s = b1s + H;                    // [H]
  float* dpre_s = w2s + H;                 // [128][H+4] dPre1 of the current chunk
  float* adp_s = dpre_s + DEC_THREADS * DPS;  // [128][H+4] act1(pre1) * dPre2 of the current chunk (for dw2)
  float* enc_t = adp_s + DEC_THREADS * DPS;   // [K0][129]  enc of the current chunk, transposed, padded
  for (int e = threadIdx.x; e < K0 * H; e += DEC_THREADS) {
    const float v = __ldg(w1 + e);
    const int j = e / K0, k = e - j * K0;
    w1s[e] = v;
    w1t[k * H + j] = v;
  }
  for (int e = threadIdx.x; e < H; e += DEC_THREADS) {
    b1s[e] = __ldg(b1 + e);
    w2s[e] = __ldg(w2 + e);
  }
  __syncthreads();

  float acc_w1[JPT];
#pragma unroll
  for (int t = 0; t < JPT; ++t) acc_w1[t] = 0.0f;
  float acc_b1 = 0.0f, acc_w2 = 0.0f, acc_b2 = 0.0f;  // column sums owned by threads < H (b2: every thread)
  const int my_k = threadIdx.x % K0;
  const int my_j0 = (threadIdx.x / K0) * JPT;

  const int64_t chunks = (n + DEC_THREADS - 1) / DEC_THREADS;
  for (int64_t chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x) {
    const int64_t i = chunk * DEC_THREADS + threadIdx.x;
    const bool live = i < n;
    // ---- phase A: one coordinate per thread ----
    {
      float h[H];
      float dp2 = 0.0f;
      float4 row[K0 / 4];
#pragma unroll
      for (int q = 0; q < K0 / 4; ++q) row[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        load_row<K0>(enc + i * K0, row);
        hidden_pre<K0, H>(row, w1t, b1s, h);
        dp2 = __ldg(gy + i) * activate_grad_rt(act2, __ldg(pre2 + i), 1.0f);
        acc_b2 += dp2;
      } else {
#pragma unroll
        for (int j = 0; j < H; ++j) h[j] = 0.0f;
      }
      float4* dp_row = reinterpret_cast<float4*>(dpre_s + threadIdx.x * DPS);
      float4* ad_row = reinterpret_cast<float4*>(adp_s + threadIdx.x * DPS);
#pragma unroll
      for (int q = 0; q < H / 4; ++q) {
        float a[4], g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          act_and_grad<ACT1>(h[4 * q + u], a[u], g[u]);
          h[4 * q + u] = dp2 * w2s[4 * q + u] * g[u];  // dPre1
          a[u] *= dp2;
        }
        dp_row[q] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
        ad_row[q] = make_float4(a[0], a[1], a[2], a[3]);
      }
      // dEnc[k] = sum_j dPre1[j] W1[j][k], four k at a time; park enc (transposed) for the dW1 phase
#pragma unroll
      for (int q = 0; q < K0 / 4; ++q) {
        const float4 e4 = row[q];
        if (live) {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j = 0; j < H; ++j) {
            const float4 w = reinterpret_cast<const float4*>(w1s + j * K0)[q];
            acc.x = fmaf(h[j], w.x, acc.x); acc.y = fmaf(h[j], w.y, acc.y);
            acc.z = fmaf(h[j], w.z, acc.z); acc.w = fmaf(h[j], w.w, acc.w);
          }
          reinterpret_cast<float4*>(denc + i * K0)[q] = acc;
        }
        enc_t[(4 * q + 0) * (DEC_THREADS + 1) + threadIdx.x] = e4.x;
        enc_t[(4 * q + 1) * (DEC_THREADS + 1) + threadIdx.x] = e4.y;
        enc_t[(4 * q + 2) * (DEC_THREADS + 1) + threadIdx.x] = e4.z;
        enc_t[(4 * q + 3) * (DEC_THREADS + 1) + threadIdx.x] = e4.w;
      }
    }
    __syncthreads();
    // ---- phase B: dW1[j][k] += sum_c dPre1[c][j] enc[c][k]; thread owns (k = my_k, j in [my_j0, +JPT)) ----
#pragma unroll 2
    for (int c = 0; c < DEC_THREADS; ++c) {
      const float ek = enc_t[my_k * (DEC_THREADS + 1) + c];
      const float4* dp = reinterpret_cast<const float4*>(dpre_s + c * DPS + my_j0);
#pragma unroll
      for (int q = 0; q < JPT / 4; ++q) {
        const float4 d = dp[q];
        acc_w1[4 * q + 0] = fmaf(d.x, ek, acc_w1[4 * q + 0]);
        acc_w1[4 * q + 1] = fmaf(d.y, ek, acc_w1[4 * q + 1]);
        acc_w1[4 * q + 2] = fmaf(d.z, ek, acc_w1[4 * q + 2]);
        acc_w1[4 * q + 3] = fmaf(d.w, ek, acc_w1[4 * q + 3]);
      }
    }
    // column sums: db1[j] += sum_c dPre1[c][j], dw2[j] += sum_c act1(pre1[c][j]) dPre2[c]
    if (threadIdx.x < H) {
#pragma unroll 8
      for (int c = 0; c < DEC_THREADS; ++c) {
        acc_b1 += dpre_s[c * DPS + threadIdx.x];
        acc_w2 += adp_s[c * DPS + threadIdx.x];
      }
    }
    __syncthreads();
  }

  // ---- flush: one round of atomics per block ----
#pragma unroll
  for (int t = 0; t < JPT; ++t) red_add_f32(gw1 + (my_j0 + t) * K0 + my_k, acc_w1[t]);
  if (threadIdx.x < H) {
    red_add_f32(gb1 + threadIdx.x, acc_b1);
    red_add_f32(gw2 + threadIdx.x, acc_w2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc_b2 += __shfl_xor_sync(0xffffffffu, acc_b2, o);
  if ((threadIdx.x & 31) == 0) red_add_f32(gb2, acc_b2);
}

// ================================ tensor-core (mma.sync) variant ====================================
// The hidden layer enc(K0) -> H is a (32 coords x K0) x (K0 x H) product per warp: tiny for tcgen05/TMEM tiles, but
// a good fit for warp-level mma.sync.m16n8k16 (bf16 inputs, fp32 accumulate) with the same split-precision trick
// as the SIREN path (x = hi + lo; A_lo*B_hi + A_hi*B_lo + A_hi*B_hi), which keeps fp32 parity.  Fragments come
// straight from global memory (A: float2 loads in the fragment's own (row g, cols 2t) pattern) and from a padded
// bf16 copy of W1 in shared memory (B); bias, GELU, the H -> 1 output layer and its quad reduction happen on the
// accumulator fragments in registers.  ~3x fewer issued instructions than the CUDA-core kernel above.
template <int K0, int H, int ACT1>
__global__ void __launch_bounds__(DEC_THREADS, 3) decoder2_mma_fwd_kernel(const float* __restrict__ enc, int64_t n,
                                                                           const float* __restrict__ w1, const float* __restrict__ b1,
                                                                           const float* __restrict__ w2, const float* __restrict__ b2,
                                                                           int act2, float* __restrict__ y, float* __restrict__ pre2_out) {
  constexpr int WS = K0 + MMA_PAD;
  __shared__ __align__(16) __nv_bfloat16 w_hi[H * WS];
  __shared__ __align__(16) __nv_bfloat16 w_lo[H * WS];
  __shared__ float b1s[H];
  __shared__ float w2s[H];
  stage_planes<H, K0>(w1, w_hi, w_lo, false);
  for (int e = threadIdx.x; e < H; e += DEC_THREADS) {
    b1s[e] = __ldg(b1 + e);
    w2s[e] = __ldg(w2 + e);
  }
  const float b2v = __ldg(b2);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t tiles = (n + 15) / 16;  // 16-coordinate m-tiles, one per warp iteration
  for (int64_t tile = static_cast<int64_t>(blockIdx.x) * (DEC_THREADS / 32) + warp; tile < tiles;
       tile += static_cast<int64_t>(gridDim.x) * (DEC_THREADS / 32)) {
    const int64_t row0 = tile * 16;
    uint32_t a_hi[K0 / 16][4], a_lo[K0 / 16][4];
    load_a_frags<K0>(enc, row0, n, g, t, a_hi, a_lo);
    float acc[H / 8][4];
    hidden_mma<K0, H>(a_hi, a_lo, w_hi, w_lo, b1s, g, t, acc);
    // output layer on the fragments: this thread owns rows (g, g+8) x columns {8nt+2t, 8nt+2t+1}
    float s_lo = 0.0f, s_hi = 0.0f;
#pragma unroll
    for (int nt = 0; nt < H / 8; ++nt) {
      const float wl = w2s[8 * nt + 2 * t], wh = w2s[8 * nt + 2 * t + 1];
      s_lo = fmaf(activate<ACT1>(acc[nt][0], 1.0f), wl, s_lo);
      s_lo = fmaf(activate<ACT1>(acc[nt][1], 1.0f), wh, s_lo);
      s_hi = fmaf(activate<ACT1>(acc[nt][2], 1.0f), wl, s_hi);
      s_hi = fmaf(activate<ACT1>(acc[nt][3], 1.0f), wh, s_hi);
    }
    s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 1); s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 2);
    s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 1); s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 2);
    if (t == 0) {
      const int64_t r0 = row0 + g, r1 = row0 + g + 8;
      if (r0 < n) { const float p = s_lo + b2v; y[r0] = activate_rt(act2, p, 1.0f); if (pre2_out) pre2_out[r0] = p; }
      if (r1 < n) { const float p = s_hi + b2v; y[r1] = activate_rt(act2, p, 1.0f); if (pre2_out) pre2_out[r1] = p; }
    }
  }
}

template <int K0, int H, int ACT1>
int launch_mma_fwd(const float* enc, int64_t n, const float* w1, const float* b1, const float* w2, const float* b2, int act2,
                   float* y, float* pre2, cudaStream_t s) {
  int64_t blocks = ((n + 15) / 16 + 3) / 4;
  const int64_t cap = 6LL * sm_count();
  if (blocks > cap) blocks = cap;
  decoder2_mma_fwd_kernel<K0, H, ACT1><<<static_cast<int>(blocks), DEC_THREADS, 0, s>>>(enc, n, w1, b1, w2, b2, act2, y, pre2);
  MRI_LAUNCH_OK("decoder2_mma_fwd_kernel");
  return MRI_OK;
}


// Backward on the tensor cores.  Per warp and per 32 coordinates:
//   (1) recompute pre1 = enc W1^T + b1 (mma, as in the forward), form dPre1 = dPre2 w2 act1'(pre1) on the fragments
//   (2) dEnc = dPre1 W1: the accumulator fragments ARE the A fragments of the next mma (no data movement)
//   (3) dW1 += dPre1^T enc: both operands go through warp-private shared memory and come back transposed with
//       ldmatrix.trans; the 64 x 32 accumulator stays in registers for the whole persistent loop
// db1 / dw2 / db2 are per-thread column partials reduced once at the end.
template <int K0, int H, int ACT1>
__global__ void __launch_bounds__(DEC_THREADS, 2) decoder2_mma_bwd_kernel(const float* __restrict__ enc, int64_t n,
                                                                           const float* __restrict__ w1, const float* __restrict__ b1,
                                                                           const float* __restrict__ w2, const float* __restrict__ pre2,
                                                                           const float* __restrict__ gy, int act2,
                                                                           float* __restrict__ denc, float* __restrict__ gw1,
                                                                           float* __restrict__ gb1, float* __restrict__ gw2,
                                                                           float* __restrict__ gb2) {
  constexpr int WS = K0 + MMA_PAD;   // row stride of W1 / enc planes (bf16 elements)
  constexpr int TS = H + MMA_PAD;    // row stride of W1^T / dPre1 planes
  constexpr int NWARP = DEC_THREADS / 32;
  extern __shared__ __align__(16) uint8_t msm[];
  __nv_bfloat16* w_hi = reinterpret_cast<__nv_bfloat16*>(msm);   // [H][WS]
  __nv_bfloat16* w_lo = w_hi + H * WS;
  __nv_bfloat16* wt_hi = w_lo + H * WS;                           // [K0][TS]  (W1 transposed)
  __nv_bfloat16* wt_lo = wt_hi + K0 * TS;
  float* b1s = reinterpret_cast<float*>(wt_lo + K0 * TS);
  float* w2s = b1s + H;
  __nv_bfloat16* warp_base = reinterpret_cast<__nv_bfloat16*>(w2s + H);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  __nv_bfloat16* dp_hi = warp_base + warp * (2 * 32 * (TS + WS));  // [32][TS]
  __nv_bfloat16* dp_lo = dp_hi + 32 * TS;
  __nv_bfloat16* e_hi = dp_lo + 32 * TS;                            // [32][WS]
  __nv_bfloat16* e_lo = e_hi + 32 * WS;

  stage_planes<H, K0>(w1, w_hi, w_lo, false);
  stage_planes<H, K0>(w1, wt_hi, wt_lo, true);
  for (int e = threadIdx.x; e < H; e += DEC_THREADS) {
    b1s[e] = __ldg(b1 + e);
    w2s[e] = __ldg(w2 + e);
  }
  __syncthreads();

  float wacc[H / 16][K0 / 8][4];
#pragma unroll
  for (int a = 0; a < H / 16; ++a)
#pragma unroll
    for (int b = 0; b < K0 / 8; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) wacc[a][b][c] = 0.0f;
  float pb1[H / 8][2], pw2[H / 8][2];
#pragma unroll
  for (int nt = 0; nt < H / 8; ++nt) { pb1[nt][0] = pb1[nt][1] = 0.0f; pw2[nt][0] = pw2[nt][1] = 0.0f; }
  float pb2 = 0.0f;

  const int64_t chunks = (n + 31) / 32;
  for (int64_t chunk = static_cast<int64_t>(blockIdx.x) * NWARP + warp; chunk < chunks;
       chunk += static_cast<int64_t>(gridDim.x) * NWARP) {
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
      const int64_t row0 = chunk * 32 + 16 * mt;
      const int64_t r_lo = row0 + g, r_hi = row0 + g + 8;
      float acc[H / 8][4];
      {
        uint32_t a_hi[K0 / 16][4], a_lo[K0 / 16][4];
        load_a_frags<K0>(enc, row0, n, g, t, a_hi, a_lo);
#pragma unroll
        for (int kt = 0; kt < K0 / 16; ++kt)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int col = 16 * kt + 8 * half + 2 * t;
            *reinterpret_cast<uint32_t*>(e_hi + (16 * mt + g) * WS + col) = a_hi[kt][2 * half];
            *reinterpret_cast<uint32_t*>(e_lo + (16 * mt + g) * WS + col) = a_lo[kt][2 * half];
            *reinterpret_cast<uint32_t*>(e_hi + (16 * mt + g + 8) * WS + col) = a_hi[kt][2 * half + 1];
            *reinterpret_cast<uint32_t*>(e_lo + (16 * mt + g + 8) * WS + col) = a_lo[kt][2 * half + 1];
          }
        hidden_mma<K0, H>(a_hi, a_lo, w_hi, w_lo, b1s, g, t, acc);
      }
      float dp2_lo = 0.0f, dp2_hi = 0.0f;
      if (r_lo < n) dp2_lo = __ldg(gy + r_lo) * activate_grad_rt(act2, __ldg(pre2 + r_lo), 1.0f);
      if (r_hi < n) dp2_hi = __ldg(gy + r_hi) * activate_grad_rt(act2, __ldg(pre2 + r_hi), 1.0f);
      if (t == 0) pb2 += dp2_lo + dp2_hi;
      uint32_t da_hi[H / 16][4], da_lo[H / 16][4];
#pragma unroll
      for (int nt = 0; nt < H / 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float dp2 = e < 2 ? dp2_lo : dp2_hi;
          float a, gp;
          act_and_grad<ACT1>(acc[nt][e], a, gp);
          pw2[nt][e & 1] = fmaf(dp2, a, pw2[nt][e & 1]);
          const float dpre = dp2 * w2s[8 * nt + 2 * t + (e & 1)] * gp;
          pb1[nt][e & 1] += dpre;
          acc[nt][e] = dpre;
        }
        uint32_t h0, l0, h1, l1;
        split_pair(acc[nt][0], acc[nt][1], h0, l0);  // row g
        split_pair(acc[nt][2], acc[nt][3], h1, l1);  // row g + 8
        const int col = 8 * nt + 2 * t;
        *reinterpret_cast<uint32_t*>(dp_hi + (16 * mt + g) * TS + col) = h0;
        *reinterpret_cast<uint32_t*>(dp_lo + (16 * mt + g) * TS + col) = l0;
        *reinterpret_cast<uint32_t*>(dp_hi + (16 * mt + g + 8) * TS + col) = h1;
        *reinterpret_cast<uint32_t*>(dp_lo + (16 * mt + g + 8) * TS + col) = l1;
        // accumulator fragment -> A fragment of the dEnc product (k-tile nt/2, halves by nt parity)
        da_hi[nt / 2][2 * (nt & 1) + 0] = h0; da_hi[nt / 2][2 * (nt & 1) + 1] = h1;
        da_lo[nt / 2][2 * (nt & 1) + 0] = l0; da_lo[nt / 2][2 * (nt & 1) + 1] = l1;
      }
      // dEnc (16 x K0) = dPre1 (16 x H) . W1 (H x K0); B[k = j][n = kenc] = W1^T planes [kenc][j]
#pragma unroll
      for (int nt2 = 0; nt2 < K0 / 8; ++nt2) {
        float dacc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int kt2 = 0; kt2 < H / 16; ++kt2) {
          const int off = (8 * nt2 + g) * TS + 16 * kt2 + 2 * t;
          const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(wt_hi + off), bh1 = *reinterpret_cast<const uint32_t*>(wt_hi + off + 8);
          const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(wt_lo + off), bl1 = *reinterpret_cast<const uint32_t*>(wt_lo + off + 8);
          mma_bf16_16816(dacc, da_lo[kt2], bh0, bh1);
          mma_bf16_16816(dacc, da_hi[kt2], bl0, bl1);
          mma_bf16_16816(dacc, da_hi[kt2], bh0, bh1);
        }
        const int col = 8 * nt2 + 2 * t;
        if (r_lo < n) *reinterpret_cast<float2*>(denc + r_lo * K0 + col) = make_float2(dacc[0], dacc[1]);
        if (r_hi < n) *reinterpret_cast<float2*>(denc + r_hi * K0 + col) = make_float2(dacc[2], dacc[3]);
      }
    }
    __syncwarp();
    // dW1 (H x K0) += dPre1^T (H x 32) . enc (32 x K0), operands transposed on the way out of shared memory
    const int lm = lane >> 3, lr = lane & 7;
#pragma unroll
    for (int nt2 = 0; nt2 < K0 / 8; ++nt2) {
      uint32_t bh[4], bl[4];  // {b0,b1} of coordinate k-tile 0, {b0,b1} of k-tile 1
      ldmatrix_x4_trans(bh, e_hi + (8 * lm + lr) * WS + 8 * nt2);
      ldmatrix_x4_trans(bl, e_lo + (8 * lm + lr) * WS + 8 * nt2);
#pragma unroll
      for (int jt = 0; jt < H / 16; ++jt) {
#pragma unroll
        for (int ct = 0; ct < 2; ++ct) {
          uint32_t ah[4], al[4];
          const int off = (16 * ct + 8 * (lm >> 1) + lr) * TS + 16 * jt + 8 * (lm & 1);
          ldmatrix_x4_trans(ah, dp_hi + off);
          ldmatrix_x4_trans(al, dp_lo + off);
          mma_bf16_16816(wacc[jt][nt2], al, bh[2 * ct], bh[2 * ct + 1]);
          mma_bf16_16816(wacc[jt][nt2], ah, bl[2 * ct], bl[2 * ct + 1]);
          mma_bf16_16816(wacc[jt][nt2], ah, bh[2 * ct], bh[2 * ct + 1]);
        }
      }
    }
    __syncwarp();
  }

  // ---- flush ----
  __syncthreads();
  float* red = reinterpret_cast<float*>(warp_base);  // reuse the per-warp staging area: [NWARP][H*K0] floats
  static_assert(sizeof(float) * H * K0 <= 2 * 32 * ((H + MMA_PAD) + (K0 + MMA_PAD)) * sizeof(__nv_bfloat16), "staging too small");
  float* mine = reinterpret_cast<float*>(dp_hi);
#pragma unroll
  for (int jt = 0; jt < H / 16; ++jt)
#pragma unroll
    for (int nt2 = 0; nt2 < K0 / 8; ++nt2) {
      const int j = 16 * jt + g, k = 8 * nt2 + 2 * t;
      mine[j * K0 + k] = wacc[jt][nt2][0];
      mine[j * K0 + k + 1] = wacc[jt][nt2][1];
      mine[(j + 8) * K0 + k] = wacc[jt][nt2][2];
      mine[(j + 8) * K0 + k + 1] = wacc[jt][nt2][3];
    }
  __syncthreads();
  constexpr int WARP_STRIDE_F = 2 * 32 * (TS + WS) * static_cast<int>(sizeof(__nv_bfloat16)) / static_cast<int>(sizeof(float));
  for (int e = threadIdx.x; e < H * K0; e += DEC_THREADS) {
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) s += red[w * WARP_STRIDE_F + e];
    red_add_f32(gw1 + e, s);
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
}

template <int K0, int H, int ACT1>
int launch_mma_bwd(const float* enc, int64_t n, const float* w1, const float* b1, const float* w2, const float* pre2,
                   const float* gy, int act2, float* denc, float* gw1, float* gb1, float* gw2, float* gb2, cudaStream_t s) {
  constexpr size_t smem = mma_bwd_smem_bytes<K0, H>();
  MRI_CUDA_OK(cudaFuncSetAttribute(decoder2_mma_bwd_kernel<K0, H, ACT1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
  int64_t blocks = ((n + 31) / 32 + 3) / 4;
  const int64_t cap = 2LL * sm_count();
  if (blocks > cap) blocks = cap;
  decoder2_mma_bwd_kernel<K0, H, ACT1><<<static_cast<int>(blocks), DEC_THREADS, smem, s>>>(enc, n, w1, b1, w2, pre2, gy, act2,
                                                                                            denc, gw1, gb1, gw2, gb2);
  MRI_LAUNCH_OK("decoder2_mma_bwd_kernel");
  return MRI_OK;
}

template <int K0, int H>
size_t bwd_smem_bytes() {
  return sizeof(float) * (2 * K0 * H + 2 * H + 2 * DEC_THREADS * (H + DP_PAD) + K0 * (DEC_THREADS + 1));
}

template <int K0, int H, int ACT1>
int launch_fwd(const float* enc, int64_t n, const float* w1, const float* b1, const float* w2, const float* b2, int act2,
               float* y, float* pre2, cudaStream_t s) {
  int64_t blocks = (n + DEC_THREADS - 1) / DEC_THREADS;
  const int64_t cap = 8LL * sm_count();
  if (blocks > cap) blocks = cap;
  decoder2_fwd_kernel<K0, H, ACT1><<<static_cast<int>(blocks), DEC_THREADS, 0, s>>>(enc, n, w1, b1, w2, b2, act2, y, pre2);
  MRI_LAUNCH_OK("decoder2_fwd_kernel");
  return MRI_OK;
}

template <int K0, int H, int ACT1>
int launch_bwd(const float* enc, int64_t n, const float* w1, const float* b1, const float* w2, const float* pre2,
               const float* gy, int act2, float* denc, float* gw1, float* gb1, float* gw2, float* gb2, cudaStream_t s) {
  const size_t smem = bwd_smem_bytes<K0, H>();
  MRI_CUDA_OK(cudaFuncSetAttribute(decoder2_bwd_kernel<K0, H, ACT1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
  int64_t blocks = (n + DEC_THREADS - 1) / DEC_THREADS;
  const int64_t cap = 2LL * sm_count();
  if (blocks > cap) blocks = cap;
  decoder2_bwd_kernel<K0, H, ACT1><<<static_cast<int>(blocks), DEC_THREADS, smem, s>>>(enc, n, w1, b1, w2, pre2, gy, act2,
                                                                                        denc, gw1, gb1, gw2, gb2);
  MRI_LAUNCH_OK("decoder2_bwd_kernel");
  return MRI_OK;
}



// all (H, activation) variants of one input width; `cuda_cores` selects the thread-per-coordinate kernels (profiling switch)
template <int K0V>
int decoder2_forward_k(bool cuda_cores, int h, int act1, const float* enc, int64_t n, const float* w1, const float* b1,
                       const float* w2, const float* b2, int act2, float* y, float* pre2, cudaStream_t s) {
#define MRI_DEC_ONE(HV, ACTV)                                                                                          \
  if (h == HV && act1 == ACTV)                                                                                         \
    return cuda_cores ? launch_fwd<K0V, HV, ACTV>(enc, n, w1, b1, w2, b2, act2, y, pre2, s)                             \
                      : launch_mma_fwd<K0V, HV, ACTV>(enc, n, w1, b1, w2, b2, act2, y, pre2, s);
  MRI_DEC_ONE(32, MRI_ACT_GELU) MRI_DEC_ONE(64, MRI_ACT_GELU) MRI_DEC_ONE(32, MRI_ACT_RELU) MRI_DEC_ONE(64, MRI_ACT_RELU)
#undef MRI_DEC_ONE
  return fail(MRI_ERR_UNSUPPORTED, "decoder2_forward: dispatch miss");
}
template <int K0V>
int decoder2_backward_k(bool cuda_cores, int h, int act1, const float* enc, int64_t n, const float* w1, const float* b1,
                        const float* w2, const float* pre2, const float* grad_y, int act2, float* grad_enc, float* grad_w1,
                        float* grad_b1, float* grad_w2, float* grad_b2, cudaStream_t s) {
#define MRI_DEC_ONE(HV, ACTV)                                                                                                      \
  if (h == HV && act1 == ACTV)                                                                                                     \
    return cuda_cores                                                                                                              \
               ? launch_bwd<K0V, HV, ACTV>(enc, n, w1, b1, w2, pre2, grad_y, act2, grad_enc, grad_w1, grad_b1, grad_w2, grad_b2, s)  \
               : launch_mma_bwd<K0V, HV, ACTV>(enc, n, w1, b1, w2, pre2, grad_y, act2, grad_enc, grad_w1, grad_b1, grad_w2, grad_b2, s);
  MRI_DEC_ONE(32, MRI_ACT_GELU) MRI_DEC_ONE(64, MRI_ACT_GELU) MRI_DEC_ONE(32, MRI_ACT_RELU) MRI_DEC_ONE(64, MRI_ACT_RELU)
#undef MRI_DEC_ONE
  return fail(MRI_ERR_UNSUPPORTED, "decoder2_backward: dispatch miss");
}

}  // namespace

#define MRI_DECODER_K_DECL(K0V)                                                                                                  \
  int decoder2_forward_k##K0V(bool cuda_cores, int h, int act1, const float* enc, int64_t n, const float* w1, const float* b1,   \
                              const float* w2, const float* b2, int act2, float* y, float* pre2, cudaStream_t s);               \
  int decoder2_backward_k##K0V(bool cuda_cores, int h, int act1, const float* enc, int64_t n, const float* w1, const float* b1,  \
                               const float* w2, const float* pre2, const float* grad_y, int act2, float* grad_enc,             \
                               float* grad_w1, float* grad_b1, float* grad_w2, float* grad_b2, cudaStream_t s);
#define MRI_DECODER_K_DEFINE(K0V)                                                                                                \
  int decoder2_forward_k##K0V(bool cuda_cores, int h, int act1, const float* enc, int64_t n, const float* w1, const float* b1,   \
                              const float* w2, const float* b2, int act2, float* y, float* pre2, cudaStream_t s) {              \
    return decoder2_forward_k<K0V>(cuda_cores, h, act1, enc, n, w1, b1, w2, b2, act2, y, pre2, s);                              \
  }                                                                                                                              \
  int decoder2_backward_k##K0V(bool cuda_cores, int h, int act1, const float* enc, int64_t n, const float* w1, const float* b1,  \
                               const float* w2, const float* pre2, const float* grad_y, int act2, float* grad_enc,             \
                               float* grad_w1, float* grad_b1, float* grad_w2, float* grad_b2, cudaStream_t s) {               \
    return decoder2_backward_k<K0V>(cuda_cores, h, act1, enc, n, w1, b1, w2, pre2, grad_y, act2, grad_enc, grad_w1, grad_b1,     \
                                    grad_w2, grad_b2, s);                                                                       \
  }

}  // namespace mri
