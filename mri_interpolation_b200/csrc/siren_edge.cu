// First / last layers of a wide SIREN around the tensor-core hidden layers (csrc/siren_tc.cu).
//
// The first layer has K = dim_in (2..4) and the last one dim_out (1..4) outputs: neither is a GEMM worth a
// tensor-core tile, both are pure HBM streams over the (n, H) activation matrix.  These kernels read/write the
// bf16 (hi, lo) planes the tensor-core layers consume directly, so no fp32 copy of the (n, H) activations and no
// separate split pass exist on the forward path, and the backward needs no generic SGEMM:
//   siren_first_fwd_kernel   planes(sin(w0 (x W0^T + b0)))  [+ w0 cos(.) for the backward]
//   rowdot_planes_kernel     y = (hi + lo) W_last^T + b_last
//   outer_mul_split_kernel   dPre = (gy W_last) * aux -> planes, + column sums (bias gradient of the last hidden layer)
//   wcolsum_kernel           out[q][j] += sum_r src[r][j] * wgt[r][q]   (dW_last from planes, dW0 from fp32 dPre0)
#include <cuda_bf16.h>

#include "common.cuh"

namespace mri {
namespace {

constexpr int EDGE_MAX_Q = 4;  // dim_in / dim_out handled by these kernels

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
  const __nv_bfloat162 h = __halves2bfloat162(ah, bh);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - __bfloat162float(ah), b - __bfloat162float(bh));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// A block owns a slab of rows; a thread owns 4 adjacent columns (its 4 x D weights and biases stay in registers) and
// every RL-th row of the slab (same thread layout as the backward stream kernels below).  The first version decomposed a
// flat element index with a 64-bit division per thread and re-read the weights per element: 533 us for 2^18 x 256
// outputs (1 TB/s); this one streams.
template <int D>
__global__ void __launch_bounds__(256) siren_first_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                              const float* __restrict__ b, int64_t n, int h, float w0,
                                                              int64_t rows_per_block, __nv_bfloat16* __restrict__ out_hi,
                                                              __nv_bfloat16* __restrict__ out_lo, float* __restrict__ aux) {
  const int cg = h / 4 < 256 ? h / 4 : 256, rl = 256 / cg;
  const int rlane = threadIdx.x / cg, clane = threadIdx.x - rlane * cg;
  if (rlane >= rl) return;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = (r0 + rows_per_block < n) ? r0 + rows_per_block : n;
  for (int c = 4 * clane; c < h; c += 4 * cg) {
    float wv[4][D], bv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      bv[u] = b ? __ldg(b + c + u) : 0.0f;
#pragma unroll
      for (int d = 0; d < D; ++d) wv[u][d] = __ldg(w + (c + u) * D + d);
    }
#pragma unroll 2
    for (int64_t r = r0 + rlane; r < r1; r += rl) {
      float xv[D];
#pragma unroll
      for (int d = 0; d < D; ++d) xv[d] = __ldg(x + r * ldx + d);
      float s[4], cs[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float pre = bv[u];
#pragma unroll
        for (int d = 0; d < D; ++d) pre = fmaf(xv[d], wv[u][d], pre);
        const float rr = reduce_2pi(w0 * pre);
        s[u] = __sinf(rr);
        cs[u] = w0 * __cosf(rr);
      }
      uint32_t h0, l0, h1, l1;
      split2(s[0], s[1], h0, l0);
      split2(s[2], s[3], h1, l1);
      const int64_t off = r * h + c;
      *reinterpret_cast<uint2*>(out_hi + off) = make_uint2(h0, h1);
      if (out_lo) *reinterpret_cast<uint2*>(out_lo + off) = make_uint2(l0, l1);
      if (aux) *reinterpret_cast<float4*>(aux + off) = make_float4(cs[0], cs[1], cs[2], cs[3]);
    }
  }
}

// y[r][o] = b[o] + sum_k (hi + lo)[r][k] w[o][k]; one warp per row, Q outputs
template <int Q>
__global__ void __launch_bounds__(256) rowdot_planes_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                                            const float* __restrict__ w, const float* __restrict__ b, int64_t n,
                                                            int h, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp; r < n; r += nwarps) {
    float acc[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[q] = 0.0f;
    for (int k = lane * 8; k < h; k += 256) {  // 8 bf16 = 16 bytes per lane per plane
      const uint4 vh = __ldg(reinterpret_cast<const uint4*>(hi + r * h + k));
      uint4 vl = make_uint4(0, 0, 0, 0);
      if (lo) vl = __ldg(reinterpret_cast<const uint4*>(lo + r * h + k));
      const uint32_t hw[4] = {vh.x, vh.y, vh.z, vh.w}, lw[4] = {vl.x, vl.y, vl.z, vl.w};
      float v[8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&hw[u]);
        const __nv_bfloat162 ll = *reinterpret_cast<const __nv_bfloat162*>(&lw[u]);
        v[2 * u] = __bfloat162float(hh.x) + __bfloat162float(ll.x);
        v[2 * u + 1] = __bfloat162float(hh.y) + __bfloat162float(ll.y);
      }
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const float4 w0v = __ldg(reinterpret_cast<const float4*>(w + q * h + k));
        const float4 w1v = __ldg(reinterpret_cast<const float4*>(w + q * h + k + 4));
        acc[q] = fmaf(v[0], w0v.x, acc[q]); acc[q] = fmaf(v[1], w0v.y, acc[q]);
        acc[q] = fmaf(v[2], w0v.z, acc[q]); acc[q] = fmaf(v[3], w0v.w, acc[q]);
        acc[q] = fmaf(v[4], w1v.x, acc[q]); acc[q] = fmaf(v[5], w1v.y, acc[q]);
        acc[q] = fmaf(v[6], w1v.z, acc[q]); acc[q] = fmaf(v[7], w1v.w, acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < Q; ++q) y[r * Q + q] = acc[q] + (b ? __ldg(b + q) : 0.0f);
    }
  }
}

// Thread layout of the two backward stream kernels below: a block owns a slab of rows; a thread owns 4 adjacent columns
// (16-byte fp32 / 8-byte plane accesses); CG = min(h / 4, 256) such threads sit side by side and RL = 256 / CG "row lanes"
// are stacked on top, each taking every RL-th row of the slab, so all 256 threads stream for any width (the first
// version gave a thread 2 columns and ALL rows: half the block idle at h = 256 and one 8-byte load in flight per thread -
// 1.4 TB/s; ncu launch list profiles/r02_launches_siren_ankle.csv).  Per-column partial sums meet in shared memory
// (RL-way red.shared), one global reduction per column, quantity and block at the end.
constexpr int EDGE_MAX_H = 2048;  // (Q + 1) * h floats of shared memory must stay below the 48 KB static limit

__device__ __forceinline__ void edge_layout(int h, int& cg, int& rl) {
  cg = h / 4 < 256 ? h / 4 : 256;
  rl = 256 / cg;
}

// dPre[r][j] = (sum_q gy[r][q] w[q][j]) * aux[r][j] -> planes; colsum[j] += sum_r dPre[r][j]
template <int Q>
__global__ void __launch_bounds__(256) outer_mul_split_kernel(const float* __restrict__ gy, const float* __restrict__ w,
                                                              const float* __restrict__ aux, int64_t n, int h, int64_t rows_per_block,
                                                              __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                                                              float* __restrict__ colsum) {
  extern __shared__ float edge_sm[];  // [h] column sums of this block
  int cg, rl;
  edge_layout(h, cg, rl);
  for (int j = threadIdx.x; j < h; j += 256) edge_sm[j] = 0.0f;
  __syncthreads();
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = (r0 + rows_per_block < n) ? r0 + rows_per_block : n;
  const int rlane = threadIdx.x / cg, clane = threadIdx.x - rlane * cg;
  if (rlane < rl) {
    for (int j = 4 * clane; j < h; j += 4 * cg) {
      float wq[Q][4];
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(w + q * h + j));
        wq[q][0] = t.x; wq[q][1] = t.y; wq[q][2] = t.z; wq[q][3] = t.w;
      }
      float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
      for (int64_t r = r0 + rlane; r < r1; r += rl) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(aux + r * h + j));
        float d[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const float g = __ldg(gy + r * Q + q);
#pragma unroll
          for (int u = 0; u < 4; ++u) d[u] = fmaf(g, wq[q][u], d[u]);
        }
        d[0] *= a.x; d[1] *= a.y; d[2] *= a.z; d[3] *= a.w;
        uint32_t h0, l0, h1, l1;
        split2(d[0], d[1], h0, l0);
        split2(d[2], d[3], h1, l1);
        *reinterpret_cast<uint2*>(out_hi + r * h + j) = make_uint2(h0, h1);
        if (out_lo) *reinterpret_cast<uint2*>(out_lo + r * h + j) = make_uint2(l0, l1);
#pragma unroll
        for (int u = 0; u < 4; ++u) s[u] += d[u];
      }
      if (colsum) {
#pragma unroll
        for (int u = 0; u < 4; ++u) atomicAdd(edge_sm + j + u, s[u]);
      }
    }
  }
  __syncthreads();
  if (colsum)
    for (int j = threadIdx.x; j < h; j += 256) red_add_f32(colsum + j, edge_sm[j]);
}

// out[q][j] (or out[j][q] when transposed_out) += sum_r src[r][j] * wgt[r][q];  colsum[j] += sum_r src[r][j]
// src is either an fp32 matrix or a (hi, lo) plane pair.
template <int Q, bool PLANES>
__global__ void __launch_bounds__(256) wcolsum_kernel(const float* __restrict__ src_f32, const __nv_bfloat16* __restrict__ src_hi,
                                                      const __nv_bfloat16* __restrict__ src_lo, const float* __restrict__ wgt,
                                                      int64_t ldw, int64_t n, int h, int64_t rows_per_block, int transposed_out,
                                                      float* __restrict__ out, float* __restrict__ colsum) {
  extern __shared__ float edge_sm[];  // [(Q + 1)][h]: Q weighted sums, then the plain column sums
  int cg, rl;
  edge_layout(h, cg, rl);
  for (int j = threadIdx.x; j < (Q + 1) * h; j += 256) edge_sm[j] = 0.0f;
  __syncthreads();
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = (r0 + rows_per_block < n) ? r0 + rows_per_block : n;
  const int rlane = threadIdx.x / cg, clane = threadIdx.x - rlane * cg;
  if (rlane < rl) {
    for (int j = 4 * clane; j < h; j += 4 * cg) {
      float acc[Q][4], s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int q = 0; q < Q; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0f;
#pragma unroll 4
      for (int64_t r = r0 + rlane; r < r1; r += rl) {
        float v[4];
        if constexpr (PLANES) {
          const uint2 hw = __ldg(reinterpret_cast<const uint2*>(src_hi + r * h + j));
          v[0] = __uint_as_float(hw.x << 16); v[1] = __uint_as_float(hw.x & 0xffff0000u);
          v[2] = __uint_as_float(hw.y << 16); v[3] = __uint_as_float(hw.y & 0xffff0000u);
          if (src_lo) {
            const uint2 lw = __ldg(reinterpret_cast<const uint2*>(src_lo + r * h + j));
            v[0] += __uint_as_float(lw.x << 16); v[1] += __uint_as_float(lw.x & 0xffff0000u);
            v[2] += __uint_as_float(lw.y << 16); v[3] += __uint_as_float(lw.y & 0xffff0000u);
          }
        } else {
          const float4 t = __ldg(reinterpret_cast<const float4*>(src_f32 + r * h + j));
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) s[u] += v[u];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const float g = __ldg(wgt + r * ldw + q);
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[q][u] = fmaf(v[u], g, acc[q][u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int q = 0; q < Q; ++q) atomicAdd(edge_sm + q * h + j + u, acc[q][u]);
        if (colsum) atomicAdd(edge_sm + Q * h + j + u, s[u]);
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < Q * h; e += 256) {
    const int q = e / h, j = e - q * h;
    red_add_f32(transposed_out ? out + static_cast<int64_t>(j) * Q + q : out + static_cast<int64_t>(q) * h + j, edge_sm[e]);
  }
  if (colsum)
    for (int j = threadIdx.x; j < h; j += 256) red_add_f32(colsum + j, edge_sm[Q * h + j]);
}

// column sums of a narrow (n, q <= 4) fp32 matrix: bias gradient of the output layer
__global__ void __launch_bounds__(256) small_colsum_kernel(const float* __restrict__ g, int64_t n, int q, float* __restrict__ out) {
  float acc[EDGE_MAX_Q] = {0.0f, 0.0f, 0.0f, 0.0f};
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < n; r += stride)
    for (int c = 0; c < q; ++c) acc[c] += __ldg(g + r * q + c);
  for (int c = 0; c < q; ++c) {
    float v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red_add_f32(out + c, v);
  }
}

int slab_rows(int64_t n, int64_t* blocks) {
  int64_t b = 8LL * sm_count();
  int64_t rpb = (n + b - 1) / b;
  if (rpb < 16) rpb = 16;
  *blocks = (n + rpb - 1) / rpb;
  return static_cast<int>(rpb);
}

}  // namespace
}  // namespace mri

using namespace mri;

extern "C" int mri_siren_first_forward(const float* x, int64_t ldx, const float* w, const float* b, int64_t n, int dim_in, int h,
                                       float w0, void* out_hi, void* out_lo, float* aux, void* stream) {
  if (n < 0) return fail(MRI_ERR_INVALID, "siren_first_forward: negative n");
  if (n == 0) return MRI_OK;
  if (!x || !w || !out_hi) return fail(MRI_ERR_INVALID, "siren_first_forward: null pointer");
  if (dim_in < 1 || dim_in > EDGE_MAX_Q || h % 8 != 0) return fail(MRI_ERR_UNSUPPORTED, "siren_first_forward: dim_in=%d h=%d", dim_in, h);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int64_t blocks;
  const int64_t rpb = slab_rows(n, &blocks);
  __nv_bfloat16* oh = static_cast<__nv_bfloat16*>(out_hi);
  __nv_bfloat16* ol = static_cast<__nv_bfloat16*>(out_lo);
  switch (dim_in) {
    case 1: siren_first_fwd_kernel<1><<<static_cast<int>(blocks), 256, 0, s>>>(x, ldx, w, b, n, h, w0, rpb, oh, ol, aux); break;
    case 2: siren_first_fwd_kernel<2><<<static_cast<int>(blocks), 256, 0, s>>>(x, ldx, w, b, n, h, w0, rpb, oh, ol, aux); break;
    case 3: siren_first_fwd_kernel<3><<<static_cast<int>(blocks), 256, 0, s>>>(x, ldx, w, b, n, h, w0, rpb, oh, ol, aux); break;
    default: siren_first_fwd_kernel<4><<<static_cast<int>(blocks), 256, 0, s>>>(x, ldx, w, b, n, h, w0, rpb, oh, ol, aux); break;
  }
  MRI_LAUNCH_OK("siren_first_fwd_kernel");
  return MRI_OK;
}

extern "C" int mri_siren_last_forward(const void* hi, const void* lo, const float* w, const float* b, int64_t n, int h, int m_out,
                                      float* y, void* stream) {
  if (n < 0) return fail(MRI_ERR_INVALID, "siren_last_forward: negative n");
  if (n == 0) return MRI_OK;
  if (!hi || !w || !y) return fail(MRI_ERR_INVALID, "siren_last_forward: null pointer");
  if (m_out < 1 || m_out > EDGE_MAX_Q || h % 8 != 0) return fail(MRI_ERR_UNSUPPORTED, "siren_last_forward: m_out=%d h=%d", m_out, h);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int64_t want = (n + 7) / 8;  // 8 warps per block, one row per warp iteration
  const int64_t cap = 16LL * sm_count();
  if (want > cap) want = cap;
  const __nv_bfloat16* ph = static_cast<const __nv_bfloat16*>(hi);
  const __nv_bfloat16* pl = static_cast<const __nv_bfloat16*>(lo);
  switch (m_out) {
    case 1: rowdot_planes_kernel<1><<<static_cast<int>(want), 256, 0, s>>>(ph, pl, w, b, n, h, y); break;
    case 2: rowdot_planes_kernel<2><<<static_cast<int>(want), 256, 0, s>>>(ph, pl, w, b, n, h, y); break;
    case 3: rowdot_planes_kernel<3><<<static_cast<int>(want), 256, 0, s>>>(ph, pl, w, b, n, h, y); break;
    default: rowdot_planes_kernel<4><<<static_cast<int>(want), 256, 0, s>>>(ph, pl, w, b, n, h, y); break;
  }
  MRI_LAUNCH_OK("rowdot_planes_kernel");
  return MRI_OK;
}

extern "C" int mri_siren_last_backward(const float* grad_y, const float* w, const float* aux, const void* act_hi, const void* act_lo,
                                       int64_t n, int h, int m_out, void* dpre_hi, void* dpre_lo, float* grad_b_hidden,
                                       float* grad_w_last, float* grad_b_last, void* stream) {
  if (n < 0) return fail(MRI_ERR_INVALID, "siren_last_backward: negative n");
  if (n == 0) return MRI_OK;
  if (!grad_y || !w || !aux || !act_hi || !dpre_hi || !grad_w_last) return fail(MRI_ERR_INVALID, "siren_last_backward: null pointer");
  if (m_out < 1 || m_out > EDGE_MAX_Q || h % 8 != 0 || h > EDGE_MAX_H) return fail(MRI_ERR_UNSUPPORTED, "siren_last_backward: m_out=%d h=%d", m_out, h);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int64_t blocks;
  const int rpb = slab_rows(n, &blocks);
  __nv_bfloat16* dh = static_cast<__nv_bfloat16*>(dpre_hi);
  __nv_bfloat16* dl = static_cast<__nv_bfloat16*>(dpre_lo);
  const __nv_bfloat16* ah = static_cast<const __nv_bfloat16*>(act_hi);
  const __nv_bfloat16* al = static_cast<const __nv_bfloat16*>(act_lo);
#define LAUNCH(Q)                                                                                                         \
  outer_mul_split_kernel<Q><<<static_cast<int>(blocks), 256, h * sizeof(float), s>>>(grad_y, w, aux, n, h, rpb, dh, dl, grad_b_hidden);   \
  wcolsum_kernel<Q, true><<<static_cast<int>(blocks), 256, (Q + 1) * h * sizeof(float), s>>>(nullptr, ah, al, grad_y, Q, n, h, rpb, 0, grad_w_last, nullptr);
  switch (m_out) {
    case 1: LAUNCH(1) break;
    case 2: LAUNCH(2) break;
    case 3: LAUNCH(3) break;
    default: LAUNCH(4) break;
  }
#undef LAUNCH
  MRI_LAUNCH_OK("siren_last_backward kernels");
  if (grad_b_last) {
    int64_t want = (n + 255) / 256;
    const int64_t cap = 2LL * sm_count();
    if (want > cap) want = cap;
    small_colsum_kernel<<<static_cast<int>(want), 256, 0, s>>>(grad_y, n, m_out, grad_b_last);
    MRI_LAUNCH_OK("small_colsum_kernel");
  }
  return MRI_OK;
}

extern "C" int mri_siren_first_backward(const float* dpre0, const float* x, int64_t ldx, int64_t n, int dim_in, int h,
                                        float* grad_w0, float* grad_b0, void* stream) {
  if (n < 0) return fail(MRI_ERR_INVALID, "siren_first_backward: negative n");
  if (n == 0) return MRI_OK;
  if (!dpre0 || !x || !grad_w0) return fail(MRI_ERR_INVALID, "siren_first_backward: null pointer");
  if (dim_in < 1 || dim_in > EDGE_MAX_Q || h % 8 != 0 || h > EDGE_MAX_H) return fail(MRI_ERR_UNSUPPORTED, "siren_first_backward: dim_in=%d h=%d", dim_in, h);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int64_t blocks;
  const int rpb = slab_rows(n, &blocks);
  // grad_w0 is (h, dim_in) row-major -> transposed_out
  switch (dim_in) {
    case 1: wcolsum_kernel<1, false><<<static_cast<int>(blocks), 256, (1 + 1) * h * sizeof(float), s>>>(dpre0, nullptr, nullptr, x, ldx, n, h, rpb, 1, grad_w0, grad_b0); break;
    case 2: wcolsum_kernel<2, false><<<static_cast<int>(blocks), 256, (2 + 1) * h * sizeof(float), s>>>(dpre0, nullptr, nullptr, x, ldx, n, h, rpb, 1, grad_w0, grad_b0); break;
    case 3: wcolsum_kernel<3, false><<<static_cast<int>(blocks), 256, (3 + 1) * h * sizeof(float), s>>>(dpre0, nullptr, nullptr, x, ldx, n, h, rpb, 1, grad_w0, grad_b0); break;
    default: wcolsum_kernel<4, false><<<static_cast<int>(blocks), 256, (4 + 1) * h * sizeof(float), s>>>(dpre0, nullptr, nullptr, x, ldx, n, h, rpb, 1, grad_w0, grad_b0); break;
  }
  MRI_LAUNCH_OK("wcolsum_kernel");
  return MRI_OK;
}
