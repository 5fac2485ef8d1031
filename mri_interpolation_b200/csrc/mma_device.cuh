// Warp-level tensor-core building blocks (mma.sync.m16n8k16, bf16 hi/lo split planes, fp32 accumulate) shared by the
// decoder kernels (decoder.cu) and the encoder+decoder forward fusion (hashdecoder_fwd.cu).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace mri {

constexpr int DEC_THREADS = 128;

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (x, y) -> packed bf16 pairs hi = (bf16(x), bf16(y)) and lo = (bf16(x - hi.x), bf16(y - hi.y)), x in the low half.
// Two packed conversions (F2FP.BF16.PACK_AB) with the high parts read back by a shift / mask: 6 instructions, where
// two scalar conversions + a permute took 8 (the split runs on every A fragment of every product).
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(y), "f"(x));
  const float xh = __uint_as_float(hi << 16), yh = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(y - yh), "f"(x - xh));
}

constexpr int MMA_PAD = 8;  // bf16 elements of row padding: conflict-free 32-bit fragment loads

// W (rows x cols, fp32, row-major in global) -> two padded bf16 planes in shared memory.  KP >= COLS: column count the
// untransposed planes are laid out for (row stride KP + MMA_PAD, columns COLS .. KP-1 zero) when the product runs on
// more k columns than the matrix has (K0 = 8 inside one 16-wide mma k-tile).
template <int ROWS, int COLS, int KP = COLS>
__device__ __forceinline__ void stage_planes(const float* __restrict__ w, __nv_bfloat16* hi, __nv_bfloat16* lo, bool transpose) {
  for (int e = threadIdx.x; e < ROWS * COLS; e += blockDim.x) {
    const int r = e / COLS, c = e - r * COLS;
    const float v = __ldg(w + e);
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const int dst = transpose ? (c * (ROWS + MMA_PAD) + r) : (r * (KP + MMA_PAD) + c);
    hi[dst] = h;
    lo[dst] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
  if constexpr (KP > COLS) {
    if (!transpose)
      for (int e = threadIdx.x; e < ROWS * (KP - COLS); e += blockDim.x) {
        const int r = e / (KP - COLS), c = COLS + e - r * (KP - COLS);
        hi[r * (KP + MMA_PAD) + c] = __float2bfloat16_rn(0.0f);
        lo[r * (KP + MMA_PAD) + c] = __float2bfloat16_rn(0.0f);
      }
  }
}

// A fragments (hi/lo) of one 16-coordinate m-tile: rows (g, g+8) of `enc`, all K0 columns
template <int K0>
__device__ __forceinline__ void load_a_frags(const float* __restrict__ enc, int64_t row0, int64_t n, int g, int t,
                                             uint32_t (&a_hi)[K0 / 16][4], uint32_t (&a_lo)[K0 / 16][4]) {
  const int64_t r_lo = row0 + g, r_hi = row0 + g + 8;
#pragma unroll
  for (int kt = 0; kt < K0 / 16; ++kt) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int col = 16 * kt + 8 * half + 2 * t;
      float2 v0 = make_float2(0.f, 0.f), v1 = make_float2(0.f, 0.f);
      if (r_lo < n) v0 = __ldg(reinterpret_cast<const float2*>(enc + r_lo * K0 + col));
      if (r_hi < n) v1 = __ldg(reinterpret_cast<const float2*>(enc + r_hi * K0 + col));
      split_pair(v0.x, v0.y, a_hi[kt][2 * half + 0], a_lo[kt][2 * half + 0]);
      split_pair(v1.x, v1.y, a_hi[kt][2 * half + 1], a_lo[kt][2 * half + 1]);
    }
  }
}

// acc[nt] (16 x 8 tiles over the H hidden units) = bias + A . W1^T with the 3-pass split product
template <int K0, int H>
__device__ __forceinline__ void hidden_mma(const uint32_t (&a_hi)[K0 / 16][4], const uint32_t (&a_lo)[K0 / 16][4],
                                           const __nv_bfloat16* __restrict__ w_hi, const __nv_bfloat16* __restrict__ w_lo,
                                           const float* __restrict__ b1s, int g, int t, float (&acc)[H / 8][4]) {
  constexpr int WS = K0 + MMA_PAD;
#pragma unroll
  for (int nt = 0; nt < H / 8; ++nt) {
    const float bl = b1s[8 * nt + 2 * t], bh = b1s[8 * nt + 2 * t + 1];
    acc[nt][0] = bl; acc[nt][1] = bh; acc[nt][2] = bl; acc[nt][3] = bh;
#pragma unroll
    for (int kt = 0; kt < K0 / 16; ++kt) {
      const int off = (8 * nt + g) * WS + 16 * kt + 2 * t;
      const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(w_hi + off), bh1 = *reinterpret_cast<const uint32_t*>(w_hi + off + 8);
      const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(w_lo + off), bl1 = *reinterpret_cast<const uint32_t*>(w_lo + off + 8);
      mma_bf16_16816(acc[nt], a_lo[kt], bh0, bh1);
      mma_bf16_16816(acc[nt], a_hi[kt], bl0, bl1);
      mma_bf16_16816(acc[nt], a_hi[kt], bh0, bh1);
    }
  }
}

template <int ACT>
__device__ __forceinline__ void act_and_grad(float pre, float& a, float& g) {
  if constexpr (ACT == MRI_ACT_GELU) {
    float cdf, pdf;
    gelu_cdf_pdf(pre, cdf, pdf);
    a = pre * cdf;
    g = cdf + pre * pdf;
  } else if constexpr (ACT == MRI_ACT_RELU) {
    a = fmaxf(pre, 0.0f);
    g = pre > 0.0f ? 1.0f : 0.0f;
  } else {
    a = pre;
    g = 1.0f;
  }
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

template <int K0, int H>
constexpr size_t mma_bwd_smem_bytes() {
  return 2 * (H * (K0 + MMA_PAD) + K0 * (H + MMA_PAD)) * sizeof(__nv_bfloat16) + 2 * H * sizeof(float) +
         (DEC_THREADS / 32) * 2 * 32 * ((H + MMA_PAD) + (K0 + MMA_PAD)) * sizeof(__nv_bfloat16);
}

}  // namespace mri
