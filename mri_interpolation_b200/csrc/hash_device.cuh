// Device-side building blocks of the hash-grid kernels, shared by hashgrid.cu (stand-alone gather / scatter) and
// decoder.cu (decoder backward fused with the scatter).
#pragma once
#include "common.cuh"

namespace mri {

template <int F>
struct Feat {
  float v[F];
};

template <int F>
__device__ __forceinline__ Feat<F> gather_row(const float* __restrict__ row) {
  Feat<F> r;
  if constexpr (F == 1) {
    r.v[0] = __ldg(row);
  } else if constexpr (F == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(row));
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
#pragma unroll
    for (int q = 0; q < F / 4; ++q) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(row) + q);
      r.v[4 * q + 0] = t.x; r.v[4 * q + 1] = t.y; r.v[4 * q + 2] = t.z; r.v[4 * q + 3] = t.w;
    }
  }
  return r;
}

template <int F>
__device__ __forceinline__ void scatter_row(float* row, const Feat<F>& g, float w) {
  if constexpr (F == 1) {
    red_add_f32(row, g.v[0] * w);
  } else if constexpr (F == 2) {
    red_add_v2(row, g.v[0] * w, g.v[1] * w);
  } else {
#pragma unroll
    for (int q = 0; q < F / 4; ++q)
      red_add_v4(row + 4 * q, g.v[4 * q] * w, g.v[4 * q + 1] * w, g.v[4 * q + 2] * w, g.v[4 * q + 3] * w);
  }
}

template <int F>
__device__ __forceinline__ void store_feat(float* dst, const Feat<F>& a) {
  if constexpr (F == 1) {
    dst[0] = a.v[0];
  } else if constexpr (F == 2) {
    *reinterpret_cast<float2*>(dst) = make_float2(a.v[0], a.v[1]);
  } else {
#pragma unroll
    for (int q = 0; q < F / 4; ++q)
      reinterpret_cast<float4*>(dst)[q] = make_float4(a.v[4 * q], a.v[4 * q + 1], a.v[4 * q + 2], a.v[4 * q + 3]);
  }
}

// ---- "pair-lane" mapping -----------------------------------------------------------------------
// Two adjacent lanes share one (coordinate, level): lane bit 0 selects the lower/upper cell index on axis 0
// (PRIME[0] = 1), each lane walks the 2^(D-1) corners of the remaining axes.  The two corners of an axis-0
// pair hash to rows h and h' = h ^ (x0 ^ (x0+1)); for x0 % 4 != 3 they fall into the same 32-byte sector
// (F = 2: 8-byte rows), and because they now sit in neighbouring lanes of the SAME load/red instruction the
// LSU merges them into one L1 wavefront / one L2 request: ~10 instead of 16 sectors per (coordinate, level).
template <int D, int F, bool POW2>
__device__ __forceinline__ Feat<F> encode_half_level(const Cell<D>& cell, int b0, const LevelDev& lv,
                                                     const float* __restrict__ tbl) {
  constexpr int CH = 1 << (D - 1);
  const uint32_t t0 = cell.lo[0] + static_cast<uint32_t>(b0);
  const float w0 = b0 ? cell.wu[0] : cell.wl[0];
  Feat<F> rows[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    uint32_t h = t0;
#pragma unroll
    for (int d = 1; d < D; ++d) h ^= ((c >> (d - 1)) & 1) ? (cell.lo[d] + prime(d)) : cell.lo[d];
    rows[c] = gather_row<F>(tbl + static_cast<size_t>(wrap_rows<POW2>(h, lv)) * F);
  }
  Feat<F> acc;
#pragma unroll
  for (int f = 0; f < F; ++f) acc.v[f] = 0.0f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    float w = w0;
#pragma unroll
    for (int d = 1; d < D; ++d) w = __fmul_rn(w, ((c >> (d - 1)) & 1) ? cell.wu[d] : cell.wl[d]);
#pragma unroll
    for (int f = 0; f < F; ++f) acc.v[f] = fmaf(rows[c].v[f], w, acc.v[f]);
  }
  return acc;
}

template <int D, int F, bool POW2>
__device__ __forceinline__ void scatter_half_level(const Cell<D>& cell, int b0, const LevelDev& lv, float* tbl, const Feat<F>& g) {
  constexpr int CH = 1 << (D - 1);
  const uint32_t t0 = cell.lo[0] + static_cast<uint32_t>(b0);
  const float w0 = b0 ? cell.wu[0] : cell.wl[0];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    uint32_t h = t0;
    float w = w0;
#pragma unroll
    for (int d = 1; d < D; ++d) {
      const bool up = (c >> (d - 1)) & 1;
      h ^= up ? (cell.lo[d] + prime(d)) : cell.lo[d];
      w = __fmul_rn(w, up ? cell.wu[d] : cell.wl[d]);
    }
    scatter_row<F>(tbl + static_cast<size_t>(wrap_rows<POW2>(h, lv)) * F, g, w);
  }
}


}  // namespace mri
