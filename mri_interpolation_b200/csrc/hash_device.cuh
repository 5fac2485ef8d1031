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
// `row_sink` (parity instrumentation, nullptr in every production launch and then compiled away): the table row each
// gather of THIS code path addresses, stored at the reference's corner number (bit d = upper cell on axis d).
template <int D, int F, bool POW2>
__device__ __forceinline__ Feat<F> encode_half_level(const Cell<D>& cell, int b0, const LevelDev& lv,
                                                     const float* __restrict__ tbl, uint32_t* row_sink = nullptr) {
  constexpr int CH = 1 << (D - 1);
  const uint32_t t0 = cell.lo[0] + static_cast<uint32_t>(b0);
  const float w0 = b0 ? cell.wu[0] : cell.wl[0];
  Feat<F> rows[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    uint32_t h = t0;
#pragma unroll
    for (int d = 1; d < D; ++d) h ^= ((c >> (d - 1)) & 1) ? (cell.lo[d] + prime(d)) : cell.lo[d];
    const uint32_t row = wrap_rows<POW2>(h, lv);
    if (row_sink != nullptr) row_sink[b0 | (c << 1)] = row;
    rows[c] = gather_row<F>(tbl + static_cast<size_t>(row) * F);
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

// Both rows of an m-tile at one level (F = 2), for the fused forward: one branch per level, both rows inside it, so all
// 2 * 2^(D-1) gathers of the two rows are in flight together.
template <int D>
__device__ __forceinline__ void encode_half_level_rows(const Cell<D>& ca, const Cell<D>& cb, int b0, const LevelDev& lv,
                                                       const float* __restrict__ tbl, Feat<2>& oa, Feat<2>& ob) {
  if (lv.is_pow2) {
    oa = encode_half_level<D, 2, true>(ca, b0, lv, tbl);
    ob = encode_half_level<D, 2, true>(cb, b0, lv, tbl);
  } else {  // a few coarse levels (res^D < T): exact modulo by multiplication (wrap_rows<false>), same straight-line shape
    oa = encode_half_level<D, 2, false>(ca, b0, lv, tbl);
    ob = encode_half_level<D, 2, false>(cb, b0, lv, tbl);
  }
}

// ---- run merging for locality-ordered batches -------------------------------------------------------
// A batch ordered along axis 0 (functional.locality_sort: voxels of one axis-0 line are neighbours, axis-0 index
// ascending) puts samples that differ ONLY in their axis-0 coordinate into consecutive rows of an m-tile.  At a coarse
// level several of them fall into the same or into adjacent axis-0 cells, so their lower/upper corners are the SAME
// table rows for every one of the 2^(D-1) corner combinations of the other axes - and those combinations' weights are
// bit-identical too (same coordinates).  red.global.add does not merge identical addresses inside an instruction (ncu:
// every duplicate is its own L2 sector operation), so the duplicates are summed in registers first:
//   * lanes are laid out as in the fused kernels: lane = 4 g + t, row g of an 8-row group, t & 1 = axis-0 half
//   * `kk` = axis-0 cell index this lane updates (lower cell + half), `line_mask` bit (4g + t) = rows g and g+1 are live
//     and agree bit-for-bit on every coordinate but axis 0
//   * contiguous rows with equal kk form a run whose first lane (the head) ends up with the run's sum (3 doubling
//     steps); then the head of a lower-half run absorbs the upper-half run that ends on the previous row when that run
//     updates the same cell (cell(g-1) + 1 == cell(g)), and the absorbed run issues nothing.
// Works for ANY input order: rows that do not form contiguous runs simply stay separate (correct, just not merged).
struct MergedHalf {
  float v0, v1;
  bool active;
};
__device__ __forceinline__ MergedHalf merge_line_runs(uint32_t kk, float v0, float v1, bool live, uint32_t line_mask, int lane) {
  constexpr uint32_t full = 0xffffffffu;
  const int g = lane >> 2, t = lane & 3, b0 = t & 1;
  const uint32_t kk_next = __shfl_down_sync(full, kk, 4);
  const bool eq = ((line_mask >> lane) & 1u) && kk_next == kk;  // line_mask has no bits for g == 7
  const uint32_t m = __ballot_sync(full, eq);
  const uint32_t col = (m >> t) & 0x11111111u;                  // bit 4g': row g' continues into row g'+1 (my level, my half)
  const int cnt = (__ffs(~((col >> (4 * g)) | 0xEEEEEEEEu)) - 1) >> 2;  // rows after mine in my run
  const bool head = g == 0 || !((col >> (4 * (g - 1))) & 1u);
#pragma unroll
  for (int d = 1; d <= 4; d <<= 1) {
    const float a0 = __shfl_down_sync(full, v0, 4 * d), a1 = __shfl_down_sync(full, v1, 4 * d);
    if (d <= cnt) { v0 += a0; v1 += a1; }
  }
  // link between an upper-half run ending on row e and the lower-half run starting on row e+1
  const int e = g + cnt;
  const bool can = b0 ? (e < 7) : (g > 0);
  const int key_src = can ? (b0 ? 4 * (e + 1) + (t - 1) : lane - 3) : lane;
  const uint32_t kk_x = __shfl_sync(full, kk, key_src);
  const int line_bit = b0 ? 4 * e + t : (g > 0 ? lane - 4 : 0);
  const bool link = can && ((line_mask >> line_bit) & 1u) && kk_x == kk;
  // head lane of the upper-half run that contains row g-1 (column t | 1)
  const int gm1 = g > 0 ? g - 1 : 0;
  const uint32_t colp = (m >> (t | 1)) & 0x11111111u;
  const uint32_t below = ~colp & 0x11111111u & ((1u << (4 * gm1)) - 1u);
  const int s = below ? ((31 - __clz(below)) >> 2) + 1 : 0;
  const int v_src = (link && !b0) ? 4 * s + (t | 1) : lane;
  const float p0 = __shfl_sync(full, v0, v_src), p1 = __shfl_sync(full, v1, v_src);
  if (link && !b0) { v0 += p0; v1 += p1; }
  MergedHalf r;
  r.v0 = v0; r.v1 = v1;
  r.active = live && head && !(b0 && link);
  return r;
}

// scatter of one half level whose value (v0, v1) already carries the axis-0 weight (F = 2).  The hashes of all corners
// are formed first; power-of-two tables mask them (fused into the XOR by the compiler), the few non-power-of-two coarse
// levels reduce them with the multiply-based exact modulo, and ONE copy of the weight / reduction code serves both.
template <int D>
__device__ __forceinline__ void scatter_half_level_weighted(const Cell<D>& cell, int b0, const LevelDev& lv, float* tbl, float v0, float v1) {
  constexpr int CH = 1 << (D - 1);
  const uint32_t t0 = cell.lo[0] + static_cast<uint32_t>(b0);
  uint32_t row[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    uint32_t h = t0;
#pragma unroll
    for (int d = 1; d < D; ++d) h ^= ((c >> (d - 1)) & 1) ? (cell.lo[d] + prime(d)) : cell.lo[d];
    row[c] = h;
  }
  if (lv.is_pow2) {
#pragma unroll
    for (int c = 0; c < CH; ++c) row[c] &= lv.pow2_mask;
  } else {
#pragma unroll
    for (int c = 0; c < CH; ++c) row[c] = exact_mod(row[c], lv.rows, lv.magic);
  }
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    float w = 1.0f;
#pragma unroll
    for (int d = 1; d < D; ++d) {
      const float wd = ((c >> (d - 1)) & 1) ? cell.wu[d] : cell.wl[d];
      w = d == 1 ? wd : __fmul_rn(w, wd);
    }
    red_add_v2(tbl + static_cast<size_t>(row[c]) * 2, v0 * w, v1 * w);
  }
}

}  // namespace mri
