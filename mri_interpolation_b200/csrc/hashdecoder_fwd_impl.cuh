// Kernel template of the encoder + decoder forward fusion (see hashdecoder_fwd.cu for the design notes); included by
// hashdecoder_fwd.cu (headline geometry F = 2, L = 16, H = 64 with compile-time activation, plus the dense-sweep
// coordinate sources) and hashdecoder_fwd_geo.cu (the other F = 2 geometries: L = 4 / 8 / 16, H = 64 / 128).
#pragma once
#include <stdlib.h>
#include "common.cuh"
#include "grid_device.cuh"
#include "hash_device.cuh"
#include "mma_device.cuh"

namespace mri {
namespace {

// coordinates of rows (row0 + g, row0 + g + 8) of a tile, from a (n, D) batch ...
template <int D>
struct BatchCoords {
  // tile walk: a warp's consecutive tiles are a grid stride apart.  Measured on the B200 (2^19 locality-ordered
  // coordinates): 0.230 ms against 0.258 ms with a contiguous range per warp and 0.239 ms with one per block - a
  // random-subset batch has little to reuse from one tile to the next
  static constexpr bool kContiguousTiles = false;
  const float* x;
  __device__ __forceinline__ void load_pair(int64_t row0, int64_t n, int lane, float (&lo)[D], float (&hi)[D]) const {
    const int64_t r_lo = row0 + (lane >> 2), r_hi = r_lo + 8;
    load_coord<D>(x, r_lo < n ? r_lo : 0, lo);
    load_coord<D>(x, r_hi < n ? r_hi : 0, hi);
  }
  __device__ __forceinline__ int64_t out_index(int64_t row) const { return row; }
};
// ... or synthesised from the flat voxel index of a dense grid: lanes 0-15 each decompose one index, the quads pick
// their two rows up with shuffles (no 4x redundant integer divisions)
template <int D>
struct SweepCoords {
  static constexpr bool kContiguousTiles = true;
  const float* axes;
  GridDesc gd;
  int64_t first;
  __device__ __forceinline__ void load_pair(int64_t row0, int64_t n, int lane, float (&lo)[D], float (&hi)[D]) const {
    const int64_t r = row0 + (lane & 15);
    float v[D];
    voxel_coord<D>(axes, gd, first + (r < n ? r : 0), v);
    const int g = lane >> 2;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      lo[d] = __shfl_sync(0xffffffffu, v[d], g);
      hi[d] = __shfl_sync(0xffffffffu, v[d], g + 8);
    }
  }
  __device__ __forceinline__ int64_t out_index(int64_t row) const { return row; }
};
// ... or from a whole-plane box [plane0, plane0 + planes) x (other axes) of the grid walked with the axis-0 index
// FASTEST (then axis 1, 2, 3): axis 0 is the one axis whose hash prime is 1, so the 16 voxels of an m-tile - neighbours
// along axis 0 - gather from neighbouring table rows (same 128-byte lines / 32-byte sectors on all but the finest
// levels), where a C-order walk (last axis fastest) lands every voxel on unrelated rows.  The result is stored at the
// voxel's C-order position, so the output volume is the same array.
template <int D>
struct SweepCoordsAxis0 {
  // dense walk: every warp takes its own contiguous range of tiles, so the coarse-level rows gathered for one tile are
  // still in the SM's L1 for the next (measured: 6.50 vs 7.07 ms for the 21.6 M-voxel sweep; see the kernel's tile loop)
  static constexpr bool kContiguousTiles = true;
  const float* axes;
  GridDesc gd;
  int64_t out_base;   // C-order index of the box's first voxel minus the C-order index out[0] stands for
  uint32_t plane0, planes;
  __device__ __forceinline__ void decompose(uint32_t r, uint32_t (&i)[D]) const {
    uint32_t q = r / planes;
    i[0] = plane0 + (r - q * planes);
#pragma unroll
    for (int d = 1; d < D - 1; ++d) {
      const uint32_t q2 = q / static_cast<uint32_t>(gd.shape[d]);
      i[d] = q - q2 * static_cast<uint32_t>(gd.shape[d]);
      q = q2;
    }
    i[D - 1] = q;
  }
  __device__ __forceinline__ void load_pair(int64_t row0, int64_t n, int lane, float (&lo)[D], float (&hi)[D]) const {
    const int64_t r = row0 + (lane & 15);
    uint32_t i[D];
    decompose(static_cast<uint32_t>(r < n ? r : 0), i);
    float v[D];
#pragma unroll
    for (int d = 0; d < D; ++d) v[d] = __ldg(axes + gd.axis_off[d] + i[d]);
    const int g = lane >> 2;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      lo[d] = __shfl_sync(0xffffffffu, v[d], g);
      hi[d] = __shfl_sync(0xffffffffu, v[d], g + 8);
    }
  }
  __device__ __forceinline__ int64_t out_index(int64_t row) const {
    uint32_t i[D];
    decompose(static_cast<uint32_t>(row), i);
    int64_t flat = i[0] - plane0;
#pragma unroll
    for (int d = 1; d < D; ++d) flat = flat * gd.shape[d] + i[d];
    return out_base + flat;
  }
};

// K0 = L * 2 in {8, 16, 32} (L = 4, 8, 16 levels of 2 features); the decoder product runs on KP = max(K0, 16) columns
// (one mma k-tile at least: for L = 4 the upper half of the tile is zero in both operands).  ACT1 >= 0: compile-time
// activation (headline geometry); ACT1 = ACT_RUNTIME: `act1` selects GELU / ReLU at run time (one uniform branch per value).
constexpr int ACT_RUNTIME = -1;
template <int ACT1>
__device__ __forceinline__ float hidden_act(float pre, int act1) {
  if constexpr (ACT1 >= 0) return activate<ACT1>(pre, 1.0f);
  else return act1 == MRI_ACT_GELU ? activate<MRI_ACT_GELU>(pre, 1.0f) : activate<MRI_ACT_RELU>(pre, 1.0f);
}
template <int D, int K0, int H, int ACT1, class Coords>
__global__ void __launch_bounds__(DEC_THREADS, H == 64 ? 5 : 3) hashdecoder_mma_fwd_kernel(const Coords src, int64_t n,
                                                                             const float* __restrict__ tables,
                                                                             const __grid_constant__ LevelTable T,
                                                                             const float* __restrict__ w1, const float* __restrict__ b1,
                                                                             const float* __restrict__ w2, const float* __restrict__ b2,
                                                                             int act1, int act2, float* __restrict__ enc_out,
                                                                             float* __restrict__ y, float* __restrict__ pre2_out) {
  static_assert(K0 == 8 || K0 == 16 || K0 == 32, "F = 2 and L = 4, 8 or 16");
  constexpr int KP = K0 < 16 ? 16 : K0;
  constexpr int WS = KP + MMA_PAD;
  __shared__ __align__(16) __nv_bfloat16 w_hi[H * WS];
  __shared__ __align__(16) __nv_bfloat16 w_lo[H * WS];
  __shared__ float b1s[H];
  __shared__ float w2s[H];
  __shared__ LevelDev lvs[K0 / 2];  // lanes of one instruction work on two different levels: shared memory, not c[] replays
  stage_planes<H, K0, KP>(w1, w_hi, w_lo, false);
  for (int e = threadIdx.x; e < H; e += DEC_THREADS) {
    b1s[e] = __ldg(b1 + e);
    w2s[e] = __ldg(w2 + e);
  }
  if (threadIdx.x < K0 / 2) lvs[threadIdx.x] = T.lv[threadIdx.x];
  const float b2v = __ldg(b2);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int b0 = t & 1;  // axis-0 half of the pair-lane mapping
  const int64_t tiles = (n + 15) / 16;
  constexpr bool contiguous_tiles = Coords::kContiguousTiles;  // see the coordinate sources above
  constexpr int WARPS = DEC_THREADS / 32;
  const int64_t n_warps = static_cast<int64_t>(gridDim.x) * WARPS;
  const int64_t wid = static_cast<int64_t>(blockIdx.x) * WARPS + warp;
  // contiguous: the BLOCK owns a contiguous range of tiles and its four warps interleave inside it - neighbouring tiles at
  // the same time (shared L1 lines now) and in sequence (still in L1 next).  Sweep of 21.6 M voxels: grid stride 7.07 ms,
  // one range per warp 6.62 ms, one range per block 6.50 ms.
  const int64_t blk_begin = (static_cast<int64_t>(blockIdx.x) * tiles) / gridDim.x;
  const int64_t blk_end = ((static_cast<int64_t>(blockIdx.x) + 1) * tiles) / gridDim.x;
  const int64_t tile_begin = contiguous_tiles ? blk_begin + warp : wid;
  const int64_t tile_end = contiguous_tiles ? blk_end : tiles;
  const int64_t tile_step = contiguous_tiles ? WARPS : n_warps;
  for (int64_t tile = tile_begin; tile < tile_end; tile += tile_step) {
    const int64_t row0 = tile * 16;
    const int64_t rows[2] = {row0 + g, row0 + g + 8};
    float xv[2][D];
    src.load_pair(row0, n, lane, xv[0], xv[1]);
    uint32_t a_hi[KP / 16][4], a_lo[KP / 16][4];
    if constexpr (K0 < KP) a_hi[0][2] = a_hi[0][3] = a_lo[0][2] = a_lo[0][3] = 0u;  // L = 4: columns 8-15 of the k-tile are padding
    // A REAL loop over the 8-column halves (levels 4q .. 4q+3): unrolled, the gather of the four halves was 80-86 KB of SASS
    // per kernel (33-38 KB now); the fragment registers are picked with a compile-time if-chain below.  Measured neutral
    // on the B200 (0.228 vs 0.230 ms, sweep 3.1 vs 3.2 G voxels/s): ncu's stall_no_instruction (0.45-1.4 per issue) was not
    // what held the kernel back - kept for the smaller binary.
#pragma unroll 1
    for (int q = 0; q < K0 / 8; ++q) {  // q-th 8-column half: levels 4q .. 4q+3, this lane ends up with level 4q + t
      Feat<2> part[2][2];               // [level of the pair: even / odd][row g / g+8], this lane's axis-0 half
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const LevelDev lv = lvs[4 * q + (t & 2) + which];
        const float* __restrict__ tbl = tables + lv.offset;
        // one branch per level (a few coarse levels have non-power-of-two row counts), both rows inside it: the 16
        // gathers of a level are straight-line code and go out back-to-back
        encode_half_level_rows<D>(make_cell<D>(xv[0], lv), make_cell<D>(xv[1], lv), b0, lv, tbl, part[which][0], part[which][1]);
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        float full[2];
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          const float mine = b0 ? part[1][rr].v[f] : part[0][rr].v[f];
          const float send = b0 ? part[0][rr].v[f] : part[1][rr].v[f];
          full[f] = mine + __shfl_xor_sync(0xffffffffu, send, 1);
        }
        uint32_t hi_bits, lo_bits;
        split_pair(full[0], full[1], hi_bits, lo_bits);
#pragma unroll
        for (int qq = 0; qq < K0 / 8; ++qq)  // static register indices: predicated moves, no local memory
          if (qq == q) { a_hi[qq >> 1][2 * (qq & 1) + rr] = hi_bits; a_lo[qq >> 1][2 * (qq & 1) + rr] = lo_bits; }
        if (enc_out != nullptr && rows[rr] < n)
          *reinterpret_cast<float2*>(enc_out + rows[rr] * K0 + 2 * (4 * q + t)) = make_float2(full[0], full[1]);
      }
    }
    float acc[H / 8][4];
    hidden_mma<KP, H>(a_hi, a_lo, w_hi, w_lo, b1s, g, t, acc);
    float s_lo = 0.0f, s_hi = 0.0f;
#pragma unroll
    for (int nt = 0; nt < H / 8; ++nt) {
      const float wl = w2s[8 * nt + 2 * t], wh = w2s[8 * nt + 2 * t + 1];
      s_lo = fmaf(hidden_act<ACT1>(acc[nt][0], act1), wl, s_lo);
      s_lo = fmaf(hidden_act<ACT1>(acc[nt][1], act1), wh, s_lo);
      s_hi = fmaf(hidden_act<ACT1>(acc[nt][2], act1), wl, s_hi);
      s_hi = fmaf(hidden_act<ACT1>(acc[nt][3], act1), wh, s_hi);
    }
    s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 1); s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 2);
    s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 1); s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 2);
    if (t == 0) {
      if (rows[0] < n) { const float p = s_lo + b2v; const int64_t o = src.out_index(rows[0]); y[o] = activate_rt(act2, p, 1.0f); if (pre2_out) pre2_out[o] = p; }
      if (rows[1] < n) { const float p = s_hi + b2v; const int64_t o = src.out_index(rows[1]); y[o] = activate_rt(act2, p, 1.0f); if (pre2_out) pre2_out[o] = p; }
    }
  }
}

template <int D, int K0, int H, int ACT1, class Coords>
int launch_fused_fwd(const Coords& src, int64_t n, const float* tables, const LevelTable& T, const float* w1, const float* b1,
                     const float* w2, const float* b2, int act1, int act2, float* enc, float* y, float* pre2, cudaStream_t s) {
  auto kernel = hashdecoder_mma_fwd_kernel<D, K0, H, ACT1, Coords>;
  // persistent grid = exactly one wave (blocks walk the tiles with a grid stride): a cap that is not a multiple of
  // the resident block count costs a whole extra pass of the tail blocks.  Cached per device.
  static DeviceCache resident_cache;
  const int dev = DeviceCache::device();
  int resident = resident_cache.slot[dev].load(std::memory_order_acquire);
  if (resident == 0) {
    int per_sm = 0;
    MRI_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, DEC_THREADS, 0));
    resident = (per_sm > 0 ? per_sm : 1) * sm_count();
    resident_cache.slot[dev].store(resident, std::memory_order_release);
  }
  int64_t blocks = ((n + 15) / 16 + 3) / 4;
  if (blocks > resident) blocks = resident;
  kernel<<<static_cast<int>(blocks), DEC_THREADS, 0, s>>>(src, n, tables, T, w1, b1, w2, b2, act1, act2, enc, y, pre2);
  MRI_LAUNCH_OK("hashdecoder_mma_fwd_kernel");
  return MRI_OK;
}


// Dense sweep of a slab [first, first + count) of the C-order voxel grid with the fused kernel: whole axis-0 planes inside
// the slab are walked axis-0-fastest (SweepCoordsAxis0); a ragged head / tail (slab boundaries that cut a plane) keeps the
// C-order walk.  MRI_SWEEP_ORDER=c restores the C-order walk everywhere (A/B runs).  Packed decoder: W1 (H x K0), b1 (H),
// w2 (H), b2 (1).
template <int D, int K0, int H, int ACT1>
int sweep_fused(const float* axes, const GridDesc& gd, int64_t first, int64_t count, const float* tables, const LevelTable& T,
                const float* decoder, int act1, int last_act, float* out, cudaStream_t s) {
  const float *w1 = decoder, *b1 = decoder + H * K0, *w2 = b1 + H, *b2 = w2 + H;
  static const bool axis0_walk = [] { const char* e = getenv("MRI_SWEEP_ORDER"); return !(e && e[0] == 'c'); }();
  int64_t plane = 1;
  for (int d = 1; d < D; ++d) plane *= gd.shape[d];
  const int64_t p_begin = (first + plane - 1) / plane, p_end = (first + count) / plane;
  const bool boxed = axis0_walk && p_end > p_begin && (p_end - p_begin) * plane < (int64_t{1} << 32);
  const int64_t head = boxed ? p_begin * plane - first : count;
  const int64_t box = boxed ? (p_end - p_begin) * plane : 0;
  const int64_t tail = count - head - box;
  int st = MRI_OK;
  if (head > 0)
    st = launch_fused_fwd<D, K0, H, ACT1>(SweepCoords<D>{axes, gd, first}, head, tables, T, w1, b1, w2, b2, act1, last_act, nullptr, out,
                                          nullptr, s);
  if (st == MRI_OK && box > 0)
    st = launch_fused_fwd<D, K0, H, ACT1>(
        SweepCoordsAxis0<D>{axes, gd, head, static_cast<uint32_t>(p_begin), static_cast<uint32_t>(p_end - p_begin)}, box, tables, T, w1,
        b1, w2, b2, act1, last_act, nullptr, out, nullptr, s);
  if (st == MRI_OK && tail > 0)
    st = launch_fused_fwd<D, K0, H, ACT1>(SweepCoords<D>{axes, gd, first + head + box}, tail, tables, T, w1, b1, w2, b2, act1, last_act,
                                          nullptr, out + head + box, nullptr, s);
  return st;
}

}  // namespace
}  // namespace mri
