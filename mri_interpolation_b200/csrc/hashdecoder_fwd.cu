// Hash-grid encoder fused with the 2-layer decoder, forward direction (HashMLP.forward: models.py:712-739 on top of
// encoding.py:190-191), for the headline geometry F = 2, L = 16 (K0 = 32), H = 64, D = 3 / 4.
//
// One warp owns a 16-coordinate m-tile of the mma.sync.m16n8k16 product enc(16 x 32) . W1^T(32 x 64).  The A fragment
// of lane (g, t) holds rows g / g+8 and, per 8-column half of a k-tile, columns 2t, 2t+1 - i.e. exactly the two
// features of ONE level (4q + t for the q-th half).  So the gather is laid out to produce fragments directly:
// lanes t and t^1 form the pair of the pair-lane mapping (hash_device.cuh), each walks its axis-0 half of the corners
// of the pair's two levels for both rows, one shuffle completes the level sums, and the interpolated features are
// split into bf16 hi/lo planes in registers.  The (n, 32) encoding is written only when the backward needs it
// (training); inference (the dense sweep) writes 4 bytes per voxel.  The gather kernel alone is bound by the L1/L2
// sector rate with the issue slots 80 % idle - the tensor-core and GELU work of the decoder hides in those slots.
//
// Summation order of a level (lower axis-0 half + upper half) is the stand-alone kernel's, so `enc` is bit-identical
// to mri_hashgrid_forward; the decoder arithmetic is decoder2_mma_fwd_kernel's (3-pass split product, fp32 parity).
#include "common.cuh"
#include "grid_device.cuh"
#include "hash_device.cuh"
#include "mma_device.cuh"

namespace mri {
namespace {

// coordinates of rows (row0 + g, row0 + g + 8) of a tile, from a (n, D) batch ...
template <int D>
struct BatchCoords {
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

template <int D, int K0, int H, int ACT1, class Coords>
__global__ void __launch_bounds__(DEC_THREADS, 5) hashdecoder_mma_fwd_kernel(const Coords src, int64_t n,
                                                                             const float* __restrict__ tables,
                                                                             const __grid_constant__ LevelTable T,
                                                                             const float* __restrict__ w1, const float* __restrict__ b1,
                                                                             const float* __restrict__ w2, const float* __restrict__ b2,
                                                                             int act2, float* __restrict__ enc_out,
                                                                             float* __restrict__ y, float* __restrict__ pre2_out) {
  static_assert(K0 == 32, "two k-tiles: 16 levels of 2 features");
  constexpr int WS = K0 + MMA_PAD;
  __shared__ __align__(16) __nv_bfloat16 w_hi[H * WS];
  __shared__ __align__(16) __nv_bfloat16 w_lo[H * WS];
  __shared__ float b1s[H];
  __shared__ float w2s[H];
  __shared__ LevelDev lvs[K0 / 2];  // lanes of one instruction work on two different levels: shared memory, not c[] replays
  stage_planes<H, K0>(w1, w_hi, w_lo, false);
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
  for (int64_t tile = static_cast<int64_t>(blockIdx.x) * (DEC_THREADS / 32) + warp; tile < tiles;
       tile += static_cast<int64_t>(gridDim.x) * (DEC_THREADS / 32)) {
    const int64_t row0 = tile * 16;
    const int64_t rows[2] = {row0 + g, row0 + g + 8};
    float xv[2][D];
    src.load_pair(row0, n, lane, xv[0], xv[1]);
    uint32_t a_hi[K0 / 16][4], a_lo[K0 / 16][4];
#pragma unroll
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
        split_pair(full[0], full[1], a_hi[q >> 1][2 * (q & 1) + rr], a_lo[q >> 1][2 * (q & 1) + rr]);
        if (enc_out != nullptr && rows[rr] < n)
          *reinterpret_cast<float2*>(enc_out + rows[rr] * K0 + 2 * (4 * q + t)) = make_float2(full[0], full[1]);
      }
    }
    float acc[H / 8][4];
    hidden_mma<K0, H>(a_hi, a_lo, w_hi, w_lo, b1s, g, t, acc);
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
      if (rows[0] < n) { const float p = s_lo + b2v; const int64_t o = src.out_index(rows[0]); y[o] = activate_rt(act2, p, 1.0f); if (pre2_out) pre2_out[o] = p; }
      if (rows[1] < n) { const float p = s_hi + b2v; const int64_t o = src.out_index(rows[1]); y[o] = activate_rt(act2, p, 1.0f); if (pre2_out) pre2_out[o] = p; }
    }
  }
}

template <int D, int ACT1, class Coords>
int launch_fused_fwd(const Coords& src, int64_t n, const float* tables, const LevelTable& T, const float* w1, const float* b1,
                     const float* w2, const float* b2, int act2, float* enc, float* y, float* pre2, cudaStream_t s) {
  auto kernel = hashdecoder_mma_fwd_kernel<D, 32, 64, ACT1, Coords>;
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
  kernel<<<static_cast<int>(blocks), DEC_THREADS, 0, s>>>(src, n, tables, T, w1, b1, w2, b2, act2, enc, y, pre2);
  MRI_LAUNCH_OK("hashdecoder_mma_fwd_kernel");
  return MRI_OK;
}

bool geometry_ok(int dim, int n_levels, int n_features, int h, int act) {
  return (dim == 3 || dim == 4) && n_levels == 16 && n_features == 2 && h == 64 && (act == MRI_ACT_GELU || act == MRI_ACT_RELU);
}

}  // namespace

bool sweep_mma_supported(int dim, int n_levels, int n_features, int h, int act) {
  return geometry_ok(dim, n_levels, n_features, h, act);
}

int launch_sweep_mma(const float* axes, const GridDesc& gd, int dim, int64_t first, int64_t count, const float* tables,
                     const LevelTable& T, const float* decoder, int act, int last_act, float* out, cudaStream_t s) {
  constexpr int K0 = 32, H = 64;  // packed decoder: W1 (H x K0), b1 (H), w2 (H), b2 (1)
  const float *w1 = decoder, *b1 = decoder + H * K0, *w2 = b1 + H, *b2 = w2 + H;
  // whole axis-0 planes inside [first, first + count) are walked axis-0-fastest; a ragged head / tail (slab boundaries
  // that cut a plane) keeps the C-order walk.  MRI_SWEEP_ORDER=c restores the C-order walk everywhere (A/B runs).
  static const bool axis0_walk = [] { const char* e = getenv("MRI_SWEEP_ORDER"); return !(e && e[0] == 'c'); }();
  int64_t plane = 1;
  for (int d = 1; d < dim; ++d) plane *= gd.shape[d];
  const int64_t p_begin = (first + plane - 1) / plane, p_end = (first + count) / plane;
  const bool boxed = axis0_walk && p_end > p_begin && (p_end - p_begin) * plane < (int64_t{1} << 32);
  const int64_t head = boxed ? p_begin * plane - first : count;
  const int64_t box = boxed ? (p_end - p_begin) * plane : 0;
  const int64_t tail = count - head - box;
#define CALL_C(DV, ACTV, FIRST, COUNT, OUT)                                                                              \
  launch_fused_fwd<DV, ACTV>(SweepCoords<DV>{axes, gd, FIRST}, COUNT, tables, T, w1, b1, w2, b2, last_act, nullptr, OUT, \
                             nullptr, s)
#define CALL_B(DV, ACTV)                                                                                                          \
  launch_fused_fwd<DV, ACTV>(SweepCoordsAxis0<DV>{axes, gd, head, static_cast<uint32_t>(p_begin), static_cast<uint32_t>(p_end - p_begin)}, \
                             box, tables, T, w1, b1, w2, b2, last_act, nullptr, out, nullptr, s)
#define RUN(CALLEXPR)                              \
  do {                                             \
    const int st_ = (CALLEXPR);                    \
    if (st_ != MRI_OK) return st_;                 \
  } while (0)
#define DISPATCH(DV, ACTV)                                                              \
  do {                                                                                  \
    if (head > 0) RUN(CALL_C(DV, ACTV, first, head, out));                              \
    if (box > 0) RUN(CALL_B(DV, ACTV));                                                 \
    if (tail > 0) RUN(CALL_C(DV, ACTV, first + head + box, tail, out + head + box));    \
    return MRI_OK;                                                                      \
  } while (0)
  if (dim == 3) {
    if (act == MRI_ACT_GELU) DISPATCH(3, MRI_ACT_GELU);
    DISPATCH(3, MRI_ACT_RELU);
  }
  if (act == MRI_ACT_GELU) DISPATCH(4, MRI_ACT_GELU);
  DISPATCH(4, MRI_ACT_RELU);
#undef DISPATCH
#undef RUN
#undef CALL_B
#undef CALL_C
}

}  // namespace mri

using namespace mri;

extern "C" int mri_hashdecoder_forward(const float* x, int64_t n, int dim, const float* tables, const mri_level_t* host_levels,
                                       int n_levels, int n_features, int k0, int h, const float* w1, const float* b1,
                                       const float* w2, const float* b2, int act1, int act2, float* enc, float* y, float* pre2,
                                       void* stream) {
  if (n < 0) return fail(MRI_ERR_INVALID, "hashdecoder_forward: negative n");
  if (n == 0) return MRI_OK;
  if (!x || !tables || !host_levels || !w1 || !b1 || !w2 || !b2 || !y)
    return fail(MRI_ERR_INVALID, "hashdecoder_forward: null pointer");
  if (k0 != 2 * n_levels || !geometry_ok(dim, n_levels, n_features, h, act1))
    return fail(MRI_ERR_UNSUPPORTED, "hashdecoder_forward: fused kernel covers F=2, L=16, H=64, dim 3/4, GELU/ReLU "
                                     "(got F=%d L=%d H=%d dim=%d act=%d)", n_features, n_levels, h, dim, act1);
  const uintptr_t need = dim == 4 ? 15 : 3;
  if ((reinterpret_cast<uintptr_t>(x) & need) || (reinterpret_cast<uintptr_t>(tables) & 15) || (reinterpret_cast<uintptr_t>(enc) & 15))
    return fail(MRI_ERR_INVALID, "hashdecoder_forward: misaligned pointer");
  for (int l = 0; l < n_levels; ++l)
    if (host_levels[l].offset % 2) return fail(MRI_ERR_INVALID, "hashdecoder_forward: level %d offset not aligned", l);
  LevelTable T;
  int st = make_level_table(host_levels, n_levels, dim, &T);
  if (st != MRI_OK) return st;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define CALL(DV, ACTV) launch_fused_fwd<DV, ACTV>(BatchCoords<DV>{x}, n, tables, T, w1, b1, w2, b2, act2, enc, y, pre2, s)
  if (dim == 3) return act1 == MRI_ACT_GELU ? CALL(3, MRI_ACT_GELU) : CALL(3, MRI_ACT_RELU);
  return act1 == MRI_ACT_GELU ? CALL(4, MRI_ACT_GELU) : CALL(4, MRI_ACT_RELU);
#undef CALL
}
