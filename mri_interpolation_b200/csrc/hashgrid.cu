// Multiresolution hash-grid encoding kernels for sm_100a (B200).
//
//   hashgrid_fwd_kernel   K1: two lanes per (coordinate, level), each walking half of the 2^D corners with
//                         its gathers issued back-to-back (2^(D-1) independent LDG.64/128 in flight per
//                         thread), no intermediates in memory.  grid = (ceil(n/128), L): blocks are scheduled
//                         level-major so one level's table is the L2/L1 working set at a time.
//   hashgrid_bwd_kernel   K2: recomputes hashes/weights and scatters w*dOut with vector
//                         reductions red.global.add.v2/v4.f32 (SASS REDG.E.ADD.F32x2/x4).
//   hashgrid_corners_kernel  parity probe: dumps hashes and weights in the reference's corner order.
//
// Arithmetic follows encoding.py:108-128 (see common.cuh::make_cell); hash indices are bit-exact.
#include "common.cuh"
#include "hash_device.cuh"

namespace mri {

namespace {

// SINK = true is the same kernel with the row of every gather also written out (mri_hashgrid_forward_rows: the parity
// tests assert the indices of the path that ships, not of a separate probe); SINK = false compiles the sink away.
template <int D, int F, bool SINK = false>
__global__ void __launch_bounds__(256) hashgrid_fwd_kernel(const float* __restrict__ x, const float* __restrict__ tables,
                                                           const __grid_constant__ LevelTable T, int64_t n,
                                                           int out_stride, float* __restrict__ out,
                                                           uint32_t* __restrict__ rows_out = nullptr) {
  const int level = blockIdx.y;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 128 + (threadIdx.x >> 1);
  const int b0 = threadIdx.x & 1;
  const bool live = i < n;
  const LevelDev& lv = T.lv[level];
  float xv[D];
  load_coord<D>(x, live ? i : 0, xv);
  const Cell<D> cell = make_cell<D>(xv, lv);
  const float* __restrict__ tbl = tables + lv.offset;
  uint32_t* sink = nullptr;
  if constexpr (SINK) sink = live ? rows_out + (i * gridDim.y + level) * (1 << D) : nullptr;
  Feat<F> acc = lv.is_pow2 ? encode_half_level<D, F, true>(cell, b0, lv, tbl, sink) : encode_half_level<D, F, false>(cell, b0, lv, tbl, sink);
#pragma unroll
  for (int f = 0; f < F; ++f) acc.v[f] += __shfl_xor_sync(0xffffffffu, acc.v[f], 1);
  if (live && b0 == 0) store_feat<F>(out + i * out_stride + level * F, acc);
}

template <int D, int F>
__global__ void __launch_bounds__(256) hashgrid_bwd_kernel(const float* __restrict__ x, const float* __restrict__ grad_out,
                                                           const __grid_constant__ LevelTable T, int64_t n,
                                                           int out_stride, int level_begin, float* __restrict__ grad_tables) {
  const int level = blockIdx.y + level_begin;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 128 + (threadIdx.x >> 1);
  const int b0 = threadIdx.x & 1;
  if (i >= n) return;
  const LevelDev& lv = T.lv[level];
  float xv[D];
  load_coord<D>(x, i, xv);
  const Cell<D> cell = make_cell<D>(xv, lv);
  float* tbl = grad_tables + lv.offset;
  const Feat<F> g = gather_row<F>(grad_out + i * out_stride + level * F);
  if (lv.is_pow2) scatter_half_level<D, F, true>(cell, b0, lv, tbl, g);
  else scatter_half_level<D, F, false>(cell, b0, lv, tbl, g);
}

template <int D>
__global__ void __launch_bounds__(256) hashgrid_corners_kernel(const float* __restrict__ x, const __grid_constant__ LevelTable T,
                                                               int64_t n, int n_levels, uint32_t* __restrict__ hashes,
                                                               float* __restrict__ weights) {
  constexpr int C = 1 << D;
  const int level = blockIdx.y;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const LevelDev& lv = T.lv[level];
  float xv[D];
  load_coord<D>(x, i, xv);
  const Cell<D> cell = make_cell<D>(xv, lv);
  const int64_t base = (i * n_levels + level) * C;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    if (hashes)
      hashes[base + c] = lv.is_pow2 ? wrap_rows<true>(corner_hash<D>(cell, c), lv) : wrap_rows<false>(corner_hash<D>(cell, c), lv);
    if (weights) weights[base + c] = corner_weight<D>(cell, c);
  }
}

template <int D, int F>
int launch_fwd(const float* x, const float* tables, const LevelTable& T, int64_t n, int n_levels, float* out,
               cudaStream_t s) {
  dim3 grid(static_cast<unsigned>((n + 127) / 128), n_levels);  // 2 lanes per coordinate
  hashgrid_fwd_kernel<D, F><<<grid, 256, 0, s>>>(x, tables, T, n, n_levels * F, out);
  MRI_LAUNCH_OK("hashgrid_fwd_kernel");
  return MRI_OK;
}
template <int D, int F>
int launch_fwd_rows(const float* x, const float* tables, const LevelTable& T, int64_t n, int n_levels, float* out,
                    uint32_t* rows_out, cudaStream_t s) {
  dim3 grid(static_cast<unsigned>((n + 127) / 128), n_levels);
  hashgrid_fwd_kernel<D, F, true><<<grid, 256, 0, s>>>(x, tables, T, n, n_levels * F, out, rows_out);
  MRI_LAUNCH_OK("hashgrid_fwd_kernel<rows>");
  return MRI_OK;
}
template <int D, int F>
int launch_bwd(const float* x, const float* go, const LevelTable& T, int64_t n, int n_levels, int level_begin,
               int level_count, float* gt, cudaStream_t s) {
  dim3 grid(static_cast<unsigned>((n + 127) / 128), level_count);  // 2 lanes per coordinate
  hashgrid_bwd_kernel<D, F><<<grid, 256, 0, s>>>(x, go, T, n, n_levels * F, level_begin, gt);
  MRI_LAUNCH_OK("hashgrid_bwd_kernel");
  return MRI_OK;
}

int check_common(const void* x, int64_t n, int dim, const mri_level_t* lv, int n_levels, int n_features) {
  if (n < 0) return fail(MRI_ERR_INVALID, "hashgrid: negative n");
  if ((!x && n > 0) || !lv) return fail(MRI_ERR_INVALID, "hashgrid: null pointer");
  if (dim < 2 || dim > MRI_MAX_DIM) return fail(MRI_ERR_UNSUPPORTED, "hashgrid: dim %d not in 2..4", dim);
  if (n_levels < 1 || n_levels > MRI_MAX_LEVELS)
    return fail(MRI_ERR_UNSUPPORTED, "hashgrid: n_levels %d not in 1..%d", n_levels, MRI_MAX_LEVELS);
  if (n_features != 1 && n_features != 2 && n_features != 4 && n_features != 8)
    return fail(MRI_ERR_UNSUPPORTED, "hashgrid: n_features %d not in {1,2,4,8}", n_features);
  const uintptr_t need = dim == 4 ? 15 : dim == 2 ? 7 : 3;  // float4 / float2 / scalar coordinate loads
  if ((reinterpret_cast<uintptr_t>(x) & need) != 0)
    return fail(MRI_ERR_INVALID, "hashgrid: x must be %d-byte aligned for dim=%d", static_cast<int>(need + 1), dim);
  return MRI_OK;
}

#define MRI_DISPATCH_DF(dim, nf, CALL)                                   \
  switch ((dim) * 16 + (nf)) {                                           \
    case 2 * 16 + 1: return CALL(2, 1);                                  \
    case 2 * 16 + 2: return CALL(2, 2);                                  \
    case 2 * 16 + 4: return CALL(2, 4);                                  \
    case 2 * 16 + 8: return CALL(2, 8);                                  \
    case 3 * 16 + 1: return CALL(3, 1);                                  \
    case 3 * 16 + 2: return CALL(3, 2);                                  \
    case 3 * 16 + 4: return CALL(3, 4);                                  \
    case 3 * 16 + 8: return CALL(3, 8);                                  \
    case 4 * 16 + 1: return CALL(4, 1);                                  \
    case 4 * 16 + 2: return CALL(4, 2);                                  \
    case 4 * 16 + 4: return CALL(4, 4);                                  \
    case 4 * 16 + 8: return CALL(4, 8);                                  \
    default: return fail(MRI_ERR_UNSUPPORTED, "hashgrid: no kernel for dim=%d F=%d", (dim), (nf)); \
  }

}  // namespace

int make_level_table(const mri_level_t* host_levels, int n_levels, int dim, LevelTable* out) {
  memset(out, 0, sizeof(*out));
  for (int l = 0; l < n_levels; ++l) {
    const mri_level_t& h = host_levels[l];
    if (h.rows == 0) return fail(MRI_ERR_INVALID, "hashgrid: level %d has 0 rows", l);
    // exact_mod's estimate leaves a remainder below 2 * rows, which must fit 32 bits for non-power-of-two tables
    if ((h.rows & (h.rows - 1)) != 0 && h.rows > 0x80000000u)
      return fail(MRI_ERR_UNSUPPORTED, "hashgrid: level %d has %u rows; non-power-of-two tables are limited to 2^31 rows", l, h.rows);
    LevelDev& d = out->lv[l];
    for (int a = 0; a < MRI_MAX_DIM; ++a) d.res[a] = a < dim ? h.resolution[a] : 0.0f;
    d.rows = h.rows;
    d.is_pow2 = (h.rows & (h.rows - 1)) == 0 ? 1u : 0u;
    d.pow2_mask = h.rows - 1;
    d.magic = h.rows > 1 ? static_cast<uint32_t>((uint64_t{1} << 32) / h.rows) : 0xffffffffu;
    d.offset = h.offset;
  }
  return MRI_OK;
}

}  // namespace mri

using namespace mri;

extern "C" int mri_hashgrid_forward(const float* x, int64_t n, int dim, const float* tables,
                                    const mri_level_t* host_levels, int n_levels, int n_features, float* out,
                                    void* stream) {
  int st = check_common(x, n, dim, host_levels, n_levels, n_features);
  if (st != MRI_OK) return st;
  if (n == 0) return MRI_OK;  // empty batch: nothing to do (pointers of empty tensors may be null)
  if (!tables || !out) return fail(MRI_ERR_INVALID, "hashgrid_forward: null tables/out");
  if ((reinterpret_cast<uintptr_t>(tables) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return fail(MRI_ERR_INVALID, "hashgrid_forward: tables/out must be 16-byte aligned");
  for (int l = 0; l < n_levels; ++l)
    if (host_levels[l].offset % (n_features >= 4 ? 4 : n_features))
      return fail(MRI_ERR_INVALID, "hashgrid_forward: level %d offset not aligned to the feature vector", l);
  if (n == 0) return MRI_OK;
  LevelTable T;
  st = make_level_table(host_levels, n_levels, dim, &T);
  if (st != MRI_OK) return st;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define CALL(D, F) launch_fwd<D, F>(x, tables, T, n, n_levels, out, s)
  MRI_DISPATCH_DF(dim, n_features, CALL)
#undef CALL
}

extern "C" int mri_hashgrid_forward_rows(const float* x, int64_t n, int dim, const float* tables,
                                         const mri_level_t* host_levels, int n_levels, int n_features, float* out,
                                         uint32_t* rows_out, void* stream) {
  int st = check_common(x, n, dim, host_levels, n_levels, n_features);
  if (st != MRI_OK) return st;
  if (n == 0) return MRI_OK;
  if (!tables || !out || !rows_out) return fail(MRI_ERR_INVALID, "hashgrid_forward_rows: null tables/out/rows_out");
  if ((reinterpret_cast<uintptr_t>(tables) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return fail(MRI_ERR_INVALID, "hashgrid_forward_rows: tables/out must be 16-byte aligned");
  for (int l = 0; l < n_levels; ++l)
    if (host_levels[l].offset % (n_features >= 4 ? 4 : n_features))
      return fail(MRI_ERR_INVALID, "hashgrid_forward_rows: level %d offset not aligned to the feature vector", l);
  LevelTable T;
  st = make_level_table(host_levels, n_levels, dim, &T);
  if (st != MRI_OK) return st;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define CALL(D, F) launch_fwd_rows<D, F>(x, tables, T, n, n_levels, out, rows_out, s)
  MRI_DISPATCH_DF(dim, n_features, CALL)
#undef CALL
}

extern "C" int mri_hashgrid_backward_levels(const float* x, int64_t n, int dim, const float* grad_out, float* grad_tables,
                                            const mri_level_t* host_levels, int n_levels, int n_features, int level_begin,
                                            int level_count, void* stream) {
  int st = check_common(x, n, dim, host_levels, n_levels, n_features);
  if (st != MRI_OK) return st;
  if (level_begin < 0 || level_count < 0 || level_begin + level_count > n_levels)
    return fail(MRI_ERR_INVALID, "hashgrid_backward: level range [%d, +%d) outside 0..%d", level_begin, level_count, n_levels);
  if (level_count == 0) return MRI_OK;
  if (n == 0) return MRI_OK;
  if (!grad_out || !grad_tables) return fail(MRI_ERR_INVALID, "hashgrid_backward: null grad pointer");
  if ((reinterpret_cast<uintptr_t>(grad_tables) & 15) || (reinterpret_cast<uintptr_t>(grad_out) & 15))
    return fail(MRI_ERR_INVALID, "hashgrid_backward: grads must be 16-byte aligned");
  for (int l = 0; l < n_levels; ++l)
    if (host_levels[l].offset % (n_features >= 4 ? 4 : n_features))
      return fail(MRI_ERR_INVALID, "hashgrid_backward: level %d offset not aligned to the feature vector", l);
  if (n == 0) return MRI_OK;
  LevelTable T;
  st = make_level_table(host_levels, n_levels, dim, &T);
  if (st != MRI_OK) return st;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define CALL(D, F) launch_bwd<D, F>(x, grad_out, T, n, n_levels, level_begin, level_count, grad_tables, s)
  MRI_DISPATCH_DF(dim, n_features, CALL)
#undef CALL
}

extern "C" int mri_hashgrid_backward(const float* x, int64_t n, int dim, const float* grad_out, float* grad_tables,
                                     const mri_level_t* host_levels, int n_levels, int n_features, void* stream) {
  return mri_hashgrid_backward_levels(x, n, dim, grad_out, grad_tables, host_levels, n_levels, n_features, 0, n_levels, stream);
}

extern "C" int mri_hashgrid_corners(const float* x, int64_t n, int dim, const mri_level_t* host_levels, int n_levels,
                                    uint32_t* hashes, float* weights, void* stream) {
  int st = check_common(x, n, dim, host_levels, n_levels, 1);
  if (st != MRI_OK) return st;
  if (n == 0) return MRI_OK;
  LevelTable T;
  st = make_level_table(host_levels, n_levels, dim, &T);
  if (st != MRI_OK) return st;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid(static_cast<unsigned>((n + 255) / 256), n_levels);
  switch (dim) {
    case 2: hashgrid_corners_kernel<2><<<grid, 256, 0, s>>>(x, T, n, n_levels, hashes, weights); break;
    case 3: hashgrid_corners_kernel<3><<<grid, 256, 0, s>>>(x, T, n, n_levels, hashes, weights); break;
    default: hashgrid_corners_kernel<4><<<grid, 256, 0, s>>>(x, T, n, n_levels, hashes, weights); break;
  }
  MRI_LAUNCH_OK("hashgrid_corners_kernel");
  return MRI_OK;
}
