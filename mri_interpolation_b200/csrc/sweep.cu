// Dense-grid query sweep (K7) for sm_100a.
//
// The reference materialises the whole (prod(shape), D) coordinate grid on the host
// (launcher.py:191-202) and pushes it through DataLoader batches.  Here coordinates are
// synthesised in-kernel from the flat C-order voxel index and D tiny per-axis vectors (the
// caller passes torch.linspace's output, so the floats are the reference's own), i.e. zero
// coordinate traffic.  hashmlp_sweep_kernel additionally fuses the all-level hash encoding with
// the 2-layer decoder: per voxel only the 4-byte result is written.
#include <stdlib.h>

#include "common.cuh"
#include "grid_device.cuh"

namespace mri {
namespace {

template <int D>
__global__ void __launch_bounds__(256) grid_coords_kernel(const float* __restrict__ axes, const GridDesc gd, int64_t first,
                                                          int64_t count, float* __restrict__ coords) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float v[D];
  voxel_coord<D>(axes, gd, first + i, v);
#pragma unroll
  for (int d = 0; d < D; ++d) coords[i * D + d] = v[d];
}

// Training batch from voxel indices: coordinates synthesised from the flat index (no (M, D) coordinate
// table in HBM), intensities gathered from the normalised volume (datamodules.py:130-172 semantics).
template <int D>
__global__ void __launch_bounds__(256) gather_voxels_kernel(const float* __restrict__ axes, const GridDesc gd,
                                                            const int64_t* __restrict__ index, const float* __restrict__ pixels,
                                                            int64_t count, float* __restrict__ coords, float* __restrict__ values) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int64_t v = __ldg(index + i);
  float xv[D];
  voxel_coord<D>(axes, gd, v, xv);
  if constexpr (D == 4) {
    reinterpret_cast<float4*>(coords)[i] = make_float4(xv[0], xv[1], xv[2], xv[3]);
  } else if constexpr (D == 2) {
    reinterpret_cast<float2*>(coords)[i] = make_float2(xv[0], xv[1]);
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) coords[i * D + d] = xv[d];
  }
  if (values) values[i] = __ldg(pixels + v);
}

template <int F>
__device__ __forceinline__ void gather_feat(const float* __restrict__ row, float (&r)[F]) {
  if constexpr (F == 1) {
    r[0] = __ldg(row);
  } else if constexpr (F == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(row));
    r[0] = t.x; r[1] = t.y;
  } else {
#pragma unroll
    for (int q = 0; q < F / 4; ++q) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(row) + q);
      r[4 * q] = t.x; r[4 * q + 1] = t.y; r[4 * q + 2] = t.z; r[4 * q + 3] = t.w;
    }
  }
}

// One thread per voxel: coords -> L levels x 2^D gathers -> first decoder layer accumulated level by
// level in registers (weights broadcast from shared memory) -> activation -> H->1 output layer.
template <int D, int F, int H, int ACT1>
__global__ void __launch_bounds__(128) hashmlp_sweep_kernel(const float* __restrict__ axes, const GridDesc gd, int64_t first,
                                                            int64_t count, const float* __restrict__ tables,
                                                            const __grid_constant__ LevelTable T, int n_levels,
                                                            const float* __restrict__ decoder, int last_act,
                                                            float* __restrict__ out) {
  constexpr int C = 1 << D;
  extern __shared__ __align__(16) float smem[];
  const int K0 = n_levels * F;
  float* w0t = smem;             // [K0][H]  (transposed: broadcast float4 reads along H)
  float* b0 = w0t + K0 * H;      // [H]
  float* w1 = b0 + H;            // [H]
  float* b1 = w1 + H;            // [1]
  for (int e = threadIdx.x; e < K0 * H; e += blockDim.x) {
    const int j = e / K0, kk = e - j * K0;  // decoder W0 is (H, K0) row-major
    w0t[kk * H + j] = __ldg(decoder + e);
  }
  for (int e = threadIdx.x; e < H; e += blockDim.x) {
    b0[e] = __ldg(decoder + K0 * H + e);
    w1[e] = __ldg(decoder + K0 * H + H + e);
  }
  if (threadIdx.x == 0) b1[0] = __ldg(decoder + K0 * H + 2 * H);
  __syncthreads();

  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    float xv[D];
    voxel_coord<D>(axes, gd, first + i, xv);
    float h[H];
#pragma unroll
    for (int j = 0; j < H; ++j) h[j] = b0[j];
#pragma unroll 1
    for (int l = 0; l < n_levels; ++l) {
      const LevelDev& lv = T.lv[l];
      const Cell<D> cell = make_cell<D>(xv, lv);
      const float* __restrict__ tbl = tables + lv.offset;
      float rows[C][F];
      if (lv.is_pow2) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const uint32_t hh = wrap_rows<true>(corner_hash<D>(cell, c), lv);
          gather_feat<F>(tbl + static_cast<size_t>(hh) * F, rows[c]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const uint32_t hh = wrap_rows<false>(corner_hash<D>(cell, c), lv);
          gather_feat<F>(tbl + static_cast<size_t>(hh) * F, rows[c]);
        }
      }
      float enc[F];
#pragma unroll
      for (int f = 0; f < F; ++f) enc[f] = 0.0f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float w = corner_weight<D>(cell, c);
#pragma unroll
        for (int f = 0; f < F; ++f) enc[f] = fmaf(rows[c][f], w, enc[f]);
      }
#pragma unroll
      for (int f = 0; f < F; ++f) {
        const float4* wrow = reinterpret_cast<const float4*>(w0t + (l * F + f) * H);
#pragma unroll
        for (int q = 0; q < H / 4; ++q) {
          const float4 w = wrow[q];
          h[4 * q + 0] = fmaf(enc[f], w.x, h[4 * q + 0]);
          h[4 * q + 1] = fmaf(enc[f], w.y, h[4 * q + 1]);
          h[4 * q + 2] = fmaf(enc[f], w.z, h[4 * q + 2]);
          h[4 * q + 3] = fmaf(enc[f], w.w, h[4 * q + 3]);
        }
      }
    }
    float y = b1[0];
#pragma unroll
    for (int j = 0; j < H; ++j) y = fmaf(activate<ACT1>(h[j], 1.0f), w1[j], y);  // compile-time activation: 3x less code
    out[i] = activate_rt(last_act, y, 1.0f);
  }
}

template <int D, int F, int H, int ACT1>
int launch_sweep_act(const float* axes, const GridDesc& gd, int64_t first, int64_t count, const float* tables,
                     const LevelTable& T, int n_levels, const float* decoder, int last_act, float* out, cudaStream_t s) {
  const size_t smem = (static_cast<size_t>(n_levels) * F * H + 2 * H + 4) * sizeof(float);
  if (smem > 48 * 1024)
    MRI_CUDA_OK(cudaFuncSetAttribute(hashmlp_sweep_kernel<D, F, H, ACT1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  int64_t blocks = (count + 127) / 128;
  const int64_t cap = 16LL * sm_count();
  if (blocks > cap) blocks = cap;
  hashmlp_sweep_kernel<D, F, H, ACT1><<<static_cast<int>(blocks), 128, smem, s>>>(axes, gd, first, count, tables, T, n_levels,
                                                                                  decoder, last_act, out);
  MRI_LAUNCH_OK("hashmlp_sweep_kernel");
  return MRI_OK;
}

template <int D, int F, int H>
int launch_sweep(const float* axes, const GridDesc& gd, int64_t first, int64_t count, const float* tables,
                 const LevelTable& T, int n_levels, const float* decoder, int act, int last_act, float* out,
                 cudaStream_t s) {
  if (act == MRI_ACT_GELU) return launch_sweep_act<D, F, H, MRI_ACT_GELU>(axes, gd, first, count, tables, T, n_levels, decoder, last_act, out, s);
  if (act == MRI_ACT_RELU) return launch_sweep_act<D, F, H, MRI_ACT_RELU>(axes, gd, first, count, tables, T, n_levels, decoder, last_act, out, s);
  return fail(MRI_ERR_UNSUPPORTED, "hashmlp_sweep: fused kernel covers GELU / ReLU hidden activations (got %d)", act);
}

template <int D, int F>
int dispatch_h(int H, const float* axes, const GridDesc& gd, int64_t first, int64_t count, const float* tables,
               const LevelTable& T, int n_levels, const float* decoder, int act, int last_act, float* out,
               cudaStream_t s) {
  switch (H) {
    case 16: return launch_sweep<D, F, 16>(axes, gd, first, count, tables, T, n_levels, decoder, act, last_act, out, s);
    case 32: return launch_sweep<D, F, 32>(axes, gd, first, count, tables, T, n_levels, decoder, act, last_act, out, s);
    case 64: return launch_sweep<D, F, 64>(axes, gd, first, count, tables, T, n_levels, decoder, act, last_act, out, s);
    case 128: return launch_sweep<D, F, 128>(axes, gd, first, count, tables, T, n_levels, decoder, act, last_act, out, s);
    default: return fail(MRI_ERR_UNSUPPORTED, "hashmlp_sweep: hidden width %d not in {16,32,64,128}", H);
  }
}

}  // namespace
}  // namespace mri

using namespace mri;

extern "C" int mri_grid_coords(const float* axes, const int32_t* host_shape, int dim, int64_t first, int64_t count,
                               float* coords, void* stream) {
  if (!axes || !host_shape || !coords) return fail(MRI_ERR_INVALID, "grid_coords: null pointer");
  if (dim < 1 || dim > MRI_MAX_DIM) return fail(MRI_ERR_UNSUPPORTED, "grid_coords: dim %d not in 1..4", dim);
  GridDesc gd;
  int64_t total;
  int st = make_grid_desc(host_shape, dim, &gd, &total);
  if (st != MRI_OK) return st;
  if (first < 0 || count < 0 || first + count > total) return fail(MRI_ERR_INVALID, "grid_coords: range outside the grid");
  if (count == 0) return MRI_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>((count + 255) / 256);
  switch (dim) {
    case 1: grid_coords_kernel<1><<<grid, 256, 0, s>>>(axes, gd, first, count, coords); break;
    case 2: grid_coords_kernel<2><<<grid, 256, 0, s>>>(axes, gd, first, count, coords); break;
    case 3: grid_coords_kernel<3><<<grid, 256, 0, s>>>(axes, gd, first, count, coords); break;
    default: grid_coords_kernel<4><<<grid, 256, 0, s>>>(axes, gd, first, count, coords); break;
  }
  MRI_LAUNCH_OK("grid_coords_kernel");
  return MRI_OK;
}

extern "C" int mri_gather_voxels(const float* axes, const int32_t* host_shape, int dim, const int64_t* index,
                                 int64_t count, const float* pixels, float* coords, float* values, void* stream) {
  if (!axes || !host_shape || !index || !coords) return fail(MRI_ERR_INVALID, "gather_voxels: null pointer");
  if ((values != nullptr) != (pixels != nullptr)) return fail(MRI_ERR_INVALID, "gather_voxels: pixels and values go together");
  if (dim < 1 || dim > MRI_MAX_DIM) return fail(MRI_ERR_UNSUPPORTED, "gather_voxels: dim %d not in 1..4", dim);
  if (count < 0) return fail(MRI_ERR_INVALID, "gather_voxels: negative count");
  GridDesc gd;
  int64_t total;
  int st = make_grid_desc(host_shape, dim, &gd, &total);
  if (st != MRI_OK) return st;
  if (count == 0) return MRI_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>((count + 255) / 256);
  switch (dim) {
    case 1: gather_voxels_kernel<1><<<grid, 256, 0, s>>>(axes, gd, index, pixels, count, coords, values); break;
    case 2: gather_voxels_kernel<2><<<grid, 256, 0, s>>>(axes, gd, index, pixels, count, coords, values); break;
    case 3: gather_voxels_kernel<3><<<grid, 256, 0, s>>>(axes, gd, index, pixels, count, coords, values); break;
    default: gather_voxels_kernel<4><<<grid, 256, 0, s>>>(axes, gd, index, pixels, count, coords, values); break;
  }
  MRI_LAUNCH_OK("gather_voxels_kernel");
  return MRI_OK;
}

extern "C" int mri_hashmlp_sweep(const float* axes, const int32_t* host_shape, int dim, int64_t first, int64_t count,
                                 const float* tables, const mri_level_t* host_levels, int n_levels, int n_features,
                                 const float* decoder, const int32_t* host_dims, int n_dense, int act, int last_act,
                                 float* out, void* stream) {
  if (!axes || !host_shape || !tables || !host_levels || !decoder || !host_dims || !out)
    return fail(MRI_ERR_INVALID, "hashmlp_sweep: null pointer");
  if (dim < 2 || dim > MRI_MAX_DIM) return fail(MRI_ERR_UNSUPPORTED, "hashmlp_sweep: dim %d not in 2..4", dim);
  if (n_levels < 1 || n_levels > MRI_MAX_LEVELS) return fail(MRI_ERR_UNSUPPORTED, "hashmlp_sweep: n_levels %d", n_levels);
  if (n_dense != 2 || host_dims[2] != 1)
    return fail(MRI_ERR_UNSUPPORTED, "hashmlp_sweep: fused kernel needs a 2-layer decoder with one output "
                                     "(got %d layers); use the unfused chunked sweep", n_dense);
  if (host_dims[0] != n_levels * n_features) return fail(MRI_ERR_INVALID, "hashmlp_sweep: decoder input != L*F");
  if ((reinterpret_cast<uintptr_t>(tables) & 15) != 0) return fail(MRI_ERR_INVALID, "hashmlp_sweep: tables must be 16-byte aligned");
  for (int l = 0; l < n_levels; ++l)
    if (host_levels[l].offset % (n_features >= 4 ? 4 : n_features))
      return fail(MRI_ERR_INVALID, "hashmlp_sweep: level %d offset not aligned to the feature vector", l);
  GridDesc gd;
  int64_t total;
  int st = make_grid_desc(host_shape, dim, &gd, &total);
  if (st != MRI_OK) return st;
  if (first < 0 || count < 0 || first + count > total) return fail(MRI_ERR_INVALID, "hashmlp_sweep: range outside the grid");
  if (count == 0) return MRI_OK;
  LevelTable T;
  st = make_level_table(host_levels, n_levels, dim, &T);
  if (st != MRI_OK) return st;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int H = host_dims[1];
  static const bool use_cuda_cores = getenv("MRI_SWEEP_CUDA_CORES") != nullptr;  // debugging/profiling switch
  if (!use_cuda_cores && sweep_mma_supported(dim, n_levels, n_features, H, act))
    return launch_sweep_mma(axes, gd, dim, n_levels, H, first, count, tables, T, decoder, act, last_act, out, s);
#define CALL(D, F) dispatch_h<D, F>(H, axes, gd, first, count, tables, T, n_levels, decoder, act, last_act, out, s)
  switch (dim * 16 + n_features) {
    case 2 * 16 + 1: return CALL(2, 1);
    case 2 * 16 + 2: return CALL(2, 2);
    case 2 * 16 + 4: return CALL(2, 4);
    case 3 * 16 + 1: return CALL(3, 1);
    case 3 * 16 + 2: return CALL(3, 2);
    case 3 * 16 + 4: return CALL(3, 4);
    case 4 * 16 + 1: return CALL(4, 1);
    case 4 * 16 + 2: return CALL(4, 2);
    case 4 * 16 + 4: return CALL(4, 4);
    default: return fail(MRI_ERR_UNSUPPORTED, "hashmlp_sweep: no fused kernel for dim=%d F=%d", dim, n_features);
  }
#undef CALL
}
