// Dense-grid descriptor: coordinates of a voxel from its flat C-order index and the per-axis coordinate vectors
// (torch.linspace output, so the floats are the reference's own - launcher.py:191-202, utils.py:14-23).
#pragma once
#include "common.cuh"

namespace mri {

struct GridDesc {
  int shape[MRI_MAX_DIM];
  int axis_off[MRI_MAX_DIM];
};

template <int D>
__device__ __forceinline__ void voxel_coord(const float* __restrict__ axes, const GridDesc& gd, int64_t idx, float (&v)[D]) {
  if ((static_cast<uint64_t>(idx) >> 32) == 0) {  // 32-bit divisions for every grid below 2^32 voxels
    uint32_t rem = static_cast<uint32_t>(idx);
#pragma unroll
    for (int d = D - 1; d >= 1; --d) {
      const uint32_t q = rem / static_cast<uint32_t>(gd.shape[d]);
      v[d] = __ldg(axes + gd.axis_off[d] + (rem - q * static_cast<uint32_t>(gd.shape[d])));
      rem = q;
    }
    v[0] = __ldg(axes + gd.axis_off[0] + rem);
    return;
  }
  uint64_t rem = static_cast<uint64_t>(idx);
#pragma unroll
  for (int d = D - 1; d >= 1; --d) {
    const uint64_t q = rem / static_cast<uint32_t>(gd.shape[d]);
    const uint32_t r = static_cast<uint32_t>(rem - q * static_cast<uint32_t>(gd.shape[d]));
    v[d] = __ldg(axes + gd.axis_off[d] + r);
    rem = q;
  }
  v[0] = __ldg(axes + gd.axis_off[0] + static_cast<uint32_t>(rem));
}

inline int make_grid_desc(const int32_t* host_shape, int dim, GridDesc* gd, int64_t* total) {
  int off = 0;
  int64_t t = 1;
  for (int d = 0; d < MRI_MAX_DIM; ++d) {
    gd->shape[d] = 1;
    gd->axis_off[d] = 0;
  }
  for (int d = 0; d < dim; ++d) {
    if (host_shape[d] < 1) return fail(MRI_ERR_INVALID, "sweep: shape[%d] = %d", d, host_shape[d]);
    gd->shape[d] = host_shape[d];
    gd->axis_off[d] = off;
    off += host_shape[d];
    t *= host_shape[d];
  }
  *total = t;
  return MRI_OK;
}

// hashdecoder_fwd.cu / hashdecoder_fwd_geo.cu: tensor-core variant of the fused sweep (F = 2, L = 4 / 8 / 16, H = 64 / 128,
// D = 3 / 4, GELU / ReLU)
bool sweep_mma_supported(int dim, int n_levels, int n_features, int h, int act);
int launch_sweep_mma(const float* axes, const GridDesc& gd, int dim, int n_levels, int h, int64_t first, int64_t count,
                     const float* tables, const LevelTable& T, const float* decoder, int act, int last_act, float* out,
                     cudaStream_t s);

}  // namespace mri
