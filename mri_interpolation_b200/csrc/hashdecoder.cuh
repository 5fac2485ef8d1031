// Cross-file declarations of the fused HashMLP kernels (F = 2, L in {4, 8, 16}, H in {64, 128}, D in {3, 4}, GELU / ReLU).
// The headline geometry (L = 16, H = 64) is instantiated with a compile-time activation in hashdecoder_fwd.cu /
// hashdecoder_bwd.cu; the other five (K0, H) combinations live in hashdecoder_{fwd,bwd}_geo.cu with the activation chosen at
// run time, so that each translation unit compiles in about a minute and the files build in parallel.
#pragma once
#include "common.cuh"

namespace mri {

inline bool fused_geometry_supported(int dim, int n_levels, int n_features, int h, int act1) {
  return n_features == 2 && (n_levels == 4 || n_levels == 8 || n_levels == 16) && (h == 64 || h == 128) && (dim == 3 || dim == 4) &&
         (act1 == MRI_ACT_GELU || act1 == MRI_ACT_RELU);
}
inline bool fused_geometry_is_headline(int n_levels, int h) { return n_levels == 16 && h == 64; }

int launch_fused_fwd_geo(const float* x, int64_t n, int dim, int k0, int h, const float* tables, const LevelTable& T, const float* w1,
                         const float* b1, const float* w2, const float* b2, int act1, int act2, float* enc, float* y, float* pre2,
                         cudaStream_t s);
struct GridDesc;
int launch_sweep_mma_geo(const float* axes, const GridDesc& gd, int dim, int k0, int h, int64_t first, int64_t count,
                         const float* tables, const LevelTable& T, const float* decoder, int act, int last_act, float* out,
                         cudaStream_t s);
int launch_fused_bwd_geo(const float* enc, int64_t n, int dim, int k0, int h, const float* w1, const float* b1, const float* w2,
                         const float* pre2, const float* gy, int act1, int act2, const float* x, const LevelTable& T, float* grad_tables,
                         float* gw1, float* gb1, float* gw2, float* gb2, int merge_nt2, cudaStream_t s);

}  // namespace mri
