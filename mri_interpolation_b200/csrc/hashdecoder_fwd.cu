// Hash-grid encoder fused with the 2-layer decoder, forward direction (HashMLP.forward: models.py:712-739 on top of
// encoding.py:190-191), for F = 2, L = 4 / 8 / 16 (K0 = 8 / 16 / 32), H = 64 / 128, D = 3 / 4 (kernel template in
// hashdecoder_fwd_impl.cuh; this file holds the headline geometry L = 16, H = 64, the dense-sweep variants and the C entry).
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
#include "hashdecoder.cuh"
#include "hashdecoder_fwd_impl.cuh"

namespace mri {
namespace {

// headline geometry: compile-time activation here; the other F = 2 geometries live in hashdecoder_fwd_geo.cu
bool geometry_ok(int dim, int n_levels, int n_features, int h, int act) {
  return (dim == 3 || dim == 4) && n_levels == 16 && n_features == 2 && h == 64 && (act == MRI_ACT_GELU || act == MRI_ACT_RELU);
}

}  // namespace

bool sweep_mma_supported(int dim, int n_levels, int n_features, int h, int act) {
  return fused_geometry_supported(dim, n_levels, n_features, h, act);
}

int launch_sweep_mma(const float* axes, const GridDesc& gd, int dim, int n_levels, int h, int64_t first, int64_t count,
                     const float* tables, const LevelTable& T, const float* decoder, int act, int last_act, float* out,
                     cudaStream_t s) {
  if (!fused_geometry_is_headline(n_levels, h))  // the other F = 2 geometries: hashdecoder_fwd_geo.cu
    return launch_sweep_mma_geo(axes, gd, dim, 2 * n_levels, h, first, count, tables, T, decoder, act, last_act, out, s);
  if (dim == 3)
    return act == MRI_ACT_GELU ? sweep_fused<3, 32, 64, MRI_ACT_GELU>(axes, gd, first, count, tables, T, decoder, act, last_act, out, s)
                               : sweep_fused<3, 32, 64, MRI_ACT_RELU>(axes, gd, first, count, tables, T, decoder, act, last_act, out, s);
  return act == MRI_ACT_GELU ? sweep_fused<4, 32, 64, MRI_ACT_GELU>(axes, gd, first, count, tables, T, decoder, act, last_act, out, s)
                             : sweep_fused<4, 32, 64, MRI_ACT_RELU>(axes, gd, first, count, tables, T, decoder, act, last_act, out, s);
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
  if (k0 != 2 * n_levels || !fused_geometry_supported(dim, n_levels, n_features, h, act1))
    return fail(MRI_ERR_UNSUPPORTED, "hashdecoder_forward: fused kernel covers F=2, L=4/8/16, H=64/128, dim 3/4, GELU/ReLU "
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
  if (!geometry_ok(dim, n_levels, n_features, h, act1))
    return launch_fused_fwd_geo(x, n, dim, k0, h, tables, T, w1, b1, w2, b2, act1, act2, enc, y, pre2, s);
#define CALL(DV, ACTV) launch_fused_fwd<DV, 32, 64, ACTV>(BatchCoords<DV>{x}, n, tables, T, w1, b1, w2, b2, ACTV, act2, enc, y, pre2, s)
  if (dim == 3) return act1 == MRI_ACT_GELU ? CALL(3, MRI_ACT_GELU) : CALL(3, MRI_ACT_RELU);
  return act1 == MRI_ACT_GELU ? CALL(4, MRI_ACT_GELU) : CALL(4, MRI_ACT_RELU);
#undef CALL
}
