// Backward of the whole HashMLP in ONE kernel: decoder backward on the tensor cores (decoder.cu's
// decoder2_mma_bwd_kernel arithmetic) with the hash-table scatter fused in (autograd of models.py:741-744 on top of
// encoding.py:127-128), for F = 2, L = 4 / 8 / 16 (K0 = 8 / 16 / 32), H = 64 / 128, D = 3 / 4.  Kernel template in
// hashdecoder_bwd_impl.cuh; this file holds the headline geometry (L = 16, H = 64) and the C entry points.
#include "hashdecoder.cuh"
#include "hashdecoder_bwd_impl.cuh"

using namespace mri;

extern "C" int mri_hashdecoder_backward(const float* x, int64_t n, int dim, const float* enc, int k0, int h, const float* w1,
                                        const float* b1, const float* w2, const float* pre2, const float* grad_y, int act1, int act2,
                                        float* grad_tables, const mri_level_t* host_levels, int n_levels, int n_features,
                                        float* grad_w1, float* grad_b1, float* grad_w2, float* grad_b2, void* stream) {
  if (n < 0) return fail(MRI_ERR_INVALID, "hashdecoder_backward: negative n");
  if (n == 0) return MRI_OK;
  if (!x || !enc || !w1 || !b1 || !w2 || !pre2 || !grad_y || !grad_tables || !host_levels || !grad_w1 || !grad_b1 || !grad_w2 || !grad_b2)
    return fail(MRI_ERR_INVALID, "hashdecoder_backward: null pointer");
  if (k0 != 2 * n_levels || !fused_geometry_supported(dim, n_levels, n_features, h, act1))
    return fail(MRI_ERR_UNSUPPORTED, "hashdecoder_backward: fused kernel covers F=2, L=4/8/16, H=64/128, dim 3/4, GELU/ReLU "
                                     "(got F=%d L=%d H=%d dim=%d act=%d)", n_features, n_levels, h, dim, act1);
  const uintptr_t need = dim == 4 ? 15 : 3;
  if ((reinterpret_cast<uintptr_t>(x) & need) || (reinterpret_cast<uintptr_t>(grad_tables) & 15) || (reinterpret_cast<uintptr_t>(enc) & 15))
    return fail(MRI_ERR_INVALID, "hashdecoder_backward: misaligned pointer");
  for (int l = 0; l < n_levels; ++l)
    if (host_levels[l].offset % 2) return fail(MRI_ERR_INVALID, "hashdecoder_backward: level %d offset not aligned", l);
  LevelTable T;
  int st = make_level_table(host_levels, n_levels, dim, &T);
  if (st != MRI_OK) return st;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // MRI_BWD_MERGE_LEVELS = 0 | 4 | 8: number of coarse levels whose axis-0 duplicates are merged before the reduction
  // (default 8; 0 switches the merging off for A/B runs)
  static const int merge_nt2 = [] {
    const char* e = getenv("MRI_BWD_MERGE_LEVELS");
    const int v = e ? atoi(e) : 8;
    return v <= 0 ? 0 : v <= 4 ? 1 : 2;
  }();
  if (!fused_geometry_is_headline(n_levels, h))
    return launch_fused_bwd_geo(enc, n, dim, k0, h, w1, b1, w2, pre2, grad_y, act1, act2, x, T, grad_tables, grad_w1, grad_b1, grad_w2,
                                grad_b2, merge_nt2, s);
#define CALL(DV, ACTV, MV, CV) launch_fused_bwd<DV, 32, 64, ACTV, MV, CV>(enc, n, w1, b1, w2, pre2, grad_y, ACTV, act2, x, T, grad_tables, grad_w1, grad_b1, grad_w2, grad_b2, s)
#define CALL_M(DV, ACTV) (merge_nt2 == 0 ? CALL(DV, ACTV, 0, true) : merge_nt2 == 1 ? CALL(DV, ACTV, 1, true) : CALL(DV, ACTV, 2, true))
  if (dim == 4 && act1 == MRI_ACT_GELU) return CALL_M(4, MRI_ACT_GELU);
  if (dim == 4 && act1 == MRI_ACT_RELU) return CALL_M(4, MRI_ACT_RELU);
  if (dim == 3 && act1 == MRI_ACT_GELU) return CALL_M(3, MRI_ACT_GELU);
  return CALL_M(3, MRI_ACT_RELU);
#undef CALL_M
#undef CALL
}

extern "C" int mri_hashdecoder_supported(int dim, int n_levels, int n_features, int h, int act1) {
  return fused_geometry_supported(dim, n_levels, n_features, h, act1) ? 1 : 0;
}
