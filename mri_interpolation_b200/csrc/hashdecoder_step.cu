// The whole training step of HashMLP under the mean-squared-error loss as ONE kernel: gather + decoder forward + loss +
// decoder backward + table scatter (BaseMLP.training_step models.py:61-66 on HashMLP.forward :741-744, and the autograd of
// both).  The kernel is the fused backward (hashdecoder_bwd_impl.cuh) instantiated with STEP = true; headline geometry
// F = 2, L = 16, hidden 64, D = 3 / 4, GELU / ReLU.  Gradients ACCUMULATE into the caller's buffers like every backward
// entry point of this library; *loss (device) receives sum_i (y_i - target_i)^2 * inv_count.
#include "hashdecoder.cuh"
#include "hashdecoder_bwd_impl.cuh"

using namespace mri;

extern "C" int mri_hashmlp_mse_step_supported(int dim, int n_levels, int n_features, int h, int act1) {
  return (fused_geometry_supported(dim, n_levels, n_features, h, act1) && fused_geometry_is_headline(n_levels, h)) ? 1 : 0;
}

extern "C" int mri_hashmlp_mse_step(const float* x, const float* target, int64_t n, int dim, const float* tables,
                                    const mri_level_t* host_levels, int n_levels, int n_features, int k0, int h, const float* w1,
                                    const float* b1, const float* w2, const float* b2, int act1, int act2, float inv_count,
                                    float* grad_tables, const mri_level_t* host_grad_levels, float* grad_w1, float* grad_b1,
                                    float* grad_w2, float* grad_b2, float* loss, float* y, void* stream) {
  if (n < 0) return fail(MRI_ERR_INVALID, "hashmlp_mse_step: negative n");
  if (n == 0) return MRI_OK;
  if (!x || !target || !tables || !host_levels || !host_grad_levels || !w1 || !b1 || !w2 || !b2 || !grad_tables || !grad_w1 || !grad_b1 ||
      !grad_w2 || !grad_b2 || !loss)
    return fail(MRI_ERR_INVALID, "hashmlp_mse_step: null pointer");
  if (k0 != 2 * n_levels || !mri_hashmlp_mse_step_supported(dim, n_levels, n_features, h, act1))
    return fail(MRI_ERR_UNSUPPORTED, "hashmlp_mse_step: fused step covers F=2, L=16, H=64, dim 3/4, GELU/ReLU "
                                     "(got F=%d L=%d H=%d dim=%d act=%d)", n_features, n_levels, h, dim, act1);
  const uintptr_t need = dim == 4 ? 15 : 3;
  if ((reinterpret_cast<uintptr_t>(x) & need) || (reinterpret_cast<uintptr_t>(tables) & 15) || (reinterpret_cast<uintptr_t>(grad_tables) & 15))
    return fail(MRI_ERR_INVALID, "hashmlp_mse_step: misaligned pointer");
  // the parameter tables and their gradient buffers are two arenas with (possibly) different level offsets: the kernel
  // takes ONE level table, so both must share the layout (true for the flat arenas: identical views of two buffers)
  for (int l = 0; l < n_levels; ++l) {
    if (host_levels[l].offset % 2) return fail(MRI_ERR_INVALID, "hashmlp_mse_step: level %d offset not aligned", l);
    if (host_levels[l].offset != host_grad_levels[l].offset || host_levels[l].rows != host_grad_levels[l].rows)
      return fail(MRI_ERR_UNSUPPORTED, "hashmlp_mse_step: tables and gradient tables must share one level layout (level %d)", l);
  }
  LevelTable T;
  int st = make_level_table(host_levels, n_levels, dim, &T);
  if (st != MRI_OK) return st;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static const bool merge = [] { const char* e = getenv("MRI_BWD_MERGE_LEVELS"); return !e || atoi(e) > 0; }();
#define CALL(DV, ACTV, MV) launch_fused_step<DV, 32, 64, ACTV, MV>(x, target, n, tables, T, w1, b1, w2, b2, ACTV, act2, inv_count, grad_tables, grad_w1, grad_b1, grad_w2, grad_b2, loss, y, s)
#define CALL_M(DV, ACTV) (merge ? CALL(DV, ACTV, 2) : CALL(DV, ACTV, 0))
  if (dim == 4 && act1 == MRI_ACT_GELU) return CALL_M(4, MRI_ACT_GELU);
  if (dim == 4 && act1 == MRI_ACT_RELU) return CALL_M(4, MRI_ACT_RELU);
  if (dim == 3 && act1 == MRI_ACT_GELU) return CALL_M(3, MRI_ACT_GELU);
  return CALL_M(3, MRI_ACT_RELU);
#undef CALL_M
#undef CALL
}
