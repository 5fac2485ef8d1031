// Fused HashMLP backward for the non-headline F = 2 geometries (see hashdecoder.cuh).  Axis-0 run merging covers the
// 8 coarsest levels (all of them for L = 4 / 8) unless MRI_BWD_MERGE_LEVELS=0 switches it off.
#include "hashdecoder.cuh"
#include "hashdecoder_bwd_impl.cuh"

namespace mri {

int launch_fused_bwd_geo(const float* enc, int64_t n, int dim, int k0, int h, const float* w1, const float* b1, const float* w2,
                         const float* pre2, const float* gy, int act1, int act2, const float* x, const LevelTable& T, float* grad_tables,
                         float* gw1, float* gb1, float* gw2, float* gb2, int merge_nt2, cudaStream_t s) {
#define CALL(DV, KV, HV, MV) \
  launch_fused_bwd<DV, KV, HV, ACT_RUNTIME, MV, true>(enc, n, w1, b1, w2, pre2, gy, act1, act2, x, T, grad_tables, gw1, gb1, gw2, gb2, s)
#define BY_DIM(KV, HV, MV) return dim == 3 ? (merge_nt2 ? CALL(3, KV, HV, MV) : CALL(3, KV, HV, 0)) : (merge_nt2 ? CALL(4, KV, HV, MV) : CALL(4, KV, HV, 0))
  switch (k0 * 1000 + h) {
    case 8 * 1000 + 64: BY_DIM(8, 64, 1);
    case 16 * 1000 + 64: BY_DIM(16, 64, 2);
    case 8 * 1000 + 128: BY_DIM(8, 128, 1);
    case 16 * 1000 + 128: BY_DIM(16, 128, 2);
    case 32 * 1000 + 128: BY_DIM(32, 128, 2);
    default: return fail(MRI_ERR_UNSUPPORTED, "hashdecoder_backward: no fused kernel for K0=%d H=%d", k0, h);
  }
#undef BY_DIM
#undef CALL
}

}  // namespace mri
