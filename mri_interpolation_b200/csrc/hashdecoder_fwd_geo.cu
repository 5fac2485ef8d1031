// Encoder + decoder forward fusion for the non-headline F = 2 geometries (see hashdecoder.cuh): the notebook's L = 8
// anisotropic V2 grid (K0 = 16, nb cell 37), 128-wide decoders (config/hash_config.json's n_neurons), L = 4.
#include "hashdecoder.cuh"
#include "hashdecoder_fwd_impl.cuh"

namespace mri {

int launch_fused_fwd_geo(const float* x, int64_t n, int dim, int k0, int h, const float* tables, const LevelTable& T, const float* w1,
                         const float* b1, const float* w2, const float* b2, int act1, int act2, float* enc, float* y, float* pre2,
                         cudaStream_t s) {
#define CALL(DV, KV, HV) \
  launch_fused_fwd<DV, KV, HV, ACT_RUNTIME>(BatchCoords<DV>{x}, n, tables, T, w1, b1, w2, b2, act1, act2, enc, y, pre2, s)
#define BY_DIM(KV, HV) return dim == 3 ? CALL(3, KV, HV) : CALL(4, KV, HV)
  switch (k0 * 1000 + h) {
    case 8 * 1000 + 64: BY_DIM(8, 64);
    case 16 * 1000 + 64: BY_DIM(16, 64);
    case 8 * 1000 + 128: BY_DIM(8, 128);
    case 16 * 1000 + 128: BY_DIM(16, 128);
    case 32 * 1000 + 128: BY_DIM(32, 128);
    default: return fail(MRI_ERR_UNSUPPORTED, "hashdecoder_forward: no fused kernel for K0=%d H=%d", k0, h);
  }
#undef BY_DIM
#undef CALL
}

int launch_sweep_mma_geo(const float* axes, const GridDesc& gd, int dim, int k0, int h, int64_t first, int64_t count,
                         const float* tables, const LevelTable& T, const float* decoder, int act, int last_act, float* out,
                         cudaStream_t s) {
#define SWEEP(DV, KV, HV) sweep_fused<DV, KV, HV, ACT_RUNTIME>(axes, gd, first, count, tables, T, decoder, act, last_act, out, s)
#define BY_DIM(KV, HV) return dim == 3 ? SWEEP(3, KV, HV) : SWEEP(4, KV, HV)
  switch (k0 * 1000 + h) {
    case 8 * 1000 + 64: BY_DIM(8, 64);
    case 16 * 1000 + 64: BY_DIM(16, 64);
    case 8 * 1000 + 128: BY_DIM(8, 128);
    case 16 * 1000 + 128: BY_DIM(16, 128);
    case 32 * 1000 + 128: BY_DIM(32, 128);
    default: return fail(MRI_ERR_UNSUPPORTED, "hashmlp_sweep: no tensor-core sweep kernel for K0=%d H=%d", k0, h);
  }
#undef BY_DIM
#undef SWEEP
}

}  // namespace mri
