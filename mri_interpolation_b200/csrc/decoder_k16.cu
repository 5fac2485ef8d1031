// Stand-alone fused 2-layer decoder, input width K0 = 16: H in {32, 64} x GELU / ReLU x (mma.sync, CUDA-core) x (fwd, bwd).
#include "decoder_impl.cuh"

namespace mri {
MRI_DECODER_K_DEFINE(16)
}  // namespace mri
