// Fused 2-layer decoder of the hash-grid models (HashMLP: enc(K0) -> H -> 1, models.py:712-739 / nb cell 37).
//
// The layers are far too small for tensor-core tiles (K0 = L*F <= 64, H <= 64, one output), so the whole
// decoder runs per coordinate in registers with the weights broadcast from shared memory:
//   decoder2_fwd_kernel : y = act2(b2 + w2 . act1(W1 enc + b1)); only y and the scalar pre-activation of the
//                         output layer are written (8 B/coord) - the hidden layer is recomputed in backward.
//   decoder2_bwd_kernel : recompute hidden layer, dEnc = W1^T (dh * act1'), and the parameter gradients.
//                         dW1 (H x K0, a reduction over the batch) is formed per 128-coordinate chunk from
//                         shared-memory copies of dPre1 and enc, accumulated in registers across the
//                         persistent block's chunks and flushed once per block with red.global.add.
#include <stdlib.h>

#include "common.cuh"

namespace mri {
// defined in decoder_k16.cu / decoder_k32.cu / decoder_k64.cu (kernel templates: decoder_impl.cuh)
#define MRI_DECODER_K_DECL(K0V)                                                                                                  \
  int decoder2_forward_k##K0V(bool cuda_cores, int h, int act1, const float* enc, int64_t n, const float* w1, const float* b1,   \
                              const float* w2, const float* b2, int act2, float* y, float* pre2, cudaStream_t s);               \
  int decoder2_backward_k##K0V(bool cuda_cores, int h, int act1, const float* enc, int64_t n, const float* w1, const float* b1,  \
                               const float* w2, const float* pre2, const float* grad_y, int act2, float* grad_enc,             \
                               float* grad_w1, float* grad_b1, float* grad_w2, float* grad_b2, cudaStream_t s);
MRI_DECODER_K_DECL(16)
MRI_DECODER_K_DECL(32)
MRI_DECODER_K_DECL(64)
#undef MRI_DECODER_K_DECL
namespace {
bool shape_ok(int k0, int h) { return (k0 == 16 || k0 == 32 || k0 == 64) && (h == 32 || h == 64); }
}  // namespace
}  // namespace mri

using namespace mri;

extern "C" int mri_decoder2_supported(int k0, int h, int act1) {
  return (shape_ok(k0, h) && (act1 == MRI_ACT_GELU || act1 == MRI_ACT_RELU)) ? 1 : 0;
}

extern "C" int mri_decoder2_forward(const float* enc, int64_t n, int k0, int h, const float* w1, const float* b1,
                                    const float* w2, const float* b2, int act1, int act2, float* y, float* pre2,
                                    void* stream) {
  if (n < 0) return fail(MRI_ERR_INVALID, "decoder2_forward: negative n");
  if (n == 0) return MRI_OK;
  if (!enc || !w1 || !b1 || !w2 || !b2 || !y) return fail(MRI_ERR_INVALID, "decoder2_forward: null pointer");
  if (!mri_decoder2_supported(k0, h, act1))
    return fail(MRI_ERR_UNSUPPORTED, "decoder2_forward: no fused kernel for K0=%d H=%d act=%d", k0, h, act1);
  if (reinterpret_cast<uintptr_t>(enc) & 15) return fail(MRI_ERR_INVALID, "decoder2_forward: enc must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static const bool use_cuda_cores = getenv("MRI_DECODER_CUDA_CORES") != nullptr;  // debugging/profiling switch
  if (k0 == 16) return decoder2_forward_k16(use_cuda_cores, h, act1, enc, n, w1, b1, w2, b2, act2, y, pre2, s);
  if (k0 == 32) return decoder2_forward_k32(use_cuda_cores, h, act1, enc, n, w1, b1, w2, b2, act2, y, pre2, s);
  return decoder2_forward_k64(use_cuda_cores, h, act1, enc, n, w1, b1, w2, b2, act2, y, pre2, s);
}

extern "C" int mri_decoder2_backward(const float* enc, int64_t n, int k0, int h, const float* w1, const float* b1,
                                     const float* w2, const float* pre2, const float* grad_y, int act1, int act2,
                                     float* grad_enc, float* grad_w1, float* grad_b1, float* grad_w2, float* grad_b2,
                                     void* stream) {
  if (n < 0) return fail(MRI_ERR_INVALID, "decoder2_backward: negative n");
  if (n == 0) return MRI_OK;
  if (!enc || !w1 || !b1 || !w2 || !pre2 || !grad_y || !grad_enc || !grad_w1 || !grad_b1 || !grad_w2 || !grad_b2)
    return fail(MRI_ERR_INVALID, "decoder2_backward: null pointer");
  if (!mri_decoder2_supported(k0, h, act1))
    return fail(MRI_ERR_UNSUPPORTED, "decoder2_backward: no fused kernel for K0=%d H=%d act=%d", k0, h, act1);
  if ((reinterpret_cast<uintptr_t>(enc) | reinterpret_cast<uintptr_t>(grad_enc)) & 15)
    return fail(MRI_ERR_INVALID, "decoder2_backward: enc/grad_enc must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static const bool use_cuda_cores = getenv("MRI_DECODER_CUDA_CORES") != nullptr;  // debugging/profiling switch
  if (k0 == 16)
    return decoder2_backward_k16(use_cuda_cores, h, act1, enc, n, w1, b1, w2, pre2, grad_y, act2, grad_enc, grad_w1, grad_b1, grad_w2, grad_b2, s);
  if (k0 == 32)
    return decoder2_backward_k32(use_cuda_cores, h, act1, enc, n, w1, b1, w2, pre2, grad_y, act2, grad_enc, grad_w1, grad_b1, grad_w2, grad_b2, s);
  return decoder2_backward_k64(use_cuda_cores, h, act1, enc, n, w1, b1, w2, pre2, grad_y, act2, grad_enc, grad_w1, grad_b1, grad_w2, grad_b2, s);
}
