/* mri_b200.h - C ABI of the B200-native coordinate-network hot path.
 *
 * Drop-in boundary for Benjamin-Fouquet/mri_interpolation.  The reference has no FFI
 * layer: its hot path is the torch arithmetic inside the nn.Module classes of
 * encoding.py / models.py.  Each entry point below replaces one such piece (cited
 * as reference file:line); the Python classes in mri_interpolation_b200/ keep the
 * reference's constructor/forward surface and bind these symbols through ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; buffers are owned by the caller;
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*);
 *     no hidden synchronisation, no allocation, re-entrant across streams/devices;
 *   - return value: 0 = ok, <0 = mri_status error; mri_last_error() gives the message
 *     (thread-local); nothing throws or exits across this boundary;
 *   - row-major fp32 tensors, coordinates (n, dim), encodings (n, n_levels*n_features).
 */
#ifndef MRI_B200_H
#define MRI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRI_B200_VERSION 100
#define MRI_MAX_DIM 4      /* hash grid input dims supported by the kernels: 2, 3, 4 (x,y,z,t) */
#define MRI_MAX_LEVELS 32

enum mri_status {
  MRI_OK = 0,
  MRI_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, ...) */
  MRI_ERR_UNSUPPORTED = -2, /* shape / dim / feature count without a kernel */
  MRI_ERR_CUDA = -3         /* CUDA runtime error (message in mri_last_error) */
};

enum mri_activation {
  MRI_ACT_IDENTITY = 0,
  MRI_ACT_SINE = 1, /* sin(w0 * pre)          models.py:108-114 */
  MRI_ACT_GELU = 2, /* exact erf GELU         nn.GELU default, models.py:672 */
  MRI_ACT_RELU = 3  /* BaseMLP default stack  models.py:31 */
};

/* One resolution level of the multiresolution hash grid.
 * encoding.py:81-96 (_HashGrid) / 194-214 (_HashGridV2): per-axis resolution (V1: all
 * equal), `rows` = hashmap_size (table rows, each n_features floats), `offset` = position
 * of the level's table inside the flat table arena, in floats. */
typedef struct mri_level {
  float resolution[MRI_MAX_DIM];
  uint32_t rows;
  uint32_t reserved;
  uint64_t offset;
} mri_level_t;

int mri_version(void);
const char* mri_last_error(void);
/* number of SMs of the current device (grid sizing on the host side) */
int mri_sm_count(void);

/* ---- multiresolution hash grid ------------------------------------------------------- */

/* Forward of MultiResHashGrid / MultiResHashGridV2: encoding.py:108-128 (per level) and
 * :190-191 (concat).  out[i, l*F + f] = sum_c tables[level l][hash(corner c)][f] * w_c. */
int mri_hashgrid_forward(const float* x, int64_t n, int dim, const float* tables,
                         const mri_level_t* host_levels, int n_levels, int n_features,
                         float* out, void* stream);

/* Backward w.r.t. the tables (autograd of encoding.py:127-128): grad_tables[h_c] += w_c *
 * grad_out, ACCUMULATING into grad_tables (same arena layout as `tables`; caller zero-fills). */
int mri_hashgrid_backward(const float* x, int64_t n, int dim, const float* grad_out,
                          float* grad_tables, const mri_level_t* host_levels, int n_levels,
                          int n_features, void* stream);

/* Same, restricted to levels [level_begin, level_begin + level_count): lets the caller launch the backward
 * in level groups and start the data-parallel all-reduce of a finished group while the next one runs. */
int mri_hashgrid_backward_levels(const float* x, int64_t n, int dim, const float* grad_out,
                                 float* grad_tables, const mri_level_t* host_levels, int n_levels,
                                 int n_features, int level_begin, int level_count, void* stream);

/* mri_hashgrid_forward run through the SAME gather code with the addressed table row of every corner also stored:
 * rows_out (n, L, 2^dim) uint32 in the reference's corner order (bit d of the corner number = upper cell on axis d,
 * encoding.py:101-106).  Parity instrumentation: the indices asserted bit-exact are those of the path that ships. */
int mri_hashgrid_forward_rows(const float* x, int64_t n, int dim, const float* tables,
                              const mri_level_t* host_levels, int n_levels, int n_features,
                              float* out, uint32_t* rows_out, void* stream);

/* Parity probe: corner hashes (n, L, 2^dim) uint32 and weights (n, L, 2^dim) f32 in the
 * reference's corner order (encoding.py:69-78, 101-106, 121-124). Either output may be NULL. */
int mri_hashgrid_corners(const float* x, int64_t n, int dim, const mri_level_t* host_levels,
                         int n_levels, uint32_t* hashes, float* weights, void* stream);

/* ---- dense layers (SIREN layer / decoder Linear) ---------------------------------------- */

/* y = act(x W^T + b): SirenLayer.forward models.py:153-156, decoder Linear+GELU
 * models.py:730-736.  x (n,k) row stride ldx; W (m,k); b (m) or NULL; y (n,m).
 * `pre` (n,m) or NULL receives the pre-activation needed by mri_dense_backward. */
int mri_dense_forward(const float* x, int64_t ldx, const float* w, const float* b, int64_t n,
                      int k, int m, int act, float w0, float* y, float* pre, void* stream);

/* Backward of the layer above.  dpre = grad_y * act'(pre) is written to `dpre` (n,m scratch);
 * grad_w (m,k) += dpre^T x; grad_b (m) += sum_n dpre; grad_x (n,k) = dpre W (skipped if NULL).
 * grad_b may be NULL. */
int mri_dense_backward(const float* x, int64_t ldx, const float* w, const float* pre,
                       const float* grad_y, int64_t n, int k, int m, int act, float w0,
                       float* dpre, float* grad_x, float* grad_w, float* grad_b, void* stream);

/* ---- fused 2-layer decoder of the hash-grid models ------------------------------------------------- */

/* 1 if the fused kernels take enc width k0, hidden width h and hidden activation act1. */
int mri_decoder2_supported(int k0, int h, int act1);

/* y = act2(b2 + w2 . act1(W1 enc + b1)): the HashMLP decoder (models.py:712-739, intended forward nb cell 37)
 * for n_layers = 2, dim_out = 1.  enc (n, k0); W1 (h, k0); b1 (h); w2 (h) [= Linear(h,1).weight]; b2 (1).
 * y (n); pre2 (n) or NULL receives the output layer's pre-activation for the backward pass. */
int mri_decoder2_forward(const float* enc, int64_t n, int k0, int h, const float* w1, const float* b1,
                         const float* w2, const float* b2, int act1, int act2, float* y, float* pre2,
                         void* stream);

/* Backward of the above: grad_enc (n, k0) is written; grad_w1/b1/w2/b2 are ACCUMULATED. */
int mri_decoder2_backward(const float* enc, int64_t n, int k0, int h, const float* w1, const float* b1,
                          const float* w2, const float* pre2, const float* grad_y, int act1, int act2,
                          float* grad_enc, float* grad_w1, float* grad_b1, float* grad_w2, float* grad_b2,
                          void* stream);

/* Decoder backward FUSED with the hash-grid scatter (autograd of models.py:741-744 in one kernel): as
 * mri_decoder2_backward, but dEnc never goes to memory - it is scattered straight into grad_tables
 * (grad_tables[h_c] += w_c * dEnc, same arena layout as mri_hashgrid_backward).  Covers the headline geometry
 * (F = 2, L = 16, hidden 64, dim 3/4, GELU/ReLU): query with mri_hashdecoder_supported. */
int mri_hashdecoder_supported(int dim, int n_levels, int n_features, int h, int act1);
int mri_hashdecoder_backward(const float* x, int64_t n, int dim, const float* enc, int k0, int h, const float* w1,
                             const float* b1, const float* w2, const float* pre2, const float* grad_y, int act1,
                             int act2, float* grad_tables, const mri_level_t* host_levels, int n_levels,
                             int n_features, float* grad_w1, float* grad_b1, float* grad_w2, float* grad_b2,
                             void* stream);

/* Encoder + decoder forward in ONE kernel for the same geometry (models.py:741-744: decoder(encoder(x))): the
 * gather produces the tensor-core fragments of the first decoder layer in registers.  enc (n, 32) is written when
 * non-NULL (bit-identical to mri_hashgrid_forward; the backward needs it), pre2 (n) likewise; y (n) always. */
int mri_hashdecoder_forward(const float* x, int64_t n, int dim, const float* tables, const mri_level_t* host_levels,
                            int n_levels, int n_features, int k0, int h, const float* w1, const float* b1,
                            const float* w2, const float* b2, int act1, int act2, float* enc, float* y, float* pre2,
                            void* stream);

/* The WHOLE training step of HashMLP under the mean-squared-error loss in one kernel (BaseMLP.training_step
 * models.py:61-66 = F.mse_loss(y, model(x)) on HashMLP.forward :741-744, plus the autograd of both): gather, decoder
 * forward, loss, decoder backward and table scatter per 16-coordinate tile, nothing but the gradients written.
 * target (n) regression targets; inv_count = 1 / (number of elements the mean runs over, normally n);
 * *loss (device float, caller zero-fills) += sum (y - target)^2 * inv_count; y (n) optional predictions or NULL.
 * Gradients accumulate into grad_tables (level layout host_grad_levels, which must equal host_levels' offsets and
 * rows) and grad_w1 / grad_b1 / grad_w2 / grad_b2.  Headline geometry only (mri_hashmlp_mse_step_supported). */
int mri_hashmlp_mse_step_supported(int dim, int n_levels, int n_features, int h, int act1);
int mri_hashmlp_mse_step(const float* x, const float* target, int64_t n, int dim, const float* tables,
                         const mri_level_t* host_levels, int n_levels, int n_features, int k0, int h,
                         const float* w1, const float* b1, const float* w2, const float* b2, int act1, int act2,
                         float inv_count, float* grad_tables, const mri_level_t* host_grad_levels, float* grad_w1,
                         float* grad_b1, float* grad_w2, float* grad_b2, float* loss, float* y, void* stream);

/* ---- wide SIREN layers on tcgen05 tensor cores ------------------------------------------------ */

/* 1 if a layer with `k` inputs and `m` outputs is taken by the tcgen05 path (multiples of 64). */
int mri_siren_tc_supported(int k, int m);

/* fp32 -> two bf16 planes, x = hi + lo (hi = bf16(x), lo = bf16(x - hi)); lo may be NULL. */
int mri_siren_tc_split(const float* src, int64_t count, void* hi, void* lo, void* stream);

/* One dense layer on the tensor cores (SirenLayer.forward, models.py:153-156, and its dgrad):
 *   acc[n, j] = sum_k A[n, k] W[j, k]   A = (a_hi [+ a_lo]) (n, k) bf16, W = (w_hi [+ w_lo]) (m, k) bf16
 *   passes = 3: A_hi W_hi + A_lo W_hi + A_hi W_lo (fp32-parity mode);  passes = 1: bf16 mode
 *   pre = acc + bias;  out = act == SINE ? sin(w0 pre) : pre;  aux = w0 cos(w0 pre) (SINE only)
 *   out *= mul (optional, (n, m) f32) - used by the backward pass: dPre_prev = (dPre W) * act'_prev
 * Outputs (each optional): out_hi/out_lo bf16 planes (next layer's A operand), out_f32, aux_f32. */
int mri_siren_tc_layer(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo,
                       const float* bias, int64_t n, int k, int m, int act, float w0, int passes,
                       const float* mul, void* out_hi, void* out_lo, float* out_f32, float* aux_f32,
                       void* stream);

/* Input gradient of the layer above (autograd of models.py:153-156): dX (n, k) = G (n, m) . W (m, k), with W's
 * planes exactly as the forward uses them (no transpose: W is read as an MN-major tensor-core operand);
 * optional "* mul" (n, k) = the previous layer's activation derivative, so the result is that layer's dPre;
 * colsum (k) or NULL += column sums of the result = that layer's bias gradient (no separate reduction pass). */
int mri_siren_tc_dgrad(const void* g_hi, const void* g_lo, const void* w_hi, const void* w_lo, int64_t n,
                       int k, int m, int passes, const float* mul, void* out_hi, void* out_lo,
                       float* out_f32, float* colsum, void* stream);

/* Weight/bias gradient of the layer above on the tensor cores (autograd of models.py:153-156):
 *   grad_w[j, i] += sum_n G[n, j] X[n, i]   G = dPre planes (n, m), X = layer-input planes (n, k)
 *   grad_b[j]    += sum_n G[n, j]           (grad_b may be NULL)
 * Split-K over the batch with fp32 TMEM accumulation; needs m % 128 == 0 and k % 64 == 0. */
int mri_siren_tc_wgrad(const void* g_hi, const void* g_lo, const void* x_hi, const void* x_lo, int64_t n,
                       int k, int m, int passes, float* grad_w, float* grad_b, void* stream);

/* (hi, lo) bf16 planes of a * b (b may be NULL): dPre = dOut * act' at the head of the backward pass. */
int mri_siren_tc_mul_split(const float* a, const float* b, int64_t count, void* hi, void* lo, void* stream);

/* First layer of a wide SIREN (K = dim_in <= 4): planes of sin(w0 (x W0^T + b0)) written directly in the bf16
 * (hi, lo) format the tensor-core layers read; aux (optional) = w0 cos(.) for the backward pass. */
int mri_siren_first_forward(const float* x, int64_t ldx, const float* w, const float* b, int64_t n, int dim_in,
                            int h, float w0, void* out_hi, void* out_lo, float* aux, void* stream);

/* Backward of the first layer's parameters from dPre0 (n, h) fp32: grad_w0 (h, dim_in) += dPre0^T x,
 * grad_b0 (h) += column sums (grad_b0 may be NULL). */
int mri_siren_first_backward(const float* dpre0, const float* x, int64_t ldx, int64_t n, int dim_in, int h,
                             float* grad_w0, float* grad_b0, void* stream);

/* Output layer (dim_out <= 4, identity activation) straight from the last hidden layer's planes:
 * y (n, m_out) = (hi + lo) W_last^T + b_last. */
int mri_siren_last_forward(const void* hi, const void* lo, const float* w, const float* b, int64_t n, int h,
                           int m_out, float* y, void* stream);

/* Backward of the output layer fused with the head of the hidden backward:
 *   dPre (n, h) = (grad_y W_last) * aux  -> (dpre_hi, dpre_lo) planes;  grad_b_hidden (h) += column sums of dPre
 *   grad_w_last (m_out, h) += grad_y^T (act_hi + act_lo);  grad_b_last (m_out) += column sums of grad_y
 * grad_b_hidden / grad_b_last may be NULL. */
int mri_siren_last_backward(const float* grad_y, const float* w, const float* aux, const void* act_hi,
                            const void* act_lo, int64_t n, int h, int m_out, void* dpre_hi, void* dpre_lo,
                            float* grad_b_hidden, float* grad_w_last, float* grad_b_last, void* stream);

/* ---- loss / optimiser ------------------------------------------------------------------ */

/* F.mse_loss(y, y_pred) (models.py:64): *loss += sum((pred-target)^2) * inv_count and
 * grad_pred = 2 (pred-target) * inv_count; caller zeroes *loss; grad_pred may be NULL.
 * inv_count = 1/numel of the GLOBAL batch (data-parallel ranks pass the global count). */
int mri_mse_loss_grad(const float* pred, const float* target, int64_t count, float inv_count,
                      float* grad_pred, float* loss, void* stream);

/* One dense torch.optim.Adam step (models.py:68-70) over a flat parameter arena.
 * step is 1-based; hyper-parameters are doubles like torch's python floats (bias corrections
 * are evaluated in double on the host); weight_decay is torch's coupled L2 term.  grad_scale multiplies g first
 * (1.0 normally).  If zero_grad != 0 the gradient is cleared in the same pass. */
int mri_adam_step(float* p, float* g, float* m, float* v, int64_t count, int64_t step, double lr,
                  double beta1, double beta2, double eps, double weight_decay, double grad_scale,
                  int zero_grad, void* stream);

/* The same step for CUDA-graph capture: the 1-based step counter lives in device memory (*step_dev, int64) and is
 * advanced by the call itself (a one-thread kernel that also evaluates the bias corrections in double into
 * hyper_dev[0..1]), so a captured training step can be replayed without host-side arguments going stale. */
int mri_adam_step_captured(float* p, float* g, float* m, float* v, int64_t count, int64_t* step_dev, float* hyper_dev,
                           double lr, double beta1, double beta2, double eps, double weight_decay, double grad_scale,
                           int zero_grad, void* stream);

/* Data-parallel variant of the step above, ONE kernel per rank over NVLink peer memory: for its shard
 * [shard_begin, shard_begin + shard_len) of the flat arena the rank sums the gradients of all `world` ranks by
 * loading their gradient arenas directly (host_peer_grads[r] = device pointer of rank r's arena, P2P-mapped, e.g.
 * torch symmetric memory), applies Adam with its local moment shards and stores the new parameters into every
 * rank's parameter arena (host_peer_params[r]).  = reduce-scatter + Adam + all-gather fused; the caller brackets it
 * with cross-rank barriers (gradients complete before, parameters visible after).  grad_scale = 1/world for means.
 * grad_multicast / param_multicast: NVSwitch multicast (NVLS) mappings of the two arenas, or 0; when given, the sum
 * is ONE multimem.ld_reduce (in-switch reduction) and the broadcast ONE multimem.st per 16 bytes.
 * zero_grad != 0: the shard owner also clears its slice of every rank's gradient arena (no separate memset). */
int mri_adam_step_sharded(const uint64_t* host_peer_grads, const uint64_t* host_peer_params,
                          uint64_t grad_multicast, uint64_t param_multicast, int world, int rank, float* m_shard, float* v_shard, int64_t shard_begin, int64_t shard_len, int64_t step,
                          double lr, double beta1, double beta2, double eps, double weight_decay,
                          double grad_scale, int zero_grad, void* stream);

/* Same step with the two cross-rank barriers folded into the kernel: host_peer_flags[r] = device pointer of rank r's
 * 32-int flag buffer (P2P-mapped symmetric memory, zero-initialised once).  The kernel waits until every rank's
 * gradients are complete before touching them and does not complete on the stream before every rank has stored its
 * shard everywhere - no separate barrier launches.  `step` must be the same on all ranks and grow by one per call. */
int mri_adam_step_sharded_sync(const uint64_t* host_peer_grads, const uint64_t* host_peer_params,
                               uint64_t grad_multicast, uint64_t param_multicast, const uint64_t* host_peer_flags,
                               int world, int rank, float* m_shard, float* v_shard, int64_t shard_begin,
                               int64_t shard_len, int64_t step, double lr, double beta1, double beta2, double eps,
                               double weight_decay, double grad_scale, int zero_grad, void* stream);

/* ---- dense-grid sweep -------------------------------------------------------------------- */

/* Coordinates of voxels [first, first+count) of a C-order grid of `shape` (launcher.py:191-202,
 * utils.py:14-23): coords[i, d] = axis_d[idx_d].  `axes` = concatenated per-axis value vectors
 * (the caller passes torch.linspace output so the floats are bit-identical to the reference). */
int mri_grid_coords(const float* axes, const int32_t* host_shape, int dim, int64_t first,
                    int64_t count, float* coords, void* stream);

/* Training batch from voxel indices (MriImage semantics, datamodules.py:130-172): coords[i,:] are the
 * grid coordinates of flat C-order voxel index[i] (synthesised, no coordinate table), values[i] =
 * pixels[index[i]].  index is int64 on the device; pixels/values may both be NULL. */
int mri_gather_voxels(const float* axes, const int32_t* host_shape, int dim, const int64_t* index,
                      int64_t count, const float* pixels, float* coords, float* values, void* stream);

/* Fused sweep of a hash-grid + MLP-decoder model (models.py:741-751 intended forward, nb cell 37):
 * voxel index -> coordinates -> all-level encoding -> decoder, nothing but `out` (count x m_out)
 * is written.  Decoder = n_dense layers of Linear+act, parameters packed as
 * [W0 (m0,k0) | b0 (m0) | W1 | b1 ...] in `decoder`; host_dims = [k0, m0, m1, ...]. */
int mri_hashmlp_sweep(const float* axes, const int32_t* host_shape, int dim, int64_t first,
                      int64_t count, const float* tables, const mri_level_t* host_levels,
                      int n_levels, int n_features, const float* decoder, const int32_t* host_dims,
                      int n_dense, int act, int last_act, float* out, void* stream);

/* ---- image-quality metrics and the linear-in-time baseline ------------------------ */

/* *sum_out += sum_i (a[i] - b[i])^2 in double (device pointer, caller zero-fills): skimage mean_squared_error /
 * peak_signal_noise_ratio of legacy_code/hash_experimentation.py:445-453 are sum/n and 10 log10(range^2 n / sum). */
int mri_sq_err_sum(const float* a, const float* b, int64_t n, double* sum_out, void* stream);

/* *sum_out += sum of the per-pixel SSIM index (skimage structural_similarity defaults: win x win uniform window,
 * K1 = 0.01, K2 = 0.03, sample covariance) over the interior (borders of (win-1)/2 cropped) of every (axis 0, axis 1)
 * plane of a, b shaped (nx, ny, planes) C-order; mean SSIM = sum / ((nx-win+1)(ny-win+1) planes). */
int mri_ssim_sum(const float* a, const float* b, int nx, int ny, int64_t planes, int win, double data_range,
                 double* sum_out, void* stream);

/* interp.py:35-52 baseline: data, out (outer, t) C-order; out[o, f] = linear interpolation of the kept frames
 * data[o, ::2] at continuous index min(f / 2, ceil(t / 2) - 1). */
int mri_linear_time_interp(const float* data, int64_t outer, int t, float* out, void* stream);

/* ---- measurement probe (bench.py only) ------------------------------------------------------------ */

/* Peak rate of 32-byte sector reductions at the L2: red.global.add.v2.f32 from every lane into pseudo-random sectors of
 * `table` (table_floats floats, L2-resident sizes), `lanes_per_sector` = 1 (32 sector operations per warp instruction)
 * or 2 (adjacent lanes share a sector, as the pair-lane scatter does).  *sector_ops (host) = operations one launch
 * issues; the caller times the launch.  The roofline the fused backward kernel is reported against besides HBM. */
int mri_probe_red_rate(float* table, int64_t table_floats, int iters, int lanes_per_sector, int64_t* sector_ops,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MRI_B200_H */
