// MSE loss/gradient (K5) and the dense Adam step over a flat parameter arena (K6), sm_100a.
//
// Both are pure HBM streams: float4 loads/stores, grid sized as a multiple of the SM count,
// grid-stride loops.  Adam moves 28 B/param (read p,g,m,v; write p,m,v) - 32 B when it also
// clears the gradient in the same pass (saves the separate 4 B/param memset + one launch).
#include <math.h>

#include <stdlib.h>

#include "common.cuh"

namespace mri {
namespace {

struct AdamArgs {
  float step_size;      // lr / (1 - beta1^t)
  float bc2_sqrt;       // sqrt(1 - beta2^t)
  float beta1, beta2, one_minus_beta1, one_minus_beta2;
  float eps, weight_decay, grad_scale;
  int zero_grad;
};

// torch.optim.Adam single-tensor path: exp_avg.lerp_(g, 1-b1); exp_avg_sq.mul_(b2).addcmul_(g,g,1-b2);
// denom = sqrt(v)/sqrt(bc2) + eps; p.addcdiv_(m, denom, value=-lr/bc1)
__device__ __forceinline__ void adam_update(float& p, float& g, float& m, float& v, const AdamArgs& a) {
  float gg = g * a.grad_scale;
  if (a.weight_decay != 0.0f) gg = fmaf(a.weight_decay, p, gg);
  m = fmaf(a.one_minus_beta1, gg - m, m);
  v = fmaf(a.one_minus_beta2 * gg, gg, v * a.beta2);
  const float denom = __fdiv_rn(sqrtf(v), a.bc2_sqrt) + a.eps;  // torch: (sqrt(v) / sqrt(bc2)).add_(eps)
  p = p - a.step_size * __fdiv_rn(m, denom);                    // torch: p.addcdiv_(m, denom, value=-lr/bc1)
  if (a.zero_grad) g = 0.0f;
}

// CUDA-graph friendly step counter: lives in device memory and is advanced by the launch itself, so a captured
// step can be replayed; the bias corrections are evaluated in double like the host path.
__global__ void adam_hyper_kernel(int64_t* __restrict__ step, float* __restrict__ hyper, double lr, double beta1, double beta2) {
  const int64_t t = *step + 1;
  *step = t;
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(t));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(t));
  hyper[0] = static_cast<float>(lr / bc1);
  hyper[1] = static_cast<float>(sqrt(bc2));
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n4, int64_t n, AdamArgs a,
                                                   const float* __restrict__ hyper) {
  if (hyper != nullptr) {  // {lr / bc1, sqrt(bc2)} written by adam_hyper_kernel earlier on this stream
    a.step_size = __ldg(hyper);
    a.bc2_sqrt = __ldg(hyper + 1);
  }
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_update(pp.x, gg.x, mm.x, vv.x, a);
    adam_update(pp.y, gg.y, mm.y, vv.y, a);
    adam_update(pp.z, gg.z, mm.z, vv.z, a);
    adam_update(pp.w, gg.w, mm.w, vv.w, a);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (a.zero_grad) reinterpret_cast<float4*>(g)[i] = gg;
  }
  // scalar tail (count not a multiple of 4)
  const int64_t tail = n4 * 4 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tail < n) adam_update(p[tail], g[tail], m[tail], v[tail], a);
}

__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                  int64_t n, float inv_count, float* __restrict__ grad,
                                                  float* __restrict__ loss) {
  float acc = 0.0f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = pred[i] - target[i];
    acc = fmaf(d, d, acc);
    if (grad) grad[i] = 2.0f * d * inv_count;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float warp_sum[8];
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = warp_sum[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, s * inv_count);
  }
}

// ---- fused reduce-scatter + Adam + all-gather over NVLink peer memory -------------------------------------
// Every rank owns one contiguous shard of the flat arena.  For its shard it (1) sums the gradient over ALL ranks by
// loading the peers' gradient arenas directly (P2P loads through NVLink/NVSwitch, symmetric-memory pointers),
// (2) applies the Adam update with its local moments, (3) stores the new parameters into EVERY rank's parameter
// arena (P2P stores).  One kernel replaces ncclAllReduce(61 MB) + a full-arena Adam pass: each GPU moves
// (W-1)/W of the arena in and out instead of 2(W-1)/W through NCCL's staging, and the optimiser work shrinks by W.
constexpr int MAX_PEERS = 8;
struct PeerPtrs {
  const float* grad[MAX_PEERS];
  float* param[MAX_PEERS];
  const float* grad_mc;  // NVSwitch multicast mapping of the gradient arenas (0 = not available)
  float* param_mc;       // NVSwitch multicast mapping of the parameter arenas
  int* flags[MAX_PEERS]; // in-kernel cross-rank synchronisation (SYNC = true): per rank 32 ints of symmetric memory,
                         // [0..7] "gradients of rank r are complete" and [8..15] "rank r has stored its shard everywhere"
                         // (both hold the step number), [16] count of finished CTAs of the local launch
  int step;
};

__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// spin until *p >= want; a rank that never arrives ends the process (trap) after ~60 s instead of hanging the GPU
__device__ __forceinline__ void wait_flag(const int* p, int want) {
  const long long t0 = clock64();
  while (ld_acquire_sys(p) < want) {
    __nanosleep(64);
    if (clock64() - t0 > 120000000000LL) __trap();
  }
}

// in-switch (NVLS) reduction: one load returns the sum over every rank's copy of the address
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc_addr) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc_addr) : "memory");
  return r;
}
// multicast store: the switch writes the value into every rank's copy
__device__ __forceinline__ void multimem_st(float* mc_addr, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// U float4 per thread and iteration: all U gradient reductions (and the parameter / moment loads) are issued before the
// first one is consumed, so a thread keeps U NVLink round trips in flight instead of one.
//
// SYNC = true folds the two cross-rank barriers into the kernel: CTA 0 publishes "my gradient arena is complete" (the
// kernel runs after this rank's backward on the stream) to every peer, every CTA waits until all peers have published
// theirs before it touches their arenas; at the end the last CTA to finish publishes "my shard is stored everywhere"
// and does not retire before all peers have published the same - so when the kernel completes on the stream, every
// replica holds all new parameters and every slice of the local gradient arena has been consumed by its owner.
template <int U, bool SYNC>
__global__ void __launch_bounds__(256) adam_sharded_kernel(const PeerPtrs peers, int world, int rank, float* __restrict__ m,
                                                           float* __restrict__ v, int64_t shard_begin4, int64_t shard_len4,
                                                           AdamArgs a) {
  if constexpr (SYNC) {
    if (blockIdx.x == 0 && threadIdx.x < world) st_release_sys(peers.flags[threadIdx.x] + rank, peers.step);
    if (threadIdx.x < world) wait_flag(peers.flags[rank] + threadIdx.x, peers.step);
    __syncthreads();
  }
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  AdamArgs b = a;
  b.zero_grad = 0;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < shard_len4; i0 += stride * U) {
    float4 gg[U], pp[U], mm[U], vv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      gg[u] = zero;
      if (i < shard_len4) {
        const int64_t gi = shard_begin4 + i;  // float4 index inside the arena
        if (peers.grad_mc) {
          gg[u] = multimem_ld_reduce_add(peers.grad_mc + 4 * gi);
        } else {
#pragma unroll
          for (int r = 0; r < MAX_PEERS; ++r) {
            if (r < world) {
              const float4 t = reinterpret_cast<const float4*>(peers.grad[r])[gi];
              gg[u].x += t.x; gg[u].y += t.y; gg[u].z += t.z; gg[u].w += t.w;
            }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < shard_len4) {
        pp[u] = reinterpret_cast<const float4*>(peers.param[rank])[shard_begin4 + i];
        mm[u] = reinterpret_cast<float4*>(m)[i];
        vv[u] = reinterpret_cast<float4*>(v)[i];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= shard_len4) continue;
      const int64_t gi = shard_begin4 + i;
      adam_update(pp[u].x, gg[u].x, mm[u].x, vv[u].x, b);
      adam_update(pp[u].y, gg[u].y, mm[u].y, vv[u].y, b);
      adam_update(pp[u].z, gg[u].z, mm[u].z, vv[u].z, b);
      adam_update(pp[u].w, gg[u].w, mm[u].w, vv[u].w, b);
      reinterpret_cast<float4*>(m)[i] = mm[u];
      reinterpret_cast<float4*>(v)[i] = vv[u];
      if (peers.param_mc) {
        multimem_st(peers.param_mc + 4 * gi, pp[u]);
        if (a.zero_grad) multimem_st(const_cast<float*>(peers.grad_mc) + 4 * gi, zero);  // clears this slice on every rank
      } else {
#pragma unroll
        for (int r = 0; r < MAX_PEERS; ++r)
          if (r < world) {
            reinterpret_cast<float4*>(peers.param[r])[gi] = pp[u];
            if (a.zero_grad) reinterpret_cast<float4*>(const_cast<float*>(peers.grad[r]))[gi] = zero;
          }
      }
    }
  }
  if constexpr (SYNC) {
    // one system fence per CTA, after the CTA barrier (the cooperative-groups grid-sync pattern: the barrier orders the
    // other threads' stores before thread 0's fence).  A fence in every thread made the kernel 3.6x slower (0.42 vs
    // 0.12 ms at W = 2): MEMBAR.SYS is serviced one at a time per SM.
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      int* local = peers.flags[rank];
      const int ticket = atomicAdd(local + 16, 1);
      if (ticket == static_cast<int>(gridDim.x) - 1) {
        local[16] = 0;  // every CTA of this launch has been counted: ready for the next launch
        __threadfence_system();
        for (int r = 0; r < world; ++r) st_release_sys(peers.flags[r] + 8 + rank, peers.step);
        for (int r = 0; r < world; ++r) wait_flag(local + 8 + r, peers.step);
      }
    }
  }
}

}  // namespace
}  // namespace mri

using namespace mri;

namespace {
int sharded_step(const uint64_t* host_peer_grads, const uint64_t* host_peer_params, uint64_t grad_multicast, uint64_t param_multicast,
                 const uint64_t* host_peer_flags, int world, int rank, float* m_shard, float* v_shard, int64_t shard_begin,
                 int64_t shard_len, int64_t step, double lr, double beta1, double beta2, double eps, double weight_decay,
                 double grad_scale, int zero_grad, void* stream);
}

extern "C" int mri_adam_step_sharded(const uint64_t* host_peer_grads, const uint64_t* host_peer_params, uint64_t grad_multicast,
                                     uint64_t param_multicast, int world, int rank, float* m_shard, float* v_shard, int64_t shard_begin, int64_t shard_len, int64_t step,
                                     double lr, double beta1, double beta2, double eps, double weight_decay,
                                     double grad_scale, int zero_grad, void* stream) {
  return sharded_step(host_peer_grads, host_peer_params, grad_multicast, param_multicast, nullptr, world, rank, m_shard, v_shard,
                      shard_begin, shard_len, step, lr, beta1, beta2, eps, weight_decay, grad_scale, zero_grad, stream);
}

extern "C" int mri_adam_step_sharded_sync(const uint64_t* host_peer_grads, const uint64_t* host_peer_params, uint64_t grad_multicast,
                                          uint64_t param_multicast, const uint64_t* host_peer_flags, int world, int rank,
                                          float* m_shard, float* v_shard, int64_t shard_begin, int64_t shard_len, int64_t step,
                                          double lr, double beta1, double beta2, double eps, double weight_decay,
                                          double grad_scale, int zero_grad, void* stream) {
  if (!host_peer_flags) return fail(MRI_ERR_INVALID, "adam_sharded_sync: null flag pointers");
  if (step >= (int64_t{1} << 31)) return fail(MRI_ERR_INVALID, "adam_sharded_sync: step counter exceeds the 32-bit flag range");
  return sharded_step(host_peer_grads, host_peer_params, grad_multicast, param_multicast, host_peer_flags, world, rank, m_shard,
                      v_shard, shard_begin, shard_len, step, lr, beta1, beta2, eps, weight_decay, grad_scale, zero_grad, stream);
}

namespace {
int sharded_step(const uint64_t* host_peer_grads, const uint64_t* host_peer_params, uint64_t grad_multicast, uint64_t param_multicast,
                 const uint64_t* host_peer_flags, int world, int rank, float* m_shard, float* v_shard, int64_t shard_begin,
                 int64_t shard_len, int64_t step, double lr, double beta1, double beta2, double eps, double weight_decay,
                 double grad_scale, int zero_grad, void* stream) {
  if (!host_peer_grads || !host_peer_params || !m_shard || !v_shard) return fail(MRI_ERR_INVALID, "adam_sharded: null pointer");
  if (world < 1 || world > MAX_PEERS || rank < 0 || rank >= world) return fail(MRI_ERR_UNSUPPORTED, "adam_sharded: world=%d rank=%d", world, rank);
  if (shard_begin < 0 || shard_len < 0 || (shard_begin & 3) || (shard_len & 3) || step < 1)
    return fail(MRI_ERR_INVALID, "adam_sharded: shard must be a multiple of 4 floats and step >= 1");
  if (shard_len == 0 && !host_peer_flags) return MRI_OK;
  if (shard_len == 0) return fail(MRI_ERR_INVALID, "adam_sharded_sync: empty shard (every rank must take part in the flags)");
  PeerPtrs peers{};
  peers.step = static_cast<int>(step);
  for (int r = 0; r < world; ++r) {
    if (host_peer_flags) {
      if (host_peer_flags[r] & 3) return fail(MRI_ERR_INVALID, "adam_sharded_sync: flag buffers must be 4-byte aligned");
      peers.flags[r] = reinterpret_cast<int*>(host_peer_flags[r]);
    }
    if ((host_peer_grads[r] | host_peer_params[r]) & 15) return fail(MRI_ERR_INVALID, "adam_sharded: peer arenas must be 16-byte aligned");
    peers.grad[r] = reinterpret_cast<const float*>(host_peer_grads[r]);
    peers.param[r] = reinterpret_cast<float*>(host_peer_params[r]);
  }
  if ((grad_multicast | param_multicast) & 15) return fail(MRI_ERR_INVALID, "adam_sharded: multicast pointers must be 16-byte aligned");
  peers.grad_mc = reinterpret_cast<const float*>(grad_multicast);
  peers.param_mc = reinterpret_cast<float*>(param_multicast);
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  AdamArgs a;
  a.step_size = static_cast<float>(lr / bc1);
  a.bc2_sqrt = static_cast<float>(sqrt(bc2));
  a.beta1 = static_cast<float>(beta1);
  a.beta2 = static_cast<float>(beta2);
  a.one_minus_beta1 = static_cast<float>(1.0 - beta1);
  a.one_minus_beta2 = static_cast<float>(1.0 - beta2);
  a.eps = static_cast<float>(eps);
  a.weight_decay = static_cast<float>(weight_decay);
  a.grad_scale = static_cast<float>(grad_scale);
  a.zero_grad = zero_grad;  // the owner of a slice clears it in EVERY rank's gradient arena after reducing it
  const int64_t n4 = shard_len / 4;
  // MRI_DP_UNROLL = 1 | 2 | 4 float4 per thread and iteration (default 2), MRI_DP_BLOCKS_PER_SM caps the grid (default 8)
  static const int unroll = [] { const char* e = getenv("MRI_DP_UNROLL"); const int u = e ? atoi(e) : 2; return u >= 4 ? 4 : u <= 1 ? 1 : 2; }();
  static const int per_sm = [] { const char* e = getenv("MRI_DP_BLOCKS_PER_SM"); const int u = e ? atoi(e) : 8; return u < 1 ? 1 : u > 16 ? 16 : u; }();
  int64_t want = (n4 + 256LL * unroll - 1) / (256LL * unroll);
  const int64_t cap = static_cast<int64_t>(per_sm) * sm_count();
  if (want > cap) want = cap;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (host_peer_flags) {
    // every CTA spins on the peers' flags at its start: the grid must be resident at once (<= 8 CTAs of 256 threads per SM)
    const int64_t resident = 6LL * sm_count();
    if (want > resident) want = resident;
    if (unroll == 4) adam_sharded_kernel<4, true><<<static_cast<int>(want), 256, 0, s>>>(peers, world, rank, m_shard, v_shard, shard_begin / 4, n4, a);
    else if (unroll == 2) adam_sharded_kernel<2, true><<<static_cast<int>(want), 256, 0, s>>>(peers, world, rank, m_shard, v_shard, shard_begin / 4, n4, a);
    else adam_sharded_kernel<1, true><<<static_cast<int>(want), 256, 0, s>>>(peers, world, rank, m_shard, v_shard, shard_begin / 4, n4, a);
  } else if (unroll == 4) adam_sharded_kernel<4, false><<<static_cast<int>(want), 256, 0, s>>>(peers, world, rank, m_shard, v_shard, shard_begin / 4, n4, a);
  else if (unroll == 2) adam_sharded_kernel<2, false><<<static_cast<int>(want), 256, 0, s>>>(peers, world, rank, m_shard, v_shard, shard_begin / 4, n4, a);
  else adam_sharded_kernel<1, false><<<static_cast<int>(want), 256, 0, s>>>(peers, world, rank, m_shard, v_shard, shard_begin / 4, n4, a);
  MRI_LAUNCH_OK("adam_sharded_kernel");
  return MRI_OK;
}
}  // namespace

extern "C" int mri_mse_loss_grad(const float* pred, const float* target, int64_t count, float inv_count,
                                 float* grad_pred, float* loss, void* stream) {
  if (!pred || !target) return fail(MRI_ERR_INVALID, "mse: null pointer");
  if (count < 0) return fail(MRI_ERR_INVALID, "mse: negative count");
  if (count == 0) return MRI_OK;
  const int64_t want = (count + 255) / 256;
  const int grid = static_cast<int>(want < 4LL * sm_count() ? want : 4LL * sm_count());
  mse_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(pred, target, count, inv_count, grad_pred, loss);
  MRI_LAUNCH_OK("mse_kernel");
  return MRI_OK;
}

namespace {
int launch_adam(float* p, float* g, float* m, float* v, int64_t count, double step_size, double bc2_sqrt, double beta1,
                double beta2, double eps, double weight_decay, double grad_scale, int zero_grad, const float* hyper,
                cudaStream_t stream) {
  AdamArgs a;
  a.step_size = static_cast<float>(step_size);
  a.bc2_sqrt = static_cast<float>(bc2_sqrt);
  a.beta1 = static_cast<float>(beta1);
  a.beta2 = static_cast<float>(beta2);
  a.one_minus_beta1 = static_cast<float>(1.0 - beta1);
  a.one_minus_beta2 = static_cast<float>(1.0 - beta2);
  a.eps = static_cast<float>(eps);
  a.weight_decay = static_cast<float>(weight_decay);
  a.grad_scale = static_cast<float>(grad_scale);
  a.zero_grad = zero_grad;
  const int64_t n4 = count / 4;
  int64_t want = (n4 + 255) / 256;
  if (want < 1) want = 1;
  const int64_t cap = 8LL * sm_count();
  const int grid = static_cast<int>(want < cap ? want : cap);
  adam_kernel<<<grid, 256, 0, stream>>>(p, g, m, v, n4, count, a, hyper);
  MRI_LAUNCH_OK("adam_kernel");
  return MRI_OK;
}

int check_adam_arenas(const float* p, const float* g, const float* m, const float* v, int64_t count) {
  if (!p || !g || !m || !v) return fail(MRI_ERR_INVALID, "adam: null pointer");
  if (count < 0) return fail(MRI_ERR_INVALID, "adam: count must be >= 0");
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
       reinterpret_cast<uintptr_t>(v)) & 15)
    return fail(MRI_ERR_INVALID, "adam: arenas must be 16-byte aligned");
  return MRI_OK;
}
}  // namespace

extern "C" int mri_adam_step(float* p, float* g, float* m, float* v, int64_t count, int64_t step, double lr,
                             double beta1, double beta2, double eps, double weight_decay, double grad_scale,
                             int zero_grad, void* stream) {
  int st = check_adam_arenas(p, g, m, v, count);
  if (st != MRI_OK) return st;
  if (step < 1) return fail(MRI_ERR_INVALID, "adam: step must be >= 1");
  if (count == 0) return MRI_OK;
  // bias corrections in double, like torch's python-float arithmetic
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  return launch_adam(p, g, m, v, count, lr / bc1, sqrt(bc2), beta1, beta2, eps, weight_decay, grad_scale, zero_grad, nullptr,
                     static_cast<cudaStream_t>(stream));
}

extern "C" int mri_adam_step_captured(float* p, float* g, float* m, float* v, int64_t count, int64_t* step_dev,
                                      float* hyper_dev, double lr, double beta1, double beta2, double eps,
                                      double weight_decay, double grad_scale, int zero_grad, void* stream) {
  int st = check_adam_arenas(p, g, m, v, count);
  if (st != MRI_OK) return st;
  if (!step_dev || !hyper_dev) return fail(MRI_ERR_INVALID, "adam_captured: null step/hyper buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  adam_hyper_kernel<<<1, 1, 0, s>>>(step_dev, hyper_dev, lr, beta1, beta2);  // advances the counter even for count == 0
  MRI_LAUNCH_OK("adam_hyper_kernel");
  if (count == 0) return MRI_OK;
  return launch_adam(p, g, m, v, count, 0.0, 1.0, beta1, beta2, eps, weight_decay, grad_scale, zero_grad, hyper_dev, s);
}
