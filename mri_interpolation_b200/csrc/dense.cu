// fp32 CUDA-core dense layers (exact-fp32 path for narrow layers: HashMLP decoder, first/last
// SIREN layers, and any layer shape the tcgen05 path does not take).
//
// One strided, tiled SGEMM  C(i,j) (+)= sum_k A(i,k) B(k,j)  serves forward, dgrad and wgrad:
//   forward : y   = act(x W^T + b)        (optionally storing the pre-activation)
//   dgrad   : dx  = dpre W
//   wgrad   : dW += dpre^T x              (split-K over the batch, atomic accumulate)
// The small output dimension is always mapped to the GEMM's M role (BM=16 tile) so that
// out-features = 1 or dim_in = 3/4 do not waste a 64-wide tile.
#include "common.cuh"

namespace mri {
namespace {

struct GemmArgs {
  const float* A; int64_t sa0, sa1;   // A(i,k) = A[i*sa0 + k*sa1],  i < M, k < K
  const float* B; int64_t sb0, sb1;   // B(k,j) = B[k*sb0 + j*sb1],  j < N
  float* C; int64_t sc0, sc1;         // C(i,j)
  float* P;                           // optional pre-activation output, same strides as C
  const float* bias; int bias_mode;   // 0 none, 1 bias[i], 2 bias[j]
  int64_t M, N, K;
  int64_t k_chunk;                    // K elements per blockIdx.y (split-K)
  unsigned tiles_n;                   // tiles along N; blockIdx.x = tile_m * tiles_n + tile_n
  int act; float w0;
  int accumulate;                     // 1: atomic += into C (split-K), epilogue skipped
};

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) sgemm_kernel(const GemmArgs g) {
  constexpr int NT = (BM / TM) * (BN / TN);
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN);  // along N
  const int ty = tid / (BN / TN);  // along M
  // N tiles on blockIdx.x (fastest) so neighbouring blocks share the A rows
  const int64_t i0 = static_cast<int64_t>(blockIdx.x / g.tiles_n) * BM;
  const int64_t j0 = static_cast<int64_t>(blockIdx.x % g.tiles_n) * BN;
  const int64_t k_begin = static_cast<int64_t>(blockIdx.y) * g.k_chunk;
  const int64_t k_end = (k_begin + g.k_chunk < g.K) ? (k_begin + g.k_chunk) : g.K;

  float acc[TM][TN];
#pragma unroll
  for (int a = 0; a < TM; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) acc[a][b] = 0.0f;

  const bool a_k_contig = (g.sa1 == 1);
  const bool b_n_contig = (g.sb1 == 1);

  for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll 2
    for (int e = tid; e < BM * BK; e += NT) {
      int mm, kk;
      if (a_k_contig) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      const int64_t gi = i0 + mm, gk = k0 + kk;
      As[kk][mm] = (gi < g.M && gk < k_end) ? __ldg(g.A + gi * g.sa0 + gk * g.sa1) : 0.0f;
    }
#pragma unroll 2
    for (int e = tid; e < BN * BK; e += NT) {
      int nn, kk;
      if (b_n_contig) { nn = e % BN; kk = e / BN; } else { kk = e % BK; nn = e / BK; }
      const int64_t gj = j0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gj < g.N && gk < k_end) ? __ldg(g.B + gk * g.sb0 + gj * g.sb1) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int t = 0; t < TM; ++t) a[t] = As[kk][ty * TM + t];
#pragma unroll
      for (int t = 0; t < TN; ++t) b[t] = Bs[kk][tx + t * (BN / TN)];  // strided columns: conflict-free reads
#pragma unroll
      for (int s = 0; s < TM; ++s)
#pragma unroll
        for (int t = 0; t < TN; ++t) acc[s][t] = fmaf(a[s], b[t], acc[s][t]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int s = 0; s < TM; ++s) {
    const int64_t gi = i0 + ty * TM + s;
    if (gi >= g.M) continue;
#pragma unroll
    for (int t = 0; t < TN; ++t) {
      const int64_t gj = j0 + tx + t * (BN / TN);
      if (gj >= g.N) continue;
      const int64_t off = gi * g.sc0 + gj * g.sc1;
      if (g.accumulate) {
        red_add_f32(g.C + off, acc[s][t]);
      } else {
        float pre = acc[s][t];
        if (g.bias_mode == 1) pre += __ldg(g.bias + gi);
        else if (g.bias_mode == 2) pre += __ldg(g.bias + gj);
        if (g.P) g.P[off] = pre;
        g.C[off] = activate_rt(g.act, pre, g.w0);
      }
    }
  }
}

int launch_gemm(GemmArgs g, bool split_k, cudaStream_t s) {
  // the small dimension sits on M: pick the narrow tile when M <= 16
  const bool narrow = g.M <= 16;
  const int BM = narrow ? 16 : 64, BN = narrow ? 128 : 64, BK = 16;
  const int64_t tiles_m = (g.M + BM - 1) / BM, tiles_n = (g.N + BN - 1) / BN;
  int64_t splits = 1;
  if (split_k) {
    const int64_t target = 4LL * sm_count();
    splits = (target + tiles_m * tiles_n - 1) / (tiles_m * tiles_n);
    const int64_t max_splits = (g.K + 8 * BK - 1) / (8 * BK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
  }
  int64_t chunk = (g.K + splits - 1) / splits;
  chunk = (chunk + BK - 1) / BK * BK;
  splits = (g.K + chunk - 1) / chunk;
  g.k_chunk = chunk;
  if (tiles_m * tiles_n > 0x7fffffffLL)
    return fail(MRI_ERR_UNSUPPORTED, "dense: problem too large for one launch (%lld tiles)", (long long)(tiles_m * tiles_n));
  g.tiles_n = static_cast<unsigned>(tiles_n);
  dim3 grid(static_cast<unsigned>(tiles_m * tiles_n), static_cast<unsigned>(splits), 1);
  if (narrow) sgemm_kernel<16, 128, 16, 2, 4><<<grid, 256, 0, s>>>(g);
  else sgemm_kernel<64, 64, 16, 4, 4><<<grid, 256, 0, s>>>(g);
  MRI_LAUNCH_OK("sgemm_kernel");
  return MRI_OK;
}

// dpre = grad_y * act'(pre), plus per-column sums for the bias gradient.
// Path A (256 % m == 0): a thread always sees the same column -> register accumulation.
__global__ void __launch_bounds__(256) dpre_fixedcol_kernel(const float* __restrict__ pre, const float* __restrict__ gy,
                                                            int64_t total, int m, int act, float w0,
                                                            float* __restrict__ dpre, float* __restrict__ grad_b) {
  __shared__ float part[256];
  float acc = 0.0f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * 256;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; e < total; e += stride) {
    const float d = gy[e] * activate_grad_rt(act, pre[e], w0);
    dpre[e] = d;
    acc += d;
  }
  if (!grad_b) return;
  part[threadIdx.x] = acc;
  __syncthreads();
  // column of thread t is t % m (stride and block offset are multiples of m)
  if (threadIdx.x < m) {
    float s = 0.0f;
    for (int t = threadIdx.x; t < 256; t += m) s += part[t];
    red_add_f32(grad_b + threadIdx.x, s);
  }
}

// Path B (any m): block owns a slab of rows, threads stride over columns (coalesced).
__global__ void __launch_bounds__(256) dpre_rowslab_kernel(const float* __restrict__ pre, const float* __restrict__ gy,
                                                           int64_t n, int m, int64_t rows_per_block, int act, float w0,
                                                           float* __restrict__ dpre, float* __restrict__ grad_b) {
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = (r0 + rows_per_block < n) ? r0 + rows_per_block : n;
  for (int j = threadIdx.x; j < m; j += 256) {
    float acc = 0.0f;
    for (int64_t r = r0; r < r1; ++r) {
      const int64_t e = r * m + j;
      const float d = gy[e] * activate_grad_rt(act, pre[e], w0);
      dpre[e] = d;
      acc += d;
    }
    if (grad_b) red_add_f32(grad_b + j, acc);
  }
}

}  // namespace
}  // namespace mri

using namespace mri;

extern "C" int mri_dense_forward(const float* x, int64_t ldx, const float* w, const float* b, int64_t n, int k, int m,
                                 int act, float w0, float* y, float* pre, void* stream) {
  if (n == 0 && w) return MRI_OK;
  if (!x || !w || !y) return fail(MRI_ERR_INVALID, "dense_forward: null pointer");
  if (n < 0 || k < 1 || m < 1 || ldx < k) return fail(MRI_ERR_INVALID, "dense_forward: bad sizes n=%lld k=%d m=%d ldx=%lld",
                                                      (long long)n, k, m, (long long)ldx);
  if (act < MRI_ACT_IDENTITY || act > MRI_ACT_RELU) return fail(MRI_ERR_INVALID, "dense_forward: unknown activation %d", act);
  if (n == 0) return MRI_OK;
  GemmArgs g{};
  g.K = k; g.act = act; g.w0 = w0; g.bias = b; g.accumulate = 0; g.P = pre; g.C = y;
  if (m <= 16) {
    // C^T (m x n) = W (m x k) * x^T (k x n)
    g.A = w; g.sa0 = k; g.sa1 = 1;
    g.B = x; g.sb0 = 1; g.sb1 = ldx;
    g.sc0 = 1; g.sc1 = m;
    g.M = m; g.N = n; g.bias_mode = b ? 1 : 0;
  } else {
    g.A = x; g.sa0 = ldx; g.sa1 = 1;
    g.B = w; g.sb0 = 1; g.sb1 = k;
    g.sc0 = m; g.sc1 = 1;
    g.M = n; g.N = m; g.bias_mode = b ? 2 : 0;
  }
  return launch_gemm(g, false, static_cast<cudaStream_t>(stream));
}

extern "C" int mri_dense_backward(const float* x, int64_t ldx, const float* w, const float* pre, const float* grad_y,
                                  int64_t n, int k, int m, int act, float w0, float* dpre, float* grad_x, float* grad_w,
                                  float* grad_b, void* stream) {
  if (n == 0 && w && grad_w) return MRI_OK;
  if (!x || !w || !grad_y || !dpre || !grad_w) return fail(MRI_ERR_INVALID, "dense_backward: null pointer");
  if (act != MRI_ACT_IDENTITY && !pre) return fail(MRI_ERR_INVALID, "dense_backward: pre-activation required");
  if (n < 0 || k < 1 || m < 1 || ldx < k) return fail(MRI_ERR_INVALID, "dense_backward: bad sizes");
  if (act < MRI_ACT_IDENTITY || act > MRI_ACT_RELU) return fail(MRI_ERR_INVALID, "dense_backward: unknown activation %d", act);
  if (n == 0) return MRI_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = n * m;
  const float* pre_or_gy = pre ? pre : grad_y;  // identity never reads it
  if (m <= 256 && 256 % m == 0) {
    int64_t want = (total + 256 * 8 - 1) / (256 * 8);
    const int64_t cap = 8LL * sm_count();
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    dpre_fixedcol_kernel<<<static_cast<int>(want), 256, 0, s>>>(pre_or_gy, grad_y, total, m, act, w0, dpre, grad_b);
  } else {
    int64_t blocks = 8LL * sm_count();
    int64_t rpb = (n + blocks - 1) / blocks;
    if (rpb < 4) rpb = 4;
    blocks = (n + rpb - 1) / rpb;
    dpre_rowslab_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(pre_or_gy, grad_y, n, m, rpb, act, w0, dpre, grad_b);
  }
  MRI_LAUNCH_OK("dpre_kernel");

  if (grad_x) {
    GemmArgs g{};
    g.K = m; g.act = MRI_ACT_IDENTITY; g.w0 = 1.0f; g.bias = nullptr; g.bias_mode = 0; g.accumulate = 0; g.P = nullptr;
    g.C = grad_x;
    if (k <= 16) {
      // dx^T (k x n) = W^T (k x m) * dpre^T (m x n)
      g.A = w; g.sa0 = 1; g.sa1 = k;
      g.B = dpre; g.sb0 = 1; g.sb1 = m;
      g.sc0 = 1; g.sc1 = k;
      g.M = k; g.N = n;
    } else {
      g.A = dpre; g.sa0 = m; g.sa1 = 1;
      g.B = w; g.sb0 = k; g.sb1 = 1;
      g.sc0 = k; g.sc1 = 1;
      g.M = n; g.N = k;
    }
    int st = launch_gemm(g, false, s);
    if (st != MRI_OK) return st;
  }
  {
    // dW (m x k) += dpre^T (m x n) * x (n x k); the smaller of (m, k) goes on M
    GemmArgs g{};
    g.K = n; g.act = MRI_ACT_IDENTITY; g.w0 = 1.0f; g.bias = nullptr; g.bias_mode = 0; g.accumulate = 1; g.P = nullptr;
    g.C = grad_w;
    if (m <= k) {
      g.A = dpre; g.sa0 = 1; g.sa1 = m;
      g.B = x; g.sb0 = ldx; g.sb1 = 1;
      g.sc0 = k; g.sc1 = 1;
      g.M = m; g.N = k;
    } else {
      // dW^T (k x m) += x^T (k x n) * dpre (n x m)
      g.A = x; g.sa0 = 1; g.sa1 = ldx;
      g.B = dpre; g.sb0 = m; g.sb1 = 1;
      g.sc0 = 1; g.sc1 = k;
      g.M = k; g.N = m;
    }
    int st = launch_gemm(g, true, s);
    if (st != MRI_OK) return st;
  }
  return MRI_OK;
}
