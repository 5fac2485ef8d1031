// Image-quality metrics and the linear-in-time baseline on the GPU.
//
//   sq_err_kernel        sum (a - b)^2 in double: MSE / PSNR of legacy_code/hash_experimentation.py:445-453
//                        (skimage.metrics.mean_squared_error / peak_signal_noise_ratio)
//   ssim_kernel          sum of the per-pixel SSIM index over the interior of every (axis 0, axis 1) plane:
//                        skimage.metrics.structural_similarity with its defaults (7 x 7 uniform window, K1 = 0.01,
//                        K2 = 0.03, sample covariance, borders of (win - 1) / 2 pixels cropped before the mean)
//   linear_time_kernel   interp.py:35-52: keep frames ::2 and re-interpolate linearly at the continuous index t / 2
//                        (the ITK LinearInterpolateImageFunction of the reference, clamped at the last kept frame)
//
// All three are streaming kernels over volumes that are a few tens of MB: HBM-bound by construction, one pass each.
#include "common.cuh"

namespace mri {
namespace {

constexpr int MET_THREADS = 256;

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double part[MET_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x < MET_THREADS / 32) s = part[threadIdx.x];
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = MET_THREADS / 64; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  return s;  // valid in thread 0
}

__global__ void __launch_bounds__(MET_THREADS) sq_err_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                             double* __restrict__ out) {
  double acc = 0.0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * MET_THREADS;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * MET_THREADS + threadIdx.x; i < n; i += stride) {
    const double d = static_cast<double>(__ldg(a + i)) - static_cast<double>(__ldg(b + i));
    acc = fma(d, d, acc);
  }
  const double s = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

// a, b: (nx, ny, planes) C-order (planes = product of the trailing axes: the metric is applied slice by slice over
// the first two axes).  One thread per interior pixel; neighbouring threads take neighbouring planes, so every one of
// the win^2 window reads of a warp is one contiguous 128-byte line.
__global__ void __launch_bounds__(MET_THREADS) ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, int nx, int ny,
                                                           int64_t planes, int win, double c1, double c2, double* __restrict__ out) {
  const int p = (win - 1) / 2;
  const int64_t inner_x = nx - 2 * p, inner_y = ny - 2 * p;
  const int64_t total = inner_x * inner_y * planes;
  const double inv = 1.0 / (static_cast<double>(win) * win);
  const double norm = (static_cast<double>(win) * win) / (static_cast<double>(win) * win - 1.0);
  double acc = 0.0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * MET_THREADS;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * MET_THREADS + threadIdx.x; e < total; e += stride) {
    const int64_t k = e % planes, ij = e / planes;
    const int64_t j = ij % inner_y + p, i = ij / inner_y + p;
    double sa = 0.0, sb = 0.0, saa = 0.0, sbb = 0.0, sab = 0.0;
    for (int di = -p; di <= p; ++di) {
      const int64_t row = ((i + di) * ny + (j - p)) * planes + k;
      for (int dj = 0; dj < win; ++dj) {
        const double va = __ldg(a + row + dj * planes), vb = __ldg(b + row + dj * planes);
        sa += va; sb += vb;
        saa = fma(va, va, saa); sbb = fma(vb, vb, sbb); sab = fma(va, vb, sab);
      }
    }
    const double ma = sa * inv, mb = sb * inv;
    const double va = norm * (saa * inv - ma * ma), vb = norm * (sbb * inv - mb * mb), vab = norm * (sab * inv - ma * mb);
    acc += ((2.0 * ma * mb + c1) * (2.0 * vab + c2)) / ((ma * ma + mb * mb + c1) * (va + vb + c2));
  }
  const double s = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

// data, out: (outer, t) C-order (time is the last axis).  Kept frames are data[..., ::2]; output frame f samples them
// at pos = min(f / 2, t_in - 1): v[lo] * (1 - a) + v[hi] * a with separately rounded fp32 products (numpy's arithmetic).
__global__ void __launch_bounds__(MET_THREADS) linear_time_kernel(const float* __restrict__ data, int64_t outer, int t,
                                                                  float* __restrict__ out) {
  const int t_in = (t + 1) / 2;
  const int64_t total = outer * t;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * MET_THREADS;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * MET_THREADS + threadIdx.x; e < total; e += stride) {
    const int64_t o = e / t;
    const int f = static_cast<int>(e - o * t);
    const double pos = fmin(f * 0.5, static_cast<double>(t_in - 1));
    const int lo = static_cast<int>(floor(pos));
    const int hi = min(lo + 1, t_in - 1);
    const float w = static_cast<float>(pos - lo);
    const float vlo = __ldg(data + o * t + 2 * lo), vhi = __ldg(data + o * t + 2 * hi);
    out[e] = __fadd_rn(__fmul_rn(vlo, __fsub_rn(1.0f, w)), __fmul_rn(vhi, w));
  }
}

int grid_for(int64_t work) {
  int64_t blocks = (work + MET_THREADS - 1) / MET_THREADS;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  return static_cast<int>(blocks < 1 ? 1 : blocks > cap ? cap : blocks);
}

}  // namespace
}  // namespace mri

using namespace mri;

extern "C" int mri_sq_err_sum(const float* a, const float* b, int64_t n, double* sum_out, void* stream) {
  if (n < 0) return fail(MRI_ERR_INVALID, "sq_err_sum: negative n");
  if (!sum_out) return fail(MRI_ERR_INVALID, "sq_err_sum: null output");
  if (n == 0) return MRI_OK;
  if (!a || !b) return fail(MRI_ERR_INVALID, "sq_err_sum: null input");
  sq_err_kernel<<<grid_for(n), MET_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a, b, n, sum_out);
  MRI_LAUNCH_OK("sq_err_kernel");
  return MRI_OK;
}

extern "C" int mri_ssim_sum(const float* a, const float* b, int nx, int ny, int64_t planes, int win, double data_range,
                            double* sum_out, void* stream) {
  if (!a || !b || !sum_out) return fail(MRI_ERR_INVALID, "ssim_sum: null pointer");
  if (win < 3 || win % 2 == 0) return fail(MRI_ERR_INVALID, "ssim_sum: window %d must be odd and >= 3", win);
  if (nx < win || ny < win || planes < 1)
    return fail(MRI_ERR_INVALID, "ssim_sum: image %d x %d (x %lld planes) smaller than the %d x %d window", nx, ny,
                static_cast<long long>(planes), win, win);
  const double c1 = (0.01 * data_range) * (0.01 * data_range), c2 = (0.03 * data_range) * (0.03 * data_range);
  const int64_t total = static_cast<int64_t>(nx - win + 1) * (ny - win + 1) * planes;
  ssim_kernel<<<grid_for(total), MET_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a, b, nx, ny, planes, win, c1, c2, sum_out);
  MRI_LAUNCH_OK("ssim_kernel");
  return MRI_OK;
}

extern "C" int mri_linear_time_interp(const float* data, int64_t outer, int t, float* out, void* stream) {
  if (outer < 0 || t < 1) return fail(MRI_ERR_INVALID, "linear_time_interp: bad shape");
  if (outer == 0) return MRI_OK;
  if (!data || !out) return fail(MRI_ERR_INVALID, "linear_time_interp: null pointer");
  linear_time_kernel<<<grid_for(outer * t), MET_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(data, outer, t, out);
  MRI_LAUNCH_OK("linear_time_kernel");
  return MRI_OK;
}
