// Shared device/host helpers for the sm_100a kernels behind include/mri_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mri_b200.h"

namespace mri {

// thread-local error text behind mri_last_error()
char* error_buffer();
int fail(int status, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int sm_count();

// One-time per-DEVICE kernel setup (dynamic shared-memory attribute, occupancy query): function attributes and
// occupancy belong to the device that is current at the call, so the cache is indexed by device id.  The cached
// computations are idempotent, so concurrent first calls from several host threads are benign.
constexpr int MRI_MAX_DEVICES = 64;
struct DeviceCache {
  std::atomic<int> slot[MRI_MAX_DEVICES];
  DeviceCache() {
    for (auto& s : slot) s.store(0, std::memory_order_relaxed);
  }
  static int device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MRI_MAX_DEVICES) d = 0;
    return d;
  }
};

#define MRI_CUDA_OK(expr)                                   \
  do {                                                      \
    int _s = ::mri::check_cuda((expr), #expr);              \
    if (_s != MRI_OK) return _s;                            \
  } while (0)

#define MRI_LAUNCH_OK(name)                                 \
  do {                                                      \
    int _s = ::mri::check_cuda(cudaGetLastError(), name);   \
    if (_s != MRI_OK) return _s;                            \
  } while (0)

// ---- hash grid level table passed by value to kernels -------------------------------------
struct LevelDev {
  float res[MRI_MAX_DIM];
  uint32_t rows;
  uint32_t pow2_mask;  // rows-1 if rows is a power of two, else 0xFFFFFFFF marker handled by is_pow2
  uint32_t is_pow2;
  uint32_t magic;      // floor(2^32 / rows): h % rows without a division (wrap_rows<false>)
  uint64_t offset;  // floats
};
struct LevelTable {
  LevelDev lv[MRI_MAX_LEVELS];
};

int make_level_table(const mri_level_t* host_levels, int n_levels, int dim, LevelTable* out);

// encoding.py:40 - the first four of PRIMES (dims beyond 4 are not supported by the kernels)
__host__ __device__ constexpr uint32_t prime(int d) {
  return d == 0 ? 1u : d == 1 ? 2654435761u : d == 2 ? 805459861u : 3674653429u;
}

// Per-coordinate, per-level cell decomposition (encoding.py:111-113, 121-122):
//   xs = x*res (one IEEE multiply, never fused), xi = trunc(xs), xf = xs - float(xi)
//   axis term for the lower corner: (uint32)xi * prime, for the upper: + prime (mod 2^32)
template <int D>
struct Cell {
  uint32_t lo[D];  // hashed lower-corner term per axis
  float wl[D];     // weight of the lower corner (1 - xf)
  float wu[D];     // weight of the upper corner (xf)
};

template <int D>
__device__ __forceinline__ Cell<D> make_cell(const float (&x)[D], const LevelDev& lv) {
  Cell<D> c;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float xs = __fmul_rn(x[d], lv.res[d]);
    const int xi = __float2int_rz(xs);
    const float xf = __fsub_rn(xs, __int2float_rn(xi));
    c.lo[d] = static_cast<uint32_t>(xi) * prime(d);
    c.wu[d] = xf;
    c.wl[d] = __fsub_rn(1.0f, xf);
  }
  return c;
}

// corner n: bit d set -> upper index on axis d (encoding.py:101-106 bin_mask is the negation)
template <int D>
__device__ __forceinline__ uint32_t corner_hash(const Cell<D>& c, int n) {
  uint32_t h = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) h ^= ((n >> d) & 1) ? (c.lo[d] + prime(d)) : c.lo[d];
  return h;
}
template <int D>
__device__ __forceinline__ float corner_weight(const Cell<D>& c, int n) {
  float w = ((n & 1) ? c.wu[0] : c.wl[0]);
#pragma unroll
  for (int d = 1; d < D; ++d) w = __fmul_rn(w, ((n >> d) & 1) ? c.wu[d] : c.wl[d]);
  return w;
}
// encoding.py:78 - true modulo; power-of-two tables take the mask shortcut.  POW2 is a template
// parameter so the (block-uniform) choice is one branch per thread, not a predicate per corner.
// Non-power-of-two tables: q = mulhi(h, floor(2^32 / rows)) is the true quotient or one less (the estimate falls short
// of h / rows by less than 1), so one conditional subtraction gives exactly h % rows - 4 instructions instead of the
// ~25 of a 32-bit division.
__device__ __forceinline__ uint32_t exact_mod(uint32_t h, uint32_t rows, uint32_t magic) {
  const uint32_t r = h - __umulhi(h, magic) * rows;
  return r >= rows ? r - rows : r;
}
template <bool POW2>
__device__ __forceinline__ uint32_t wrap_rows(uint32_t h, const LevelDev& lv) {
  if constexpr (POW2) return h & lv.pow2_mask;
  else return exact_mod(h, lv.rows, lv.magic);
}

template <int D>
__device__ __forceinline__ void load_coord(const float* __restrict__ x, int64_t i, float (&v)[D]) {
  if constexpr (D == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(x) + i);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (D == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(x) + i);
    v[0] = t.x; v[1] = t.y;
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) v[d] = __ldg(x + i * D + d);
  }
}

// erf-GELU (nn.GELU default, approximate='none') and its derivative.  Phi(x) = 0.5 (1 + erf(x/sqrt2)) is
// evaluated with the Abramowitz-Stegun 7.1.26 rational form (|abs err| <= 1.5e-7, far inside the 1e-3 parity
// bound): one MUFU.RCP + one MUFU.EX2 + 5 FMA, and exp(-x^2/2) is shared with the Gaussian pdf of the derivative.
// The SFU approximations are issued directly (rcp.approx / ex2.approx, ~1-2 ulp): __frcp_rn adds a Newton step and
// a range-check branch with a slow-path call per element, __expf a denormal fix-up - in the fused backward kernel this
// function is evaluated 32 times per lane and tile and was a quarter of all issued instructions (ncu source page).
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void gelu_cdf_pdf(float x, float& cdf, float& pdf) {
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x), 1.0f));
  const float e = ex2_approx(x * x * -0.72134752044448170368f);  // exp(-x^2 / 2) = exp(-z^2), z = |x| / sqrt(2)
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);  // the 0.5 of 0.5 erfc(z) folded into the coefficients
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  const float half_erfc = poly * t * e;  // 0.5 * erfc(z)
  cdf = x >= 0.0f ? 1.0f - half_erfc : half_erfc;
  pdf = 0.39894228040143267794f * e;
}
__device__ __forceinline__ float gelu_f(float x) {
  float cdf, pdf;
  gelu_cdf_pdf(x, cdf, pdf);
  return x * cdf;
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float cdf, pdf;
  gelu_cdf_pdf(x, cdf, pdf);
  return cdf + x * pdf;
}

// sin/cos of w0*pre with an explicit 2-term Cody-Waite reduction in front of the SFU, so the
// result stays ~1e-6 absolute for the |arg| <~ 1e3 a SIREN produces (MUFU alone degrades with |arg|).
__device__ __forceinline__ float reduce_2pi(float a) {
  const float k = rintf(a * 0.15915494309189533577f);
  float r = fmaf(-k, 6.2831854820251464844f, a);     // 2*pi rounded to f32
  r = fmaf(-k, -1.7484555314695172e-7f, r);           // 2*pi - f32(2*pi)
  return r;
}
__device__ __forceinline__ float fast_sin(float a) { return __sinf(reduce_2pi(a)); }
__device__ __forceinline__ float fast_cos(float a) { return __cosf(reduce_2pi(a)); }

template <int ACT>
__device__ __forceinline__ float activate(float pre, float w0) {
  if constexpr (ACT == MRI_ACT_SINE) return fast_sin(w0 * pre);
  else if constexpr (ACT == MRI_ACT_GELU) return gelu_f(pre);
  else if constexpr (ACT == MRI_ACT_RELU) return fmaxf(pre, 0.0f);
  else return pre;
}
template <int ACT>
__device__ __forceinline__ float activate_grad(float pre, float w0) {
  if constexpr (ACT == MRI_ACT_SINE) return w0 * fast_cos(w0 * pre);
  else if constexpr (ACT == MRI_ACT_GELU) return gelu_grad_f(pre);
  else if constexpr (ACT == MRI_ACT_RELU) return pre > 0.0f ? 1.0f : 0.0f;
  else return 1.0f;
}
__device__ __forceinline__ float activate_rt(int act, float pre, float w0) {
  switch (act) {
    case MRI_ACT_SINE: return activate<MRI_ACT_SINE>(pre, w0);
    case MRI_ACT_GELU: return activate<MRI_ACT_GELU>(pre, w0);
    case MRI_ACT_RELU: return activate<MRI_ACT_RELU>(pre, w0);
    default: return pre;
  }
}
__device__ __forceinline__ float activate_grad_rt(int act, float pre, float w0) {
  switch (act) {
    case MRI_ACT_SINE: return activate_grad<MRI_ACT_SINE>(pre, w0);
    case MRI_ACT_GELU: return activate_grad<MRI_ACT_GELU>(pre, w0);
    case MRI_ACT_RELU: return activate_grad<MRI_ACT_RELU>(pre, w0);
    default: return 1.0f;
  }
}

__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

}  // namespace mri
