// Wide SIREN layers on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
//   D[n, j] = sum_k A[n, k] * B[j, k]        A = activations (rows = coordinates), B = weights (K-major both)
//
// fp32 parity on bf16 tensor cores: every fp32 operand is stored as two bf16 planes, x = hi + lo
// (hi = bf16(x), lo = bf16(x - hi)); the product is accumulated in fp32 TMEM as
//   A_hi*B_hi + A_lo*B_hi + A_hi*B_lo          (3 tcgen05.mma per K-slice, "passes = 3")
// which keeps ~16 mantissa bits per operand (dropped term ~2^-18).  passes = 1 is the plain bf16 mode.
//
// Kernel structure (persistent, one CTA per SM, 192 threads):
//   warp 0     TMA producer: cp.async.bulk.tensor.2d (SWIZZLE_128B boxes of 64 K-elements) into a ring of
//              smem stages, completion on "full" mbarriers
//   warp 1     MMA issuer: one lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=block_n, K=16) against
//              smem descriptors, tcgen05.commit frees the stage ("empty") and publishes the accumulator
//   warps 2-9  epilogue (two warps per TMEM lane quarter, half of the columns each): tcgen05.ld the 128 x block_n
//              fp32 accumulator out of TMEM (double-buffered: 2 x block_n columns), + bias, sin(w0 .) / w0 cos(w0 .)
//              or * mul, stores bf16 hi/lo planes + f32
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace mri {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;          // wgrad kernel: producer, MMA, 4 epilogue warps
constexpr int LAYER_THREADS = 320;        // layer kernel: producer, MMA, 8 epilogue warps (2 per TMEM lane quarter)
constexpr int SMEM_LIMIT = 227 * 1024;

struct LayerParams {
  int64_t n_rows;
  int k, m, block_n, passes, stages;
  int num_m_tiles, num_n_tiles;
  uint32_t stage_bytes, a_plane_bytes, b_plane_bytes;
  int act;
  int b_mn_major;          // 0: B = W (m, k) K-major (forward); 1: B given as (k, m) row-major, i.e. MN-major (dgrad on W itself)
  float w0;
  const float* bias;       // (m) or null
  const float* mul;        // (n_rows, m) or null: elementwise factor applied to the result (dgrad: act')
  __nv_bfloat16* out_hi;   // (n_rows, m) or null
  __nv_bfloat16* out_lo;   // (n_rows, m) or null
  float* out_f32;          // (n_rows, m) or null
  float* aux_f32;          // (n_rows, m) or null: w0*cos(w0*pre) for the backward pass
  float* colsum;           // (m) or null: += column sums of the final `out` (bias gradient when out = dPre)
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"): 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);  // start address, 16-byte units
  d |= static_cast<uint64_t>(1) << 16;                     // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;             // stride byte offset: 8 rows x 128 B
  d |= static_cast<uint64_t>(1) << 46;                     // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                     // SWIZZLE_128B
  return d;
}
// instruction descriptor: bf16 x bf16 -> f32, both operands K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;  // stride between 64-element MN groups
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                   // stride between 8-row K groups
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;                           // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_mnmajor(int m, int n) {
  return make_idesc_bf16(m, n) | (1u << 15) | (1u << 16);  // a_major = b_major = MN
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}


// ---- coalesced epilogue I/O ---------------------------------------------------------------------
// After tcgen05.ld a lane owns one ROW of the tile (32 consecutive columns).  Storing that directly makes every
// STG.128 touch 32 different 128-byte lines (32 L1 wavefronts per instruction), which - not the sine ALU work -
// bounded the epilogue.  Each epilogue warp therefore transposes through a private 32 x 33-word shared-memory
// patch so that one instruction writes 4 full rows of 128 contiguous bytes (f32) or 8 rows of 64 bytes (bf16).
constexpr int EPI_PATCH_WORDS = 32 * 33;

__device__ __forceinline__ void store_tile_f32(const float (&v)[32], float* patch, float* __restrict__ gbase, int64_t ld,
                                               int rows_valid, int lane) {
#pragma unroll
  for (int j = 0; j < 32; ++j) patch[lane * 33 + j] = v[j];
  __syncwarp();
  const int sub_row = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + sub_row;
    const float* src = patch + r * 33 + c4;
    const float4 o = make_float4(src[0], src[1], src[2], src[3]);
    if (r < rows_valid) *reinterpret_cast<float4*>(gbase + static_cast<int64_t>(r) * ld + c4) = o;
  }
  __syncwarp();
}
__device__ __forceinline__ void load_tile_f32(float (&v)[32], float* patch, const float* __restrict__ gbase, int64_t ld,
                                              int rows_valid, int lane) {
  const int sub_row = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + sub_row;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows_valid) o = __ldg(reinterpret_cast<const float4*>(gbase + static_cast<int64_t>(r) * ld + c4));
    float* dst = patch + r * 33 + c4;
    dst[0] = o.x; dst[1] = o.y; dst[2] = o.z; dst[3] = o.w;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = patch[lane * 33 + j];
  __syncwarp();
}
// 32 bf16 per row = 16 packed words
__device__ __forceinline__ void store_tile_bf16(const uint32_t (&w)[16], uint32_t* patch, __nv_bfloat16* __restrict__ gbase,
                                                int64_t ld, int rows_valid, int lane) {
#pragma unroll
  for (int j = 0; j < 16; ++j) patch[lane * 17 + j] = w[j];
  __syncwarp();
  const int sub_row = lane >> 2, w4 = (lane & 3) * 4;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + sub_row;
    const uint32_t* src = patch + r * 17 + w4;
    const uint4 o = make_uint4(src[0], src[1], src[2], src[3]);
    if (r < rows_valid) *reinterpret_cast<uint4*>(gbase + static_cast<int64_t>(r) * ld + 2 * w4) = o;
  }
  __syncwarp();
}

// ---- the GEMM + fused epilogue ------------------------------------------------------------------
__global__ void __launch_bounds__(LAYER_THREADS, 1)
siren_tc_layer_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                      const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                      const LayerParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = p.k / BLOCK_K;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const uint32_t tmem_cols = 2u * static_cast<uint32_t>(p.block_n);  // power of two >= 128 (block_n in {64,128,256})
  // per-epilogue-warp transposition patches live behind the operand ring
  float* epi_patch = reinterpret_cast<float*>(smem + static_cast<size_t>(p.stages) * p.stage_bytes) + (warp >= 2 ? (warp - 2) : 0) * EPI_PATCH_WORDS;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 8);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b_hi)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.num_n_tiles, n_tile = tile % p.num_n_tiles;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* base = smem + static_cast<size_t>(stage) * p.stage_bytes;
          mbar_expect_tx(&full_bar[stage], p.stage_bytes);
          tma_load_2d(base, &map_a_hi, &full_bar[stage], kb * BLOCK_K, m_tile * BLOCK_M);
          uint8_t* bptr = base + p.a_plane_bytes;
          if (p.passes == 3) {
            tma_load_2d(bptr, &map_a_lo, &full_bar[stage], kb * BLOCK_K, m_tile * BLOCK_M);
            bptr += p.a_plane_bytes;
          }
          if (!p.b_mn_major) {
            tma_load_2d(bptr, &map_b_hi, &full_bar[stage], kb * BLOCK_K, n_tile * p.block_n);
            if (p.passes == 3)
              tma_load_2d(bptr + p.b_plane_bytes, &map_b_lo, &full_bar[stage], kb * BLOCK_K, n_tile * p.block_n);
          } else {
            // B stored (k, m) row-major: boxes of 64 output columns x 64 k-rows -> canonical MN-major SW128 groups
            const int groups = p.block_n / 64;
            for (int g = 0; g < groups; ++g)
              tma_load_2d(bptr + g * (BLOCK_K * 128), &map_b_hi, &full_bar[stage], n_tile * p.block_n + g * 64, kb * BLOCK_K);
            if (p.passes == 3)
              for (int g = 0; g < groups; ++g)
                tma_load_2d(bptr + p.b_plane_bytes + g * (BLOCK_K * 128), &map_b_lo, &full_bar[stage],
                            n_tile * p.block_n + g * 64, kb * BLOCK_K);
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = make_idesc_bf16(BLOCK_M, p.block_n) | (p.b_mn_major ? (1u << 16) : 0u);
    int stage = 0;
    uint32_t phase = 0;
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc_stage], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc_stage * p.block_n);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t a_hi = smem_u32(smem + static_cast<size_t>(stage) * p.stage_bytes);
          const uint32_t a_lo = a_hi + p.a_plane_bytes;
          const uint32_t b_hi = a_hi + (p.passes == 3 ? 2 : 1) * p.a_plane_bytes;
          const uint32_t b_lo = b_hi + p.b_plane_bytes;
#pragma unroll
          for (int ks = 0; ks < BLOCK_K / UMMA_K; ++ks) {
            const uint32_t koff = ks * UMMA_K * 2;  // bytes along the swizzled 128-byte row
            const uint32_t koff_mn = ks * UMMA_K * 128;  // MN-major B: 16 k-rows of 128 bytes
            const uint64_t da_hi = make_kmajor_sw128_desc(a_hi + koff);
            const uint64_t db_hi = p.b_mn_major ? make_mnmajor_sw128_desc(b_hi + koff_mn, BLOCK_K * 128)
                                                : make_kmajor_sw128_desc(b_hi + koff);
            if (p.passes == 3) {
              const uint64_t db_lo = p.b_mn_major ? make_mnmajor_sw128_desc(b_lo + koff_mn, BLOCK_K * 128)
                                                  : make_kmajor_sw128_desc(b_lo + koff);
              // small cross terms first, then the dominant hi*hi term
              tcgen05_mma_f16(d_tmem, make_kmajor_sw128_desc(a_lo + koff), db_hi, idesc, (kb | ks) != 0);
              tcgen05_mma_f16(d_tmem, da_hi, db_lo, idesc, 1);
              tcgen05_mma_f16(d_tmem, da_hi, db_hi, idesc, 1);
            } else {
              tcgen05_mma_f16(d_tmem, da_hi, db_hi, idesc, (kb | ks) != 0);
            }
          }
          tcgen05_commit(&empty_bar[stage]);                            // smem stage reusable once these MMAs retire
          if (kb == num_kb - 1) tcgen05_commit(&tmem_full_bar[acc_stage]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
  } else {
    // ================= epilogue (warps 2..9) =================
    const int quarter = warp & 3;          // TMEM lanes [32*quarter, +32) are the ones this warp may read
    const int col_half = (warp - 2) >> 2;  // two warps per quarter: each takes half of the tile's columns
    const int c_begin = col_half * (p.block_n / 2), c_end = c_begin + p.block_n / 2;
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.num_n_tiles, n_tile = tile % p.num_n_tiles;
      mbar_wait(&tmem_full_bar[acc_stage], acc_phase);
      tcgen05_fence_after();
      const int64_t row_base = static_cast<int64_t>(m_tile) * BLOCK_M + quarter * 32;  // first row of this warp's patch
      const int64_t rows_left = p.n_rows - row_base;
      const int rows_valid = rows_left < 0 ? 0 : (rows_left > 32 ? 32 : static_cast<int>(rows_left));
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(acc_stage * p.block_n + c0);
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        if (rows_valid > 0) {  // warp-uniform
          const int col0 = n_tile * p.block_n + c0;
          const int64_t off = row_base * p.m + col0;  // element offset of the patch's top-left corner
          float out[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float acc = __uint_as_float(v[j]);
            if (p.bias) acc += __ldg(p.bias + col0 + j);
            out[j] = acc;
          }
          if (p.act == MRI_ACT_SINE) {
            float aux[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float r = reduce_2pi(p.w0 * out[j]);
              aux[j] = p.w0 * __cosf(r);
              out[j] = __sinf(r);
            }
            if (p.aux_f32) store_tile_f32(aux, epi_patch, p.aux_f32 + off, p.m, rows_valid, lane);
          }
          if (p.mul) {
            float f[32];
            load_tile_f32(f, epi_patch, p.mul + off, p.m, rows_valid, lane);
#pragma unroll
            for (int j = 0; j < 32; ++j) out[j] *= f[j];
          }
          if (p.out_f32) store_tile_f32(out, epi_patch, p.out_f32 + off, p.m, rows_valid, lane);
          if (p.colsum) {  // bias gradient: column sums of this 32 x 32 patch, one atomic per column
#pragma unroll
            for (int j = 0; j < 32; ++j) epi_patch[lane * 33 + j] = out[j];
            __syncwarp();
            float cs = 0.0f;
            for (int r = 0; r < rows_valid; ++r) cs += epi_patch[r * 33 + lane];
            red_add_f32(p.colsum + col0 + lane, cs);
            __syncwarp();
          }
          if (p.out_hi) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float a = out[2 * j], b = out[2 * j + 1];
              const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
              hi[j] = pack_bf16x2(a, b);
              lo[j] = pack_bf16x2(a - __bfloat162float(ah), b - __bfloat162float(bh));
            }
            store_tile_bf16(hi, reinterpret_cast<uint32_t*>(epi_patch), p.out_hi + off, p.m, rows_valid, lane);
            if (p.out_lo) store_tile_bf16(lo, reinterpret_cast<uint32_t*>(epi_patch), p.out_lo + off, p.m, rows_valid, lane);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc_stage]);
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}


// ---- weight gradient: dW[j, i] += sum_n G[n, j] * X[n, i]  (both operands MN-major, K = batch) -------
//
// G = dPre planes (n, m) and X = layer-input planes (n, k) are row-major with the batch index slowest, so
// for this GEMM (M = j, N = i, K = n) both tensor-core operands are "MN-major".  TMA boxes of 64 columns x
// 64 batch rows (SWIZZLE_128B) give the canonical MN-major SW128 layout: 64-element MN groups LBO bytes
// apart, 8-row K groups 1024 B apart.  Split-K over the batch: grid = tiles x splits, each CTA reduces its
// slice of the batch in TMEM and adds the 128 x block_n partial tile into dW with red.global.add.f32.
struct WgradParams {
  int64_t n_rows;
  int m, k, block_n, passes, stages;
  int num_m_tiles, num_n_tiles, splits;
  int64_t rows_per_split;  // multiple of 64
  uint32_t stage_bytes, a_plane_bytes, b_plane_bytes;
  float* grad_w;  // (m, k) f32, accumulated
};

constexpr int WG_BLOCK_K = 64;  // batch rows per stage


__global__ void __launch_bounds__(NUM_THREADS, 1)
siren_tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_g_hi, const __grid_constant__ CUtensorMap map_g_lo,
                      const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                      const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int tile = blockIdx.x % tiles;
  const int split = blockIdx.x / tiles;
  const int m_tile = tile / p.num_n_tiles, n_tile = tile % p.num_n_tiles;
  const int64_t row_begin = static_cast<int64_t>(split) * p.rows_per_split;
  int64_t row_end = row_begin + p.rows_per_split;
  if (row_end > p.n_rows) row_end = p.n_rows;
  const int num_kb = row_end > row_begin ? static_cast<int>((row_end - row_begin + WG_BLOCK_K - 1) / WG_BLOCK_K) : 0;
  const uint32_t tmem_cols = p.block_n < 32 ? 32u : static_cast<uint32_t>(p.block_n);
  const int a_groups = BLOCK_M / 64, b_groups = p.block_n / 64;
  constexpr uint32_t GROUP_BYTES = WG_BLOCK_K * 128;  // one 64-column x 64-row box

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (lane == 0) {
        uint8_t* base = smem + static_cast<size_t>(stage) * p.stage_bytes;
        const int r0 = static_cast<int>(row_begin + static_cast<int64_t>(kb) * WG_BLOCK_K);
        mbar_expect_tx(&full_bar[stage], p.stage_bytes);
        uint8_t* dst = base;
        for (int g = 0; g < a_groups; ++g, dst += GROUP_BYTES) tma_load_2d(dst, &map_g_hi, &full_bar[stage], m_tile * BLOCK_M + g * 64, r0);
        if (p.passes == 3)
          for (int g = 0; g < a_groups; ++g, dst += GROUP_BYTES) tma_load_2d(dst, &map_g_lo, &full_bar[stage], m_tile * BLOCK_M + g * 64, r0);
        for (int g = 0; g < b_groups; ++g, dst += GROUP_BYTES) tma_load_2d(dst, &map_x_hi, &full_bar[stage], n_tile * p.block_n + g * 64, r0);
        if (p.passes == 3)
          for (int g = 0; g < b_groups; ++g, dst += GROUP_BYTES) tma_load_2d(dst, &map_x_lo, &full_bar[stage], n_tile * p.block_n + g * 64, r0);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16_mnmajor(BLOCK_M, p.block_n);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(smem + static_cast<size_t>(stage) * p.stage_bytes);
        const uint32_t a_lo = a_hi + p.a_plane_bytes;
        const uint32_t b_hi = a_hi + (p.passes == 3 ? 2 : 1) * p.a_plane_bytes;
        const uint32_t b_lo = b_hi + p.b_plane_bytes;
#pragma unroll
        for (int ks = 0; ks < WG_BLOCK_K / UMMA_K; ++ks) {
          const uint32_t koff = ks * UMMA_K * 128;  // 16 batch rows of 128 B
          const uint64_t da_hi = make_mnmajor_sw128_desc(a_hi + koff, GROUP_BYTES);
          const uint64_t db_hi = make_mnmajor_sw128_desc(b_hi + koff, GROUP_BYTES);
          if (p.passes == 3) {
            tcgen05_mma_f16(tmem_base, make_mnmajor_sw128_desc(a_lo + koff, GROUP_BYTES), db_hi, idesc, (kb | ks) != 0);
            tcgen05_mma_f16(tmem_base, da_hi, make_mnmajor_sw128_desc(b_lo + koff, GROUP_BYTES), idesc, 1);
            tcgen05_mma_f16(tmem_base, da_hi, db_hi, idesc, 1);
          } else {
            tcgen05_mma_f16(tmem_base, da_hi, db_hi, idesc, (kb | ks) != 0);
          }
        }
        tcgen05_commit(&empty_bar[stage]);
        if (kb == num_kb - 1) tcgen05_commit(&tmem_full_bar);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (num_kb > 0) {
    const int quarter = warp & 3;
    mbar_wait(&tmem_full_bar, 0);
    tcgen05_fence_after();
    const int j = m_tile * BLOCK_M + quarter * 32 + lane;  // output row of dW (always < m: m % 128 == 0)
    for (int c0 = 0; c0 < p.block_n; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c0), v);
      tmem_ld_wait();
      float* dst = p.grad_w + static_cast<int64_t>(j) * p.k + n_tile * p.block_n + c0;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        red_add_v4(dst + 4 * q, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                   __uint_as_float(v[4 * q + 3]));
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// column sums of a (hi [+ lo]) bf16 plane pair: out[j] += sum_n (hi + lo)[n, j]   (bias gradient)
__global__ void __launch_bounds__(256) colsum_planes_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                                            int64_t n, int m, int64_t rows_per_block, float* __restrict__ out) {
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = (r0 + rows_per_block < n) ? r0 + rows_per_block : n;
  for (int j = threadIdx.x; j < m; j += 256) {
    float acc = 0.0f;
    for (int64_t r = r0; r < r1; ++r) {
      acc += __bfloat162float(hi[r * m + j]);
      if (lo) acc += __bfloat162float(lo[r * m + j]);
    }
    red_add_f32(out + j, acc);
  }
}

// (hi, lo) planes of a * b  (first step of the tensor-core backward: dPre = dOut * act')
__global__ void __launch_bounds__(256) mul_split_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t count,
                                                        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    const float x = a[i] * (b ? b[i] : 1.0f);
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[i] = h;
    if (lo) lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

// fp32 -> (hi, lo) bf16 planes
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ src, int64_t count, __nv_bfloat16* __restrict__ hi,
                                                    __nv_bfloat16* __restrict__ lo) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride) {
    const float x = src[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[i] = h;
    if (lo) lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 2-D bf16 row-major (rows, cols) tensor, box = (box_rows, 64 cols), SWIZZLE_128B
int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int box_rows) {
  if (reinterpret_cast<uintptr_t>(ptr) & 15) return fail(MRI_ERR_INVALID, "siren_tc: operand planes must be 16-byte aligned");
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(MRI_ERR_CUDA, "siren_tc: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BLOCK_K), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MRI_ERR_CUDA, "siren_tc: cuTensorMapEncodeTiled failed with %d", static_cast<int>(r));
  return MRI_OK;
}

int pick_block_n(int m) {
  if (m % 256 == 0) return 256;
  if (m % 128 == 0) return 128;
  if (m % 64 == 0) return 64;
  return 0;
}

}  // namespace tc
}  // namespace mri

using namespace mri;

extern "C" int mri_siren_tc_supported(int k, int m) {
  return (k >= 64 && k % 64 == 0 && tc::pick_block_n(m) != 0) ? 1 : 0;
}

extern "C" int mri_siren_tc_split(const float* src, int64_t count, void* hi, void* lo, void* stream) {
  if (!src || !hi) return fail(MRI_ERR_INVALID, "siren_tc_split: null pointer");
  if (count < 0) return fail(MRI_ERR_INVALID, "siren_tc_split: negative count");
  if (count == 0) return MRI_OK;
  int64_t want = (count + 255) / 256;
  const int64_t cap = 16LL * sm_count();
  if (want > cap) want = cap;
  tc::split_kernel<<<static_cast<int>(want), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, count, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo));
  MRI_LAUNCH_OK("split_kernel");
  return MRI_OK;
}

static int siren_tc_layer_impl(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                               int64_t n, int k, int m, int act, float w0, int passes, const float* mul, void* out_hi,
                               void* out_lo, float* out_f32, float* aux_f32, int b_mn_major, float* colsum, void* stream) {
  if (!a_hi || !w_hi) return fail(MRI_ERR_INVALID, "siren_tc_layer: null operand");
  if (passes != 1 && passes != 3) return fail(MRI_ERR_INVALID, "siren_tc_layer: passes must be 1 (bf16) or 3 (split fp32)");
  if (passes == 3 && (!a_lo || !w_lo)) return fail(MRI_ERR_INVALID, "siren_tc_layer: lo planes required for passes=3");
  if (!mri_siren_tc_supported(k, m)) return fail(MRI_ERR_UNSUPPORTED, "siren_tc_layer: k=%d m=%d not tileable (need multiples of 64)", k, m);
  if (act != MRI_ACT_IDENTITY && act != MRI_ACT_SINE) return fail(MRI_ERR_UNSUPPORTED, "siren_tc_layer: activation %d", act);
  if (n < 0) return fail(MRI_ERR_INVALID, "siren_tc_layer: negative n");
  if (n == 0) return MRI_OK;
  if (!out_hi && !out_f32 && !colsum) return fail(MRI_ERR_INVALID, "siren_tc_layer: no output requested");

  tc::LayerParams p{};
  p.n_rows = n; p.k = k; p.m = m; p.passes = passes; p.act = act; p.w0 = w0; p.b_mn_major = b_mn_major;
  p.block_n = tc::pick_block_n(m);
  p.num_m_tiles = static_cast<int>((n + tc::BLOCK_M - 1) / tc::BLOCK_M);
  p.num_n_tiles = m / p.block_n;
  p.a_plane_bytes = tc::BLOCK_M * tc::BLOCK_K * 2;
  p.b_plane_bytes = static_cast<uint32_t>(p.block_n) * tc::BLOCK_K * 2;
  p.stage_bytes = static_cast<uint32_t>(passes == 3 ? 2 : 1) * (p.a_plane_bytes + p.b_plane_bytes);
  const int epi_bytes = 8 * tc::EPI_PATCH_WORDS * static_cast<int>(sizeof(float));  // 8 epilogue warps
  const int budget = tc::SMEM_LIMIT - 1024 - epi_bytes;
  p.stages = budget / static_cast<int>(p.stage_bytes);
  if (p.stages > 8) p.stages = 8;
  if (p.stages < 2) return fail(MRI_ERR_UNSUPPORTED, "siren_tc_layer: tile does not fit shared memory");
  p.bias = bias; p.mul = mul;
  p.out_hi = static_cast<__nv_bfloat16*>(out_hi); p.out_lo = static_cast<__nv_bfloat16*>(out_lo);
  p.out_f32 = out_f32; p.aux_f32 = aux_f32; p.colsum = colsum;

  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int st;
  if ((st = tc::make_map(&ma_hi, a_hi, n, k, tc::BLOCK_M)) != MRI_OK) return st;
  if ((st = tc::make_map(&ma_lo, passes == 3 ? a_lo : a_hi, n, k, tc::BLOCK_M)) != MRI_OK) return st;
  if (!b_mn_major) {
    if ((st = tc::make_map(&mb_hi, w_hi, m, k, p.block_n)) != MRI_OK) return st;
    if ((st = tc::make_map(&mb_lo, passes == 3 ? w_lo : w_hi, m, k, p.block_n)) != MRI_OK) return st;
  } else {  // operand stored (k, m) row-major: boxes of 64 columns x 64 rows
    if ((st = tc::make_map(&mb_hi, w_hi, k, m, tc::BLOCK_K)) != MRI_OK) return st;
    if ((st = tc::make_map(&mb_lo, passes == 3 ? w_lo : w_hi, k, m, tc::BLOCK_K)) != MRI_OK) return st;
  }

  const size_t smem = static_cast<size_t>(p.stages) * p.stage_bytes + 1024 + epi_bytes;
  MRI_CUDA_OK(cudaFuncSetAttribute(tc::siren_tc_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  tc::siren_tc_layer_kernel<<<grid, tc::LAYER_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
  MRI_LAUNCH_OK("siren_tc_layer_kernel");
  return MRI_OK;
}

extern "C" int mri_siren_tc_wgrad(const void* g_hi, const void* g_lo, const void* x_hi, const void* x_lo, int64_t n, int k,
                                  int m, int passes, float* grad_w, float* grad_b, void* stream) {
  if (!g_hi || !x_hi || !grad_w) return fail(MRI_ERR_INVALID, "siren_tc_wgrad: null pointer");
  if (passes != 1 && passes != 3) return fail(MRI_ERR_INVALID, "siren_tc_wgrad: passes must be 1 or 3");
  if (passes == 3 && (!g_lo || !x_lo)) return fail(MRI_ERR_INVALID, "siren_tc_wgrad: lo planes required for passes=3");
  if (m % 128 != 0 || tc::pick_block_n(k) == 0)
    return fail(MRI_ERR_UNSUPPORTED, "siren_tc_wgrad: need m %% 128 == 0 and k %% 64 == 0 (m=%d k=%d)", m, k);
  if (n < 0) return fail(MRI_ERR_INVALID, "siren_tc_wgrad: negative n");
  if (n == 0) return MRI_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  tc::WgradParams p{};
  p.n_rows = n; p.m = m; p.k = k; p.passes = passes; p.grad_w = grad_w;
  p.block_n = tc::pick_block_n(k);
  p.num_m_tiles = m / tc::BLOCK_M;
  p.num_n_tiles = k / p.block_n;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  // two full waves of CTAs: round DOWN so that tiles * splits <= 2 * SMs (ncu: 320 CTAs on 148 SMs left the third
  // wave 84% empty - tensor pipe 78% while active but 56% over the launch)
  int splits = (2 * sm_count()) / tiles;
  const int64_t max_splits = (n + 4 * tc::WG_BLOCK_K - 1) / (4 * tc::WG_BLOCK_K);
  if (splits > max_splits) splits = static_cast<int>(max_splits);
  if (splits < 1) splits = 1;
  int64_t rps = (n + splits - 1) / splits;
  rps = (rps + tc::WG_BLOCK_K - 1) / tc::WG_BLOCK_K * tc::WG_BLOCK_K;
  splits = static_cast<int>((n + rps - 1) / rps);
  p.splits = splits; p.rows_per_split = rps;
  p.a_plane_bytes = tc::BLOCK_M * tc::WG_BLOCK_K * 2;
  p.b_plane_bytes = static_cast<uint32_t>(p.block_n) * tc::WG_BLOCK_K * 2;
  p.stage_bytes = static_cast<uint32_t>(passes == 3 ? 2 : 1) * (p.a_plane_bytes + p.b_plane_bytes);
  p.stages = (tc::SMEM_LIMIT - 2048) / static_cast<int>(p.stage_bytes);
  if (p.stages > 8) p.stages = 8;
  if (p.stages < 2) return fail(MRI_ERR_UNSUPPORTED, "siren_tc_wgrad: tile does not fit shared memory");
  CUtensorMap mg_hi, mg_lo, mx_hi, mx_lo;
  int st;
  if ((st = tc::make_map(&mg_hi, g_hi, n, m, tc::WG_BLOCK_K)) != MRI_OK) return st;
  if ((st = tc::make_map(&mg_lo, passes == 3 ? g_lo : g_hi, n, m, tc::WG_BLOCK_K)) != MRI_OK) return st;
  if ((st = tc::make_map(&mx_hi, x_hi, n, k, tc::WG_BLOCK_K)) != MRI_OK) return st;
  if ((st = tc::make_map(&mx_lo, passes == 3 ? x_lo : x_hi, n, k, tc::WG_BLOCK_K)) != MRI_OK) return st;
  const size_t smem = static_cast<size_t>(p.stages) * p.stage_bytes + 1024;
  MRI_CUDA_OK(cudaFuncSetAttribute(tc::siren_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  tc::siren_tc_wgrad_kernel<<<tiles * splits, tc::NUM_THREADS, smem, s>>>(mg_hi, mg_lo, mx_hi, mx_lo, p);
  MRI_LAUNCH_OK("siren_tc_wgrad_kernel");
  if (grad_b) {
    int64_t blocks = 4LL * sm_count();
    int64_t rpb = (n + blocks - 1) / blocks;
    if (rpb < 8) rpb = 8;
    blocks = (n + rpb - 1) / rpb;
    tc::colsum_planes_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(g_hi),
                                                                     static_cast<const __nv_bfloat16*>(passes == 3 ? g_lo : nullptr),
                                                                     n, m, rpb, grad_b);
    MRI_LAUNCH_OK("colsum_planes_kernel");
  }
  return MRI_OK;
}

extern "C" int mri_siren_tc_mul_split(const float* a, const float* b, int64_t count, void* hi, void* lo, void* stream) {
  if (!a || !hi) return fail(MRI_ERR_INVALID, "siren_tc_mul_split: null pointer");
  if (count < 0) return fail(MRI_ERR_INVALID, "siren_tc_mul_split: negative count");
  if (count == 0) return MRI_OK;
  int64_t want = (count + 255) / 256;
  const int64_t cap = 16LL * sm_count();
  if (want > cap) want = cap;
  tc::mul_split_kernel<<<static_cast<int>(want), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a, b, count, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo));
  MRI_LAUNCH_OK("mul_split_kernel");
  return MRI_OK;
}

extern "C" int mri_siren_tc_layer(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                                  int64_t n, int k, int m, int act, float w0, int passes, const float* mul, void* out_hi,
                                  void* out_lo, float* out_f32, float* aux_f32, void* stream) {
  return siren_tc_layer_impl(a_hi, a_lo, w_hi, w_lo, bias, n, k, m, act, w0, passes, mul, out_hi, out_lo, out_f32, aux_f32, 0,
                             nullptr, stream);
}

extern "C" int mri_siren_tc_dgrad(const void* g_hi, const void* g_lo, const void* w_hi, const void* w_lo, int64_t n, int k,
                                  int m, int passes, const float* mul, void* out_hi, void* out_lo, float* out_f32,
                                  float* colsum, void* stream) {
  // dX (n, k) = G (n, m) . W (m, k): W is consumed as stored (rows = the GEMM's K), i.e. as an MN-major B operand
  return siren_tc_layer_impl(g_hi, g_lo, w_hi, w_lo, nullptr, n, /*gemm K=*/m, /*gemm N=*/k, MRI_ACT_IDENTITY, 1.0f, passes, mul,
                             out_hi, out_lo, out_f32, nullptr, 1, colsum, stream);
}
