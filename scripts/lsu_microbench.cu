// LSU / L2 cost model probe for the hash-grid kernels (sm_100a): what does one warp-level gather (LDG.64/128) or
// reduction (RED.v2/v4.f32) cost as a function of how its 32 lane addresses fall into 128-byte lines, 32-byte sectors
// and active lanes?  The answer decides which batch order / lane mapping / aggregation pays in hashdecoder_mma_*.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/lsu_microbench scripts/lsu_microbench.cu
//   ./scripts/lsu_microbench > gpurun_out/lsu_microbench.txt
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

enum Pattern { SPREAD = 0, PAIR16B, QUAD_SECTOR, HALF_LINE, CONTIG, SAME, SPREAD_HALF, SPREAD_QUARTER, PAIRLANE_X0, OCT_64B, PERM_LINE, CHAIN_LINE, CHAIN_LINE_HOT, SPREAD_HOT, DUP2_SPREAD, N_PATTERNS };
static const char* pattern_name[] = {"spread(32 rows, 32 lines)", "pairs in one 16B slot", "quads in one 32B sector", "16 lanes per 128B line",
                                     "contiguous 256B", "same row", "spread, 16 lanes active", "spread, 8 lanes active",
                                     "pair-lane h / h^(x0^(x0+1))", "8 lanes per 64B",
                                     "16 lanes per line, xor-permuted", "chain base^k, base^(k+1), k=lane/2", "chain, 1024 hot lines grid-wide",
                                     "spread over 8192 hot rows", "spread, every address twice"};

// row index (8-byte rows) of this lane for op `k` of iteration `it`
__device__ __forceinline__ uint32_t lane_row(int pattern, uint32_t warp_seed, int lane, uint32_t rows_mask) {
  switch (pattern) {
    case SPREAD: case SPREAD_HALF: case SPREAD_QUARTER: return mix(warp_seed * 32u + lane) & rows_mask;
    case PAIR16B: return ((mix(warp_seed * 32u + (lane >> 1)) << 1) | (lane & 1)) & rows_mask;
    case QUAD_SECTOR: return ((mix(warp_seed * 32u + (lane >> 2)) << 2) | (lane & 3)) & rows_mask;
    case OCT_64B: return ((mix(warp_seed * 32u + (lane >> 3)) << 3) | (lane & 7)) & rows_mask;
    case HALF_LINE: return ((mix(warp_seed * 32u + (lane >> 4)) << 4) | (lane & 15)) & rows_mask;
    case CONTIG: return ((mix(warp_seed) << 5) | lane) & rows_mask;
    case SAME: return mix(warp_seed) & rows_mask;
    case PERM_LINE: {
      const uint32_t h = mix(warp_seed * 32u + (lane >> 4));
      return ((h << 4) | ((lane & 15) ^ (h >> 28))) & rows_mask;
    }
    case CHAIN_LINE: case CHAIN_LINE_HOT: {  // sorted x0-line at a coarse level: 16 samples in cells k = 0..15, lower / upper corner
      uint32_t h = mix(warp_seed);
      if (pattern == CHAIN_LINE_HOT) h &= 1023u;
      const uint32_t base = (h << 5) | ((h >> 27) & 31u);
      return (base ^ ((lane >> 1) + (lane & 1))) & rows_mask;
    }
    case SPREAD_HOT: return mix(warp_seed * 32u + lane) & 8191u;
    case DUP2_SPREAD: return mix(warp_seed * 32u + (lane >> 1)) & rows_mask;
    case PAIRLANE_X0: {
      const uint32_t h = mix(warp_seed * 32u + (lane >> 1));
      const uint32_t x0 = h >> 20;
      return (h ^ ((lane & 1) ? (x0 + 1) : x0)) & rows_mask;
    }
  }
  return 0;
}

__device__ __forceinline__ bool lane_active(int pattern, int lane) {
  if (pattern == SPREAD_HALF) return (lane & 1) == 0;
  if (pattern == SPREAD_QUARTER) return (lane & 3) == 0;
  return true;
}

enum Op { RED_V2 = 0, RED_V4, RED_F32, LDG_64, LDG_128, N_OPS };
static const char* op_name[] = {"red.v2.f32", "red.v4.f32", "red.f32", "ld.v2.f32", "ld.v4.f32"};

template <int OP, int U>
__global__ void __launch_bounds__(128) probe(float* __restrict__ table, uint32_t rows_mask, int pattern, int iters, float* sink, int sm_mod) {
  if (sm_mod > 1) {  // only every sm_mod-th SM works: separates a per-SM (LSU) limit from a chip-wide (L2) one
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (smid % sm_mod) return;
  }
  const int lane = threadIdx.x & 31;
  const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const bool active = lane_active(pattern, lane);
  float acc = 0.0f;
  for (int it = 0; it < iters; ++it) {
    uint32_t rows[U];
#pragma unroll
    for (int k = 0; k < U; ++k) rows[k] = lane_row(pattern, (warp_id * 9973u + it) * U + k, lane, rows_mask);
    if (!active) continue;
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if constexpr (OP == RED_V2) {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(table + 2ull * rows[k]), "f"(1.0f), "f"(2.0f) : "memory");
      } else if constexpr (OP == RED_V4) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(table + 2ull * (rows[k] & ~1u)), "f"(1.0f), "f"(2.0f), "f"(3.0f), "f"(4.0f) : "memory");
      } else if constexpr (OP == RED_F32) {
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(table + 2ull * rows[k]), "f"(1.0f) : "memory");
      } else if constexpr (OP == LDG_64) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(table + 2ull * rows[k]));
        acc += v.x + v.y;
      } else {
        const float4 v = __ldg(reinterpret_cast<const float4*>(table + 2ull * (rows[k] & ~1u)));
        acc += v.x + v.y + v.z + v.w;
      }
    }
  }
  if (acc == 123.456f) *sink = acc;
}

template <int OP>
float run(float* table, uint32_t rows_mask, int pattern, int blocks, int iters, float* sink, int sm_mod = 1) {
  constexpr int U = 8;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  probe<OP, U><<<blocks, 128>>>(table, rows_mask, pattern, iters / 4, sink, sm_mod);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(a));
    probe<OP, U><<<blocks, 128>>>(table, rows_mask, pattern, iters, sink, sm_mod);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  int dev = 0, sms = 0, khz = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  const uint32_t rows = 1u << 23;  // 8-byte rows: 64 MB, L2-resident like the 61 MB of level tables
  float* table;
  float* sink;
  CK(cudaMalloc(&table, 8ull * rows));
  CK(cudaMemset(table, 0, 8ull * rows));
  CK(cudaMalloc(&sink, 4));
  printf("SMs %d, clock %d kHz (nominal), table 64 MB\n", sms, khz);
  printf("%-12s %-34s %7s %6s %9s %12s %14s\n", "op", "pattern", "blk/SM", "iters", "ms", "ns/warp-op/SM", "cyc/warp-op/SM");
  for (int per_sm = 2; per_sm <= 4; per_sm += 2) {
    for (int op = 0; op < N_OPS; ++op) {
      for (int p = 0; p < N_PATTERNS; ++p) {
        const int iters = 256;
        const int blocks = per_sm * sms;
        float ms = 0;
        switch (op) {
          case RED_V2: ms = run<RED_V2>(table, rows - 1, p, blocks, iters, sink); break;
          case RED_V4: ms = run<RED_V4>(table, rows - 1, p, blocks, iters, sink); break;
          case RED_F32: ms = run<RED_F32>(table, rows - 1, p, blocks, iters, sink); break;
          case LDG_64: ms = run<LDG_64>(table, rows - 1, p, blocks, iters, sink); break;
          case LDG_128: ms = run<LDG_128>(table, rows - 1, p, blocks, iters, sink); break;
        }
        const double warp_ops_per_sm = double(per_sm) * 4 * iters * 8;
        const double ns = ms * 1e6 / warp_ops_per_sm;
        printf("%-12s %-34s %7d %6d %9.4f %12.2f %14.1f\n", op_name[op], pattern_name[p], per_sm, iters, ms, ns, ns * 1.965);
      }
    }
  }
  printf("\n-- per-SM or chip-wide?  red.v2.f32 with only every k-th SM active (4 blocks of 128 threads per SM) --\n");
  for (int sm_mod = 1; sm_mod <= 8; sm_mod *= 2)
    for (int p : {int(SPREAD), int(PAIR16B), int(CONTIG), int(CHAIN_LINE)}) {
      const float ms = run<RED_V2>(table, rows - 1, p, 4 * sms, 256, sink, sm_mod);
      const double ns = ms * 1e6 / (4.0 * 4 * 256 * 8);
      printf("every %d-th SM  %-34s %9.4f ms %8.2f ns/warp-op/SM %8.1f cyc\n", sm_mod, pattern_name[p], ms, ns, ns * 1.965);
    }
  printf("\n-- per-SM or chip-wide?  red.v2.f32 with only every k-th SM active (4 blocks of 128 threads per SM) --\n");
  for (int sm_mod = 1; sm_mod <= 8; sm_mod *= 2)
    for (int p : {int(SPREAD), int(PAIR16B), int(CONTIG), int(CHAIN_LINE)}) {
      const float ms = run<RED_V2>(table, rows - 1, p, 4 * sms, 256, sink, sm_mod);
      const double ns = ms * 1e6 / (4.0 * 4 * 256 * 8);
      printf("every %d-th SM  %-34s %9.4f ms %8.2f ns/warp-op/SM %8.1f cyc\n", sm_mod, pattern_name[p], ms, ns, ns * 1.965);
    }
  return 0;
}
