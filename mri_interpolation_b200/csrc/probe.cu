// Measurement probe (bench.py / scripts only, no product path calls it): the chip-wide rate at which the L2 atomic units
// retire 32-byte sector reductions.  The fused decoder-backward + hash-scatter kernel issues one red.global.add.v2.f32
// per (coordinate, level, corner); ncu shows the 61 MB gradient arena L2-resident (DRAM < 10 % busy), so the resource
// that bounds it is not HBM but the number of sector reductions per second the L2 can apply.  This kernel measures that
// peak directly: every lane of every warp instruction reduces into a different, pseudo-random 32-byte sector of an
// L2-resident table (32 sector operations per instruction, no two lanes share a sector, no hot addresses).
#include "common.cuh"

namespace mri {
namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// table: sectors * 8 floats.  lanes_per_sector = 1 (spread) or 2 (the axis-0 pair of the pair-lane mapping: two adjacent
// lanes reduce into neighbouring 8-byte rows of one sector, merged by the LSU into one sector operation).
__global__ void __launch_bounds__(128) red_rate_kernel(float* __restrict__ table, uint32_t sector_mask, int iters, int lanes_per_sector) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t group = lanes_per_sector == 2 ? tid >> 1 : tid;
  const uint32_t sub = lanes_per_sector == 2 ? (tid & 1u) : 0u;
  for (int it = 0; it < iters; ++it) {
    uint32_t s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = mix32(group * 2654435761u + static_cast<uint32_t>(it * 8 + k)) & sector_mask;
#pragma unroll
    for (int k = 0; k < 8; ++k) red_add_v2(table + 8ull * s[k] + 2u * sub, 1.0f, 2.0f);
  }
}

}  // namespace
}  // namespace mri

using namespace mri;

extern "C" int mri_probe_red_rate(float* table, int64_t table_floats, int iters, int lanes_per_sector, int64_t* sector_ops,
                                  void* stream) {
  if (!table || table_floats < 8 || iters < 1 || (lanes_per_sector != 1 && lanes_per_sector != 2))
    return fail(MRI_ERR_INVALID, "probe_red_rate: bad arguments");
  uint32_t sectors = 1;
  while (2ull * sectors * 8 <= static_cast<uint64_t>(table_floats)) sectors *= 2;  // largest power of two that fits
  const int blocks = 4 * sm_count();
  red_rate_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(table, sectors - 1, iters, lanes_per_sector);
  MRI_LAUNCH_OK("red_rate_kernel");
  if (sector_ops) *sector_ops = static_cast<int64_t>(blocks) * 128 * iters * 8 / lanes_per_sector;
  return MRI_OK;
}
