// b200seg — device helpers shared by the tcgen05 convolution kernels (conv_igemm.cu, conv_c64.cu): TMA stores, the
// staged output tile's layout, the BatchNorm-statistics accumulation and the magic-multiplier division.
#pragma once
#include "common.cuh"

namespace b2 {

static constexpr int kTileM = 128;
static constexpr int kKBlock = 64;               // channels per K step (128 B of bf16)

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, const void* smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* tm, const void* smem, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }


// byte offset of (row, channel) inside the staged output tile (swizzled so that both the row-wise 16 B stores of
// the TMEM drain and the column-wise reads of the statistics pass are bank-conflict free; for block_n >= 64 it is
// exactly the TMA SWIZZLE_128B layout of 64-channel panels)
__device__ __forceinline__ uint32_t ctile_off(int block_n, int row, int ch) {
  if (block_n >= 64) {
    const int panel = ch >> 6, cc = ch & 63;
    return (uint32_t)(panel * (kTileM * 128) + row * 128 + ((((cc >> 3) ^ (row & 7))) << 4) + ((cc & 7) << 1));
  }
  return (uint32_t)(row * 64 + ((((ch >> 3) ^ ((row >> 1) & 3))) << 4) + ((ch & 7) << 1));
}

// t / d through the precomputed magic multiplier mg = ceil(2^32 / d); exact while t * d < 2^32 (checked on the host)
__device__ __forceinline__ int fast_div(int t, int d, uint32_t mg) {
  return d == 1 ? t : (int)__umulhi((uint32_t)t, mg);
}

// BatchNorm statistics: one staged 16 B chunk (8 channels of one row) joins the sum / sum of squares of the ROUNDED
// values (fp32; folded into the fp64 accumulators every few tiles)
__device__ __forceinline__ void stats_accum(const uint4 u, float* s, float* q) {
  const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float a = bf16lo(w4[j]), b = bf16hi(w4[j]);
    s[2 * j] += a;
    s[2 * j + 1] += b;
    q[2 * j] = fmaf(a, a, q[2 * j]);
    q[2 * j + 1] = fmaf(b, b, q[2 * j + 1]);
  }
}

}  // namespace b2
