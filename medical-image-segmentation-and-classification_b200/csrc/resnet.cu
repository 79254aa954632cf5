// b200seg — kernels that only the ResNet-50 encoder of ResNetUnet needs (ResnetUnet.py:32-43; torchvision resnet50):
//   * stem convolution 7x7 / stride 2 / pad 3, 3 -> 64, no bias (CUDA cores: K = 147 with 3 input channels)
//   * MaxPool2d(3, stride 2, pad 1)
// The encoder is frozen in the reference default (ResnetUnet.py:30,45-46,60-66), so these are forward-only.
#include "common.cuh"

namespace b2 {

// x4: NHWC bf16 with 4 (zero padded) channels; wk fp32 [cout=64][49][4]; y NHWC bf16 [n][h/2][w/2][64]
// 256 threads = 64 output pixels x 4 groups of 16 output channels
__global__ void __launch_bounds__(256) stem7x7_fprop_kernel(const uint2* __restrict__ x4, int n, int h, int w,
                                                            const float* __restrict__ wk,
                                                            __nv_bfloat16* __restrict__ y, int ldy) {
  __shared__ __align__(16) float ws[49 * 3 * 64];   // [tap][c][co]
  for (int i = threadIdx.x; i < 49 * 3 * 64; i += blockDim.x) {
    const int co = i & 63, c = (i >> 6) % 3, tap = i / 192;
    ws[i] = bf16_round(wk[(co * 49 + tap) * 4 + c]);
  }
  __syncthreads();
  const int ho = h / 2, wo = w / 2;
  const long long total = (long long)n * ho * wo;
  const int grp = threadIdx.x >> 6;                 // 16-channel group
  for (long long base = (long long)blockIdx.x * 64; base < total; base += (long long)gridDim.x * 64) {
    const long long p = base + (threadIdx.x & 63);
    if (p >= total) continue;
    const int xo = (int)(p % wo);
    const int yo = (int)((p / wo) % ho);
    const long long b = p / ((long long)wo * ho);
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    for (int r = 0; r < 7; ++r) {
      const int yi = 2 * yo + r - 3;
      if (yi < 0 || yi >= h) continue;
#pragma unroll
      for (int s = 0; s < 7; ++s) {
        const int xi = 2 * xo + s - 3;
        if (xi < 0 || xi >= w) continue;
        const uint2 u = __ldg(x4 + (b * h + yi) * w + xi);
        const float in[3] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y)};
        const float* wp = ws + ((r * 7 + s) * 3) * 64 + grp * 16;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 wv = *reinterpret_cast<const float4*>(wp + c * 64 + q * 4);
            acc[q * 4 + 0] = fmaf(in[c], wv.x, acc[q * 4 + 0]);
            acc[q * 4 + 1] = fmaf(in[c], wv.y, acc[q * 4 + 1]);
            acc[q * 4 + 2] = fmaf(in[c], wv.z, acc[q * 4 + 2]);
            acc[q * 4 + 3] = fmaf(in[c], wv.w, acc[q * 4 + 3]);
          }
        }
      }
    }
    uint4* yp = reinterpret_cast<uint4*>(y + p * ldy + grp * 16);
    yp[0] = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                       pack_bf16x2(acc[6], acc[7]));
    yp[1] = make_uint4(pack_bf16x2(acc[8], acc[9]), pack_bf16x2(acc[10], acc[11]), pack_bf16x2(acc[12], acc[13]),
                       pack_bf16x2(acc[14], acc[15]));
  }
}

__device__ __forceinline__ void r_unpack8(const uint4& u, float* f) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}

__global__ void maxpool3x3s2_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int n, int h, int w, int cg,
                                        __nv_bfloat16* __restrict__ y, int ldy) {
  const int ho = h / 2, wo = w / 2;
  const long long total = (long long)n * ho * wo * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long p = i / cg;
    const int xo = (int)(p % wo); p /= wo;
    const int yo = (int)(p % ho);
    const long long b = p / ho;
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int r = -1; r <= 1; ++r) {
      const int yi = 2 * yo + r;
      if (yi < 0 || yi >= h) continue;
      for (int s = -1; s <= 1; ++s) {
        const int xi = 2 * xo + s;
        if (xi < 0 || xi >= w) continue;
        float f[8];
        r_unpack8(__ldg(reinterpret_cast<const uint4*>(x + ((b * h + yi) * w + xi) * ldx + g * 8)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
      }
    }
    *reinterpret_cast<uint4*>(y + ((b * ho + yo) * wo + xo) * ldy + g * 8) =
        make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]),
                   pack_bf16x2(m[6], m[7]));
  }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_stem7x7_fprop(const void* x4, int32_t n, int32_t h, int32_t w, const float* wk, int32_t cout,
                                void* y, int32_t ldy, b2_stream_t stream) {
  B2_REQUIRE(cout == 64, B2_ERR_SHAPE, "7x7 stem supports cout == 64 (got %d)", cout);
  B2_REQUIRE(h % 2 == 0 && w % 2 == 0, B2_ERR_SHAPE, "7x7/s2 stem needs even extents");
  B2_REQUIRE(ldy % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, B2_ERR_ALIGN, "y misaligned");
  const long long total = (long long)n * (h / 2) * (w / 2);
  long long grid = (total + 63) / 64;
  const long long cap = (long long)num_sms() * 16;
  if (grid > cap) grid = cap;
  stem7x7_fprop_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const uint2*)x4, n, h, w, wk,
                                                                         (__nv_bfloat16*)y, ldy);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_maxpool3x3s2_fwd(const void* x, int32_t ldx, int32_t n, int32_t h, int32_t w, int32_t c, void* y,
                                   int32_t ldy, b2_stream_t stream) {
  B2_REQUIRE(c % 8 == 0 && h % 2 == 0 && w % 2 == 0, B2_ERR_SHAPE, "maxpool3x3s2 needs c%%8==0 and even h,w");
  B2_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(y) & 15) == 0,
             B2_ERR_ALIGN, "maxpool operands misaligned");
  const long long total = (long long)n * (h / 2) * (w / 2) * (c / 8);
  long long grid = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (grid > cap) grid = cap;
  maxpool3x3s2_fwd_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ldx, n, h, w,
                                                                            c / 8, (__nv_bfloat16*)y, ldy);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
