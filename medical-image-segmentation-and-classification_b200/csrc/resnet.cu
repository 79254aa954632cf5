// b200seg — kernels that only the ResNet-50 encoder of ResNetUnet needs (ResnetUnet.py:32-43; torchvision resnet50):
//   * stem convolution 7x7 / stride 2 / pad 3, 3 -> 64, no bias (CUDA cores: K = 147 with 3 input channels)
//   * MaxPool2d(3, stride 2, pad 1)
// The encoder is frozen in the reference default (ResnetUnet.py:30,45-46,60-66): the frozen forward uses the two
// kernels above directly.  ResNetUnet(freeze=False) trains the encoder too (ResnetUnet.py:29-30); its extra pieces:
//   * stem_im2col_kernel: the 7x7/s2 stem as a K = 152 GEMM on the tcgen05 kernels (fprop with BN statistics, wgrad)
//   * maxpool3x3s2_bwd_kernel (gather form: deterministic, first maximum wins as in ATen)
//   * zero_insert2x_kernel: dY of a stride-2 convolution on the stride-1 grid, so that dgrad / wgrad of the strided
//     3x3 and 1x1 convolutions (three of each in ResNet-50) run on the ordinary stride-1 tensor-core kernels
//   * relu_mask_kernel: gradient of relu(bn3(z) + identity) w.r.t. its pre-activation (the Bottleneck residual)
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

// xc[n][yo][xo][tap * c + ch] = x[n][ch][yo*stride + r - pad][xo*stride + s - pad] (0 outside / beyond ks*ks*c);
// one thread per (pixel, 8-column chunk)
__global__ void stem_im2col_kernel(const float* __restrict__ x, int n, int c, int h, int w, int ks, int stride, int pad,
                                   int ho, int wo, int cols, __nv_bfloat16* __restrict__ xc) {
  const int chunks = cols / 8;
  const long long total = (long long)n * ho * wo * chunks;
  const int used = ks * ks * c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ck = (int)(i % chunks);
    long long p = i / chunks;
    const int xo = (int)(p % wo); p /= wo;
    const int yo = (int)(p % ho);
    const long long b = p / ho;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = ck * 8 + j;
      float f = 0.f;
      if (col < used) {
        const int tap = col / c, ch = col % c;
        const int yi = yo * stride + tap / ks - pad, xi = xo * stride + tap % ks - pad;
        if (yi >= 0 && yi < h && xi >= 0 && xi < w) f = __ldg(x + ((b * c + ch) * h + yi) * (long long)w + xi);
      }
      v[j] = f;
    }
    *reinterpret_cast<uint4*>(xc + (i / chunks) * cols + ck * 8) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

// dx[pixel] = sum over the (<= 4) pooling windows that contain it and whose FIRST maximum (row-major scan, ATen's
// rule) sits at this pixel of dy[window]
__global__ void maxpool3x3s2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int lddy,
                                        const __nv_bfloat16* __restrict__ x, int ldx, int n, int h, int w, int cg,
                                        __nv_bfloat16* __restrict__ dx, int lddx) {
  const int ho = h / 2, wo = w / 2;
  const long long total = (long long)n * h * w * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long p = i / cg;
    const int xi = (int)(p % w); p /= w;
    const int yi = (int)(p % h);
    const long long b = p / h;
    float me[8], acc[8];
    r_unpack8(__ldg(reinterpret_cast<const uint4*>(x + ((b * h + yi) * w + xi) * ldx + g * 8)), me);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // windows (yo, xo) with 2*yo - 1 <= yi <= 2*yo + 1
    for (int yo = yi / 2; yo <= (yi + 1) / 2; ++yo) {
      if (yo < 0 || yo >= ho) continue;
      for (int xo = xi / 2; xo <= (xi + 1) / 2; ++xo) {
        if (xo < 0 || xo >= wo) continue;
        bool win[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) win[j] = true;
        for (int r = -1; r <= 1; ++r) {
          const int y2 = 2 * yo + r;
          if (y2 < 0 || y2 >= h) continue;
          for (int s2 = -1; s2 <= 1; ++s2) {
            const int x2 = 2 * xo + s2;
            if (x2 < 0 || x2 >= w || (y2 == yi && x2 == xi)) continue;
            float f[8];
            r_unpack8(__ldg(reinterpret_cast<const uint4*>(x + ((b * h + y2) * w + x2) * ldx + g * 8)), f);
            const bool before = y2 < yi || (y2 == yi && x2 < xi);     // scanned before this pixel: ties go to it
#pragma unroll
            for (int j = 0; j < 8; ++j) win[j] = win[j] && (before ? f[j] < me[j] : f[j] <= me[j]);
          }
        }
        float d[8];
        r_unpack8(__ldg(reinterpret_cast<const uint4*>(dy + ((b * ho + yo) * wo + xo) * lddy + g * 8)), d);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += win[j] ? d[j] : 0.f;
      }
    }
    *reinterpret_cast<uint4*>(dx + ((b * h + yi) * w + xi) * lddx + g * 8) =
        make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                   pack_bf16x2(acc[6], acc[7]));
  }
}

// y[n][2h'][2w'] = x[n][h'][w'], zero elsewhere (16-byte chunks)
__global__ void zero_insert2x_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int n, int h, int w, int cg,
                                     __nv_bfloat16* __restrict__ y, int ldy) {
  const long long total = (long long)n * 4 * h * w * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long p = i / cg;
    const int xo = (int)(p % (2 * w)); p /= 2 * w;
    const int yo = (int)(p % (2 * h));
    const long long b = p / (2 * h);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (((xo | yo) & 1) == 0)
      v = __ldg(reinterpret_cast<const uint4*>(x + ((b * h + yo / 2) * w + xo / 2) * ldx + g * 8));
    *reinterpret_cast<uint4*>(y + ((b * 2 * h + yo) * 2 * w + xo) * ldy + g * 8) = v;
  }
}

// g = out > 0 ? dy : 0
__global__ void relu_mask_kernel(const __nv_bfloat16* __restrict__ dy, int lddy, const __nv_bfloat16* __restrict__ out,
                                 int ldo, long long npix, int cg, __nv_bfloat16* __restrict__ g, int ldg) {
  const long long total = npix * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cg);
    const long long p = i / cg;
    float d[8], o[8];
    r_unpack8(__ldg(reinterpret_cast<const uint4*>(dy + p * lddy + c * 8)), d);
    r_unpack8(__ldg(reinterpret_cast<const uint4*>(out + p * ldo + c * 8)), o);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = o[j] > 0.f ? d[j] : 0.f;
    *reinterpret_cast<uint4*>(g + p * ldg + c * 8) =
        make_uint4(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]), pack_bf16x2(d[4], d[5]), pack_bf16x2(d[6], d[7]));
  }
}

static int r_grid(long long total) {
  long long g = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}
static bool r_al(const void* p, int ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 8 == 0; }

}  // namespace b2

using namespace b2;

extern "C" int b2_stem_im2col(const float* x, int32_t n, int32_t c, int32_t h, int32_t w, int32_t ksize,
                              int32_t stride, int32_t pad, int32_t cols, void* xc, b2_stream_t stream) {
  B2_REQUIRE(n > 0 && c > 0 && ksize >= 1 && ksize <= 7 && stride >= 1, B2_ERR_SHAPE, "bad im2col request");
  B2_REQUIRE(cols % 8 == 0 && cols >= ksize * ksize * c, B2_ERR_SHAPE, "cols=%d must be a multiple of 8 and >= %d", cols,
             ksize * ksize * c);
  B2_REQUIRE((reinterpret_cast<uintptr_t>(xc) & 15) == 0, B2_ERR_ALIGN, "xc misaligned");
  const int ho = (h + 2 * pad - ksize) / stride + 1, wo = (w + 2 * pad - ksize) / stride + 1;
  stem_im2col_kernel<<<r_grid((long long)n * ho * wo * (cols / 8)), 256, 0, (cudaStream_t)stream>>>(
      x, n, c, h, w, ksize, stride, pad, ho, wo, cols, (__nv_bfloat16*)xc);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_maxpool3x3s2_bwd(const void* dy, int32_t lddy, const void* x, int32_t ldx, int32_t n, int32_t h,
                                   int32_t w, int32_t c, void* dx, int32_t lddx, b2_stream_t stream) {
  B2_REQUIRE(c % 8 == 0 && h % 2 == 0 && w % 2 == 0, B2_ERR_SHAPE, "maxpool3x3s2 needs c%%8==0 and even h,w");
  B2_REQUIRE(r_al(dy, lddy) && r_al(x, ldx) && r_al(dx, lddx), B2_ERR_ALIGN, "maxpool operands misaligned");
  maxpool3x3s2_bwd_kernel<<<r_grid((long long)n * h * w * (c / 8)), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)x, ldx, n, h, w, c / 8, (__nv_bfloat16*)dx, lddx);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_zero_insert2x(const void* x, int32_t ldx, int32_t n, int32_t h, int32_t w, int32_t c, void* y,
                                int32_t ldy, b2_stream_t stream) {
  B2_REQUIRE(c % 8 == 0 && n > 0 && h > 0 && w > 0, B2_ERR_SHAPE, "bad zero-insert extent");
  B2_REQUIRE(r_al(x, ldx) && r_al(y, ldy), B2_ERR_ALIGN, "zero-insert operands misaligned");
  zero_insert2x_kernel<<<r_grid((long long)n * 4 * h * w * (c / 8)), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, ldx, n, h, w, c / 8, (__nv_bfloat16*)y, ldy);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_relu_mask(const void* dy, int32_t lddy, const void* out, int32_t ldo, int64_t npix, int32_t c,
                            void* g, int32_t ldg, b2_stream_t stream) {
  B2_REQUIRE(c % 8 == 0 && npix > 0, B2_ERR_SHAPE, "bad relu-mask extent");
  B2_REQUIRE(r_al(dy, lddy) && r_al(out, ldo) && r_al(g, ldg), B2_ERR_ALIGN, "relu-mask operands misaligned");
  relu_mask_kernel<<<r_grid(npix * (c / 8)), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)out, ldo, npix, c / 8, (__nv_bfloat16*)g, ldg);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

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
