// b200seg — memory-bound helpers: weight packing, 2x2 max-pool, nearest 2x upsample, add, layout conversion.
// Reference call sites: nn.MaxPool2d(2,2) AttentionUNet.py:61 / R2U_Net.py:54; nn.Upsample(scale_factor=2)
// AttentionUNet.py:18 / R2U_Net.py:25; x + x1 R2U_Net.py:19,48; input .to(device) helpers.py:318.
#include "common.cuh"

namespace b2 {

__device__ __forceinline__ void unpack8p(const uint4& u, float* f) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8p(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}

// fp32 [cout][cin][k][k] (any strides) -> bf16 [tap][cout][cin] and flipped/transposed [taps-1-tap][cin][cout]
__global__ void pack_weights_kernel(const float* __restrict__ w, int cout, int cin, int taps, int ks, long long s_co,
                                    long long s_ci, long long s_kh, long long s_kw,
                                    __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd,
                                    const float* __restrict__ oscale) {
  const long long total = (long long)taps * cout * cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const int co = (int)((i / cin) % cout);
    const int tap = (int)(i / ((long long)cin * cout));
    const int kh = tap / ks, kw = tap % ks;
    // oscale: per-output-channel factor folded into the weights (eval-mode BatchNorm folding, gamma / sqrt(var+eps))
    const __nv_bfloat16 v =
        __float2bfloat16_rn(w[co * s_co + ci * s_ci + kh * s_kh + kw * s_kw] * (oscale ? oscale[co] : 1.f));
    if (wf) wf[i] = v;
    if (wd) wd[((long long)(taps - 1 - tap) * cin + ci) * cout + co] = v;
  }
}

// UpConv folding: which 3x3 taps collapse onto tap u (0/1) of phase a (0/1): a=0 -> {0},{1,2}; a=1 -> {0,1},{2}
__device__ __forceinline__ bool upfold_member(int a, int u, int r) {
  return a == 0 ? (u == 0 ? r == 0 : r >= 1) : (u == 0 ? r <= 1 : r == 2);
}

// fp32 [cout][cin][3][3] -> bf16 wf [phase][tap][cout][cin], wd [phase][3 - tap][cin][cout]
__global__ void pack_weights_upfold_kernel(const float* __restrict__ w, int cout, int cin, long long s_co,
                                           long long s_ci, long long s_kh, long long s_kw,
                                           __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd,
                                           const float* __restrict__ oscale) {
  const long long total = 16ll * cout * cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const int co = (int)((i / cin) % cout);
    const int pt = (int)(i / ((long long)cin * cout));   // phase * 4 + tap
    const int phase = pt >> 2, tap = pt & 3;
    const int a = phase >> 1, b = phase & 1, u = tap >> 1, v = tap & 1;
    float s = 0.f;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        if (upfold_member(a, u, r) && upfold_member(b, v, c)) s += w[co * s_co + ci * s_ci + r * s_kh + c * s_kw];
    const __nv_bfloat16 val = __float2bfloat16_rn(s * (oscale ? oscale[co] : 1.f));
    if (wf) wf[i] = val;
    if (wd) wd[(((long long)phase * 4 + (3 - tap)) * cin + ci) * cout + co] = val;
  }
}

// dw[co][r*3+c][ci] = sum over (a,u) containing r and (b,v) containing c of dweff[a*2+b][co][u*2+v][ci]
__global__ void fold_upconv_wgrad_kernel(const float* __restrict__ dweff, int cout, int cin,
                                         float* __restrict__ dw) {
  const long long total = 9ll * cout * cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const int tap = (int)((i / cin) % 9);
    const int co = (int)(i / (9ll * cin));
    const int r = tap / 3, c = tap % 3;
    float s = 0.f;
    for (int a = 0; a < 2; ++a)
      for (int u = 0; u < 2; ++u) {
        if (!upfold_member(a, u, r)) continue;
        for (int b = 0; b < 2; ++b)
          for (int v = 0; v < 2; ++v) {
            if (!upfold_member(b, v, c)) continue;
            s += dweff[(((long long)(a * 2 + b) * cout + co) * 4 + (u * 2 + v)) * cin + ci];
          }
      }
    dw[i] = s;
  }
}

__global__ void maxpool2x2_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int n, int h, int w, int cg,
                                      __nv_bfloat16* __restrict__ y, int ldy) {
  pdl_enter();
  const int ho = h / 2, wo = w / 2;
  const long long total = (long long)n * ho * wo * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long p = i / cg;
    const int xo = (int)(p % wo); p /= wo;
    const int yo = (int)(p % ho);
    const int b = (int)(p / ho);
    const long long base = ((long long)(b * h + 2 * yo) * w + 2 * xo);
    float m[8], f[8];
    unpack8p(__ldg(reinterpret_cast<const uint4*>(x + base * ldx + g * 8)), m);
    unpack8p(__ldg(reinterpret_cast<const uint4*>(x + (base + 1) * ldx + g * 8)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
    unpack8p(__ldg(reinterpret_cast<const uint4*>(x + (base + w) * ldx + g * 8)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
    unpack8p(__ldg(reinterpret_cast<const uint4*>(x + (base + w + 1) * ldx + g * 8)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
    *reinterpret_cast<uint4*>(y + ((long long)(b * ho + yo) * wo + xo) * ldy + g * 8) = pack8p(m);
  }
}

// arg-max recomputed from the saved input; first maximum in window scan order wins (ATen's tie rule)
__global__ void maxpool2x2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int lddy,
                                      const __nv_bfloat16* __restrict__ x, int ldx, int n, int h, int w, int cg,
                                      __nv_bfloat16* __restrict__ dx, int lddx,
                                      const __nv_bfloat16* __restrict__ addend, int ldadd) {
  pdl_enter();
  const int ho = h / 2, wo = w / 2;
  const long long total = (long long)n * ho * wo * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long p = i / cg;
    const int xo = (int)(p % wo); p /= wo;
    const int yo = (int)(p % ho);
    const int b = (int)(p / ho);
    const long long base = ((long long)(b * h + 2 * yo) * w + 2 * xo);
    const long long off[4] = {base, base + 1, base + w, base + w + 1};
    float v[4][8], d[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) unpack8p(__ldg(reinterpret_cast<const uint4*>(x + off[k] * ldx + g * 8)), v[k]);
    unpack8p(__ldg(reinterpret_cast<const uint4*>(dy + ((long long)(b * ho + yo) * wo + xo) * lddy + g * 8)), d);
    float o[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int best = 0;
      float bv = v[0][j];
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        if (v[k][j] > bv) {
          bv = v[k][j];
          best = k;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k][j] = (k == best) ? d[j] : 0.f;
    }
    if (addend != nullptr) {     // dx = pool gradient + the gradient x received from its other consumer (one rounding)
      float a[4][8];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        unpack8p(__ldg(reinterpret_cast<const uint4*>(addend + off[k] * ldadd + g * 8)), a[k]);
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) o[k][j] += a[k][j];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(dx + off[k] * lddx + g * 8) = pack8p(o[k]);
  }
}

__global__ void upsample2x_fwd_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int n, int h, int w, int cg,
                                      __nv_bfloat16* __restrict__ y, int ldy) {
  const long long total = (long long)n * h * w * cg;
  const int w2 = 2 * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long p = i / cg;
    const int xi = (int)(p % w); p /= w;
    const int yi = (int)(p % h);
    const int b = (int)(p / h);
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + ((long long)(b * h + yi) * w + xi) * ldx + g * 8));
    const long long ob = ((long long)(b * 2 * h + 2 * yi) * w2 + 2 * xi);
    *reinterpret_cast<uint4*>(y + ob * ldy + g * 8) = u;
    *reinterpret_cast<uint4*>(y + (ob + 1) * ldy + g * 8) = u;
    *reinterpret_cast<uint4*>(y + (ob + w2) * ldy + g * 8) = u;
    *reinterpret_cast<uint4*>(y + (ob + w2 + 1) * ldy + g * 8) = u;
  }
}

__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int lddy, int n, int h, int w, int cg,
                                      __nv_bfloat16* __restrict__ dx, int lddx) {
  const long long total = (long long)n * h * w * cg;
  const int w2 = 2 * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long p = i / cg;
    const int xi = (int)(p % w); p /= w;
    const int yi = (int)(p % h);
    const int b = (int)(p / h);
    const long long ob = ((long long)(b * 2 * h + 2 * yi) * w2 + 2 * xi);
    float s[8], f[8];
    unpack8p(__ldg(reinterpret_cast<const uint4*>(dy + ob * lddy + g * 8)), s);
    unpack8p(__ldg(reinterpret_cast<const uint4*>(dy + (ob + 1) * lddy + g * 8)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += f[j];
    unpack8p(__ldg(reinterpret_cast<const uint4*>(dy + (ob + w2) * lddy + g * 8)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += f[j];
    unpack8p(__ldg(reinterpret_cast<const uint4*>(dy + (ob + w2 + 1) * lddy + g * 8)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += f[j];
    *reinterpret_cast<uint4*>(dx + ((long long)(b * h + yi) * w + xi) * lddx + g * 8) = pack8p(s);
  }
}

__global__ void add_kernel(const __nv_bfloat16* __restrict__ a, int lda, const __nv_bfloat16* __restrict__ b,
                           int ldb, long long npix, int cg, __nv_bfloat16* __restrict__ out, int ldo) {
  pdl_enter();
  const long long total = npix * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    const long long p = i / cg;
    float fa[8], fb[8];
    unpack8p(__ldg(reinterpret_cast<const uint4*>(a + p * lda + g * 8)), fa);
    unpack8p(__ldg(reinterpret_cast<const uint4*>(b + p * ldb + g * 8)), fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) fa[j] += fb[j];
    *reinterpret_cast<uint4*>(out + p * ldo + g * 8) = pack8p(fa);
  }
}

// NCHW fp32 -> NHWC bf16 with the channel dimension zero-padded to ldy (ldy in {4, 8, ...})
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int n, int c, long long hw,
                                    __nv_bfloat16* __restrict__ y, int ldy) {
  const long long total = (long long)n * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / hw, p = i % hw;
    for (int ch = 0; ch < ldy; ++ch) {
      const float v = ch < c ? __ldg(x + (b * c + ch) * hw + p) : 0.f;
      y[i * ldy + ch] = __float2bfloat16_rn(v);
    }
  }
}

static int ew_grid(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}
static bool al16(const void* p, int ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 8 == 0; }

// Image stem as a GEMM: im2col of the 3x3 neighbourhood of a <= 3-channel fp32 NCHW image into 32 bf16 columns per
// pixel (column = tap * CIN + c, zero padded), so that the first convolution (AttentionUNet.py:6, K = 27) and its
// weight gradient run on the tensor-core kernels as a 1x1 convolution with K = 32.
template <int CIN>
__global__ void __launch_bounds__(256) stem_im2col3x3_kernel(const float* __restrict__ x, int n, int h, int w,
                                                             __nv_bfloat16* __restrict__ xc) {
  const long long total = (long long)n * h * w;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int xx = (int)(p % w);
  const int yy = (int)((p / w) % h);
  const long long img = p / ((long long)w * h);
  const float* xi = x + img * CIN * (long long)h * w;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int y2 = yy + t / 3 - 1, x2 = xx + t % 3 - 1;
    const bool in = y2 >= 0 && y2 < h && x2 >= 0 && x2 < w;
#pragma unroll
    for (int c = 0; c < CIN; ++c) v[t * CIN + c] = in ? __ldg(xi + ((long long)c * h + y2) * w + x2) : 0.f;
  }
  uint4* dst = reinterpret_cast<uint4*>(xc + p * 32);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    dst[q] = make_uint4(pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]), pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]),
                        pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]));
}

}  // namespace b2

using namespace b2;

extern "C" int b2_pack_weights(const float* w, int32_t cout, int32_t cin, int32_t ksize, int64_t s_co,
                               int64_t s_ci, int64_t s_kh, int64_t s_kw, void* w_fprop, void* w_dgrad,
                               b2_stream_t stream) {
  B2_REQUIRE(cout > 0 && cin > 0 && ksize >= 1 && ksize <= 7, B2_ERR_SHAPE, "bad weight shape");
  const int taps = ksize * ksize;
  const long long total = (long long)taps * cout * cin;
  pack_weights_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
      w, cout, cin, taps, ksize, s_co, s_ci, s_kh, s_kw, (__nv_bfloat16*)w_fprop, (__nv_bfloat16*)w_dgrad, nullptr);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

// eval-mode BatchNorm folding: W' = W * scale[co] (bf16 fprop packing), b' = b * scale + shift
__global__ void fold_bn_bias_kernel(const float* __restrict__ bias, const float* __restrict__ scale,
                                    const float* __restrict__ shift, int c, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c) out[i] = (bias ? bias[i] : 0.f) * scale[i] + shift[i];
}

extern "C" int b2_pack_weights_folded(const float* w, int32_t cout, int32_t cin, int32_t ksize, int64_t s_co,
                                      int64_t s_ci, int64_t s_kh, int64_t s_kw, const float* scale,
                                      const float* shift, const float* bias, int32_t upfold, void* w_fprop,
                                      float* bias_out, b2_stream_t stream) {
  B2_REQUIRE(cout > 0 && cin > 0 && ksize >= 1 && ksize <= 7 && scale != nullptr && shift != nullptr,
             B2_ERR_SHAPE, "bad folded-weight request");
  if (upfold) {
    B2_REQUIRE(ksize == 3, B2_ERR_SHAPE, "UpConv folding needs a 3x3 kernel");
    pack_weights_upfold_kernel<<<ew_grid(16ll * cout * cin, 256), 256, 0, (cudaStream_t)stream>>>(
        w, cout, cin, s_co, s_ci, s_kh, s_kw, (__nv_bfloat16*)w_fprop, nullptr, scale);
  } else {
    const int taps = ksize * ksize;
    pack_weights_kernel<<<ew_grid((long long)taps * cout * cin, 256), 256, 0, (cudaStream_t)stream>>>(
        w, cout, cin, taps, ksize, s_co, s_ci, s_kh, s_kw, (__nv_bfloat16*)w_fprop, nullptr, scale);
  }
  B2_LAUNCH_CHECK();
  fold_bn_bias_kernel<<<(cout + 127) / 128, 128, 0, (cudaStream_t)stream>>>(bias, scale, shift, cout, bias_out);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_pack_weights_upfold(const float* w, int32_t cout, int32_t cin, int64_t s_co, int64_t s_ci,
                                      int64_t s_kh, int64_t s_kw, void* w_fprop, void* w_dgrad,
                                      b2_stream_t stream) {
  B2_REQUIRE(cout > 0 && cin > 0, B2_ERR_SHAPE, "bad weight shape");
  pack_weights_upfold_kernel<<<ew_grid(16ll * cout * cin, 256), 256, 0, (cudaStream_t)stream>>>(
      w, cout, cin, s_co, s_ci, s_kh, s_kw, (__nv_bfloat16*)w_fprop, (__nv_bfloat16*)w_dgrad, nullptr);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_fold_upconv_wgrad(const float* dweff, int32_t cout, int32_t cin, float* dw, b2_stream_t stream) {
  B2_REQUIRE(cout > 0 && cin > 0, B2_ERR_SHAPE, "bad weight shape");
  fold_upconv_wgrad_kernel<<<ew_grid(9ll * cout * cin, 256), 256, 0, (cudaStream_t)stream>>>(dweff, cout, cin, dw);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_maxpool2x2_fwd(const void* x, int32_t ldx, int32_t n, int32_t h, int32_t w, int32_t c, void* y,
                                 int32_t ldy, b2_stream_t stream) {
  B2_REQUIRE(c % 8 == 0 && h % 2 == 0 && w % 2 == 0, B2_ERR_SHAPE, "maxpool needs c%%8==0 and even h,w");
  B2_REQUIRE(al16(x, ldx) && al16(y, ldy), B2_ERR_ALIGN, "maxpool operands misaligned");
  const long long total = (long long)n * (h / 2) * (w / 2) * (c / 8);
  B2_CHECK_CUDA(launch_chain(maxpool2x2_fwd_kernel, dim3(ew_grid(total, 256)), dim3(256), (size_t)(0),
      (cudaStream_t)stream, 1, total * 64, (const __nv_bfloat16*)x, ldx, n, h, w, c / 8, (__nv_bfloat16*)y, ldy));
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_maxpool2x2_bwd(const void* dy, int32_t lddy, const void* x, int32_t ldx, int32_t n, int32_t h,
                                 int32_t w, int32_t c, void* dx, int32_t lddx, b2_stream_t stream) {
  B2_REQUIRE(c % 8 == 0 && h % 2 == 0 && w % 2 == 0, B2_ERR_SHAPE, "maxpool needs c%%8==0 and even h,w");
  B2_REQUIRE(al16(x, ldx) && al16(dy, lddy) && al16(dx, lddx), B2_ERR_ALIGN, "maxpool operands misaligned");
  const long long total = (long long)n * (h / 2) * (w / 2) * (c / 8);
  B2_CHECK_CUDA(launch_chain(maxpool2x2_bwd_kernel, dim3(ew_grid(total, 256)), dim3(256), (size_t)(0),
      (cudaStream_t)stream, 1, total * 64, (const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)x, ldx, n, h, w,
      c / 8, (__nv_bfloat16*)dx, lddx, nullptr, 0));
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_maxpool2x2_bwd_add(const void* dy, int32_t lddy, const void* x, int32_t ldx, int32_t n, int32_t h,
                                     int32_t w, int32_t c, const void* addend, int32_t ldadd, void* dx, int32_t lddx,
                                     b2_stream_t stream) {
  B2_REQUIRE(c % 8 == 0 && h % 2 == 0 && w % 2 == 0, B2_ERR_SHAPE, "maxpool needs c%%8==0 and even h,w");
  B2_REQUIRE(al16(x, ldx) && al16(dy, lddy) && al16(dx, lddx) && addend != nullptr && al16(addend, ldadd),
             B2_ERR_ALIGN, "maxpool operands misaligned / addend missing");
  const long long total = (long long)n * (h / 2) * (w / 2) * (c / 8);
  B2_CHECK_CUDA(launch_chain(maxpool2x2_bwd_kernel, dim3(ew_grid(total, 256)), dim3(256), (size_t)(0),
      (cudaStream_t)stream, 1, total * 64, (const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)x, ldx, n, h, w,
      c / 8, (__nv_bfloat16*)dx, lddx, (const __nv_bfloat16*)addend, ldadd));
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_upsample2x_fwd(const void* x, int32_t ldx, int32_t n, int32_t h, int32_t w, int32_t c, void* y,
                                 int32_t ldy, b2_stream_t stream) {
  B2_REQUIRE(c % 8 == 0, B2_ERR_SHAPE, "upsample needs c%%8==0");
  B2_REQUIRE(al16(x, ldx) && al16(y, ldy), B2_ERR_ALIGN, "upsample operands misaligned");
  const long long total = (long long)n * h * w * (c / 8);
  upsample2x_fwd_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ldx, n, h,
                                                                                w, c / 8, (__nv_bfloat16*)y, ldy);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_upsample2x_bwd(const void* dy, int32_t lddy, int32_t n, int32_t h, int32_t w, int32_t c,
                                 void* dx, int32_t lddx, b2_stream_t stream) {
  B2_REQUIRE(c % 8 == 0, B2_ERR_SHAPE, "upsample needs c%%8==0");
  B2_REQUIRE(al16(dy, lddy) && al16(dx, lddx), B2_ERR_ALIGN, "upsample operands misaligned");
  const long long total = (long long)n * h * w * (c / 8);
  upsample2x_bwd_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, lddy, n,
                                                                                h, w, c / 8, (__nv_bfloat16*)dx,
                                                                                lddx);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_add(const void* a, int32_t lda, const void* b, int32_t ldb, int64_t npix, int32_t c, void* out,
                      int32_t ldo, b2_stream_t stream) {
  B2_REQUIRE(c % 8 == 0, B2_ERR_SHAPE, "add needs c%%8==0");
  B2_REQUIRE(al16(a, lda) && al16(b, ldb) && al16(out, ldo), B2_ERR_ALIGN, "add operands misaligned");
  B2_CHECK_CUDA(launch_chain(add_kernel, dim3(ew_grid(npix * (c / 8), 256)), dim3(256), (size_t)(0),
      (cudaStream_t)stream, 1, (long long)npix * c * 2, (const __nv_bfloat16*)a, lda, (const __nv_bfloat16*)b, ldb,
      npix, c / 8, (__nv_bfloat16*)out, ldo));
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_nchw_f32_to_nhwc_bf16(const float* x, int32_t n, int32_t c, int32_t h, int32_t w, void* y,
                                        int32_t ldy, b2_stream_t stream) {
  B2_REQUIRE(c > 0 && ldy >= c && ldy <= 64, B2_ERR_SHAPE, "layout conversion supports c <= ldy <= 64");
  const long long total = (long long)n * h * w;
  nchw_to_nhwc_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(x, n, c, (long long)h * w,
                                                                              (__nv_bfloat16*)y, ldy);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

// ------------------------------------------------------------------------------------------------------------
// general layout adapters at the module boundary: NCHW fp32 <-> NHWC bf16 (32x32 smem tile transpose)
// ------------------------------------------------------------------------------------------------------------
namespace b2 {
// x [n][c][hw] fp32 -> y [n][hw][ldy] bf16
__global__ void nchw2nhwc_tile_kernel(const float* __restrict__ x, int c, long long hw, __nv_bfloat16* __restrict__ y,
                                      int ldy) {
  __shared__ float tile[32][33];
  const long long b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int ch = c0 + i;
    const long long p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (ch < c && p < hw) ? x[(b * c + ch) * hw + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long p = p0 + i;
    const int ch = c0 + threadIdx.x;
    if (p < hw && ch < c) y[(b * hw + p) * ldy + ch] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}
// x [n][hw][ldx] bf16 -> y [n][c][hw] fp32
__global__ void nhwc2nchw_tile_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int c, long long hw,
                                      float* __restrict__ y) {
  __shared__ float tile[32][33];
  const long long b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long p = p0 + i;
    const int ch = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < hw && ch < c) ? __bfloat162float(x[(b * hw + p) * ldx + ch]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int ch = c0 + i;
    const long long p = p0 + threadIdx.x;
    if (ch < c && p < hw) y[(b * c + ch) * hw + p] = tile[threadIdx.x][i];
  }
}
}  // namespace b2

extern "C" int b2_layout_nchw_to_nhwc(const float* x, int32_t n, int32_t c, int32_t h, int32_t w, void* y,
                                      int32_t ldy, b2_stream_t stream) {
  B2_REQUIRE(n > 0 && c > 0 && ldy >= c, B2_ERR_SHAPE, "bad layout conversion extent");
  const long long hw = (long long)h * w;
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((c + 31) / 32), (unsigned)n);
  b2::nchw2nhwc_tile_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, c, hw, (__nv_bfloat16*)y, ldy);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_layout_nhwc_to_nchw(const void* x, int32_t ldx, int32_t n, int32_t c, int32_t h, int32_t w,
                                      float* y, b2_stream_t stream) {
  B2_REQUIRE(n > 0 && c > 0 && ldx >= c, B2_ERR_SHAPE, "bad layout conversion extent");
  const long long hw = (long long)h * w;
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((c + 31) / 32), (unsigned)n);
  b2::nhwc2nchw_tile_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ldx, c, hw, y);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_stem_im2col3x3(const float* x, int32_t n, int32_t c, int32_t h, int32_t w, void* xc,
                                 b2_stream_t stream) {
  B2_REQUIRE(n > 0 && h > 0 && w > 0 && c >= 1 && c <= 3, B2_ERR_SHAPE, "stem im2col: %d channels unsupported (1..3)",
             c);
  B2_REQUIRE((reinterpret_cast<uintptr_t>(xc) & 15) == 0, B2_ERR_ALIGN, "xc misaligned");
  const long long total = (long long)n * h * w;
  const unsigned grid = (unsigned)((total + 255) / 256);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(xc);
  if (c == 1) stem_im2col3x3_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, h, w, out);
  else if (c == 2) stem_im2col3x3_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, h, w, out);
  else stem_im2col3x3_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, h, w, out);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
