// b200seg — fp32 parity mode ("fp32-accumulate mode within 1e-4", BASELINE.json north_star): the INFERENCE path of the
// four models with fp32 activation storage and fp32 FMA accumulation, for callers that need the reference's own fp32
// results (utils/pipeline.py:340-357 runs the segmentation model in fp32 without autocast; its `sigmoid > 0.5` masks
// flip at threshold pixels under bf16 storage).  CUDA cores only: bf16 tensor-core products carry 8 mantissa bits, and
// a 3-term bf16 split of both operands would need every pointwise kernel of the path in a split format as well; this
// mode is about bit-level agreement, not speed (~5 TFLOP/s, AttU_Net 256^2 in ~30 ms per image).
//
//   f32_conv_kernel      direct convolution as a register-tiled GEMM (64 pixels x 64 channels per block, K in chunks
//                        of 16 channels per tap), any ksize <= 7, stride 1|2, explicit padding, K from two tensors
//                        (elided torch.cat), + bias (eval-mode BatchNorm folded in) (+ addend) (ReLU), strided output
//                        placement (ConvTranspose2d k2 s2 = four 1x1 convolutions)
//   f32_gate_tail_kernel sigmoid(BN1(psi . relu(g1 + x1))) * x        (AttentionUNet.py:48-54, after the two 1x1 convs)
//   f32 pool / upsample / layout adapters / weight folding
#include "common.cuh"

namespace b2 {

struct F32Conv {
  const float* x0;
  const float* x1;
  int c0, c1, ld0, ld1;
  int n, hi, wi, ho, wo;
  int ks, stride, pad_h, pad_w;
  const float* w;       // [taps][c0 + c1][cout]
  const float* bias;
  const float* addend;
  int ldadd, add_after_act, relu, cout;
  float* y;
  long long y_sn, y_sh, y_sw;
};

static constexpr int F_BM = 64, F_BN = 64, F_BK = 16;

__global__ void __launch_bounds__(256) f32_conv_kernel(const F32Conv p) {
  __shared__ float As[F_BK][F_BM + 4];
  __shared__ __align__(16) float Bs[F_BK][F_BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;              // 16 x 16 threads, 4 x 4 outputs each
  const long long npix = (long long)p.n * p.ho * p.wo;
  const long long pix0 = (long long)blockIdx.x * F_BM;
  const int co0 = blockIdx.y * F_BN;
  const int ctot = p.c0 + p.c1;
  // A-tile loader role: pixel a_px (0..63), channel quad a_cq (0..3) of the 16-channel chunk
  const int a_px = tid >> 2, a_cq = tid & 3;
  const long long a_pix = pix0 + a_px;
  const bool a_live = a_pix < npix;
  int a_n = 0, a_h = 0, a_w = 0;
  if (a_live) {
    a_w = (int)(a_pix % p.wo);
    a_h = (int)((a_pix / p.wo) % p.ho);
    a_n = (int)(a_pix / ((long long)p.wo * p.ho));
  }
  // B-tile loader role: row b_k (0..15), column quad b_cq (0..15)
  const int b_k = tid >> 4, b_cq = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int taps = p.ks * p.ks;
  const int kchunks = (ctot + F_BK - 1) / F_BK;
  for (int tap = 0; tap < taps; ++tap) {
    const int ih = a_h * p.stride + tap / p.ks - p.pad_h;
    const int iw = a_w * p.stride + tap % p.ks - p.pad_w;
    const bool inb = a_live && ih >= 0 && ih < p.hi && iw >= 0 && iw < p.wi;
    const long long ipix = ((long long)a_n * p.hi + ih) * p.wi + iw;
    for (int kc = 0; kc < kchunks; ++kc) {
      // ---- stage A: 64 pixels x 16 channels (zero padded), stored channel-major
      {
        const int c = kc * F_BK + a_cq * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (inb && c < ctot) {
          const float* src;
          int cc, cend;
          if (c < p.c0) { src = p.x0 + ipix * p.ld0; cc = c; cend = p.c0; }
          else { src = p.x1 + ipix * p.ld1; cc = c - p.c0; cend = p.c1; }
          if (cc + 3 < cend && ((reinterpret_cast<uintptr_t>(src + cc) & 15) == 0)) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(src + cc));
            v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (cc + j < cend) v[j] = __ldg(src + cc + j);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) As[a_cq * 4 + j][a_px] = v[j];
      }
      // ---- stage B: 16 K rows x 64 output channels
      {
        const int k = kc * F_BK + b_k;
        const int co = co0 + b_cq * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (k < ctot) {
          const float* src = p.w + ((long long)tap * ctot + k) * p.cout;
          if (co + 3 < p.cout && (p.cout & 3) == 0) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(src + co));
            v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (co + j < p.cout) v[j] = __ldg(src + co + j);
          }
        }
        *reinterpret_cast<float4*>(&Bs[b_k][b_cq * 4]) = make_float4(v[0], v[1], v[2], v[3]);
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < F_BK; ++k) {
        float a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long pix = pix0 + ty * 4 + i;
    if (pix >= npix) continue;
    const int ow = (int)(pix % p.wo);
    const int oh = (int)((pix / p.wo) % p.ho);
    const int on = (int)(pix / ((long long)p.wo * p.ho));
    float* dst = p.y + on * p.y_sn + oh * p.y_sh + ow * p.y_sw;
    const float* add = p.addend ? p.addend + pix * p.ldadd : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co >= p.cout) continue;
      float v = acc[i][j] + (p.bias ? __ldg(p.bias + co) : 0.f);
      if (add != nullptr && !p.add_after_act) v += __ldg(add + co);
      if (p.relu) v = fmaxf(v, 0.f);
      if (add != nullptr && p.add_after_act) v += __ldg(add + co);
      dst[co] = v;
    }
  }
}

// w [cout][cin][k][k] (any strides) * scale[co] -> out [tap][cin][cout];  bias_out = bias * scale + shift
__global__ void f32_pack_kernel(const float* __restrict__ w, int cout, int cin, int ks, long long s_co, long long s_ci,
                                long long s_kh, long long s_kw, const float* __restrict__ scale,
                                const float* __restrict__ shift, const float* __restrict__ bias,
                                float* __restrict__ out, float* __restrict__ bias_out) {
  const long long total = (long long)ks * ks * cin * cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout);
    const int ci = (int)((i / cout) % cin);
    const int tap = (int)(i / ((long long)cout * cin));
    out[i] = w[co * s_co + ci * s_ci + (tap / ks) * s_kh + (tap % ks) * s_kw] * (scale ? scale[co] : 1.f);
  }
  if (bias_out != nullptr) {
    for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < cout; co += gridDim.x * blockDim.x)
      bias_out[co] = (bias ? bias[co] : 0.f) * (scale ? scale[co] : 1.f) + (shift ? shift[co] : 0.f);
  }
}

// out[p][c] = x[p][c] * sigmoid(s1 * (bpsi + sum_f wpsi[f] * relu(g1[p][f] + x1[p][f])) + h1); one warp per pixel
__global__ void __launch_bounds__(256) f32_gate_tail_kernel(const float* __restrict__ g1, const float* __restrict__ x1,
                                                            int fint, const float* __restrict__ wpsi,
                                                            const float* __restrict__ bpsi,
                                                            const float* __restrict__ scale1,
                                                            const float* __restrict__ shift1,
                                                            const float* __restrict__ x, int ldx, int c, long long npix,
                                                            float* __restrict__ out, int ldo) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float s1 = __ldg(scale1), h1 = __ldg(shift1), bp = bpsi ? __ldg(bpsi) : 0.f;
  for (long long p = warp; p < npix; p += nwarps) {
    float s = 0.f;
    for (int f = lane; f < fint; f += 32)
      s = fmaf(fmaxf(__ldg(g1 + p * fint + f) + __ldg(x1 + p * fint + f), 0.f), __ldg(wpsi + f), s);
    s = warp_sum(s);
    const float q = fmaf(s + bp, s1, h1);
    const float psi = 1.f / (1.f + expf(-q));
    for (int k = lane; k < c; k += 32) out[p * ldo + k] = __ldg(x + p * ldx + k) * psi;
  }
}

__global__ void f32_maxpool_kernel(const float* __restrict__ x, int n, int h, int w, int c, int ks, int stride,
                                   int pad, int ho, int wo, float* __restrict__ y) {
  const long long total = (long long)n * ho * wo * c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    long long q = i / c;
    const int xo = (int)(q % wo); q /= wo;
    const int yo = (int)(q % ho);
    const int b = (int)(q / ho);
    float m = -INFINITY;
    for (int dy = 0; dy < ks; ++dy)
      for (int dx = 0; dx < ks; ++dx) {
        const int yy = yo * stride + dy - pad, xx = xo * stride + dx - pad;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) m = fmaxf(m, __ldg(x + (((long long)b * h + yy) * w + xx) * c + ch));
      }
    y[i] = m;
  }
}

__global__ void f32_upsample2x_kernel(const float* __restrict__ x, int n, int h, int w, int c, float* __restrict__ y) {
  const long long total = (long long)n * 4 * h * w * c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    long long q = i / c;
    const int xo = (int)(q % (2 * w)); q /= 2 * w;
    const int yo = (int)(q % (2 * h));
    const int b = (int)(q / (2 * h));
    y[i] = __ldg(x + (((long long)b * h + yo / 2) * w + xo / 2) * c + ch);
  }
}

// dir 0: NCHW -> NHWC, dir 1: NHWC -> NCHW (fp32, small tensors: plain gather)
__global__ void f32_layout_kernel(const float* __restrict__ x, int n, int c, long long hw, int dir,
                                  float* __restrict__ y) {
  const long long total = (long long)n * c * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    if (dir == 0) {      // i indexes NHWC
      const int ch = (int)(i % c);
      const long long p = (i / c) % hw;
      const long long b = i / (c * hw);
      y[i] = __ldg(x + (b * c + ch) * hw + p);
    } else {             // i indexes NCHW
      const long long p = i % hw;
      const int ch = (int)((i / hw) % c);
      const long long b = i / (hw * c);
      y[i] = __ldg(x + (b * hw + p) * c + ch);
    }
  }
}

static int f_grid(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)num_sms() * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace b2

using namespace b2;

extern "C" int b2_f32_conv(const b2_f32_conv_args* a, b2_stream_t stream) {
  int rc = b2_arch_check();
  if (rc) return rc;
  B2_REQUIRE(a != nullptr && a->n > 0 && a->hi > 0 && a->wi > 0 && a->ho > 0 && a->wo > 0, B2_ERR_SHAPE, "bad extent");
  B2_REQUIRE(a->ksize >= 1 && a->ksize <= 7 && (a->stride == 1 || a->stride == 2), B2_ERR_SHAPE,
             "ksize %d / stride %d unsupported", a->ksize, a->stride);
  B2_REQUIRE(a->c0 > 0 && a->c1 >= 0 && a->cout > 0, B2_ERR_SHAPE, "bad channel counts");
  B2_REQUIRE(a->c1 == 0 || a->c0 % F_BK == 0, B2_ERR_SHAPE, "c0=%d must be a multiple of %d when c1 > 0", a->c0, F_BK);
  F32Conv p;
  p.x0 = a->x0; p.x1 = a->x1; p.c0 = a->c0; p.c1 = a->c1; p.ld0 = a->ldx0; p.ld1 = a->ldx1;
  p.n = a->n; p.hi = a->hi; p.wi = a->wi; p.ho = a->ho; p.wo = a->wo;
  p.ks = a->ksize; p.stride = a->stride; p.pad_h = a->pad_h; p.pad_w = a->pad_w;
  p.w = a->w; p.bias = a->bias; p.addend = a->addend; p.ldadd = a->ldadd; p.add_after_act = a->add_after_act;
  p.relu = a->relu; p.cout = a->cout;
  const int out_mul = a->out_mul == 0 ? 1 : a->out_mul;
  const long long ow = (long long)a->wo * out_mul, oh = (long long)a->ho * out_mul;
  p.y_sw = (long long)out_mul * a->ldy;
  p.y_sh = (long long)out_mul * ow * a->ldy;
  p.y_sn = oh * ow * a->ldy;
  p.y = a->y + ((long long)a->out_off_h * ow + a->out_off_w) * a->ldy;
  const long long npix = (long long)a->n * a->ho * a->wo;
  dim3 grid((unsigned)((npix + F_BM - 1) / F_BM), (unsigned)((a->cout + F_BN - 1) / F_BN));
  f32_conv_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_f32_pack_weights(const float* w, int32_t cout, int32_t cin, int32_t ksize, int64_t s_co, int64_t s_ci,
                                   int64_t s_kh, int64_t s_kw, const float* scale, const float* shift,
                                   const float* bias, float* w_out, float* bias_out, b2_stream_t stream) {
  B2_REQUIRE(cout > 0 && cin > 0 && ksize >= 1 && ksize <= 7, B2_ERR_SHAPE, "bad weight shape");
  const long long total = (long long)ksize * ksize * cin * cout;
  f32_pack_kernel<<<f_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(w, cout, cin, ksize, s_co, s_ci, s_kh, s_kw,
                                                                       scale, shift, bias, w_out, bias_out);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_f32_gate_tail(const float* g1, const float* x1, int32_t fint, const float* wpsi, const float* bpsi,
                                const float* scale1, const float* shift1, const float* x, int32_t ldx, int32_t c,
                                int64_t npix, float* out, int32_t ldo, b2_stream_t stream) {
  B2_REQUIRE(fint > 0 && c > 0 && npix > 0, B2_ERR_SHAPE, "bad gate extent");
  f32_gate_tail_kernel<<<f_grid(npix * 32, 256), 256, 0, (cudaStream_t)stream>>>(g1, x1, fint, wpsi, bpsi, scale1,
                                                                                shift1, x, ldx, c, npix, out, ldo);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_f32_maxpool(const float* x, int32_t n, int32_t h, int32_t w, int32_t c, int32_t ksize,
                              int32_t stride, int32_t pad, float* y, b2_stream_t stream) {
  B2_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && ksize >= 1 && stride >= 1, B2_ERR_SHAPE, "bad pool extent");
  const int ho = (h + 2 * pad - ksize) / stride + 1, wo = (w + 2 * pad - ksize) / stride + 1;
  f32_maxpool_kernel<<<f_grid((long long)n * ho * wo * c, 256), 256, 0, (cudaStream_t)stream>>>(x, n, h, w, c, ksize,
                                                                                               stride, pad, ho, wo, y);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_f32_upsample2x(const float* x, int32_t n, int32_t h, int32_t w, int32_t c, float* y,
                                 b2_stream_t stream) {
  B2_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0, B2_ERR_SHAPE, "bad upsample extent");
  f32_upsample2x_kernel<<<f_grid((long long)n * 4 * h * w * c, 256), 256, 0, (cudaStream_t)stream>>>(x, n, h, w, c, y);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_f32_layout(const float* x, int32_t n, int32_t c, int64_t hw, int32_t to_nchw, float* y,
                             b2_stream_t stream) {
  B2_REQUIRE(n > 0 && c > 0 && hw > 0, B2_ERR_SHAPE, "bad layout extent");
  f32_layout_kernel<<<f_grid((long long)n * c * hw, 256), 256, 0, (cudaStream_t)stream>>>(x, n, c, hw, to_nchw, y);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
