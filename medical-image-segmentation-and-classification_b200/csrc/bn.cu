// b200seg — BatchNorm2d kernels (memory-bound; 16-byte vector accesses, warp/block reductions, fp64 accumulators).
// Reference call sites: nn.BatchNorm2d at models/segmentation_models/AttentionUNet.py:7,10,21,34,38,42,
// R2U_Net.py:11,28, ResnetUnet.py:8,11,55 (eps 1e-5, momentum 0.1, biased variance for normalisation, unbiased for
// running_var); backward = native_batch_norm_backward + threshold_backward fused.
//
// Thread mapping shared by every kernel here: a thread owns ONE group of 8 consecutive channels (one uint4 of
// bf16) and walks over pixels, so per-channel coefficients live in registers and a warp reads consecutive 16-byte
// chunks of consecutive pixels (fully coalesced for ld == C).
#include <stdlib.h>

#include <string.h>

#include "common.cuh"

namespace b2 {

static constexpr int kBlock = 256;

struct ChanMap {
  int tpp;    // threads per pixel == number of 8-channel groups
  int rows;   // pixels per block iteration
};

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ void load8f(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Sum acc[0..7] over the `rows` threads that own the same channel group; result valid in threads with r == 0.
__device__ __forceinline__ void rows_reduce8(float* acc, int tpp, int rows, int g, int r, bool active, float* red) {
  __syncthreads();
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[(r * tpp + g) * 8 + j] = acc[j];
  }
  __syncthreads();
  if (active && r == 0) {
    for (int rr = 1; rr < rows; ++rr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += red[(rr * tpp + g) * 8 + j];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// per-channel sum / sum of squares of a bf16 tensor (used where the conv epilogue could not produce them)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) channel_stats_kernel(const __nv_bfloat16* __restrict__ z, int ldz,
                                                               long long npix, int C, ChanMap m,
                                                               double* __restrict__ stats, int want_sq, DetBuf det) {
  pdl_enter();
  __shared__ float red[kBlock * 8];
  const int t = threadIdx.x;
  const int g = t % m.tpp, r = t / m.tpp;
  const bool active = r < m.rows;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (active) {
    for (long long p = (long long)blockIdx.x * m.rows + r; p < npix; p += (long long)gridDim.x * m.rows) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(z + p * ldz + g * 8));
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        q[j] += f[j] * f[j];
      }
    }
  }
  rows_reduce8(s, m.tpp, m.rows, g, r, active, red);
  if (active && r == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red_out(stats, det, g * 8 + j, (double)s[j]);
  }
  if (want_sq) {
    rows_reduce8(q, m.tpp, m.rows, g, r, active, red);
    if (active && r == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red_out(stats, det, C + g * 8 + j, (double)q[j]);
    }
  }
}

// db[c] = sum_p dy[p][c]  (two-stage: fp64 scratch is avoided by a single-block-per-channel-slab final pass)
__global__ void __launch_bounds__(kBlock) channel_sum_kernel(const __nv_bfloat16* __restrict__ dy, int lddy,
                                                             long long npix, int C, ChanMap m,
                                                             float* __restrict__ db, DetBuf det) {
  pdl_enter();
  __shared__ float red[kBlock * 8];
  // gridDim.y channel slices of m.tpp * 8 channels: a block ends in one atomic per channel of its slice, and with ~1200
  // blocks on the same C addresses those atomics were a fixed ~100 us per launch whatever the tensor size
  const int c_off = blockIdx.y * (m.tpp * 8);
  dy += c_off;
  const int t = threadIdx.x;
  const int g = t % m.tpp, r = t / m.tpp;
  const bool active = r < m.rows && c_off + g * 8 < C;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (active) {
    const long long step = (long long)gridDim.x * m.rows;
    for (long long p = (long long)blockIdx.x * m.rows + r; p < npix; p += 2 * step) {      // two pixels in flight
      const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(dy + p * lddy + g * 8));
      const uint4 u1 = p + step < npix ? __ldg(reinterpret_cast<const uint4*>(dy + (p + step) * lddy + g * 8))
                                       : make_uint4(0u, 0u, 0u, 0u);
      float f[8], h[8];
      unpack8(u0, f);
      unpack8(u1, h);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += f[j] + h[j];
    }
  }
  rows_reduce8(s, m.tpp, m.rows, g, r, active, red);
  if (active && r == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red_out(db, det, c_off + g * 8 + j, s[j]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// finalize: batch mean / invstd, affine coefficients, running statistics (momentum update, unbiased variance)
// ------------------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const double* __restrict__ stats, int C, long long count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* running_mean, float* running_var,
                                   long long* num_batches_tracked, float* mean, float* invstd, float* scale,
                                   float* shift) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
  if (c >= C) return;
  const double m = stats[c] / (double)count;
  double var = __fma_rn(-m, m, stats[C + c] / (double)count);
  if (var < 0.0) var = 0.0;
  const float fm = (float)m;
  const float is = (float)(1.0 / sqrt(var + (double)eps));
  if (mean != nullptr) {   // (NULL outputs: running-statistics update only)
    mean[c] = fm;
    invstd[c] = is;
    const float sc = (gamma ? gamma[c] : 1.f) * is;
    scale[c] = sc;
    shift[c] = (beta ? beta[c] : 0.f) - fm * sc;
  }
  if (running_mean != nullptr) {
    const double unbiased = count > 1 ? var * (double)count / (double)(count - 1) : var;
    // (explicit rounding points: the multi-layer kernel below must give the same bits)
    running_mean[c] = __fmaf_rn(momentum, fm, __fmul_rn(1.f - momentum, running_mean[c]));
    running_var[c] = __fmaf_rn(momentum, (float)unbiased, __fmul_rn(1.f - momentum, running_var[c]));
  }
}

// running statistics of MANY BatchNorm layers in one launch (b2_bn_update_running_multi): block (layer, 256-channel slab)
struct BnRunRef {      // mirrors b2_bn_run_ref
  const double* stats;
  float* running_mean;
  float* running_var;
  long long* num_batches_tracked;
  long long count;
  int c;
  float momentum;
};
static constexpr int kRunBatch = 64;      // layers per launch: the references travel as a kernel PARAMETER (3 KB), so
struct BnRunBatch {                       // there is no device table to build or keep alive — safe under graph capture
  BnRunRef r[kRunBatch];
};
__global__ void __launch_bounds__(256) bn_update_running_multi_kernel(const __grid_constant__ BnRunBatch batch) {
  const BnRunRef& r = batch.r[blockIdx.x];
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c == 0) *r.num_batches_tracked += 1;
  if (c >= r.c) return;
  const double m = r.stats[c] / (double)r.count;
  double var = __fma_rn(-m, m, r.stats[r.c + c] / (double)r.count);
  if (var < 0.0) var = 0.0;
  const double unbiased = r.count > 1 ? var * (double)r.count / (double)(r.count - 1) : var;
  r.running_mean[c] = __fmaf_rn(r.momentum, (float)m, __fmul_rn(1.f - r.momentum, r.running_mean[c]));
  r.running_var[c] = __fmaf_rn(r.momentum, (float)unbiased, __fmul_rn(1.f - r.momentum, r.running_var[c]));
}

__global__ void bn_eval_coeffs_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps, int C,
                                      float* mean, float* invstd, float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float is = rsqrtf(rv[c] + eps);
  const float isx = (float)(1.0 / sqrt((double)rv[c] + (double)eps));
  (void)is;
  mean[c] = rm[c];
  invstd[c] = isx;
  const float sc = (gamma ? gamma[c] : 1.f) * isx;
  scale[c] = sc;
  shift[c] = (beta ? beta[c] : 0.f) - rm[c] * sc;
}

// ------------------------------------------------------------------------------------------------------------
// apply: y = act(z*scale + shift)   [+ ysum = y + addend]
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) bn_apply_kernel(const __nv_bfloat16* __restrict__ z, int ldz,
                                                          long long npix, ChanMap m,
                                                          const float* __restrict__ scale,
                                                          const float* __restrict__ shift, int relu,
                                                          __nv_bfloat16* __restrict__ y, int ldy,
                                                          const __nv_bfloat16* __restrict__ addend, int ldadd,
                                                          __nv_bfloat16* __restrict__ ysum, int ldysum) {
  pdl_enter();
  const int t = threadIdx.x;
  const int g = t % m.tpp, r = t / m.tpp;
  if (r >= m.rows) return;
  float sc[8], sh[8];
  load8f(scale + g * 8, sc);
  load8f(shift + g * 8, sh);
  if (addend == nullptr) {
    // plain y = act(z * scale + shift): four pixels in flight per thread (with one, a thread had a single 16-byte load
    // outstanding and the kernel ran at 82 % of the copy bandwidth)
    constexpr int U = 4;
    const long long step = (long long)gridDim.x * m.rows;
    for (long long p0 = (long long)blockIdx.x * m.rows + r; p0 < npix; p0 += U * step) {
      uint4 u[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const long long p = p0 + k * step;
        u[k] = p < npix ? __ldg(reinterpret_cast<const uint4*>(z + p * ldz + g * 8)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const long long p = p0 + k * step;
        if (p >= npix) break;
        float f[8];
        unpack8(u[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[j] = fmaf(f[j], sc[j], sh[j]);
          if (relu) f[j] = fmaxf(f[j], 0.f);
        }
        *reinterpret_cast<uint4*>(y + p * ldy + g * 8) = pack8(f);
      }
    }
    return;
  }
  if (ysum != nullptr) {
    // x + act(bn(z)) (Recurrent_block, R2U_Net.py:19): two pixels = four 16-byte loads in flight per thread
    constexpr int U = 2;
    const long long step = (long long)gridDim.x * m.rows;
    for (long long p0 = (long long)blockIdx.x * m.rows + r; p0 < npix; p0 += U * step) {
      uint4 u[U], a[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const long long p = p0 + k * step;
        const bool ok = p < npix;
        u[k] = ok ? __ldg(reinterpret_cast<const uint4*>(z + p * ldz + g * 8)) : make_uint4(0u, 0u, 0u, 0u);
        a[k] = ok ? __ldg(reinterpret_cast<const uint4*>(addend + p * ldadd + g * 8)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const long long p = p0 + k * step;
        if (p >= npix) break;
        float f[8], fa[8], fo[8];
        unpack8(u[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[j] = fmaf(f[j], sc[j], sh[j]);
          if (relu) f[j] = fmaxf(f[j], 0.f);
        }
        const uint4 o = pack8(f);
        if (y != nullptr) *reinterpret_cast<uint4*>(y + p * ldy + g * 8) = o;
        unpack8(a[k], fa);
        unpack8(o, fo);   // the sum is taken on the ROUNDED activation, as torch does (x + x1 on bf16 tensors)
#pragma unroll
        for (int j = 0; j < 8; ++j) fo[j] += fa[j];
        *reinterpret_cast<uint4*>(ysum + p * ldysum + g * 8) = pack8(fo);
      }
    }
    return;
  }
  // residual mode (torchvision Bottleneck: relu(bn3(z) + identity)): the addend joins BEFORE the activation
  for (long long p = (long long)blockIdx.x * m.rows + r; p < npix; p += (long long)gridDim.x * m.rows) {
    float f[8], fa[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(z + p * ldz + g * 8)), f);
    unpack8(__ldg(reinterpret_cast<const uint4*>(addend + p * ldadd + g * 8)), fa);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = bf16_round(fmaf(f[j], sc[j], sh[j])) + fa[j];
      if (relu) f[j] = fmaxf(f[j], 0.f);
    }
    *reinterpret_cast<uint4*>(y + p * ldy + g * 8) = pack8(f);
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward pass 1: sums[0][c] = sum dy*mask, sums[1][c] = sum dy*mask*xhat
// The loop only needs the mask coefficients: sum dy*mask*xhat = invstd*(sum dy*mask*z - mean*sum dy*mask), taken in
// fp64 at the end.  Two pixels per iteration keep four 16-byte loads in flight per thread.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock, 3) bn_bwd_reduce_kernel(
    const __nv_bfloat16* __restrict__ dy, int lddy, const __nv_bfloat16* __restrict__ z, int ldz, long long npix,
    int C, ChanMap m, const float* __restrict__ scale, const float* __restrict__ shift,
    const float* __restrict__ mean, const float* __restrict__ invstd, int relu, double* __restrict__ sums,
    DetBuf det) {
  pdl_enter();
  __shared__ float red[kBlock * 8];
  const int t = threadIdx.x;
  const int g = t % m.tpp, r = t / m.tpp;
  const bool active = r < m.rows;
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s0[j] = s1[j] = 0.f;
  if (active) {
    float sc[8], sh[8];
    load8f(scale + g * 8, sc);
    load8f(shift + g * 8, sh);
    const long long step = (long long)gridDim.x * m.rows;
    long long p = (long long)blockIdx.x * m.rows + r;
    for (; p + step < npix; p += 2 * step) {
      const uint4 ud0 = __ldg(reinterpret_cast<const uint4*>(dy + p * lddy + g * 8));
      const uint4 uz0 = __ldg(reinterpret_cast<const uint4*>(z + p * ldz + g * 8));
      const uint4 ud1 = __ldg(reinterpret_cast<const uint4*>(dy + (p + step) * lddy + g * 8));
      const uint4 uz1 = __ldg(reinterpret_cast<const uint4*>(z + (p + step) * ldz + g * 8));
      float d[8], f[8];
      unpack8(ud0, d);
      unpack8(uz0, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dd = (!relu || fmaf(f[j], sc[j], sh[j]) > 0.f) ? d[j] : 0.f;
        s0[j] += dd;
        s1[j] = fmaf(dd, f[j], s1[j]);
      }
      unpack8(ud1, d);
      unpack8(uz1, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dd = (!relu || fmaf(f[j], sc[j], sh[j]) > 0.f) ? d[j] : 0.f;
        s0[j] += dd;
        s1[j] = fmaf(dd, f[j], s1[j]);
      }
    }
    for (; p < npix; p += step) {
      const uint4 ud = __ldg(reinterpret_cast<const uint4*>(dy + p * lddy + g * 8));
      const uint4 uz = __ldg(reinterpret_cast<const uint4*>(z + p * ldz + g * 8));
      float d[8], f[8];
      unpack8(ud, d);
      unpack8(uz, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dd = (!relu || fmaf(f[j], sc[j], sh[j]) > 0.f) ? d[j] : 0.f;
        s0[j] += dd;
        s1[j] = fmaf(dd, f[j], s1[j]);
      }
    }
  }
  rows_reduce8(s0, m.tpp, m.rows, g, r, active, red);
  float keep0[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) keep0[j] = s0[j];
  rows_reduce8(s1, m.tpp, m.rows, g, r, active, red);
  if (active && r == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      const double a0 = (double)keep0[j];
      const double a1 = (double)invstd[c] * ((double)s1[j] - (double)mean[c] * a0);
      red_out(sums, det, c, a0);
      red_out(sums, det, C + c, a1);
    }
  }
}

// backward pass 2: dz = gamma*invstd*(dy*mask - [training](s0/m + xhat*s1/m)) = A*dy*mask + B*z + Cc per channel;
// optionally dbias[c] += sum_p dz (the bias gradient of the convolution that produced z, taken on the rounded dz
// the conv backward consumes).  Two pixels per iteration.
__global__ void __launch_bounds__(kBlock, 2) bn_bwd_apply_kernel(
    const __nv_bfloat16* __restrict__ dy, int lddy, const __nv_bfloat16* __restrict__ z, int ldz, long long npix,
    int C, ChanMap m, const float* __restrict__ scale, const float* __restrict__ shift,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma, int relu,
    int training, const double* __restrict__ sums, __nv_bfloat16* __restrict__ dz, int lddz,
    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias, DetBuf det) {
  pdl_enter();
  __shared__ float red[kBlock * 8];
  const int t = threadIdx.x;
  const int g = t % m.tpp, r = t / m.tpp;
  const bool active = r < m.rows;
  float bsum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bsum[j] = 0.f;
  if (active) {
    float sc[8], sh[8], ka[8], kb[8], kc[8];
    load8f(scale + g * 8, sc);
    load8f(shift + g * 8, sh);
    const double inv_m = 1.0 / (double)npix;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      const float ga = gamma ? __ldg(gamma + c) : 1.f;
      const float is = __ldg(invstd + c), mu = __ldg(mean + c);
      const double a0 = sums[c], a1 = sums[C + c];
      const float k0 = training ? (float)(a0 * inv_m) : 0.f;
      const float k1 = training ? (float)(a1 * inv_m) : 0.f;
      ka[j] = ga * is;
      kb[j] = -ka[j] * k1 * is;
      kc[j] = -ka[j] * k0 - kb[j] * mu;
      if (blockIdx.x == 0 && r == 0) {
        if (dgamma) dgamma[c] = (float)a1;
        if (dbeta) dbeta[c] = (float)a0;
      }
    }
    const long long step = (long long)gridDim.x * m.rows;
    long long p = (long long)blockIdx.x * m.rows + r;
    for (; p < npix; p += 2 * step) {
      const bool two = p + step < npix;
      const long long p1 = two ? p + step : p;
      const uint4 ud0 = __ldg(reinterpret_cast<const uint4*>(dy + p * lddy + g * 8));
      const uint4 uz0 = __ldg(reinterpret_cast<const uint4*>(z + p * ldz + g * 8));
      const uint4 ud1 = __ldg(reinterpret_cast<const uint4*>(dy + p1 * lddy + g * 8));
      const uint4 uz1 = __ldg(reinterpret_cast<const uint4*>(z + p1 * ldz + g * 8));
      float d[8], f[8];
      unpack8(ud0, d);
      unpack8(uz0, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dd = (!relu || fmaf(f[j], sc[j], sh[j]) > 0.f) ? d[j] : 0.f;
        f[j] = fmaf(ka[j], dd, fmaf(kb[j], f[j], kc[j]));
      }
      uint4 o = pack8(f);
      *reinterpret_cast<uint4*>(dz + p * lddz + g * 8) = o;
      if (dbias != nullptr) {
        unpack8(o, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) bsum[j] += f[j];
      }
      if (two) {
        unpack8(ud1, d);
        unpack8(uz1, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dd = (!relu || fmaf(f[j], sc[j], sh[j]) > 0.f) ? d[j] : 0.f;
          f[j] = fmaf(ka[j], dd, fmaf(kb[j], f[j], kc[j]));
        }
        o = pack8(f);
        *reinterpret_cast<uint4*>(dz + p1 * lddz + g * 8) = o;
        if (dbias != nullptr) {
          unpack8(o, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) bsum[j] += f[j];
        }
      }
    }
  }
  if (dbias != nullptr) {
    rows_reduce8(bsum, m.tpp, m.rows, g, r, active, red);
    if (active && r == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red_out(dbias, det, g * 8 + j, bsum[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// "Light" BatchNorm-backward kernels: 4 channels per thread (8-byte loads), four pixels in flight.  Half the
// per-channel state of the 8-channel kernels -> <= 64 registers, so several blocks fit next to a persistent
// weight-gradient CTA that runs on the side stream and the kernels overlap with it.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack4(const uint2& u, float* f) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
}
__device__ __forceinline__ void load4f(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
}
// Sum acc[0..3] over the `rows` threads that own the same channel group; result valid in threads with r == 0.
__device__ __forceinline__ void rows_reduce4(float* acc, int tpp, int rows, int g, int r, bool active, float* red) {
  __syncthreads();
  if (active) {
#pragma unroll
    for (int j = 0; j < 4; ++j) red[(r * tpp + g) * 4 + j] = acc[j];
  }
  __syncthreads();
  if (active && r == 0) {
    for (int rr = 1; rr < rows; ++rr) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += red[(rr * tpp + g) * 4 + j];
    }
  }
}

static constexpr int kLightPix = 4;     // pixels in flight per thread

__global__ void __launch_bounds__(kBlock, 4) bn_bwd_reduce_light_kernel(
    const __nv_bfloat16* __restrict__ dy, int lddy, const __nv_bfloat16* __restrict__ z, int ldz, long long npix,
    int C, ChanMap m, const float* __restrict__ scale, const float* __restrict__ shift,
    const float* __restrict__ mean, const float* __restrict__ invstd, int relu, double* __restrict__ sums,
    DetBuf det) {
  pdl_enter();
  __shared__ float red[kBlock * 4];
  // gridDim.y channel slices of m.tpp * 4 channels: a block ends in 2 * (slice width) atomics instead of 2 C, which is
  // what bounded the many-channel layers (1184 blocks x 1024 fp64 atomics at C = 512: 0.115 ms for 0.024 ms of data)
  const int c_off = blockIdx.y * (m.tpp * 4);
  dy += c_off;
  z += c_off;
  scale += c_off;
  shift += c_off;
  const int t = threadIdx.x;
  const int g = t % m.tpp, r = t / m.tpp;
  const bool active = r < m.rows;
  float s0[4], s1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) s0[j] = s1[j] = 0.f;
  if (active) {
    float sc[4], sh[4];
    load4f(scale + g * 4, sc);
    load4f(shift + g * 4, sh);
    const long long step = (long long)gridDim.x * m.rows;
    for (long long p0 = (long long)blockIdx.x * m.rows + r; p0 < npix; p0 += kLightPix * step) {
      uint2 ud[kLightPix], uz[kLightPix];
#pragma unroll
      for (int u = 0; u < kLightPix; ++u) {
        const long long p = p0 + u * step;
        const bool ok = p < npix;
        ud[u] = ok ? __ldg(reinterpret_cast<const uint2*>(dy + p * lddy + g * 4)) : make_uint2(0u, 0u);
        uz[u] = ok ? __ldg(reinterpret_cast<const uint2*>(z + p * ldz + g * 4)) : make_uint2(0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < kLightPix; ++u) {
        float d[4], f[4];
        unpack4(ud[u], d);          // out-of-range pixels were loaded as dy = 0: they add nothing
        unpack4(uz[u], f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float dd = (!relu || fmaf(f[j], sc[j], sh[j]) > 0.f) ? d[j] : 0.f;
          s0[j] += dd;
          s1[j] = fmaf(dd, f[j], s1[j]);
        }
      }
    }
  }
  rows_reduce4(s0, m.tpp, m.rows, g, r, active, red);
  float keep0[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) keep0[j] = s0[j];
  rows_reduce4(s1, m.tpp, m.rows, g, r, active, red);
  if (active && r == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c_off + g * 4 + j;
      const double a0 = (double)keep0[j];
      const double a1 = (double)invstd[c] * ((double)s1[j] - (double)mean[c] * a0);
      red_out(sums, det, c, a0);
      red_out(sums, det, C + c, a1);
    }
  }
}

__global__ void __launch_bounds__(kBlock, 4) bn_bwd_apply_light_kernel(
    const __nv_bfloat16* __restrict__ dy, int lddy, const __nv_bfloat16* __restrict__ z, int ldz, long long npix,
    int C, ChanMap m, const float* __restrict__ scale, const float* __restrict__ shift,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma, int relu,
    int training, const double* __restrict__ sums, __nv_bfloat16* __restrict__ dz, int lddz,
    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias, DetBuf det) {
  pdl_enter();
  __shared__ float red[kBlock * 4];
  // gridDim.y channel slices of m.tpp * 4 channels (as in the reduce kernel): with C / 4 threads per pixel a thread of a
  // 512- or 1024-channel layer met only a handful of pixels, and its prologue (six per-channel coefficient loads, two of
  // them fp64, and the fp64 coefficient algebra) cost more than its share of the tensor: 32^2 x 512 ran at 2.2 TB/s,
  // 16^2 x 1024 at 1.1 TB/s.  A 64-channel slice gives every thread 16x the pixels for the same prologue.
  const int c_off = blockIdx.y * (m.tpp * 4);
  dy += c_off;
  z += c_off;
  dz += c_off;
  const int t = threadIdx.x;
  const int g = t % m.tpp, r = t / m.tpp;
  const bool active = r < m.rows;
  float bsum[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bsum[j] = 0.f;
  if (active) {
    float sc[4], sh[4], ka[4], kb[4], kc[4];
    load4f(scale + c_off + g * 4, sc);
    load4f(shift + c_off + g * 4, sh);
    const double inv_m = 1.0 / (double)npix;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c_off + g * 4 + j;
      const float ga = gamma ? __ldg(gamma + c) : 1.f;
      const float is = __ldg(invstd + c), mu = __ldg(mean + c);
      const double a0 = sums[c], a1 = sums[C + c];
      const float k0 = training ? (float)(a0 * inv_m) : 0.f;
      const float k1 = training ? (float)(a1 * inv_m) : 0.f;
      ka[j] = ga * is;
      kb[j] = -ka[j] * k1 * is;
      kc[j] = -ka[j] * k0 - kb[j] * mu;
      if (blockIdx.x == 0 && r == 0) {
        if (dgamma) dgamma[c] = (float)a1;
        if (dbeta) dbeta[c] = (float)a0;
      }
    }
    const long long step = (long long)gridDim.x * m.rows;
    for (long long p0 = (long long)blockIdx.x * m.rows + r; p0 < npix; p0 += kLightPix * step) {
      uint2 ud[kLightPix], uz[kLightPix];
#pragma unroll
      for (int u = 0; u < kLightPix; ++u) {
        const long long p = p0 + u * step;
        const bool ok = p < npix;
        ud[u] = ok ? __ldg(reinterpret_cast<const uint2*>(dy + p * lddy + g * 4)) : make_uint2(0u, 0u);
        uz[u] = ok ? __ldg(reinterpret_cast<const uint2*>(z + p * ldz + g * 4)) : make_uint2(0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < kLightPix; ++u) {
        const long long p = p0 + u * step;
        if (p >= npix) break;
        float d[4], f[4];
        unpack4(ud[u], d);
        unpack4(uz[u], f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float dd = (!relu || fmaf(f[j], sc[j], sh[j]) > 0.f) ? d[j] : 0.f;
          f[j] = fmaf(ka[j], dd, fmaf(kb[j], f[j], kc[j]));
        }
        const uint2 o = make_uint2(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]));
        *reinterpret_cast<uint2*>(dz + p * lddz + g * 4) = o;
        if (dbias != nullptr) {
          unpack4(o, f);
#pragma unroll
          for (int j = 0; j < 4; ++j) bsum[j] += f[j];
        }
      }
    }
  }
  if (dbias != nullptr) {
    rows_reduce4(bsum, m.tpp, m.rows, g, r, active, red);
    if (active && r == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) red_out(dbias, det, c_off + g * 4 + j, bsum[j]);
    }
  }
}

// channel map of the light kernels: C / 4 threads per pixel (C <= 4 * kBlock)
static bool light_map(int C, ChanMap* m) {
  static const bool enabled = [] {
    const char* e = getenv("B200SEG_BN_LIGHT");
    return !(e != nullptr && atoi(e) == 0);
  }();
  if (!enabled || C % 4 != 0 || C / 4 > kBlock) return false;
  m->tpp = C / 4;
  m->rows = kBlock / m->tpp;
  return true;
}

static int make_map(int C, ChanMap* m) {
  B2_REQUIRE(C > 0 && C % 8 == 0, B2_ERR_SHAPE, "channel count %d must be a positive multiple of 8", C);
  const int cg = C / 8;
  B2_REQUIRE(cg <= kBlock, B2_ERR_SHAPE, "channel count %d > %d unsupported", C, kBlock * 8);
  m->tpp = cg;
  m->rows = kBlock / cg;
  return B2_OK;
}
static int chan_grid(long long npix, const ChanMap& m, int waves) {
  long long need = (npix + m.rows - 1) / m.rows;
  long long cap = (long long)num_sms() * waves;
  if (need > cap) need = cap;
  if (need < 1) need = 1;
  return (int)need;
}
static bool aligned16(const void* p, int ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 8 == 0; }

}  // namespace b2

using namespace b2;

extern "C" int b2_channel_stats(const void* z, int32_t ldz, int64_t npix, int32_t c, double* stats,
                                b2_stream_t stream) {
  ChanMap m;
  int rc = make_map(c, &m);
  if (rc) return rc;
  B2_REQUIRE(aligned16(z, ldz), B2_ERR_ALIGN, "z misaligned");
  const int grid = chan_grid(npix, m, 8);
  DetBuf det;
  rc = det_begin(&det, grid, 2 * c, (cudaStream_t)stream);
  if (rc) return rc;
  B2_CHECK_CUDA(launch_chain(channel_stats_kernel, dim3(grid), dim3(kBlock), (size_t)(0), (cudaStream_t)stream, 1,
      (long long)npix * c * 2, (const __nv_bfloat16*)z, ldz, npix, c, m, stats, 1, det));
  B2_LAUNCH_CHECK();
  if (det.partial) return det_finish(det.partial, grid, det.n, 2 * c, stats, (cudaStream_t)stream);
  return B2_OK;
}

extern "C" int b2_channel_sum(const void* dy, int32_t lddy, int64_t npix, int32_t c, float* db,
                              b2_stream_t stream) {
  ChanMap m;
  int rc = make_map(c, &m);
  if (rc) return rc;
  B2_REQUIRE(aligned16(dy, lddy), B2_ERR_ALIGN, "dy misaligned");
  B2_CHECK_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * c, (cudaStream_t)stream));
  int slices = 1;
  if (c > 64 && env_switch("B200SEG_BN_SLICE", 1) != 0) {      // 64-channel slices: 8 threads per pixel, 32 pixels per pass
    slices = (c + 63) / 64;
    m.tpp = 8;
    m.rows = kBlock / 8;
  }
  int grid = chan_grid(npix, m, 4);
  grid = (grid + slices - 1) / slices;
  if (grid < 1) grid = 1;
  DetBuf det;
  rc = det_begin(&det, grid, c, (cudaStream_t)stream);
  if (rc) return rc;
  B2_CHECK_CUDA(launch_chain(channel_sum_kernel, dim3(grid, slices), dim3(kBlock), (size_t)(0), (cudaStream_t)stream,
      1, (long long)npix * c * 2, (const __nv_bfloat16*)dy, lddy, npix, c, m, db, det));
  B2_LAUNCH_CHECK();
  if (det.partial) return det_finish(det.partial, grid, det.n, c, db, (cudaStream_t)stream);
  return B2_OK;
}

extern "C" int b2_bn_finalize(const double* stats, int32_t c, int64_t count, const float* gamma,
                              const float* beta, float eps, float momentum, float* running_mean,
                              float* running_var, int64_t* num_batches_tracked, float* mean, float* invstd,
                              float* scale, float* shift, b2_stream_t stream) {
  B2_REQUIRE(c > 0 && count > 0, B2_ERR_SHAPE, "bad bn_finalize extent");
  B2_CHECK_CUDA(launch_chain(bn_finalize_kernel, dim3((c + 127) / 128), dim3(128), (size_t)(0), (cudaStream_t)stream,
      1, 0, stats, c, count, gamma, beta, eps, momentum, running_mean, running_var,
      reinterpret_cast<long long*>(num_batches_tracked), mean, invstd, scale, shift));
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_bn_update_running_multi(const b2_bn_run_ref* refs, int32_t n, int32_t max_c, b2_stream_t stream) {
  static_assert(sizeof(BnRunRef) == sizeof(b2_bn_run_ref), "BnRunRef must mirror b2_bn_run_ref");
  B2_REQUIRE(refs != nullptr && n > 0 && max_c > 0, B2_ERR_SHAPE, "empty running-statistics launch");
  for (int i0 = 0; i0 < n; i0 += kRunBatch) {
    const int m = n - i0 < kRunBatch ? n - i0 : kRunBatch;
    BnRunBatch batch;
    memcpy(batch.r, reinterpret_cast<const BnRunRef*>(refs) + i0, sizeof(BnRunRef) * (size_t)m);
    bn_update_running_multi_kernel<<<dim3(m, (max_c + 255) / 256), 256, 0, (cudaStream_t)stream>>>(batch);
    B2_LAUNCH_CHECK();
  }
  return B2_OK;
}

extern "C" int b2_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean,
                                 const float* running_var, float eps, int32_t c, float* mean, float* invstd,
                                 float* scale, float* shift, b2_stream_t stream) {
  B2_REQUIRE(c > 0, B2_ERR_SHAPE, "bad channel count");
  bn_eval_coeffs_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean, running_var,
                                                                          eps, c, mean, invstd, scale, shift);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_bn_apply(const void* z, int32_t ldz, int64_t npix, int32_t c, const float* scale,
                           const float* shift, int32_t relu, void* y, int32_t ldy, const void* addend,
                           int32_t ldadd, void* ysum, int32_t ldysum, b2_stream_t stream) {
  ChanMap m;
  int rc = make_map(c, &m);
  if (rc) return rc;
  B2_REQUIRE(aligned16(z, ldz) && (y == nullptr ? ysum != nullptr : aligned16(y, ldy)), B2_ERR_ALIGN,
             "bn_apply operands misaligned (y may be NULL only when ysum is given)");
  B2_REQUIRE(ysum == nullptr || (addend != nullptr && aligned16(ysum, ldysum)), B2_ERR_ALIGN,
             "bn_apply ysum misaligned or addend missing");
  B2_REQUIRE(addend == nullptr || aligned16(addend, ldadd), B2_ERR_ALIGN, "bn_apply addend misaligned");
  B2_CHECK_CUDA(launch_chain(bn_apply_kernel, dim3(chan_grid(npix, m, 16)), dim3(kBlock), (size_t)(0),
      (cudaStream_t)stream, 1, (long long)npix * c * 2, (const __nv_bfloat16*)z, ldz, npix, m, scale, shift, relu,
      (__nv_bfloat16*)y, ldy, (const __nv_bfloat16*)addend, ldadd, (__nv_bfloat16*)ysum, ldysum));
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_bn_bwd_reduce(const void* dy, int32_t lddy, const void* z, int32_t ldz, int64_t npix,
                                int32_t c, const float* scale, const float* shift, const float* mean,
                                const float* invstd, int32_t relu, double* sums, b2_stream_t stream) {
  ChanMap m;
  int rc = make_map(c, &m);
  if (rc) return rc;
  B2_REQUIRE(aligned16(dy, lddy) && aligned16(z, ldz), B2_ERR_ALIGN, "bn_bwd_reduce operands misaligned");
  ChanMap lm;
  const bool light = light_map(c, &lm);
  int slices = 1;
  if (light && c % 64 == 0 && c > 64 && env_switch("B200SEG_BN_SLICE", 1) != 0) {
    slices = c / 64;                 // 64-channel slices: 16 threads per pixel, 16 pixels per block iteration
    lm.tpp = 16;
    lm.rows = kBlock / 16;
  }
  int grid = light ? chan_grid(npix, lm, 8) : chan_grid(npix, m, 4);
  if (slices > 1) {
    grid = (grid + slices - 1) / slices;
    if (grid < 1) grid = 1;
  }
  DetBuf det;
  rc = det_begin(&det, grid, 2 * c, (cudaStream_t)stream);
  if (rc) return rc;
  if (light) {
    B2_CHECK_CUDA(launch_chain(bn_bwd_reduce_light_kernel, dim3(grid, slices), dim3(kBlock), (size_t)(0),
        (cudaStream_t)stream, 1, (long long)npix * c * 2, (const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)z,
        ldz, npix, c, lm, scale, shift, mean, invstd, relu, sums, det));
  } else {
    B2_CHECK_CUDA(launch_chain(bn_bwd_reduce_kernel, dim3(grid), dim3(kBlock), (size_t)(0), (cudaStream_t)stream, 1,
        (long long)npix * c * 2, (const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)z, ldz, npix, c, m, scale,
        shift, mean, invstd, relu, sums, det));
  }
  B2_LAUNCH_CHECK();
  if (det.partial) return det_finish(det.partial, grid, det.n, 2 * c, sums, (cudaStream_t)stream);
  return B2_OK;
}

extern "C" int b2_bn_bwd_apply(const void* dy, int32_t lddy, const void* z, int32_t ldz, int64_t npix, int32_t c,
                               const float* scale, const float* shift, const float* mean, const float* invstd,
                               const float* gamma, int32_t relu, int32_t training, const double* sums, void* dz,
                               int32_t lddz, float* dgamma, float* dbeta, float* dbias, b2_stream_t stream) {
  ChanMap m;
  int rc = make_map(c, &m);
  if (rc) return rc;
  B2_REQUIRE(aligned16(dy, lddy) && aligned16(z, ldz) && aligned16(dz, lddz), B2_ERR_ALIGN,
             "bn_bwd_apply operands misaligned");
  ChanMap lm;
  const bool light = light_map(c, &lm);
  int slices = 1;
  if (light && c % 64 == 0 && c > 64 && env_switch("B200SEG_BN_SLICE", 1) != 0) {
    slices = c / 64;                 // 64-channel slices: 16 threads per pixel, 16 pixels per block iteration
    lm.tpp = 16;
    lm.rows = kBlock / 16;
  }
  int grid = light ? chan_grid(npix, lm, 8) : chan_grid(npix, m, 4);
  if (slices > 1) {
    grid = (grid + slices - 1) / slices;
    if (grid < 1) grid = 1;
  }
  DetBuf det;
  det.partial = nullptr;
  det.n = c;
  if (dbias != nullptr) {
    rc = det_begin(&det, grid, c, (cudaStream_t)stream);
    if (rc) return rc;
  }
  if (light) {
    B2_CHECK_CUDA(launch_chain(bn_bwd_apply_light_kernel, dim3(grid, slices), dim3(kBlock), (size_t)(0),
        (cudaStream_t)stream, 1, (long long)npix * c * 2, (const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)z,
        ldz, npix, c, lm, scale, shift, mean, invstd, gamma, relu, training, sums, (__nv_bfloat16*)dz, lddz, dgamma,
        dbeta, dbias, det));
  } else {
    B2_CHECK_CUDA(launch_chain(bn_bwd_apply_kernel, dim3(grid), dim3(kBlock), (size_t)(0), (cudaStream_t)stream, 1,
        (long long)npix * c * 2, (const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)z, ldz, npix, c, m, scale,
        shift, mean, invstd, gamma, relu, training, sums, (__nv_bfloat16*)dz, lddz, dgamma, dbeta, dbias, det));
  }
  B2_LAUNCH_CHECK();
  if (det.partial) return det_finish(det.partial, grid, det.n, c, dbias, (cudaStream_t)stream);
  return B2_OK;
}
