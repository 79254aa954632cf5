// b200seg — segmentation losses on fp32 logits.
//   BCEWithLogits, mean reduction          : nn.BCEWithLogitsLoss(), utils/helpers.py:244-246,327
//   0.5*BCE + 0.5*Dice (global, smooth=1)  : CombinedLoss / DiceLoss, utils/clip_seg_finetuner.py:40-74
// One reduction kernel (warp-shuffle + fp64 atomics) yields every sum either loss needs plus the IoU counts of
// utils/helpers.py:223-227; one elementwise kernel yields d loss / d logits.
#include "common.cuh"

namespace b2 {

__device__ __forceinline__ float stable_sigmoid(float z) {
  const float e = expf(-fabsf(z));
  return z >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
}

__global__ void __launch_bounds__(256) loss_fwd_kernel(const float* __restrict__ z, const float* __restrict__ t,
                                                       long long count, double* __restrict__ sums, DetBuf det) {
  float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const float zi = __ldg(z + i), ti = __ldg(t + i);
    a[0] += fmaxf(zi, 0.f) - zi * ti + log1pf(expf(-fabsf(zi)));
    const float s = stable_sigmoid(zi);
    a[1] += s * ti;
    a[2] += s;
    a[3] += ti;
    // iou() of utils/helpers.py:223-227 on the RAW mask: inter = sum(p * mask), union = #((p + mask) > 0), p in {0,1}
    const float pf = zi > 0.f ? 1.f : 0.f;
    a[4] += pf * ti;
    a[5] += (pf + ti > 0.f) ? 1.f : 0.f;
  }
  __shared__ float sh[6][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    a[k] = warp_sum(a[k]);
    if (lane == 0) sh[k][warp] = a[k];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[threadIdx.x][w];
    red_out(sums, det, (int)threadIdx.x, (double)s);
  }
}

__global__ void loss_finalize_kernel(const double* __restrict__ sums, long long count, float w_bce, float w_dice,
                                     float smooth, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double l = 0.0;
    if (w_bce != 0.f) l += (double)w_bce * sums[0] / (double)count;
    if (w_dice != 0.f) {
      const double dice = (2.0 * sums[1] + smooth) / (sums[2] + sums[3] + smooth);
      l += (double)w_dice * (1.0 - dice);
    }
    *loss = (float)l;
  }
}

__global__ void __launch_bounds__(256) loss_bwd_kernel(const float* __restrict__ z, const float* __restrict__ t,
                                                       long long count, const double* __restrict__ sums,
                                                       float w_bce, float w_dice, float smooth,
                                                       const float* __restrict__ grad_out,
                                                       float* __restrict__ dz) {
  const float go = grad_out ? __ldg(grad_out) : 1.f;
  const float kb = w_bce / (float)count;
  float num = 0.f, den = 1.f;
  if (w_dice != 0.f) {
    num = (float)(2.0 * sums[1] + smooth);
    den = (float)(sums[2] + sums[3] + smooth);
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const float zi = __ldg(z + i), ti = __ldg(t + i);
    const float s = stable_sigmoid(zi);
    float g = kb * (s - ti);
    if (w_dice != 0.f) {
      // d/dz [1 - num/den] = -(2 t den - num) s(1-s) / den^2
      g -= w_dice * (2.f * ti * den - num) * s * (1.f - s) / (den * den);
    }
    dz[i] = go * g;
  }
}

// Per-sample confusion counts for the validation / test metrics (utils/tester.py:92-193, utils/helpers.py:223-227):
// pred = z > thr_logit  (== sigmoid(z) > threshold), target = t > thr_target; counts[n] = {TP, #pred, #target}.
// Everything the reference derives per sample (IoU, Dice, pixel accuracy, precision, recall, F1) is a function of
// these three integers and H*W, so one pass over the logits replaces ~10 reductions and 8 host syncs per sample.
__global__ void __launch_bounds__(256) seg_counts_kernel(const float* __restrict__ z, const float* __restrict__ t,
                                                         long long per_sample, float thr_logit, float thr_target,
                                                         unsigned long long* __restrict__ counts) {
  const int n = blockIdx.y;
  const float* zn = z + (long long)n * per_sample;
  const float* tn = t + (long long)n * per_sample;
  unsigned tp = 0, np = 0, nt = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample;
       i += (long long)gridDim.x * blockDim.x) {
    const bool pr = __ldg(zn + i) > thr_logit, tg = __ldg(tn + i) > thr_target;
    tp += (pr && tg) ? 1u : 0u;
    np += pr ? 1u : 0u;
    nt += tg ? 1u : 0u;
  }
  tp = __reduce_add_sync(0xffffffffu, tp);
  np = __reduce_add_sync(0xffffffffu, np);
  nt = __reduce_add_sync(0xffffffffu, nt);
  __shared__ unsigned sh[3][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh[0][warp] = tp;
    sh[1][warp] = np;
    sh[2][warp] = nt;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    unsigned long long s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[threadIdx.x][w];
    if (s) atomicAdd(&counts[n * 3 + threadIdx.x], s);
  }
}

// binary mask image of utils/pipeline.py:352-354: (sigmoid(z) > threshold) * 255 as uint8
__global__ void __launch_bounds__(256) logits_to_mask_kernel(const float* __restrict__ z, long long count,
                                                             float thr_logit, uint8_t* __restrict__ mask) {
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < count) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(z + i4));
    uchar4 m;
    m.x = v.x > thr_logit ? 255 : 0;
    m.y = v.y > thr_logit ? 255 : 0;
    m.z = v.z > thr_logit ? 255 : 0;
    m.w = v.w > thr_logit ? 255 : 0;
    *reinterpret_cast<uchar4*>(mask + i4) = m;
  } else {
    for (long long i = i4; i < count; ++i) mask[i] = __ldg(z + i) > thr_logit ? 255 : 0;
  }
}

static int l_grid(long long count) {
  long long g = (count + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace b2

using namespace b2;

extern "C" int b2_loss_fwd(const float* z, const float* t, int64_t count, double* sums, b2_stream_t stream) {
  B2_REQUIRE(count > 0, B2_ERR_SHAPE, "empty loss input");
  const int grid = l_grid(count);
  DetBuf det;
  int rc = det_begin(&det, grid, 6, (cudaStream_t)stream);
  if (rc) return rc;
  loss_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, t, count, sums, det);
  B2_LAUNCH_CHECK();
  if (det.partial) return det_finish(det.partial, grid, 6, 6, sums, (cudaStream_t)stream);
  return B2_OK;
}

extern "C" int b2_loss_finalize(const double* sums, int64_t count, float w_bce, float w_dice, float smooth,
                                float* loss, b2_stream_t stream) {
  B2_REQUIRE(count > 0, B2_ERR_SHAPE, "empty loss input");
  loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, count, w_bce, w_dice, smooth, loss);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_loss_bwd(const float* z, const float* t, int64_t count, const double* sums, float w_bce,
                           float w_dice, float smooth, const float* grad_out, float* dz, b2_stream_t stream) {
  B2_REQUIRE(count > 0, B2_ERR_SHAPE, "empty loss input");
  loss_bwd_kernel<<<l_grid(count), 256, 0, (cudaStream_t)stream>>>(z, t, count, sums, w_bce, w_dice, smooth,
                                                                   grad_out, dz);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_seg_counts(const float* z, const float* t, int32_t n, int64_t per_sample, float thr_logit,
                             float thr_target, uint64_t* counts, b2_stream_t stream) {
  B2_REQUIRE(n > 0 && n <= 65535 && per_sample > 0, B2_ERR_SHAPE, "bad seg_counts extent n=%d per_sample=%lld", n,
             (long long)per_sample);
  B2_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)n * 3 * sizeof(uint64_t), (cudaStream_t)stream));
  long long gx = (per_sample + 256 * 8 - 1) / (256 * 8);
  const long long cap = ((long long)num_sms() * 8 + n - 1) / n;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  seg_counts_kernel<<<dim3((unsigned)gx, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(
      z, t, per_sample, thr_logit, thr_target, reinterpret_cast<unsigned long long*>(counts));
  B2_LAUNCH_CHECK();
  return B2_OK;
}

extern "C" int b2_logits_to_mask(const float* z, int64_t count, float thr_logit, uint8_t* mask, b2_stream_t stream) {
  B2_REQUIRE(count > 0, B2_ERR_SHAPE, "empty mask input");
  B2_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(mask) & 3) == 0, B2_ERR_ALIGN,
             "logits must be 16 B aligned, mask 4 B aligned");
  const long long n4 = (count + 3) / 4;
  logits_to_mask_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(z, count, thr_logit, mask);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
