// b200seg — fused multi-tensor optimizer step: global gradient L2 norm, clip_grad_norm_(max_norm)
// and AdamW in two launches over ALL parameters, replacing the per-step ATen foreach kernels of
//   torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0); optimizer.step()      (utils/helpers.py:333-335)
// with the reference's AdamW hyper-parameters (helpers.py:251: lr, weight_decay 5e-4, betas (0.9, 0.999), eps 1e-8).
// Work is described by a device-side table of tensor references and a block -> (tensor, chunk) map, so the launch is
// independent of the number of tensors and can be captured in a CUDA graph.
#include "common.cuh"

namespace b2 {

struct TensorRef {     // mirrors b2_tensor_ref
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
};

static constexpr int kOptThreads = 256;

// sqnorm += sum g^2 ; block 0 also advances the step counter used for the bias correction
__global__ void __launch_bounds__(kOptThreads) grad_sqnorm_kernel(const TensorRef* __restrict__ refs,
                                                                  const int* __restrict__ block_tensor,
                                                                  const int* __restrict__ block_chunk,
                                                                  int chunk_elems, double* __restrict__ sqnorm,
                                                                  float* __restrict__ step, DetBuf det) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && step != nullptr) *step += 1.f;
  const TensorRef r = refs[block_tensor[blockIdx.x]];
  const long long base = (long long)block_chunk[blockIdx.x] * chunk_elems;
  long long end = base + chunk_elems;
  if (end > r.n) end = r.n;
  float s = 0.f;
  for (long long i = base + threadIdx.x; i < end; i += kOptThreads) {
    const float g = __ldg(r.g + i);
    s = fmaf(g, g, s);
  }
  __shared__ float sh[kOptThreads / 32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < kOptThreads / 32; ++i) t += sh[i];
    red_out(sqnorm, det, 0, (double)t);
  }
}

// clip coefficient = min(1, max_norm / (||g|| + 1e-6))  (torch.nn.utils.clip_grad_norm_), then decoupled-decay AdamW
__global__ void __launch_bounds__(kOptThreads) adamw_kernel(const TensorRef* __restrict__ refs,
                                                            const int* __restrict__ block_tensor,
                                                            const int* __restrict__ block_chunk, int chunk_elems,
                                                            const double* __restrict__ sqnorm, float max_norm,
                                                            const float* __restrict__ lr_ptr, float beta1, float beta2,
                                                            float eps, float weight_decay,
                                                            const float* __restrict__ step, float* total_norm_out) {
  const TensorRef r = refs[block_tensor[blockIdx.x]];
  const long long base = (long long)block_chunk[blockIdx.x] * chunk_elems;
  long long end = base + chunk_elems;
  if (end > r.n) end = r.n;
  const float total_norm = (float)sqrt(*sqnorm);
  if (blockIdx.x == 0 && threadIdx.x == 0 && total_norm_out != nullptr) *total_norm_out = total_norm;
  float clip = max_norm > 0.f ? max_norm / (total_norm + 1e-6f) : 1.f;
  if (clip > 1.f) clip = 1.f;
  const float lr = __ldg(lr_ptr);
  const float t = __ldg(step);
  const float bc1 = 1.f - powf(beta1, t);
  const float bc2 = 1.f - powf(beta2, t);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float decay = 1.f - lr * weight_decay;
  for (long long i = base + threadIdx.x; i < end; i += kOptThreads) {
    const float g = __ldg(r.g + i) * clip;
    const float m = beta1 * r.m[i] + (1.f - beta1) * g;
    const float v = beta2 * r.v[i] + (1.f - beta2) * g * g;
    r.m[i] = m;
    r.v[i] = v;
    const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
    r.p[i] = r.p[i] * decay - step_size * (m / denom);
  }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_grad_sqnorm_multi(const b2_tensor_ref* refs, const int32_t* block_tensor, const int32_t* block_chunk,
                                    int32_t nblocks, int32_t chunk_elems, double* sqnorm, float* step,
                                    b2_stream_t stream) {
  B2_REQUIRE(nblocks > 0 && chunk_elems > 0, B2_ERR_SHAPE, "empty optimizer launch");
  B2_CHECK_CUDA(cudaMemsetAsync(sqnorm, 0, sizeof(double), (cudaStream_t)stream));
  DetBuf det;
  int rc = det_begin(&det, nblocks, 1, (cudaStream_t)stream);
  if (rc) return rc;
  grad_sqnorm_kernel<<<nblocks, kOptThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const TensorRef*>(refs), block_tensor, block_chunk, chunk_elems, sqnorm, step, det);
  B2_LAUNCH_CHECK();
  if (det.partial) return det_finish(det.partial, nblocks, 1, 1, sqnorm, (cudaStream_t)stream);
  return B2_OK;
}

extern "C" int b2_adamw_multi(const b2_tensor_ref* refs, const int32_t* block_tensor, const int32_t* block_chunk,
                              int32_t nblocks, int32_t chunk_elems, const double* sqnorm, float max_norm,
                              const float* lr, float beta1, float beta2, float eps, float weight_decay,
                              const float* step, float* total_norm_out, b2_stream_t stream) {
  B2_REQUIRE(nblocks > 0 && chunk_elems > 0, B2_ERR_SHAPE, "empty optimizer launch");
  adamw_kernel<<<nblocks, kOptThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const TensorRef*>(refs), block_tensor, block_chunk, chunk_elems, sqnorm, max_norm, lr, beta1,
      beta2, eps, weight_decay, step, total_norm_out);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
