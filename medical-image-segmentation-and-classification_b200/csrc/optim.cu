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

// ------------------------------------------------------------------------------------------------------------
// Multi-tensor weight re-pack: ONE launch refreshes the bf16 MMA-layout copies of every convolution weight after the
// optimizer step (SURVEY.md §8f N1 "... + bf16 weight re-pack"), instead of one pack launch per conv call in forward
// and again in backward.  A work item is a 32 x 32 (cout x cin) tile of one tap of one tensor, moved through shared
// memory so that the reads of w, the writes of the fprop layout [tap][cout][cin] and the writes of the transposed /
// flipped dgrad layout [taps-1-tap][cin][cout] are all row-contiguous.
//   kind 0  plain:   wf[tap][co][ci] = w[co][ci][kh][kw],  wd[taps-1-tap][ci][co]
//   kind 1  UpConv folding (AttentionUNet.py:15-27): 4 phases x 4 taps of summed 3x3 taps (pack_weights_upfold_kernel)
//   kind 2  image stem as a GEMM (3x3: K = 32, 7x7/s2: K = 152): wf[co][tap * cin + c], zero padded to pad_ columns
// ------------------------------------------------------------------------------------------------------------
struct PackRef {       // mirrors b2_pack_ref
  const float* w;
  __nv_bfloat16* wf;
  __nv_bfloat16* wd;
  int cout, cin, ksize, kind;
  long long s_co, s_ci, s_kh, s_kw;
  int item_start, pad_;
};

__device__ __forceinline__ bool upfold_member_(int a, int u, int r) {
  return a == 0 ? (u == 0 ? r == 0 : r >= 1) : (u == 0 ? r <= 1 : r == 2);
}

__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const PackRef* __restrict__ refs, int nrefs) {
  __shared__ float tile[32][33];
  // find the tensor of this work item: last ref with item_start <= blockIdx.x
  int lo = 0, hi = nrefs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (refs[mid].item_start <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const PackRef r = refs[lo];
  int item = (int)blockIdx.x - r.item_start;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tco = (r.cout + 31) / 32;
  if (r.kind == 2) {                                    // stem matrix [cout][cols], cols = r.pad_ (a multiple of 8)
    const int cols = r.pad_;
    const int ctiles = (cols + 31) / 32;
    const int co0 = (item / ctiles) * 32, col = (item % ctiles) * 32 + tx;
    const int ncol = r.ksize * r.ksize * r.cin;
    if (col >= cols) return;
    for (int rr = ty; rr < 32; rr += 8) {
      const int co = co0 + rr;
      if (co >= r.cout) continue;
      float v = 0.f;
      if (col < ncol) {
        const int tap = col / r.cin, c = col % r.cin;
        v = r.w[co * r.s_co + c * r.s_ci + (tap / r.ksize) * r.s_kh + (tap % r.ksize) * r.s_kw];
      }
      r.wf[(long long)co * cols + col] = __float2bfloat16_rn(v);
    }
    return;
  }
  const int tci = (r.cin + 31) / 32;
  const int ci_t = item % tci;
  item /= tci;
  const int co_t = item % tco;
  const int slab = item / tco;                          // tap (kind 0) or phase * 4 + tap (kind 1)
  const int co0 = co_t * 32, ci0 = ci_t * 32;
  const int taps = r.ksize * r.ksize;
  for (int rr = ty; rr < 32; rr += 8) {
    const int co = co0 + rr, ci = ci0 + tx;
    float v = 0.f;
    if (co < r.cout && ci < r.cin) {
      const float* base = r.w + co * r.s_co + ci * r.s_ci;
      if (r.kind == 0) {
        v = base[(slab / r.ksize) * r.s_kh + (slab % r.ksize) * r.s_kw];
      } else {
        const int phase = slab >> 2, tap = slab & 3;
        const int a = phase >> 1, b = phase & 1, u = tap >> 1, vv = tap & 1;
        for (int rw = 0; rw < 3; ++rw)
          for (int c = 0; c < 3; ++c)
            if (upfold_member_(a, u, rw) && upfold_member_(b, vv, c)) v += base[rw * r.s_kh + c * r.s_kw];
      }
      if (r.wf != nullptr) r.wf[((long long)slab * r.cout + co) * r.cin + ci] = __float2bfloat16_rn(v);
    }
    tile[rr][tx] = v;
  }
  if (r.wd == nullptr) return;
  __syncthreads();
  const int dslab = r.kind == 0 ? (taps - 1 - slab) : ((slab & ~3) + (3 - (slab & 3)));
  for (int rr = ty; rr < 32; rr += 8) {
    const int ci = ci0 + rr, co = co0 + tx;
    if (co < r.cout && ci < r.cin)
      r.wd[((long long)dslab * r.cin + ci) * r.cout + co] = __float2bfloat16_rn(tile[tx][rr]);
  }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_pack_weights_multi(const b2_pack_ref* refs, int32_t nrefs, int32_t total_items,
                                     b2_stream_t stream) {
  static_assert(sizeof(PackRef) == sizeof(b2_pack_ref), "PackRef must mirror b2_pack_ref");
  B2_REQUIRE(nrefs > 0 && total_items > 0, B2_ERR_SHAPE, "empty pack launch");
  pack_weights_multi_kernel<<<total_items, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const PackRef*>(refs),
                                                                           nrefs);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

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
