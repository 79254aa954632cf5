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
  long long i0 = base;
  if ((reinterpret_cast<uintptr_t>(r.g) & 15) == 0 && (base & 3) == 0) {
    const long long end4 = base + ((end - base) & ~3ll);
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (long long i = base + 4ll * threadIdx.x; i < end4; i += 4ll * kOptThreads) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(r.g + i));
      s = fmaf(g.x, g.x, s);
      s1 = fmaf(g.y, g.y, s1);
      s2 = fmaf(g.z, g.z, s2);
      s3 = fmaf(g.w, g.w, s3);
    }
    s = (s + s1) + (s2 + s3);
    i0 = end4;
  }
  for (long long i = i0 + threadIdx.x; i < end; i += kOptThreads) {
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
  auto update = [&](float g, float& m, float& v, float& w) {
    g *= clip;
    m = beta1 * m + (1.f - beta1) * g;
    v = beta2 * v + (1.f - beta2) * g * g;
    const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
    w = w * decay - step_size * (m / denom);
  };
  long long i0 = base;
  // 16-byte path: four elements per thread and access (chunks start at multiples of 4 elements; the tensors' base
  // pointers are checked).  With scalar accesses the launch had ~18 KB in flight per SM and ran at 2.8 TB/s.
  if (((reinterpret_cast<uintptr_t>(r.p) | reinterpret_cast<uintptr_t>(r.g) | reinterpret_cast<uintptr_t>(r.m) |
        reinterpret_cast<uintptr_t>(r.v)) & 15) == 0 && (base & 3) == 0) {
    const long long end4 = base + ((end - base) & ~3ll);
    for (long long i = base + 4ll * threadIdx.x; i < end4; i += 4ll * kOptThreads) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(r.g + i));
      float4 m4 = *reinterpret_cast<const float4*>(r.m + i);
      float4 v4 = *reinterpret_cast<const float4*>(r.v + i);
      float4 w4 = *reinterpret_cast<const float4*>(r.p + i);
      update(g4.x, m4.x, v4.x, w4.x);
      update(g4.y, m4.y, v4.y, w4.y);
      update(g4.z, m4.z, v4.z, w4.z);
      update(g4.w, m4.w, v4.w, w4.w);
      *reinterpret_cast<float4*>(r.m + i) = m4;
      *reinterpret_cast<float4*>(r.v + i) = v4;
      *reinterpret_cast<float4*>(r.p + i) = w4;
    }
    i0 = end4;
  }
  for (long long i = i0 + threadIdx.x; i < end; i += kOptThreads) {
    float m = r.m[i], v = r.v[i], w = r.p[i];
    update(__ldg(r.g + i), m, v, w);
    r.m[i] = m;
    r.v[i] = v;
    r.p[i] = w;
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

static constexpr int kPackMaxRefs = 1024;

__device__ __forceinline__ void pack_item(const PackRef& r, int item, float (*tile)[33]);

__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const PackRef* __restrict__ refs, int nrefs,
                                                                 int total_items) {
  __shared__ float tile[32][33];
  __shared__ int starts[kPackMaxRefs];
  // The table of first work items sits in shared memory (one coalesced read): the binary search for an item's tensor
  // used to be ~7 dependent global loads per 4 KB tile, which is what the launch spent its time on (1 TB/s).
  __shared__ int which;
  for (int i = threadIdx.x; i < nrefs; i += 256) starts[i] = refs[i].item_start;
  __syncthreads();
  for (int gi = blockIdx.x; gi < total_items; gi += gridDim.x) {
    // ONE thread searches (the launch was issue-bound: 70 % of the issue slots, most of them 256 threads repeating
    // the search and the item decode for four elements each)
    if (threadIdx.x == 0) {
      int lo = 0, hi = nrefs - 1;    // last ref with item_start <= gi
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (starts[mid] <= gi) lo = mid; else hi = mid - 1;
      }
      which = lo;
    }
    __syncthreads();
    const int lo = which;
    const PackRef r = refs[lo];
    pack_item(r, gi - starts[lo], tile);
    __syncthreads();                 // the tile (and `which`) are reused by the next item
  }
}

__device__ __forceinline__ void pack_item(const PackRef& r, int item, float (*tile)[33]) {
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
  // One work item = one 32 (co) x 32 (ci) tile of ALL taps: a thread loads the k x k taps of its four elements in one
  // batch (up to 36 loads in flight), then walks the slabs (tap, or phase * 4 + tap for kind 1) out of registers.  With
  // one item per (tile, slab) the launch was issue-bound: ~650 instructions per thread and item, 450 of them the search,
  // the PackRef copy and the index algebra, for four elements.
  const int tci = (r.cin + 31) / 32;
  const int taps = r.ksize * r.ksize;             // <= 9 (checked on the host)
  const int ci_t = item % tci;
  const int co_t = item / tci;
  const int co0 = co_t * 32, ci0 = ci_t * 32;
  float wv[4][9];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int co = co0 + ty + 8 * q, ci = ci0 + tx;
    const bool in = co < r.cout && ci < r.cin;
    const float* __restrict__ base = r.w + co * r.s_co + ci * r.s_ci;
#pragma unroll
    for (int t = 0; t < 9; ++t)
      wv[q][t] = (in && t < taps) ? __ldg(base + (t / r.ksize) * r.s_kh + (t % r.ksize) * r.s_kw) : 0.f;
  }
  const int nslab = r.kind == 0 ? taps : 16;
  for (int slab = 0; slab < nslab; ++slab) {
    float vals[4];
    if (r.kind == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v = wv[q][0];
#pragma unroll
        for (int t = 1; t < 9; ++t) v = t == slab ? wv[q][t] : v;      // register select (no dynamic indexing)
        vals[q] = v;
      }
    } else {
      const int phase = slab >> 2, tap = slab & 3;
      const int a = phase >> 1, b = phase & 1, u = tap >> 1, vv = tap & 1;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v = 0.f;       // same summation order as pack_weights_upfold_kernel: rows outer, columns inner
#pragma unroll
        for (int rw = 0; rw < 3; ++rw)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (upfold_member_(a, u, rw) && upfold_member_(b, vv, c)) v += wv[q][rw * 3 + c];
        vals[q] = v;
      }
    }
    if (slab > 0) __syncthreads();               // the previous slab's transposed reads of the tile are done
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int rr = ty + 8 * q;
      const int co = co0 + rr, ci = ci0 + tx;
      if (co < r.cout && ci < r.cin && r.wf != nullptr)
        r.wf[((long long)slab * r.cout + co) * r.cin + ci] = __float2bfloat16_rn(vals[q]);
      tile[rr][tx] = vals[q];
    }
    if (r.wd == nullptr) continue;
    __syncthreads();
    const int dslab = r.kind == 0 ? (taps - 1 - slab) : ((slab & ~3) + (3 - (slab & 3)));
    for (int rr = ty; rr < 32; rr += 8) {
      const int ci = ci0 + rr, co = co0 + tx;
      if (co < r.cout && ci < r.cin)
        r.wd[((long long)dslab * r.cin + ci) * r.cout + co] = __float2bfloat16_rn(tile[tx][rr]);
    }
  }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_pack_weights_multi(const b2_pack_ref* refs, int32_t nrefs, int32_t total_items,
                                     b2_stream_t stream) {
  static_assert(sizeof(PackRef) == sizeof(b2_pack_ref), "PackRef must mirror b2_pack_ref");
  B2_REQUIRE(nrefs > 0 && total_items > 0, B2_ERR_SHAPE, "empty pack launch");
  B2_REQUIRE(nrefs <= kPackMaxRefs, B2_ERR_SHAPE, "pack launch: %d tensors > %d", nrefs, kPackMaxRefs);
  const int grid = total_items < num_sms() * 32 ? total_items : num_sms() * 32;
  pack_weights_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const PackRef*>(refs), nrefs,
                                                                    total_items);
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
