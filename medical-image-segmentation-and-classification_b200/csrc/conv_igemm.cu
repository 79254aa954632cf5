// b200seg — implicit-GEMM convolution (stride 1, 'same' padding, 1x1 / 3x3) on tcgen05.
//
// GEMM view:  D[M = pixels, N = cout] = sum over (tap, channel block)  A_tap[pixels, 64 ch] * W_tap[cout, 64 ch]^T
//   A tiles : one TMA box (64 ch, Wb, Hb, Nb) of the NHWC activation per (tap, channel block); the box origin is
//             shifted by the tap offset and TMA's out-of-bounds zero fill supplies the padding halo.  The box
//             lands in smem as 128 rows of 128 B (K-major, 128B swizzle) == the canonical UMMA A layout.
//   B tiles : TMA box (64 ch, BLOCK_N, 1) of the packed weights [tap][cout][cin], same layout.
//   D       : 128 x BLOCK_N fp32 accumulator in TMEM (lane = pixel row, column = output channel).
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> +bias/+addend/ReLU -> bf16 -> global; BN sum / sum^2 of the rounded output).
// Two CTAs are co-resident per SM so one CTA's epilogue overlaps the other's main loop.
//
// Used for fprop (reference nn.Conv2d call sites, see include/b200seg.h) and for dgrad (flipped/transposed
// weight packing).  The K dimension may span two source tensors (elided torch.cat).
#include "common.cuh"

namespace b2 {

static constexpr int kTileM = 128;
static constexpr int kKBlock = 64;               // channels per K step (128 B of bf16)
static constexpr int kABytes = kTileM * 128;     // 16 KB
static constexpr int kThreads = 192;

struct IgemmParams {
  int H, W, N;
  int Wb, Hb, Nb;     // box extents; Wb*Hb*Nb == 128
  int tw, th;         // tiles along W and H
  int taps;           // 1 or 9
  int cb0, cb1;       // 64-channel blocks in source 0 / 1
  int block_n, stages, num_k_iters;
  int cout;
  __nv_bfloat16* y;
  int ldy;
  const float* bias;
  const __nv_bfloat16* addend;
  int ldadd;
  double* stats;
  int relu;
};

__global__ void __launch_bounds__(kThreads, 2)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB, const IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024 B alignment for the 128B swizzle atoms
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = kABytes + p.block_n * 128;
  uint8_t* tail = smem + p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full_bar = empty_bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_bias = reinterpret_cast<float*>(tail + 256);          // [block_n]
  float* s_sum = s_bias + 256;                                   // [block_n]
  float* s_sq = s_sum + 256;                                     // [block_n]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates: n-tile fastest so CTAs sharing an activation tile run together
  const int n_tiles = p.cout / p.block_n;
  const int n_tile = blockIdx.x % n_tiles;
  int m_tile = blockIdx.x / n_tiles;
  const int tw_i = m_tile % p.tw;
  m_tile /= p.tw;
  const int th_i = m_tile % p.th;
  const int tn_i = m_tile / p.th;
  const int w0 = tw_i * p.Wb, h0 = th_i * p.Hb, n0 = tn_i * p.Nb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB);
    if (p.cb1 > 0) tma_prefetch_desc(&tmA1);
  }
  const uint32_t tmem_cols = p.block_n < 32 ? 32u : (uint32_t)p.block_n;
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      const int cbt = p.cb0 + p.cb1;
      int stage = 0;
      uint32_t phase = 0;
      int tap = 0, cb = 0;
      for (int it = 0; it < p.num_k_iters; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * stage_bytes;
        uint8_t* sb = sa + kABytes;
        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
        int dr = 0, ds = 0;
        if (p.taps == 9) {
          dr = tap / 3 - 1;
          ds = tap % 3 - 1;
        }
        if (cb < p.cb0) {
          tma_load_4d(sa, &tmA0, &full_bar[stage], cb * kKBlock, w0 + ds, h0 + dr, n0);
        } else {
          tma_load_4d(sa, &tmA1, &full_bar[stage], (cb - p.cb0) * kKBlock, w0 + ds, h0 + dr, n0);
        }
        tma_load_3d(sb, &tmB, &full_bar[stage], cb * kKBlock, n_tile * p.block_n, tap);
        if (++cb == cbt) {
          cb = 0;
          ++tap;
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    const uint32_t idesc = umma_idesc_bf16(kTileM, p.block_n, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < p.num_k_iters; ++it) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
        const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
        for (int k = 0; k < kKBlock / 16; ++k) {
          const uint64_t da = umma_desc_sw128(a_addr + k * 32, 16, 1024);
          const uint64_t db = umma_desc_sw128(b_addr + k * 32, 16, 1024);
          umma_bf16(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);   // frees this smem stage when the MMAs have read it
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (lane == 0) umma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    // ------------------------------ epilogue (warps 2..5) ------------------------------
    const int et = threadIdx.x - 64;  // 0..127
    const int ch_base = n_tile * p.block_n;
    for (int i = et; i < p.block_n; i += 128) {
      s_bias[i] = p.bias ? p.bias[ch_base + i] : 0.f;
      s_sum[i] = 0.f;
      s_sq[i] = 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");

    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;          // accumulator row == pixel within the tile
    const int wl = row % p.Wb;
    const int hl = (row / p.Wb) % p.Hb;
    const int nl = row / (p.Wb * p.Hb);
    const bool valid = (n0 + nl) < p.N;
    const long long pix = ((long long)(n0 + nl) * p.H + (h0 + hl)) * p.W + (w0 + wl);
    __nv_bfloat16* yrow = p.y + pix * p.ldy + ch_base;
    const __nv_bfloat16* arow = p.addend ? p.addend + pix * p.ldadd + ch_base : nullptr;

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();

    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int c = 0; c < p.block_n; c += 32) {
      float v[32];
      tmem_ld32(taddr + c, v);
      tmem_ld_wait();
      if (arow != nullptr && valid) {
        const uint4* ap = reinterpret_cast<const uint4*>(arow + c);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 u = __ldg(ap + q);
          v[q * 8 + 0] += bf16lo(u.x); v[q * 8 + 1] += bf16hi(u.x);
          v[q * 8 + 2] += bf16lo(u.y); v[q * 8 + 3] += bf16hi(u.y);
          v[q * 8 + 4] += bf16lo(u.z); v[q * 8 + 5] += bf16hi(u.z);
          v[q * 8 + 6] += bf16lo(u.w); v[q * 8 + 7] += bf16hi(u.w);
        }
      }
      uint32_t packed[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float a = v[j] + s_bias[c + j];
        float b = v[j + 1] + s_bias[c + j + 1];
        if (p.relu) {
          a = fmaxf(a, 0.f);
          b = fmaxf(b, 0.f);
        }
        packed[j >> 1] = pack_bf16x2(a, b);
      }
      if (valid) {
        uint4* yp = reinterpret_cast<uint4*>(yrow + c);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          yp[q] = make_uint4(packed[q * 4], packed[q * 4 + 1], packed[q * 4 + 2], packed[q * 4 + 3]);
      }
      if (p.stats != nullptr) {
        // statistics of the ROUNDED output (what BatchNorm sees under autocast)
        float sq[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = valid ? bf16lo(packed[j]) : 0.f;
          const float b = valid ? bf16hi(packed[j]) : 0.f;
          v[2 * j] = a;
          v[2 * j + 1] = b;
          sq[2 * j] = a * a;
          sq[2 * j + 1] = b * b;
        }
        const float cs = warp_transpose_sum32(v, lane);
        const float cq = warp_transpose_sum32(sq, lane);
        atomicAdd(&s_sum[c + lane], cs);
        atomicAdd(&s_sq[c + lane], cq);
      }
    }
    if (p.stats != nullptr) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = et; i < p.block_n; i += 128) {
        atomicAdd(&p.stats[ch_base + i], (double)s_sum[i]);
        atomicAdd(&p.stats[p.cout + ch_base + i], (double)s_sq[i]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// Tile geometry shared with the wgrad kernel: split `tile_pix` pixels into a (Wb, Hb, Nb) box.
int conv_tile_geometry(int n, int h, int w, int tile_pix, int* Wb, int* Hb, int* Nb, int* tw, int* th, int* tn) {
  if (w >= tile_pix) {
    B2_REQUIRE(w % tile_pix == 0, B2_ERR_SHAPE, "W=%d must be a multiple of %d", w, tile_pix);
    *Wb = tile_pix; *Hb = 1; *Nb = 1;
  } else {
    B2_REQUIRE(tile_pix % w == 0, B2_ERR_SHAPE, "W=%d must divide %d", w, tile_pix);
    *Wb = w;
    const int rows = tile_pix / w;
    if (h >= rows) {
      B2_REQUIRE(h % rows == 0, B2_ERR_SHAPE, "H=%d must be a multiple of %d (W=%d)", h, rows, w);
      *Hb = rows; *Nb = 1;
    } else {
      B2_REQUIRE(rows % h == 0, B2_ERR_SHAPE, "H=%d must divide %d (W=%d)", h, rows, w);
      *Hb = h; *Nb = rows / h;
    }
  }
  *tw = w / *Wb;
  *th = h / *Hb;
  *tn = (n + *Nb - 1) / *Nb;
  return B2_OK;
}

int encode_act_tmap(CUtensorMap* tm, const void* base, int c, int ld, int n, int h, int w, int Wb, int Hb, int Nb) {
  B2_REQUIRE(ld % 8 == 0 && c % 8 == 0, B2_ERR_ALIGN, "channel count %d / stride %d must be multiples of 8", c, ld);
  uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
  uint64_t str[4] = {2, (uint64_t)ld * 2, (uint64_t)ld * 2 * w, (uint64_t)ld * 2 * w * h};
  uint32_t box[4] = {64, (uint32_t)Wb, (uint32_t)Hb, (uint32_t)Nb};
  return encode_tmap_bf16(tm, base, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

static int pick_block_n(int cout) {
  if (cout % 128 == 0) return 128;
  if (cout % 64 == 0) return 64;
  if (cout % 32 == 0) return 32;
  return 0;
}

static int conv_igemm_launch(const b2_conv_args* a, cudaStream_t stream) {
  B2_REQUIRE(a != nullptr, B2_ERR_SHAPE, "null args");
  B2_REQUIRE(a->ksize == 1 || a->ksize == 3, B2_ERR_SHAPE, "ksize %d unsupported (1 or 3)", a->ksize);
  B2_REQUIRE(a->n > 0 && a->h > 0 && a->w > 0, B2_ERR_SHAPE, "bad extent n=%d h=%d w=%d", a->n, a->h, a->w);
  B2_REQUIRE(a->c0 > 0 && a->c1 >= 0, B2_ERR_SHAPE, "bad channel counts c0=%d c1=%d", a->c0, a->c1);
  B2_REQUIRE(a->c1 == 0 || a->c0 % 64 == 0, B2_ERR_SHAPE, "c0=%d must be a multiple of 64 when c1>0", a->c0);
  B2_REQUIRE(a->ktot >= a->c0 + a->c1 && a->ktot % 8 == 0, B2_ERR_SHAPE, "ktot=%d inconsistent", a->ktot);
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.block_n = pick_block_n(a->cout);
  B2_REQUIRE(p.block_n != 0, B2_ERR_SHAPE, "cout=%d must be a multiple of 32", a->cout);
  B2_REQUIRE(a->ldy % 8 == 0 && (reinterpret_cast<uintptr_t>(a->y) & 15) == 0, B2_ERR_ALIGN, "y misaligned");
  B2_REQUIRE(a->addend == nullptr || (a->ldadd % 8 == 0 && (reinterpret_cast<uintptr_t>(a->addend) & 15) == 0),
             B2_ERR_ALIGN, "addend misaligned");
  int tn;
  int rc = conv_tile_geometry(a->n, a->h, a->w, kTileM, &p.Wb, &p.Hb, &p.Nb, &p.tw, &p.th, &tn);
  if (rc) return rc;
  p.H = a->h; p.W = a->w; p.N = a->n;
  p.taps = a->ksize * a->ksize;
  p.cb0 = (a->c0 + 63) / 64;
  p.cb1 = (a->c1 + 63) / 64;
  p.num_k_iters = p.taps * (p.cb0 + p.cb1);
  p.cout = a->cout;
  p.y = static_cast<__nv_bfloat16*>(a->y);
  p.ldy = a->ldy;
  p.bias = a->bias;
  p.addend = static_cast<const __nv_bfloat16*>(a->addend);
  p.ldadd = a->ldadd;
  p.stats = a->stats;
  p.relu = a->relu;
  const int stage_bytes = kABytes + p.block_n * 128;
  int stages = (110 * 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages > p.num_k_iters) stages = p.num_k_iters;
  if (stages < 1) stages = 1;
  p.stages = stages;
  const int smem_bytes = stages * stage_bytes + 256 + 3 * 256 * 4 + 1024;

  CUtensorMap tmA0, tmA1, tmB;
  rc = encode_act_tmap(&tmA0, a->x0, a->c0, a->ldx0, a->n, a->h, a->w, p.Wb, p.Hb, p.Nb);
  if (rc) return rc;
  if (a->c1 > 0) {
    rc = encode_act_tmap(&tmA1, a->x1, a->c1, a->ldx1, a->n, a->h, a->w, p.Wb, p.Hb, p.Nb);
    if (rc) return rc;
  } else {
    tmA1 = tmA0;
  }
  {
    uint64_t dims[3] = {(uint64_t)(a->c0 + a->c1), (uint64_t)a->cout, (uint64_t)p.taps};
    uint64_t str[3] = {2, (uint64_t)a->ktot * 2, (uint64_t)a->w_tap_stride * 2};
    uint32_t box[3] = {64, (uint32_t)p.block_n, 1};
    rc = encode_tmap_bf16(&tmB, a->wpk, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 116 * 1024));
    attr_set = true;
  }
  const long long grid = (long long)p.tw * p.th * tn * (a->cout / p.block_n);
  B2_REQUIRE(grid > 0 && grid < (1ll << 31), B2_ERR_SHAPE, "grid too large");
  conv_igemm_kernel<<<(unsigned)grid, kThreads, smem_bytes, stream>>>(tmA0, tmA1, tmB, p);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

}  // namespace b2

extern "C" int b2_conv_fprop(const b2_conv_args* a, b2_stream_t stream) {
  int rc = b2_arch_check();
  if (rc) return rc;
  return b2::conv_igemm_launch(a, static_cast<cudaStream_t>(stream));
}

extern "C" int b2_conv_dgrad(const b2_conv_args* a, b2_stream_t stream) {
  int rc = b2_arch_check();
  if (rc) return rc;
  return b2::conv_igemm_launch(a, static_cast<cudaStream_t>(stream));
}
