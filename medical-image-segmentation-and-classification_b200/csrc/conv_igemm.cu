// b200seg — implicit-GEMM convolution (stride 1, 'same' padding, 1x1 / 3x3) on tcgen05: persistent, warp-specialised.
//
// GEMM view:  D[M = pixels, N = cout] = sum over (tap, channel block)  A_tap[pixels, 64 ch] * W_tap[cout, 64 ch]^T
//   A tiles : one TMA box (64 ch, Wb, Hb, Nb) of the NHWC activation per (tap, channel block); the box origin is
//             shifted by the tap offset and TMA's out-of-bounds zero fill supplies the padding halo.  The box
//             lands in smem as 128 rows of 128 B (K-major, 128B swizzle) == the canonical UMMA A layout.
//             "halo" mode (tiles that are a single image-row segment, W >= 128): ONE box of Wb+2 pixels per
//             (filter row, channel block); the three horizontal taps are three MMAs whose A descriptors start
//             0 / 128 / 256 B into that box, so the activation is fetched 3x instead of 9x.
//   B tiles : TMA box (64 ch, BLOCK_N, 1 or 3 taps) of the packed weights [tap][cout][cin], same layout.
//   D       : 128 x BLOCK_N fp32 accumulators in TMEM, double buffered (2 x BLOCK_N columns).
// One CTA per SM loops over its tiles (fixed n-tile, strided m-tiles):
//   warp 0    TMA producer (smem ring of `stages` A+B slots, full/empty mbarriers)
//   warp 1    TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit frees smem slots and publishes
//             a finished accumulator (tmem_full); waits on tmem_empty before overwriting a buffer
//   warps 2-5 epilogue: tcgen05.ld -> +bias (+addend) (ReLU) -> bf16 -> swizzled smem tile -> TMA store;
//             BatchNorm sum / sum-of-squares of the ROUNDED output are column sums of that smem tile, kept in
//             fp64 registers across all tiles of the CTA and flushed with one fp64 atomic per channel at the end.
// The epilogue of tile i overlaps the main loop of tile i+1.
//
// Used for fprop (reference nn.Conv2d call sites, see include/b200seg.h) and for dgrad (flipped/transposed
// weight packing).  The K dimension may span two source tensors (elided torch.cat).
#include <stdlib.h>

#include "common.cuh"

namespace b2 {

static constexpr int kTileM = 128;
static constexpr int kKBlock = 64;               // channels per K step (128 B of bf16)
static constexpr int kThreads = 192;
static constexpr int kMaxStages = 12;

struct IgemmParams {
  int H, W, N;
  int Wb, Hb, Nb;     // box extents; Wb*Hb*Nb == 128
  int tw, th;         // tiles along W and H
  int taps;           // 1 or 9
  int cb0, cb1;       // 64-channel blocks in source 0 / 1
  int block_n, stages, num_k_iters;
  int cout, n_tiles, m_tiles, m_stride;
  int halo, base_off_mode;
  int stride, ksize, pad_h, pad_w;
  long long y_sn, y_sh, y_sw;   // element strides of the output pixel grid (strided placement for ConvT)
  int a_stage_bytes, b_stage_bytes, a_tx_bytes;
  int tma_store;
  __nv_bfloat16* y;
  int ldy;
  const float* bias;
  const __nv_bfloat16* addend;
  int ldadd;
  double* stats;
  int relu;
  int add_after_act;
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, const void* smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint64_t umma_desc_sw128_bo(uint32_t smem_addr, uint32_t lbo, uint32_t sbo, uint32_t bo) {
  return umma_desc_sw128(smem_addr, lbo, sbo) | ((uint64_t)(bo & 7u) << 49);
}

// byte offset of (row, channel) inside the staged output tile (swizzled so that both the row-wise 16 B stores of
// the TMEM drain and the column-wise reads of the statistics pass are bank-conflict free; for block_n >= 64 it is
// exactly the TMA SWIZZLE_128B layout of 64-channel panels)
__device__ __forceinline__ uint32_t ctile_off(int block_n, int row, int ch) {
  if (block_n >= 64) {
    const int panel = ch >> 6, cc = ch & 63;
    return (uint32_t)(panel * (kTileM * 128) + row * 128 + ((((cc >> 3) ^ (row & 7))) << 4) + ((cc & 7) << 1));
  }
  return (uint32_t)(row * 64 + ((((ch >> 3) ^ ((row >> 1) & 3))) << 4) + ((ch & 7) << 1));
}

__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY,
                  const IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
  uint8_t* ctile = smem + p.stages * stage_bytes;                      // 128 x block_n bf16, 1024 B aligned
  uint8_t* tail = ctile + kTileM * p.block_n * 2;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;                    // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(tail + 256);                // [block_n]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // tiles are numbered n-fastest (CTAs that run together share an activation tile); CTA b takes b, b+G, b+2G, ...
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 4);      // one arrival per epilogue warp
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB);
    if (p.cb1 > 0) tma_prefetch_desc(&tmA1);
    if (p.tma_store) tma_prefetch_desc(&tmY);
  }
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * p.block_n) tmem_cols <<= 1;
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int cbt = p.cb0 + p.cb1;

  if (warp == 0) {
    // ------------------------------ TMA producer (warp-uniform control flow) ------------------------------
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx = (uint32_t)(p.a_tx_bytes + p.b_stage_bytes);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      int t = tile / p.n_tiles;
      const int tw_i = t % p.tw;
      t /= p.tw;
      const int th_i = t % p.th;
      const int tn_i = t / p.th;
      const int w0 = tw_i * p.Wb, h0 = th_i * p.Hb, n0 = tn_i * p.Nb;
      const int outer = p.halo ? 3 : p.taps;     // halo mode: one iteration per (filter row, channel block)
      for (int o = 0; o < outer; ++o) {
        int dr = 0, ds = 0, tap0 = o;
        if (p.halo) {
          dr = o - 1;
          ds = -1;
          tap0 = o * 3;
        } else {
          dr = o / p.ksize - p.pad_h;
          ds = o % p.ksize - p.pad_w;
        }
        for (int cb = 0; cb < cbt; ++cb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem + stage * stage_bytes;
            uint8_t* sb = sa + p.a_stage_bytes;
            mbar_arrive_expect_tx(&full_bar[stage], tx);
            if (cb < p.cb0) {
              tma_load_4d(sa, &tmA0, &full_bar[stage], cb * kKBlock, p.stride * w0 + ds, p.stride * h0 + dr, n0);
            } else {
              tma_load_4d(sa, &tmA1, &full_bar[stage], (cb - p.cb0) * kKBlock, p.stride * w0 + ds,
                          p.stride * h0 + dr, n0);
            }
            tma_load_3d(sb, &tmB, &full_bar[stage], cb * kKBlock, n_tile * p.block_n, tap0);
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    const uint32_t idesc = umma_idesc_bf16(kTileM, p.block_n, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int ti = 0;
    const int sub = p.halo ? 3 : 1;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const int buf = ti & 1;
      mbar_wait(&tmem_empty_bar[buf], (((uint32_t)ti >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.block_n);
      for (int it = 0; it < p.num_k_iters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint32_t b_addr = a_addr + p.a_stage_bytes;
          for (int s = 0; s < sub; ++s) {
            const uint32_t a_s = a_addr + s * 128;            // halo mode: shift by s pixels (rows of 128 B)
            const uint32_t b_s = b_addr + s * p.block_n * 128;
            const uint32_t bo = p.base_off_mode ? ((a_s >> 7) & 7u) : 0u;
#pragma unroll
            for (int k = 0; k < kKBlock / 16; ++k) {
              const uint64_t da = umma_desc_sw128_bo(a_s + k * 32, 16, 1024, bo);
              const uint64_t db = umma_desc_sw128(b_s + k * 32, 16, 1024);
              umma_bf16(d_tmem, da, db, idesc, (it | s | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);   // frees this smem slot when the MMAs have read it
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(&tmem_full_bar[buf]);
      __syncwarp();
    }
  } else {
    // ------------------------------ epilogue (warps 2..5) ------------------------------
    const int et = threadIdx.x - 64;  // 0..127
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;          // accumulator row == pixel within the tile
    const int wl = row % p.Wb;
    const int hl = (row / p.Wb) % p.Hb;
    const int nl = row / (p.Wb * p.Hb);
    // statistics mapping: thread owns one column pair for a slab of rows
    const int npairs = p.block_n >> 1;
    const int groups = 128 / npairs > 0 ? 128 / npairs : 1;
    const int rows_per_group = kTileM / groups;
    const int pair = et % npairs;
    const int grp = et / npairs;               // < groups when block_n <= 256
    double acc_s0 = 0.0, acc_s1 = 0.0, acc_q0 = 0.0, acc_q1 = 0.0;
    const int chunks_per_row = p.block_n >> 3;

    int ti = 0;
    int cur_n_tile = -1;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const int n_tile = tile % p.n_tiles;
      const int ch_base = n_tile * p.block_n;
      if (n_tile != cur_n_tile) {
        // new output-channel slab: flush the statistics kept for the previous one, reload the bias slice
        if (cur_n_tile >= 0 && p.stats != nullptr && grp < groups) {
          const int ch = cur_n_tile * p.block_n + pair * 2;
          atomicAdd(&p.stats[ch], acc_s0);
          atomicAdd(&p.stats[ch + 1], acc_s1);
          atomicAdd(&p.stats[p.cout + ch], acc_q0);
          atomicAdd(&p.stats[p.cout + ch + 1], acc_q1);
          acc_s0 = acc_s1 = acc_q0 = acc_q1 = 0.0;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");     // everyone is done reading the old bias slice
        for (int i = et; i < p.block_n; i += 128) s_bias[i] = p.bias ? p.bias[ch_base + i] : 0.f;
        cur_n_tile = n_tile;
      }
      int t = tile / p.n_tiles;
      const int tw_i = t % p.tw;
      t /= p.tw;
      const int th_i = t % p.th;
      const int tn_i = t / p.th;
      const int w0 = tw_i * p.Wb, h0 = th_i * p.Hb, n0 = tn_i * p.Nb;
      int valid_rows = (p.N - n0) * p.Wb * p.Hb;
      if (valid_rows > kTileM) valid_rows = kTileM;
      const int buf = ti & 1;

      // the previous tile's TMA store must have finished READING the staging tile before it is overwritten
      if (p.tma_store && et == 0) tma_store_wait_read();
      asm volatile("bar.sync 1, 128;" ::: "memory");

      mbar_wait(&tmem_full_bar[buf], ((uint32_t)ti >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * p.block_n);
      const bool valid = row < valid_rows;
      const long long pix = ((long long)(n0 + nl) * p.H + (h0 + hl)) * p.W + (w0 + wl);
      const __nv_bfloat16* arow = (p.addend && valid) ? p.addend + pix * p.ldadd + ch_base : nullptr;
      for (int c = 0; c < p.block_n; c += 32) {
        float v[32];
        tmem_ld32(taddr + c, v);
        tmem_ld_wait();
        float av[32];
        const bool has_add = arow != nullptr;
        if (has_add) {
          const uint4* ap = reinterpret_cast<const uint4*>(arow + c);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 u = __ldg(ap + q);
            av[q * 8 + 0] = bf16lo(u.x); av[q * 8 + 1] = bf16hi(u.x);
            av[q * 8 + 2] = bf16lo(u.y); av[q * 8 + 3] = bf16hi(u.y);
            av[q * 8 + 4] = bf16lo(u.z); av[q * 8 + 5] = bf16hi(u.z);
            av[q * 8 + 6] = bf16lo(u.w); av[q * 8 + 7] = bf16hi(u.w);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) av[j] = 0.f;
        }
        const bool add_first = has_add && !p.add_after_act;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e = q * 8 + 2 * j;
            float a = v[e] + s_bias[c + e] + (add_first ? av[e] : 0.f);
            float b = v[e + 1] + s_bias[c + e + 1] + (add_first ? av[e + 1] : 0.f);
            if (p.relu) {
              a = fmaxf(a, 0.f);
              b = fmaxf(b, 0.f);
            }
            if (p.add_after_act) {      // x + relu(conv(.)): the addend joins after the activation (on rounded values)
              a = bf16_round(a) + av[e];
              b = bf16_round(b) + av[e + 1];
            }
            pk[j] = pack_bf16x2(a, b);
          }
          *reinterpret_cast<uint4*>(ctile + ctile_off(p.block_n, row, c + q * 8)) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
      fence_proxy_async();                       // generic-proxy smem writes -> visible to the TMA store
      asm volatile("bar.sync 1, 128;" ::: "memory");

      if (p.tma_store) {
        if (et == 0) {
          for (int pn = 0; pn < (p.block_n >> 6); ++pn)
            tma_store_4d(&tmY, ctile + pn * (kTileM * 128), ch_base + pn * 64, w0, h0, n0);
          tma_store_commit();
        }
      } else {
        // coalesced copy-out: consecutive threads write consecutive 16 B of consecutive pixels
        for (int idx = et; idx < kTileM * chunks_per_row; idx += 128) {
          const int r = idx / chunks_per_row, ck = idx % chunks_per_row;
          if (r < valid_rows) {
            const int rwl = r % p.Wb, rhl = (r / p.Wb) % p.Hb, rnl = r / (p.Wb * p.Hb);
            const long long ro = (n0 + rnl) * p.y_sn + (h0 + rhl) * p.y_sh + (w0 + rwl) * p.y_sw;
            const uint4 u = *reinterpret_cast<const uint4*>(ctile + ctile_off(p.block_n, r, ck * 8));
            *reinterpret_cast<uint4*>(p.y + ro + ch_base + ck * 8) = u;
          }
        }
      }
      if (p.stats != nullptr && grp < groups) {
        // column sums of the ROUNDED outputs (what BatchNorm sees under autocast)
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        const int r_begin = grp * rows_per_group;
        int r_end = r_begin + rows_per_group;
        if (r_end > valid_rows) r_end = valid_rows;
        for (int r = r_begin; r < r_end; ++r) {
          const uint32_t u = *reinterpret_cast<const uint32_t*>(ctile + ctile_off(p.block_n, r, pair * 2));
          const float a = bf16lo(u), b = bf16hi(u);
          s0 += a;
          s1 += b;
          q0 = fmaf(a, a, q0);
          q1 = fmaf(b, b, q1);
        }
        acc_s0 += (double)s0;
        acc_s1 += (double)s1;
        acc_q0 += (double)q0;
        acc_q1 += (double)q1;
      }
    }
    if (p.stats != nullptr && grp < groups && cur_n_tile >= 0) {
      const int ch = cur_n_tile * p.block_n + pair * 2;
      atomicAdd(&p.stats[ch], acc_s0);
      atomicAdd(&p.stats[ch + 1], acc_s1);
      atomicAdd(&p.stats[p.cout + ch], acc_q0);
      atomicAdd(&p.stats[p.cout + ch + 1], acc_q1);
    }
    if (p.tma_store && et == 0) tma_store_wait_all();
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

// General NHWC view: element strides (s_w, s_h, s_n) of the pixel grid, optional traversal stride `es` in w and h
// (box extents are given in LOADED pixels).
int encode_act_tmap_ex(CUtensorMap* tm, const void* base, int c, int n, int h, int w, long long s_w, long long s_h,
                       long long s_n, int Wb, int Hb, int Nb, int es) {
  B2_REQUIRE(s_w % 8 == 0 && s_h % 8 == 0 && s_n % 8 == 0 && c % 8 == 0, B2_ERR_ALIGN,
             "channel count %d / pixel strides must be multiples of 8 elements", c);
  B2_REQUIRE(Wb * es <= 256 && Hb * es <= 256, B2_ERR_SHAPE, "TMA box too large for stride %d", es);
  uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
  uint64_t str[4] = {2, (uint64_t)s_w * 2, (uint64_t)s_h * 2, (uint64_t)s_n * 2};
  uint32_t box[4] = {64, (uint32_t)(Wb * es), (uint32_t)(Hb * es), (uint32_t)Nb};
  uint32_t est[4] = {1, (uint32_t)es, (uint32_t)es, 1};
  return encode_tmap_bf16(tm, base, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, est);
}

int encode_act_tmap(CUtensorMap* tm, const void* base, int c, int ld, int n, int h, int w, int Wb, int Hb, int Nb) {
  return encode_act_tmap_ex(tm, base, c, n, h, w, ld, (long long)ld * w, (long long)ld * w * h, Wb, Hb, Nb, 1);
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

static int pick_block_n(int cout, bool wide_rows) {
  const int forced = env_int("B200SEG_BLOCK_N", 0);
  // N = 256 tiles do not leave room for >= 3 pipeline stages in halo mode (3 taps of B per stage)
  if (forced > 0 && cout % forced == 0 && !(forced > 128 && wide_rows)) return forced;
  if (cout % 256 == 0 && !wide_rows) return 256;   // deep layers: halves the activation re-fetch, 96 B/clk smem feed
  if (cout % 128 == 0) return 128;
  if (cout % 64 == 0) return 64;
  if (cout % 32 == 0) return 32;
  return 0;
}

static int conv_igemm_launch(const b2_conv_args* a, cudaStream_t stream) {
  B2_REQUIRE(a != nullptr, B2_ERR_SHAPE, "null args");
  const int stride = a->stride == 0 ? 1 : a->stride;
  const int out_mul = a->out_mul == 0 ? 1 : a->out_mul;
  B2_REQUIRE(a->ksize >= 1 && a->ksize <= 3, B2_ERR_SHAPE, "ksize %d unsupported (1, 2 or 3)", a->ksize);
  B2_REQUIRE(stride == 1 || stride == 2, B2_ERR_SHAPE, "stride %d unsupported (1 or 2)", stride);
  const int in_mul = a->in_mul == 0 ? 1 : a->in_mul;
  B2_REQUIRE(a->ksize != 2 || stride == 2 || a->custom_pad != 0, B2_ERR_SHAPE,
             "ksize 2 needs stride 2 or explicit tap offsets");
  B2_REQUIRE(in_mul >= 1 && (in_mul == 1 || stride == 1) && a->in_off_h >= 0 && a->in_off_h < in_mul &&
                 a->in_off_w >= 0 && a->in_off_w < in_mul,
             B2_ERR_SHAPE, "bad input placement mul=%d off=(%d,%d)", in_mul, a->in_off_h, a->in_off_w);
  B2_REQUIRE(out_mul >= 1 && a->out_off_h >= 0 && a->out_off_h < out_mul && a->out_off_w >= 0 &&
                 a->out_off_w < out_mul,
             B2_ERR_SHAPE, "bad output placement mul=%d off=(%d,%d)", out_mul, a->out_off_h, a->out_off_w);
  B2_REQUIRE(a->n > 0 && a->h > 0 && a->w > 0, B2_ERR_SHAPE, "bad extent n=%d h=%d w=%d", a->n, a->h, a->w);
  B2_REQUIRE(a->c0 > 0 && a->c1 >= 0, B2_ERR_SHAPE, "bad channel counts c0=%d c1=%d", a->c0, a->c1);
  B2_REQUIRE(a->c1 == 0 || a->c0 % 64 == 0, B2_ERR_SHAPE, "c0=%d must be a multiple of 64 when c1>0", a->c0);
  B2_REQUIRE(a->ktot >= a->c0 + a->c1 && a->ktot % 8 == 0, B2_ERR_SHAPE, "ktot=%d inconsistent", a->ktot);
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.block_n = pick_block_n(a->cout, a->ksize == 3 && a->w >= kTileM);
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
  const int cbt = p.cb0 + p.cb1;
  p.cout = a->cout;
  p.n_tiles = a->cout / p.block_n;
  p.m_tiles = p.tw * p.th * tn;
  p.y = static_cast<__nv_bfloat16*>(a->y);
  p.ldy = a->ldy;
  p.bias = a->bias;
  p.addend = static_cast<const __nv_bfloat16*>(a->addend);
  p.ldadd = a->ldadd;
  p.stats = a->stats;
  p.relu = a->relu;
  p.add_after_act = (a->addend != nullptr && a->add_after_act) ? 1 : 0;
  // halo mode: 3x3, tile == one row segment of 128 pixels
  const int halo_env = env_int("B200SEG_HALO", 1);
  p.stride = stride;
  p.ksize = a->ksize;
  p.pad_h = a->custom_pad ? a->pad_h : (a->ksize == 3 ? 1 : 0);
  p.pad_w = a->custom_pad ? a->pad_w : (a->ksize == 3 ? 1 : 0);
  p.halo = (halo_env != 0 && p.taps == 9 && stride == 1 && !a->custom_pad && p.Hb == 1 && p.Nb == 1 &&
            p.Wb == kTileM) ? 1 : 0;
  p.base_off_mode = (halo_env == 2) ? 1 : 0;
  if (p.halo) {
    p.a_tx_bytes = (kTileM + 2) * 128;
    p.a_stage_bytes = ((p.a_tx_bytes + 1023) / 1024) * 1024;
    p.b_stage_bytes = 3 * p.block_n * 128;
    p.num_k_iters = 3 * cbt;
  } else {
    p.a_tx_bytes = kTileM * 128;
    p.a_stage_bytes = kTileM * 128;
    p.b_stage_bytes = p.block_n * 128;
    p.num_k_iters = p.taps * cbt;
  }
  p.tma_store = (p.block_n % 64 == 0 && env_int("B200SEG_TMA_STORE", 1) != 0) ? 1 : 0;
  const int stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
  const int ctile_bytes = kTileM * p.block_n * 2;
  const int tail_bytes = 256 + 256 * 4 + 256;
  const int budget = 232448 - 1024 - tail_bytes - ctile_bytes;
  int stages = budget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  B2_REQUIRE(stages >= 2, B2_ERR_SHAPE, "tile configuration does not fit shared memory");
  p.stages = stages;
  const int smem_bytes = stages * stage_bytes + ctile_bytes + tail_bytes + 1024;

  CUtensorMap tmA0, tmA1, tmB, tmY;
  const int boxw = p.halo ? p.Wb + 2 : p.Wb;
  const int ih = a->h * stride, iw = a->w * stride;      // extent of the sampled input grid
  {
    // underlying image is (ih*in_mul) x (iw*in_mul); the operand is its (in_off_h, in_off_w) sub-lattice
    const long long fw = (long long)iw * in_mul, fh = (long long)ih * in_mul;
    const long long off0 = ((long long)a->in_off_h * fw + a->in_off_w) * a->ldx0;
    rc = encode_act_tmap_ex(&tmA0, static_cast<const __nv_bfloat16*>(a->x0) + off0, a->c0, a->n, ih, iw,
                            (long long)in_mul * a->ldx0, (long long)in_mul * fw * a->ldx0, fh * fw * a->ldx0, boxw,
                            p.Hb, p.Nb, stride);
    if (rc) return rc;
  }
  if (a->c1 > 0) {
    B2_REQUIRE(in_mul == 1, B2_ERR_SHAPE, "input placement is not supported with two K sources");
    rc = encode_act_tmap_ex(&tmA1, a->x1, a->c1, a->n, ih, iw, a->ldx1, (long long)a->ldx1 * iw,
                            (long long)a->ldx1 * iw * ih, boxw, p.Hb, p.Nb, stride);
    if (rc) return rc;
  } else {
    tmA1 = tmA0;
  }
  {
    uint64_t dims[3] = {(uint64_t)(a->c0 + a->c1), (uint64_t)a->cout, (uint64_t)p.taps};
    uint64_t str[3] = {2, (uint64_t)a->ktot * 2, (uint64_t)a->w_tap_stride * 2};
    uint32_t box[3] = {64, (uint32_t)p.block_n, p.halo ? 3u : 1u};
    rc = encode_tmap_bf16(&tmB, a->wpk, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  // output pixel grid (possibly a strided sub-lattice of a larger image: ConvTranspose pixel shuffle)
  const long long ow = (long long)a->w * out_mul, oh = (long long)a->h * out_mul;
  p.y_sw = (long long)out_mul * a->ldy;
  p.y_sh = (long long)out_mul * ow * a->ldy;
  p.y_sn = oh * ow * a->ldy;
  p.y = static_cast<__nv_bfloat16*>(a->y) + ((long long)a->out_off_h * ow + a->out_off_w) * a->ldy;
  if (p.tma_store) {
    rc = encode_act_tmap_ex(&tmY, p.y, a->cout, a->n, a->h, a->w, p.y_sw, p.y_sh, p.y_sn, p.Wb, p.Hb, p.Nb, 1);
    if (rc) return rc;
  } else {
    tmY = tmA0;
  }
  static bool attr_set = false;
  if (!attr_set) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  // persistent grid: one CTA per SM.  When the grid is a multiple of n_tiles every CTA keeps one n-tile for its
  // whole life and its BN statistics stay in registers; otherwise they are flushed whenever the slab changes.
  int ctas = num_sms();
  const long long total = (long long)p.m_tiles * p.n_tiles;
  B2_REQUIRE(total < (1ll << 31), B2_ERR_SHAPE, "too many tiles");
  if (ctas > total) ctas = (int)total;
  if (ctas > p.n_tiles && (total / ctas) >= 16) ctas = (ctas / p.n_tiles) * p.n_tiles;   // many tiles: keep slabs fixed
  p.m_stride = 0;
  conv_igemm_kernel<<<(unsigned)ctas, kThreads, smem_bytes, stream>>>(tmA0, tmA1, tmB, tmY, p);
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
