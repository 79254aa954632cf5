// b200seg — row-streaming 3x3 convolution for Cout = 64 layers on wide images (W a multiple of 128): fprop of
// conv_block's 64-channel convolutions (AttentionUNet.py:4-13 at the 256^2 level, R2U_Net.py:4-21) and their dgrad.
//
// Why a second tcgen05 kernel.  With the generic tile kernel (conv_igemm.cu) an N = 64 MMA re-reads its 4 KB A slice
// (128 pixels x 16 channels) for only 64 output columns, and the shared-memory operand bandwidth (about 110 B/clk per
// SM) held these layers at 45 % tensor-pipe activity whatever was done about loads, issue rate or weights.  Here one
// INPUT row feeds the three OUTPUT rows it contributes to in a single MMA:
//     D[128 px, 192] (+)= A_row[128 px, 16 ch] * [W(kr=2) ; W(kr=1) ; W(kr=0)][192, 16 ch]^T
// columns 0..63 belong to output row r-1, 64..127 to row r, 128..191 to row r+1 — so N = 192 and every input row is
// fetched ONCE (one halo box of 130 pixels; the three horizontal taps are descriptor shifts of one pixel), instead of
// three times with N = 64.  Shared-memory bytes per output row tile: 185 KB -> the generic kernel moves 279 KB.
//
//   TMEM     : eight accumulator slots of 64 columns = a ring over output rows (local row counter lr -> slot lr & 7);
//              the MMA of input row r writes the slots of rows r-1, r, r+1 (split in two where the ring wraps).  The
//              first MMA that touches a row's slot runs with accumulate = 0 (it is issued as its own N = 64 piece).
//   weights  : all nine taps (x channel blocks) resident in shared memory, stacked [kr=2; kr=1; kr=0] per (s, block).
//   work     : the sequence of all row tiles (n, column segment, h) is cut into one contiguous range per CTA; a range
//              is walked as strips of consecutive rows of one image column segment.  Rows outside a strip are neither
//              accumulated nor stored (the boundary input rows run N = 64 / 128 pieces), so strips cost one extra input
//              row at each inner end.
//   fold     : the folded UpConv with Cout = 64 (Upsample x2 -> conv3x3 as four 2x2 phase convolutions of the COARSE
//              input, AttentionUNet.py:15-27; ops.upconv_bn_act) streams the same way.  A CTA owns one column phase b
//              (fine pixels 2w + b; half of the CTAs each): coarse input row r feeds the FOUR fine rows 2r-1 .. 2r+2
//              — stacked weights [(a=1,ty=1) ; (a=0,ty=1) ; (a=1,ty=0) ; (a=0,ty=0)] per (tx, channel block), N = 256 —
//              with the two horizontal taps tx as descriptor shifts of b + tx pixels, and the ring advances by two
//              rows per input row.  Every input row is fetched once per column phase instead of once per (phase, tap):
//              the generic kernel's merged launch pulled 5.4 GB through the L2 -> SM fabric for 0.27 GB of input.
//   warps    : 0 = TMA producer, 1 = MMA issuer, 2-5 / 6-9 = two epilogue groups taking alternate output rows
//              (tcgen05.ld -> +bias, ReLU -> bf16 -> swizzled staging tile -> TMA store, BatchNorm sum / sum of squares
//              of the rounded values as in conv_igemm.cu).
#include <mutex>

#include "common.cuh"
#include "igemm_util.cuh"

namespace b2 {

int encode_act_tmap_ex(CUtensorMap* tm, const void* base, int c, int n, int h, int w, long long s_w, long long s_h,
                       long long s_n, int Wb, int Hb, int Nb, int es);

static constexpr int kC64Threads = 64 + 256;
static constexpr int kC64MaxStages = 8;
static constexpr int kC64Slots = 8;
static constexpr int kC64AStage = 17408;          // (128 + 2) pixels x 128 B, rounded up to 1024
static constexpr int kC64ATx = 130 * 128;

struct C64Params {
  int H, W, N, tw;          // image extent, column segments per row
  int cbt, cb0;             // 64-channel blocks of K in total / in source 0
  int stages, epi_groups;
  int grid;                 // CTAs (the row tiles are split into `grid` contiguous ranges)
  long long rows_total;     // N * tw * H
  uint32_t mg_tw;           // magic multiplier of tw (fast_div)
  const float* bias;
  double* stats;
  double* stats_partial;    // deterministic mode: [grid * epi_groups][128]
  int relu;
  int add;                  // 1: an addend tile (tmAdd; same geometry as y) is TMA-loaded into the group's staging tile at
                            // the start of each row's drain and added before ReLU / rounding (dgrad + incoming gradient)
  int fold;                 // 1: folded UpConv (see the header); H, W, rows_total are those of the COARSE input
  int debug_skip;           // timing experiments only (B200SEG_C64_SKIP; results are wrong): 1 no TMA loads, 2 no MMAs,
                            // 4 no drain at all, 8 no statistics pass, 16 no TMA store
};

// One strip = consecutive output rows [ha, hb) of image n, columns [w0, w0 + 128).
struct Strip {
  int n, w0, ha, hb;
};

// Walks the strips of the row-tile range [cur, end): returns false when the range is exhausted.
__device__ __forceinline__ bool next_strip(const C64Params& p, long long& cur, long long end, Strip& s) {
  if (cur >= end) return false;
  const int col = (int)cur / p.H;                 // rows_total < 2^31; one division per strip
  const int ha = (int)cur - col * p.H;
  long long hb = (long long)ha + (end - cur);
  if (hb > p.H) hb = p.H;
  const int n = fast_div(col, p.tw, p.mg_tw);
  s.n = n;
  s.w0 = (col - n * p.tw) * kTileM;
  s.ha = ha;
  s.hb = (int)hb;
  cur += hb - ha;
  return true;
}

__global__ void __launch_bounds__(kC64Threads, 1)
conv_c64_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY,
                const __grid_constant__ CUtensorMap tmAdd, const C64Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int wres_bytes = (p.fold ? 8 : 9) * p.cbt * 8192;              // taps x cbt blocks x 64 rows x 128 B
  uint8_t* wres = smem + p.stages * kC64AStage;                        // resident weights
  uint8_t* ctile0 = wres + wres_bytes;                                 // per group: 128 x 64 bf16 staging tile
  uint8_t* tail = ctile0 + p.epi_groups * (kTileM * 128);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);              // [kC64MaxStages]
  uint64_t* empty_bar = full_bar + kC64MaxStages;                      // [kC64MaxStages]
  uint64_t* tmem_full_bar = empty_bar + kC64MaxStages;                 // [kC64Slots]
  uint64_t* tmem_empty_bar = tmem_full_bar + kC64Slots;                // [kC64Slots]
  uint64_t* wres_bar = tmem_empty_bar + kC64Slots;
  uint64_t* add_bar = wres_bar + 1;                                    // [2]: a group's addend tile has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(add_bar + 2);
  float* s_bias = reinterpret_cast<float*>(tail + 336);                // [64]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // fold: CTAs [0, grid/2) take column phase 0, the rest phase 1; each half splits all (coarse) row tiles
  const int fb = p.fold ? ((int)blockIdx.x >= p.grid / 2 ? 1 : 0) : 0;
  const int rgrid = p.fold ? (fb ? p.grid - p.grid / 2 : p.grid / 2) : p.grid;
  const int ridx = p.fold ? (fb ? (int)blockIdx.x - p.grid / 2 : (int)blockIdx.x) : (int)blockIdx.x;
  const long long range_begin = p.rows_total * (long long)ridx / rgrid;
  const long long range_end = p.rows_total * (long long)(ridx + 1) / rgrid;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < kC64Slots; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 4);         // one arrival per warp of the draining epilogue group
    }
    mbar_init(wres_bar, 1);
    mbar_init(&add_bar[0], 1);
    mbar_init(&add_bar[1], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmY);
    if (p.cbt > p.cb0) tma_prefetch_desc(&tmA1);
  }
  pdl_enter();            // (B200SEG_PDL) barrier init / descriptor prefetch above overlap the previous kernel's tail
  if (threadIdx.x >= 64 && threadIdx.x < 128) s_bias[threadIdx.x - 64] = p.bias ? p.bias[threadIdx.x - 64] : 0.f;
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      mbar_arrive_expect_tx(wres_bar, (uint32_t)wres_bytes);
      if (p.fold) {
        for (int s = 0; s < 2; ++s)            // s = tx
          for (int cb = 0; cb < p.cbt; ++cb)
            for (int j = 0; j < 4; ++j) {      // fine row 2r-1+j: row phase a = 1 - (j & 1), filter row ty = 1 - (j >> 1)
              const int wa = 1 - (j & 1), wty = 1 - (j >> 1);
              tma_load_3d(wres + ((s * p.cbt + cb) * 256 + j * 64) * 128, &tmB, wres_bar, cb * kKBlock, 0,
                          (2 * wa + fb) * 4 + 2 * wty + s);
            }
      } else
      for (int s = 0; s < 3; ++s)
        for (int cb = 0; cb < p.cbt; ++cb)
          for (int j = 0; j < 3; ++j)          // stacked [kr = 2 ; kr = 1 ; kr = 0] for horizontal tap s
            tma_load_3d(wres + ((s * p.cbt + cb) * 192 + j * 64) * 128, &tmB, wres_bar, cb * kKBlock, 0, (2 - j) * 3 + s);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    long long cur = range_begin;
    Strip st;
    while (next_strip(p, cur, range_end, st)) {
      const int r0 = st.ha > 0 ? st.ha - 1 : 0;
      const int rl = st.hb < p.H ? st.hb : p.H - 1;
      for (int r = r0; r <= rl; ++r) {
        for (int cb = 0; cb < p.cbt; ++cb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            if (p.debug_skip & 1) {
              mbar_arrive(&full_bar[stage]);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)kC64ATx);
              const CUtensorMap* tm = cb < p.cb0 ? &tmA0 : &tmA1;
              const int c = (cb < p.cb0 ? cb : cb - p.cb0) * kKBlock;
              tma_load_4d(smem + stage * kC64AStage, tm, &full_bar[stage], c, st.w0 - 1, r, st.n);
            }
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
    // The tensor pipe queues only a few MMAs ahead of the issuing thread, so whatever this thread executes between
    // the last MMA of one input row and the first of the next is time the pipe idles (measured: a first version with
    // 330 instructions per row spent 1500 of 2500 clocks per row there).  Hence: incremental bookkeeping, no 64-bit
    // arithmetic or divisions per row, and straight-line MMA sequences with constant descriptor offsets.
    const uint64_t desc0 = umma_desc_sw128(0, 16, 1024);          // K-major SW128 descriptor with start address 0
    const uint64_t desc_hi = desc0 & 0xFFFFFFFF00000000ull;
    const uint32_t desc_lo0 = (uint32_t)desc0;
    const uint32_t a_ring_lo = desc_lo0 + (smem_u32(smem) >> 4);
    const uint32_t wres_lo = desc_lo0 + (smem_u32(wres) >> 4);
    const uint32_t idesc_n0 = umma_idesc_bf16(kTileM, 0, 0, 0);
    const uint32_t idesc1 = idesc_n0 + (1u << 20), idesc2 = idesc_n0 + (2u << 20), idesc3 = idesc_n0 + (3u << 20);
    const uint32_t s_step = (uint32_t)p.cbt * 1536u;               // 192 rows x 128 B per (s, channel block), in 16 B units
    const bool no_mma = (p.debug_skip & 2) != 0;
    mbar_wait(wres_bar, 0);
    tc_fence_after();
    int stage = 0;
    uint32_t phase = 0;
    uint32_t lr_base = 0;                         // local row counter of the strip's first output row
    long long cur = range_begin;
    Strip st;
    while (p.fold && next_strip(p, cur, range_end, st)) {
      // ---- folded UpConv: coarse input row r -> fine rows 2r-1 .. 2r+2 of the strip's fine range [2 ha, 2 hb) ----
      const int r0 = st.ha > 0 ? st.ha - 1 : 0;
      const int rl = st.hb < p.H ? st.hb : p.H - 1;
      const int f0 = 2 * st.ha, f1 = 2 * st.hb - 1;                    // first / last fine row of the strip
      const uint32_t fs_step = (uint32_t)p.cbt * 2048u;                // 256 rows x 128 B per (tx, channel block)
      for (int r = r0; r <= rl; ++r) {
        const int lo = 2 * r - 1 > f0 ? 2 * r - 1 : f0;
        const int hi = 2 * r + 2 < f1 ? 2 * r + 2 : f1;
        const int cnt = hi - lo + 1;                                   // 1 .. 4
        const uint32_t lr_lo = lr_base + (uint32_t)(lo - f0);
        const uint32_t sl = lr_lo & 7u;
        const uint32_t off = (uint32_t)(lo - (2 * r - 1));             // position in the 256-row weight stack
        // fresh rows (first touched by this input row): all at the strip's first input row, else 2r+1 and 2r+2
        int nfresh = r == r0 ? cnt : hi - 2 * r;
        if (nfresh < 0) nfresh = 0;
        for (int j = cnt - nfresh; j < cnt; ++j) {
          const uint32_t lr = lr_lo + (uint32_t)j;
          mbar_wait(&tmem_empty_bar[lr & 7u], ((lr >> 3) & 1u) ^ 1u);
        }
        tc_fence_after();
        const int c0 = (int)(kC64Slots - sl) < cnt ? (int)(kC64Slots - sl) : cnt;   // rows before the ring wraps
        const uint32_t d0 = tmem_base + sl * 64u;
        for (int cb = 0; cb < p.cbt; ++cb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            // the halo box starts at coarse pixel w0 - 1: tap tx of column phase b reads pixel w + b + tx - 1
            const uint32_t a0 = a_ring_lo + (uint32_t)stage * (uint32_t)(kC64AStage >> 4) + 8u * (uint32_t)fb;
            const uint32_t b0 = wres_lo + (uint32_t)cb * 2048u + off * 512u;
            if (!no_mma) {
              if (cb == 0) {             // K step 0: one N = 64 MMA per fine row, the fresh ones overwrite
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (j < cnt)
                    umma_bf16(tmem_base + ((sl + (uint32_t)j) & 7u) * 64u, desc_hi | (uint64_t)a0,
                              desc_hi | (uint64_t)(b0 + (uint32_t)j * 512u), idesc1, j >= cnt - nfresh ? 0u : 1u);
              } else if (c0 == cnt) {
                umma_bf16(d0, desc_hi | (uint64_t)a0, desc_hi | (uint64_t)b0, idesc_n0 + ((uint32_t)cnt << 20), 1u);
              } else {
                umma_bf16(d0, desc_hi | (uint64_t)a0, desc_hi | (uint64_t)b0, idesc_n0 + ((uint32_t)c0 << 20), 1u);
                umma_bf16(tmem_base, desc_hi | (uint64_t)a0, desc_hi | (uint64_t)(b0 + (uint32_t)c0 * 512u),
                          idesc_n0 + ((uint32_t)(cnt - c0) << 20), 1u);
              }
              // K steps 1 .. 7: tap tx = kk / 4, 16 channels k = kk % 4
              if (c0 == cnt) {
                const uint32_t id = idesc_n0 + ((uint32_t)cnt << 20);
#pragma unroll
                for (int kk = 1; kk < 8; ++kk)
                  umma_bf16(d0, desc_hi | (uint64_t)(a0 + (uint32_t)(8 * (kk >> 2) + 2 * (kk & 3))),
                            desc_hi | (uint64_t)(b0 + (uint32_t)(kk >> 2) * fs_step + (uint32_t)(2 * (kk & 3))), id, 1u);
              } else {
                const uint32_t id_a = idesc_n0 + ((uint32_t)c0 << 20), id_b = idesc_n0 + ((uint32_t)(cnt - c0) << 20);
                const uint32_t b1 = b0 + (uint32_t)c0 * 512u;
#pragma unroll
                for (int kk = 1; kk < 8; ++kk) {
                  const uint64_t da = desc_hi | (uint64_t)(a0 + (uint32_t)(8 * (kk >> 2) + 2 * (kk & 3)));
                  const uint32_t bk = (uint32_t)(kk >> 2) * fs_step + (uint32_t)(2 * (kk & 3));
                  umma_bf16(d0, da, desc_hi | (uint64_t)(b0 + bk), id_a, 1u);
                  umma_bf16(tmem_base, da, desc_hi | (uint64_t)(b1 + bk), id_b, 1u);
                }
              }
            }
            umma_commit(&empty_bar[stage]);
            if (cb == p.cbt - 1) {
              // fine rows that received their last contribution: those up to 2r, everything at the last input row
              for (int f = lo; f <= hi; ++f)
                if (f <= 2 * r || r == rl) umma_commit(&tmem_full_bar[(lr_base + (uint32_t)(f - f0)) & 7u]);
            }
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      lr_base += (uint32_t)(2 * (st.hb - st.ha));
    }
    while (next_strip(p, cur, range_end, st)) {
      const int r0 = st.ha > 0 ? st.ha - 1 : 0;
      const int rl = st.hb < p.H ? st.hb : p.H - 1;
      for (int r = r0; r <= rl; ++r) {
        const int lo = r - 1 > st.ha ? r - 1 : st.ha;                  // output rows this input row contributes to
        const int hi = r + 1 < st.hb - 1 ? r + 1 : st.hb - 1;
        const int cnt = hi - lo + 1;                                   // 1 .. 3
        const uint32_t lr_lo = lr_base + (uint32_t)(lo - st.ha);
        const uint32_t sl = lr_lo & 7u;                                // slot of row lo
        const uint32_t off = (uint32_t)(lo - (r - 1));                 // its position in the 192-row weight stack
        // rows touched for the first time by this input row (their first MMA overwrites): all of them at the strip's
        // first input row, else row r + 1 if it is part of the strip
        const int nfresh = r == r0 ? cnt : (r + 1 <= hi ? 1 : 0);
        for (int j = cnt - nfresh; j < cnt; ++j) {
          const uint32_t lr = lr_lo + (uint32_t)j;
          mbar_wait(&tmem_empty_bar[lr & 7u], ((lr >> 3) & 1u) ^ 1u);
        }
        tc_fence_after();
        const int c0 = (int)(kC64Slots - sl) < cnt ? (int)(kC64Slots - sl) : cnt;   // rows before the ring wraps
        const uint32_t d0 = tmem_base + sl * 64u;
        for (int cb = 0; cb < p.cbt; ++cb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a0 = a_ring_lo + (uint32_t)stage * (uint32_t)(kC64AStage >> 4);
            const uint32_t b0 = wres_lo + (uint32_t)cb * 1536u + off * 512u;     // 64 rows x 128 B = 512 units
            if (!no_mma) {
              // ---- K step 0 ----
              if (cb == 0) {             // one N = 64 MMA per output row: the fresh ones overwrite
#pragma unroll
                for (int j = 0; j < 3; ++j)
                  if (j < cnt)
                    umma_bf16(tmem_base + ((sl + (uint32_t)j) & 7u) * 64u, desc_hi | (uint64_t)a0,
                              desc_hi | (uint64_t)(b0 + (uint32_t)j * 512u), idesc1, j >= cnt - nfresh ? 0u : 1u);
              } else if (c0 == cnt) {
                umma_bf16(d0, desc_hi | (uint64_t)a0, desc_hi | (uint64_t)b0, idesc_n0 + ((uint32_t)cnt << 20), 1u);
              } else {
                umma_bf16(d0, desc_hi | (uint64_t)a0, desc_hi | (uint64_t)b0, idesc_n0 + ((uint32_t)c0 << 20), 1u);
                umma_bf16(tmem_base, desc_hi | (uint64_t)a0, desc_hi | (uint64_t)(b0 + (uint32_t)c0 * 512u),
                          idesc_n0 + ((uint32_t)(cnt - c0) << 20), 1u);
              }
              // ---- K steps 1 .. 11: tap s = kk / 4 (one pixel = 8 units further into the halo box, the next (s, block)
              //      of the weights), 16 channels k = kk % 4 (2 units) ----
              if (c0 == cnt) {           // no wrap: ONE MMA per K step (N = 192 in the interior of a strip)
                const uint32_t id = cnt == 3 ? idesc3 : (cnt == 2 ? idesc2 : idesc1);
#pragma unroll
                for (int kk = 1; kk < 12; ++kk)
                  umma_bf16(d0, desc_hi | (uint64_t)(a0 + (uint32_t)(8 * (kk >> 2) + 2 * (kk & 3))),
                            desc_hi | (uint64_t)(b0 + (uint32_t)(kk >> 2) * s_step + (uint32_t)(2 * (kk & 3))), id, 1u);
              } else {                   // the ring wraps inside the N range: two MMAs per K step
                const uint32_t id_a = idesc_n0 + ((uint32_t)c0 << 20), id_b = idesc_n0 + ((uint32_t)(cnt - c0) << 20);
                const uint32_t b1 = b0 + (uint32_t)c0 * 512u;
#pragma unroll
                for (int kk = 1; kk < 12; ++kk) {
                  const uint64_t da = desc_hi | (uint64_t)(a0 + (uint32_t)(8 * (kk >> 2) + 2 * (kk & 3)));
                  const uint32_t bk = (uint32_t)(kk >> 2) * s_step + (uint32_t)(2 * (kk & 3));
                  umma_bf16(d0, da, desc_hi | (uint64_t)(b0 + bk), id_a, 1u);
                  umma_bf16(tmem_base, da, desc_hi | (uint64_t)(b1 + bk), id_b, 1u);
                }
              }
            }
            umma_commit(&empty_bar[stage]);
            if (cb == p.cbt - 1) {
              // output rows that received their last contribution: r - 1, and r itself at the image's last row
              if (r - 1 >= st.ha) umma_commit(&tmem_full_bar[(lr_base + (uint32_t)(r - 1 - st.ha)) & 7u]);
              if (r == rl && r <= st.hb - 1) umma_commit(&tmem_full_bar[(lr_base + (uint32_t)(r - st.ha)) & 7u]);
            }
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      lr_base += (uint32_t)(st.hb - st.ha);
    }
  } else {
    // ---------------- epilogue: group 0 = warps 2..5, group 1 = warps 6..9 (output rows alternate) ----------------
    const int g = (warp - 2) >> 2;
    const int G = p.epi_groups;
    if (g < G) {
      const int et = (threadIdx.x - 64) & 127;
      uint8_t* ctile = ctile0 + g * (kTileM * 128);
      const int bar_id = 1 + g;
      const int quad = warp & 3;
      const int row = quad * 32 + lane;
      const bool relu = p.relu != 0;
      const int schunk = lane & 7;
      const int sgrp = et >> 3;
      double acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.0;
      float fs[8], fq[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) fs[j] = fq[j] = 0.f;
      int fpending = 0;
      uint32_t add_phase = 0;
      int lr = 0;
      long long cur = range_begin;
      Strip st;
      while (next_strip(p, cur, range_end, st)) {
        // fold: the strip's output rows are the fine rows [2 ha, 2 hb) of this CTA's column phase
        const int h_begin = p.fold ? 2 * st.ha : st.ha, h_end = p.fold ? 2 * st.hb : st.hb;
        for (int h = h_begin; h < h_end; ++h, ++lr) {
          if (G == 2 && (lr & 1) != g) continue;
          const uint32_t slot = (uint32_t)lr & 7u;
          // the group's previous TMA store must have finished READING the staging tile before it is overwritten
          if (et == 0) tma_store_wait_read();
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (p.add && et == 0) {
            mbar_arrive_expect_tx(&add_bar[g], (uint32_t)(kTileM * 128));
            tma_load_4d(ctile, &tmAdd, &add_bar[g], 0, st.w0, h, st.n);
          }
          mbar_wait(&tmem_full_bar[slot], ((uint32_t)lr >> 3) & 1u);
          tc_fence_after();
          if (p.debug_skip & 4) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[slot]);
            continue;
          }
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + slot * 64u;
          float v[64];
          tmem_ld32(taddr, v);
          tmem_ld32(taddr + 32, v + 32);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[slot]);        // accumulator drained
          const uint32_t row_off = (uint32_t)row * 128u;
          const uint32_t row_x = (uint32_t)(row & 7);
          if (p.add) {
            mbar_wait(&add_bar[g], add_phase);
            add_phase ^= 1u;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint4 au = *reinterpret_cast<const uint4*>(ctile + row_off + ((((uint32_t)q) ^ row_x) << 4));
              v[q * 8 + 0] += bf16lo(au.x); v[q * 8 + 1] += bf16hi(au.x);
              v[q * 8 + 2] += bf16lo(au.y); v[q * 8 + 3] += bf16hi(au.y);
              v[q * 8 + 4] += bf16lo(au.z); v[q * 8 + 5] += bf16hi(au.z);
              v[q * 8 + 6] += bf16lo(au.w); v[q * 8 + 7] += bf16hi(au.w);
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            uint32_t pk[4];
            const float4 b0 = *reinterpret_cast<const float4*>(s_bias + q * 8);       // broadcast reads
            const float4 b1 = *reinterpret_cast<const float4*>(s_bias + q * 8 + 4);
            const float bq[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = q * 8 + 2 * j;
              float a = v[e] + bq[2 * j];
              float b = v[e + 1] + bq[2 * j + 1];
              if (relu) {
                a = fmaxf(a, 0.f);
                b = fmaxf(b, 0.f);
              }
              pk[j] = pack_bf16x2(a, b);
            }
            *reinterpret_cast<uint4*>(ctile + row_off + ((((uint32_t)q) ^ row_x) << 4)) =
                make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          fence_proxy_async();
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (et == 0 && !(p.debug_skip & 16)) {
            if (p.fold) tma_store_5d(&tmY, ctile, 0, fb, st.w0, h, st.n);     // (c, column phase, coarse w, fine h, n)
            else tma_store_4d(&tmY, ctile, 0, st.w0, h, st.n);
            tma_store_commit();
          }
          if (p.stats != nullptr && !(p.debug_skip & 8)) {
            const uint8_t* base = ctile + (uint32_t)(sgrp * 8) * 128u;
            uint4 u[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              u[i] = *reinterpret_cast<const uint4*>(base + (uint32_t)i * 128u + ((((uint32_t)schunk) ^ (uint32_t)i) << 4));
#pragma unroll
            for (int i = 0; i < 8; ++i) stats_accum(u[i], fs, fq);
            if (++fpending == 4) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                acc[j] += (double)fs[j];
                acc[8 + j] += (double)fq[j];
                fs[j] = fq[j] = 0.f;
              }
              fpending = 0;
            }
          }
        }
      }
      if (p.stats != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j] += (double)fs[j];
          acc[8 + j] += (double)fq[j];
        }
        // lanes owning the same chunk are combined by shuffles, the four warps through the staging tile
        for (int off = 8; off < 32; off <<= 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
        }
        if (et == 0) tma_store_wait_read();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        double* red = reinterpret_cast<double*>(ctile);                 // [warp][sum | sumsq][64]
        if (lane < 8) {
          double* r0 = red + ((warp - 2) & 3) * 128 + schunk * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            r0[j] = acc[j];
            r0[64 + j] = acc[8 + j];
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        {
          const int i = et;                                             // 128 threads: [sum | sumsq] x 64 channels
          double v = 0.0;
#pragma unroll
          for (int w = 0; w < 4; ++w) v += red[w * 128 + i];
          if (p.stats_partial != nullptr) p.stats_partial[(size_t)(blockIdx.x * G + g) * 128 + i] = v;
          else atomicAdd(&p.stats[i], v);
        }
      }
      if (et == 0) tma_store_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// Eligibility + launch.  Returns 1 when the layer was launched here, 0 when the caller should use the generic kernel,
// a negative error code on failure.
int conv_c64_try_launch(const b2_conv_args* a, cudaStream_t stream) {
  if (env_switch("B200SEG_C64", 1) == 0) return 0;
  const int stride = a->stride == 0 ? 1 : a->stride;
  const int out_mul = a->out_mul == 0 ? 1 : a->out_mul;
  const int in_mul = a->in_mul == 0 ? 1 : a->in_mul;
  const int fold = a->fold_mode == 1 ? 1 : 0;           // merged folded-UpConv fprop (ksize 2, y on the 2x grid)
  if (fold && (a->ksize != 2 || a->c1 != 0 || a->ldy != a->cout || env_switch("B200SEG_C64_FOLD", 1) == 0)) return 0;
  if ((!fold && (a->ksize != 3 || a->fold_mode != 0)) || a->cout != 64 || stride != 1 || out_mul != 1 || in_mul != 1 ||
      a->custom_pad != 0 || a->w % kTileM != 0 || a->c0 % 64 != 0 || a->c1 % 64 != 0)
    return 0;
  if (a->addend != nullptr && (fold || a->add_after_act || a->ldadd % 8 != 0 ||
                               (reinterpret_cast<uintptr_t>(a->addend) & 15) != 0))
    return 0;
  const int cb0 = a->c0 / 64, cbt = cb0 + a->c1 / 64;
  if (cbt < 1 || cbt > 2) return 0;
  if (env_switch("B200SEG_DEBUG_SKIP", 0) != 0 || env_switch("B200SEG_BLOCK_N", 0) != 0) return 0;
  const long long rows_total = (long long)a->n * (a->w / kTileM) * a->h;
  if (rows_total < 8ll * num_sms() || rows_total >= (1ll << 31) || a->h < 4) return 0;   // strips need some length
  B2_REQUIRE(a->ldy % 8 == 0 && (reinterpret_cast<uintptr_t>(a->y) & 15) == 0, B2_ERR_ALIGN, "y misaligned");
  B2_REQUIRE(a->ktot >= a->c0 + a->c1 && a->ktot % 8 == 0, B2_ERR_SHAPE, "ktot=%d inconsistent", a->ktot);

  C64Params p;
  memset(&p, 0, sizeof(p));
  p.H = a->h; p.W = a->w; p.N = a->n;
  p.tw = a->w / kTileM;
  p.cbt = cbt; p.cb0 = cb0;
  p.rows_total = rows_total;
  p.bias = a->bias;
  p.stats = a->stats;
  p.relu = a->relu;
  p.add = a->addend != nullptr ? 1 : 0;
  p.fold = fold;
  p.debug_skip = env_switch("B200SEG_C64_SKIP", 0);
  p.mg_tw = p.tw <= 1 ? 0u : (uint32_t)(((1ull << 32) + (uint64_t)p.tw - 1) / (uint64_t)p.tw);
  const int tail_bytes = 320 + 16 + 64 * 4 + 192;
  const int wres_bytes = (fold ? 8 : 9) * cbt * 8192;
  p.epi_groups = 2;
  int budget = 232448 - 1024 - tail_bytes - wres_bytes - 2 * kTileM * 128;
  if (budget / kC64AStage < 3) {       // two channel blocks: one staging tile, the main loop is twice as long anyway
    p.epi_groups = 1;
    budget += kTileM * 128;
  }
  int stages = budget / kC64AStage;
  if (stages > kC64MaxStages) stages = kC64MaxStages;
  if (stages < 2) return 0;
  p.stages = stages;
  const int smem_bytes = stages * kC64AStage + wres_bytes + p.epi_groups * kTileM * 128 + tail_bytes + 1024;
  p.grid = num_sms();

  CUtensorMap tmA0, tmA1, tmB, tmY;
  int rc = encode_act_tmap_ex(&tmA0, a->x0, a->c0, a->n, a->h, a->w, a->ldx0, (long long)a->ldx0 * a->w,
                              (long long)a->ldx0 * a->w * a->h, kTileM + 2, 1, 1, 1);
  if (rc) return rc;
  if (a->c1 > 0) {
    rc = encode_act_tmap_ex(&tmA1, a->x1, a->c1, a->n, a->h, a->w, a->ldx1, (long long)a->ldx1 * a->w,
                            (long long)a->ldx1 * a->w * a->h, kTileM + 2, 1, 1, 1);
    if (rc) return rc;
  } else {
    tmA1 = tmA0;
  }
  {
    uint64_t dims[3] = {(uint64_t)(a->c0 + a->c1), (uint64_t)a->cout, (uint64_t)(fold ? 16 : 9)};
    uint64_t str[3] = {2, (uint64_t)a->ktot * 2, (uint64_t)a->w_tap_stride * 2};
    uint32_t box[3] = {64, 64, 1};
    rc = encode_tmap_bf16(&tmB, a->wpk, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (fold) {      // y [n, 2h, 2w, 64] as (c, column phase, coarse w, fine h, n): a store writes every other fine pixel
    const uint64_t ld = (uint64_t)a->ldy;
    uint64_t dims[5] = {64, 2, (uint64_t)a->w, (uint64_t)a->h * 2, (uint64_t)a->n};
    uint64_t str[5] = {2, ld * 2, ld * 4, ld * 4 * a->w, ld * 8 * a->w * a->h};
    uint32_t box[5] = {64, 1, (uint32_t)kTileM, 1, 1};
    rc = encode_tmap_bf16(&tmY, a->y, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
  } else {
    rc = encode_act_tmap_ex(&tmY, a->y, a->cout, a->n, a->h, a->w, a->ldy, (long long)a->ldy * a->w,
                            (long long)a->ldy * a->w * a->h, kTileM, 1, 1, 1);
  }
  if (rc) return rc;
  CUtensorMap tmAdd = tmY;
  if (p.add) {
    rc = encode_act_tmap_ex(&tmAdd, a->addend, a->cout, a->n, a->h, a->w, a->ldadd, (long long)a->ldadd * a->w,
                            (long long)a->ldadd * a->w * a->h, kTileM, 1, 1, 1);
    if (rc) return rc;
  }
  {
    static std::once_flag attr_once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once, [] {
      attr_err = cudaFuncSetAttribute(conv_c64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    });
    B2_CHECK_CUDA(attr_err);
  }
  DetBuf det;
  det.partial = nullptr;
  const long long det_rows = (long long)p.grid * p.epi_groups;
  if (p.stats != nullptr) {
    rc = det_begin(&det, det_rows, 128, stream);
    if (rc) return rc;
  }
  p.stats_partial = det.partial;
  B2_CHECK_CUDA(launch_chain(conv_c64_kernel, dim3(p.grid), dim3(kC64Threads), (size_t)(smem_bytes), stream, 1,
      (long long)p.N * p.H * p.W * 128 * (p.fold ? 4 : 1), tmA0, tmA1, tmB, tmY, tmAdd, p));
  B2_LAUNCH_CHECK();
  if (det.partial) {
    rc = det_finish(det.partial, det_rows, 128, 128, p.stats, stream);
    if (rc) return rc;
  }
  return 1;
}

}  // namespace b2
