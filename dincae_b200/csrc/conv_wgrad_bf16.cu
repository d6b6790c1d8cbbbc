// Weight + bias gradient of the 3x3 convolutions in bf16 on the 5th-gen tensor cores
// (tcgen05.mma kind::f16, fp32 accumulators in TMEM) for any channel width.
//
//   dW[dy][dx][ci][co] = sum over pixels  U[y+dy-1][x+dx-1][ci] * G[y][x][co]
//
// The reduction runs over PIXELS.  In the planar-8 bf16 layout a pixel of a plane is 16 bytes
// (8 channels) and the pixels of a row are contiguous, which is exactly the no-swizzle
// MN-major canonical operand layout of kind::f16 (core matrix = 8 pixels x 8 channels = 128
// contiguous bytes; next 8 pixels at LBO = 128 B; next 8 channels = next plane at SBO) --
// checked on the device by tools/umma_bf16_test.cu.  So both operands are used AS STORED:
// planes that exist at the conv's resolution go from HBM to shared memory by TMA (one box
// per plane: (R+2) x (TW+2) pixels of U, R x TW pixels of G, out-of-image pixels zero-filled)
// and the MMA reads them in place: A = U (M = input channels), B = G (N = output channels),
// K = 16 consecutive pixels of a row; a tap (dy,dx) is a shift of A's start address by dy rows
// and dx pixels.  No transposition, no ALU work on the operands.  Only x2-upsampled planes
// (decoder input) and un-pooled gradients (encoder) pass through registers.
//
// Work decomposition: job = (M-tile of 64 / 128 input channels, group of taps) so that the
// accumulators (taps x N fp32 columns) fit the 512 TMEM columns; every job is split over
// `splits` CTAs by tiles (image, row block, column block).  A CTA keeps its accumulators in
// TMEM over all its tiles and writes one partial; dincae_conv3x3_wgrad_reduce sums the
// partials of a job in fixed order (deterministic).  The bias gradient is the row of a
// constant-one channel placed in the first free channel slot of U.
#include "bf16.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace dincae {

constexpr int WB_MAX_LOAD_WARPS = 16;          // loader warps (a multiple of 4) + one MMA warp; runtime choice
constexpr int WB_THREADS = 32 * (WB_MAX_LOAD_WARPS + 1);
constexpr int WB_MAX_STAGES = 3;
constexpr int WB_SMEM_BUDGET = 214 * 1024;

struct WgBfParams {
  const uint4* u0;
  const uint4* u1;
  const uint4* g;
  const uint8_t* sign;
  int p0, p1, half0, gmode;
  float slope;
  int B, H, W, Hh, Wh, Wq;
  int kp, gp;                  // planes of U and G
  int cinp4, cout_pad, wsize;  // partial layout: WP [9][cinp4][cout_pad][4] then [cout_pad]
  float* partial;
  // plan
  int R, TW, TWu, tiles_x, n_row_blocks, n_tiles;
  int M, N, slots, u_alloc, mt_count, tg_count, tpg, jobs, splits;
  int ones_mt, ones_slot;
  int ones_sep, ones_off;      // bias row as a separate all-ones MMA (the planes fill the M tiles exactly); its 256 bytes
  int sbo_u, sbo_g, u_bytes, stage_bytes, n_stages;
  int tmem_cols;
  int tma_u0, tma_u1, tma_g;
  int toff[12];                // per tap: dy * TWu + dx (16-byte units), precomputed on the host
  unsigned idesc;              // instruction descriptor and the constant halves of the operand descriptors
  unsigned long long a_hi, b_hi;
  int load_warps, poll_all;    // experiment knobs (DINCAE_WB_WARPS, DINCAE_WB_POLL_ALL)
  long long* dbg;              // DINCAE_WB_DEBUG=1: role timeline of CTA 0 ([2][256][4] clock64 values)
};

// One pixel row of a register-path plane (up to 4 x 32 pixels) in two halves, so that the loads of
// all rows a warp owns in a tile are in flight before the first one is used.
struct WbRow {
  uint4 v[4];
  unsigned sb[4];
};

template <int KIND>   // 0: x2-upsampled U plane, 1: un-pooled G plane
__device__ __forceinline__ void wb_issue(const WgBfParams& P, int b, int plane, int y, int x_first, int width,
                                         int lane, WbRow& row) {
  const bool yok = (unsigned)y < (unsigned)P.H;
  const uint4* src;
  const uint8_t* sg = nullptr;
  if (KIND == 0) {
    src = P.u0 + ((size_t)(b * P.p0 + plane) * P.Hh + (y >> 1)) * P.Wh;
  } else {
    src = P.g + ((size_t)(b * P.gp + plane) * P.Hh + (y >> 1)) * P.Wh;
    sg = P.sign + ((size_t)(b * P.gp + plane) * P.H + y) * (4 * P.Wq);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int px = lane + 32 * u, x = x_first + px;
    const bool ok = yok && px < width && (unsigned)x < (unsigned)P.W;
    row.v[u] = ldg_u4_pred(src + (x >> 1), ok);
    row.sb[u] = 0;
    if (KIND == 1) row.sb[u] = ldg_u8_pred(sg + x, ok);
  }
}

template <int KIND>
__device__ __forceinline__ void wb_commit(const WgBfParams& P, int y, int x_first, int width, uint4* dst, int lane,
                                          const WbRow& row) {
  const float inv_y = (2 * (y >> 1) + 1 < P.H) ? 0.5f : 1.0f;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int px = lane + 32 * u, x = x_first + px;
    if (px < width) {
      uint4 t = row.v[u];
      if (KIND == 1) t = bf8_unpool(t, row.sb[u], inv_y * ((2 * (x >> 1) + 1 < P.W) ? 0.5f : 1.0f), P.slope);
      dst[px] = t;
    }
  }
}

__global__ void __launch_bounds__(WB_THREADS, 1)
conv3x3_wgrad_bf16_kernel(const WgBfParams P, const __grid_constant__ CUtensorMap tmu0,
                          const __grid_constant__ CUtensorMap tmu1, const __grid_constant__ CUtensorMap tmg) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* ring = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  __shared__ __align__(8) uint64_t full_bar[WB_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[WB_MAX_STAGES];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_holder;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int WB_LOAD_WARPS = P.load_warps, WB_MMA_WARP = P.load_warps;
  const int nthreads = 32 * (P.load_warps + 1);
  const int job = blockIdx.x % P.jobs, split = blockIdx.x / P.jobs;
  const int mt = job % P.mt_count, tg = job / P.mt_count;
  const int tap0 = tg * P.tpg;
  const int ntaps = (9 - tap0) < P.tpg ? (9 - tap0) : P.tpg;
  const int plane0 = mt * P.slots;                               // first U plane of this M-tile
  int n_u = P.kp - plane0;                                       // real U planes of this M-tile
  if (n_u > P.slots) n_u = P.slots;
  const int S = P.n_stages;

  // ---- one-time setup: zero the ring, ones plane, barriers, TMEM --------------------------------
  pdl_trigger();
  {
    const int n16 = (S * P.stage_bytes + (P.slots - P.u_alloc) * P.sbo_u) >> 4;
    for (int i = tid; i < n16; i += nthreads) reinterpret_cast<uint4*>(ring)[i] = u4_zero();
  }
  __syncthreads();
  if (P.ones_sep) {
    // 256 bytes of bf16 ones behind the ring: operand A of the bias MMA (every M row, 16 pixels)
    if (tid < 16) reinterpret_cast<uint4*>(ring + P.ones_off)[tid] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
  } else if (mt == P.ones_mt) {
    // channel lane 0 of plane slot `ones_slot` = 1.0 (bf16 0x3f80) at every pixel of every stage
    const int npx = P.sbo_u >> 4;
    for (int s = 0; s < S; ++s) {
      uint4* pl = reinterpret_cast<uint4*>(ring + (size_t)s * P.stage_bytes + (size_t)P.ones_slot * P.sbo_u);
      for (int i = tid; i < npx; i += nthreads) pl[i] = make_uint4(0x3f80u, 0u, 0u, 0u);
    }
  }
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], WB_LOAD_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (P.tma_u0) asm volatile("prefetch.tensormap [%0];" :: "l"(&tmu0) : "memory");
    if (P.tma_u1) asm volatile("prefetch.tensormap [%0];" :: "l"(&tmu1) : "memory");
    if (P.tma_g) asm volatile("prefetch.tensormap [%0];" :: "l"(&tmg) : "memory");
  }
  if (warp == WB_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&tmem_holder)), "r"(P.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  pdl_wait();
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  const int per_img = P.n_row_blocks * P.tiles_x;
  const int u_row_bytes = P.TWu * 16, g_row_bytes = P.TW * 16;

  if (warp < WB_LOAD_WARPS) {
    // =============================== loader ===================================================
    // TMA planes: one elected thread; register-path planes: (plane, row) items round-robin over
    // the loader warps.  Every warp arrives once per stage.
    const int n_u_reg = P.half0 ? ((plane0 < P.p0) ? ((P.p0 - plane0 < n_u) ? P.p0 - plane0 : n_u) : 0) : 0;
    const int items_u = n_u_reg * (P.R + 2);
    const int items_g = P.gmode == 1 ? P.gp * P.R : 0;
    uint32_t tx_bytes = 0;
    for (int k = 0; k < n_u; ++k) {
      const int plane = plane0 + k;
      const bool first = plane < P.p0;
      if (first ? P.tma_u0 : P.tma_u1) tx_bytes += (uint32_t)((P.R + 2) * u_row_bytes);
    }
    if (P.tma_g) tx_bytes += (uint32_t)(P.gp * P.R * g_row_bytes);
    int s = 0;
    uint32_t ph = 0;
    int dbg_n = 0;
    for (int tile = split; tile < P.n_tiles; tile += P.splits) {
      const int b = tile / per_img, rem = tile - b * per_img;
      const int rb = rem / P.tiles_x, tx = rem - rb * P.tiles_x;
      const int y0 = rb * P.R, x0 = tx * P.TW;
      unsigned char* Us = ring + (size_t)s * P.stage_bytes;
      unsigned char* Gs = Us + P.u_bytes;
      // register-path rows: the loads of the first two items of this warp are issued before the
      // ring slot is known to be free (they only read global memory)
      constexpr int WB_BATCH = 2;
      WbRow rows[WB_BATCH];
      const int n_items = items_u + items_g;
      auto issue = [&](int it, WbRow& row) {
        if (it < items_u) {
          const int k = it / (P.R + 2), r = it - k * (P.R + 2);
          wb_issue<0>(P, b, plane0 + k, y0 - 1 + r, x0 - 1, P.TWu, lane, row);
        } else if (it < n_items) {
          const int i2 = it - items_u, k = i2 / P.R, r = i2 - k * P.R;
          wb_issue<1>(P, b, k, y0 + r, x0, P.TW, lane, row);
        }
      };
      auto commit = [&](int it, const WbRow& row) {
        if (it < items_u) {
          const int k = it / (P.R + 2), r = it - k * (P.R + 2);
          wb_commit<0>(P, y0 - 1 + r, x0 - 1, P.TWu,
                       reinterpret_cast<uint4*>(Us + (size_t)k * P.sbo_u + (size_t)r * u_row_bytes), lane, row);
        } else if (it < n_items) {
          const int i2 = it - items_u, k = i2 / P.R, r = i2 - k * P.R;
          wb_commit<1>(P, y0 + r, x0, P.TW,
                       reinterpret_cast<uint4*>(Gs + (size_t)k * P.sbo_g + (size_t)r * g_row_bytes), lane, row);
        }
      };
#pragma unroll
      for (int j = 0; j < WB_BATCH; ++j) issue(warp + j * WB_LOAD_WARPS, rows[j]);
      long long t_a = 0, t_b = 0, t_c = 0;
      if (P.dbg) t_a = clock64();
      if (P.poll_all) mbar_wait(&empty_bar[s], ph ^ 1u); else mbar_wait_lane0(&empty_bar[s], ph ^ 1u);
      if (P.dbg) t_b = clock64();
      if (warp == 0 && tx_bytes) {
        if (elect_one()) {
          mbar_expect_tx(&full_bar[s], tx_bytes);
          for (int k = 0; k < n_u; ++k) {
            const int plane = plane0 + k;
            const bool first = plane < P.p0;
            if (!(first ? P.tma_u0 : P.tma_u1)) continue;
            tma_load_4d(Us + (size_t)k * P.sbo_u, first ? &tmu0 : &tmu1, (x0 - 1) * 2, y0 - 1,
                        first ? plane : plane - P.p0, b, &full_bar[s]);
          }
          if (P.tma_g)
            for (int k = 0; k < P.gp; ++k)
              tma_load_4d(Gs + (size_t)k * P.sbo_g, &tmg, x0 * 2, y0, k, b, &full_bar[s]);
        }
        __syncwarp();
      }
#pragma unroll
      for (int j = 0; j < WB_BATCH; ++j) commit(warp + j * WB_LOAD_WARPS, rows[j]);
      for (int it0 = warp + WB_BATCH * WB_LOAD_WARPS; it0 < n_items; it0 += WB_BATCH * WB_LOAD_WARPS) {
#pragma unroll
        for (int j = 0; j < WB_BATCH; ++j) issue(it0 + j * WB_LOAD_WARPS, rows[j]);
#pragma unroll
        for (int j = 0; j < WB_BATCH; ++j) commit(it0 + j * WB_LOAD_WARPS, rows[j]);
      }
      if (P.dbg) t_c = clock64();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
      if (P.dbg && blockIdx.x == 0 && warp == 0 && lane == 0 && dbg_n < 256) {
        long long* d = P.dbg + 4 * dbg_n++;
        d[0] = t_a; d[1] = t_b; d[2] = t_c; d[3] = clock64();
      }
      if (++s == S) { s = 0; ph ^= 1u; }
    }
  } else {
    // =============================== MMA issuer ===============================================
    const uint32_t idesc = P.idesc;
    const uint64_t a_hi = P.a_hi, b_hi = P.b_hi;
    int s = 0;
    uint32_t ph = 0;
    const uint32_t TWu = (uint32_t)P.TWu, TW = (uint32_t)P.TW;
    int dbg_m = 0;
    for (int tile = split; tile < P.n_tiles; tile += P.splits) {
      const int b = tile / per_img, rem = tile - b * per_img;
      const int rb = rem / P.tiles_x, tx = rem - rb * P.tiles_x;
      const int y0 = rb * P.R, x0 = tx * P.TW;
      const int rows = (P.H - y0) < P.R ? (P.H - y0) : P.R;
      const int wpx = (P.W - x0) < P.TW ? (P.W - x0) : P.TW;
      const int ksteps = (wpx + 15) >> 4;
      const bool first_tile = tile == split;     // the accumulators hold nothing yet
      (void)b;
      long long t_a = 0, t_b = 0;
      if (P.dbg) t_a = clock64();
      if (P.poll_all) mbar_wait(&full_bar[s], ph); else mbar_wait_lane0(&full_bar[s], ph);
      if (P.dbg) t_b = clock64();
      tc_fence_after();
      if (elect_one()) {
        const uint32_t u_lo = smem_u32(ring + (size_t)s * P.stage_bytes) >> 4;
        const uint32_t g_lo = u_lo + (uint32_t)(P.u_bytes >> 4);
        // Taps innermost: consecutive MMAs go to different accumulators (MMAs into the same TMEM
        // columns serialise on the accumulate latency).  The issuing thread keeps one RUNNING
        // 64-bit descriptor per tap and advances it by 16 pixels per K step, so that no MMA waits
        // on an address computation: a dependent chain of uniform-datapath instructions per MMA
        // (indexed constant load -> add -> mask -> add) cost ~70 clk per MMA in the first version
        // (role timeline of DINCAE_WB_DEBUG=1).  Start addresses stay below 2^14 16-byte units
        // (228 KB of shared memory), so the descriptor's address field never overflows.
        uint64_t ad[9];
#pragma unroll
        for (int tl = 0; tl < 9; ++tl) ad[tl] = a_hi + (uint64_t)(u_lo + (uint32_t)P.toff[tap0 + tl]);
        uint64_t bd = b_hi + (uint64_t)g_lo;
        // bias row as its own MMA (M rows of ones x G) when the planes leave no free slot in the M tile:
        // SBO = 0, every plane slot of the operand aliases the same 256 bytes of ones
        const bool bias_mma = P.ones_sep && mt == 0 && tg == 0;
        const uint64_t od = ((uint64_t)(128u >> 4) << 16) | (1ull << 46) | (uint64_t)(smem_u32(ring + P.ones_off) >> 4);
        const uint32_t bias_col = tmem_base + (uint32_t)(ntaps * P.N);
        const uint64_t a_skip = (uint64_t)(TWu - 16u * (uint32_t)ksteps), b_skip = (uint64_t)(TW - 16u * (uint32_t)ksteps);
        uint32_t accf = first_tile ? 0u : 1u;
        for (int r = 0; r < rows; ++r) {
          for (int kx = 0; kx < ksteps; ++kx) {
#pragma unroll
            for (int tl = 0; tl < 9; ++tl)
              if (tl < ntaps) umma_f16(tmem_base + (uint32_t)(tl * P.N), ad[tl], bd, idesc, accf);
            if (bias_mma) umma_f16(bias_col, od, bd, idesc, accf);
#pragma unroll
            for (int tl = 0; tl < 9; ++tl) ad[tl] += 16u;
            bd += 16u;
            accf = 1u;
          }
#pragma unroll
          for (int tl = 0; tl < 9; ++tl) ad[tl] += a_skip;
          bd += b_skip;
        }
        umma_commit(&empty_bar[s]);
      }
      __syncwarp();
      if (P.dbg && blockIdx.x == 0 && lane == 0 && dbg_m < 256) {
        long long* d = P.dbg + 1024 + 4 * dbg_m++;
        d[0] = t_a; d[1] = t_b; d[2] = clock64(); d[3] = (long long)rows * ksteps * ntaps;
      }
      if (++s == S) { s = 0; ph ^= 1u; }
    }
    if (elect_one()) umma_commit(&done_bar);
    __syncwarp();
  }

  // =============================== epilogue (loader warps) ========================================
  if (warp < WB_LOAD_WARPS) {
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    const int q = warp & 3, part = warp >> 2;
    const int PARTS = WB_LOAD_WARPS / 4;
    const int m = P.M == 64 ? (lane < 16 ? 16 * q + lane : -1) : 32 * q + lane;
    float* out = P.partial + (size_t)split * (P.wsize + P.cout_pad);
    const int ci = 8 * plane0 + m;
    const bool row_w = m >= 0 && ci < 4 * P.cinp4 && ci < 8 * P.kp;
    const bool row_b = !P.ones_sep && mt == P.ones_mt && m == 8 * P.ones_slot;
    if (P.ones_sep && mt == 0 && tg == 0 && warp == 0) {             // row 0 of the bias accumulator
      for (int c0 = 0; c0 < P.N; c0 += 8) {
        float a[8];
        tmem_ld8(tmem_base + (uint32_t)(ntaps * P.N + c0), a);
        if (lane == 0)
#pragma unroll
          for (int t = 0; t < 8; ++t)
            if (c0 + t < P.gp * 8 && c0 + t < P.cout_pad) out[P.wsize + c0 + t] = a[t];
      }
    }
    int chunk = 0;
    for (int tl = 0; tl < ntaps; ++tl) {
      const int tap = tap0 + tl;
      for (int c0 = 0; c0 < P.N; c0 += 8, ++chunk) {
        if (chunk % PARTS != part) continue;                         // warp-uniform
        float a[8];
        tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(tl * P.N + c0), a);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int co = c0 + t;
          if (co >= P.gp * 8 || co >= P.cout_pad) continue;
          if (row_w) out[((size_t)(tap * P.cinp4 + (ci >> 2)) * P.cout_pad + co) * 4 + (ci & 3)] = a[t];
          if (row_b && tap == 4) out[P.wsize + co] = a[t];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WB_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 :: "r"(tmem_base), "r"(P.tmem_cols));
  }
}

static int wb_pow2_cols(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

static bool wgrad_bf16_plan(WgBfParams& P, int sm_count) {
  if (P.kp <= 0 || P.gp <= 0 || P.gp * 8 > 256) return false;
  if ((long long)P.B * (P.kp > P.gp ? P.kp : P.gp) * P.H * P.W >= (1ll << 31)) return false;
  // M-tiles: the real planes plus one slot for the constant-one channel
  // When the planes fill their tiles exactly (8 planes: the last decoder layer of configs[3], 32 + 28
  // channels), the bias row would cost a whole extra tile (M = 128 instead of 64, or one more job):
  // it becomes a separate MMA with an all-ones operand instead (23 clk per K step against
  // 9 x 16 clk for the larger M).
  if (P.kp <= 8) {
    P.M = 64;
    P.slots = 8;
  } else {
    P.M = 128;
    P.slots = 16;
  }
  P.ones_sep = (P.kp % P.slots == 0) ? 1 : 0;
  const int rows_needed = P.kp + (P.ones_sep ? 0 : 1);
  P.mt_count = (rows_needed + P.slots - 1) / P.slots;
  P.ones_mt = P.ones_sep ? 0 : P.kp / P.slots;
  P.ones_slot = P.ones_sep ? -1 : P.kp % P.slots;
  P.u_alloc = rows_needed < P.slots ? rows_needed : P.slots;
  P.N = P.M == 128 ? (P.gp * 8 + 15) / 16 * 16 : P.gp * 8;
  if (P.N < 16) P.N = 16;
  if (P.N > 256) return false;
  const int tpg_max = 512 / P.N - (P.ones_sep ? 1 : 0);
  if (tpg_max < 1) return false;
  P.tg_count = (9 + tpg_max - 1) / tpg_max;
  P.tpg = (9 + P.tg_count - 1) / P.tg_count;
  P.jobs = P.mt_count * P.tg_count;
  P.tmem_cols = wb_pow2_cols((P.tpg + (P.ones_sep ? 1 : 0)) * P.N);
  // column tiling: the width (multiple of 16, at most 112: the TMA box holds <= 256 elements of
  // 8 bytes per row) that stages the fewest pixels
  int best_tw = 0;
  long long best_cost = 0;
  for (int tw = 16; tw <= 112; tw += 16) {
    const int tx = (P.W + tw - 1) / tw;
    const long long cost = (long long)tx * (tw + 2);
    if (!best_tw || cost < best_cost || (cost == best_cost && tw > best_tw)) {
      best_tw = tw;
      best_cost = cost;
    }
  }
  P.TW = best_tw;
  P.TWu = P.TW + 2;
  P.tiles_x = (P.W + P.TW - 1) / P.TW;
  const int ng = P.N / 8;
  auto stage_bytes = [&](int R) {
    const long long sbo_u = ((long long)(R + 2) * P.TWu * 16 + 127) / 128 * 128;
    return (long long)P.u_alloc * sbo_u + (long long)ng * R * P.TW * 16;
  };
  auto slack = [&](int R) {
    const long long sbo_u = ((long long)(R + 2) * P.TWu * 16 + 127) / 128 * 128;
    return (long long)(P.slots - P.u_alloc) * sbo_u;
  };
  int bestR = 0, bestS = 0;
  for (int S = 2; S >= 1 && !bestR; --S)
    for (int R = 1; R <= 16 && R <= P.H; ++R)
      if (S * stage_bytes(R) + slack(R) + 128 + 256 <= WB_SMEM_BUDGET) {
        bestR = R;
        bestS = S;
      }
  if (!bestR) return false;
  P.R = bestR;
  P.n_stages = bestS;
  if (bestS == 2 && 3 * stage_bytes(bestR) + slack(bestR) + 128 + 256 <= WB_SMEM_BUDGET) P.n_stages = 3;
  P.sbo_u = (int)(((long long)(P.R + 2) * P.TWu * 16 + 127) / 128 * 128);
  P.sbo_g = P.R * P.TW * 16;
  for (int t = 0; t < 12; ++t) P.toff[t] = t < 9 ? (t / 3) * P.TWu + (t % 3) : 0;
  P.idesc = umma_idesc_bf16(P.M, P.N, 1, 1);
  // no-swizzle descriptors (umma_desc with a zero start address): LBO = 128 B, SBO = plane stride
  P.a_hi = ((unsigned long long)(128u >> 4) << 16) | ((unsigned long long)((unsigned)P.sbo_u >> 4) << 32) | (1ull << 46);
  P.b_hi = ((unsigned long long)(128u >> 4) << 16) | ((unsigned long long)((unsigned)P.sbo_g >> 4) << 32) | (1ull << 46);
  P.u_bytes = P.u_alloc * P.sbo_u;
  P.stage_bytes = (int)stage_bytes(P.R);
  P.ones_off = (int)(((long long)P.n_stages * P.stage_bytes + slack(P.R) + 127) / 128 * 128);
  P.n_row_blocks = (P.H + P.R - 1) / P.R;
  const long long nt = (long long)P.B * P.n_row_blocks * P.tiles_x;
  if (nt > (1ll << 30)) return false;
  P.n_tiles = (int)nt;
  int splits = sm_count / P.jobs;
  if (splits < 1) splits = 1;
  if (splits > P.n_tiles) splits = P.n_tiles;
  P.splits = splits;
  return true;
}

// tensor map over a planar tensor [B][planes][H][W] of 16-byte pixels seen as f64 [..][W*2]:
// box = one plane's `rows` x `cols` pixel block
static int wb_encode_map(CUtensorMap* tm, const void* base, int B, int planes, int H, int W, int rows, int cols) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return DINCAE_ECUDA;
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  const cuuint64_t dims[4] = {(cuuint64_t)W * 2, (cuuint64_t)H, (cuuint64_t)planes, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)planes * H * W * 16};
  const cuuint32_t box[4] = {(cuuint32_t)cols * 2, (cuuint32_t)rows, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void*>(base), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_cuda_error = (int)r;
    return DINCAE_ECUDA;
  }
  return DINCAE_OK;
}

}  // namespace dincae

using namespace dincae;

extern "C" size_t dincae_conv3x3_wgrad_bf16_workspace(int cinp4, int cout_pad) {
  // at most one partial per SM (148 on B200; 160 leaves head-room)
  return align_up((size_t)160 * ((size_t)9 * cinp4 * cout_pad * 4 + cout_pad) * sizeof(float), 256);
}

extern "C" int dincae_conv3x3_wgrad_bf16(const void* src0, int c0p, int src0_half, const void* src1, int c1p,
                                         const void* g_full, const void* g_pool, const uint8_t* sign,
                                         float neg_slope, int gp, int cinp4, int cout_pad,
                                         int B, int H, int W, float* dw, float* db,
                                         void* workspace, size_t workspace_bytes, void* stream) {
  if (!src0 || !dw || !db || !workspace || c0p <= 0 || gp <= 0 || B <= 0 || H <= 0 || W <= 0) return DINCAE_EINVAL;
  if (!g_full && (!g_pool || !sign)) return DINCAE_EINVAL;
  if (gp * 8 > cout_pad || (cout_pad & 3) || cinp4 != 2 * (c0p + (src1 ? c1p : 0))) return DINCAE_EINVAL;
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  WgBfParams P = {};
  P.u0 = reinterpret_cast<const uint4*>(src0);
  P.u1 = reinterpret_cast<const uint4*>(src1);
  P.p0 = c0p; P.p1 = src1 ? c1p : 0; P.half0 = src0_half;
  P.gmode = g_full ? 0 : 1;
  P.g = reinterpret_cast<const uint4*>(g_full ? g_full : g_pool);
  P.sign = sign;
  P.slope = neg_slope;
  P.B = B; P.H = H; P.W = W; P.Hh = (H + 1) / 2; P.Wh = (W + 1) / 2; P.Wq = (W + 3) / 4;
  P.kp = P.p0 + P.p1; P.gp = gp;
  P.cinp4 = cinp4; P.cout_pad = cout_pad; P.wsize = 9 * cinp4 * cout_pad * 4;
  P.partial = reinterpret_cast<float*>(workspace);
  if (!wgrad_bf16_plan(P, sm_count)) return DINCAE_EINVAL;
  if ((size_t)P.splits * ((size_t)P.wsize + cout_pad) * sizeof(float) > workspace_bytes) return DINCAE_ENOSPC;
  CUtensorMap tmu0, tmu1, tmg;
  memset(&tmu0, 0, sizeof(tmu0));
  memset(&tmu1, 0, sizeof(tmu1));
  memset(&tmg, 0, sizeof(tmg));
  if (!P.half0) {
    const int rc = wb_encode_map(&tmu0, P.u0, B, P.p0, H, W, P.R + 2, P.TWu);
    if (rc) return rc;
    P.tma_u0 = 1;
  }
  if (P.p1 > 0) {
    const int rc = wb_encode_map(&tmu1, P.u1, B, P.p1, H, W, P.R + 2, P.TWu);
    if (rc) return rc;
    P.tma_u1 = 1;
  }
  if (P.gmode == 0) {
    const int rc = wb_encode_map(&tmg, P.g, B, P.gp, H, W, P.R, P.TW);
    if (rc) return rc;
    P.tma_g = 1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         WB_SMEM_BUDGET + 1024);
    if (e != cudaSuccess) return cuda_rc(e);
    attr_set = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)P.ones_off + 256 + 128;
  {
    static int lw = 0, pa = -1;
    if (pa < 0) {
      const char* e = getenv("DINCAE_WB_WARPS");
      lw = e ? atoi(e) : 16;
      if (lw != 4 && lw != 8 && lw != 12 && lw != 16) lw = 16;
      e = getenv("DINCAE_WB_POLL_ALL");
      pa = (e && e[0] == '0') ? 0 : 1;
    }
    P.load_warps = lw;
    P.poll_all = pa;
  }
  static long long* dbg_dev = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    const char* e = getenv("DINCAE_WB_DEBUG");
    dbg_on = (e && e[0] == '1') ? 1 : 0;
    if (dbg_on) cudaMalloc(&dbg_dev, 2048 * sizeof(long long));
  }
  P.dbg = nullptr;
  if (dbg_on) {
    cudaMemsetAsync(dbg_dev, 0, 2048 * sizeof(long long), st);
    P.dbg = dbg_dev;
  }
  launch_k(conv3x3_wgrad_bf16_kernel, P.jobs * P.splits, 32 * (P.load_warps + 1), smem, st, P, tmu0, tmu1, tmg);
  int rc = check_launch();
  if (rc) return rc;
  if (dbg_on) {
    static long long h[2048];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "wgrad_bf16 timeline CTA0: M=%d N=%d R=%d TW=%d stages=%d jobs=%d splits=%d tiles=%d\n", P.M, P.N,
            P.R, P.TW, P.n_stages, P.jobs, P.splits, P.n_tiles);
    const long long t0 = h[0];
    for (int i = 0; i < 12 && h[4 * i]; ++i)
      fprintf(stderr, "  load %2d: wait_empty %7lld..%7lld  stored %7lld  arrived %7lld | mma: wait_full %7lld..%7lld issued %7lld (%lld mmas)\n",
              i, h[4 * i] - t0, h[4 * i + 1] - t0, h[4 * i + 2] - t0, h[4 * i + 3] - t0, h[1024 + 4 * i] - t0,
              h[1024 + 4 * i + 1] - t0, h[1024 + 4 * i + 2] - t0, h[1024 + 4 * i + 3]);
  }
  int gp4 = 2 * gp;
  if (gp4 * 4 > cout_pad) gp4 = cout_pad / 4;
  return dincae_conv3x3_wgrad_reduce(workspace, P.splits, cinp4, gp4, cout_pad, dw, db, stream);
}
