// 3x3 SAME convolution (forward and data gradient) as a persistent, warp-specialised
// implicit GEMM on the 5th-gen tensor cores: tcgen05.mma kind::tf32, accumulators in TMEM,
// operands in shared memory in the no-swizzle K-major canonical layout that the planar-4
// activation format maps onto directly (a pixel's 4 channels = 16 bytes = one K group).
//
// Work decomposition.  A tile is 16 rows x 8*Xb columns of one image: Xb MMA M-tiles of
// 16 rows x 8 pixels (M = 128: the 8 pixels of a row are one 128-byte core-matrix group, the
// 16 rows are the SBO stride), so that the 2x2 pooling partners of a pixel live in the same
// warp of the epilogue (lane^1, lane^8).  Tap (dy,dx) of the 3x3 stencil is a shift of the A
// descriptor's start address inside the staged (16+2) x (8*Xb+2) halo block; K runs over
// input planes (2 planes = one K=8 MMA = one pipeline stage), N over output channels.
//
// Roles (480 threads).  Warps 0-5 stage operands into a shared-memory ring: a plane whose
// source is stored at the conv's resolution is ONE TMA tensor-tile load per stage (4-D
// tensor map over [B][planes][H][W*4 floats]; out-of-bounds coordinates give the zero halo
// and the zero padding plane for free); x2-upsampled planes and un-pooled gradients
// (concat / upsample / AvgPoolGrad*LeakyReluGrad fusions) go through registers.  Weights go
// through registers (rounded to TF32).  Warp 6 issues the MMAs (one elected lane); warps
// 7-14 run the epilogue out of TMEM (bias, leaky-ReLU, sign mask, 2x2 pool / the
// data-gradient fusions) while the MMA warp already works on the next tile
// (double-buffered accumulators).
//
// TF32 contract: tcgen05 ignores the low 13 mantissa bits of an operand.  Everything the
// register path stages is rounded (rna) first; tensors that arrive by TMA must already be
// TF32-representable -- every kernel of this library that feeds a convolution rounds on
// store when the tensor-core back end is selected (conv epilogues here, the FFMA fallback,
// the batch builder, the loss head).
#include "common.cuh"
#include "conv_params.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <string.h>

namespace dincae {

constexpr int TC_LOAD_WARPS = 6;
constexpr int TC_MMA_WARP = TC_LOAD_WARPS;
constexpr int TC_EPI_WARP0 = TC_LOAD_WARPS + 1;
constexpr int TC_EPI_WARPS = 8;               // two per TMEM lane quadrant (M-tiles split by parity)
constexpr int TC_THREADS = 32 * (TC_LOAD_WARPS + 1 + TC_EPI_WARPS);
constexpr int TC_ROWS = 16;
constexpr int TC_IN_ROWS = TC_ROWS + 2;
constexpr int TC_KC = 2;                      // planes per stage (one K=8 MMA per tap)
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_A_ROWS = TC_KC * TC_IN_ROWS;
constexpr int TC_ROWS_PER_WARP = (TC_A_ROWS + TC_LOAD_WARPS - 1) / TC_LOAD_WARPS;
constexpr int TC_MAX_N = 256;                 // widest MMA N (output channels, padded to 16)
constexpr int TC_WB = 4;                      // weight float4 per loader thread and pass

struct TcPlan {
  int Xb, tiles_x, rbk, n_tiles, PWs, n_chunks, Nmma, n_stages, acc_cols, tmem_cols;
  int plane_elems;                            // smem plane stride of the A block (float4, 128-byte multiple)
  int a_elems, w_elems, stage_elems;          // float4 units
  unsigned w_div;                             // ceil(2^32 / d): idx / d == __umulhi(idx, w_div) for idx < 2^16
  int tma0, tma1;                             // source 0 / source 1 planes arrive by TMA
};

struct TcParams {
  ConvParams c;
  TcPlan m;
};

template <int MODE>
__device__ __forceinline__ void tc_load_w(float4 (&wv)[TC_WB], int base, int tid, int k0, int Kp,
                                          const ConvParams& P, const TcPlan& L) {
  constexpr int LT = 32 * TC_LOAD_WARPS;
#pragma unroll
  for (int u = 0; u < TC_WB; ++u) {
    const int idx = base + tid + u * LT;
    if (MODE == 0) {
      const int q = (int)__umulhi((unsigned)idx, L.w_div);   // idx / Nmma
      const int n = idx - q * L.Nmma;
      const int pl = q & (TC_KC - 1), tap = q >> 1;
      const bool ok = idx < L.w_elems && k0 + pl < Kp && n < P.out_planes * 4;
      wv[u] = ldg_f4_pred(reinterpret_cast<const float4*>(P.w) +
                          ((size_t)((tap * P.w_cinp + k0 + pl) * P.cout_pad + n)), ok);
    } else {
      const int col = idx & (4 * TC_KC - 1);
      const int q = idx >> 3;
      const int tap = (int)__umulhi((unsigned)q, L.w_div);   // q / (Nmma / 4)
      const int cip = q - tap * (L.Nmma >> 2);
      const int co = 4 * k0 + col;
      const bool ok = idx < L.w_elems && cip < P.out_planes && co < Kp * 4 && co < P.cout_pad;
      wv[u] = ldg_f4_pred(reinterpret_cast<const float4*>(P.w) +
                          ((size_t)(((8 - tap) * P.w_cinp + cip) * P.cout_pad + co)), ok);
    }
  }
}

template <int MODE>
__device__ __forceinline__ void tc_store_w(const float4 (&wv)[TC_WB], int base, int tid, float4* Ws,
                                           const TcPlan& L) {
  constexpr int LT = 32 * TC_LOAD_WARPS;
#pragma unroll
  for (int u = 0; u < TC_WB; ++u) {
    const int idx = base + tid + u * LT;
    if (idx >= L.w_elems) continue;
    const float4 t = to_tf32(wv[u]);
    if (MODE == 0) {
      Ws[idx] = t;
    } else {
      const int col = idx & (4 * TC_KC - 1);
      const int q = idx >> 3;
      const int tap = (int)__umulhi((unsigned)q, L.w_div);
      const int cip = q - tap * (L.Nmma >> 2);
      float* d = reinterpret_cast<float*>(Ws) + ((tap * TC_KC + (col >> 2)) * L.Nmma + 4 * cip) * 4 + (col & 3);
      d[0] = t.x; d[4] = t.y; d[8] = t.z; d[12] = t.w;
    }
  }
}

#ifdef DINCAE_TC_TIMELINE
// role timeline of CTA 0: fire-and-forget stores, one private region per role (no atomics)
__device__ long long tc_dbg[3 * 2048];
#define TC_MARK(code) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { const int r_ = (code) / 100 - 1; \
  if (tc_k[r_] < 1023) { tc_dbg[r_ * 2048 + 2 * tc_k[r_]] = clock64(); tc_dbg[r_ * 2048 + 2 * tc_k[r_] + 1] = (code); ++tc_k[r_]; } } } while (0)
#else
#define TC_MARK(code)
#endif

template <int MODE, int SRCM>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_tc_kernel(const TcParams Q, const __grid_constant__ CUtensorMap tm0,
                  const __grid_constant__ CUtensorMap tm1) {
  const ConvParams& P = Q.c;
  const TcPlan& L = Q.m;
  extern __shared__ unsigned char smem_raw[];
  // TMA destinations must be 128-byte aligned: align the ring by hand (the launch adds slack)
  float4* ring = reinterpret_cast<float4*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
  __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_holder;
  __shared__ float4 bias_s[64];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = L.n_stages;
#ifdef DINCAE_TC_TIMELINE
  int tc_k[3] = {0, 0, 0};
#endif

  pdl_trigger();
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], TC_LOAD_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], TC_EPI_WARPS);
    mbar_init(&tempty_bar[1], TC_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (L.tma0) asm volatile("prefetch.tensormap [%0];" :: "l"(&tm0) : "memory");
    if (L.tma1) asm volatile("prefetch.tensormap [%0];" :: "l"(&tm1) : "memory");
  }
  if (warp == TC_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&tmem_holder)), "r"(L.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // everything above is local to the CTA; from here on global memory written by the previous
  // kernel is read (and buffers it may still be reading are written)
  pdl_wait();
  if (MODE == 0 && tid < 64)
    bias_s[tid] = (P.bias && tid < P.out_planes) ? __ldg(reinterpret_cast<const float4*>(P.bias) + tid)
                                                 : f4_zero();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  const int PWs = L.PWs;
  const int tiles_img = L.rbk * L.tiles_x;

  if (warp < TC_LOAD_WARPS) {
    // =============================== operand loader ===========================================
    // A single warp issues ~1 dependent instruction every 5 cycles, so the per-stage
    // instruction count is what bounds this role: a TMA plane is one instruction for the whole
    // 18 x PWs halo block; register rows use offsets resolved once per tile; weights use
    // offsets resolved once per kernel.
    constexpr int LT = 32 * TC_LOAD_WARPS;
    const ConvSrc S0 = P.src;
    const int Kp = S0.p0 + S0.p1;
    const int sh0 = (SRCM == 1) ? 1 : S0.half0;
    const int ps0 = sh0 ? S0.Hh * S0.Wh : S0.H * S0.W;       // plane strides (float4)
    const int ps1 = S0.H * S0.W;
    const int pss = S0.H * S0.Wq;                            // sign-mask plane stride (u16)
    const uint32_t plane_bytes = (uint32_t)(TC_IN_ROWS * PWs) * 16u;
    int w_src[TC_WB], w_dst[TC_WB];
#pragma unroll
    for (int u = 0; u < TC_WB; ++u) {
      const int idx = tid + u * LT;
      w_src[u] = -1;
      w_dst[u] = 0;
      if (idx < L.w_elems) {
        if (MODE == 0) {
          const int q = idx / L.Nmma, n = idx - q * L.Nmma;
          const int pl = q & (TC_KC - 1), tap = q >> 1;
          if (n < P.out_planes * 4) w_src[u] = ((tap * P.w_cinp + pl) * P.cout_pad + n) | (pl << 30);
          w_dst[u] = idx * 4;
        } else {
          const int col = idx & (4 * TC_KC - 1), q = idx >> 3;
          const int np = L.Nmma >> 2;
          const int tap = q / np, cip = q - tap * np;
          if (cip < P.out_planes) w_src[u] = (((8 - tap) * P.w_cinp + cip) * P.cout_pad + col) | ((col >> 2) << 30);
          w_dst[u] = ((tap * TC_KC + (col >> 2)) * L.Nmma + 4 * cip) * 4 + (col & 3);
        }
      }
    }
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < L.n_tiles; tile += gridDim.x) {
      const int b = tile / tiles_img, rem = tile - b * tiles_img;
      const int rbi = rem / L.tiles_x, tx = rem - rbi * L.tiles_x;
      const int y0 = rbi * TC_ROWS, x0 = tx * 8 * L.Xb;
      int off0[TC_ROWS_PER_WARP], off1[TC_ROWS_PER_WARP], offs[TC_ROWS_PER_WARP];
      unsigned ok_mask = 0, edge_mask = 0;
#pragma unroll
      for (int i = 0; i < TC_ROWS_PER_WARP; ++i) {
        const int row = warp + TC_LOAD_WARPS * i;
        const int y = y0 - 1 + (row % TC_IN_ROWS);
        if (row < TC_A_ROWS && (unsigned)y < (unsigned)P.H) ok_mask |= 1u << i;
        off0[i] = (b * S0.p0 * (sh0 ? S0.Hh : S0.H) + (y >> sh0)) * (sh0 ? S0.Wh : S0.W);
        off1[i] = (b * S0.p1 * S0.H + y) * S0.W;
        offs[i] = (b * S0.p0 * S0.H + y) * S0.Wq;
        if (2 * (y >> 1) + 1 >= P.H) edge_mask |= 1u << i;       // un-pool window of one row
      }
      const int xa = x0 - 1 + lane, xb = xa + 32;
      const bool in0 = lane < PWs, in1 = lane + 32 < PWs;
      const bool xok0 = in0 && (unsigned)xa < (unsigned)P.W;
      const bool xok1 = in1 && (unsigned)xb < (unsigned)P.W;
      for (int ch = 0; ch < L.n_chunks; ++ch) {
        const int k0 = ch * TC_KC;
        float4* As = ring + (size_t)s * L.stage_elems;
        float4* Ws = As + L.a_elems;
        // which of the two planes of this chunk arrive by TMA (warp-uniform)
        bool via_tma[TC_KC];
#pragma unroll
        for (int pl = 0; pl < TC_KC; ++pl) {
          const bool first = (k0 + pl) < S0.p0 || S0.p1 == 0;
          via_tma[pl] = SRCM == 0 && (first ? L.tma0 : L.tma1);
        }
        const bool all_tma = via_tma[0] && via_tma[1];
        float4 v[TC_ROWS_PER_WARP][2];
        unsigned sb[TC_ROWS_PER_WARP][2];
        if (!all_tma) {
#pragma unroll
          for (int i = 0; i < TC_ROWS_PER_WARP; ++i) {
            const int row = warp + TC_LOAD_WARPS * i;
            const int pl = row / TC_IN_ROWS;
            const int plane = k0 + pl;
            const bool ok = ((ok_mask >> i) & 1u) && plane < Kp && !via_tma[pl];
            const bool first = plane < S0.p0;
            const float4* base = first ? S0.s0 + (off0[i] + plane * ps0) : S0.s1 + (off1[i] + (plane - S0.p0) * ps1);
            const int sh = first ? sh0 : 0;
            v[i][0] = ldg_f4_pred(base + (xa >> sh), ok && xok0);
            v[i][1] = ldg_f4_pred(base + (xb >> sh), ok && xok1);
            sb[i][0] = sb[i][1] = 0;
            if (SRCM == 1) {
              const uint16_t* sg = S0.sign + (offs[i] + plane * pss);
              sb[i][0] = ldg_u16_pred(sg + (xa >> 2), ok && xok0);
              sb[i][1] = ldg_u16_pred(sg + (xb >> 2), ok && xok1);
            }
          }
        }
        if (warp == 0) TC_MARK(103);
        // ---- weights of this K chunk; the loads of the first pass are issued before waiting
        //      for the ring slot ----------------------------------------------------------------
        float4 wv[TC_WB];
#pragma unroll
        for (int u = 0; u < TC_WB; ++u) {
          const int off = w_src[u] & 0x3fffffff, pl = (unsigned)w_src[u] >> 30;
          bool ok = w_src[u] >= 0;
          if (MODE == 0) ok = ok && k0 + pl < Kp;
          else ok = ok && (k0 + pl) < Kp && 4 * (k0 + pl) < P.cout_pad;
          wv[u] = ldg_f4_pred(reinterpret_cast<const float4*>(P.w) +
                              (off + (MODE == 0 ? k0 * P.cout_pad : 4 * k0)), ok);
        }
        if (warp == 0) TC_MARK(100);
        mbar_wait(&empty_bar[s], ph ^ 1u);          // the ring slot is free (loads already in flight)
        if (warp == 0) TC_MARK(101);
        if (SRCM == 0 && warp == 0 && (via_tma[0] || via_tma[1])) {
          if (elect_one()) {
            mbar_expect_tx(&full_bar[s], plane_bytes * ((via_tma[0] ? 1u : 0u) + (via_tma[1] ? 1u : 0u)));
#pragma unroll
            for (int pl = 0; pl < TC_KC; ++pl) {
              if (!via_tma[pl]) continue;
              const int plane = k0 + pl;
              const bool first = plane < S0.p0 || S0.p1 == 0;
              // coordinates: (float column, row, plane, image); anything outside the tensor
              // (halo, padding plane) is zero-filled by the TMA unit
              tma_load_4d(As + pl * L.plane_elems, first ? &tm0 : &tm1, (x0 - 1) * 4, y0 - 1,
                          first ? plane : plane - S0.p0, b, &full_bar[s]);
            }
          }
          __syncwarp();
        }
#pragma unroll
        for (int u = 0; u < TC_WB; ++u) {
          if (tid + u * LT < L.w_elems) {
            const float4 t = to_tf32(wv[u]);
            if (MODE == 0) {
              Ws[tid + u * LT] = t;
            } else {
              float* d = reinterpret_cast<float*>(Ws) + w_dst[u];
              d[0] = t.x; d[4] = t.y; d[8] = t.z; d[12] = t.w;
            }
          }
        }
        for (int base = TC_WB * LT; base < L.w_elems; base += TC_WB * LT) {
          tc_load_w<MODE>(wv, base, tid, k0, Kp, P, L);
          tc_store_w<MODE>(wv, base, tid, Ws, L);
        }
        if (warp == 0) TC_MARK(104);
        if (!all_tma) {
#pragma unroll
          for (int i = 0; i < TC_ROWS_PER_WARP; ++i) {
            const int row = warp + TC_LOAD_WARPS * i;
            const int pl = row / TC_IN_ROWS;
            if (row < TC_A_ROWS && !via_tma[pl]) {
              float4* dst = As + pl * L.plane_elems + (row - pl * TC_IN_ROWS) * PWs;
              float4 t0 = v[i][0], t1 = v[i][1];
              if (SRCM == 1) {
                const float inv_y = ((edge_mask >> i) & 1u) ? 1.0f : 0.5f;
                t0 = tc_unpool(t0, sb[i][0], xa, P.W, inv_y, S0.slope);
                t1 = tc_unpool(t1, sb[i][1], xb, P.W, inv_y, S0.slope);
              }
              if (in0) dst[lane] = to_tf32(t0);
              if (in1) dst[lane + 32] = to_tf32(t1);
            }
          }
        }
        if (warp == 0) TC_MARK(105);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
        if (warp == 0) TC_MARK(102);
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == TC_MMA_WARP) {
    // =============================== MMA issuer ===============================================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(L.Nmma >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    const uint64_t a_hi = umma_desc(0u, (uint32_t)L.plane_elems * 16u, (uint32_t)PWs * 16u);
    const uint64_t w_hi = umma_desc(0u, (uint32_t)L.Nmma * 16u, 128u);
    int s = 0;
    uint32_t ph = 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < L.n_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      mbar_wait(&tempty_bar[buf], (((uint32_t)lt >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(buf * L.acc_cols);
      for (int ch = 0; ch < L.n_chunks; ++ch) {
        TC_MARK(200);
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        TC_MARK(201);
        if (elect_one()) {
          const uint32_t a_lo = (smem_u32(ring + (size_t)s * L.stage_elems) >> 4);
          const uint32_t w_lo = a_lo + (uint32_t)L.a_elems;
          for (int j = 0; j < L.Xb; ++j) {
            const uint32_t d = d0 + (uint32_t)(j * L.Nmma);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t at = a_lo + (uint32_t)((tap / 3) * PWs + (tap % 3) + 8 * j);
              const uint32_t wt = w_lo + (uint32_t)(tap * TC_KC * L.Nmma);
              umma_tf32(d, a_hi | (uint64_t)(at & 0x3fffu), w_hi | (uint64_t)(wt & 0x3fffu), idesc,
                        (ch != 0 || tap != 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[s]);
          if (ch == L.n_chunks - 1) umma_commit(&tfull_bar[buf]);
        }
        __syncwarp();
        TC_MARK(202);
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // =============================== epilogue ================================================
    // Thread = one output pixel of the M-tile (TMEM lane): row g, column xi of the 16 x 8 block;
    // its 2x2 pooling partners are lanes ^1 (x) and ^8 (y).  All output offsets are resolved
    // once per tile; per plane they advance by a constant stride.
    const int q = warp & 3;                       // TMEM lane quadrant this warp may read
    const int jh = (warp - TC_EPI_WARP0) >> 2;    // which share of the M-tiles this warp takes
    const int g = 4 * q + (lane >> 3), xi = lane & 7;
    const int Hh = (P.H + 1) >> 1, Wh = (P.W + 1) >> 1, Wq = (P.W + 3) >> 2;
    const int s_full = P.H * P.W, s_half = Hh * Wh, s_sign = P.H * Wq;
    const bool store_lane = (lane & 9) == 0;
    const float slope = P.slope;
    int lt = 0;
    for (int tile = blockIdx.x; tile < L.n_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int b = tile / tiles_img, rem = tile - b * tiles_img;
      const int rbi = rem / L.tiles_x, tx = rem - rbi * L.tiles_x;
      const int y = rbi * TC_ROWS + g;
      const bool row_ok = y < P.H;
      const int xt = tx * 8 * L.Xb + xi;
      const float inv_cy = (y + 1 < P.H) ? 0.5f : 1.0f;
      int o_full, o_half, o_sign = 0, o_hi = 0;
      if (MODE == 0) {
        o_full = (b * P.out_planes * P.H + y) * P.W + xt;
        o_half = (b * P.out_planes * Hh + (y >> 1)) * Wh + (xt >> 1);
        o_sign = (b * P.out_planes * P.H + y) * Wq + (xt >> 2);
      } else {
        o_full = 0;
        o_half = (b * P.c_lo * Hh + (y >> 1)) * Wh + (xt >> 1);
        o_hi = (b * P.c_hi * P.H + y) * P.W + xt;
      }
      if (warp == TC_EPI_WARP0) TC_MARK(300);
      mbar_wait(&tfull_bar[buf], ((uint32_t)lt >> 1) & 1u);
      tc_fence_after();
      if (warp == TC_EPI_WARP0) TC_MARK(301);
      for (int j = jh; j < L.Xb; j += TC_EPI_WARPS / 4) {
        const int x = xt + 8 * j;
        const bool valid = row_ok && (x < P.W);
        const float pool_scale = inv_cy * ((x + 1 < P.W) ? 0.5f : 1.0f);
        const uint32_t t0 = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * L.acc_cols + j * L.Nmma);
        // Four planes (16 accumulator columns) per step, processed phase by phase so that the
        // independent shuffles / stores of the four planes overlap (a single warp has little
        // else to hide their latency with).
        for (int p0 = 0; p0 < P.out_planes; p0 += 4) {
          float acc[16];
          tmem_ld16(t0 + (uint32_t)(4 * p0), acc);
          float4 a[4];
#pragma unroll
          for (int qq = 0; qq < 4; ++qq)
            a[qq] = make_float4(acc[4 * qq], acc[4 * qq + 1], acc[4 * qq + 2], acc[4 * qq + 3]);
          if (MODE == 0) {
            unsigned wb[4];
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              a[qq] = f4_add(a[qq], bias_s[p0 + qq]);
              const unsigned nib = (a[qq].x > 0.f ? 1u : 0u) | (a[qq].y > 0.f ? 2u : 0u) |
                                   (a[qq].z > 0.f ? 4u : 0u) | (a[qq].w > 0.f ? 8u : 0u);
              wb[qq] = valid ? nib << (4 * (lane & 3)) : 0u;
              a[qq].x = a[qq].x > 0.f ? a[qq].x : slope * a[qq].x;
              a[qq].y = a[qq].y > 0.f ? a[qq].y : slope * a[qq].y;
              a[qq].z = a[qq].z > 0.f ? a[qq].z : slope * a[qq].z;
              a[qq].w = a[qq].w > 0.f ? a[qq].w : slope * a[qq].w;
              if (!valid) a[qq] = f4_zero();
            }
            if (P.out_sign) {
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) wb[qq] |= __shfl_xor_sync(0xffffffffu, wb[qq], 1);
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) wb[qq] |= __shfl_xor_sync(0xffffffffu, wb[qq], 2);
#pragma unroll
              for (int qq = 0; qq < 4; ++qq)
                if ((lane & 3) == 0 && valid && p0 + qq < P.out_planes)
                  P.out_sign[o_sign + (p0 + qq) * s_sign + 2 * j] = (uint16_t)wb[qq];
            }
            if (P.out_full) {
#pragma unroll
              for (int qq = 0; qq < 4; ++qq)
                if (valid && p0 + qq < P.out_planes)
                  P.out_full[o_full + (p0 + qq) * s_full + 8 * j] = to_tf32(a[qq]);
            }
            if (P.out_pool) {
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) {
                a[qq].x += __shfl_xor_sync(0xffffffffu, a[qq].x, 1);
                a[qq].y += __shfl_xor_sync(0xffffffffu, a[qq].y, 1);
                a[qq].z += __shfl_xor_sync(0xffffffffu, a[qq].z, 1);
                a[qq].w += __shfl_xor_sync(0xffffffffu, a[qq].w, 1);
              }
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) {
                a[qq].x += __shfl_xor_sync(0xffffffffu, a[qq].x, 8);
                a[qq].y += __shfl_xor_sync(0xffffffffu, a[qq].y, 8);
                a[qq].z += __shfl_xor_sync(0xffffffffu, a[qq].z, 8);
                a[qq].w += __shfl_xor_sync(0xffffffffu, a[qq].w, 8);
              }
#pragma unroll
              for (int qq = 0; qq < 4; ++qq)
                if (store_lane && valid && p0 + qq < P.out_planes)
                  P.out_pool[o_half + (p0 + qq) * s_half + 4 * j] = to_tf32(f4_scale(a[qq], pool_scale));
            }
          } else {
            // planes below c_lo: gradient of the x2 nearest upsample (2x2 sum) times the previous
            // activation's slope; planes from c_lo on: skip gradient at full resolution
            float4 h[4], side[4];
            // side inputs of the four planes (previous activation for its slope, or the skip
            // gradient to add) are requested first: their latency hides behind the shuffles
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              const int p = p0 + qq;
              h[qq] = valid ? a[qq] : f4_zero();
              const bool lo = p < P.c_lo;
              const float4* sp = lo ? P.prev_act + (o_half + p * s_half + 4 * j)
                                    : P.dx_add + (o_hi + (p - P.c_lo) * s_full + 8 * j);
              const bool want = lo ? (P.prev_act != nullptr && store_lane && valid)
                                   : (P.dx_add != nullptr && valid && p < P.out_planes);
              side[qq] = ldg_f4_pred(sp, want);
            }
            if (p0 < P.c_lo) {
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) {
                h[qq].x += __shfl_xor_sync(0xffffffffu, h[qq].x, 1);
                h[qq].y += __shfl_xor_sync(0xffffffffu, h[qq].y, 1);
                h[qq].z += __shfl_xor_sync(0xffffffffu, h[qq].z, 1);
                h[qq].w += __shfl_xor_sync(0xffffffffu, h[qq].w, 1);
              }
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) {
                h[qq].x += __shfl_xor_sync(0xffffffffu, h[qq].x, 8);
                h[qq].y += __shfl_xor_sync(0xffffffffu, h[qq].y, 8);
                h[qq].z += __shfl_xor_sync(0xffffffffu, h[qq].z, 8);
                h[qq].w += __shfl_xor_sync(0xffffffffu, h[qq].w, 8);
              }
            }
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              const int p = p0 + qq;
              if (p < P.c_lo) {
                if (store_lane && valid) {
                  const int at = o_half + p * s_half + 4 * j;
                  float4 t = h[qq];
                  if (P.prev_act) {
                    const float4 pa = side[qq];
                    t.x *= pa.x > 0.f ? 1.0f : P.prev_slope;
                    t.y *= pa.y > 0.f ? 1.0f : P.prev_slope;
                    t.z *= pa.z > 0.f ? 1.0f : P.prev_slope;
                    t.w *= pa.w > 0.f ? 1.0f : P.prev_slope;
                  }
                  P.g_prev[at] = to_tf32(t);
                }
              } else if (p < P.out_planes && valid) {
                const int at = o_hi + (p - P.c_lo) * s_full + 8 * j;
                float4 t = a[qq];
                if (P.dx_add) t = f4_add(t, side[qq]);
                P.dx_hi[at] = to_tf32(t);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      if (warp == TC_EPI_WARP0) TC_MARK(302);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 :: "r"(tmem_base), "r"(L.tmem_cols));
  }
}

static int tc_pow2_cols(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

// Tile shape, pipeline depth and TMEM budget for one layer.  Returns false when the shape
// does not fit the tensor-core path (the caller then uses the FFMA kernel).
bool conv_tc_plan(int mode, int Kp, int out_planes, int B, int H, int W, TcPlan& L) {
  L.Nmma = (out_planes * 4 + 15) / 16 * 16;
  if (L.Nmma > TC_MAX_N || out_planes > 64) return false;
  int tiles_x = (W + 55) / 56;
  int Xb;
  for (;; ++tiles_x) {
    const int wt = (W + tiles_x - 1) / tiles_x;
    Xb = (wt + 7) / 8;
    if (Xb <= 7 && 2 * Xb * L.Nmma <= 512) break;
    if (Xb <= 1) return false;
  }
  L.Xb = Xb;
  L.tiles_x = (W + 8 * Xb - 1) / (8 * Xb);
  L.rbk = (H + TC_ROWS - 1) / TC_ROWS;
  if ((long long)B * (Kp > out_planes ? Kp : out_planes) * H * W >= (1ll << 31)) return false;   // int offsets
  const long long nt = (long long)B * L.rbk * L.tiles_x;
  if (nt > (1ll << 30)) return false;
  L.n_tiles = (int)nt;
  L.PWs = 8 * Xb + 2;
  L.n_chunks = (Kp + TC_KC - 1) / TC_KC;
  L.plane_elems = (TC_IN_ROWS * L.PWs + 7) / 8 * 8;
  L.a_elems = TC_KC * L.plane_elems;
  L.w_elems = 9 * TC_KC * L.Nmma;
  {
    const unsigned d = mode == 0 ? (unsigned)L.Nmma : (unsigned)(L.Nmma / 4);
    L.w_div = (unsigned)((0x100000000ull + d - 1) / d);
  }
  L.stage_elems = (L.a_elems + L.w_elems + 7) / 8 * 8;
  int ns = (200 * 1024) / (L.stage_elems * 16);
  if (ns > TC_MAX_STAGES) ns = TC_MAX_STAGES;
  if (ns < 2) return false;
  L.n_stages = ns;
  L.tma0 = L.tma1 = 0;
  L.acc_cols = Xb * L.Nmma;
  L.tmem_cols = tc_pow2_cols(2 * L.acc_cols);
  return true;
}

// 4-D tensor map over a planar-4 tensor [B][planes][H][W] of float4, seen as f32
// [B][planes][H][W*4]; the box is one plane's (16+2) x PWs halo block.
static int tc_encode_map(CUtensorMap* tm, const void* base, int B, int planes, int H, int W, int PWs) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return DINCAE_ECUDA;
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  const cuuint64_t dims[4] = {(cuuint64_t)W * 4, (cuuint64_t)H, (cuuint64_t)planes, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)planes * H * W * 16};
  const cuuint32_t box[4] = {(cuuint32_t)PWs * 4, (cuuint32_t)TC_IN_ROWS, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_cuda_error = (int)r;
    return DINCAE_ECUDA;
  }
  return DINCAE_OK;
}

int launch_conv_tc(int mode, const ConvParams& P, cudaStream_t st) {
  TcParams Q;
  Q.c = P;
  if (!conv_tc_plan(mode, P.src.planes(), P.out_planes, P.B, P.H, P.W, Q.m)) return 1;
  static int sm_count = 0;
  static bool attr_set[3] = {false, false, false};
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  const int srcm = P.src.mode;
  const int which = mode == 0 ? 0 : 1 + srcm;
  if (mode == 0 && srcm != 0) return 1;
  // sources stored at this resolution arrive by TMA (row pitch W*16 bytes is always a
  // multiple of 16; the base must be 16-byte aligned, which float4 tensors are)
  CUtensorMap tm0, tm1;
  memset(&tm0, 0, sizeof(tm0));
  memset(&tm1, 0, sizeof(tm1));
  if (srcm == 0 && (size_t)P.W * 4 <= 0xffffffffull) {
    if (!P.src.half0 && P.src.p0 > 0 && ((uintptr_t)P.src.s0 & 15) == 0) {
      const int rc = tc_encode_map(&tm0, P.src.s0, P.B, P.src.p0, P.H, P.W, Q.m.PWs);
      if (rc) return rc;
      Q.m.tma0 = 1;
    }
    if (P.src.p1 > 0 && ((uintptr_t)P.src.s1 & 15) == 0) {
      const int rc = tc_encode_map(&tm1, P.src.s1, P.B, P.src.p1, P.H, P.W, Q.m.PWs);
      if (rc) return rc;
      Q.m.tma1 = 1;
    }
  }
  const size_t smem = (size_t)Q.m.n_stages * Q.m.stage_elems * 16 + 128;
  if (!attr_set[which]) {
    cudaError_t e =
        which == 0 ? cudaFuncSetAttribute(conv3x3_tc_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024)
        : which == 1 ? cudaFuncSetAttribute(conv3x3_tc_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024)
                     : cudaFuncSetAttribute(conv3x3_tc_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    if (e != cudaSuccess) return cuda_rc(e);
    attr_set[which] = true;
  }
  const int grid = Q.m.n_tiles < sm_count ? Q.m.n_tiles : sm_count;
  if (which == 0)
    launch_k(conv3x3_tc_kernel<0, 0>, grid, TC_THREADS, smem, st, Q, tm0, tm1);
  else if (which == 1)
    launch_k(conv3x3_tc_kernel<1, 0>, grid, TC_THREADS, smem, st, Q, tm0, tm1);
  else
    launch_k(conv3x3_tc_kernel<1, 1>, grid, TC_THREADS, smem, st, Q, tm0, tm1);
  return check_launch();
}

}  // namespace dincae

#ifdef DINCAE_TC_TIMELINE
extern "C" int dincae_debug_timeline(long long* host_out, int max_pairs, int reset) {
  // returns 3 regions of 1024 (clock, code) pairs; unused entries are zero
  if (host_out && max_pairs >= 3 * 1024) cudaMemcpyFromSymbol(host_out, dincae::tc_dbg, sizeof(long long) * 3 * 2048);
  if (reset) {
    static long long zeros[3 * 2048];
    cudaMemcpyToSymbol(dincae::tc_dbg, zeros, sizeof(zeros));
  }
  return 3 * 1024;
}
#endif
