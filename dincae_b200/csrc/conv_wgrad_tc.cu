// Weight + bias gradient of the 3x3 convolutions on the 5th-gen tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM).
//
//   dW[dy][dx][ci][co] = sum over pixels  U[y+dy-1][x+dx-1][ci] * G[y][x][co]
//
// is a GEMM whose reduction dimension is the PIXEL index, so both operands are needed
// "K-major" with pixels contiguous -- the transpose of the planar-4 activation layout (tf32
// MN-major descriptors are not usable without swizzle: tools/umma_decode.cu).  The loader
// warps therefore transpose on the way into shared memory (float4 of 4 channels in, 4 scalar
// stores out) into [pixel group of 4][channel group of 8][8][4] tiles: core matrices of
// 8 channels x 4 pixels (128 B) at a 144-byte pitch (SBO), pixel groups at an odd multiple of
// 16 bytes (LBO) so that the scalar stores of a warp (32 consecutive pixels of one channel)
// hit 32 different banks.  MMA rows past the real channel groups (M is 64 or 128 whatever
// Cin) alias the following pixel groups: in bounds, and their D rows are never read.
//
// Per K step (8 pixels of one image row) and per dy one MMA:
//   A = U rows  (M = input channels, start address shifted by dy rows),
//   B = three dx-shifted copies of G stacked along N (N = 3 * Cout8), written by the loader,
// so that 3 MMAs per 8 pixels produce all 9 taps.  One extra A row of constant ones gives the
// bias gradient (sum of G) as row `ones_row` of the dy = 1 accumulator.  Each CTA keeps its
// three accumulators in TMEM over all its tiles and writes ONE partial at the end;
// wgrad_reduce_kernel (conv.cu) sums the partials in fixed order (deterministic).
#include "common.cuh"
#include "conv_params.cuh"
#include "umma.cuh"
#include <math.h>

namespace dincae {

constexpr int WG_LOAD_WARPS = 20;
constexpr int WG_MMA_WARP = WG_LOAD_WARPS;
constexpr int WG_THREADS = 32 * (WG_LOAD_WARPS + 1);
constexpr int WG_STAGES = 2;
constexpr int WG_CM = 36;                 // floats per (8 channel x 4 pixel) core matrix incl. pad
constexpr int WG_MAX_ITEMS = 512;

struct WgTcParams {
  ConvSrc u;                // conv input (concat / upsample fused)
  ConvSrc g;                // gradient w.r.t. the pre-activation (direct or un-pooled)
  int B, H, W;
  int kp, gp;               // planes of U and G
  int cinp, cout_pad;       // packed weight shape [9][cinp][cout_pad][4]
  int wsize;                // 9 * cinp * cout_pad * 4
  float* partial;           // [gridDim.x][wsize + cout_pad]
  // plan
  int R, TW, n_row_blocks, n_tiles;
  int tiles_x, TWt, x_off;  // tiles per image row; image columns per tile; tile column of its first own pixel
  int cgu, cgg;             // channel groups (of 8) of U (incl. ones row) and of one G copy
  int M, N, ones_row;
  int npg_u, npg_g;         // pixel groups (of 4) in the U / G tile
  int pitch_u, pitch_g;     // floats per pixel group (all channel groups + bank pad)
  int u_floats, stage_floats;
  int n_items_u, n_items;   // loader items per stage (plane x 128-pixel chunk)
  int chunks_u, chunks_g;
  unsigned inv_tw;          // ceil(2^32 / TW): p / TW == __umulhi(p, inv_tw) for the p used here
  int tmem_cols;
};

__device__ __forceinline__ void wg_store_t(float* tile, int pitch, int plane, int p, float4 v) {
  // channels 4*plane .. 4*plane+3 of pixel p -> [p/4][plane/2][(plane&1)*4 + j][p%4]
  float* d = tile + (p >> 2) * pitch + (plane >> 1) * WG_CM + ((plane & 1) << 4) + (p & 3);
  d[0] = v.x; d[4] = v.y; d[8] = v.z; d[12] = v.w;
}

struct WgPass {
  int tile, it;
};

// tile -> image, first row, and the image column that tile column 0 stands for
// (XT = 0: one tile spans the image width -- no column arithmetic in the loader's hot loop)
template <int XT>
__device__ __forceinline__ void wg_tile(const WgTcParams& P, int tile, int& b, int& y0, int& x0) {
  if (XT == 0) {
    b = tile / P.n_row_blocks;
    y0 = (tile - b * P.n_row_blocks) * P.R;
    x0 = 0;
    return;
  }
  const int per_img = P.n_row_blocks * P.tiles_x;
  b = tile / per_img;
  const int rem = tile - b * per_img;
  const int rb = rem / P.tiles_x;
  y0 = rb * P.R;
  x0 = (rem - rb * P.tiles_x) * P.TWt - P.x_off;
}

// next pass of a warp that takes items first, first + step, ... of every `tile_stride`-th tile
__device__ __forceinline__ WgPass wg_next(WgPass c, int n_pad, int first, int step, int tile_stride) {
  c.it += step;
  if (c.it >= n_pad) {
    c.it = first;
    c.tile += tile_stride;
  }
  return c;
}

// Issue the 4 loads of one pass (item c.it of tile c.tile) of this lane.  All offsets are
// 32-bit (the plan rejects tensors of 2^31 float4 or more).
template <int GMODE, int XT>
__device__ __forceinline__ void wg_issue(const WgTcParams& P, const WgPass& c, int lane,
                                         const uint16_t* item_tab, float4 (&v)[4], unsigned (&sb)[4]) {
  const bool on = c.it < P.n_items;
  const unsigned e = on ? item_tab[c.it] : 0u;
  const int plane = (e >> 6) & 0x1ff, chunk = e & 63;
  const bool is_g = (e & 0x8000u) != 0;
  int b, y0, x0;
  wg_tile<XT>(P, c.tile, b, y0, x0);
  const ConvSrc& U = P.u;
  const ConvSrc& G = P.g;
  // plane base offsets (float4 units), resolved once per pass
  const float4* ubase;
  int ush = 0, uw = U.W;
  if (plane < U.p0) {
    ush = U.half0;
    uw = ush ? U.Wh : U.W;
    ubase = U.s0 + (b * U.p0 + plane) * (ush ? U.Hh * U.Wh : U.H * U.W);
  } else {
    ubase = U.s1 + (b * U.p1 + (plane - U.p0)) * (U.H * U.W);
  }
  const float4* gbase = G.s0 + (b * G.p0 + plane) * (GMODE == 0 ? G.H * G.W : G.Hh * G.Wh);
  const uint16_t* sbase = GMODE == 1 ? G.sign + (b * G.p0 + plane) * (G.H * G.Wq) : nullptr;
  const int npix = is_g ? P.R * P.TW : (P.R + 2) * P.TW;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int p = chunk * 128 + lane + 32 * u;
    const int r = (int)__umulhi((unsigned)p, P.inv_tw), xt = p - r * P.TW;
    const int x = x0 + xt;                        // image column (negative / >= W in the halo)
    sb[u] = 0;
    if (!is_g) {
      const int y = y0 - 1 + r;
      const bool ok = on && p < npix && (unsigned)y < (unsigned)P.H && (unsigned)x < (unsigned)P.W;
      v[u] = ldg_f4_pred(ubase + ((y >> ush) * uw + (x >> ush)), ok);
    } else {
      // a tile takes the gradient of its own TWt columns only (the halo columns stay zero), so
      // that every pixel of G enters the sum exactly once
      const int y = y0 + r;
      const bool ok = on && p < npix && y < P.H && x < P.W &&
                      (XT == 0 || (xt >= P.x_off && xt < P.x_off + P.TWt));
      if (GMODE == 0) {
        v[u] = ldg_f4_pred(gbase + (y * G.W + x), ok);
      } else {
        v[u] = ldg_f4_pred(gbase + ((y >> 1) * G.Wh + (x >> 1)), ok);
        sb[u] = ldg_u16_pred(sbase + (y * G.Wq + (x >> 2)), ok);
      }
    }
  }
}

// Convert (TF32) and store one pass into its stage: U transposed, G as three dx-shifted copies.
template <int GMODE, int XT>
__device__ __forceinline__ void wg_commit(const WgTcParams& P, const WgPass& c, int lane,
                                          const uint16_t* item_tab, float* stage,
                                          const float4 (&v)[4], const unsigned (&sb)[4]) {
  if (c.it >= P.n_items) return;
  const unsigned e = item_tab[c.it];
  const int plane = (e >> 6) & 0x1ff, chunk = e & 63;
  const bool is_g = (e & 0x8000u) != 0;
  const int TW = P.TW;
  float* Ut = stage;
  float* Gt = stage + P.u_floats;
  int b, y0, x0;
  wg_tile<XT>(P, c.tile, b, y0, x0);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int p = chunk * 128 + lane + 32 * u;
    if (!is_g) {
      if (p < (P.R + 2) * TW) wg_store_t(Ut, P.pitch_u, plane, p, to_tf32(v[u]));
    } else if (p < P.R * TW) {
      const int r = (int)__umulhi((unsigned)p, P.inv_tw), x = p - r * TW;
      float4 t = v[u];
      if (GMODE == 1) {
        const int y = y0 + r;
        t = tc_unpool(t, sb[u], x0 + x, P.W, (2 * (y >> 1) + 1 < P.H) ? 0.5f : 1.0f, P.g.slope);
      }
      t = to_tf32(t);
      // copy dx holds G[x' - dx + 1] at column x', i.e. G[x] lands at column x + dx - 1 (the
      // cells no G[x] lands on, column TW-1 of copy 0 and column 0 of copy 2, stay zero from
      // the initial clear)
      if (x >= 1) wg_store_t(Gt, P.pitch_g, plane, p - 1, t);
      wg_store_t(Gt, P.pitch_g, 2 * P.cgg + plane, p, t);
      if (x + 1 < TW) wg_store_t(Gt, P.pitch_g, 4 * P.cgg + plane, p + 1, t);
    }
  }
}

#ifdef DINCAE_TC_TIMELINE
__device__ long long wg_dbg[3 * 2048];
#define WG_MARK(code) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { const int r_ = (code) / 100 - 1; \
  if (wg_k[r_] < 1023) { wg_dbg[r_ * 2048 + 2 * wg_k[r_]] = clock64(); wg_dbg[r_ * 2048 + 2 * wg_k[r_] + 1] = (code); ++wg_k[r_]; } } } while (0)
#else
#define WG_MARK(code)
#endif

template <int GMODE, int XT>
__global__ void __launch_bounds__(WG_THREADS, 1) conv3x3_wgrad_tc_kernel(const WgTcParams P) {
#ifdef DINCAE_TC_TIMELINE
  int wg_k[3] = {0, 0, 0};
#endif
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t full_bar[WG_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[WG_STAGES];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_holder;
  __shared__ uint16_t item_tab[WG_MAX_ITEMS];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup: zero the ring, ones row, item table, barriers, TMEM ------------------
  pdl_trigger();
  for (int i = tid; i < (WG_STAGES * P.stage_floats) >> 2; i += WG_THREADS)
    reinterpret_cast<float4*>(ring)[i] = f4_zero();
  for (int i = tid; i < P.n_items; i += WG_THREADS) {
    // item = (plane, 128-pixel chunk): U items first, then G items
    int is_g = i >= P.n_items_u, k = is_g ? i - P.n_items_u : i;
    const int chunks = is_g ? P.chunks_g : P.chunks_u;
    item_tab[i] = (uint16_t)((is_g << 15) | ((k / chunks) << 6) | (k % chunks));
  }
  __syncthreads();
  {
    const int cg = P.ones_row >> 3, c8 = P.ones_row & 7;
    for (int s = 0; s < WG_STAGES; ++s)
      for (int i = tid; i < P.npg_u * 4; i += WG_THREADS)
        ring[s * P.stage_floats + (i >> 2) * P.pitch_u + cg * WG_CM + c8 * 4 + (i & 3)] = 1.0f;
  }
  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(&full_bar[s], WG_LOAD_WARPS / 2);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WG_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&tmem_holder)), "r"(P.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  pdl_wait();                 // the set-up above touches no global memory
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  const int TW = P.TW, R = P.R;

  if (warp < WG_LOAD_WARPS) {
    // =============================== transposing loader ======================================
    // Item = 128 consecutive pixels (tile-linear index p = row * TW + x) of one plane.  The
    // loader warps form two groups: group g stages the CTA's tiles g, g+2, ... into stage g,
    // so that while one group waits for its loads the other converts and stores.  Inside a
    // tile a warp's passes are software-pipelined (loads of pass n+1 in flight while pass n
    // is stored); the loads of the group's NEXT tile are only issued after the stage has
    // been published, because the proxy fence in front of the mbarrier arrive (MEMBAR) would
    // otherwise wait for them.
    constexpr int GW = WG_LOAD_WARPS / 2;                    // warps per group
    const int grp = warp / GW, wl = warp - grp * GW;
    const int n_pad = P.n_items > GW ? P.n_items : GW;       // every warp arrives per tile
    const int tstride = 2 * gridDim.x;
    float* stage = ring + grp * P.stage_floats;
    WgPass cur, nxt;
    cur.tile = blockIdx.x + grp * gridDim.x;
    cur.it = wl;
    float4 va[4], vb[4];
    unsigned sa[4], sbb[4];
    uint32_t ph = 0;
    if (cur.tile < P.n_tiles) wg_issue<GMODE, XT>(P, cur, lane, item_tab, va, sa);
    while (cur.tile < P.n_tiles) {
      // ---- pass A ---------------------------------------------------------------------------
      nxt = wg_next(cur, n_pad, wl, GW, tstride);
      bool last = nxt.tile != cur.tile;
      if (!last) wg_issue<GMODE, XT>(P, nxt, lane, item_tab, vb, sbb);
      if (warp == 0) WG_MARK(100);
      if (cur.it == wl) mbar_wait(&empty_bar[grp], ph ^ 1u);
      if (warp == 0) WG_MARK(101);
      wg_commit<GMODE, XT>(P, cur, lane, item_tab, stage, va, sa);
      if (warp == 0) WG_MARK(102);
      if (last) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[grp]);
        if (warp == 0) WG_MARK(103);
        ph ^= 1u;
        if (nxt.tile < P.n_tiles) wg_issue<GMODE, XT>(P, nxt, lane, item_tab, vb, sbb);
      }
      cur = nxt;
      if (cur.tile >= P.n_tiles) break;
      // ---- pass B (buffers swapped) ------------------------------------------------------------
      nxt = wg_next(cur, n_pad, wl, GW, tstride);
      last = nxt.tile != cur.tile;
      if (!last) wg_issue<GMODE, XT>(P, nxt, lane, item_tab, va, sa);
      if (warp == 0) WG_MARK(100);
      if (cur.it == wl) mbar_wait(&empty_bar[grp], ph ^ 1u);
      if (warp == 0) WG_MARK(101);
      wg_commit<GMODE, XT>(P, cur, lane, item_tab, stage, vb, sbb);
      if (warp == 0) WG_MARK(102);
      if (last) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[grp]);
        if (warp == 0) WG_MARK(103);
        ph ^= 1u;
        if (nxt.tile < P.n_tiles) wg_issue<GMODE, XT>(P, nxt, lane, item_tab, va, sa);
      }
      cur = nxt;
    }
  } else {
    // =============================== MMA issuer ===============================================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P.N >> 3) << 17) |
                           ((uint32_t)(P.M >> 4) << 24);
    const uint64_t a_hi = umma_desc(0u, (uint32_t)P.pitch_u * 4u, WG_CM * 4u);
    const uint64_t b_hi = umma_desc(0u, (uint32_t)P.pitch_g * 4u, WG_CM * 4u);
    const uint32_t ua = (uint32_t)(P.pitch_u >> 2), ga = (uint32_t)(P.pitch_g >> 2);   // 16-byte units
    int s = 0;
    uint32_t ph = 0;
    uint32_t first = 0;                         // 0 until the accumulators hold a value
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      WG_MARK(200);
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      WG_MARK(201);
      if (elect_one()) {
        const uint32_t u_lo = smem_u32(ring + s * P.stage_floats) >> 4;
        const uint32_t g_lo = u_lo + (uint32_t)(P.u_floats >> 2);
        // running descriptor offsets (16-byte units): a K step advances two pixel groups, a
        // row TW/4 of them; dy shifts the A start by whole rows
        const uint32_t a_row = (uint32_t)(TW >> 2) * ua, a_k = 2u * ua, b_k = 2u * ga;
        const int ksteps = R * (TW >> 3);
        uint32_t at = u_lo, bt = g_lo;
#pragma unroll 2
        for (int k = 0; k < ksteps; ++k) {
          umma_tf32(tmem_base, a_hi | (uint64_t)(at & 0x3fffu), b_hi | (uint64_t)(bt & 0x3fffu), idesc, first);
          umma_tf32(tmem_base + (uint32_t)P.N, a_hi | (uint64_t)((at + a_row) & 0x3fffu),
                    b_hi | (uint64_t)(bt & 0x3fffu), idesc, first);
          umma_tf32(tmem_base + (uint32_t)(2 * P.N), a_hi | (uint64_t)((at + 2u * a_row) & 0x3fffu),
                    b_hi | (uint64_t)(bt & 0x3fffu), idesc, first);
          first = 1u;
          at += a_k;
          bt += b_k;
        }
        umma_commit(&empty_bar[s]);
      }
      __syncwarp();
      WG_MARK(202);
      if (++s == WG_STAGES) { s = 0; ph ^= 1u; }
    }
    if (elect_one()) umma_commit(&done_bar);
    __syncwarp();
  }

  // =============================== epilogue (all loader warps) ================================
  // TMEM lane quadrant = warp % 4; the warps of a quadrant share its 8-column chunks.
  if (warp < WG_LOAD_WARPS) {
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    const int q = warp & 3, part = warp >> 2;
    constexpr int PARTS = WG_LOAD_WARPS / 4;
    const int m = P.M == 64 ? (lane < 16 ? 16 * q + lane : -1) : 32 * q + lane;
    float* out = P.partial + (size_t)blockIdx.x * (P.wsize + P.cout_pad);
    const int cout8 = P.cgg * 8;
    const bool row_w = m >= 0 && m < 4 * P.kp;
    const bool row_b = m == P.ones_row;
    int chunk = 0;
    for (int dy = 0; dy < 3; ++dy) {
      for (int dx = 0; dx < 3; ++dx) {
        const int tap = dy * 3 + dx;
        for (int c0 = 0; c0 < cout8; c0 += 8, ++chunk) {
          if (chunk % PARTS != part) continue;                       // warp-uniform
          float a[8];
          tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(dy * P.N + dx * cout8 + c0), a);
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int co = c0 + t;
            if (co >= P.gp * 4) continue;
            if (row_w) out[((size_t)(tap * P.cinp + (m >> 2)) * P.cout_pad + co) * 4 + (m & 3)] = a[t];
            if (row_b && tap == 4) out[P.wsize + co] = a[t];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WG_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 :: "r"(tmem_base), "r"(P.tmem_cols));
  }
}

static int wg_pow2_cols(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

// Shape plan; false when the layer does not fit this path (caller uses the FFMA kernel).
bool conv_wgrad_tc_plan(WgTcParams& P, int sm_count) {
  const int Cin4 = 4 * P.kp;
  if (P.kp > 31 || P.gp > 21) return false;
  P.ones_row = Cin4;                                // first row after the packed channels
  P.cgu = (Cin4 + 1 + 7) / 8;
  P.cgg = (P.gp * 4 + 7) / 8;
  P.M = P.cgu * 8 <= 64 ? 64 : 128;
  P.N = 3 * P.cgg * 8;
  if (P.M == 128 && (P.N & 15)) {                   // M = 128 needs N % 16 == 0: pad one group
    P.cgg += 1;
    P.N = 3 * P.cgg * 8;
  }
  if (P.N > 256 || 3 * P.N > 512) return false;
  if ((long long)P.B * (P.kp > P.gp ? P.kp : P.gp) * P.H * P.W >= (1ll << 31)) return false;
  P.tmem_cols = wg_pow2_cols(3 * P.N);
  // pixel-group pitch: all channel groups, padded to an odd number of 16-byte units
  P.pitch_u = (P.cgu * 9 + ((P.cgu & 1) ? 0 : 1)) * 4;
  P.pitch_g = (3 * P.cgg * 9 + (((3 * P.cgg) & 1) ? 0 : 1)) * 4;
  const int slack = 16 * WG_CM;                     // over-read of the unused MMA row groups
  // Tile = R image rows x TWt image columns.  Up to 128 columns a tile spans the image width
  // (no column halo: the three dx-shifted copies of G take care of the taps, and U is zero
  // outside the image).  Wider images are cut into column tiles whose staged block carries 4
  // halo columns on either side (a pixel group, so that K steps stay aligned): U is loaded for
  // the whole block, G only for the tile's own columns.
  auto fits = [&](int R, int TW) {
    const long long uf = (long long)((R + 2) * TW / 4) * P.pitch_u + slack;
    const long long gf = (long long)(R * TW / 4) * P.pitch_g + slack;
    const int cu = ((R + 2) * TW + 127) / 128, cg = (R * TW + 127) / 128;
    const int items = P.kp * cu + P.gp * cg;
    return (uf + gf) * 4 * WG_STAGES <= 200 * 1024 && items <= WG_MAX_ITEMS && cu <= 63;
  };
  int best = 0;
  if (P.W <= 128) {
    P.TW = (P.W + 7) / 8 * 8;
    P.TWt = P.TW;
    P.x_off = 0;
    P.tiles_x = 1;
    // rows per stage: largest R (<= 16) whose two stages fit in ~200 KB
    for (int R = 1; R <= 16 && R <= P.H; ++R)
      if (fits(R, P.TW)) best = R;
  } else {
    // the shape with the least staged traffic per useful pixel
    double best_score = 0.0;
    int best_twt = 0;
    for (int twt = 32; twt <= 128; twt += 16) {
      const int TW = twt + 8;
      for (int R = 1; R <= 16 && R <= P.H; ++R) {
        if (!fits(R, TW)) continue;
        const double score = (double)R * twt * (P.kp + P.gp) /
                             ((double)(R + 2) * TW * P.kp + (double)R * TW * P.gp);
        if (score > best_score + 1e-9) {
          best_score = score;
          best = R;
          best_twt = twt;
        }
      }
    }
    if (best) {
      P.TWt = best_twt;
      P.TW = best_twt + 8;
      P.x_off = 4;
      P.tiles_x = (P.W + best_twt - 1) / best_twt;
    }
  }
  if (!best) return false;
  P.R = best;
  P.npg_u = (P.R + 2) * P.TW / 4;
  P.npg_g = P.R * P.TW / 4;
  P.u_floats = P.npg_u * P.pitch_u + slack;
  P.stage_floats = P.u_floats + P.npg_g * P.pitch_g + slack;
  if ((long long)P.stage_floats * 4 * WG_STAGES >= (1 << 18)) return false;
  P.chunks_u = ((P.R + 2) * P.TW + 127) / 128;
  P.chunks_g = (P.R * P.TW + 127) / 128;
  P.n_items_u = P.kp * P.chunks_u;
  P.n_items = P.n_items_u + P.gp * P.chunks_g;
  P.inv_tw = (unsigned)((0x100000000ull + P.TW - 1) / P.TW);
  P.n_row_blocks = (P.H + P.R - 1) / P.R;
  if ((long long)P.B * P.n_row_blocks * P.tiles_x > (1ll << 30)) return false;
  P.n_tiles = P.B * P.n_row_blocks * P.tiles_x;
  (void)sm_count;
  return true;
}

static int wgrad_tc_blocks(const WgTcParams& P, int sm_count) {
  // one CTA per SM while every CTA still gets at least two tiles (the partial write and the
  // set-up are per CTA); the loader, not the MMA, bounds the per-tile time
  int nblk = P.n_tiles / 2;
  if (nblk > sm_count) nblk = sm_count;
  if (nblk < 1) nblk = 1;
  return nblk;
}

int wgrad_tc(const ConvSrc& u, const ConvSrc& g, int B, int H, int W, int cinp, int gp, int cout_pad,
             float* partial, size_t partial_bytes, int* nblk_out, cudaStream_t st) {
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  if (u.mode != 0) return 1;
  WgTcParams P;
  P.u = u; P.g = g; P.B = B; P.H = H; P.W = W;
  P.kp = cinp; P.gp = gp; P.cinp = cinp; P.cout_pad = cout_pad;
  P.wsize = 9 * cinp * cout_pad * 4;
  P.partial = partial;
  if (!conv_wgrad_tc_plan(P, sm_count)) return 1;
  const int nblk = wgrad_tc_blocks(P, sm_count);
  if ((size_t)nblk * ((size_t)P.wsize + cout_pad) * sizeof(float) > partial_bytes) return DINCAE_ENOSPC;
  const int gmode = g.mode;
  const int xt = P.tiles_x > 1 ? 1 : 0;
  void (*kern)(const WgTcParams) =
      gmode == 0 ? (xt ? conv3x3_wgrad_tc_kernel<0, 1> : conv3x3_wgrad_tc_kernel<0, 0>)
                 : (xt ? conv3x3_wgrad_tc_kernel<1, 1> : conv3x3_wgrad_tc_kernel<1, 0>);
  static bool attr_set[4] = {false, false, false, false};
  if (!attr_set[2 * gmode + xt]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    if (e != cudaSuccess) return cuda_rc(e);
    attr_set[2 * gmode + xt] = true;
  }
  const size_t smem = (size_t)WG_STAGES * P.stage_floats * 4;
  launch_k(kern, nblk, WG_THREADS, smem, st, P);
  *nblk_out = nblk;
  return check_launch();
}

}  // namespace dincae

extern "C" int dincae_conv3x3_wgrad_on_tensor_cores(int cinp, int gp, int B, int H, int W) {
  if (cinp <= 0 || gp <= 0 || B <= 0 || H <= 0 || W <= 0) return 0;
  dincae::WgTcParams P = {};
  P.B = B; P.H = H; P.W = W; P.kp = cinp; P.gp = gp;
  return dincae::conv_wgrad_tc_plan(P, 148) ? 1 : 0;
}

#ifdef DINCAE_TC_TIMELINE
extern "C" int dincae_debug_timeline_wg(long long* host_out, int reset) {
  if (host_out) cudaMemcpyFromSymbol(host_out, dincae::wg_dbg, sizeof(long long) * 3 * 2048);
  if (reset) {
    static long long zeros[3 * 2048];
    cudaMemcpyToSymbol(dincae::wg_dbg, zeros, sizeof(zeros));
  }
  return 3 * 1024;
}
#endif
