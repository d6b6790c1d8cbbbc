// bf16 variant of the 3x3 SAME convolution (forward and data gradient) for BASELINE configs[3]
// (512x512, encoder [32,48,72,108,162]): tcgen05.mma kind::f16 with bf16 operands, fp32
// accumulators in TMEM, fp32 master weights converted once per step (dincae_conv_weights_bf16).
//
// Same decomposition as conv_tc.cu (persistent CTA, tile = 16 rows x 8*Xb columns, M = 128 =
// 16 rows x 8 pixels, a 3x3 tap = a shift of the A descriptor inside the staged halo block),
// because a pixel of a planar-8 bf16 plane is the same 16 bytes as a pixel of a planar-4 fp32
// plane: K-major no-swizzle core matrices of 8 pixels x 16 bytes.  One pipeline stage = 2
// planes = 16 input channels = ONE K=16 MMA per tap and M-tile, i.e. half the MMAs and half
// the shared-memory operand bytes of the TF32 kernel for the same channels.
//
// Differences from conv_tc.cu: weights arrive pre-packed (bf16, per K chunk one contiguous
// block [9 taps][2 planes][N][8]) by one 1-D bulk copy per stage -- no register staging, no
// transposition in the data gradient (a second pre-packed copy); activations stored at the
// conv's resolution arrive by TMA; only x2-upsampled planes (decoder) and un-pooled gradients
// (encoder backward) go through registers.  Sign masks are one byte per pixel and plane.
#include "bf16.cuh"
#include "conv_params.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace dincae {

constexpr int BF_LOAD_WARPS = 6;
constexpr int BF_MMA_WARP = BF_LOAD_WARPS;
constexpr int BF_EPI_WARP0 = BF_LOAD_WARPS + 1;
constexpr int BF_EPI_WARPS = 16;             // four per TMEM lane quadrant: the epilogue is latency-bound
constexpr int BF_THREADS = 32 * (BF_LOAD_WARPS + 1 + BF_EPI_WARPS);
constexpr int BF_ROWS = 16;
constexpr int BF_IN_ROWS = BF_ROWS + 2;
constexpr int BF_KC = 2;                      // planes (of 8 channels) per stage = one K=16 MMA per tap
constexpr int BF_MAX_STAGES = 8;
constexpr int BF_A_ROWS = BF_KC * BF_IN_ROWS;
constexpr int BF_ROWS_PER_WARP = (BF_A_ROWS + BF_LOAD_WARPS - 1) / BF_LOAD_WARPS;
constexpr int BF_MAX_N = 256;

struct BfPlan {
  int Xb, tiles_x, rbk, n_tiles, PWs, n_chunks, Nmma, n_stages, acc_cols, tmem_cols;
  int plane_elems;                            // smem plane stride of the A block (16-byte units, 128-byte multiple)
  int a_elems, w_elems, stage_elems;          // 16-byte units
  int tma0, tma1;
  int tma_only;                               // every input plane arrives by TMA: one loader warp does it all
};

struct BfConvParams {
  // logical input = concat(s0 [p0 planes, optionally half resolution], s1 [p1 planes]), or
  // (srcm 1) the un-pooled gradient s0[y>>1][x>>1] / window * lrelu'(sign)
  const uint4* s0;
  const uint4* s1;
  const uint8_t* sign;     // [B][p0][H][4*Wq] bytes, bit j = channel lane j positive
  int p0, p1, half0;
  float src_slope;
  int B, H, W, Hh, Wh, Wq;
  const uint4* w;          // pre-packed chunks [n_chunks][9][2][Nmma][8 bf16]
  const float* bias;       // forward: [>= 8*out_planes] floats (may be null)
  int out_planes;          // planes of 8 output channels
  float slope;
  // forward outputs (any may be null)
  uint4* out_full;
  uint4* out_pool;
  uint32_t* out_sign;      // the same byte array written 4 pixels at a time
  float4* out_f32;         // planar-4 fp32 copy of the activated output, out_f32_planes planes
  int out_f32_planes;
  // data gradient
  int c_lo, c_hi;
  const uint4* prev_act;
  float prev_slope;
  uint4* g_prev;
  uint4* dx_hi;
  const uint4* dx_add;
  long long* dbg;          // DINCAE_BF_DEBUG=1: role timeline of CTA 0 ([3][256][4] clock64 values)
};

struct BfParams {
  BfConvParams c;
  BfPlan m;
};

template <int MODE, int SRCM>
__global__ void __launch_bounds__(BF_THREADS, 1)
conv3x3_bf16_kernel(const BfParams Q, const __grid_constant__ CUtensorMap tm0,
                    const __grid_constant__ CUtensorMap tm1) {
  const BfConvParams& P = Q.c;
  const BfPlan& L = Q.m;
  extern __shared__ unsigned char smem_raw[];
  uint4* ring = reinterpret_cast<uint4*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
  __shared__ __align__(8) uint64_t full_bar[BF_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[BF_MAX_STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) float bias_s[BF_MAX_N + 8];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = L.n_stages;

  pdl_trigger();
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], L.tma_only ? 1 : BF_LOAD_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], BF_EPI_WARPS);
    mbar_init(&tempty_bar[1], BF_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (L.tma0) asm volatile("prefetch.tensormap [%0];" :: "l"(&tm0) : "memory");
    if (L.tma1) asm volatile("prefetch.tensormap [%0];" :: "l"(&tm1) : "memory");
  }
  if (warp == BF_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&tmem_holder)), "r"(L.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  pdl_wait();
  if (MODE == 0)
    for (int i = tid; i < BF_MAX_N + 8; i += BF_THREADS)
      bias_s[i] = (P.bias && i < 8 * P.out_planes) ? __ldg(P.bias + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  const int PWs = L.PWs;
  const int tiles_img = L.rbk * L.tiles_x;

  if (warp < BF_LOAD_WARPS) {
    // =============================== operand loader ===========================================
    // (with every plane on the TMA path warps 1.. have nothing to stage and stay out of the way)
    if (!L.tma_only || warp == 0) {
    const int Kp = P.p0 + P.p1;
    const int sh0 = (SRCM == 1) ? 1 : P.half0;
    const int ps0 = sh0 ? P.Hh * P.Wh : P.H * P.W;          // plane strides (pixels)
    const int ps1 = P.H * P.W;
    const int sw = 4 * P.Wq;                                 // sign row pitch (bytes)
    const int pss = P.H * sw;
    const uint32_t plane_bytes = (uint32_t)(BF_IN_ROWS * PWs) * 16u;
    const uint32_t w_bytes = (uint32_t)L.w_elems * 16u;
    int s = 0;
    uint32_t ph = 0;
    int dbg_n = 0;
    for (int tile = blockIdx.x; tile < L.n_tiles; tile += gridDim.x) {
      const int b = tile / tiles_img, rem = tile - b * tiles_img;
      const int rbi = rem / L.tiles_x, tx = rem - rbi * L.tiles_x;
      const int y0 = rbi * BF_ROWS, x0 = tx * 8 * L.Xb;
      int off0[BF_ROWS_PER_WARP], off1[BF_ROWS_PER_WARP], offs[BF_ROWS_PER_WARP];
      unsigned ok_mask = 0, edge_mask = 0;
#pragma unroll
      for (int i = 0; i < BF_ROWS_PER_WARP; ++i) {
        const int row = warp + BF_LOAD_WARPS * i;
        const int y = y0 - 1 + (row % BF_IN_ROWS);
        if (row < BF_A_ROWS && (unsigned)y < (unsigned)P.H) ok_mask |= 1u << i;
        off0[i] = (b * P.p0 * (sh0 ? P.Hh : P.H) + (y >> sh0)) * (sh0 ? P.Wh : P.W);
        off1[i] = (b * P.p1 * P.H + y) * P.W;
        offs[i] = (b * P.p0 * P.H + y) * sw;
        if (2 * (y >> 1) + 1 >= P.H) edge_mask |= 1u << i;
      }
      const int xa = x0 - 1 + lane, xb = xa + 32;
      const bool in0 = lane < PWs, in1 = lane + 32 < PWs;
      const bool xok0 = in0 && (unsigned)xa < (unsigned)P.W;
      const bool xok1 = in1 && (unsigned)xb < (unsigned)P.W;
      for (int ch = 0; ch < L.n_chunks; ++ch) {
        const int k0 = ch * BF_KC;
        uint4* As = ring + (size_t)s * L.stage_elems;
        uint4* Ws = As + L.a_elems;
        bool via_tma[BF_KC];
#pragma unroll
        for (int pl = 0; pl < BF_KC; ++pl) {
          const bool first = (k0 + pl) < P.p0 || P.p1 == 0;
          via_tma[pl] = SRCM == 0 && (first ? L.tma0 : L.tma1);
        }
        const bool all_tma = via_tma[0] && via_tma[1];
        uint4 v[BF_ROWS_PER_WARP][2];
        unsigned sb[BF_ROWS_PER_WARP][2];
        if (!all_tma) {
#pragma unroll
          for (int i = 0; i < BF_ROWS_PER_WARP; ++i) {
            const int row = warp + BF_LOAD_WARPS * i;
            const int pl = row / BF_IN_ROWS;
            const int plane = k0 + pl;
            const bool ok = ((ok_mask >> i) & 1u) && plane < Kp && !via_tma[pl];
            const bool first = plane < P.p0;
            const uint4* base = first ? P.s0 + (off0[i] + plane * ps0) : P.s1 + (off1[i] + (plane - P.p0) * ps1);
            const int sh = first ? sh0 : 0;
            v[i][0] = ldg_u4_pred(base + (xa >> sh), ok && xok0);
            v[i][1] = ldg_u4_pred(base + (xb >> sh), ok && xok1);
            sb[i][0] = sb[i][1] = 0;
            if (SRCM == 1) {
              const uint8_t* sg = P.sign + (offs[i] + plane * pss);
              sb[i][0] = ldg_u8_pred(sg + xa, ok && xok0);
              sb[i][1] = ldg_u8_pred(sg + xb, ok && xok1);
            }
          }
        }
        long long t_a = 0, t_b = 0;
        if (P.dbg) t_a = clock64();
        mbar_wait(&empty_bar[s], ph ^ 1u);
        if (P.dbg) t_b = clock64();
        if (warp == 0) {
          if (elect_one()) {
            mbar_expect_tx(&full_bar[s], w_bytes + plane_bytes * ((via_tma[0] ? 1u : 0u) + (via_tma[1] ? 1u : 0u)));
            bulk_g2s(Ws, P.w + (size_t)ch * L.w_elems, w_bytes, &full_bar[s]);
#pragma unroll
            for (int pl = 0; pl < BF_KC; ++pl) {
              if (!via_tma[pl]) continue;
              const int plane = k0 + pl;
              const bool first = plane < P.p0 || P.p1 == 0;
              tma_load_4d(As + pl * L.plane_elems, first ? &tm0 : &tm1, (x0 - 1) * 4, y0 - 1,
                          first ? plane : plane - P.p0, b, &full_bar[s]);
            }
          }
          __syncwarp();
        }
        if (!all_tma) {
#pragma unroll
          for (int i = 0; i < BF_ROWS_PER_WARP; ++i) {
            const int row = warp + BF_LOAD_WARPS * i;
            const int pl = row / BF_IN_ROWS;
            if (row < BF_A_ROWS && !via_tma[pl]) {
              uint4* dst = As + pl * L.plane_elems + (row - pl * BF_IN_ROWS) * PWs;
              uint4 t0 = v[i][0], t1 = v[i][1];
              if (SRCM == 1) {
                const float inv_y = ((edge_mask >> i) & 1u) ? 1.0f : 0.5f;
                t0 = bf8_unpool(t0, sb[i][0], inv_y * ((2 * (xa >> 1) + 1 < P.W) ? 0.5f : 1.0f), P.src_slope);
                t1 = bf8_unpool(t1, sb[i][1], inv_y * ((2 * (xb >> 1) + 1 < P.W) ? 0.5f : 1.0f), P.src_slope);
              }
              if (in0) dst[lane] = t0;
              if (in1) dst[lane + 32] = t1;
            }
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
        if (P.dbg && blockIdx.x == 0 && warp == 0 && lane == 0 && dbg_n < 256) {
          long long* d = P.dbg + 4 * dbg_n++;
          d[0] = t_a; d[1] = t_b; d[2] = clock64(); d[3] = tile;
        }
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
    }
  } else if (warp == BF_MMA_WARP) {
    // =============================== MMA issuer ===============================================
    const uint32_t idesc = umma_idesc_bf16(128, L.Nmma, 0, 0);
    const uint64_t a_hi = umma_desc(0u, (uint32_t)L.plane_elems * 16u, (uint32_t)PWs * 16u);
    const uint64_t w_hi = umma_desc(0u, (uint32_t)L.Nmma * 16u, 128u);
    int s = 0;
    uint32_t ph = 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < L.n_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      long long t_a = 0, t_b = 0, t_c = 0;
      if (P.dbg) t_a = clock64();
      mbar_wait(&tempty_bar[buf], (((uint32_t)lt >> 1) & 1u) ^ 1u);
      if (P.dbg) t_b = clock64();
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(buf * L.acc_cols);
      for (int ch = 0; ch < L.n_chunks; ++ch) {
        mbar_wait(&full_bar[s], ph);
        if (P.dbg && ch == 0) t_c = clock64();
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = (smem_u32(ring + (size_t)s * L.stage_elems) >> 4);
          const uint32_t w_lo = a_lo + (uint32_t)L.a_elems;
          for (int j = 0; j < L.Xb; ++j) {
            const uint32_t d = d0 + (uint32_t)(j * L.Nmma);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t at = a_lo + (uint32_t)((tap / 3) * PWs + (tap % 3) + 8 * j);
              const uint32_t wt = w_lo + (uint32_t)(tap * BF_KC * L.Nmma);
              umma_f16(d, a_hi | (uint64_t)(at & 0x3fffu), w_hi | (uint64_t)(wt & 0x3fffu), idesc,
                       (ch != 0 || tap != 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[s]);
          if (ch == L.n_chunks - 1) umma_commit(&tfull_bar[buf]);
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1u; }
      }
      if (P.dbg && blockIdx.x == 0 && lane == 0 && lt < 256) {
        long long* d = P.dbg + 1024 + 4 * lt;
        d[0] = t_a; d[1] = t_b; d[2] = t_c; d[3] = clock64();
      }
    }
  } else {
    // =============================== epilogue ================================================
    // thread = one output pixel of the M-tile (TMEM lane): row g, column xi of the 16 x 8 block;
    // 2x2 pooling partners are lanes ^1 (x) and ^8 (y); 16 accumulator columns = 2 planes
    const int q = warp & 3;
    const int jh = (warp - BF_EPI_WARP0) >> 2;
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
      const int y = rbi * BF_ROWS + g;
      const bool row_ok = y < P.H;
      const int xt = tx * 8 * L.Xb + xi;
      const float inv_cy = (y + 1 < P.H) ? 0.5f : 1.0f;
      int o_full, o_half, o_sign = 0, o_hi = 0, o_f32 = 0;
      if (MODE == 0) {
        o_full = (b * P.out_planes * P.H + y) * P.W + xt;
        o_half = (b * P.out_planes * Hh + (y >> 1)) * Wh + (xt >> 1);
        o_sign = (b * P.out_planes * P.H + y) * Wq + (xt >> 2);
        o_f32 = (b * P.out_f32_planes * P.H + y) * P.W + xt;
      } else {
        o_full = 0;
        o_half = (b * P.c_lo * Hh + (y >> 1)) * Wh + (xt >> 1);
        o_hi = (b * P.c_hi * P.H + y) * P.W + xt;
      }
      long long t_a = 0, t_b = 0;
      if (P.dbg) t_a = clock64();
      mbar_wait(&tfull_bar[buf], ((uint32_t)lt >> 1) & 1u);
      if (P.dbg) t_b = clock64();
      tc_fence_after();
      // work units = (M-tile j, pair of output planes), dealt round-robin to the warps of a quadrant
      const int npair = (P.out_planes + 1) >> 1;
      for (int unit = jh; unit < L.Xb * npair; unit += BF_EPI_WARPS / 4) {
        const int j = unit / npair, p0 = 2 * (unit - j * npair);
        const int x = xt + 8 * j;
        const bool valid = row_ok && (x < P.W);
        const float pool_scale = inv_cy * ((x + 1 < P.W) ? 0.5f : 1.0f);
        const uint32_t t0 = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * L.acc_cols + j * L.Nmma);
        {
          float acc[16];
          tmem_ld16(t0 + (uint32_t)(8 * p0), acc);
          if (MODE == 0) {
            unsigned wb[2];
#pragma unroll
            for (int qq = 0; qq < 2; ++qq) {
              unsigned bits = 0;
              const float4 b0 = reinterpret_cast<const float4*>(bias_s)[2 * (p0 + qq)];
              const float4 b1 = reinterpret_cast<const float4*>(bias_s)[2 * (p0 + qq) + 1];
              const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                float a = acc[8 * qq + c] + bv[c];
                bits |= (a > 0.f ? 1u : 0u) << c;
                a = a > 0.f ? a : slope * a;
                acc[8 * qq + c] = valid ? a : 0.f;
              }
              wb[qq] = valid ? bits << (8 * (lane & 3)) : 0u;
            }
            if (P.out_sign) {
#pragma unroll
              for (int qq = 0; qq < 2; ++qq) wb[qq] |= __shfl_xor_sync(0xffffffffu, wb[qq], 1);
#pragma unroll
              for (int qq = 0; qq < 2; ++qq) wb[qq] |= __shfl_xor_sync(0xffffffffu, wb[qq], 2);
#pragma unroll
              for (int qq = 0; qq < 2; ++qq)
                if ((lane & 3) == 0 && valid && p0 + qq < P.out_planes)
                  P.out_sign[o_sign + (p0 + qq) * s_sign + 2 * j] = wb[qq];
            }
            if (P.out_full) {
#pragma unroll
              for (int qq = 0; qq < 2; ++qq)
                if (valid && p0 + qq < P.out_planes) {
                  float f[8];
#pragma unroll
                  for (int c = 0; c < 8; ++c) f[c] = acc[8 * qq + c];
                  P.out_full[o_full + (p0 + qq) * s_full + 8 * j] = bf8_pack(f);
                }
            }
            if (P.out_f32) {
#pragma unroll
              for (int p4 = 0; p4 < 4; ++p4)
                if (valid && 2 * p0 + p4 < P.out_f32_planes)
                  P.out_f32[o_f32 + (2 * p0 + p4) * s_full + 8 * j] =
                      make_float4(acc[4 * p4], acc[4 * p4 + 1], acc[4 * p4 + 2], acc[4 * p4 + 3]);
            }
            if (P.out_pool) {
#pragma unroll
              for (int c = 0; c < 16; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
#pragma unroll
              for (int c = 0; c < 16; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 8);
#pragma unroll
              for (int qq = 0; qq < 2; ++qq)
                if (store_lane && valid && p0 + qq < P.out_planes) {
                  float f[8];
#pragma unroll
                  for (int c = 0; c < 8; ++c) f[c] = acc[8 * qq + c] * pool_scale;
                  P.out_pool[o_half + (p0 + qq) * s_half + 4 * j] = bf8_pack(f);
                }
            }
          } else {
            // planes below c_lo: gradient of the x2 nearest upsample (2x2 sum) times the previous
            // activation's slope; planes from c_lo on: skip gradient at full resolution
            uint4 side[2];
#pragma unroll
            for (int qq = 0; qq < 2; ++qq) {
              const int p = p0 + qq;
              const bool lo = p < P.c_lo;
              const uint4* sp = lo ? P.prev_act + (o_half + p * s_half + 4 * j)
                                   : P.dx_add + (o_hi + (p - P.c_lo) * s_full + 8 * j);
              const bool want = lo ? (P.prev_act != nullptr && store_lane && valid)
                                   : (P.dx_add != nullptr && valid && p < P.out_planes);
              side[qq] = ldg_u4_pred(sp, want);
            }
            float h[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) h[c] = valid ? acc[c] : 0.f;
            if (p0 < P.c_lo) {
#pragma unroll
              for (int c = 0; c < 16; ++c) h[c] += __shfl_xor_sync(0xffffffffu, h[c], 1);
#pragma unroll
              for (int c = 0; c < 16; ++c) h[c] += __shfl_xor_sync(0xffffffffu, h[c], 8);
            }
#pragma unroll
            for (int qq = 0; qq < 2; ++qq) {
              const int p = p0 + qq;
              float f[8], sd[8];
              bf8_unpack(side[qq], sd);
              if (p < P.c_lo) {
                if (store_lane && valid) {
#pragma unroll
                  for (int c = 0; c < 8; ++c) {
                    f[c] = h[8 * qq + c];
                    if (P.prev_act) f[c] *= sd[c] > 0.f ? 1.0f : P.prev_slope;
                  }
                  P.g_prev[o_half + p * s_half + 4 * j] = bf8_pack(f);
                }
              } else if (p < P.out_planes && valid) {
#pragma unroll
                for (int c = 0; c < 8; ++c) f[c] = acc[8 * qq + c] + (P.dx_add ? sd[c] : 0.f);
                P.dx_hi[o_hi + (p - P.c_lo) * s_full + 8 * j] = bf8_pack(f);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      if (P.dbg && blockIdx.x == 0 && warp == BF_EPI_WARP0 && lane == 0 && lt < 256) {
        long long* d = P.dbg + 2048 + 4 * lt;
        d[0] = t_a; d[1] = t_b; d[2] = clock64(); d[3] = tile;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == BF_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 :: "r"(tmem_base), "r"(L.tmem_cols));
  }
}

static int bf_pow2_cols(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

// Tile shape, pipeline depth and TMEM budget of one layer; false = shape not supported.
static bool conv_bf16_plan(int Kp, int out_planes, int B, int H, int W, BfPlan& L) {
  L.Nmma = (out_planes * 8 + 15) / 16 * 16;
  if (L.Nmma > BF_MAX_N) return false;
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
  L.rbk = (H + BF_ROWS - 1) / BF_ROWS;
  if ((long long)B * (Kp > out_planes ? Kp : out_planes) * H * W >= (1ll << 31)) return false;
  const long long nt = (long long)B * L.rbk * L.tiles_x;
  if (nt > (1ll << 30)) return false;
  L.n_tiles = (int)nt;
  L.PWs = 8 * Xb + 2;
  L.n_chunks = (Kp + BF_KC - 1) / BF_KC;
  L.plane_elems = (BF_IN_ROWS * L.PWs + 7) / 8 * 8;
  L.a_elems = BF_KC * L.plane_elems;
  L.w_elems = 9 * BF_KC * L.Nmma;
  L.stage_elems = (L.a_elems + L.w_elems + 7) / 8 * 8;
  int ns = (200 * 1024) / (L.stage_elems * 16);
  if (ns > BF_MAX_STAGES) ns = BF_MAX_STAGES;
  if (ns < 2) return false;
  L.n_stages = ns;
  L.tma0 = L.tma1 = L.tma_only = 0;
  L.acc_cols = Xb * L.Nmma;
  L.tmem_cols = bf_pow2_cols(2 * L.acc_cols);
  return true;
}

// 4-D tensor map over a planar tensor [B][planes][H][W] of 16-byte pixels, seen as f32
// [B][planes][H][W*4]; the box is one plane's (16+2) x PWs halo block.
static int bf_encode_map(CUtensorMap* tm, const void* base, int B, int planes, int H, int W, int PWs) {
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
  const cuuint32_t box[4] = {(cuuint32_t)PWs * 4, (cuuint32_t)BF_IN_ROWS, 1, 1};
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

static int launch_conv_bf16(int mode, int srcm, BfConvParams& P, cudaStream_t st) {
  BfParams Q;
  if (!conv_bf16_plan(P.p0 + P.p1, P.out_planes, P.B, P.H, P.W, Q.m)) return DINCAE_EINVAL;
  static int sm_count = 0;
  static bool attr_set[3] = {false, false, false};
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  const int which = mode == 0 ? 0 : 1 + srcm;
  if (mode == 0 && srcm != 0) return DINCAE_EINVAL;
  CUtensorMap tm0, tm1;
  memset(&tm0, 0, sizeof(tm0));
  memset(&tm1, 0, sizeof(tm1));
  if (srcm == 0) {
    if (!P.half0 && P.p0 > 0 && ((uintptr_t)P.s0 & 15) == 0) {
      const int rc = bf_encode_map(&tm0, P.s0, P.B, P.p0, P.H, P.W, Q.m.PWs);
      if (rc) return rc;
      Q.m.tma0 = 1;
    }
    if (P.p1 > 0 && ((uintptr_t)P.s1 & 15) == 0) {
      const int rc = bf_encode_map(&tm1, P.s1, P.B, P.p1, P.H, P.W, Q.m.PWs);
      if (rc) return rc;
      Q.m.tma1 = 1;
    }
  }
  Q.m.tma_only = (srcm == 0 && Q.m.tma0 && (P.p1 == 0 || Q.m.tma1)) ? 1 : 0;
  static long long* dbg_dev = nullptr;
  static int dbg_on = -1;
  if (dbg_on < 0) {
    const char* e = getenv("DINCAE_BF_DEBUG");
    dbg_on = (e && e[0] == '1') ? 1 : 0;
    if (dbg_on) cudaMalloc(&dbg_dev, 3072 * sizeof(long long));
  }
  P.dbg = nullptr;
  if (dbg_on) {
    cudaMemsetAsync(dbg_dev, 0, 3072 * sizeof(long long), st);
    P.dbg = dbg_dev;
  }
  Q.c = P;
  const size_t smem = (size_t)Q.m.n_stages * Q.m.stage_elems * 16 + 128;
  if (!attr_set[which]) {
    cudaError_t e =
        which == 0 ? cudaFuncSetAttribute(conv3x3_bf16_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024)
        : which == 1 ? cudaFuncSetAttribute(conv3x3_bf16_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024)
                     : cudaFuncSetAttribute(conv3x3_bf16_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    if (e != cudaSuccess) return cuda_rc(e);
    attr_set[which] = true;
  }
  const int grid = Q.m.n_tiles < sm_count ? Q.m.n_tiles : sm_count;
  if (which == 0)
    launch_k(conv3x3_bf16_kernel<0, 0>, grid, BF_THREADS, smem, st, Q, tm0, tm1);
  else if (which == 1)
    launch_k(conv3x3_bf16_kernel<1, 0>, grid, BF_THREADS, smem, st, Q, tm0, tm1);
  else
    launch_k(conv3x3_bf16_kernel<1, 1>, grid, BF_THREADS, smem, st, Q, tm0, tm1);
  if (dbg_on) {
    static long long h[3072];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "conv_bf16 timeline CTA0: mode=%d srcm=%d Xb=%d N=%d chunks=%d stages=%d tiles=%d tma_only=%d\n", mode,
            srcm, Q.m.Xb, Q.m.Nmma, Q.m.n_chunks, Q.m.n_stages, Q.m.n_tiles, Q.m.tma_only);
    const long long t0 = h[1024];
    for (int i = 2; i < 10 && h[1024 + 4 * i]; ++i)
      fprintf(stderr, "  tile %2d mma: wait_tempty %7lld..%7lld first_full %7lld issued %7lld | epi: wait_tfull %7lld..%7lld done %7lld | load(chunk %d): wait_empty %7lld..%7lld arrived %7lld\n",
              i, h[1024 + 4 * i] - t0, h[1024 + 4 * i + 1] - t0, h[1024 + 4 * i + 2] - t0, h[1024 + 4 * i + 3] - t0,
              h[2048 + 4 * i] - t0, h[2048 + 4 * i + 1] - t0, h[2048 + 4 * i + 2] - t0, Q.m.n_chunks * i,
              h[4 * Q.m.n_chunks * i] - t0, h[4 * Q.m.n_chunks * i + 1] - t0, h[4 * Q.m.n_chunks * i + 2] - t0);
  }
  return check_launch();
}

// ---------------------------------------------------------------------------------------------
// Weight pre-pack: fp32 master WP [9][cinp4][cout_pad][4] -> bf16 operand blocks.
//   wf (forward, K = input channels):  [ceil(cin8/2)][9][2][cout_pad][8]
//       wf[ch][tap][pl][n][j] = W[tap][ci = 8*(2ch+pl)+j][co = n]
//   wd (data gradient, K = output channels, taps flipped): [ceil(co8/2)][9][2][nd][8]
//       wd[ch][tap][pl][n][j] = W[8-tap][ci = ci_first + n][co = 8*(2ch+pl)+j]   (n < nd, a multiple of 16)
// One thread per 16-byte output element.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_weights_bf16_kernel(const float* __restrict__ wp, int cinp4,
                                                                 int cout_pad, int co8, int ci_first, int nd,
                                                                 uint4* __restrict__ wf, uint4* __restrict__ wd) {
  pdl_trigger();
  pdl_wait();
  const int cin8 = (cinp4 + 1) >> 1;
  const int nf = wf ? ((cin8 + 1) >> 1) * 9 * 2 * cout_pad : 0;
  const int ndt = wd ? ((co8 + 1) >> 1) * 9 * 2 * nd : 0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float f[8];
  if (i < nf) {
    const int n = i % cout_pad;
    int r = i / cout_pad;
    const int pl = r & 1;
    r >>= 1;
    const int tap = r % 9, ch = r / 9;
    const int p4 = 2 * (2 * ch + pl);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 v = f4_zero();
      if (p4 + h < cinp4) v = __ldg(reinterpret_cast<const float4*>(wp) + ((size_t)(tap * cinp4 + p4 + h) * cout_pad + n));
      f[4 * h] = v.x; f[4 * h + 1] = v.y; f[4 * h + 2] = v.z; f[4 * h + 3] = v.w;
    }
    wf[i] = bf8_pack(f);
  } else if (i < nf + ndt) {
    const int k = i - nf;
    const int nl = k % nd, n = ci_first + nl;
    int r = k / nd;
    const int pl = r & 1;
    r >>= 1;
    const int tap = r % 9, ch = r / 9;
    const int co0 = 8 * (2 * ch + pl);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = co0 + j;
      f[j] = (co < cout_pad && (n >> 2) < cinp4)
                 ? __ldg(wp + ((size_t)((8 - tap) * cinp4 + (n >> 2)) * cout_pad + co) * 4 + (n & 3))
                 : 0.f;
    }
    wd[k] = bf8_pack(f);
  }
}

// planar-8 bf16 [B][C8][H][W][8] <-> planar-4 fp32 [B][C4][H][W][4]; c4 <= 2*c8 planes of the
// fp32 tensor are read / written (a missing odd plane reads as zero)
__global__ void __launch_bounds__(256) p4_to_p8_kernel(const float4* __restrict__ src, int c4, int c8, int HW,
                                                       size_t n, uint4* __restrict__ dst) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // over [B][c8][HW]
  if (i >= n) return;
  const int p = (int)(i % HW);
  const size_t r = i / HW;
  const int pl = (int)(r % c8);
  const size_t b = r / c8;
  float f[8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float4 v = f4_zero();
    if (2 * pl + h < c4) v = __ldg(src + ((b * c4 + 2 * pl + h) * HW + p));
    f[4 * h] = v.x; f[4 * h + 1] = v.y; f[4 * h + 2] = v.z; f[4 * h + 3] = v.w;
  }
  dst[i] = bf8_pack(f);
}

__global__ void __launch_bounds__(256) p8_to_p4_kernel(const uint4* __restrict__ src, int c4, int c8, int HW,
                                                       size_t n, float4* __restrict__ dst) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // over [B][c8][HW]
  if (i >= n) return;
  const int p = (int)(i % HW);
  const size_t r = i / HW;
  const int pl = (int)(r % c8);
  const size_t b = r / c8;
  float f[8];
  bf8_unpack(__ldg(src + i), f);
#pragma unroll
  for (int h = 0; h < 2; ++h)
    if (2 * pl + h < c4)
      dst[(b * c4 + 2 * pl + h) * HW + p] = make_float4(f[4 * h], f[4 * h + 1], f[4 * h + 2], f[4 * h + 3]);
}

}  // namespace dincae

using namespace dincae;

extern "C" size_t dincae_conv_weights_bf16_elems(int cinp4, int cout_pad, int nd, size_t* wd_elems) {
  const int cin8 = (cinp4 + 1) / 2, co8 = cout_pad / 8;
  if (wd_elems) *wd_elems = (size_t)((co8 + 1) / 2) * 9 * 2 * nd * 8;
  return (size_t)((cin8 + 1) / 2) * 9 * 2 * cout_pad * 8;
}

extern "C" int dincae_conv_weights_bf16(const float* wpack, int cinp4, int cout_pad, int ci_first, int nd,
                                        void* wf, void* wd, void* stream) {
  if (!wpack || (!wf && !wd) || cinp4 <= 0 || cout_pad <= 0 || (cout_pad & 15) ||
      (wd && (nd <= 0 || (nd & 15) || ci_first < 0)))
    return DINCAE_EINVAL;
  size_t nde = 0;
  const size_t nfe = dincae_conv_weights_bf16_elems(cinp4, cout_pad, wd ? nd : 16, &nde);
  const size_t total = (wf ? nfe / 8 : 0) + (wd ? nde / 8 : 0);
  launch_k(conv_weights_bf16_kernel, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream,
           wpack, cinp4, cout_pad, cout_pad / 8, ci_first, nd, reinterpret_cast<uint4*>(wf),
           reinterpret_cast<uint4*>(wd));
  return check_launch();
}

extern "C" int dincae_p4_to_p8(const float* src, int c4, int c8, int B, int H, int W, void* dst, void* stream) {
  if (!src || !dst || c4 <= 0 || c8 <= 0 || c4 > 2 * c8 || B <= 0 || H <= 0 || W <= 0) return DINCAE_EINVAL;
  const size_t n = (size_t)B * c8 * H * W;
  launch_k(p4_to_p8_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream,
           reinterpret_cast<const float4*>(src), c4, c8, H * W, n, reinterpret_cast<uint4*>(dst));
  return check_launch();
}

extern "C" int dincae_p8_to_p4(const void* src, int c4, int c8, int B, int H, int W, float* dst, void* stream) {
  if (!src || !dst || c4 <= 0 || c8 <= 0 || c4 > 2 * c8 || B <= 0 || H <= 0 || W <= 0) return DINCAE_EINVAL;
  const size_t n = (size_t)B * c8 * H * W;
  launch_k(p8_to_p4_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream,
           reinterpret_cast<const uint4*>(src), c4, c8, H * W, n, reinterpret_cast<float4*>(dst));
  return check_launch();
}

extern "C" int dincae_conv3x3_fwd_bf16(const void* src0, int c0p, int src0_half, const void* src1, int c1p,
                                       const void* wf, const float* bias, int cout_pad,
                                       int B, int H, int W, int coutp, float neg_slope,
                                       void* out_full, void* out_pool, uint8_t* out_sign,
                                       float* out_f32, int out_f32_planes, void* stream) {
  if (!src0 || !wf || c0p <= 0 || B <= 0 || H <= 0 || W <= 0 || coutp <= 0) return DINCAE_EINVAL;
  if (coutp * 8 > cout_pad || (cout_pad & 15) || cout_pad != (coutp * 8 + 15) / 16 * 16) return DINCAE_EINVAL;
  if (!out_full && !out_pool && !out_f32) return DINCAE_EINVAL;
  if (out_f32 && (out_f32_planes <= 0 || out_f32_planes > 2 * coutp)) return DINCAE_EINVAL;
  BfConvParams P = {};
  P.s0 = reinterpret_cast<const uint4*>(src0);
  P.s1 = reinterpret_cast<const uint4*>(src1);
  P.p0 = c0p; P.p1 = src1 ? c1p : 0; P.half0 = src0_half;
  P.B = B; P.H = H; P.W = W; P.Hh = (H + 1) / 2; P.Wh = (W + 1) / 2; P.Wq = (W + 3) / 4;
  P.w = reinterpret_cast<const uint4*>(wf); P.bias = bias; P.out_planes = coutp; P.slope = neg_slope;
  P.out_full = reinterpret_cast<uint4*>(out_full);
  P.out_pool = reinterpret_cast<uint4*>(out_pool);
  P.out_sign = reinterpret_cast<uint32_t*>(out_sign);
  P.out_f32 = reinterpret_cast<float4*>(out_f32);
  P.out_f32_planes = out_f32 ? out_f32_planes : 0;
  return launch_conv_bf16(0, 0, P, (cudaStream_t)stream);
}

extern "C" int dincae_conv3x3_dgrad_bf16(const void* g_full, const void* g_pool, const uint8_t* sign,
                                         float neg_slope, int gp, const void* wd, int nd,
                                         int B, int H, int W,
                                         int c_lo, const void* prev_act, float prev_slope, void* g_prev,
                                         int c_hi, void* dx_hi, const void* dx_add, void* stream) {
  if (!wd || gp <= 0 || B <= 0 || H <= 0 || W <= 0) return DINCAE_EINVAL;
  if (c_lo < 0 || c_hi < 0 || c_lo + c_hi == 0 || nd != ((c_lo + c_hi) * 8 + 15) / 16 * 16) return DINCAE_EINVAL;
  if ((c_lo > 0 && !g_prev) || (c_hi > 0 && !dx_hi)) return DINCAE_EINVAL;
  if (!g_full && (!g_pool || !sign)) return DINCAE_EINVAL;
  BfConvParams P = {};
  const int srcm = g_full ? 0 : 1;
  P.s0 = reinterpret_cast<const uint4*>(g_full ? g_full : g_pool);
  P.s1 = nullptr;
  P.sign = sign;
  P.p0 = gp; P.p1 = 0; P.half0 = srcm; P.src_slope = neg_slope;
  P.B = B; P.H = H; P.W = W; P.Hh = (H + 1) / 2; P.Wh = (W + 1) / 2; P.Wq = (W + 3) / 4;
  P.w = reinterpret_cast<const uint4*>(wd); P.bias = nullptr; P.out_planes = c_lo + c_hi; P.slope = 0.f;
  P.c_lo = c_lo; P.c_hi = c_hi;
  P.prev_act = reinterpret_cast<const uint4*>(prev_act); P.prev_slope = prev_slope;
  P.g_prev = reinterpret_cast<uint4*>(g_prev);
  P.dx_hi = reinterpret_cast<uint4*>(dx_hi);
  P.dx_add = reinterpret_cast<const uint4*>(dx_add);
  return launch_conv_bf16(1, srcm, P, (cudaStream_t)stream);
}
