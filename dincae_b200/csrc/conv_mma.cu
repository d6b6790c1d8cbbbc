// 3x3 SAME convolution (forward and data gradient) as an implicit GEMM on the 5th-gen
// tensor cores: tcgen05.mma kind::tf32, accumulators in TMEM, operands in shared memory in
// the no-swizzle canonical layout that the planar-4 activation format maps onto directly.
//
// Formulation.  A CTA owns R output rows of one image.  Its input rows (R+2, zero halo)
// are staged as [plane][padded linear pixel][4 channels] with a row pitch P >= W+2.  With
// the GEMM M index running over *padded linear* output positions m = r*P + x, the A operand
// of tap (dy,dx) is the same buffer shifted by (dy*P + dx) pixels: one descriptor
// start-address change, no im2col copy.  Columns x >= W of every row are computed and
// discarded (<= 4 %).  K runs over input planes (2 planes = one K=8 tf32 MMA), N over the
// output channels (padded to 16).  Per M-tile of 128 positions the 9 taps x K-steps
// accumulate into one TMEM region; all M-tiles of the CTA sit side by side in TMEM.
//
// The loader is the same conv_src_load() gather as the FFMA kernels (upsample, concat,
// un-pool + leaky-ReLU gradient fused), rounding to TF32 (rna) on the way into shared
// memory; the epilogue applies bias / leaky-ReLU / sign mask / 2x2 pool or the data-gradient
// fusions straight from TMEM.
#include "common.cuh"
#include "conv_params.cuh"
#include "umma.cuh"
#include <math.h>

namespace dincae {

struct MmaPlan {
  int R, P, n_mt, Kc, n_chunks, Nmma, aps, tmem_cols;   // aps: A plane stride in pixels
  int stage_planes;                                     // planes staged for the 2x2 reduction
  size_t smem_bytes;
};

struct MmaConvParams {
  ConvParams c;
  MmaPlan m;
};

constexpr int MMA_THREADS = 256;
constexpr int MMA_ISSUERS = 4;

template <int MODE>
__global__ void __launch_bounds__(MMA_THREADS) conv3x3_mma_kernel(const MmaConvParams Q) {
  const ConvParams& P = Q.c;
  const MmaPlan& L = Q.m;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* As = reinterpret_cast<float4*>(smem_raw);                 // [Kc][aps]
  float4* Ws = As + (size_t)L.Kc * L.aps;                           // [9][Kc][Nmma]
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_holder;
  __shared__ float4 bias_s[64];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int y0 = blockIdx.x * L.R;
  const int Kp = P.src.planes();
  const int PW = L.P;
  const int n_issuers = L.n_mt < MMA_ISSUERS ? L.n_mt : MMA_ISSUERS;   // only warps that own a tile

  if (MODE == 0 && tid < 64)
    bias_s[tid] = (P.bias && tid < P.out_planes) ? __ldg(reinterpret_cast<const float4*>(P.bias) + tid)
                                                 : f4_zero();
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(n_issuers));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&tmem_holder)), "r"(L.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_holder;

  // instruction descriptor: D=f32, A=B=tf32, both K-major, N = Nmma, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(L.Nmma >> 3) << 17) |
                         ((uint32_t)(128 >> 4) << 24);
  const int rows_in = L.R + 2;

  for (int ch = 0; ch < L.n_chunks; ++ch) {
    const int k0 = ch * L.Kc;
    if (ch > 0) mbar_wait(&mbar, (uint32_t)((ch - 1) & 1));   // previous MMAs done: smem is free
    // ---- A chunk: [Kc planes][rows_in][PW]; a warp per row, 4 loads in flight per lane -----
    const int a_rows = L.Kc * rows_in;
    for (int row = warp; row < a_rows; row += MMA_THREADS / 32) {
      const int pl = row / rows_in, ry = row - pl * rows_in;
      const bool plane_ok = (k0 + pl) < Kp;
      float4* dst = As + (size_t)pl * L.aps + ry * PW;
      const int yy = y0 - 1 + ry;
      for (int xp = lane; xp < PW; xp += 128) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int xq = xp + 32 * u;
          v[u] = (plane_ok && xq < PW) ? conv_src_load(P.src, b, k0 + pl, yy, xq - 1) : f4_zero();
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int xq = xp + 32 * u;
          if (xq < PW) dst[xq] = to_tf32(v[u]);
        }
      }
    }
    // ---- weight chunk: [9][Kc][Nmma] -------------------------------------------------------
    const int w_elems = 9 * L.Kc * L.Nmma;
    for (int idx = tid; idx < w_elems; idx += MMA_THREADS) {
      const int n = idx % L.Nmma;
      const int r = idx / L.Nmma;
      const int pl = r % L.Kc, tap = r / L.Kc;
      float4 w = f4_zero();
      if (k0 + pl < Kp && n < P.out_planes * 4) {
        if (MODE == 0) {
          w = __ldg(reinterpret_cast<const float4*>(P.w) +
                    ((size_t)(tap * P.w_cinp + (k0 + pl)) * P.cout_pad + n));
        } else {
          const float* base = P.w + (((size_t)((8 - tap) * P.w_cinp + (n >> 2)) * P.cout_pad +
                                      4 * (k0 + pl)) * 4 + (n & 3));
          w = make_float4(__ldg(base), __ldg(base + 4), __ldg(base + 8), __ldg(base + 12));
        }
        w = to_tf32(w);
      }
      Ws[idx] = w;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // Four elected threads (lane 0 of warps 0-3) issue the MMAs, M-tiles round-robin: a
    // single thread needs ~50 cycles per tcgen05.mma (descriptor arithmetic + issue), more
    // than the tensor pipe needs to execute one, so one issuer would starve it.  The other
    // lanes of an issuing warp must not enter the mbarrier wait before their lane 0 is done:
    // try_wait suspends the whole warp in hardware and would stall the issuing lane until
    // the suspend time limit expires.
    if (warp < n_issuers) {
      if (lane == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // start-address field (16-byte units) is the low 14 bits: tap / plane / tile steps are
        // plain additions on the descriptor
        const uint64_t adesc0 = umma_desc(smem_u32(As), (uint32_t)L.aps * 16u, 128u);
        const uint64_t wdesc0 = umma_desc(smem_u32(Ws), (uint32_t)L.Nmma * 16u, 128u);
        const uint64_t a_hi = adesc0 & ~0x3fffull, w_hi = wdesc0 & ~0x3fffull;
        const uint32_t a_lo0 = (uint32_t)(adesc0 & 0x3fffull), w_lo0 = (uint32_t)(wdesc0 & 0x3fffull);
        const uint32_t a_step = 2u * (uint32_t)L.aps, w_step = 2u * (uint32_t)L.Nmma;
        for (int j = warp; j < L.n_mt; j += n_issuers) {
          const uint32_t d = tmem_base + (uint32_t)(j * L.Nmma);
          const uint32_t aj = a_lo0 + (uint32_t)(128 * j);
          uint32_t wt = w_lo0;
          uint32_t acc = ch != 0 ? 1u : 0u;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            uint32_t at = aj + (uint32_t)((tap / 3) * PW + (tap % 3));
            for (int kp = 0; kp < L.Kc; kp += 2) {
              umma_tf32(d, a_hi | (uint64_t)(at & 0x3fffu), w_hi | (uint64_t)(wt & 0x3fffu), idesc, acc);
              acc = 1u;
              at += a_step;
              wt += w_step;
            }
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                     :: "r"(smem_u32(&mbar)) : "memory");
      }
      __syncwarp();
    }
  }
  mbar_wait(&mbar, (uint32_t)((L.n_chunks - 1) & 1));
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: TMEM -> registers -> global / staging -----------------------------------------
  float4* stage = reinterpret_cast<float4*>(smem_raw);       // aliases As/Ws (MMAs are complete)
  const int PH2 = PW >> 1;
  const int wq = warp & 3, wg = warp >> 2;
  const int Hh = (P.H + 1) >> 1, Wh = (P.W + 1) >> 1;
  const int n_c16 = (P.out_planes + 3) >> 2;
  for (int j = wg; j < L.n_mt; j += MMA_THREADS / 128) {
    const int m = 128 * j + 32 * wq + lane;
    const int r = m / PW, x = m - r * PW;
    const int y = y0 + r;
    const bool valid = (r < L.R) && (x < P.W) && (y < P.H);
    for (int c16 = 0; c16 < n_c16; ++c16) {
      float acc[16];
      tmem_ld16(tmem_base + ((uint32_t)(32 * wq) << 16) + (uint32_t)(j * L.Nmma + 16 * c16), acc);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int p = 4 * c16 + q;
        if (p >= P.out_planes) continue;                       // warp-uniform
        float4 a = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        if (MODE == 0) {
          a = f4_add(a, bias_s[p]);
          unsigned nib = 0;
          if (valid) {
            nib = (a.x > 0.f ? 1u : 0u) | (a.y > 0.f ? 2u : 0u) | (a.z > 0.f ? 4u : 0u) |
                  (a.w > 0.f ? 8u : 0u);
          }
          a.x = a.x > 0.f ? a.x : P.slope * a.x;
          a.y = a.y > 0.f ? a.y : P.slope * a.y;
          a.z = a.z > 0.f ? a.z : P.slope * a.z;
          a.w = a.w > 0.f ? a.w : P.slope * a.w;
          if (!valid) a = f4_zero();
          if (valid && P.out_full)
            P.out_full[((size_t)(b * P.out_planes + p) * P.H + y) * P.W + x] = a;
          if (P.out_sign) {
            unsigned wbits = nib << (4 * (lane & 3));
            wbits |= __shfl_xor_sync(0xffffffffu, wbits, 1);
            wbits |= __shfl_xor_sync(0xffffffffu, wbits, 2);
            if ((lane & 3) == 0 && valid) {
              const int Wq = (P.W + 3) >> 2;
              P.out_sign[((size_t)(b * P.out_planes + p) * P.H + y) * Wq + (x >> 2)] = (uint16_t)wbits;
            }
          }
          if (P.out_pool) {
            float4 h = a;
            h.x += __shfl_xor_sync(0xffffffffu, a.x, 1);
            h.y += __shfl_xor_sync(0xffffffffu, a.y, 1);
            h.z += __shfl_xor_sync(0xffffffffu, a.z, 1);
            h.w += __shfl_xor_sync(0xffffffffu, a.w, 1);
            if ((lane & 1) == 0 && r < L.R) stage[((size_t)p * L.R + r) * PH2 + (x >> 1)] = h;
          }
        } else {
          if (p < P.c_lo) {
            if (!valid) a = f4_zero();
            float4 h = a;
            h.x += __shfl_xor_sync(0xffffffffu, a.x, 1);
            h.y += __shfl_xor_sync(0xffffffffu, a.y, 1);
            h.z += __shfl_xor_sync(0xffffffffu, a.z, 1);
            h.w += __shfl_xor_sync(0xffffffffu, a.w, 1);
            if ((lane & 1) == 0 && r < L.R) stage[((size_t)p * L.R + r) * PH2 + (x >> 1)] = h;
          } else if (valid) {
            const size_t at = ((size_t)(b * P.c_hi + (p - P.c_lo)) * P.H + y) * P.W + x;
            if (P.dx_add) a = f4_add(a, __ldg(&P.dx_add[at]));
            P.dx_hi[at] = a;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 :: "r"(tmem_base), "r"(L.tmem_cols));
  }
  // ---- second pass of the 2x2 reductions (rows pair up across warps) ---------------------------
  const int sp = L.stage_planes;
  if (sp > 0) {
    const int half_rows = L.R >> 1;
    const int total = sp * half_rows * Wh;
    for (int idx = tid; idx < total; idx += MMA_THREADS) {
      const int p = idx / (half_rows * Wh);
      const int rem = idx - p * (half_rows * Wh);
      const int rq = rem / Wh, xh = rem - rq * Wh;
      const int y = y0 + 2 * rq;
      if (y >= P.H) continue;
      float4 s = stage[((size_t)p * L.R + 2 * rq) * PH2 + xh];
      const bool two_rows = (y + 1 < P.H);
      if (two_rows) s = f4_add(s, stage[((size_t)p * L.R + 2 * rq + 1) * PH2 + xh]);
      if (MODE == 0) {
        const int cy = two_rows ? 2 : 1;
        const int cx = (2 * xh + 1 < P.W) ? 2 : 1;
        P.out_pool[((size_t)(b * P.out_planes + p) * Hh + (y >> 1)) * Wh + xh] =
            f4_scale(s, 1.0f / (float)(cx * cy));
      } else {
        const size_t at = ((size_t)(b * P.c_lo + p) * Hh + (y >> 1)) * Wh + xh;
        if (P.prev_act) {
          const float4 pa = __ldg(&P.prev_act[at]);
          s.x *= pa.x > 0.f ? 1.0f : P.prev_slope;
          s.y *= pa.y > 0.f ? 1.0f : P.prev_slope;
          s.z *= pa.z > 0.f ? 1.0f : P.prev_slope;
          s.w *= pa.w > 0.f ? 1.0f : P.prev_slope;
        }
        P.g_prev[at] = s;
      }
    }
  }
}

static int next_pow2(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

// Pick rows per CTA (R, even) and planes per K chunk (Kc, even) under the shared-memory and
// TMEM budgets.  Returns false when the shape does not fit the MMA path.
bool conv_mma_plan(int Kp, int out_planes, int stage_planes, int B, int H, int W, MmaPlan& best) {
  const int P = (W + 2 + 3) / 4 * 4;
  const int Nmma = (out_planes * 4 + 15) / 16 * 16;
  if (Nmma > 256 || out_planes > 64) return false;
  const int Kp2 = (Kp + 1) / 2 * 2;
  double best_cost = -1.0;
  const int Hmax = (H + 1) / 2 * 2;
  for (int R = 2; R <= Hmax && R <= 64; R += 2) {
    const int n_mt = (R * P + 127) / 128;
    if (n_mt * Nmma > 512) break;
    int aps = (R + 2) * P;
    const int need = 128 * n_mt + 2 * P + 2;
    if (need > aps) aps = need;
    aps = (aps + 7) / 8 * 8;
    if ((size_t)aps * 16 >= (1u << 18)) break;
    const size_t stage_bytes = (size_t)stage_planes * R * (P / 2) * 16;
    const int tmem_cols = next_pow2(n_mt * Nmma);
    for (int Kc = Kp2; Kc >= 2; Kc -= 2) {
      const size_t op_bytes = (size_t)Kc * aps * 16 + (size_t)9 * Kc * Nmma * 16;
      const size_t smem = op_bytes > stage_bytes ? op_bytes : stage_bytes;
      if (smem > 110 * 1024) continue;
      const int n_chunks = (Kp2 + Kc - 1) / Kc;
      const int blocks = (H + R - 1) / R;
      // rough per-CTA latency (cycles): fixed + loads (4 in flight) + MMA issue + epilogue
      int occ = (int)((220 * 1024) / (smem + 1024));
      if (occ > 512 / tmem_cols) occ = 512 / tmem_cols;
      if (occ > 8) occ = 8;
      if (occ < 1) occ = 1;
      const double loads = (double)n_chunks * (Kc * (R + 2) * P + 9 * Kc * Nmma) / 256.0;
      const double t_cta = 4000.0 + loads * 150.0 + n_chunks * 1500.0 +
                           (double)n_mt * 9 * (Kp2 / 2) * 40.0 +
                           (double)((n_mt + 1) / 2) * ((out_planes + 3) / 4) * 500.0;
      const double grid = (double)blocks * B;
      const double waves = ceil(grid / (148.0 * occ));
      // co-resident CTAs overlap their phases only partially
      const double cost = waves * t_cta * (1.0 + 0.35 * (occ - 1));
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best.R = R; best.P = P; best.n_mt = n_mt; best.Kc = Kc; best.n_chunks = n_chunks;
        best.Nmma = Nmma; best.aps = aps; best.tmem_cols = tmem_cols;
        best.stage_planes = stage_planes; best.smem_bytes = smem;
      }
      break;      // largest Kc that fits for this R
    }
  }
  return best_cost > 0.0;
}

int launch_conv_mma(int mode, const ConvParams& P, int stage_planes, cudaStream_t st) {
  MmaConvParams Q;
  Q.c = P;
  if (!conv_mma_plan(P.src.planes(), P.out_planes, stage_planes, P.B, P.H, P.W, Q.m)) return 1;
  static bool attr_set[2] = {false, false};
  if (!attr_set[mode]) {
    cudaError_t e = mode == 0
        ? cudaFuncSetAttribute(conv3x3_mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
        : cudaFuncSetAttribute(conv3x3_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return cuda_rc(e);
    attr_set[mode] = true;
  }
  dim3 grid((P.H + Q.m.R - 1) / Q.m.R, P.B);
  if (mode == 0)
    conv3x3_mma_kernel<0><<<grid, MMA_THREADS, Q.m.smem_bytes, st>>>(Q);
  else
    conv3x3_mma_kernel<1><<<grid, MMA_THREADS, Q.m.smem_bytes, st>>>(Q);
  return check_launch();
}

}  // namespace dincae
