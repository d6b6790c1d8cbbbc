// 3x3 SAME convolutions of the DINCAE encoder/decoder on planar-4 activations, fp32 FFMA
// path: forward (+bias, leaky-ReLU, 2x2 SAME average pool, sign mask), data gradient
// (+nearest-upsample gradient, activation gradients, skip-gradient add) and weight/bias
// gradient.  The logical input of every kernel is assembled on the fly by conv_src_load()
// (concatenation, x2 nearest upsample, un-pool + leaky-ReLU gradient), so none of those
// intermediate tensors of the TF graph (DINCAE/__init__.py:442-452, 496-513) ever exists
// in HBM.
#include "common.cuh"
#include "conv_params.cuh"

namespace dincae {

constexpr int CPB = 4;   // logical output planes per block (one per 64 threads)
constexpr int CK = 4;    // logical input planes staged in shared memory per step


// MODE 0: forward.  MODE 1: data gradient (weights read transposed, taps mirrored).
// Block = 256 threads = 64 pixel-quads (4 consecutive x each) x CPB output planes; thread
// tile 4 pixels x 4 output channels.  Lanes run along y first so that (a) the float4
// shared-memory reads of a quarter-warp fall in 8 different rows of odd pitch (conflict
// free) and (b) the vertical pooling partner is lane^1.
template <int TWQ, int TH, int MODE>
__global__ void __launch_bounds__(256) conv3x3_kernel(const ConvParams P) {
  pdl_trigger();
  pdl_wait();
  constexpr int TW = 4 * TWQ;
  constexpr int LW = TW + 2;           // loaded width (halo)
  constexpr int PW = LW | 1;           // odd pitch
  constexpr int PH = TH + 2;
  static_assert(TWQ * TH == 64, "64 pixel-quads per block");
  static_assert((TH & 1) == 0, "tile rows pair up for pooling");
  __shared__ float4 tile[CK][PH][PW];
  __shared__ float4 wsm[9][CK][CPB * 4];

  const int tiles_x = ceil_div(P.W, TW);
  const int tx = blockIdx.x % tiles_x, tyb = blockIdx.x / tiles_x;
  const int b = blockIdx.z;
  const int op0 = blockIdx.y * CPB;
  const int tid = threadIdx.x;
  const int pt = tid & 63, cpl = tid >> 6;
  const int ty = pt % TH, qx = pt / TH;
  const int x0 = tx * TW, y0 = tyb * TH;
  const int Kp = P.src.planes();
  const bool plane_on = (op0 + cpl) < P.out_planes;     // warp-uniform

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Kp; k0 += CK) {
    const int kc = min(CK, Kp - k0);
    __syncthreads();
    for (int idx = tid; idx < kc * PH * LW; idx += 256) {
      const int pl = idx / (PH * LW);
      const int r = idx - pl * (PH * LW);
      const int yy = r / LW, xx = r - yy * LW;
      tile[pl][yy][xx] = conv_src_load(P.src, b, k0 + pl, y0 + yy - 1, x0 + xx - 1);
    }
    for (int idx = tid; idx < 9 * kc * (CPB * 4); idx += 256) {
      const int col = idx & (CPB * 4 - 1);
      const int r = idx / (CPB * 4);
      const int pl = r % kc, tap = r / kc;
      const int oc = op0 * 4 + col;
      float4 w = f4_zero();
      if (oc < P.out_planes * 4) {
        if (MODE == 0) {
          w = __ldg(reinterpret_cast<const float4*>(P.w) +
                    ((size_t)(tap * P.w_cinp + (k0 + pl)) * P.cout_pad + oc));
        } else {
          const float* base = P.w + (((size_t)((8 - tap) * P.w_cinp + (oc >> 2)) * P.cout_pad +
                                      4 * (k0 + pl)) * 4 + (oc & 3));
          w = make_float4(__ldg(base), __ldg(base + 4), __ldg(base + 8), __ldg(base + 12));
        }
      }
      wsm[tap][pl][col] = w;
    }
    __syncthreads();
    if (plane_on) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        for (int pl = 0; pl < kc; ++pl) {
          float4 in[6];
#pragma unroll
          for (int k = 0; k < 6; ++k) in[k] = tile[pl][ty + dy][4 * qx + k];
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
            for (int co = 0; co < 4; ++co) {
              const float4 w = wsm[dy * 3 + dx][pl][cpl * 4 + co];
#pragma unroll
              for (int px = 0; px < 4; ++px) acc[px][co] = f4_dot(in[px + dx], w, acc[px][co]);
            }
          }
        }
      }
    }
  }

  if (!plane_on) return;
  const int op = op0 + cpl;
  const int y = y0 + ty;
  const int xb = x0 + 4 * qx;
  const int Hh = (P.H + 1) >> 1, Wh = (P.W + 1) >> 1;

  if (MODE == 0) {
    float4 bias4 = f4_zero();
    if (P.bias) bias4 = __ldg(reinterpret_cast<const float4*>(P.bias) + op);
    unsigned bits = 0;
    float4 a[4];
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      float v[4];
      const bool inb = (y < P.H) && (xb + px < P.W);
#pragma unroll
      for (int co = 0; co < 4; ++co) {
        float z = acc[px][co] + f4_get(bias4, co);
        if (z > 0.f && inb) bits |= 1u << (4 * px + co);
        v[co] = z > 0.f ? z : P.slope * z;
      }
      a[px] = inb ? make_float4(v[0], v[1], v[2], v[3]) : f4_zero();
      if (inb && P.out_full)
        P.out_full[((size_t)(b * P.out_planes + op) * P.H + y) * P.W + xb + px] = P.round_out ? to_tf32(a[px]) : a[px];
    }
    if (P.out_sign && y < P.H && xb < P.W) {
      const int Wq = (P.W + 3) >> 2;
      P.out_sign[((size_t)(b * P.out_planes + op) * P.H + y) * Wq + (xb >> 2)] = (uint16_t)bits;
    }
    if (P.out_pool) {
      float4 h[2] = {f4_add(a[0], a[1]), f4_add(a[2], a[3])};
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        h[q].x += __shfl_xor_sync(0xffffffffu, h[q].x, 1);
        h[q].y += __shfl_xor_sync(0xffffffffu, h[q].y, 1);
        h[q].z += __shfl_xor_sync(0xffffffffu, h[q].z, 1);
        h[q].w += __shfl_xor_sync(0xffffffffu, h[q].w, 1);
      }
      if ((ty & 1) == 0 && y < P.H) {
        const int cy = (y + 1 < P.H) ? 2 : 1;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int x = xb + 2 * q;
          if (x < P.W) {
            const int cx = (x + 1 < P.W) ? 2 : 1;
            const float4 pv = f4_scale(h[q], 1.0f / (float)(cx * cy));
            P.out_pool[((size_t)(b * P.out_planes + op) * Hh + (y >> 1)) * Wh + (x >> 1)] =
                P.round_out ? to_tf32(pv) : pv;
          }
        }
      }
    }
  } else {
    if (op < P.c_lo) {
      float4 a[4];
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        const bool inb = (y < P.H) && (xb + px < P.W);
        a[px] = inb ? make_float4(acc[px][0], acc[px][1], acc[px][2], acc[px][3]) : f4_zero();
      }
      float4 h[2] = {f4_add(a[0], a[1]), f4_add(a[2], a[3])};
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        h[q].x += __shfl_xor_sync(0xffffffffu, h[q].x, 1);
        h[q].y += __shfl_xor_sync(0xffffffffu, h[q].y, 1);
        h[q].z += __shfl_xor_sync(0xffffffffu, h[q].z, 1);
        h[q].w += __shfl_xor_sync(0xffffffffu, h[q].w, 1);
      }
      if ((ty & 1) == 0 && y < P.H) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int x = xb + 2 * q;
          if (x < P.W) {
            const size_t at = ((size_t)(b * P.c_lo + op) * Hh + (y >> 1)) * Wh + (x >> 1);
            float4 g = h[q];
            if (P.prev_act) {
              const float4 pa = __ldg(&P.prev_act[at]);
              g.x *= pa.x > 0.f ? 1.0f : P.prev_slope;
              g.y *= pa.y > 0.f ? 1.0f : P.prev_slope;
              g.z *= pa.z > 0.f ? 1.0f : P.prev_slope;
              g.w *= pa.w > 0.f ? 1.0f : P.prev_slope;
            }
            P.g_prev[at] = P.round_out ? to_tf32(g) : g;
          }
        }
      }
    } else {
      const int hp = op - P.c_lo;
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        if (y < P.H && xb + px < P.W) {
          const size_t at = ((size_t)(b * P.c_hi + hp) * P.H + y) * P.W + xb + px;
          float4 v = make_float4(acc[px][0], acc[px][1], acc[px][2], acc[px][3]);
          if (P.dx_add) v = f4_add(v, __ldg(&P.dx_add[at]));
          P.dx_hi[at] = P.round_out ? to_tf32(v) : v;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Weight gradient.  dW[tap][ci][co] = sum over pixels of U[pix + tap] * G[pix].
// grid = (pixel blocks, panels); a panel is CIB input planes x COB gradient planes.
// Thread = (item, slice): item = (dy, input plane, gradient plane) owning the 3 dx taps
// x 4 ci x 4 co accumulators; 8 slices (lanes of an octet) stride over the pixels of a
// 16x8 tile so a quarter-warp always reads 8 consecutive float4 from shared memory.
// Blocks loop over tiles, keep accumulators in registers, and write ONE partial each;
// wgrad_reduce_kernel then sums the partials in fixed order (deterministic).
// ---------------------------------------------------------------------------------------
struct WgradParams {
  ConvSrc u;
  ConvSrc g;
  int B, H, W;
  int cinp, gp, cout_pad;
  int cib, cob;            // panel shape
  int panels_ci, panels_co;
  float* partial;          // [gridDim.x][wsize + cout_pad]
  int wsize;               // 9*cinp*cout_pad*4
};

constexpr int WG_TW = 16, WG_TH = 8, WG_PW = 19, WG_PH = 10;

__global__ void __launch_bounds__(256) conv3x3_wgrad_kernel(const WgradParams P) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float4 smem[];
  float4* Us = smem;                                   // [cib][WG_PH][WG_PW]
  float4* Gs = smem + P.cib * WG_PH * WG_PW;           // [cob][WG_TH][WG_TW]
  const int tid = threadIdx.x;
  const int oct = tid >> 3, s = tid & 7;
  const int pci = blockIdx.y % P.panels_ci, pco = blockIdx.y / P.panels_ci;
  const int ci0 = pci * P.cib, co0 = pco * P.cob;
  const int cib = min(P.cib, P.cinp - ci0), cob = min(P.cob, P.gp - co0);
  const int n_items = 3 * cib * cob;
  const bool item_ok = oct < n_items;
  const int dy = oct % 3;
  const int rr = oct / 3;
  const int cil = item_ok ? rr % cib : 0, col = item_ok ? rr / cib : 0;
  const bool want_db = item_ok && dy == 1 && cil == 0 && pci == 0;

  float acc[3][4][4];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[a][c][j] = 0.f;
  float4 gsum = f4_zero();

  const int tiles_x = ceil_div(P.W, WG_TW), tiles_y = ceil_div(P.H, WG_TH);
  const int tiles_img = tiles_x * tiles_y;
  const int ntiles = P.B * tiles_img;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int b = t / tiles_img;
    const int r = t - b * tiles_img;
    const int y0 = (r / tiles_x) * WG_TH, x0 = (r % tiles_x) * WG_TW;
    __syncthreads();
    for (int idx = tid; idx < cib * WG_PH * (WG_TW + 2); idx += 256) {
      const int pl = idx / (WG_PH * (WG_TW + 2));
      const int q = idx - pl * (WG_PH * (WG_TW + 2));
      const int yy = q / (WG_TW + 2), xx = q - yy * (WG_TW + 2);
      Us[(pl * WG_PH + yy) * WG_PW + xx] = conv_src_load(P.u, b, ci0 + pl, y0 + yy - 1, x0 + xx - 1);
    }
    for (int idx = tid; idx < cob * WG_TH * WG_TW; idx += 256) {
      const int pl = idx / (WG_TH * WG_TW);
      const int q = idx - pl * (WG_TH * WG_TW);
      const int yy = q / WG_TW, xx = q - yy * WG_TW;
      Gs[idx] = conv_src_load(P.g, b, co0 + pl, y0 + yy, x0 + xx);
    }
    __syncthreads();
    if (item_ok) {
#pragma unroll 2
      for (int yy = 0; yy < WG_TH; ++yy) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int x = s + 8 * i;
          const float4 g = Gs[(col * WG_TH + yy) * WG_TW + x];
          if (want_db) gsum = f4_add(gsum, g);
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const float4 u = Us[(cil * WG_PH + yy + dy) * WG_PW + x + dx];
            const float uu[4] = {u.x, u.y, u.z, u.w};
            const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[dx][c][j] = fmaf(uu[c], gg[j], acc[dx][c][j]);
          }
        }
      }
    }
  }
  // reduce the 8 slices of each octet
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = acc[a][c][j];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        acc[a][c][j] = v;
      }
  gsum.x += __shfl_xor_sync(0xffffffffu, gsum.x, 4);
  gsum.x += __shfl_xor_sync(0xffffffffu, gsum.x, 2);
  gsum.x += __shfl_xor_sync(0xffffffffu, gsum.x, 1);
  gsum.y += __shfl_xor_sync(0xffffffffu, gsum.y, 4);
  gsum.y += __shfl_xor_sync(0xffffffffu, gsum.y, 2);
  gsum.y += __shfl_xor_sync(0xffffffffu, gsum.y, 1);
  gsum.z += __shfl_xor_sync(0xffffffffu, gsum.z, 4);
  gsum.z += __shfl_xor_sync(0xffffffffu, gsum.z, 2);
  gsum.z += __shfl_xor_sync(0xffffffffu, gsum.z, 1);
  gsum.w += __shfl_xor_sync(0xffffffffu, gsum.w, 4);
  gsum.w += __shfl_xor_sync(0xffffffffu, gsum.w, 2);
  gsum.w += __shfl_xor_sync(0xffffffffu, gsum.w, 1);
  if (item_ok && s == 0) {
    float* out = P.partial + (size_t)blockIdx.x * (P.wsize + P.cout_pad);
    const int cip = ci0 + cil;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int tap = dy * 3 + dx;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = (co0 + col) * 4 + j;
        reinterpret_cast<float4*>(out)[(size_t)(tap * P.cinp + cip) * P.cout_pad + co] =
            make_float4(acc[dx][0][j], acc[dx][1][j], acc[dx][2][j], acc[dx][3][j]);
      }
    }
    if (want_db)
      reinterpret_cast<float4*>(out + P.wsize)[co0 + col] = gsum;
  }
}

// out[e] = sum over blocks of partial[blk][e]; entries of padded output channels are zero.
// Eight lanes share an element (lane j takes blocks j, j+8, ...), partial sums are combined
// in a fixed order by shuffles, so the result does not depend on the run.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int nblk,
                                                           int wsize, int cout_pad, int gp4,
                                                           float* __restrict__ dw, float* __restrict__ db) {
  pdl_trigger();
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = t >> 3, j = t & 7;
  const int total = wsize + cout_pad;
  const bool in = e < total;
  int co = 0;
  if (in) co = e < wsize ? (e >> 2) % cout_pad : e - wsize;
  float s0 = 0.f, s1 = 0.f;
  if (in && co < gp4) {
    int k = j;
    for (; k + 8 < nblk; k += 16) {
      s0 += __ldg(&partial[(size_t)k * total + e]);
      s1 += __ldg(&partial[(size_t)(k + 8) * total + e]);
    }
    if (k < nblk) s0 += __ldg(&partial[(size_t)k * total + e]);
  }
  float sum = s0 + s1;
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  sum += __shfl_xor_sync(0xffffffffu, sum, 4);
  if (in && j == 0) {
    if (e < wsize) dw[e] = sum; else db[co] = sum;
  }
}

static void fill_src_concat(ConvSrc& s, const float* src0, int c0p, int half0, const float* src1,
                            int c1p, int H, int W) {
  s.s0 = reinterpret_cast<const float4*>(src0);
  s.s1 = reinterpret_cast<const float4*>(src1);
  s.sign = nullptr;
  s.p0 = c0p; s.p1 = src1 ? c1p : 0; s.half0 = half0; s.mode = 0; s.slope = 0.f;
  s.H = H; s.W = W; s.Hh = (H + 1) / 2; s.Wh = (W + 1) / 2; s.Wq = (W + 3) / 4;
}

static int fill_src_grad(ConvSrc& s, const float* g_full, const float* g_pool,
                         const uint16_t* sign, float slope, int gp, int H, int W) {
  if (g_full) {
    fill_src_concat(s, g_full, gp, 0, nullptr, 0, H, W);
    return DINCAE_OK;
  }
  if (!g_pool || !sign) return DINCAE_EINVAL;
  fill_src_concat(s, g_pool, gp, 1, nullptr, 0, H, W);
  s.sign = sign; s.mode = 1; s.slope = slope;
  return DINCAE_OK;
}

template <int MODE>
static int launch_conv(const ConvParams& P, cudaStream_t st) {
  const int groups = ceil_div(P.out_planes, CPB);
  if (P.W > 16) {
    dim3 grid(ceil_div(P.W, 32) * ceil_div(P.H, 8), groups, P.B);
    launch_k(conv3x3_kernel<8, 8, MODE>, grid, 256, 0, st, P);
  } else {
    dim3 grid(ceil_div(P.W, 16) * ceil_div(P.H, 16), groups, P.B);
    launch_k(conv3x3_kernel<4, 16, MODE>, grid, 256, 0, st, P);
  }
  return check_launch();
}

static void wgrad_plan(int cinp, int gp, int B, int H, int W, int& cib, int& cob, int& panels_ci,
                       int& panels_co, int& nblk) {
  cob = gp < 5 ? gp : 5;
  cib = 32 / (3 * cob);
  if (cib > cinp) cib = cinp;
  if (cib < 1) cib = 1;
  panels_ci = ceil_div(cinp, cib);
  panels_co = ceil_div(gp, cob);
  const int ntiles = B * ceil_div(W, WG_TW) * ceil_div(H, WG_TH);
  int want = (4 * 148) / (panels_ci * panels_co);
  if (want < 1) want = 1;
  nblk = ntiles < want ? ntiles : want;
  if (nblk < 1) nblk = 1;
}

}  // namespace dincae

using namespace dincae;

namespace dincae {
int g_conv_backend = 1;
}

extern "C" int dincae_set_conv_backend(int backend) {
  if (backend != 0 && backend != 1) return DINCAE_EINVAL;
  g_conv_backend = backend;
  return DINCAE_OK;
}

extern "C" int dincae_get_conv_backend(void) { return g_conv_backend; }

extern "C" int dincae_conv3x3_fwd(const float* src0, int c0p, int src0_half, const float* src1,
                                  int c1p, const float* wpack, const float* bias, int cout_pad,
                                  int B, int H, int W, int coutp, float neg_slope,
                                  float* out_full, float* out_pool, uint16_t* out_sign,
                                  void* stream) {
  if (!src0 || !wpack || c0p <= 0 || B <= 0 || H <= 0 || W <= 0 || coutp <= 0) return DINCAE_EINVAL;
  if (coutp * 4 > cout_pad || (cout_pad & 3) || B > 65535) return DINCAE_EINVAL;
  if (!out_full && !out_pool) return DINCAE_EINVAL;
  ConvParams P = {};
  fill_src_concat(P.src, src0, c0p, src0_half, src1, c1p, H, W);
  P.w = wpack; P.bias = bias; P.w_cinp = P.src.planes(); P.cout_pad = cout_pad;
  P.B = B; P.H = H; P.W = W; P.out_planes = coutp; P.slope = neg_slope;
  P.out_full = reinterpret_cast<float4*>(out_full);
  P.out_pool = reinterpret_cast<float4*>(out_pool);
  P.out_sign = out_sign;
  P.round_out = g_conv_backend == 1;
  if (g_conv_backend == 1) {
    const int rc = launch_conv_tc(0, P, (cudaStream_t)stream);
    if (rc <= 0) return rc;          // 1 = shape not taken by the tensor-core path
  }
  return launch_conv<0>(P, (cudaStream_t)stream);
}

extern "C" int dincae_conv3x3_dgrad(const float* g_full, const float* g_pool, const uint16_t* sign,
                                    float neg_slope, int gp, const float* wpack, int cinp,
                                    int cout_pad, int B, int H, int W, int c_lo,
                                    const float* prev_act, float prev_slope, float* g_prev,
                                    int c_hi, float* dx_hi, const float* dx_add, void* stream) {
  if (!wpack || gp <= 0 || cinp <= 0 || B <= 0 || H <= 0 || W <= 0 || B > 65535) return DINCAE_EINVAL;
  if (gp * 4 > cout_pad || c_lo < 0 || c_hi < 0 || c_lo + c_hi > cinp || c_lo + c_hi == 0)
    return DINCAE_EINVAL;
  if ((c_lo > 0 && !g_prev) || (c_hi > 0 && !dx_hi)) return DINCAE_EINVAL;
  ConvParams P = {};
  int rc = fill_src_grad(P.src, g_full, g_pool, sign, neg_slope, gp, H, W);
  if (rc) return rc;
  P.w = wpack; P.bias = nullptr; P.w_cinp = cinp; P.cout_pad = cout_pad;
  P.B = B; P.H = H; P.W = W; P.out_planes = c_lo + c_hi; P.slope = 0.f;
  P.c_lo = c_lo; P.c_hi = c_hi;
  P.prev_act = reinterpret_cast<const float4*>(prev_act); P.prev_slope = prev_slope;
  P.g_prev = reinterpret_cast<float4*>(g_prev);
  P.dx_hi = reinterpret_cast<float4*>(dx_hi);
  P.dx_add = reinterpret_cast<const float4*>(dx_add);
  P.round_out = g_conv_backend == 1;
  if (g_conv_backend == 1) {
    rc = launch_conv_tc(1, P, (cudaStream_t)stream);
    if (rc <= 0) return rc;
  }
  return launch_conv<1>(P, (cudaStream_t)stream);
}

extern "C" size_t dincae_conv3x3_wgrad_workspace(int c0p, int c1p, int cout_pad, int B, int H,
                                                 int W) {
  // upper bound over every gradient plane count gp <= cout_pad/4
  const int cinp = c0p + c1p;
  size_t worst = 0;
  for (int gp = 1; gp <= cout_pad / 4; ++gp) {
    int cib, cob, pci, pco, nblk;
    wgrad_plan(cinp, gp, B, H, W, cib, cob, pci, pco, nblk);
    const size_t need = (size_t)nblk * ((size_t)9 * cinp * cout_pad * 4 + cout_pad) * sizeof(float);
    if (need > worst) worst = need;
  }
  // tensor-core path: at most one partial per SM (148 on B200; 160 leaves head-room)
  const size_t tc_need = (size_t)160 * ((size_t)9 * cinp * cout_pad * 4 + cout_pad) * sizeof(float);
  if (tc_need > worst) worst = tc_need;
  return align_up(worst, 256);
}

extern "C" int dincae_conv3x3_wgrad_partial(const float* src0, int c0p, int src0_half, const float* src1,
                                            int c1p, const float* g_full, const float* g_pool,
                                            const uint16_t* sign, float neg_slope, int gp, int cout_pad,
                                            int B, int H, int W, void* workspace, size_t workspace_bytes,
                                            int* n_partials, void* stream) {
  if (!src0 || !workspace || !n_partials || c0p <= 0 || gp <= 0 || B <= 0 || H <= 0 || W <= 0)
    return DINCAE_EINVAL;
  if (gp * 4 > cout_pad || (cout_pad & 3)) return DINCAE_EINVAL;
  WgradParams P = {};
  fill_src_concat(P.u, src0, c0p, src0_half, src1, c1p, H, W);
  int rc = fill_src_grad(P.g, g_full, g_pool, sign, neg_slope, gp, H, W);
  if (rc) return rc;
  P.B = B; P.H = H; P.W = W; P.cinp = P.u.planes(); P.gp = gp; P.cout_pad = cout_pad;
  cudaStream_t st = (cudaStream_t)stream;
  P.wsize = 9 * P.cinp * cout_pad * 4;
  if (g_conv_backend == 1) {
    int nb = 0;
    rc = wgrad_tc(P.u, P.g, B, H, W, P.cinp, gp, cout_pad, reinterpret_cast<float*>(workspace),
                  workspace_bytes, &nb, st);
    if (rc < 0) return rc;
    if (rc == 0) {
      *n_partials = nb;
      return DINCAE_OK;
    }
  }
  int nblk;
  wgrad_plan(P.cinp, gp, B, H, W, P.cib, P.cob, P.panels_ci, P.panels_co, nblk);
  const size_t need = (size_t)nblk * ((size_t)P.wsize + cout_pad) * sizeof(float);
  if (need > workspace_bytes) return DINCAE_ENOSPC;
  P.partial = reinterpret_cast<float*>(workspace);
  const size_t smem = ((size_t)P.cib * WG_PH * WG_PW + (size_t)P.cob * WG_TH * WG_TW) * sizeof(float4);
  dim3 grid(nblk, P.panels_ci * P.panels_co);
  launch_k(conv3x3_wgrad_kernel, grid, 256, smem, st, P);
  *n_partials = nblk;
  return check_launch();
}

extern "C" int dincae_conv3x3_wgrad_reduce(const void* workspace, int n_partials, int cinp, int gp,
                                           int cout_pad, float* dw, float* db, void* stream) {
  if (!workspace || !dw || !db || n_partials <= 0 || cinp <= 0 || gp <= 0 || gp * 4 > cout_pad)
    return DINCAE_EINVAL;
  const int wsize = 9 * cinp * cout_pad * 4;
  const int total = wsize + cout_pad;
  launch_k(wgrad_reduce_kernel, ceil_div(total * 8, 256), 256, 0, (cudaStream_t)stream,
           reinterpret_cast<const float*>(workspace), n_partials, wsize, cout_pad, gp * 4, dw, db);
  return check_launch();
}

extern "C" int dincae_conv3x3_wgrad(const float* src0, int c0p, int src0_half, const float* src1,
                                    int c1p, const float* g_full, const float* g_pool,
                                    const uint16_t* sign, float neg_slope, int gp, int cout_pad,
                                    int B, int H, int W, float* dw, float* db, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  if (!dw || !db) return DINCAE_EINVAL;
  int nb = 0;
  int rc = dincae_conv3x3_wgrad_partial(src0, c0p, src0_half, src1, c1p, g_full, g_pool, sign, neg_slope, gp,
                                        cout_pad, B, H, W, workspace, workspace_bytes, &nb, stream);
  if (rc) return rc;
  return dincae_conv3x3_wgrad_reduce(workspace, nb, c0p + (src1 ? c1p : 0), gp, cout_pad, dw, db, stream);
}
