// Dense-bottleneck weight gradients kept in FACTORED form (new capability; replaces the
// materialised MatMul gradient + ApplyAdam of DINCAE/__init__.py:479-483, 598-603).
//
// The gradient of a dense kernel is a rank-B outer product, dW = x^T dz with x [Bt,K] the layer's
// input and dz [Bt,N] the gradient of its pre-activation (Bt = frames of the GLOBAL batch).  At
// BASELINE configs[3] the two dense kernels hold 688 M of the 689 M parameters: writing dW, reading
// it twice (global norm, Adam) and -- data parallel -- all-reducing 2.75 GB per step is what the
// step costs.  Instead:
//   * ||dW||_F^2 = sum_{b,b'} (x_b . x_b') (dz_b . dz_b')   (two Bt x Bt Gram matrices) goes into
//     the global norm of tf.clip_by_global_norm without forming dW;
//   * the Adam kernel recomputes its tile of dW from the factors (Bt FMAs per weight) while it
//     streams p, m, v: 24 bytes per parameter instead of 36 (write dW, read it twice);
//   * data-parallel ranks exchange the FACTORS (an all-gather of Bt x (K+N) floats) instead of
//     all-reducing dW; every rank then applies the same update in the same order.
// Rows of the factors live in `world` per-rank blocks (row bt -> rank bt / rows_per_rank): the
// all-gathered buffer is used in place.
#include "bf16.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <math.h>
#include <stdlib.h>

namespace dincae {

struct Factor {
  const float* base;       // row (r, i) at base + r * rank_stride + i * ld
  size_t rank_stride;      // floats between the blocks of consecutive ranks
  int ld;                  // floats between consecutive rows of a rank
};

__device__ __forceinline__ const float* factor_row(const Factor& f, int bt, int rows_per_rank) {
  const int r = bt / rows_per_rank;
  return f.base + (size_t)r * f.rank_stride + (size_t)(bt - r * rows_per_rank) * f.ld;
}

// ---------------------------------------------------------------------------------------------
// Gram matrices: partial[which][c][b][b'] = sum over the columns of chunk c of f[b][k] f[b'][k].
// grid (chunks, 2, pair tiles): y = 0 -> x (K columns), y = 1 -> dz (N columns); a CTA owns a
// 64 x 64 tile of (b, b') pairs, a thread 4 x 4 of them.
// ---------------------------------------------------------------------------------------------
constexpr int GR_COLS = 32;       // columns staged per pass
constexpr int GR_TILE = 64;       // rows per pair-tile side
constexpr int GR_MAX_BT = 1024;

__global__ void __launch_bounds__(256) gram_partial_kernel(Factor fx, Factor fz, int Bt, int rows_per_rank,
                                                           int K, int N, int cols_per_cta,
                                                           float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  __shared__ float ga[GR_TILE][GR_COLS + 1];
  __shared__ float gb[GR_TILE][GR_COLS + 1];
  const Factor f = blockIdx.y == 0 ? fx : fz;
  const int ncols = blockIdx.y == 0 ? K : N;
  const int c0 = blockIdx.x * cols_per_cta;
  const int nt = (Bt + GR_TILE - 1) / GR_TILE;
  const int ta = blockIdx.z / nt, tb = blockIdx.z - ta * nt;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int c1 = min(ncols, c0 + cols_per_cta);
  for (int cc = c0; cc < c1; cc += GR_COLS) {
    const int w = min(GR_COLS, c1 - cc);
    for (int i = tid; i < GR_TILE * GR_COLS; i += 256) {
      const int r = i / GR_COLS, c = i - r * GR_COLS;
      const int ba = ta * GR_TILE + r, bb = tb * GR_TILE + r;
      ga[r][c] = (c < w && ba < Bt) ? __ldg(factor_row(f, ba, rows_per_rank) + cc + c) : 0.f;
      gb[r][c] = (c < w && bb < Bt) ? __ldg(factor_row(f, bb, rows_per_rank) + cc + c) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int c = 0; c < GR_COLS; ++c) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = ga[ty + 16 * i][c];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = gb[tx + 16 * j][c];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (size_t)Bt * Bt;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ba = ta * GR_TILE + ty + 16 * i, bb = tb * GR_TILE + tx + 16 * j;
      if (ba < Bt && bb < Bt) out[(size_t)ba * Bt + bb] = c0 < ncols ? acc[i][j] : 0.f;
    }
}

// sumsq[0] += sum over pairs of Gx[pair] * Gz[pair], the Gram entries summed over the chunk
// partials.  One thread per pair (its chunk loads are independent and all in flight -- a single
// CTA walking the chunks one dependent load after the other took 75 us at config A), fixed-order
// block sums, and the last CTA adds the block sums in block order: deterministic.
struct GramWork {
  unsigned int counter;
  unsigned int pad[3];
  double block_sum[1];
};

__global__ void __launch_bounds__(256) gram_finish_kernel(const float* __restrict__ partial, int chunks_x,
                                                          int chunks_z, int chunks_stride, int npairs,
                                                          double* __restrict__ sumsq, GramWork* __restrict__ work) {
  pdl_trigger();
  pdl_wait();
  const int pr = blockIdx.x * blockDim.x + threadIdx.x;
  double prod = 0.0;
  if (pr < npairs) {
    float gx0 = 0.f, gx1 = 0.f, gx2 = 0.f, gx3 = 0.f, gz0 = 0.f, gz1 = 0.f, gz2 = 0.f, gz3 = 0.f;
    double gx = 0.0, gz = 0.0;
    const float* px = partial + pr;
    const float* pz = partial + (size_t)chunks_stride * npairs + pr;
    int c = 0;
    for (; c + 4 <= chunks_x; c += 4) {
      gx0 = __ldg(px + (size_t)c * npairs); gx1 = __ldg(px + (size_t)(c + 1) * npairs);
      gx2 = __ldg(px + (size_t)(c + 2) * npairs); gx3 = __ldg(px + (size_t)(c + 3) * npairs);
      gx += ((double)gx0 + (double)gx1) + ((double)gx2 + (double)gx3);
    }
    for (; c < chunks_x; ++c) gx += (double)__ldg(px + (size_t)c * npairs);
    c = 0;
    for (; c + 4 <= chunks_z; c += 4) {
      gz0 = __ldg(pz + (size_t)c * npairs); gz1 = __ldg(pz + (size_t)(c + 1) * npairs);
      gz2 = __ldg(pz + (size_t)(c + 2) * npairs); gz3 = __ldg(pz + (size_t)(c + 3) * npairs);
      gz += ((double)gz0 + (double)gz1) + ((double)gz2 + (double)gz3);
    }
    for (; c < chunks_z; ++c) gz += (double)__ldg(pz + (size_t)c * npairs);
    prod = gx * gz;
  }
  prod = warp_sum(prod);
  __shared__ double red[8];
  __shared__ bool is_last;
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = prod;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    work->block_sum[blockIdx.x] = t;
    __threadfence();
    is_last = atomicAdd(&work->counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (unsigned k = 0; k < gridDim.x; ++k) t += *((volatile double*)&work->block_sum[k]);
    sumsq[0] += t;
    work->counter = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// Adam update of a dense kernel w [K][ldw] (N columns used) from its gradient factors.
// grid (ceil(N/256), ceil(K/16)); 256 threads; thread = 4 rows x 4 columns.  The thread's p, m, v
// (12 float4) are requested BEFORE the factor loop, so their HBM latency hides behind the Bt FMAs
// per weight; ~100 registers -> two CTAs per SM, one computing while the other streams.  (The
// first version -- 8 rows per thread, loads after the FMAs, 193 registers, one CTA per SM -- ran
// at 2.1-3.2 TB/s at configs[3].)
// ---------------------------------------------------------------------------------------------
constexpr int AF_ROWS = 4;                 // rows per thread
constexpr int AF_TK = 4 * AF_ROWS;         // rows per CTA

__global__ void __launch_bounds__(256, 2) adam_factored_kernel(float* __restrict__ w, float* __restrict__ m,
                                                               float* __restrict__ v, int ldw, Factor fx, Factor fz,
                                                               int Bt, int rows_per_rank, int K, int N,
                                                               float grad_scale, const double* __restrict__ denom,
                                                               float beta1, float beta2, float eps,
                                                               const float* __restrict__ state,
                                                               uint2* __restrict__ w8) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float xs[16][AF_TK];
  __shared__ __align__(16) float zs[16][256];
  const int tid = threadIdx.x, tn = tid & 63, tk = tid >> 6;
  const int k0 = blockIdx.y * AF_TK, n0 = blockIdx.x * 256;
  const int n = n0 + 4 * tn;
  const bool col_ok = n < N;
  float4 pp[AF_ROWS], mm[AF_ROWS], vv[AF_ROWS];
#pragma unroll
  for (int i = 0; i < AF_ROWS; ++i) {
    const int k = k0 + AF_ROWS * tk + i;
    const size_t at = (size_t)k * ldw + n;
    const bool ok = col_ok && k < K;
    pp[i] = ldg_f4_pred(reinterpret_cast<const float4*>(w + at), ok);
    mm[i] = ldg_f4_pred(reinterpret_cast<const float4*>(m + at), ok);
    vv[i] = ldg_f4_pred(reinterpret_cast<const float4*>(v + at), ok);
  }
  float4 acc[AF_ROWS];
#pragma unroll
  for (int i = 0; i < AF_ROWS; ++i) acc[i] = f4_zero();
  for (int b0 = 0; b0 < Bt; b0 += 16) {
    if (tid < 16 * AF_TK / 4) {                        // x chunk: 16 rows x AF_TK columns as float4
      const int bb = tid / (AF_TK / 4), c = tid - bb * (AF_TK / 4);
      const int k = k0 + 4 * c;
      const bool ok = b0 + bb < Bt && k < K;
      const float* row = ok ? factor_row(fx, b0 + bb, rows_per_rank) : fx.base;
      *reinterpret_cast<float4*>(&xs[bb][4 * c]) = ldg_f4_pred(reinterpret_cast<const float4*>(row + k), ok);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {                      // dz chunk: 16 x 256 floats as float4
      const int idx = tid + 256 * r;
      const int bb = idx >> 6, c = idx & 63;
      const int nn = n0 + 4 * c;
      const bool ok = b0 + bb < Bt && nn < N;
      const float* row = ok ? factor_row(fz, b0 + bb, rows_per_rank) : fz.base;
      *reinterpret_cast<float4*>(&zs[bb][4 * c]) = ldg_f4_pred(reinterpret_cast<const float4*>(row + nn), ok);
    }
    __syncthreads();
#pragma unroll
    for (int bb = 0; bb < 16; ++bb) {
      const float4 zv = *reinterpret_cast<const float4*>(&zs[bb][4 * tn]);
      const float4 x0 = *reinterpret_cast<const float4*>(&xs[bb][AF_ROWS * tk]);
      acc[0] = f4_fma(x0.x, zv, acc[0]); acc[1] = f4_fma(x0.y, zv, acc[1]);
      acc[2] = f4_fma(x0.z, zv, acc[2]); acc[3] = f4_fma(x0.w, zv, acc[3]);
    }
    __syncthreads();
  }
  if (denom) grad_scale /= (float)(*denom);
  const float gs = grad_scale * state[4];            // 1/n normalisation times the clip scale
  const float lr_t = state[3];
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
#pragma unroll
  for (int i = 0; i < AF_ROWS; ++i) {
    float pa[4] = {pp[i].x, pp[i].y, pp[i].z, pp[i].w};
    float ma[4] = {mm[i].x, mm[i].y, mm[i].z, mm[i].w};
    float va[4] = {vv[i].x, vv[i].y, vv[i].z, vv[i].w};
    const float ga[4] = {acc[i].x, acc[i].y, acc[i].z, acc[i].w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float gk = ga[c] * gs;
      ma[c] = ma[c] + (gk - ma[c]) * omb1;
      va[c] = va[c] + (gk * gk - va[c]) * omb2;
      pa[c] = pa[c] - (ma[c] * lr_t) / (sqrtf(va[c]) + eps);
    }
    const int k = k0 + AF_ROWS * tk + i;
    if (k < K && col_ok) {
      const size_t at = (size_t)k * ldw + n;
      *reinterpret_cast<float4*>(w + at) = make_float4(pa[0], pa[1], pa[2], pa[3]);
      *reinterpret_cast<float4*>(m + at) = make_float4(ma[0], ma[1], ma[2], ma[3]);
      *reinterpret_cast<float4*>(v + at) = make_float4(va[0], va[1], va[2], va[3]);
      // tiled bf16 copy for the tensor-core GEMMs (dense_tc.cu), staged in shared memory (the zs
      // buffer is free by now) so that it leaves the SM as whole 128-byte blocks
      if (w8)
        reinterpret_cast<uint2*>(&zs[0][0])[(((AF_ROWS * tk + i) >> 3) * 32 + tn / 2) * 16 +
                                            ((AF_ROWS * tk + i) & 7) * 2 + (tn & 1)] =
            make_uint2(bf2_pack(pa[0], pa[1]), bf2_pack(pa[2], pa[3]));
    }
  }
  if (w8) {
    // CTA tile = 16 rows x 256 columns = 2 k-blocks x 32 n-blocks of 128 bytes: [kb][nb][8][8] in
    // shared memory, copied out 16 bytes per thread and pass (two passes), whole blocks contiguous
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(&zs[0][0]);
    const int NB = N >> 3;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int idx = tid + 256 * r;                   // 16-byte unit: kb * 256 + nb * 8 + row
      const int kb = idx >> 8, nb = (idx >> 3) & 31, row = idx & 7;
      const int k = k0 + 8 * kb + row, nn = n0 + 8 * nb;
      if (k < K && nn < N)
        reinterpret_cast<uint4*>(w8)[((size_t)(k >> 3) * NB + (nn >> 3)) * 8 + row] = src[idx];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The same update with the rank-Bt product on the tensor cores (mma.sync m16n8k8, TF32 operands,
// fp32 accumulation): data parallel, Bt = world x B frames enter every weight's gradient, and the
// FFMA kernel above turns FMA-bound (128 FMAs per weight at 8 x 16 frames: 6.1 ms per configs[3]
// kernel against 1.9 ms at Bt = 16).  CTA = 16 rows x 256 columns, 8 warps x (16 x 32): per 8
// frames a warp issues 4 MMAs for 12 shared-memory loads.  In the bf16 network the factors are
// bf16 values, which TF32 holds exactly, so the products are exact; fp32 factors are truncated to
// TF32 by the tensor core (used only in the tf32 / bf16 precision modes).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int AM_XP = 24;      // padded row lengths of the staged factors: frames t = 0..3 of a fragment start 24 floats
                               // apart = banks 0, 24, 16, 8 (20 put t = 3 on banks 28..35, a 2-way conflict with t = 0: ncu)
constexpr int AM_ZP = 264;

__global__ void __launch_bounds__(256, 2) adam_factored_mma_kernel(float* __restrict__ w, float* __restrict__ m,
                                                                   float* __restrict__ v, int ldw, Factor fx,
                                                                   Factor fz, int Bt, int rows_per_rank, int K, int N,
                                                                   float grad_scale, const double* __restrict__ denom,
                                                                   float beta1, float beta2, float eps,
                                                                   const float* __restrict__ state,
                                                                   uint32_t* __restrict__ w8) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float xs[16][AM_XP];
  __shared__ __align__(16) float zs[16][AM_ZP];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int k0 = blockIdx.y * 16, n0 = blockIdx.x * 256;
  const int cw = 32 * warp;                                   // this warp's first column in the tile
  // p, m, v of the thread's 2 rows x 4 n-tiles x 2 columns: requested before the factor loop
  float2 pp[2][4], mm[2][4], vv[2][4];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + g + 8 * r, n = n0 + cw + 8 * j + 2 * t;
      const bool ok = k < K && n < N;
      const size_t at = (size_t)k * ldw + n;
      pp[r][j] = ok ? *reinterpret_cast<const float2*>(w + at) : make_float2(0.f, 0.f);
      mm[r][j] = ok ? *reinterpret_cast<const float2*>(m + at) : make_float2(0.f, 0.f);
      vv[r][j] = ok ? *reinterpret_cast<const float2*>(v + at) : make_float2(0.f, 0.f);
    }
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
  for (int b0 = 0; b0 < Bt; b0 += 16) {
    if (tid < 64) {                                           // x chunk: 16 frames x 16 rows as float4
      const int bb = tid >> 2, c = tid & 3;
      const int k = k0 + 4 * c;
      const bool ok = b0 + bb < Bt && k < K;
      const float* row = ok ? factor_row(fx, b0 + bb, rows_per_rank) : fx.base;
      *reinterpret_cast<float4*>(&xs[bb][4 * c]) = ldg_f4_pred(reinterpret_cast<const float4*>(row + k), ok);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {                             // dz chunk: 16 frames x 256 columns as float4
      const int idx = tid + 256 * r;
      const int bb = idx >> 6, c = idx & 63;
      const int nn = n0 + 4 * c;
      const bool ok = b0 + bb < Bt && nn < N;
      const float* row = ok ? factor_row(fz, b0 + bb, rows_per_rank) : fz.base;
      *reinterpret_cast<float4*>(&zs[bb][4 * c]) = ldg_f4_pred(reinterpret_cast<const float4*>(row + nn), ok);
    }
    __syncthreads();
#pragma unroll
    for (int s8 = 0; s8 < 16; s8 += 8) {
      // A = x^T: A[row][frame] = xs[frame][row]; fragment (g, t), (g+8, t), (g, t+4), (g+8, t+4)
      const uint32_t a[4] = {__float_as_uint(xs[s8 + t][g]), __float_as_uint(xs[s8 + t][g + 8]),
                             __float_as_uint(xs[s8 + t + 4][g]), __float_as_uint(xs[s8 + t + 4][g + 8])};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // B[frame][col] = zs[frame][col]; fragment (frame t, col g), (frame t+4, col g)
        const uint32_t bq0 = __float_as_uint(zs[s8 + t][cw + 8 * j + g]);
        const uint32_t bq1 = __float_as_uint(zs[s8 + t + 4][cw + 8 * j + g]);
        mma_tf32_16x8x8(acc[j], a, bq0, bq1);
      }
    }
    __syncthreads();
  }
  if (denom) grad_scale /= (float)(*denom);
  const float gs = grad_scale * state[4];
  const float lr_t = state[3];
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  uint32_t* stage = reinterpret_cast<uint32_t*>(&zs[0][0]);      // [2 kb][32 nb][8 rows][8 n] bf16 = 8 KB
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // accumulator fragment: c0, c1 = (row g, cols 2t, 2t+1); c2, c3 = (row g+8, same columns)
      float pa[2] = {pp[r][j].x, pp[r][j].y}, ma[2] = {mm[r][j].x, mm[r][j].y}, va[2] = {vv[r][j].x, vv[r][j].y};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float gk = acc[j][2 * r + c] * gs;
        ma[c] = ma[c] + (gk - ma[c]) * omb1;
        va[c] = va[c] + (gk * gk - va[c]) * omb2;
        pa[c] = pa[c] - (ma[c] * lr_t) / (sqrtf(va[c]) + eps);
      }
      const int row = g + 8 * r, col = cw + 8 * j + 2 * t;
      const int k = k0 + row, n = n0 + col;
      if (k < K && n < N) {
        const size_t at = (size_t)k * ldw + n;
        *reinterpret_cast<float2*>(w + at) = make_float2(pa[0], pa[1]);
        *reinterpret_cast<float2*>(m + at) = make_float2(ma[0], ma[1]);
        *reinterpret_cast<float2*>(v + at) = make_float2(va[0], va[1]);
      }
      if (w8) stage[((row >> 3) * 32 + (col >> 3)) * 32 + (row & 7) * 4 + ((col & 7) >> 1)] = bf2_pack(pa[0], pa[1]);
    }
  if (w8) {
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(&zs[0][0]);
    const int NB = N >> 3;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int idx = tid + 256 * r;                   // 16-byte unit: kb * 256 + nb * 8 + row
      const int kb = idx >> 8, nb = (idx >> 3) & 31, row = idx & 7;
      const int k = k0 + 8 * kb + row, nn = n0 + 8 * nb;
      if (k < K && nn < N)
        reinterpret_cast<uint4*>(w8)[((size_t)(k >> 3) * NB + (nn >> 3)) * 8 + row] = src[idx];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Persistent variant for 16 < Bt <= 128: a CTA owns a block of 256 columns, stages dz[:, block]
// ONCE (Bt x 256 floats in shared memory) and walks down its share of the rows; per 16-row tile
// only the x tile (Bt x 16) is new, and it is prefetched into registers during the previous
// tile together with the tile's p, m, v.  (The per-tile staging of the kernel above re-read the dz
// block for every 16 rows -- 139 KB of factors per 106 KB of weight traffic at Bt = 128 -- with one
// exposed load latency per 16 frames: 6.3 ms per configs[3] kernel, no faster than FFMA.)
// ---------------------------------------------------------------------------------------------
constexpr int AP_MAX_BT = 128;

__global__ void __launch_bounds__(256, 1) adam_factored_persist_kernel(float* __restrict__ w, float* __restrict__ m,
                                                                       float* __restrict__ v, int ldw, Factor fx,
                                                                       Factor fz, int Bt, int rows_per_rank, int K,
                                                                       int N, int tiles_per_cta, float grad_scale,
                                                                       const double* __restrict__ denom, float beta1,
                                                                       float beta2, float eps,
                                                                       const float* __restrict__ state,
                                                                       uint32_t* __restrict__ w8) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float ap_smem[];
  const int Bp = (Bt + 7) & ~7;
  float* zs = ap_smem;                                  // [Bp][AM_ZP]
  float* xs0 = zs + (size_t)Bp * AM_ZP;                 // [2][Bp][AM_XP]
  uint32_t* stage = reinterpret_cast<uint32_t*>(xs0 + 2 * (size_t)Bp * AM_XP);     // 8 KB
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * 256;
  const int cw = 32 * warp;
  const int n_tiles = (K + 15) >> 4;
  const int tile0 = blockIdx.y * tiles_per_cta;
  int tile1 = tile0 + tiles_per_cta;
  if (tile1 > n_tiles) tile1 = n_tiles;
  // ---- dz block, once ----------------------------------------------------------------------------
  for (int idx = tid; idx < Bp * 64; idx += 256) {
    const int bb = idx >> 6, c = idx & 63;
    const int nn = n0 + 4 * c;
    const bool ok = bb < Bt && nn < N;
    const float* row = ok ? factor_row(fz, bb, rows_per_rank) : fz.base;
    *reinterpret_cast<float4*>(&zs[(size_t)bb * AM_ZP + 4 * c]) = ldg_f4_pred(reinterpret_cast<const float4*>(row + nn), ok);
  }
  // x tile loader: element e = tid + 256 * u  ->  frame e / 4, float4 column e % 4
  auto load_x = [&](int tile, float4 (&xr)[2]) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int e = tid + 256 * u;
      const int bb = e >> 2, c = e & 3;
      const int k = 16 * tile + 4 * c;
      const bool ok = bb < Bt && k < K && tile < tile1;
      const float* row = ok ? factor_row(fx, bb, rows_per_rank) : fx.base;
      xr[u] = ldg_f4_pred(reinterpret_cast<const float4*>(row + k), ok);
    }
  };
  auto store_x = [&](int buf, const float4 (&xr)[2]) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int e = tid + 256 * u;
      const int bb = e >> 2, c = e & 3;
      if (bb < Bp) *reinterpret_cast<float4*>(&xs0[((size_t)buf * Bp + bb) * AM_XP + 4 * c]) = xr[u];
    }
  };
  float gscale = grad_scale;
  if (denom) gscale /= (float)(*denom);
  const float gs = gscale * state[4];
  const float lr_t = state[3];
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  float4 xr[2];
  load_x(tile0, xr);
  store_x(0, xr);
  __syncthreads();
  int buf = 0;
  for (int tile = tile0; tile < tile1; ++tile, buf ^= 1) {
    const int k0 = 16 * tile;
    float2 pp[2][4], mm[2][4], vv[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + g + 8 * r, n = n0 + cw + 8 * j + 2 * t;
        const bool ok = k < K && n < N;
        const size_t at = (size_t)k * ldw + n;
        pp[r][j] = ok ? *reinterpret_cast<const float2*>(w + at) : make_float2(0.f, 0.f);
        mm[r][j] = ok ? *reinterpret_cast<const float2*>(m + at) : make_float2(0.f, 0.f);
        vv[r][j] = ok ? *reinterpret_cast<const float2*>(v + at) : make_float2(0.f, 0.f);
      }
    load_x(tile + 1, xr);                                  // next tile's x: in flight during the MMAs
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
    const float* xs = xs0 + (size_t)buf * Bp * AM_XP;
    for (int s8 = 0; s8 < Bp; s8 += 8) {
      const uint32_t a[4] = {__float_as_uint(xs[(s8 + t) * AM_XP + g]), __float_as_uint(xs[(s8 + t) * AM_XP + g + 8]),
                             __float_as_uint(xs[(s8 + t + 4) * AM_XP + g]),
                             __float_as_uint(xs[(s8 + t + 4) * AM_XP + g + 8])};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t bq0 = __float_as_uint(zs[(size_t)(s8 + t) * AM_ZP + cw + 8 * j + g]);
        const uint32_t bq1 = __float_as_uint(zs[(size_t)(s8 + t + 4) * AM_ZP + cw + 8 * j + g]);
        mma_tf32_16x8x8(acc[j], a, bq0, bq1);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float pa[2] = {pp[r][j].x, pp[r][j].y}, ma[2] = {mm[r][j].x, mm[r][j].y}, va[2] = {vv[r][j].x, vv[r][j].y};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float gk = acc[j][2 * r + c] * gs;
          ma[c] = ma[c] + (gk - ma[c]) * omb1;
          va[c] = va[c] + (gk * gk - va[c]) * omb2;
          pa[c] = pa[c] - (ma[c] * lr_t) / (sqrtf(va[c]) + eps);
        }
        const int row = g + 8 * r, col = cw + 8 * j + 2 * t;
        const int k = k0 + row, n = n0 + col;
        if (k < K && n < N) {
          const size_t at = (size_t)k * ldw + n;
          *reinterpret_cast<float2*>(w + at) = make_float2(pa[0], pa[1]);
          *reinterpret_cast<float2*>(m + at) = make_float2(ma[0], ma[1]);
          *reinterpret_cast<float2*>(v + at) = make_float2(va[0], va[1]);
        }
        if (w8) stage[((row >> 3) * 32 + (col >> 3)) * 32 + (row & 7) * 4 + ((col & 7) >> 1)] = bf2_pack(pa[0], pa[1]);
      }
    store_x(buf ^ 1, xr);
    __syncthreads();                                       // next x tile and the staged bf16 tile are complete
    if (w8) {
      const uint4* src = reinterpret_cast<const uint4*>(stage);
      const int NB = N >> 3;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int idx = tid + 256 * r;
        const int kb = idx >> 8, nb = (idx >> 3) & 31, row = idx & 7;
        const int k = k0 + 8 * kb + row, nn = n0 + 8 * nb;
        if (k < K && nn < N)
          reinterpret_cast<uint4*>(w8)[((size_t)(k >> 3) * NB + (nn >> 3)) * 8 + row] = src[idx];
      }
      __syncthreads();                                     // the stage buffer is free again
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The single-GPU update (Bt <= 16) as a TMA pipeline.  adam_factored_kernel above keeps p, m, v
// of a tile in registers between the load and the store, so a CTA has no bytes in flight while it
// computes and the kernel tops out at 4.4-4.9 TB/s (67-75 % of the measured copy bandwidth).
// Here one persistent CTA per SM walks a contiguous range of 16 x 256 tiles in ROW-MAJOR order
// (along the columns of a 16-row block, then the next block: every CTA is a contiguous stream
// through w, m and v -- walking down a column block instead touches 1 KB segments 33 KB apart and
// measured 2.5 TB/s):
//   * thread 0 keeps the p / m / v tiles of the NEXT TWO tiles in flight (3 stages x 48 KB,
//     cp.async.bulk.tensor loads on an mbarrier per stage; out-of-range rows / columns are
//     zero-filled by the tensor map and clipped again by the store);
//   * the 512 threads form the gradient tile from the factors while those loads fly -- the x
//     values of a thread's two rows stay in REGISTERS for the whole row block (32 registers;
//     the CTA has the SM to itself), the 16 x 256 dz chunk of the next tile arrives one tile
//     ahead by cp.async into a double-buffered 16 KB array;
//   * p, m, v are updated in place in shared memory and leave through cp.async.bulk.tensor stores,
//     the tiled bf16 copy through two 4 KB bulk copies; a stage is refilled only after the
//     store that read it has finished reading (cp.async.bulk.wait_group.read).
// Measured at configs[3] (43008 x 8296 / 8296 x 43008, 16 frames): 1.58 / 1.50 ms = 5.9 / 6.2 TB/s
// (90 / 94 % of the 6.55 TB/s copy peak; the same pipeline with the arithmetic removed: 5.7 / 6.2)
// against 2.09 / 1.93 ms for adam_factored_kernel.  256 threads x 4 rows: 4.4 TB/s (issue bound),
// 1024 x 1: 4.9-5.2 TB/s; L2 evict-first hints on the loads / stores: no change.
// Summation order and the m / v updates are those of adam_factored_kernel (bit-identical m, v);
// the weight step uses the approximate sqrt / division below.
// ---------------------------------------------------------------------------------------------
constexpr int AT_STAGES = 3;
constexpr int AT_TILE = 16 * 256;                                   // floats of one array's tile
constexpr int AT_OFF_W8 = AT_STAGES * 3 * AT_TILE * 4;               // 2 x 8 KB bf16 staging
constexpr int AT_OFF_ZS = AT_OFF_W8 + 2 * 8192;                      // 2 x [16][256] floats
constexpr int AT_OFF_BAR = AT_OFF_ZS + 2 * AT_TILE * 4;
constexpr int AT_SMEM = AT_OFF_BAR + 64 + 128;                       // + alignment slack

__device__ __forceinline__ void cp_async16_zfill(void* dst_smem, const void* src, bool ok) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
               :: "r"(smem_u32(dst_smem)), "l"(src), "r"(ok ? 16 : 0) : "memory");
}

// MUFU.SQRT / MUFU.RCP forms (relative error 2^-23 and 2 ulp): the IEEE sqrtf and division of the
// other Adam kernels cost ~30 instructions per weight with slow-path branches and made this
// kernel issue bound (measured 5.4 TB/s against 5.7-6.1 TB/s with the arithmetic removed); the
// step lr * m / (sqrt(v) + eps) changes by < 4e-7 of itself, below one ulp of the weight.
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int AT_ROWS = 2;                                          // rows of the tile per thread
constexpr int AT_THREADS = 64 * (16 / AT_ROWS);

__global__ void __launch_bounds__(AT_THREADS, 1) adam_factored_tma_kernel(
    const __grid_constant__ CUtensorMap tmw, const __grid_constant__ CUtensorMap tmm,
    const __grid_constant__ CUtensorMap tmv, Factor fx, Factor fz, int Bt, int rows_per_rank, int K, int N,
    float grad_scale, const double* __restrict__ denom, float beta1, float beta2, float eps,
    const float* __restrict__ state, uint4* __restrict__ w8) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ unsigned char at_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(at_raw) + 127) & ~(uintptr_t)127);
  float* tiles = reinterpret_cast<float*>(base);
  uint2* w8s = reinterpret_cast<uint2*>(base + AT_OFF_W8);
  float* zs = reinterpret_cast<float*>(base + AT_OFF_ZS);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + AT_OFF_BAR);
  const int tid = threadIdx.x, tn = tid & 63, tk = tid >> 6;
  const int RT = (K + 15) >> 4, CB = (N + 255) >> 8, NB = N >> 3;
  const long long T = (long long)RT * CB;
  const long long t0 = T * blockIdx.x / gridDim.x, t1 = T * (blockIdx.x + 1) / gridDim.x;
  if (t0 >= t1) return;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < AT_STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue_load = [&](long long t, int s) {                          // thread 0
    const int rt = (int)(t / CB), cb = (int)(t - (long long)rt * CB);
    float* dst = tiles + (size_t)s * 3 * AT_TILE;
    mbar_expect_tx(&full[s], 3 * AT_TILE * 4);
    tma_load_2d(dst, &tmw, cb * 256, rt * 16, &full[s]);
    tma_load_2d(dst + AT_TILE, &tmm, cb * 256, rt * 16, &full[s]);
    tma_load_2d(dst + 2 * AT_TILE, &tmv, cb * 256, rt * 16, &full[s]);
    mbar_arrive(&full[s]);
  };
  if (tid == 0) {
    issue_load(t0, 0);
    if (t0 + 1 < t1) issue_load(t0 + 1, 1);
  }
  // dz chunk of a tile: 16 frames x 256 columns, four 16-byte cp.async per thread (zero fill
  // outside the batch / the matrix)
  auto fetch_z = [&](int cb, int buf, bool valid) {
#pragma unroll
    for (int r = 0; r < 1024 / AT_THREADS; ++r) {
      const int idx = tid + AT_THREADS * r;
      const int bb = idx >> 6, c = idx & 63;
      const int nn = cb * 256 + 4 * c;
      const bool ok = valid && bb < Bt && nn < N;
      const float* row = ok ? factor_row(fz, bb, rows_per_rank) + nn : fz.base;
      cp_async16_zfill(zs + buf * AT_TILE + bb * 256 + 4 * c, row, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float gscale = grad_scale;
  if (denom) gscale /= (float)(*denom);
  const float gs = gscale * state[4];                 // 1/n normalisation times the clip scale
  const float lr_t = state[3];
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  int rt = (int)(t0 / CB), cb = (int)(t0 - (long long)rt * CB);
  fetch_z(cb, 0, true);
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  float x[16][AT_ROWS];
  int x_rt = -1;
  int i = 0;
  for (long long t = t0; t < t1; ++t, ++i) {
    const int s = i % AT_STAGES;
    const uint32_t ph = (uint32_t)(i / AT_STAGES) & 1u;
    const int n0 = cb * 256, k0 = rt * 16;
    int rt_n = rt, cb_n = cb + 1;
    if (cb_n == CB) { cb_n = 0; ++rt_n; }
    fetch_z(cb_n, (i + 1) & 1, t + 1 < t1);            // next tile's dz chunk: lands during this tile
    if (rt != x_rt) {                                  // x of this thread's rows, all frames
      const int k = k0 + AT_ROWS * tk;
#pragma unroll
      for (int bb = 0; bb < 16; ++bb) {
        const bool ok = bb < Bt && k < K;                // K is a multiple of 4: the AT_ROWS rows are in or out together
        const float* row = ok ? factor_row(fx, bb, rows_per_rank) : fx.base;
#pragma unroll
        for (int r = 0; r < AT_ROWS; ++r) x[bb][r] = ok ? __ldg(row + k + r) : 0.f;
      }
      x_rt = rt;
    }
    float4 acc[AT_ROWS];
#pragma unroll
    for (int r = 0; r < AT_ROWS; ++r) acc[r] = f4_zero();
    const float* zcur = zs + (i & 1) * AT_TILE;
#pragma unroll
    for (int bb = 0; bb < 16; ++bb) {
      const float4 zv = *reinterpret_cast<const float4*>(&zcur[bb * 256 + 4 * tn]);
#pragma unroll
      for (int r = 0; r < AT_ROWS; ++r) acc[r] = f4_fma(x[bb][r], zv, acc[r]);
    }
    mbar_wait(&full[s], ph);
    float* tw = tiles + (size_t)s * 3 * AT_TILE;
    uint2* wst = w8s + (i & 1) * 1024;
#pragma unroll
    for (int r = 0; r < AT_ROWS; ++r) {
      const int row = AT_ROWS * tk + r, at = row * 256 + 4 * tn;
      const float4 p4 = *reinterpret_cast<const float4*>(tw + at);
      const float4 m4 = *reinterpret_cast<const float4*>(tw + AT_TILE + at);
      const float4 v4 = *reinterpret_cast<const float4*>(tw + 2 * AT_TILE + at);
      float pa[4] = {p4.x, p4.y, p4.z, p4.w}, ma[4] = {m4.x, m4.y, m4.z, m4.w}, va[4] = {v4.x, v4.y, v4.z, v4.w};
      const float ga[4] = {acc[r].x, acc[r].y, acc[r].z, acc[r].w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float gk = ga[c] * gs;
        ma[c] = ma[c] + (gk - ma[c]) * omb1;
        va[c] = va[c] + (gk * gk - va[c]) * omb2;
        pa[c] = pa[c] - __fdividef(ma[c] * lr_t, sqrt_approx(va[c]) + eps);
      }
      *reinterpret_cast<float4*>(tw + at) = make_float4(pa[0], pa[1], pa[2], pa[3]);
      *reinterpret_cast<float4*>(tw + AT_TILE + at) = make_float4(ma[0], ma[1], ma[2], ma[3]);
      *reinterpret_cast<float4*>(tw + 2 * AT_TILE + at) = make_float4(va[0], va[1], va[2], va[3]);
      if (w8)                                          // [kb][nb][8 rows][8 columns] bf16
        wst[((row >> 3) * 32 + (tn >> 1)) * 16 + (row & 7) * 2 + (tn & 1)] =
            make_uint2(bf2_pack(pa[0], pa[1]), bf2_pack(pa[2], pa[3]));
    }
    asm volatile("cp.async.wait_all;" ::: "memory");  // next dz chunk
    fence_async_smem();                                // the updated tiles become visible to the TMA engine
    if (tid == 0) bulk_wait_read<0>();                 // the previous tile's stores have left shared memory
    __syncthreads();
    if (tid == 0) {
      tma_store_2d(&tmw, n0, k0, tw);
      tma_store_2d(&tmm, n0, k0, tw + AT_TILE);
      tma_store_2d(&tmv, n0, k0, tw + 2 * AT_TILE);
      if (w8) {
        const int nbs = min(32, NB - (n0 >> 3));
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
          if (k0 + 8 * kb < K)
            bulk_s2g(w8 + ((size_t)((k0 >> 3) + kb) * NB + (n0 >> 3)) * 8, wst + kb * 512, (uint32_t)nbs * 128u);
      }
      bulk_commit();
      if (t + 2 < t1) issue_load(t + 2, (i + 2) % AT_STAGES);
    }
    cb = cb_n;
    rt = rt_n;
  }
  if (tid == 0) bulk_wait<0>();
}

static int at_encode_map(CUtensorMap* tm, const float* base, int K, int N, int ldw) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return DINCAE_ECUDA;
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)K};
  const cuuint64_t strides[1] = {(cuuint64_t)ldw * 4};
  const cuuint32_t box[2] = {256, 16};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_cuda_error = (int)r;
    return DINCAE_ECUDA;
  }
  return DINCAE_OK;
}

// ---------------------------------------------------------------------------------------------
// The data-parallel update (16 < Bt <= 128 frames) on the same TMA pipeline, gradient tile by
// mma.sync (TF32, as adam_factored_mma_kernel).  Differences to the kernel above:
//   * the dz operand of Bt frames x 256 columns does not fit next to the p / m / v stages, so the
//     walk is DOWN a column block and each thread keeps its B fragments (dz of the warp's 16
//     columns, all frames: 4 Bt/8 registers) for the whole task; the 16 x Bt x-chunk of the next
//     tile arrives by cp.async.  To keep DRAM pages open the tasks (row range, column block) are
//     dealt column block fastest over the CTAs, so the CTAs of a wave walk down adjacent column
//     blocks of the same rows in step;
//   * accumulator fragments own (row g / g+8, columns 2t, 2t+1): on a row-major tile those are
//     8-way bank conflicts, so the p / m / v tiles travel as eight 16 x 32 boxes with the 128-byte
//     TMA swizzle (16-byte chunk ^= row % 8), which makes the float2 accesses conflict free;
//   * a 17th warp issues the 48 tensor-map loads / stores per tile (the compute warps hand a
//     finished stage over through an mbarrier), w8 staging is per stage.
// ---------------------------------------------------------------------------------------------
constexpr int AX_XP = 24;                                 // padded frame pitch of the x chunk (floats; see AM_XP)
constexpr int AX_CWARPS = 16;
constexpr int AX_THREADS = 32 * AX_CWARPS + 32;
constexpr int AX_OFF_W8 = AT_STAGES * 3 * AT_TILE * 4;     // 3 x 8 KB bf16 staging (one per stage)
constexpr int AX_OFF_XS = AX_OFF_W8 + AT_STAGES * 8192;
__host__ __device__ constexpr int ax_smem_bytes(int nk) { return AX_OFF_XS + 2 * nk * 8 * AX_XP * 4 + 64 + 1024; }

struct AxWalk {                                            // this CTA's tile sequence
  int CB, RT, TPS, ntasks, G;
  int task, cb, rt, rt1;
  __device__ void start(int cta) { task = cta - G; rt = 0; rt1 = 0; next(); }
  __device__ bool valid() const { return task < ntasks; }
  __device__ void next() {
    if (++rt < rt1) return;
    task += G;
    if (task >= ntasks) return;
    const int rs = task / CB;
    cb = task - rs * CB;
    rt = rs * TPS;
    rt1 = min(RT, rt + TPS);
  }
};

template <int NK>
__global__ void __launch_bounds__(AX_THREADS, 1) adam_factored_tma_mma_kernel(
    const __grid_constant__ CUtensorMap tmw, const __grid_constant__ CUtensorMap tmm,
    const __grid_constant__ CUtensorMap tmv, Factor fx, Factor fz, int Bt, int rows_per_rank, int K, int N, int RS,
    float grad_scale, const double* __restrict__ denom, float beta1, float beta2, float eps,
    const float* __restrict__ state, uint4* __restrict__ w8) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ unsigned char ax_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ax_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int BP = NK * 8;
  float* xs = reinterpret_cast<float*>(base + AX_OFF_XS);                  // [2][BP][AX_XP]
  uint64_t* full = reinterpret_cast<uint64_t*>(base + AX_OFF_XS + 2 * BP * AX_XP * 4);
  uint64_t* done = full + AT_STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  AxWalk wk;
  wk.CB = (N + 255) >> 8; wk.RT = (K + 15) >> 4; wk.TPS = (wk.RT + RS - 1) / RS;
  wk.ntasks = wk.CB * RS; wk.G = gridDim.x;
  const int NB = N >> 3;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < AT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], AX_CWARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == AX_CWARPS) {
    // ---------------- TMA warp: loads two tiles ahead, stores behind ---------------------------
    if (lane != 0) return;
    auto issue_load = [&](const AxWalk& a, int s) {
      const int n0 = a.cb * 256, k0 = a.rt * 16;
      const int nbox = min(8, (N - n0 + 31) >> 5);
      unsigned char* dst = base + (size_t)s * 3 * AT_TILE * 4;
      mbar_expect_tx(&full[s], (uint32_t)(3 * nbox * 2048));
      for (int b = 0; b < nbox; ++b) {
        tma_load_2d(dst + b * 2048, &tmw, n0 + 32 * b, k0, &full[s]);
        tma_load_2d(dst + AT_TILE * 4 + b * 2048, &tmm, n0 + 32 * b, k0, &full[s]);
        tma_load_2d(dst + 2 * AT_TILE * 4 + b * 2048, &tmv, n0 + 32 * b, k0, &full[s]);
      }
      mbar_arrive(&full[s]);
    };
    AxWalk ld = wk, st = wk;
    ld.start(blockIdx.x);
    st.start(blockIdx.x);
    for (int q = 0; q < 2 && ld.valid(); ++q) { issue_load(ld, q); ld.next(); }
    for (int i = 0; st.valid(); ++i) {
      const int s = i % AT_STAGES;
      mbar_wait(&done[s], (uint32_t)(i / AT_STAGES) & 1u);
      const int n0 = st.cb * 256, k0 = st.rt * 16;
      const int nbox = min(8, (N - n0 + 31) >> 5);
      const unsigned char* src = base + (size_t)s * 3 * AT_TILE * 4;
      for (int b = 0; b < nbox; ++b) {
        tma_store_2d(&tmw, n0 + 32 * b, k0, src + b * 2048);
        tma_store_2d(&tmm, n0 + 32 * b, k0, src + AT_TILE * 4 + b * 2048);
        tma_store_2d(&tmv, n0 + 32 * b, k0, src + 2 * AT_TILE * 4 + b * 2048);
      }
      if (w8) {
        const int nbs = min(32, NB - (n0 >> 3));
        const unsigned char* wsrc = base + AX_OFF_W8 + s * 8192;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
          if (k0 + 8 * kb < K)
            bulk_s2g(w8 + ((size_t)((k0 >> 3) + kb) * NB + (n0 >> 3)) * 8, wsrc + kb * 4096, (uint32_t)nbs * 128u);
      }
      bulk_commit();
      st.next();
      if (ld.valid()) {
        bulk_wait_read<1>();                           // the stage of tile i - 1 has been read out
        issue_load(ld, (i + 2) % AT_STAGES);
        ld.next();
      }
    }
    bulk_wait<0>();
    return;
  }

  // ---------------- compute warps ------------------------------------------------------------------
  const int g = lane >> 2, t = lane & 3;
  // columns of this warp's two n-tiles: a warp pair shares a 32-column box, the even warp takes box
  // columns 0-7 and 16-23, the odd warp 8-15 and 24-31 (n-tile j at cw + 16 j) -- see toff below
  const int cw = 32 * (warp >> 1) + 8 * (warp & 1);
  float gscale = grad_scale;
  if (denom) gscale /= (float)(*denom);
  const float gs = gscale * state[4];
  const float lr_t = state[3];
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  // x chunk of a tile: BP frames x 16 rows, one 16-byte cp.async per thread (zero fill outside)
  auto fetch_x = [&](int rt, int buf, bool valid) {
    if (tid < BP * 4) {
      const int bb = tid >> 2, c = tid & 3;
      const int k = 16 * rt + 4 * c;
      const bool ok = valid && bb < Bt && k < K;
      const float* src = ok ? factor_row(fx, bb, rows_per_rank) + k : fx.base;
      cp_async16_zfill(xs + ((size_t)buf * BP + bb) * AX_XP + 4 * c, src, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  AxWalk it = wk;
  it.start(blockIdx.x);
  if (!it.valid()) return;
  fetch_x(it.rt, 0, true);
  asm volatile("cp.async.wait_all;" ::: "memory");
  named_bar_sync(1, 32 * AX_CWARPS);
  uint32_t bfr[NK][2][2];
  int cur_task = -1;
  // Swizzled position of this thread's float2 in a 16 x 256 tile (eight 16 x 32 boxes of 2 KB,
  // 128-byte rows, 16-byte chunk index ^= row % 8), rows g and g + 8.  In pass q the lanes with
  // t < 2 work on n-tile q and the lanes with t >= 2 on n-tile q ^ 1: the two halves of a half
  // warp then touch chunks of different 4-chunk groups (the n-tiles are 16 columns apart), the
  // row phases g = 0..3 spread each over its group, and the 64-bit accesses are conflict free.
  // (Both halves on the same n-tile hit adjacent chunks c, c + 1, which the same four phases map
  // onto the same four positions: 2-way conflicts on every access, profiles/r2_adam_ncu.md.)
  int toff[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int j = q ^ (t >> 1);
    const int chunk = 2 * (warp & 1) + 4 * j + (t >> 1);
    toff[q] = (warp >> 1) * 2048 + g * 128 + ((chunk ^ g) << 4) + (t & 1) * 8;
  }
  const bool hi_half = (t >> 1) != 0;
  for (int i = 0; it.valid(); ++i) {
    const int s = i % AT_STAGES;
    if (it.task != cur_task) {                                // B fragments: dz of the warp's columns, all frames
      const int n0 = it.cb * 256;
#pragma unroll
      for (int ks = 0; ks < NK; ++ks)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int f = 8 * ks + t + 4 * h, col = n0 + cw + 16 * j + g;
            const bool ok = f < Bt && col < N;
            bfr[ks][j][h] = ok ? __float_as_uint(__ldg(factor_row(fz, f, rows_per_rank) + col)) : 0u;
          }
      cur_task = it.task;
    }
    AxWalk nx = it;
    nx.next();
    fetch_x(nx.rt, (i + 1) & 1, nx.valid());                  // next tile's x chunk: lands during this tile
    float acc[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
    const float* xc = xs + (size_t)(i & 1) * BP * AX_XP;
#pragma unroll
    for (int ks = 0; ks < NK; ++ks) {
      // A = x^T: A[row][frame] = xc[frame][row]; fragment (g, t), (g+8, t), (g, t+4), (g+8, t+4)
      const uint32_t a[4] = {__float_as_uint(xc[(8 * ks + t) * AX_XP + g]), __float_as_uint(xc[(8 * ks + t) * AX_XP + g + 8]),
                             __float_as_uint(xc[(8 * ks + t + 4) * AX_XP + g]),
                             __float_as_uint(xc[(8 * ks + t + 4) * AX_XP + g + 8])};
#pragma unroll
      for (int j = 0; j < 2; ++j) mma_tf32_16x8x8(acc[j], a, bfr[ks][j][0], bfr[ks][j][1]);
    }
    mbar_wait(&full[s], (uint32_t)(i / AT_STAGES) & 1u);
    unsigned char* tw = base + (size_t)s * 3 * AT_TILE * 4;
    uint32_t* wst = reinterpret_cast<uint32_t*>(base + AX_OFF_W8 + s * 8192);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int j = q ^ (hi_half ? 1 : 0);                  // this lane's n-tile in pass q
        const int at = toff[q] + r * 1024;                    // row g + 8 r: same swizzle phase
        const float2 p2 = *reinterpret_cast<const float2*>(tw + at);
        const float2 m2 = *reinterpret_cast<const float2*>(tw + AT_TILE * 4 + at);
        const float2 v2 = *reinterpret_cast<const float2*>(tw + 2 * AT_TILE * 4 + at);
        float pa[2] = {p2.x, p2.y}, ma[2] = {m2.x, m2.y}, va[2] = {v2.x, v2.y};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float gk = ((q == 0) != hi_half ? acc[0][2 * r + c] : acc[1][2 * r + c]) * gs;
          ma[c] = ma[c] + (gk - ma[c]) * omb1;
          va[c] = va[c] + (gk * gk - va[c]) * omb2;
          pa[c] = pa[c] - __fdividef(ma[c] * lr_t, sqrt_approx(va[c]) + eps);
        }
        *reinterpret_cast<float2*>(tw + at) = make_float2(pa[0], pa[1]);
        *reinterpret_cast<float2*>(tw + AT_TILE * 4 + at) = make_float2(ma[0], ma[1]);
        *reinterpret_cast<float2*>(tw + 2 * AT_TILE * 4 + at) = make_float2(va[0], va[1]);
        const int row = g + 8 * r, col = cw + 16 * j + 2 * t;
        if (w8) wst[((row >> 3) * 32 + (col >> 3)) * 32 + (row & 7) * 4 + ((col & 7) >> 1)] = bf2_pack(pa[0], pa[1]);
      }
    asm volatile("cp.async.wait_all;" ::: "memory");          // next x chunk
    fence_async_smem();                                        // updated tile visible to the TMA engine
    __syncwarp();
    if (lane == 0) mbar_arrive(&done[s]);
    named_bar_sync(1, 32 * AX_CWARPS);                         // x double buffer
    it = nx;
  }
}

static int ax_encode_map(CUtensorMap* tm, const float* base, int K, int N, int ldw) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return DINCAE_ECUDA;
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)K};
  const cuuint64_t strides[1] = {(cuuint64_t)ldw * 4};
  const cuuint32_t box[2] = {32, 16};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_cuda_error = (int)r;
    return DINCAE_ECUDA;
  }
  return DINCAE_OK;
}

// row splits of the task grid: the smallest count whose waves are >= 96 % full (tasks are
// column blocks x row splits, dealt round robin over the CTAs), at least 8 tiles per task
static int ax_row_splits(int CB, int RT, int G) {
  int best = 1;
  double best_eff = 0.0;
  for (int rs = 1; rs <= RT && rs <= 4096; ++rs) {
    if ((RT + rs - 1) / rs < 8 && rs > 1) break;
    const long long tasks = (long long)CB * rs;
    const long long waves = (tasks + G - 1) / G;
    const double eff = (double)tasks / (double)(waves * G);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = rs; }
    if (eff >= 0.96) return rs;
  }
  return best;
}

template <int NK>
static int ax_launch(int sm_count, const CUtensorMap& tmw, const CUtensorMap& tmm, const CUtensorMap& tmv, Factor fx,
                     Factor fz, int Bt, int rows_per_rank, int K, int N, float grad_scale, const double* denom,
                     float beta1, float beta2, float eps, const float* state, void* w8, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(adam_factored_tma_mma_kernel<NK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         ax_smem_bytes(NK));
    if (e != cudaSuccess) return cuda_rc(e);
    attr_set = true;
  }
  const int RS = ax_row_splits(ceil_div(N, 256), ceil_div(K, 16), sm_count);
  launch_k(adam_factored_tma_mma_kernel<NK>, sm_count, AX_THREADS, ax_smem_bytes(NK), st, tmw, tmm, tmv, fx, fz, Bt,
           rows_per_rank, K, N, RS, grad_scale, denom, beta1, beta2, eps, state, reinterpret_cast<uint4*>(w8));
  return DINCAE_OK;
}

// dw = x^T dz written out (tests / callers that want the matrix): the same tiling without Adam
__global__ void __launch_bounds__(256) factored_expand_kernel(float* __restrict__ dw, int ldw, Factor fx, Factor fz,
                                                              int Bt, int rows_per_rank, int K, int N) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.x * 256 + threadIdx.x, k = blockIdx.y;
  if (n >= N || k >= K) return;
  float s = 0.f;
  for (int b = 0; b < Bt; ++b)
    s = fmaf(__ldg(factor_row(fx, b, rows_per_rank) + k), __ldg(factor_row(fz, b, rows_per_rank) + n), s);
  dw[(size_t)k * ldw + n] = s;
}

// out[n] = sum_b dz[b][n] (bias gradient of a dense layer), rows of one rank
__global__ void __launch_bounds__(256) dense_bias_grad_kernel(const float* __restrict__ dz, int B, int N, int ld,
                                                              float* __restrict__ db) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += __ldg(dz + (size_t)b * ld + n);
  db[n] = s;
}

// ---------------------------------------------------------------------------------------------
// Global norm and Adam over the materialised part of the gradient, given as `world` per-rank
// copies that are summed on the fly in rank order (an all-gathered buffer used in place; world = 1
// is the plain case), plus `extra` (device double): sum of squares of the factored gradients.
// ---------------------------------------------------------------------------------------------
struct RedWork2 {
  unsigned int counter;
  unsigned int pad[3];
  double partial[1];
};

static int red2_blocks(size_t n) {
  size_t b = (n + 4095) / 4096;
  if (b > 8 * 148) b = 8 * 148;
  if (b < 1) b = 1;
  return (int)b;
}

__global__ void __launch_bounds__(256) sumsq_multi_kernel(const float* __restrict__ g, size_t rank_stride, int world,
                                                          size_t n, float grad_scale, const double* __restrict__ denom,
                                                          const double* __restrict__ extra, float clip_norm,
                                                          const float* __restrict__ lr, float beta1, float beta2,
                                                          float* __restrict__ state, RedWork2* __restrict__ work) {
  pdl_trigger();
  pdl_wait();
  if (denom) grad_scale /= (float)(*denom);
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int r = 0; r < world; ++r) a += __ldg(g + (size_t)r * rank_stride + i);
    s = fmaf(a, a, s);
  }
  double d = warp_sum((double)s);
  __shared__ double red[8];
  __shared__ bool is_last;
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    work->partial[blockIdx.x] = t;
    __threadfence();
    is_last = atomicAdd(&work->counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double mine = 0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) mine += *((volatile double*)&work->partial[b]);
  mine = warp_sum(mine);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    if (extra) t += *extra;
    const float norm = (float)(sqrt(t) * (double)grad_scale);
    float scale = 1.0f;
    if (clip_norm > 0.f) scale = clip_norm * fminf(1.0f / norm, 1.0f / clip_norm);
    const float b1p = state[0], b2p = state[1];
    state[2] = norm;
    state[3] = (*lr) * sqrtf(1.0f - b2p) / (1.0f - b1p);
    state[4] = scale;
    state[0] = b1p * beta1;
    state[1] = b2p * beta2;
    work->counter = 0;
  }
}

__global__ void __launch_bounds__(256) adam_multi_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                         size_t rank_stride, int world, float* __restrict__ m,
                                                         float* __restrict__ v, size_t n, float grad_scale,
                                                         const double* __restrict__ denom, float beta1, float beta2,
                                                         float eps, const float* __restrict__ state) {
  pdl_trigger();
  pdl_wait();
  if (denom) grad_scale /= (float)(*denom);
  const float gs = grad_scale * state[4];
  const float lr_t = state[3];
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int r = 0; r < world; ++r) a += __ldg(g + (size_t)r * rank_stride + i);
    const float gk = a * gs;
    const float mk = m[i] + (gk - m[i]) * omb1;
    const float vk = v[i] + (gk * gk - v[i]) * omb2;
    m[i] = mk;
    v[i] = vk;
    p[i] = p[i] - (mk * lr_t) / (sqrtf(vk) + eps);
  }
}

// out[k] = sum over ranks of in[r * rank_stride + k] (doubles), k < count: loss sums of all ranks
__global__ void sum_ranks_f64_kernel(const double* __restrict__ in, size_t rank_stride, int world, int count,
                                     double* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int k = threadIdx.x;
  if (k >= count) return;
  double s = 0.0;
  for (int r = 0; r < world; ++r) s += in[(size_t)r * rank_stride + k];
  out[k] = s;
}

// doubles <-> (hi, lo) float pairs, so that a float32 all-reduce can carry the loss sums exactly
// enough (the pair sums recombine to ~1e-14 relative): out[2k] = float(x), out[2k+1] = float(x - hi)
__global__ void f64_to_hilo_kernel(const double* __restrict__ in, int count, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int k = threadIdx.x;
  if (k >= count) return;
  const double x = in[k];
  const float hi = (float)x;
  out[2 * k] = hi;
  out[2 * k + 1] = (float)(x - (double)hi);
}
__global__ void hilo_to_f64_kernel(const float* __restrict__ in, int count, double* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int k = threadIdx.x;
  if (k >= count) return;
  out[k] = (double)in[2 * k] + (double)in[2 * k + 1];
}

}  // namespace dincae

using namespace dincae;

static size_t gram_work_bytes(int Bt) { return align_up(16 + (size_t)((Bt * Bt + 255) / 256) * sizeof(double), 256); }

static int gram_chunks(int ncols, int Bt, int* cols_per_cta) {
  // about two CTAs per SM over (column chunks x pair tiles); chunks are multiples of GR_COLS columns
  const int nt = (Bt + GR_TILE - 1) / GR_TILE;
  int want = (2 * 148) / (nt * nt);
  if (want < 1) want = 1;
  int per = (ncols + want - 1) / want;
  per = (per + GR_COLS - 1) / GR_COLS * GR_COLS;
  if (per < GR_COLS) per = GR_COLS;
  *cols_per_cta = per;
  return (ncols + per - 1) / per;
}

extern "C" size_t dincae_dense_factored_workspace(int Bt, int K, int N) {
  if (Bt <= 0 || K <= 0 || N <= 0) return 0;
  int per;
  const int wide = K > N ? K : N;
  const int chunks = gram_chunks(wide, Bt, &per);
  return align_up((size_t)2 * chunks * Bt * Bt * sizeof(float), 256) + gram_work_bytes(Bt);
}

extern "C" int dincae_dense_factored_sumsq(const float* x, size_t x_rank_stride, int ldx,
                                           const float* dz, size_t dz_rank_stride, int lddz,
                                           int world, int rows_per_rank, int K, int N, double* sumsq,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  const int Bt = world * rows_per_rank;
  if (!x || !dz || !sumsq || !workspace || world <= 0 || rows_per_rank <= 0 || K <= 0 || N <= 0) return DINCAE_EINVAL;
  if (Bt > GR_MAX_BT) return DINCAE_EINVAL;
  if (workspace_bytes < dincae_dense_factored_workspace(Bt, K, N)) return DINCAE_ENOSPC;
  int per;
  const int wide = K > N ? K : N;
  const int chunks = gram_chunks(wide, Bt, &per);
  const int chunks_x = (K + per - 1) / per, chunks_z = (N + per - 1) / per;
  const int nt = (Bt + GR_TILE - 1) / GR_TILE;
  Factor fx = {x, x_rank_stride, ldx}, fz = {dz, dz_rank_stride, lddz};
  cudaStream_t st = (cudaStream_t)stream;
  launch_k(gram_partial_kernel, dim3(chunks, 2, nt * nt), 256, 0, st, fx, fz, Bt, rows_per_rank, K, N, per,
           reinterpret_cast<float*>(workspace));
  int rc = check_launch();
  if (rc) return rc;
  // the finish kernel's counter + block sums sit behind the partials; the counter is zero on entry
  // (zeroed here once per call: cheap, and robust against an aborted earlier launch)
  unsigned char* wk = reinterpret_cast<unsigned char*>(workspace) +
                      align_up((size_t)2 * chunks * Bt * Bt * sizeof(float), 256);
  rc = cuda_rc(cudaMemsetAsync(wk, 0, 16, st));
  if (rc) return rc;
  launch_k(gram_finish_kernel, (Bt * Bt + 255) / 256, 256, 0, st, reinterpret_cast<const float*>(workspace), chunks_x,
           chunks_z, chunks, Bt * Bt, sumsq, reinterpret_cast<GramWork*>(wk));
  return check_launch();
}

extern "C" int dincae_adam_update_factored(float* w, float* m, float* v, int ldw,
                                           const float* x, size_t x_rank_stride, int ldx,
                                           const float* dz, size_t dz_rank_stride, int lddz,
                                           int world, int rows_per_rank, int K, int N,
                                           float grad_scale, const double* denom, float beta1, float beta2,
                                           float eps, const float* state, void* w8, void* stream) {
  if (!w || !m || !v || !x || !dz || !state || world <= 0 || rows_per_rank <= 0 || K <= 0 || N <= 0) return DINCAE_EINVAL;
  if (w8 && ((K & 7) || (N & 7))) return DINCAE_EINVAL;
  if ((ldw & 3) || (N & 3) || (K & 3) || (ldx & 3) || (lddz & 3) || (x_rank_stride & 3) || (dz_rank_stride & 3) ||
      (((uintptr_t)w | (uintptr_t)m | (uintptr_t)v | (uintptr_t)x | (uintptr_t)dz) & 15))
    return DINCAE_EINVAL;
  Factor fx = {x, x_rank_stride, ldx}, fz = {dz, dz_rank_stride, lddz};
  // Bt > 16 frames: the rank-Bt product goes to the tensor cores (TF32 operands; exact for the bf16
  // network's factors).  DINCAE_ADAM_MMA=0 / 1 forces the FFMA / the tensor-core kernel.
  static int force = -2;
  if (force == -2) {
    const char* e = getenv("DINCAE_ADAM_MMA");
    force = e ? atoi(e) : -1;
  }
  const int Bt = world * rows_per_rank;
  const bool use_mma = force == 1 || (force != 0 && Bt > 16);
  // measured per configs[3] kernel (B200): Bt 32: 2.8 ms per-tile staging vs 3.0-3.8 persistent;
  // Bt 64: 4.05 vs 3.3-4.0; Bt 128: 6.3 vs 3.9-4.7
  const char* e_tma = getenv("DINCAE_ADAM_TMA");       // read per call: the tests switch between the kernels
  const bool use_tma = e_tma ? atoi(e_tma) != 0 : true;
  if (use_mma && use_tma && Bt <= 128) {
    static int sm_count_x = 0;
    if (!sm_count_x) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sm_count_x, cudaDevAttrMultiProcessorCount, dev);
      if (sm_count_x <= 0) sm_count_x = 148;
    }
    CUtensorMap tmw, tmm, tmv;
    int rc = ax_encode_map(&tmw, w, K, N, ldw);
    if (!rc) rc = ax_encode_map(&tmm, m, K, N, ldw);
    if (!rc) rc = ax_encode_map(&tmv, v, K, N, ldw);
    if (rc) return rc;
    const cudaStream_t st = (cudaStream_t)stream;
    if (Bt <= 32)
      rc = ax_launch<4>(sm_count_x, tmw, tmm, tmv, fx, fz, Bt, rows_per_rank, K, N, grad_scale, denom, beta1, beta2, eps, state, w8, st);
    else if (Bt <= 64)
      rc = ax_launch<8>(sm_count_x, tmw, tmm, tmv, fx, fz, Bt, rows_per_rank, K, N, grad_scale, denom, beta1, beta2, eps, state, w8, st);
    else
      rc = ax_launch<16>(sm_count_x, tmw, tmm, tmv, fx, fz, Bt, rows_per_rank, K, N, grad_scale, denom, beta1, beta2, eps, state, w8, st);
    if (rc) return rc;
  } else if (use_mma && Bt > 48 && Bt <= AP_MAX_BT) {
    static int sm_count = 0;
    if (!sm_count) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
      if (sm_count <= 0) sm_count = 148;
    }
    const int Bp = (Bt + 7) & ~7;
    const size_t smem = ((size_t)Bp * AM_ZP + 2 * (size_t)Bp * AM_XP) * sizeof(float) + 8192;
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(adam_factored_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(((size_t)AP_MAX_BT * AM_ZP + 2 * (size_t)AP_MAX_BT * AM_XP) * sizeof(float) + 8192));
      if (e != cudaSuccess) return cuda_rc(e);
      attr_set = true;
    }
    const int col_blocks = ceil_div(N, 256), n_tiles = ceil_div(K, 16);
    int row_splits = ceil_div(2 * sm_count, col_blocks);          // about two waves of CTAs
    if (row_splits > n_tiles) row_splits = n_tiles;
    const int tiles_per_cta = ceil_div(n_tiles, row_splits);
    row_splits = ceil_div(n_tiles, tiles_per_cta);
    launch_k(adam_factored_persist_kernel, dim3(col_blocks, row_splits), 256, smem, (cudaStream_t)stream, w, m, v, ldw,
             fx, fz, Bt, rows_per_rank, K, N, tiles_per_cta, grad_scale, denom, beta1, beta2, eps, state,
             reinterpret_cast<uint32_t*>(w8));
  } else if (use_mma)
    launch_k(adam_factored_mma_kernel, dim3(ceil_div(N, 256), ceil_div(K, 16)), 256, 0, (cudaStream_t)stream, w, m, v,
             ldw, fx, fz, Bt, rows_per_rank, K, N, grad_scale, denom, beta1, beta2, eps, state,
             reinterpret_cast<uint32_t*>(w8));
  else {
    // Bt <= 16: the TMA pipeline (DINCAE_ADAM_TMA=0 selects the register-staged kernel)
    if (use_tma && Bt <= 16) {
      static int sm_count = 0;
      static bool attr_set = false;
      if (!attr_set) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (sm_count <= 0) sm_count = 148;
        cudaError_t e = cudaFuncSetAttribute(adam_factored_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
        if (e != cudaSuccess) return cuda_rc(e);
        attr_set = true;
      }
      CUtensorMap tmw, tmm, tmv;
      int rc = at_encode_map(&tmw, w, K, N, ldw);
      if (!rc) rc = at_encode_map(&tmm, m, K, N, ldw);
      if (!rc) rc = at_encode_map(&tmv, v, K, N, ldw);
      if (rc) return rc;
      launch_k(adam_factored_tma_kernel, sm_count, AT_THREADS, AT_SMEM, (cudaStream_t)stream, tmw, tmm, tmv, fx, fz, Bt,
               rows_per_rank, K, N, grad_scale, denom, beta1, beta2, eps, state, reinterpret_cast<uint4*>(w8));
    } else
      launch_k(adam_factored_kernel, dim3(ceil_div(N, 256), ceil_div(K, AF_TK)), 256, 0, (cudaStream_t)stream, w, m, v,
               ldw, fx, fz, Bt, rows_per_rank, K, N, grad_scale, denom, beta1, beta2, eps, state,
               reinterpret_cast<uint2*>(w8));
  }
  return check_launch();
}

extern "C" int dincae_dense_factored_expand(const float* x, size_t x_rank_stride, int ldx,
                                            const float* dz, size_t dz_rank_stride, int lddz,
                                            int world, int rows_per_rank, int K, int N, float* dw, int ldw,
                                            void* stream) {
  if (!x || !dz || !dw || world <= 0 || rows_per_rank <= 0 || K <= 0 || N <= 0 || K > 65535) return DINCAE_EINVAL;
  Factor fx = {x, x_rank_stride, ldx}, fz = {dz, dz_rank_stride, lddz};
  launch_k(factored_expand_kernel, dim3(ceil_div(N, 256), K), 256, 0, (cudaStream_t)stream, dw, ldw, fx, fz,
           world * rows_per_rank, rows_per_rank, K, N);
  return check_launch();
}

extern "C" int dincae_dense_bias_grad(const float* dz, int B, int N, int ld, float* db, void* stream) {
  if (!dz || !db || B <= 0 || N <= 0) return DINCAE_EINVAL;
  launch_k(dense_bias_grad_kernel, ceil_div(N, 256), 256, 0, (cudaStream_t)stream, dz, B, N, ld, db);
  return check_launch();
}

extern "C" size_t dincae_adam_multi_workspace(size_t n) {
  return align_up(16 + (size_t)red2_blocks(n) * sizeof(double), 256);
}

extern "C" int dincae_adam_step_multi(float* params, const float* grads, size_t grad_rank_stride, int world,
                                      float* m, float* v, size_t n, float clip_norm, const float* lr,
                                      float beta1, float beta2, float eps, float grad_scale, const double* denom,
                                      const double* extra_sumsq, float* state,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  if (!params || !grads || !m || !v || !lr || !state || !workspace || n == 0 || world <= 0) return DINCAE_EINVAL;
  if (workspace_bytes < dincae_adam_multi_workspace(n)) return DINCAE_ENOSPC;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = cuda_rc(cudaMemsetAsync(workspace, 0, 16, st));
  if (rc) return rc;
  launch_k(sumsq_multi_kernel, red2_blocks(n), 256, 0, st, grads, grad_rank_stride, world, n, grad_scale, denom,
           extra_sumsq, clip_norm, lr, beta1, beta2, state, reinterpret_cast<RedWork2*>(workspace));
  rc = check_launch();
  if (rc) return rc;
  size_t ub = (n + 255) / 256;
  if (ub > 8 * 148) ub = 8 * 148;
  launch_k(adam_multi_kernel, (unsigned)ub, 256, 0, st, params, grads, grad_rank_stride, world, m, v, n, grad_scale,
           denom, beta1, beta2, eps, state);
  return check_launch();
}

extern "C" int dincae_sum_ranks_f64(const double* in, size_t rank_stride, int world, int count, double* out,
                                    void* stream) {
  if (!in || !out || world <= 0 || count <= 0 || count > 32) return DINCAE_EINVAL;
  launch_k(sum_ranks_f64_kernel, 1, 32, 0, (cudaStream_t)stream, in, rank_stride, world, count, out);
  return check_launch();
}

extern "C" int dincae_f64_to_hilo(const double* in, int count, float* out, void* stream) {
  if (!in || !out || count <= 0 || count > 32) return DINCAE_EINVAL;
  launch_k(f64_to_hilo_kernel, 1, 32, 0, (cudaStream_t)stream, in, count, out);
  return check_launch();
}

extern "C" int dincae_hilo_to_f64(const float* in, int count, double* out, void* stream) {
  if (!in || !out || count <= 0 || count > 32) return DINCAE_EINVAL;
  launch_k(hilo_to_f64_kernel, 1, 32, 0, (cudaStream_t)stream, in, count, out);
  return check_launch();
}
