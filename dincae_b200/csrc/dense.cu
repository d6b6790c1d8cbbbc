// Dense bottleneck (DINCAE/__init__.py:465-485): three skinny fp32 GEMMs per step
// (forward, data gradient, weight gradient).  The batch is the small dimension, so all
// three are bound by streaming the weight matrix once; split-K spreads that stream over
// the SMs and a second tiny kernel reduces the splits in fixed order and applies
// bias / ReLU / ReLU-gradient.
#include "common.cuh"
#include "umma.cuh"
#include <stdlib.h>

namespace dincae {

constexpr int GT = 64;    // tile M = tile N
constexpr int GK = 16;    // tile K

// C[M,N] (partial over a K range) = A x Bm, with
//   TA == 0: A[m][k] = a[m*lda + k]      TA == 1: A[m][k] = a[k*lda + m]
//   TB == 0: Bm[k][n] = b[k*ldb + n]     TB == 1: Bm[k][n] = b[n*ldb + k]
// 256 threads, 4x4 outputs each.  blockIdx.z selects the K split; out is [splits][M][N].
template <int TA, int TB>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ a, int lda,
                                                    const float* __restrict__ b, int ldb,
                                                    int M, int N, int K, int k_per_split,
                                                    float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float As[GK][GT + 4];
  __shared__ __align__(16) float Bs[GK][GT + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += GK) {
    // A tile: GT x GK
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + 256 * r;
      int m, k;
      if (TA == 0) { k = idx & (GK - 1); m = idx >> 4; }     // k fastest (contiguous in a)
      else         { m = idx & (GT - 1); k = idx >> 6; }     // m fastest
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < kend) v = TA == 0 ? __ldg(&a[(size_t)gm * lda + gk]) : __ldg(&a[(size_t)gk * lda + gm]);
      As[k][m] = v;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + 256 * r;
      int n, k;
      if (TB == 0) { n = idx & (GT - 1); k = idx >> 6; }     // n fastest (contiguous in b)
      else         { k = idx & (GK - 1); n = idx >> 4; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < kend) v = TB == 0 ? __ldg(&b[(size_t)gk * ldb + gn]) : __ldg(&b[(size_t)gn * ldb + gk]);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][tm]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tn]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* o = out + (size_t)blockIdx.z * M * N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + tm + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tn + j;
      if (gn < N) o[(size_t)gm * N + gn] = acc[i][j];
    }
  }
}

// Vectorised variant used when every row pitch and pointer is 16-byte aligned (always the case
// for the padded dense layers): float4 global loads along the contiguous dimension, the next
// K tile prefetched into registers while the current one is multiplied, and optional fused
// work so that the skinny GEMMs of the bottleneck need no follow-up kernels:
//   EPI 0: raw partial of a K split -> out[split][M][N]
//   EPI 1: y = epi(acc + bias) with ReLU / gate (single split)
//   colsum != null and blockIdx.y == 0: colsum[n] = sum over k of Bm[k][n]  (bias gradient
//   of the weight-gradient GEMM, whose B operand is dz)
constexpr int VK = 32;
template <int TA, int TB, int EPI>
__global__ void __launch_bounds__(256) sgemm_v4_kernel(const float* __restrict__ a, int lda,
                                                       const float* __restrict__ b, int ldb,
                                                       int M, int N, int K, int k_per_split,
                                                       float* __restrict__ out,
                                                       const float* __restrict__ bias, int relu,
                                                       const float* __restrict__ gate,
                                                       float* __restrict__ colsum) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float As[VK][GT + 4];
  __shared__ __align__(16) float Bs[VK][GT + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float csum = 0.f;                         // column tid (< GT) of the B tile, summed over k
  const bool want_cs = colsum != nullptr && blockIdx.y == 0;

  float4 ra[2], rb[2];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int f = tid + 256 * r;
      if (TA == 0) {                          // a[m][k], float4 along k
        const int m = f >> 3, k = k0 + 4 * (f & 7);
        ra[r] = (m0 + m < M && k < kend) ? __ldg(reinterpret_cast<const float4*>(a + (size_t)(m0 + m) * lda + k))
                                         : f4_zero();
      } else {                                // a[k][m], float4 along m
        const int k = k0 + (f >> 4), m = m0 + 4 * (f & 15);
        ra[r] = (m < M && k < kend) ? __ldg(reinterpret_cast<const float4*>(a + (size_t)k * lda + m))
                                    : f4_zero();
      }
      if (TB == 0) {                          // b[k][n], float4 along n
        const int k = k0 + (f >> 4), n = n0 + 4 * (f & 15);
        rb[r] = (n < N && k < kend) ? __ldg(reinterpret_cast<const float4*>(b + (size_t)k * ldb + n))
                                    : f4_zero();
      } else {                                // b[n][k], float4 along k
        const int n = f >> 3, k = k0 + 4 * (f & 7);
        rb[r] = (n0 + n < N && k < kend) ? __ldg(reinterpret_cast<const float4*>(b + (size_t)(n0 + n) * ldb + k))
                                         : f4_zero();
      }
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int f = tid + 256 * r;
      if (TA == 0) {
        const int m = f >> 3, k = 4 * (f & 7);
        As[k][m] = ra[r].x; As[k + 1][m] = ra[r].y; As[k + 2][m] = ra[r].z; As[k + 3][m] = ra[r].w;
      } else {
        *reinterpret_cast<float4*>(&As[f >> 4][4 * (f & 15)]) = ra[r];
      }
      if (TB == 0) {
        *reinterpret_cast<float4*>(&Bs[f >> 4][4 * (f & 15)]) = rb[r];
      } else {
        const int n = f >> 3, k = 4 * (f & 7);
        Bs[k][n] = rb[r].x; Bs[k + 1][n] = rb[r].y; Bs[k + 2][n] = rb[r].z; Bs[k + 3][n] = rb[r].w;
      }
    }
  };

  fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += VK) {
    stash();
    __syncthreads();
    if (k0 + VK < kend) fetch(k0 + VK);
#pragma unroll
    for (int k = 0; k < VK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][tm]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tn]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    if (want_cs && tid < GT) {
#pragma unroll
      for (int k = 0; k < VK; ++k) csum += Bs[k][tid];
    }
    __syncthreads();
  }
  if (want_cs && tid < GT && n0 + tid < N) colsum[n0 + tid] = csum;
  float* o = EPI == 0 ? out + (size_t)blockIdx.z * M * N : out;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + tm + i;
    if (gm >= M) continue;
    const int gn = n0 + tn;
    if (gn >= N) continue;                       // N is a multiple of 4: the float4 is in or out
    float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (EPI == 1) {
      if (bias) v = f4_add(v, __ldg(reinterpret_cast<const float4*>(bias + gn)));
      if (relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
      if (gate) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gate + (size_t)gm * N + gn));
        v = make_float4(g.x > 0.f ? v.x : 0.f, g.y > 0.f ? v.y : 0.f, g.z > 0.f ? v.z : 0.f, g.w > 0.f ? v.w : 0.f);
      }
    }
    *reinterpret_cast<float4*>(o + (size_t)gm * N + gn) = v;
  }
}


// ---------------------------------------------------------------------------------------
// Small-batch ("skinny") variants, M = batch <= 16.  With a handful of rows the three GEMMs
// are pure streams over the weight matrix (config D: 1.38 GB per dense layer), so the 64-row
// tiles above waste 4-8x their FMAs and shared-memory bandwidth on padding.  Here every
// weight element is fetched once, straight from global memory into registers (8-16
// independent 16-byte loads in flight per thread), and meets all M rows of the small operand,
// which lives in shared memory (broadcast reads) -- HBM-bound by construction.
// ---------------------------------------------------------------------------------------
constexpr int SK_KC = 64;      // k per staged chunk of the small operand (skinny_nn)
constexpr int SK_NC = 512;     // n per staged chunk of the small operand (skinny_nt)

// out[m][n] = sum_k A[m][k] * W[k][n]   (A [M][lda], W [K][N], N % 4 == 0, M <= RM)
// grid (ceil(N/128), 1, splits); 8 warps = 8 interleaved k groups, lane = 4 columns.
// EPI 0: partial -> out[split][M][N];  EPI 1: y = relu?(acc + bias) (single split).
template <int RM, int EPI>
__global__ void __launch_bounds__(256) skinny_nn_kernel(const float* __restrict__ a, int lda,
                                                        const float* __restrict__ w, int M, int N, int K,
                                                        int k_per_split, float* __restrict__ out,
                                                        const float* __restrict__ bias, int relu) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float As[2][SK_KC][RM];
  __shared__ __align__(16) float red[RM][128];
  const int tid = threadIdx.x, lane = tid & 31, kg = tid >> 5;
  const int n = blockIdx.x * 128 + 4 * lane;
  const bool n_ok = n < N;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  float4 acc[RM];
#pragma unroll
  for (int m = 0; m < RM; ++m) acc[m] = f4_zero();

  auto stage = [&](int buf, int k0) {
#pragma unroll
    for (int r = 0; r < SK_KC * RM / 256; ++r) {
      const int idx = tid + 256 * r;
      const int kk = idx % SK_KC, m = idx / SK_KC;
      const int k = k0 + kk;
      As[buf][kk][m] = (m < M && k < kend) ? __ldg(a + (size_t)m * lda + k) : 0.f;
    }
  };
  int buf = 0;
  if (kbeg < kend) stage(0, kbeg);
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += SK_KC) {
    float4 wv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + kg + 8 * j;
      wv[j] = ldg_f4_pred(reinterpret_cast<const float4*>(w + (size_t)k * N + n), n_ok && k < kend);
    }
    if (k0 + SK_KC < kend) stage(buf ^ 1, k0 + SK_KC);      // other buffer: no hazard with the reads below
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* ar = As[buf][kg + 8 * j];
#pragma unroll
      for (int m4 = 0; m4 < RM / 4; ++m4) {
        const float4 av = *reinterpret_cast<const float4*>(ar + 4 * m4);
        acc[4 * m4 + 0] = f4_fma(av.x, wv[j], acc[4 * m4 + 0]);
        acc[4 * m4 + 1] = f4_fma(av.y, wv[j], acc[4 * m4 + 1]);
        acc[4 * m4 + 2] = f4_fma(av.z, wv[j], acc[4 * m4 + 2]);
        acc[4 * m4 + 3] = f4_fma(av.w, wv[j], acc[4 * m4 + 3]);
      }
    }
    __syncthreads();
    buf ^= 1;
  }
  // ordered (deterministic) sum over the 8 k groups
  for (int g = 0; g < 8; ++g) {
    if (kg == g) {
#pragma unroll
      for (int m = 0; m < RM; ++m) {
        float4* r = reinterpret_cast<float4*>(&red[m][4 * lane]);
        *r = g == 0 ? acc[m] : f4_add(*r, acc[m]);
      }
    }
    __syncthreads();
  }
  float* o = EPI == 0 ? out + (size_t)blockIdx.z * M * N : out;
  for (int m = kg; m < M; m += 8) {
    if (!n_ok) continue;
    float4 v = *reinterpret_cast<const float4*>(&red[m][4 * lane]);
    if (EPI == 1) {
      if (bias) v = f4_add(v, __ldg(reinterpret_cast<const float4*>(bias + n)));
      if (relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    }
    *reinterpret_cast<float4*>(o + (size_t)m * N + n) = v;
  }
}

// out[m][k] = sum_n D[m][n] * W[k][n]   (D [M][N], W [K][N], N % 4 == 0, M <= RM)
// grid (ceil(K/32), 1, splits over n); warp = 4 rows of W, lanes stride along n; the rows of D
// are staged chunk-wise in shared memory and reused by the 4 rows of every warp.
// EPI 0: partial -> out[split][M][K];  EPI 1: dx = acc * (gate > 0) (single split).
template <int RM, int EPI>
__global__ void __launch_bounds__(256) skinny_nt_kernel(const float* __restrict__ d,
                                                        const float* __restrict__ w, int M, int N, int K,
                                                        int n_per_split, float* __restrict__ out,
                                                        const float* __restrict__ gate) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float Ds[RM][SK_NC];
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  const int krow = blockIdx.x * 32 + 4 * wp;
  const int nbeg = blockIdx.z * n_per_split;
  const int nend = min(N, nbeg + n_per_split);
  float acc[4][RM];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int m = 0; m < RM; ++m) acc[r][m] = 0.f;
  const float* wr[4];
  bool row_ok[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    row_ok[r] = krow + r < K;
    wr[r] = w + (size_t)(row_ok[r] ? krow + r : 0) * N;
  }
  for (int n0 = nbeg; n0 < nend; n0 += SK_NC) {
    // stage D[:, n0 .. n0 + SK_NC)
#pragma unroll
    for (int r = 0; r < RM * SK_NC / 4 / 256; ++r) {
      const int idx = tid + 256 * r;
      const int m = idx / (SK_NC / 4), c = idx % (SK_NC / 4);
      const int nn = n0 + 4 * c;
      *reinterpret_cast<float4*>(&Ds[m][4 * c]) =
          ldg_f4_pred(reinterpret_cast<const float4*>(d + (size_t)m * N + nn), m < M && nn < nend);
    }
    __syncthreads();
    float4 wv[4][SK_NC / 128];
#pragma unroll
    for (int i = 0; i < SK_NC / 128; ++i) {
      const int nn = n0 + 128 * i + 4 * lane;
#pragma unroll
      for (int r = 0; r < 4; ++r)
        wv[r][i] = ldg_f4_pred(reinterpret_cast<const float4*>(wr[r] + nn), row_ok[r] && nn < nend);
    }
#pragma unroll
    for (int i = 0; i < SK_NC / 128; ++i) {
#pragma unroll
      for (int m = 0; m < RM; ++m) {
        const float4 dv = *reinterpret_cast<const float4*>(&Ds[m][128 * i + 4 * lane]);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          acc[r][m] = fmaf(dv.x, wv[r][i].x, acc[r][m]);
          acc[r][m] = fmaf(dv.y, wv[r][i].y, acc[r][m]);
          acc[r][m] = fmaf(dv.z, wv[r][i].z, acc[r][m]);
          acc[r][m] = fmaf(dv.w, wv[r][i].w, acc[r][m]);
        }
      }
    }
    __syncthreads();
  }
  // butterfly sum over the lanes (fixed order: deterministic), result in every lane
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int m = 0; m < RM; ++m) {
      float v = acc[r][m];
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
      acc[r][m] = v;
    }
  float* o = EPI == 0 ? out + (size_t)blockIdx.z * M * K : out;
  // lane l writes (r, m) = (l / RM..): spread the 4 * RM results of the warp over its lanes
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int m = 0; m < RM; ++m) {
      if (lane == ((r * RM + m) & 31) && m < M && row_ok[r]) {
        float v = acc[r][m];
        const size_t at = (size_t)m * K + krow + r;
        if (EPI == 1 && gate) v = __ldg(gate + at) > 0.f ? v : 0.f;
        o[at] = v;
      }
    }
}

// dw[k][n] = sum_b x[b][k] * dz[b][n], db[n] = sum_b dz[b][n]   (any B; K, N % 4 == 0)
// grid (ceil(N/256), ceil(K/32)); thread = 8 rows x 4 columns; exactly B FMAs per output, so the
// kernel is bound by writing dw once.
__global__ void __launch_bounds__(256) outer_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                    int B, int K, int N, float* __restrict__ dw,
                                                    float* __restrict__ db) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float xs[16][32];
  __shared__ __align__(16) float zs[16][256];
  const int tid = threadIdx.x, tn = tid & 63, tk = tid >> 6;
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 256;
  const int n = n0 + 4 * tn;
  float4 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = f4_zero();
  float4 zsum = f4_zero();
  for (int b0 = 0; b0 < B; b0 += 16) {
    if (tid < 128) {                                    // x chunk: 16 x 32 floats as float4
      const int bb = tid >> 3, c = tid & 7;
      const int k = k0 + 4 * c;
      *reinterpret_cast<float4*>(&xs[bb][4 * c]) =
          ldg_f4_pred(reinterpret_cast<const float4*>(x + (size_t)(b0 + bb) * K + k), b0 + bb < B && k < K);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {                       // dz chunk: 16 x 256 floats as float4
      const int idx = tid + 256 * r;
      const int bb = idx >> 6, c = idx & 63;
      const int nn = n0 + 4 * c;
      *reinterpret_cast<float4*>(&zs[bb][4 * c]) =
          ldg_f4_pred(reinterpret_cast<const float4*>(dz + (size_t)(b0 + bb) * N + nn), b0 + bb < B && nn < N);
    }
    __syncthreads();
#pragma unroll
    for (int bb = 0; bb < 16; ++bb) {
      const float4 zv = *reinterpret_cast<const float4*>(&zs[bb][4 * tn]);
      const float4 x0 = *reinterpret_cast<const float4*>(&xs[bb][8 * tk]);
      const float4 x1 = *reinterpret_cast<const float4*>(&xs[bb][8 * tk + 4]);
      acc[0] = f4_fma(x0.x, zv, acc[0]); acc[1] = f4_fma(x0.y, zv, acc[1]);
      acc[2] = f4_fma(x0.z, zv, acc[2]); acc[3] = f4_fma(x0.w, zv, acc[3]);
      acc[4] = f4_fma(x1.x, zv, acc[4]); acc[5] = f4_fma(x1.y, zv, acc[5]);
      acc[6] = f4_fma(x1.z, zv, acc[6]); acc[7] = f4_fma(x1.w, zv, acc[7]);
      zsum = f4_add(zsum, zv);
    }
    __syncthreads();
  }
  if (n < N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = k0 + 8 * tk + i;
      if (k < K) *reinterpret_cast<float4*>(dw + (size_t)k * N + n) = acc[i];
    }
    if (db && blockIdx.y == 0 && tk == 0) *reinterpret_cast<float4*>(db + n) = zsum;
  }
}

constexpr int SKINNY_MAX_M = 16;

// splits so that about two waves of CTAs stream the matrix; ranges are multiples of `quantum`
static int skinny_splits(int strips, int depth, int quantum) {
  int splits = (2 * 148 + strips - 1) / strips;
  const int max_splits = ceil_div(depth, 4 * quantum);
  if (splits > max_splits) splits = max_splits;
  return splits < 1 ? 1 : splits;
}
static int skinny_nn_splits(int N, int K) { return skinny_splits(ceil_div(N, 128), K, SK_KC); }
static int skinny_nt_splits(int K, int N) { return skinny_splits(ceil_div(K, 32), N, SK_NC); }

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// y[m][n] = epi(sum_s part[s][m][n] + bias[n]);  epi: relu and/or multiply by (gate > 0).
__global__ void splitk_epilogue_kernel(const float* __restrict__ part, int splits, int M, int N,
                                       const float* __restrict__ bias, int relu,
                                       const float* __restrict__ gate, float* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)M * N;
  if (e >= total) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += part[(size_t)k * total + e];
  if (bias) s += bias[e % N];
  if (relu) s = fmaxf(s, 0.f);
  if (gate) s = gate[e] > 0.f ? s : 0.f;
  y[e] = s;
}

// db[n] = sum_b dz[b][n]
__global__ void colsum_kernel(const float* __restrict__ dz, int B, int N, float* __restrict__ db) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += dz[(size_t)b * N + n];
  db[n] = s;
}

static int pick_splits(int M, int N, int K) {
  const int tiles = ceil_div(M, GT) * ceil_div(N, GT);
  if (tiles >= 120) return 1;                // enough CTAs: keep the epilogue fused
  int splits = (148 + tiles - 1) / tiles;
  const int max_splits = ceil_div(K, 4 * GK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  return splits;
}

static size_t splitk_bytes(int M, int N, int K) {
  return align_up((size_t)pick_splits(M, N, K) * M * N * sizeof(float), 256);
}

static int g_outer_max_b = -1;        // largest batch that takes outer_kernel (DINCAE_OUTER_MAX_B)
static int outer_max_b() {
  if (g_outer_max_b < 0) {
    const char* e = getenv("DINCAE_OUTER_MAX_B");
    g_outer_max_b = e ? atoi(e) : SKINNY_MAX_M;
  }
  return g_outer_max_b;
}

}  // namespace dincae

using namespace dincae;

extern "C" size_t dincae_dense_workspace(int B, int K, int N) {
  size_t a = splitk_bytes(B, N, K), b = splitk_bytes(B, K, N);
  if (B <= SKINNY_MAX_M) {
    const size_t sa = align_up((size_t)skinny_nn_splits(N, K) * B * N * sizeof(float), 256);
    const size_t sb = align_up((size_t)skinny_nt_splits(K, N) * B * K * sizeof(float), 256);
    a = sa > a ? sa : a;
    b = sb > b ? sb : b;
  }
  return a > b ? a : b;
}

extern "C" int dincae_dense_fwd(const float* x, const float* w, const float* bias, int B, int K,
                                int N, int relu, float* y, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (!x || !w || !y || !workspace || B <= 0 || K <= 0 || N <= 0) return DINCAE_EINVAL;
  const int splits = pick_splits(B, N, K);
  if ((size_t)splits * B * N * sizeof(float) > workspace_bytes) return DINCAE_ENOSPC;
  int kps = ceil_div(ceil_div(K, splits), VK) * VK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ceil_div(N, GT), ceil_div(B, GT), ceil_div(K, kps));
  const bool vec = !(K & 3) && !(N & 3) && aligned16(x) && aligned16(w) && aligned16(y) &&
                   aligned16(workspace) && (!bias || aligned16(bias));
  if (vec && B <= SKINNY_MAX_M) {
    const int sp = skinny_nn_splits(N, K);
    if ((size_t)sp * B * N * sizeof(float) > workspace_bytes) return DINCAE_ENOSPC;
    const int kq = ceil_div(ceil_div(K, sp), SK_KC) * SK_KC;
    dim3 g(ceil_div(N, 128), 1, ceil_div(K, kq));
    if (g.z == 1) {
      if (B <= 8) launch_k(skinny_nn_kernel<8, 1>, g, 256, 0, st, x, K, w, B, N, K, kq, y, bias, relu);
      else launch_k(skinny_nn_kernel<16, 1>, g, 256, 0, st, x, K, w, B, N, K, kq, y, bias, relu);
      return check_launch();
    }
    if (B <= 8) launch_k(skinny_nn_kernel<8, 0>, g, 256, 0, st, x, K, w, B, N, K, kq, (float*)workspace, (const float*)nullptr, 0);
    else launch_k(skinny_nn_kernel<16, 0>, g, 256, 0, st, x, K, w, B, N, K, kq, (float*)workspace, (const float*)nullptr, 0);
    int rc = check_launch();
    if (rc) return rc;
    const size_t total = (size_t)B * N;
    launch_k(splitk_epilogue_kernel, (unsigned)((total + 255) / 256), 256, 0, st,
             (const float*)workspace, (int)g.z, B, N, bias, relu, nullptr, y);
    return check_launch();
  }
  if (vec && grid.z == 1) {
    launch_k(sgemm_v4_kernel<0, 0, 1>, grid, 256, 0, st, x, K, w, N, B, N, K, kps, y, bias, relu,
             (const float*)nullptr, (float*)nullptr);
    return check_launch();
  }
  if (vec)
    launch_k(sgemm_v4_kernel<0, 0, 0>, grid, 256, 0, st, x, K, w, N, B, N, K, kps, (float*)workspace,
             (const float*)nullptr, 0, (const float*)nullptr, (float*)nullptr);
  else
    launch_k(sgemm_kernel<0, 0>, grid, 256, 0, st, x, K, w, N, B, N, K, kps, (float*)workspace);
  int rc = check_launch();
  if (rc) return rc;
  const size_t total = (size_t)B * N;
  launch_k(splitk_epilogue_kernel, (unsigned)((total + 255) / 256), 256, 0, st, 
      (const float*)workspace, (int)grid.z, B, N, bias, relu, nullptr, y);
  return check_launch();
}

extern "C" int dincae_dense_bwd_data(const float* dz, const float* w, const float* act_in, int B,
                                     int K, int N, float* dx, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  if (!dz || !w || !dx || !workspace || B <= 0 || K <= 0 || N <= 0) return DINCAE_EINVAL;
  // dx[B,K] = dz[B,N] * w^T : GEMM with M=B, N'=K, K'=N, Bm[k'][n'] = w[n'][k']
  const int splits = pick_splits(B, K, N);
  if ((size_t)splits * B * K * sizeof(float) > workspace_bytes) return DINCAE_ENOSPC;
  int kps = ceil_div(ceil_div(N, splits), VK) * VK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ceil_div(K, GT), ceil_div(B, GT), ceil_div(N, kps));
  const bool vec = !(K & 3) && !(N & 3) && aligned16(dz) && aligned16(w) && aligned16(dx) &&
                   aligned16(workspace) && (!act_in || aligned16(act_in));
  if (vec && B <= SKINNY_MAX_M) {
    const int sp = skinny_nt_splits(K, N);
    if ((size_t)sp * B * K * sizeof(float) > workspace_bytes) return DINCAE_ENOSPC;
    const int nq = ceil_div(ceil_div(N, sp), SK_NC) * SK_NC;
    dim3 g(ceil_div(K, 32), 1, ceil_div(N, nq));
    if (g.z == 1) {
      if (B <= 8) launch_k(skinny_nt_kernel<8, 1>, g, 256, 0, st, dz, w, B, N, K, nq, dx, act_in);
      else launch_k(skinny_nt_kernel<16, 1>, g, 256, 0, st, dz, w, B, N, K, nq, dx, act_in);
      return check_launch();
    }
    if (B <= 8) launch_k(skinny_nt_kernel<8, 0>, g, 256, 0, st, dz, w, B, N, K, nq, (float*)workspace, (const float*)nullptr);
    else launch_k(skinny_nt_kernel<16, 0>, g, 256, 0, st, dz, w, B, N, K, nq, (float*)workspace, (const float*)nullptr);
    int rc = check_launch();
    if (rc) return rc;
    const size_t total = (size_t)B * K;
    launch_k(splitk_epilogue_kernel, (unsigned)((total + 255) / 256), 256, 0, st,
             (const float*)workspace, (int)g.z, B, K, nullptr, 0, act_in, dx);
    return check_launch();
  }
  if (vec && grid.z == 1) {
    launch_k(sgemm_v4_kernel<0, 1, 1>, grid, 256, 0, st, dz, N, w, N, B, K, N, kps, dx,
             (const float*)nullptr, 0, act_in, (float*)nullptr);
    return check_launch();
  }
  if (vec)
    launch_k(sgemm_v4_kernel<0, 1, 0>, grid, 256, 0, st, dz, N, w, N, B, K, N, kps, (float*)workspace,
             (const float*)nullptr, 0, (const float*)nullptr, (float*)nullptr);
  else
    launch_k(sgemm_kernel<0, 1>, grid, 256, 0, st, dz, N, w, N, B, K, N, kps, (float*)workspace);
  int rc = check_launch();
  if (rc) return rc;
  const size_t total = (size_t)B * K;
  launch_k(splitk_epilogue_kernel, (unsigned)((total + 255) / 256), 256, 0, st, 
      (const float*)workspace, (int)grid.z, B, K, nullptr, 0, act_in, dx);
  return check_launch();
}

extern "C" int dincae_dense_bwd_weight(const float* x, const float* dz, int B, int K, int N,
                                       float* dw, float* db, void* stream) {
  if (!x || !dz || !dw || !db || B <= 0 || K <= 0 || N <= 0) return DINCAE_EINVAL;
  // dw[K,N] = x^T dz : M=K, N'=N, K'=B, A[m][k'] = x[k'][m]
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ceil_div(N, GT), ceil_div(K, GT), 1);
  if (B <= outer_max_b() && !(K & 3) && !(N & 3) && aligned16(x) && aligned16(dz) && aligned16(dw) &&
      aligned16(db)) {
    launch_k(outer_kernel, dim3(ceil_div(N, 256), ceil_div(K, 32)), 256, 0, st, x, dz, B, K, N, dw, db);
    return check_launch();
  }
  if (!(K & 3) && !(N & 3) && aligned16(x) && aligned16(dz) && aligned16(dw)) {
    // weight gradient with the bias gradient (column sums of dz) fused into the first row of CTAs
    launch_k(sgemm_v4_kernel<1, 0, 1>, grid, 256, 0, st, x, K, dz, N, K, N, B, ceil_div(B, VK) * VK, dw,
             (const float*)nullptr, 0, (const float*)nullptr, db);
    return check_launch();
  }
  launch_k(sgemm_kernel<1, 0>, grid, 256, 0, st, x, K, dz, N, K, N, B, ceil_div(B, GK) * GK, dw);
  int rc = check_launch();
  if (rc) return rc;
  launch_k(colsum_kernel, ceil_div(N, 128), 128, 0, st, dz, B, N, db);
  return check_launch();
}
