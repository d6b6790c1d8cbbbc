// Dense bottleneck (DINCAE/__init__.py:465-485): three skinny fp32 GEMMs per step
// (forward, data gradient, weight gradient).  The batch is the small dimension, so all
// three are bound by streaming the weight matrix once; split-K spreads that stream over
// the SMs and a second tiny kernel reduces the splits in fixed order and applies
// bias / ReLU / ReLU-gradient.
#include "common.cuh"

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
  int splits = (2 * 148 + tiles - 1) / tiles;
  const int max_splits = ceil_div(K, 4 * GK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  return splits;
}

static size_t splitk_bytes(int M, int N, int K) {
  return align_up((size_t)pick_splits(M, N, K) * M * N * sizeof(float), 256);
}

}  // namespace dincae

using namespace dincae;

extern "C" size_t dincae_dense_workspace(int B, int K, int N) {
  const size_t a = splitk_bytes(B, N, K), b = splitk_bytes(B, K, N);
  return a > b ? a : b;
}

extern "C" int dincae_dense_fwd(const float* x, const float* w, const float* bias, int B, int K,
                                int N, int relu, float* y, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (!x || !w || !y || !workspace || B <= 0 || K <= 0 || N <= 0) return DINCAE_EINVAL;
  const int splits = pick_splits(B, N, K);
  if ((size_t)splits * B * N * sizeof(float) > workspace_bytes) return DINCAE_ENOSPC;
  int kps = ceil_div(ceil_div(K, splits), GK) * GK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ceil_div(N, GT), ceil_div(B, GT), ceil_div(K, kps));
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
  int kps = ceil_div(ceil_div(N, splits), GK) * GK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ceil_div(K, GT), ceil_div(B, GT), ceil_div(N, kps));
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
  launch_k(sgemm_kernel<1, 0>, grid, 256, 0, st, x, K, dz, N, K, N, B, ceil_div(B, GK) * GK, dw);
  int rc = check_launch();
  if (rc) return rc;
  launch_k(colsum_kernel, ceil_div(N, 128), 128, 0, st, dz, B, N, db);
  return check_launch();
}
