// Dense bottleneck (tf.layers.dense, DINCAE/__init__.py:479-483) on the 5th-gen tensor cores for
// the bf16 path: y = x W and dx = dz W^T as tcgen05.mma kind::f16 GEMMs (fp32 accumulation in TMEM)
// that stream a bf16 copy of the kernel once.
//
// At BASELINE configs[3] the two kernels are 43008 x 8296 (1.38 GB each in fp32) and the batch is
// small, so both products are pure streams over W: the weight tile is the 128-row MMA operand A,
// the batch (padded to a multiple of 16) is the MMA N.  The bf16 copy is kept TILED,
//     W8 [K/8][N/8][8 k][8 n]           (one 8 x 8 block = 128 contiguous bytes),
// because that one block is at the same time
//   * the K-major core matrix of A = W   (rows k, 16 bytes of n)  -> dx = dz W^T, contraction over n
//   * the MN-major core matrix of A = W^T (8 k-rows of 16 bytes of n) -> y = x W, contraction over k
// of the no-swizzle canonical layouts (both checked on the device, tools/umma_bf16_test.cu), so ONE
// copy feeds both GEMMs and a 3-D TMA box over [K/8][N/8][128 B] lands in shared memory ready to
// use.  The small operand (x or dz) is converted to the same blocked form by a tiny kernel.
// Split over the contraction dimension across CTAs; partials [split][B][rows] are summed (fixed
// order) with bias / ReLU / ReLU-gate by the epilogue kernel of dense.cu's split path.
#include "bf16.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <string.h>

namespace dincae {

constexpr int DT_KS = 16;                  // contraction blocks (of 8) per pipeline stage = 8 MMAs of K = 16
constexpr int DT_STAGES = 5;                // at most; fewer when the batch operand is wide
constexpr int DT_THREADS = 32 * 6;         // warp 0: TMA, warp 1: MMA, warps 2-5: epilogue

struct DenseTcParams {
  const uint4* xb;          // small operand, blocked [cb][Bp/8][8 rows][8 contraction] bf16
  float* partial;           // [splits][B][rows_total]
  int mode;                 // 0: y = x W (rows = n, contraction = k); 1: dx = dz W^T (rows = k, contraction = n)
  int B, Bp;                // batch, padded to 16
  int rows_total;           // N (mode 0) or K (mode 1)
  int cblocks;              // contraction blocks of 8: K/8 (mode 0) or N/8 (mode 1)
  int cb_per_split;         // multiple of DT_KS
  int a_bytes, x_bytes, stage_bytes, n_stages;
};

__global__ void __launch_bounds__(DT_THREADS, 1)
dense_tc_kernel(const DenseTcParams P, const __grid_constant__ CUtensorMap tmw) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* ring = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  __shared__ __align__(8) uint64_t full_bar[DT_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[DT_STAGES];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mtile = blockIdx.x, split = blockIdx.y;
  const int cb0 = split * P.cb_per_split;
  int cb1 = cb0 + P.cb_per_split;
  if (cb1 > P.cblocks) cb1 = P.cblocks;
  const int n_stage_iters = (cb1 - cb0 + DT_KS - 1) / DT_KS;
  int tmem_cols = 32;
  while (tmem_cols < P.Bp) tmem_cols <<= 1;

  pdl_trigger();
  if (tid == 0) {
    for (int s = 0; s < P.n_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&tmw) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&tmem_holder)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;

  if (warp == 0) {
    // =============================== producer ================================================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < n_stage_iters; ++it) {
        const int cb = cb0 + it * DT_KS;
        uint32_t spins = 0;
        while (!mbar_try_wait(&empty_bar[s], ph ^ 1u))
          if (++spins > (1u << 24)) __trap();
        unsigned char* As = ring + (size_t)s * P.stage_bytes;
        unsigned char* Xs = As + P.a_bytes;
        mbar_expect_tx(&full_bar[s], (uint32_t)(P.a_bytes + P.x_bytes));
        // weight tile: coordinates (byte-in-block as 16 x f64, n-block, k-block)
        if (P.mode == 0)
          tma_load_3d(As, &tmw, 0, 16 * mtile, cb, &full_bar[s]);      // box {128 B, 16 nb, KS kb}
        else
          tma_load_3d(As, &tmw, 0, cb, 16 * mtile, &full_bar[s]);      // box {128 B, KS nb, 16 kb}
        bulk_g2s(Xs, P.xb + (size_t)cb * (P.Bp / 8) * 8, (uint32_t)P.x_bytes, &full_bar[s]);
        mbar_arrive(&full_bar[s]);                 // the one arrival; the phase completes with the bytes
        if (++s == P.n_stages) { s = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================================
    // mode 0: A = W^T tile, MN-major: n-blocks (M groups) at SBO = 128 B, k-blocks (K groups) at
    //         LBO = 16 * 128 B;  mode 1: A = W tile, K-major: k-blocks (M groups) at SBO = KS * 128 B,
    //         n-blocks (K groups) at LBO = 128 B.  B = blocked small operand, K-major: batch groups
    //         at SBO = 128 B, contraction blocks at LBO = (Bp/8) * 128 B.
    const uint32_t idesc = umma_idesc_bf16(128, P.Bp, P.mode == 0 ? 1 : 0, 0);
    const uint64_t a_hi = P.mode == 0 ? umma_desc(0u, 16u * 128u, 128u) : umma_desc(0u, 128u, (uint32_t)DT_KS * 128u);
    const uint64_t b_hi = umma_desc(0u, (uint32_t)(P.Bp / 8) * 128u, 128u);
    const uint32_t a_step = P.mode == 0 ? (2u * 16u * 128u) >> 4 : (2u * 128u) >> 4;    // two contraction blocks
    const uint32_t b_step = (2u * (uint32_t)(P.Bp / 8) * 128u) >> 4;
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < n_stage_iters; ++it) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = smem_u32(ring + (size_t)s * P.stage_bytes) >> 4;
        const uint32_t b_lo = a_lo + (uint32_t)(P.a_bytes >> 4);
#pragma unroll
        for (int j = 0; j < DT_KS / 2; ++j)
          umma_f16(tmem_base, a_hi | (uint64_t)((a_lo + j * a_step) & 0x3fffu),
                   b_hi | (uint64_t)((b_lo + j * b_step) & 0x3fffu), idesc, (it != 0 || j != 0) ? 1u : 0u);
        umma_commit(&empty_bar[s]);
        if (it == n_stage_iters - 1) umma_commit(&done_bar);
      }
      __syncwarp();
      if (++s == P.n_stages) { s = 0; ph ^= 1u; }
    }
  } else {
    // =============================== epilogue ================================================
    const int q = warp & 3;                       // TMEM lane quadrant of this warp
    const int row = 128 * mtile + 32 * q + lane;
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    float* out = P.partial + (size_t)split * P.B * P.rows_total;
    for (int c0 = 0; c0 < P.Bp; c0 += 8) {
      float a[8];
      tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, a);
      if (row < P.rows_total) {
#pragma unroll
        for (int t = 0; t < 8; ++t)
          if (c0 + t < P.B) out[(size_t)(c0 + t) * P.rows_total + row] = n_stage_iters > 0 ? a[t] : 0.f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(tmem_cols));
  }
}

// fp32 [K][ldw] (N columns) -> tiled bf16 W8 [K/8][N/8][8][8]; one thread per 8 n of one row
__global__ void __launch_bounds__(256) dense_tile_weights_kernel(const float* __restrict__ w, int ldw, int K, int N,
                                                                 uint4* __restrict__ w8) {
  pdl_trigger();
  pdl_wait();
  const int nb = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  const int NB = N >> 3;
  if (nb >= NB || k >= K) return;
  const float4 a = __ldg(reinterpret_cast<const float4*>(w + (size_t)k * ldw + 8 * nb));
  const float4 b = __ldg(reinterpret_cast<const float4*>(w + (size_t)k * ldw + 8 * nb + 4));
  const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  w8[((size_t)(k >> 3) * NB + nb) * 8 + (k & 7)] = bf8_pack(f);
}

// small operand fp32 [B][ld] (C columns) -> blocked bf16 [cb_pad][Bp/8][8 rows][8 cols], zero padded
__global__ void __launch_bounds__(256) dense_block_operand_kernel(const float* __restrict__ x, int ld, int B, int C,
                                                                  int Bp, int cb_pad, uint4* __restrict__ xb) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;       // over cb_pad * Bp
  if (i >= cb_pad * Bp) return;
  const int cb = i / Bp, b = i - cb * Bp;
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = 8 * cb + j;
    f[j] = (b < B && c < C) ? __ldg(x + (size_t)b * ld + c) : 0.f;
  }
  xb[((size_t)cb * (Bp >> 3) + (b >> 3)) * 8 + (b & 7)] = bf8_pack(f);
}

// y[b][n] = epi(sum_s part[s][b][n] + bias[n]); epi: relu and/or gate (act_in > 0)
__global__ void __launch_bounds__(256) dense_tc_epilogue_kernel(const float* __restrict__ part, int splits, int B,
                                                                int N, int ldy, const float* __restrict__ bias,
                                                                int relu, const float* __restrict__ gate,
                                                                float* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * N;
  if (e >= total) return;
  const int b = (int)(e / N), n = (int)(e - (size_t)b * N);
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += part[(size_t)k * total + e];
  if (bias) s += __ldg(bias + n);
  if (relu) s = fmaxf(s, 0.f);
  if (gate) s = __ldg(gate + (size_t)b * ldy + n) > 0.f ? s : 0.f;
  y[(size_t)b * ldy + n] = s;
}

static int dt_encode_map(CUtensorMap* tm, const void* w8, int KB, int NB, int box_nb, int box_kb) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return DINCAE_ECUDA;
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  // [KB][NB][128 bytes] seen as f64 [KB][NB][16]
  const cuuint64_t dims[3] = {16, (cuuint64_t)NB, (cuuint64_t)KB};
  const cuuint64_t strides[2] = {128, (cuuint64_t)NB * 128};
  const cuuint32_t box[3] = {16, (cuuint32_t)box_nb, (cuuint32_t)box_kb};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(w8), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_cuda_error = (int)r;
    return DINCAE_ECUDA;
  }
  return DINCAE_OK;
}

static int dt_splits(int mtiles, int cblocks, int sm_count) {
  int splits = (2 * sm_count + mtiles - 1) / mtiles;
  const int max_splits = (cblocks + 4 * DT_KS - 1) / (4 * DT_KS);
  if (splits > max_splits) splits = max_splits;
  return splits < 1 ? 1 : splits;
}

}  // namespace dincae

using namespace dincae;

extern "C" size_t dincae_dense_tc_weights_elems(int K, int N) { return (size_t)K * N; }

extern "C" int dincae_dense_tc_tile_weights(const float* w, int ldw, int K, int N, void* w8, void* stream) {
  if (!w || !w8 || K <= 0 || N <= 0 || (K & 7) || (N & 7) || (ldw & 3) || K > 65535 * 1) return DINCAE_EINVAL;
  launch_k(dense_tile_weights_kernel, dim3(ceil_div(N / 8, 256), K), 256, 0, (cudaStream_t)stream, w, ldw, K, N,
           reinterpret_cast<uint4*>(w8));
  return check_launch();
}

// split plan of one product; returns the workspace bytes it needs (blocked operand, then partials)
static size_t dt_plan(int mode, int B, int K, int N, int sm_count, DenseTcParams& P, int* n_splits_out,
                      int* cb_pad_out, size_t* xb_bytes_out) {
  P.mode = mode;
  P.B = B;
  P.Bp = (B + 15) / 16 * 16;
  P.rows_total = mode == 0 ? N : K;
  P.cblocks = (mode == 0 ? K : N) / 8;
  const int mtiles = ceil_div(P.rows_total, 128);
  const int splits = dt_splits(mtiles, P.cblocks, sm_count);
  P.cb_per_split = ceil_div(ceil_div(P.cblocks, splits), DT_KS) * DT_KS;
  const int n_splits = ceil_div(P.cblocks, P.cb_per_split);
  const int cb_pad = n_splits * P.cb_per_split;
  P.a_bytes = 16 * DT_KS * 128;
  P.x_bytes = DT_KS * (P.Bp / 8) * 128;
  P.stage_bytes = P.a_bytes + P.x_bytes;
  P.n_stages = (200 * 1024) / P.stage_bytes;
  if (P.n_stages > DT_STAGES) P.n_stages = DT_STAGES;
  const size_t xb_bytes = align_up((size_t)cb_pad * P.Bp * 16, 256);
  *n_splits_out = n_splits;
  *cb_pad_out = cb_pad;
  *xb_bytes_out = xb_bytes;
  return xb_bytes + align_up((size_t)n_splits * B * P.rows_total * sizeof(float), 256);
}

extern "C" size_t dincae_dense_tc_workspace(int B, int K, int N) {
  if (B <= 0 || K <= 0 || N <= 0) return 0;
  DenseTcParams P = {};
  int ns, cp;
  size_t xb;
  // the SM count only shapes the split; 148 (B200) and a smaller part give the same upper bound class
  const size_t a = dt_plan(0, B, K, N, 148, P, &ns, &cp, &xb);
  const size_t b = dt_plan(1, B, K, N, 148, P, &ns, &cp, &xb);
  return a > b ? a : b;
}

// mode 0: y[B][N] = act(x[B][K] W + bias); mode 1: dx[B][K] = (dz[B][N] W^T) * (gate > 0)
static int dense_tc(int mode, const float* small, int ld_small, const void* w8, int K, int N, int B,
                    const float* bias, int relu, const float* gate, float* out, int ld_out,
                    void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if ((K & 7) || (N & 7) || B > 256) return DINCAE_EINVAL;
  DenseTcParams P = {};
  int n_splits, cb_pad;
  size_t xb_bytes;
  const size_t need = dt_plan(mode, B, K, N, 148, P, &n_splits, &cb_pad, &xb_bytes);
  if (P.n_stages < 2) return DINCAE_EINVAL;
  if (workspace_bytes < need) return DINCAE_ENOSPC;
  const int mtiles = ceil_div(P.rows_total, 128);
  unsigned char* wsb = reinterpret_cast<unsigned char*>(workspace);
  P.xb = reinterpret_cast<const uint4*>(wsb);
  P.partial = reinterpret_cast<float*>(wsb + xb_bytes);
  const int C = mode == 0 ? K : N;
  launch_k(dense_block_operand_kernel, ceil_div(cb_pad * P.Bp, 256), 256, 0, st, small, ld_small, B, C, P.Bp, cb_pad,
           reinterpret_cast<uint4*>(wsb));
  int rc = check_launch();
  if (rc) return rc;
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  rc = mode == 0 ? dt_encode_map(&tm, w8, K / 8, N / 8, 16, DT_KS) : dt_encode_map(&tm, w8, K / 8, N / 8, DT_KS, 16);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return cuda_rc(e);
    attr_set = true;
  }
  const size_t smem = (size_t)P.n_stages * P.stage_bytes + 128;
  launch_k(dense_tc_kernel, dim3(mtiles, n_splits), DT_THREADS, smem, st, P, tm);
  rc = check_launch();
  if (rc) return rc;
  const size_t total = (size_t)B * P.rows_total;
  launch_k(dense_tc_epilogue_kernel, (unsigned)((total + 255) / 256), 256, 0, st, P.partial, n_splits, B, P.rows_total,
           ld_out, bias, relu, gate, out);
  return check_launch();
}

extern "C" int dincae_dense_tc_fwd(const float* x, int ldx, const void* w8, const float* bias, int B, int K, int N,
                                   int relu, float* y, int ldy, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  if (!x || !w8 || !y || !workspace || B <= 0 || K <= 0 || N <= 0) return DINCAE_EINVAL;
  return dense_tc(0, x, ldx, w8, K, N, B, bias, relu, nullptr, y, ldy, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int dincae_dense_tc_bwd_data(const float* dz, int lddz, const void* w8, const float* act_in, int B, int K,
                                        int N, float* dx, int lddx, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  if (!dz || !w8 || !dx || !workspace || B <= 0 || K <= 0 || N <= 0) return DINCAE_EINVAL;
  return dense_tc(1, dz, lddz, w8, K, N, B, nullptr, 0, act_in, dx, lddx, workspace, workspace_bytes,
                  (cudaStream_t)stream);
}
