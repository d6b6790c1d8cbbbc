// Decode which shared-memory words tcgen05.mma kind::tf32 reads for an MN-major operand:
// the other operand is an identity (K-major, known good), the probed operand's region holds
// its own word index, so D directly lists the word read for every (row, k).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../dincae_b200/csrc/umma.cuh"
using namespace dincae;

struct Cfg { int probe_a; int mn; int sbo, lbo; int M, N; };

__global__ void __launch_bounds__(128) probe(Cfg c, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* P = reinterpret_cast<float*>(smem);                 // probed operand: 8 KB of indices
  float* I = reinterpret_cast<float*>(smem + 64 * 1024);     // identity operand, K-major, SBO 128, LBO 2048
  for (int i = tid; i < 2048; i += blockDim.x) P[i] = (float)i;
  for (int i = tid; i < 8192; i += blockDim.x) I[i] = 0.f;
  __syncthreads();
  if (tid < 8) {   // row r=tid, k=tid -> 1
    const int r = tid, k = tid;
    I[((r >> 3) * 128 + (r & 7) * 16 + (k >> 2) * 2048 + (k & 3) * 4) / 4] = 1.f;
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_holder)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_holder;
  if (tid == 0) {
    const uint32_t amn = c.probe_a ? c.mn : 0, bmn = c.probe_a ? 0 : c.mn;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (amn << 15) | (bmn << 16) |
                           ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    const uint64_t pd = umma_desc(smem_u32(P), c.lbo, c.sbo);
    const uint64_t id = umma_desc(smem_u32(I), 2048, 128);
    umma_tf32(tmem, c.probe_a ? pd : id, c.probe_a ? id : pd, idesc, 0u);
    umma_commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[16];
  for (int c16 = 0; c16 < c.N / 16; ++c16) {
    tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(16 * c16), v);
    for (int j = 0; j < 16; ++j) out[tid * c.N + 16 * c16 + j] = v[j];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(64));
}

int main() {
  float* out; cudaMalloc(&out, 128 * 64 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  static float h[128 * 64];
  Cfg cases[] = {
    {0, 0, 128, 1024, 128, 32},   // B K-major reference: expect word(n,k) = (n/8)*32 + (n%8)*4 + (k/4)*256 + k%4
    {0, 1, 640, 128, 128, 32},    // B MN-major
    {0, 1, 128, 640, 128, 32},    // B MN-major swapped
    {1, 1, 640, 128, 128, 16},    // A MN-major (identity on B side: N=16, n<8)
    {1, 1, 128, 640, 128, 16},
  };
  for (auto& c : cases) {
    cudaMemset(out, 0, 128 * 64 * 4);
    probe<<<1, 128, 96 * 1024>>>(c, out);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, 128 * c.N * 4, cudaMemcpyDeviceToHost);
    printf("probe_%s mn=%d sbo=%d lbo=%d: %s\n", c.probe_a ? "A" : "B", c.mn, c.sbo, c.lbo, cudaGetErrorString(e));
    if (!c.probe_a) {
      // D[m=k][n] = word of B(n,k)
      for (int k = 0; k < 8; ++k) {
        printf("  k=%d:", k);
        for (int n = 0; n < c.N; ++n) printf(" %4g", h[k * c.N + n]);
        printf("\n");
      }
    } else {
      // D[m][n=k] = word of A(m,k); print m = 0..15 and 64..67
      for (int k = 0; k < 8; ++k) {
        printf("  k=%d:", k);
        for (int m = 0; m < 24; ++m) printf(" %4g", h[m * c.N + k]);
        printf("  | m=64..67:");
        for (int m = 64; m < 68; ++m) printf(" %4g", h[m * c.N + k]);
        printf("\n");
      }
    }
  }
  return 0;
}
