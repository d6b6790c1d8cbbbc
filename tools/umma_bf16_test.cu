// Correctness probe for tcgen05.mma kind::f16 (bf16 operands) shared-memory descriptors without
// swizzle: K-major and MN-major operands with the strides the bf16 conv kernels use, M = 128
// and M = 64.  One MMA (K = 16), D read back from TMEM, compared on the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/umma_bf16_test tools/umma_bf16_test.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../dincae_b200/csrc/umma.cuh"

using namespace dincae;

struct Cfg {
  int M, N;
  int a_mn, b_mn;
  int a_sbo, a_lbo, b_sbo, b_lbo;   // bytes
  int a_shift;                      // extra byte offset of the A start address (a tap shift)
};

__host__ __device__ inline float a_val(int m, int k) { return (float)((m * 37 + k * 11) % 127 - 63); }
__host__ __device__ inline float b_val(int n, int k) { return (float)((n * 29 + k * 13) % 61 - 30); }

// byte offset of element (r, k): r = M/N index, k = K index (0..15), 2-byte elements
__host__ __device__ inline int elem_off(int mn_major, int r, int k, int sbo, int lbo) {
  if (!mn_major) return (r >> 3) * sbo + (r & 7) * 16 + (k >> 3) * lbo + (k & 7) * 2;
  return (r >> 3) * sbo + (r & 7) * 2 + (k & 7) * 16 + (k >> 3) * lbo;
}

__global__ void __launch_bounds__(128) probe(Cfg c, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x447a447au;  // 1000
  __syncthreads();
  unsigned char* As = smem + c.a_shift;
  unsigned char* Bs = smem + 96 * 1024;
  for (int i = tid; i < c.M * 16; i += blockDim.x) {
    const int m = i / 16, k = i % 16;
    *reinterpret_cast<__nv_bfloat16*>(As + elem_off(c.a_mn, m, k, c.a_sbo, c.a_lbo)) = __float2bfloat16(a_val(m, k));
  }
  for (int i = tid; i < c.N * 16; i += blockDim.x) {
    const int n = i / 16, k = i % 16;
    *reinterpret_cast<__nv_bfloat16*>(Bs + elem_off(c.b_mn, n, k, c.b_sbo, c.b_lbo)) = __float2bfloat16(b_val(n, k));
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&tmem_holder)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(c.M, c.N, c.a_mn, c.b_mn);
    const uint64_t a0 = umma_desc(smem_u32(As), c.a_lbo, c.a_sbo);
    const uint64_t b0 = umma_desc(smem_u32(Bs), c.b_lbo, c.b_sbo);
    umma_f16(tmem, a0, b0, idesc, 0u);
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
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(64));
}

int main() {
  float* out;
  cudaMalloc(&out, 128 * 64 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  struct Named { const char* name; Cfg c; } cases[] = {
      // forward conv: A = pixels (K-major, LBO = plane stride, SBO = row pitch), B = weights
      {"M128 N16 A K-major  B K-major ", {128, 16, 0, 0, 1824, 20000, 128, 256, 0}},
      {"M128 N32 A K-major  B K-major ", {128, 32, 0, 0, 1056, 19008, 128, 512, 16}},
      // weight gradient: both operands channel-contiguous (MN-major), K = 16 pixels of a row
      {"M128 N32 A MN-major B MN-major", {128, 32, 1, 1, 4224, 128, 2048, 128, 0}},
      {"M128 N32 A MN (+16B) B MN     ", {128, 32, 1, 1, 4224, 128, 2048, 128, 16}},
      {"M64  N32 A MN-major B MN-major", {64, 32, 1, 1, 4224, 128, 2048, 128, 0}},
      {"M64  N16 A MN-major B MN-major", {64, 16, 1, 1, 4224, 128, 2048, 128, 32}},
      {"M128 N48 A MN-major B K-major ", {128, 48, 1, 0, 4224, 128, 128, 768, 0}},
      // swapped roles of LBO / SBO (in case the no-swizzle MN-major form reads them the other way)
      {"M128 N32 A MN swap  B MN swap ", {128, 32, 1, 1, 128, 4224, 128, 2048, 0}},
  };
  static float h[128 * 64];
  int fails = 0;
  for (auto& nc : cases) {
    const Cfg& c = nc.c;
    cudaMemset(out, 0xff, 128 * 64 * 4);
    probe<<<1, 128, 165 * 1024>>>(c, out);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, 128 * c.N * 4, cudaMemcpyDeviceToHost);
    int found = 0, inplace = 0;
    int lane_of_row[128];
    for (int m = 0; m < c.M; ++m) {
      float ref[64];
      for (int n = 0; n < c.N; ++n) {
        float s = 0.f;
        for (int k = 0; k < 16; ++k) s += a_val(m, k) * b_val(n, k);
        ref[n] = s;
      }
      lane_of_row[m] = -1;
      for (int l = 0; l < 128; ++l) {
        bool ok = true;
        for (int n = 0; n < c.N && ok; ++n) ok = (h[l * c.N + n] == ref[n]);
        if (ok) { lane_of_row[m] = l; break; }
      }
      if (lane_of_row[m] >= 0) ++found;
      if (lane_of_row[m] == (c.M == 128 ? m : (m / 16) * 32 + m % 16)) ++inplace;
    }
    printf("%s: %s  rows matched %d / %d (expected lane: %d)", nc.name, cudaGetErrorString(e), found, c.M, inplace);
    if (found) printf("  lane(row): 0->%d 1->%d 15->%d 16->%d 31->%d 32->%d 63->%d", lane_of_row[0], lane_of_row[1],
                      lane_of_row[15], lane_of_row[16], lane_of_row[31], lane_of_row[32], lane_of_row[63]);
    else printf("  D[0][0..3] = %g %g %g %g", h[0], h[1], h[2], h[3]);
    printf("\n");
    if (found != c.M) ++fails;
  }
  printf("%s\n", fails <= 1 ? "PROBE DONE" : "PROBE: several layouts failed");
  return 0;
}
