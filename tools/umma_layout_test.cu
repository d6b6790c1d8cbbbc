// Correctness probe for tcgen05.mma kind::tf32 shared-memory descriptors (no swizzle):
// K-major and MN-major operands with the strides the conv kernels use, M = 128 and M = 64
// (TMEM row -> lane mapping).  One MMA (K = 8), D read back from TMEM, compared on the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_layout_test umma_layout_test.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include "../dincae_b200/csrc/umma.cuh"

using namespace dincae;

struct Cfg {
  int M, N;
  int a_mn, b_mn;
  int a_sbo, a_lbo, b_sbo, b_lbo;   // bytes
};

__host__ __device__ inline float a_val(int m, int k) { return (float)((m * 37 + k * 11) % 127 - 63); }
__host__ __device__ inline float b_val0(int n, int k) { return (float)((n * 29 + k * 13) % 61 - 30); }
// with an N-group stride of 16 bytes consecutive groups alias consecutive K rows ("pixels")
__host__ __device__ inline float b_elem(int sbo, int n, int k) {
  return sbo == 16 ? b_val0(n & 3, k + (n >> 2)) : b_val0(n, k);
}

// byte offset of element (r, k) of an operand: r = M/N index, k = K index (0..7)
__host__ __device__ inline int elem_off(int mn_major, int r, int k, int sbo, int lbo) {
  if (!mn_major) return (r >> 3) * sbo + (r & 7) * 16 + (k >> 2) * lbo + (k & 3) * 4;
  return (r >> 2) * sbo + (r & 3) * 4 + (k & 7) * 16 + (k >> 3) * lbo;
}

__global__ void __launch_bounds__(128) probe(Cfg c, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1000.f;
  __syncthreads();
  unsigned char* As = smem;
  unsigned char* Bs = smem + 96 * 1024;
  for (int i = tid; i < c.M * 8; i += blockDim.x) {
    const int m = i / 8, k = i % 8;
    *reinterpret_cast<float*>(As + elem_off(c.a_mn, m, k, c.a_sbo, c.a_lbo)) = a_val(m, k);
  }
  for (int i = tid; i < c.N * 8; i += blockDim.x) {
    const int n = i / 8, k = i % 8;
    *reinterpret_cast<float*>(Bs + elem_off(c.b_mn, n, k, c.b_sbo, c.b_lbo)) = b_elem(c.b_sbo, n, k);
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
  // clear the accumulator region first (32 columns, all lanes) via an MMA on zeros? simpler:
  // use accumulate = 0 and read every lane; lanes not written keep stale data -> mark by NaN
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)c.a_mn << 15) |
                           ((uint32_t)c.b_mn << 16) | ((uint32_t)(c.N >> 3) << 17) |
                           ((uint32_t)(c.M >> 4) << 24);
    const uint64_t a0 = umma_desc(smem_u32(As), c.a_lbo, c.a_sbo);
    const uint64_t b0 = umma_desc(smem_u32(Bs), c.b_lbo, c.b_sbo);
    umma_tf32(tmem, a0, b0, idesc, 0u);
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
      {"M128 N16 A K-major  B K-major ", {128, 16, 0, 0, 1824, 20000, 128, 256}},
      {"M128 N32 A K-major  B MN-major", {128, 32, 0, 1, 1824, 20000, 640, 128}},
      {"M128 N16 A MN-major B MN-major", {128, 16, 1, 1, 1824, 128, 928, 128}},
      {"M64  N16 A MN-major B MN-major", {64, 16, 1, 1, 1824, 128, 928, 128}},
      {"M64  N16 A K-major  B K-major ", {64, 16, 0, 0, 1824, 20000, 128, 256}},
      {"M128 N16 A K-major  B MN sbo16", {128, 16, 0, 1, 1824, 20000, 16, 128}},
  };
  static float h[128 * 64];
  for (auto& nc : cases) {
    const Cfg& c = nc.c;
    cudaMemset(out, 0xff, 128 * 64 * 4);
    probe<<<1, 128, 165 * 1024>>>(c, out);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, 128 * c.N * 4, cudaMemcpyDeviceToHost);
    // find, for every logical row m, which TMEM lane holds it
    int bad = 0, found = 0;
    int lane_of_row[128];
    for (int m = 0; m < c.M; ++m) {
      float ref[64];
      for (int n = 0; n < c.N; ++n) {
        float s = 0.f;
        for (int k = 0; k < 8; ++k) {
          s += a_val(m, k) * b_elem(c.b_sbo, n, k);
        }
        ref[n] = s;
      }
      lane_of_row[m] = -1;
      for (int l = 0; l < 128; ++l) {
        bool ok = true;
        for (int n = 0; n < c.N && ok; ++n) ok = (h[l * c.N + n] == ref[n]);
        if (ok) { lane_of_row[m] = l; break; }
      }
      if (lane_of_row[m] < 0) ++bad; else ++found;
    }
    printf("%s: %s  rows matched %d / %d", nc.name, cudaGetErrorString(e), found, c.M);
    if (found) {
      printf("  lane(row): 0->%d 1->%d 15->%d 16->%d 31->%d 32->%d 63->%d", lane_of_row[0], lane_of_row[1],
             lane_of_row[15], lane_of_row[16], lane_of_row[31], lane_of_row[32], lane_of_row[63]);
      if (c.M == 128) printf(" 64->%d 127->%d", lane_of_row[64], lane_of_row[127]);
    }
    printf("\n");
  }
  return 0;
}
