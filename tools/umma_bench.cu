// Microbenchmark: cycles per tcgen05.mma kind::tf32 (M=128|64, K=8) issued by one thread,
// operands in shared memory, no-swizzle descriptors with the strides the conv kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu && ./umma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../dincae_b200/csrc/umma.cuh"

using namespace dincae;

struct Cfg {
  int M, N;            // MMA shape
  int a_mn, b_mn;      // 1 = MN-major operand
  int a_sbo, a_lbo;    // bytes
  int b_sbo, b_lbo;
  int n_mma;           // MMAs per timed batch
  int shift_units;     // A start-address step between consecutive MMAs (16-byte units)
  int n_acc;           // distinct accumulators cycled (columns = n_acc * N)
};

__global__ void __launch_bounds__(128) bench_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&tmem_holder)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)c.a_mn << 15) |
                           ((uint32_t)c.b_mn << 16) | ((uint32_t)(c.N >> 3) << 17) |
                           ((uint32_t)(c.M >> 4) << 24);
    const uint64_t a0 = umma_desc(smem_u32(smem), c.a_lbo, c.a_sbo);
    const uint64_t b0 = umma_desc(smem_u32(smem + 128 * 1024), c.b_lbo, c.b_sbo);
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      uint32_t sh = 0;
      for (int i = 0; i < c.n_mma; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t d = tmem + (uint32_t)((u & (c.n_acc - 1)) * c.N);
          umma_tf32(d, a0 + sh + (uint32_t)(u * c.shift_units), b0, idesc, 1u);
        }
        sh = (sh + 8 * c.shift_units) & 1023;
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, phase);
      phase ^= 1;
      const long long t2 = clock64();
      if (blockIdx.x == 0) { out[2 * rep] = t1 - t0; out[2 * rep + 1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

int main() {
  long long* out;
  cudaMalloc(&out, 64);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  // name, cfg
  struct Named { const char* name; Cfg c; } cases[] = {
      {"fwd  M128 N16  sbo128  (linear tile)", {128, 16, 0, 0, 128, 20000, 128, 256, 4096, 1, 8}},
      {"fwd  M128 N16  sbo1824 (114-px rows)", {128, 16, 0, 0, 1824, 20000, 128, 256, 4096, 1, 8}},
      {"fwd  M128 N16  sbo928  (58-px rows)", {128, 16, 0, 0, 928, 20000, 128, 256, 4096, 1, 8}},
      {"fwd  M128 N16  sbo1024 (64-px rows)", {128, 16, 0, 0, 1024, 20000, 128, 256, 4096, 1, 8}},
      {"fwd  M128 N32  sbo1824", {128, 32, 0, 0, 1824, 20000, 128, 512, 4096, 1, 8}},
      {"fwd  M128 N48  sbo1824", {128, 48, 0, 0, 1824, 20000, 128, 768, 4096, 1, 8}},
      {"fwd  M128 N64  sbo1824", {128, 64, 0, 0, 1824, 20000, 128, 1024, 4096, 1, 8}},
      {"fwd  M128 N128 sbo1824", {128, 128, 0, 0, 1824, 20000, 128, 2048, 4096, 1, 4}},
      {"fwd  M128 N256 sbo1824", {128, 256, 0, 0, 1824, 20000, 128, 4096, 4096, 1, 2}},
      {"fwd  M64  N16  sbo1824", {64, 16, 0, 0, 1824, 20000, 128, 256, 4096, 1, 8}},
      {"fwd  M128 N16  same acc", {128, 16, 0, 0, 1824, 20000, 128, 256, 4096, 1, 1}},
      {"fwd  M128 N16  no shift", {128, 16, 0, 0, 1824, 20000, 128, 256, 4096, 0, 8}},
      {"dgr  M128 N16  B mn-major", {128, 16, 0, 1, 1824, 20000, 256, 128, 4096, 1, 8}},
      {"dgr  M128 N96  B mn-major", {128, 96, 0, 1, 1824, 20000, 768, 128, 4096, 1, 4}},
      {"wgr  M128 N16  A,B mn-major sbo1824", {128, 16, 1, 1, 1824, 128, 928, 128, 4096, 1, 2}},
      {"wgr  M64  N16  A,B mn-major sbo1824", {64, 16, 1, 1, 1824, 128, 928, 128, 4096, 1, 2}},
      {"wgr  M128 N64  A,B mn-major sbo1824", {128, 64, 1, 1, 1824, 128, 928, 128, 4096, 1, 2}},
      {"wgr  M64  N64  A,B mn-major sbo1824", {64, 64, 1, 1, 1824, 128, 928, 128, 4096, 1, 2}},
      // K-major transposed tiles of the weight gradient: 144-byte core-matrix pitch (bank pad)
      // vs 128-byte aligned core matrices; shift = one K step (2 pixel groups)
      {"wgK  M64  N48  K-major sbo144 lbo304/880", {64, 48, 0, 0, 144, 304, 144, 880, 4096, 38, 1}},
      {"wgK  M64  N48  K-major sbo128 lbo256/768", {64, 48, 0, 0, 128, 256, 128, 768, 4096, 32, 1}},
      {"wgK  M64  N24  K-major sbo144 lbo592/432", {64, 24, 0, 0, 144, 592, 144, 432, 4096, 74, 1}},
      {"wgK  M64  N24  K-major sbo128 lbo512/384", {64, 24, 0, 0, 128, 512, 128, 384, 4096, 64, 1}},
      {"wgK  M128 N120 K-major sbo144 lbo1744", {128, 120, 0, 0, 144, 1744, 144, 2176, 4096, 218, 1}},
      {"wgK  M128 N120 K-major sbo128 lbo1536", {128, 120, 0, 0, 128, 1536, 128, 1920, 4096, 192, 1}},
  };
  for (auto& nc : cases) {
    for (int grid : {1, 148}) {
      bench_kernel<<<grid, 128, 170 * 1024>>>(nc.c, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[6];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%-42s grid %3d: issue %.1f cyc/mma, complete %.1f cyc/mma  (%s)\n", nc.name, grid,
             (double)h[4] / nc.c.n_mma, (double)h[5] / nc.c.n_mma, cudaGetErrorString(e));
    }
  }
  return 0;
}
