// Inline-PTX building blocks for the tcgen05 kernels: mbarriers, shared-memory matrix
// descriptors (no-swizzle canonical layout), tcgen05.mma kind::tf32, TMEM loads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

namespace dincae {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// true in exactly one (converged) lane of the warp; ptxas keeps the tcgen05 issue loop of the
// elected lane on uniform registers (see profiles/r1_umma_microbench.md)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// Wait for a phase of an mbarrier: every lane of the (converged) warp polls.  Measured against
// a lane-0-only poll + __syncwarp (A/B bench runs of round 2): the single-lane variant takes the
// polling traffic off the shared-memory pipe but wakes the warp up later, and the config-A step
// (many short kernels) was 6.7 % slower with it (0.736 vs 0.690 ms); the bf16 kernels at 512x512
// did not change.  So: all lanes poll.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();      // never hang the device on a lost arrive
  }
}
// lane 0 polls, the other lanes park at the warp barrier (kept for A/B measurements)
__device__ __forceinline__ void mbar_wait_lane0(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
      if (++spins > (1u << 24)) __trap();
    }
  }
  __syncwarp();
}

// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
      :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// kind::f16 (bf16 / fp16 operands, K = 16 per instruction, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
      :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// instruction descriptor of kind::f16 with bf16 operands and fp32 accumulators; a_mn / b_mn = 1
// selects an MN-major (channel-contiguous) operand
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}


__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine, no tensor map): 16-byte aligned on both
// sides, size a multiple of 16; completes bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// TMA tensor-tile load (4-D tensor map, tiled mode): box -> shared memory, completes its
// bytes on `bar`; out-of-bounds elements are zero-filled
__device__ __forceinline__ void tma_load_4d(void* dst_smem, const void* tmap, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5}], [%6];"
      :: "r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst_smem, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];"
      :: "r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];"
      :: "r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// TMA tensor-tile store (2-D map): shared memory box -> global, clipped at the tensor bounds;
// tracked by the issuing thread's bulk async-group (bulk_commit / bulk_wait*)
__device__ __forceinline__ void tma_store_2d(const void* tmap, int c0, int c1, const void* src_smem) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
               :: "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(src_smem)) : "memory");
}
// 1-D bulk async copy shared -> global (16-byte aligned, size a multiple of 16)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the N most recent bulk groups of this thread have finished READING shared memory
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// arrive on `bar` when all tcgen05 operations issued so far by this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

// 16 lanes x 16 columns (tcgen05.ld.16x256b.x2): thread t of the warp receives, for lane a = t / 4
// and column pair c = 2 * (t % 4):  v[0..1] = (lane a, cols c, c+1), v[2..3] = (lane a+8, same cols),
// v[4..5] = (lane a, cols c+8, c+9), v[6..7] = (lane a+8, cols c+8, c+9).  `taddr` carries the first
// of the 16 lanes (a multiple of 16 inside the warp's quadrant).
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// Predicated read-only loads as single instructions: the loader issues all loads of a stage
// back to back (no branches between them), then converts and stores.
__device__ __forceinline__ float4 ldg_f4_pred(const float4* p, bool ok) {
  float4 v;
  asm("{\n.reg .pred p;\nsetp.ne.b32 p, %5, 0;\n"
      "mov.f32 %0, 0f00000000;\nmov.f32 %1, 0f00000000;\nmov.f32 %2, 0f00000000;\nmov.f32 %3, 0f00000000;\n"
      "@p ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];\n}"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "r"((int)ok));
  return v;
}
__device__ __forceinline__ unsigned ldg_u16_pred(const uint16_t* p, bool ok) {
  unsigned short v;
  asm("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\nmov.b16 %0, 0;\n@p ld.global.nc.u16 %0, [%1];\n}"
      : "=h"(v) : "l"(p), "r"((int)ok));
  return v;
}

// un-pool + leaky-ReLU gradient applied to a loaded half-resolution gradient (SRCM 1)
__device__ __forceinline__ float4 tc_unpool(float4 g, unsigned bits, int x, int W, float inv_y, float slope) {
  const float inv = inv_y * ((2 * (x >> 1) + 1 < W) ? 0.5f : 1.0f);
  bits >>= 4 * (x & 3);
  g.x = g.x * inv * ((bits & 1u) ? 1.0f : slope);
  g.y = g.y * inv * ((bits & 2u) ? 1.0f : slope);
  g.z = g.z * inv * ((bits & 4u) ? 1.0f : slope);
  g.w = g.w * inv * ((bits & 8u) ? 1.0f : slope);
  return g;
}

}  // namespace dincae
