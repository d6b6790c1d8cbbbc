// bf16 helpers shared by the kind::f16 convolution kernels: the "planar-8" activation format
// (bf16 [B][CP8][H][W][8]: one pixel of one plane = 8 channels = 16 bytes, the same byte
// structure as planar-4 fp32, so tiles, TMA boxes and K-major descriptors carry over) and
// packing / unpacking of 8 channels.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace dincae {

__device__ __forceinline__ uint32_t bf2_pack(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);      // .x = lo (low 16 bits)
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ void bf8_unpack(const uint4& v, float (&f)[8]) {
  f[0] = bf_lo(v.x); f[1] = bf_hi(v.x);
  f[2] = bf_lo(v.y); f[3] = bf_hi(v.y);
  f[4] = bf_lo(v.z); f[5] = bf_hi(v.z);
  f[6] = bf_lo(v.w); f[7] = bf_hi(v.w);
}
__device__ __forceinline__ uint4 bf8_pack(const float (&f)[8]) {
  return make_uint4(bf2_pack(f[0], f[1]), bf2_pack(f[2], f[3]), bf2_pack(f[4], f[5]), bf2_pack(f[6], f[7]));
}
__device__ __forceinline__ uint4 u4_zero() { return make_uint4(0u, 0u, 0u, 0u); }

__device__ __forceinline__ uint4 ldg_u4_pred(const uint4* p, bool ok) {
  uint4 v;
  asm("{\n.reg .pred p;\nsetp.ne.b32 p, %5, 0;\n"
      "mov.b32 %0, 0;\nmov.b32 %1, 0;\nmov.b32 %2, 0;\nmov.b32 %3, 0;\n"
      "@p ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];\n}"
      : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "r"((int)ok));
  return v;
}
__device__ __forceinline__ unsigned ldg_u8_pred(const uint8_t* p, bool ok) {
  unsigned v;
  asm("{\n.reg .pred p;\n.reg .b16 t;\nsetp.ne.b32 p, %2, 0;\nmov.b16 t, 0;\n@p ld.global.nc.u8 t, [%1];\ncvt.u32.u16 %0, t;\n}"
      : "=r"(v) : "l"(p), "r"((int)ok));
  return v;
}

// un-pool + leaky-ReLU gradient of one pixel-plane: g (half-res gradient, 8 channels) times
// 1 / window count times (bit ? 1 : slope)
__device__ __forceinline__ uint4 bf8_unpool(const uint4& g, unsigned bits, float inv, float slope) {
  float f[8];
  bf8_unpack(g, f);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = f[j] * inv * (((bits >> j) & 1u) ? 1.0f : slope);
  return bf8_pack(f);
}

}  // namespace dincae
