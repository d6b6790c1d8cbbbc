// Shared helpers for the dincae_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dincae_b200.h"

namespace dincae {

extern int g_last_cuda_error;
extern long long g_launch_count;     // diagnostics only: kernels enqueued through this library

inline int check_launch() {
  ++g_launch_count;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    return DINCAE_ECUDA;
  }
  return DINCAE_OK;
}

inline int cuda_rc(cudaError_t e) {
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    return DINCAE_ECUDA;
  }
  return DINCAE_OK;
}

// Programmatic dependent launch (experimental, DINCAE_PDL=1): every kernel of the library can
// be launched with the programmatic-stream-serialization attribute; it raises
// launch_dependents at its start and executes griddepcontrol.wait before its first global
// access, so that the launch latency and the local prologue (barrier init, TMEM allocation,
// shared-memory clears) of kernel N+1 overlap the tail of kernel N.  Without the attribute
// the two instructions are no-ops.
extern int g_use_pdl;
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);      // errors: check_launch()
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_fma(float s, float4 b, float4 c) {
  return make_float4(fmaf(s, b.x, c.x), fmaf(s, b.y, c.y), fmaf(s, b.z, c.z), fmaf(s, b.w, c.w));
}
__device__ __forceinline__ float4 f4_scale(float4 a, float s) {
  return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ float f4_dot(float4 a, float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  acc = fmaf(a.w, b.w, acc);
  return acc;
}
__device__ __forceinline__ float f4_get(const float4& a, int i) {
  return i == 0 ? a.x : (i == 1 ? a.y : (i == 2 ? a.z : a.w));
}

// round-to-nearest (ties away from zero) to the 10-bit TF32 mantissa == cvt.rna.tf32.f32 for
// finite values; two full-rate integer ops instead of the multi-instruction cvt expansion
__device__ __forceinline__ float to_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ float4 to_tf32(float4 v) {
  return make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
}

// Tensor-core back end selected (dincae_set_conv_backend): every kernel that produces a tensor
// a convolution reads rounds it to TF32 on store, because tcgen05 would otherwise truncate.
extern int g_conv_backend;      // 0: fp32 FFMA kernels, 1: tcgen05 TF32 implicit GEMM

// Logical conv input U[b][plane][y][x] (one float4 = 4 channels) assembled on the fly:
//   mode 0: concat(s0 [p0 planes, optionally stored at half resolution], s1 [p1 planes])
//   mode 1: gradient of an encoder stage: gpool[y>>1][x>>1] / window_count * lrelu'(sign)
struct ConvSrc {
  const float4* s0;
  const float4* s1;
  const uint16_t* sign;
  int p0, p1;
  int half0;
  int mode;
  float slope;
  int H, W, Hh, Wh, Wq;   // full size, half size, ceil(W/4)
  __host__ __device__ __forceinline__ int planes() const { return p0 + p1; }
};

__device__ __forceinline__ float4 conv_src_load(const ConvSrc& s, int b, int plane, int y, int x) {
  if ((unsigned)y >= (unsigned)s.H || (unsigned)x >= (unsigned)s.W) return f4_zero();
  if (s.mode == 0) {
    if (plane < s.p0) {
      if (s.half0)
        return __ldg(&s.s0[((size_t)(b * s.p0 + plane) * s.Hh + (y >> 1)) * s.Wh + (x >> 1)]);
      return __ldg(&s.s0[((size_t)(b * s.p0 + plane) * s.H + y) * s.W + x]);
    }
    plane -= s.p0;
    return __ldg(&s.s1[((size_t)(b * s.p1 + plane) * s.H + y) * s.W + x]);
  }
  const int yh = y >> 1, xh = x >> 1;
  float4 g = __ldg(&s.s0[((size_t)(b * s.p0 + plane) * s.Hh + yh) * s.Wh + xh]);
  const int cy = (2 * yh + 1 < s.H) ? 2 : 1;
  const int cx = (2 * xh + 1 < s.W) ? 2 : 1;
  const float inv = 1.0f / (float)(cy * cx);
  unsigned bits = __ldg(&s.sign[((size_t)(b * s.p0 + plane) * s.H + y) * s.Wq + (x >> 2)]);
  bits >>= 4 * (x & 3);
  g.x = g.x * inv * ((bits & 1u) ? 1.0f : s.slope);
  g.y = g.y * inv * ((bits & 2u) ? 1.0f : s.slope);
  g.z = g.z * inv * ((bits & 4u) ? 1.0f : s.slope);
  g.w = g.w * inv * ((bits & 8u) ? 1.0f : s.slope);
  return g;
}

}  // namespace dincae
