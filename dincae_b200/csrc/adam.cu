// Global-norm clipping + TF-Adam over one flat parameter buffer
// (DINCAE/__init__.py:563-567, 598-603): two HBM-bound passes -- a sum-of-squares reduction
// whose last block also derives the clip scale and the bias-corrected step size, then the
// vectorised update.  28 bytes/parameter/step (read g twice, p, m, v; write p, m, v).
#include "common.cuh"

namespace dincae {

struct RedWork {
  unsigned int counter;
  unsigned int pad[3];
  double partial[1];
};

static int red_blocks(size_t n) {
  size_t b = (n + 4095) / 4096;
  if (b > 8 * 148) b = 8 * 148;
  if (b < 1) b = 1;
  return (int)b;
}

// MODE 0: sum over (g*scale + decay*p)^2 and finalise the Adam step scalars.
// MODE 1: 0.5 * sum p^2 -> out[0].
template <int MODE>
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ p,
                                                    const float* __restrict__ g, size_t n,
                                                    size_t n_decay, float l2_beta, float grad_scale,
                                                    const double* __restrict__ denom,
                                                    float clip_norm, const float* __restrict__ lr,
                                                    float beta1, float beta2, float* __restrict__ state,
                                                    RedWork* __restrict__ work) {
  pdl_trigger();
  pdl_wait();
  float s = 0.f;
  if (MODE == 0 && denom) grad_scale /= (float)(*denom);
  const bool decay = MODE == 0 && l2_beta != 0.f;
  const size_t n4 = n >> 2;
  size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (!decay) {
    // plain sum of squares of one array: four independent 16-byte loads in flight per thread
    const float4* src = reinterpret_cast<const float4*>(MODE == 0 ? g : p);
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (; i0 + 3 * stride < n4; i0 += 4 * stride) {
      const float4 a0 = __ldg(src + i0), a1 = __ldg(src + i0 + stride);
      const float4 a2 = __ldg(src + i0 + 2 * stride), a3 = __ldg(src + i0 + 3 * stride);
      s = f4_dot(a0, a0, s);
      s1 = f4_dot(a1, a1, s1);
      s2 = f4_dot(a2, a2, s2);
      s3 = f4_dot(a3, a3, s3);
    }
    s = (s + s1) + (s2 + s3);
    if (MODE == 0) s *= grad_scale * grad_scale;
  }
  for (size_t i = i0; i < n4; i += stride) {
    float4 a;
    if (MODE == 0) {
      a = f4_scale(__ldg(reinterpret_cast<const float4*>(g) + i), grad_scale);
      if (decay && 4 * i < n_decay) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p) + i);
        a.x = fmaf(l2_beta, q.x, a.x);
        if (4 * i + 1 < n_decay) a.y = fmaf(l2_beta, q.y, a.y);
        if (4 * i + 2 < n_decay) a.z = fmaf(l2_beta, q.z, a.z);
        if (4 * i + 3 < n_decay) a.w = fmaf(l2_beta, q.w, a.w);
      }
    } else {
      a = __ldg(reinterpret_cast<const float4*>(p) + i);
    }
    s = fmaf(a.x, a.x, s);
    s = fmaf(a.y, a.y, s);
    s = fmaf(a.z, a.z, s);
    s = fmaf(a.w, a.w, s);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {          // tail (n not a multiple of 4)
    const size_t i = (n4 << 2) + threadIdx.x;
    float v;
    if (MODE == 0) {
      v = g[i] * grad_scale;
      if (decay && i < n_decay) v = fmaf(l2_beta, p[i], v);
    } else {
      v = p[i];
    }
    s = fmaf(v, v, s);
  }
  double d = warp_sum((double)s);
  __shared__ double red[8];
  __shared__ bool is_last;
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    work->partial[blockIdx.x] = t;
    __threadfence();
    is_last = atomicAdd(&work->counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  // last block: sum the per-block partials in a fixed order (thread k takes k, k+256, ...)
  __threadfence();
  double mine = 0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) mine += *((volatile double*)&work->partial[b]);
  mine = warp_sum(mine);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    if (MODE == 0) {
      const float norm = (float)sqrt(t);
      float scale = 1.0f;
      if (clip_norm > 0.f) scale = clip_norm * fminf(1.0f / norm, 1.0f / clip_norm);
      const float b1p = state[0], b2p = state[1];
      state[2] = norm;
      state[3] = (*lr) * sqrtf(1.0f - b2p) / (1.0f - b1p);
      state[4] = scale;
      state[0] = b1p * beta1;
      state[1] = b2p * beta2;
    } else {
      state[0] = (float)(0.5 * t);
    }
    work->counter = 0;
  }
}

__global__ void __launch_bounds__(256) adam_update_kernel(float* __restrict__ p,
                                                          const float* __restrict__ g,
                                                          float* __restrict__ m,
                                                          float* __restrict__ v, size_t n,
                                                          size_t n_decay, float l2_beta,
                                                          float grad_scale,
                                                          const double* __restrict__ denom,
                                                          float beta1,
                                                          float beta2, float eps,
                                                          const float* __restrict__ state) {
  pdl_trigger();
  pdl_wait();
  if (denom) grad_scale /= (float)(*denom);
  const float lr_t = state[3];
  const float scale = state[4];
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  const size_t n4 = n >> 2;
  const bool decay = l2_beta != 0.f;
  auto update = [&](size_t i, float4& pp, const float4& gg, float4& mm, float4& vv) {
    float pa[4] = {pp.x, pp.y, pp.z, pp.w};
    const float ga[4] = {gg.x, gg.y, gg.z, gg.w};
    float ma[4] = {mm.x, mm.y, mm.z, mm.w};
    float va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = ga[k] * grad_scale;
      if (decay && 4 * i + k < n_decay) gk = fmaf(l2_beta, pa[k], gk);
      gk *= scale;
      ma[k] = ma[k] + (gk - ma[k]) * omb1;
      va[k] = va[k] + (gk * gk - va[k]) * omb2;
      pa[k] = pa[k] - (ma[k] * lr_t) / (sqrtf(va[k]) + eps);
    }
    pp = make_float4(pa[0], pa[1], pa[2], pa[3]);
    mm = make_float4(ma[0], ma[1], ma[2], ma[3]);
    vv = make_float4(va[0], va[1], va[2], va[3]);
  };
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  // two elements per trip: all eight 16-byte loads are issued before the first store
  for (; i + stride < n4; i += 2 * stride) {
    const size_t j = i + stride;
    float4 pa = p4[i], pb = p4[j];
    const float4 ga = __ldg(g4 + i), gb = __ldg(g4 + j);
    float4 ma = m4[i], mb = m4[j];
    float4 va = v4[i], vb = v4[j];
    update(i, pa, ga, ma, va);
    update(j, pb, gb, mb, vb);
    p4[i] = pa; m4[i] = ma; v4[i] = va;
    p4[j] = pb; m4[j] = mb; v4[j] = vb;
  }
  if (i < n4) {
    float4 pa = p4[i];
    const float4 ga = __ldg(g4 + i);
    float4 ma = m4[i], va = v4[i];
    update(i, pa, ga, ma, va);
    p4[i] = pa; m4[i] = ma; v4[i] = va;
  }
  // tail (n not a multiple of 4)
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const size_t i = (n4 << 2) + threadIdx.x;
    float gk = g[i] * grad_scale;
    if (i < n_decay && l2_beta != 0.f) gk = fmaf(l2_beta, p[i], gk);
    gk *= scale;
    const float mk = m[i] + (gk - m[i]) * omb1;
    const float vk = v[i] + (gk * gk - v[i]) * omb2;
    m[i] = mk;
    v[i] = vk;
    p[i] = p[i] - (mk * lr_t) / (sqrtf(vk) + eps);
  }
}

}  // namespace dincae

using namespace dincae;

extern "C" size_t dincae_adam_workspace(size_t n) {
  return align_up(16 + (size_t)red_blocks(n) * sizeof(double), 256);
}

extern "C" int dincae_adam_step(float* params, const float* grads, float* m, float* v, size_t n,
                                size_t n_decay, float l2_beta, float clip_norm, const float* lr,
                                float beta1, float beta2, float eps, float grad_scale,
                                const double* denom, float* state, void* workspace, size_t workspace_bytes,
                                void* stream) {
  if (!params || !grads || !m || !v || !lr || !state || !workspace || n == 0 || n_decay > n)
    return DINCAE_EINVAL;
  if (workspace_bytes < dincae_adam_workspace(n)) return DINCAE_ENOSPC;
  if ((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)m | (uintptr_t)v) & 15) != 0)
    return DINCAE_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = cuda_rc(cudaMemsetAsync(workspace, 0, 16, st));
  if (rc) return rc;
  const int blocks = red_blocks(n);
  launch_k(sumsq_kernel<0>, blocks, 256, 0, st, params, grads, n, n_decay, l2_beta, grad_scale, denom,
                                          clip_norm, lr, beta1, beta2, state,
                                          reinterpret_cast<RedWork*>(workspace));
  rc = check_launch();
  if (rc) return rc;
  size_t ub = (n / 4 + 255) / 256;
  if (ub > 8 * 148) ub = 8 * 148;
  if (ub < 1) ub = 1;
  launch_k(adam_update_kernel, (unsigned)ub, 256, 0, st, params, grads, m, v, n, n_decay, l2_beta,
                                                   grad_scale, denom, beta1, beta2, eps, state);
  return check_launch();
}

extern "C" int dincae_l2_half_sumsq(const float* params, size_t n_decay, float* out,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  if (!params || !out || !workspace || n_decay == 0) return DINCAE_EINVAL;
  if (workspace_bytes < dincae_adam_workspace(n_decay)) return DINCAE_ENOSPC;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = cuda_rc(cudaMemsetAsync(workspace, 0, 16, st));
  if (rc) return rc;
  launch_k(sumsq_kernel<1>, red_blocks(n_decay), 256, 0, st, params, nullptr, n_decay, 0, 0.f, 1.f, nullptr,
                                                       0.f, nullptr, 0.f, 0.f, out,
                                                       reinterpret_cast<RedWork*>(workspace));
  return check_launch();
}
