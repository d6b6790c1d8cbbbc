// Output head and masked Gaussian negative log-likelihood with its backward
// (DINCAE/__init__.py:520-570): one pass over xrec / xtrue, warp-shuffle + shared-memory
// reduction, fixed-order final sum by the last block (deterministic).
#include "bf16.cuh"

namespace dincae {

__device__ __forceinline__ void head_eval(float a, float lb, float& m_rec, float& s2, float& q,
                                          float& dq) {
  const float t = fminf(lb, 10.0f);
  const float p = expf(t);
  q = fmaxf(p, 1e-3f);
  s2 = 1.0f / q;
  m_rec = a * s2;
  // d q / d lb: tf.minimum / tf.maximum pass the gradient on ties
  dq = (lb <= 10.0f && p >= 1e-3f) ? p : 0.0f;
}

__global__ void __launch_bounds__(256) head_recon_kernel(const float4* __restrict__ xrec, size_t n,
                                                         float* __restrict__ m_rec,
                                                         float* __restrict__ s2_rec) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 r = __ldg(&xrec[i]);
  float m, s2, q, dq;
  head_eval(r.x, r.y, m, s2, q, dq);
  m_rec[i] = m;
  s2_rec[i] = s2;
}

// Head + what `savesample` does per batch (DINCAE/__init__.py:237-248, 288-291), written as
// the records of the output file: record b = [mean_rec[b] | sigma_rec[b]], each H*W float32,
// mean_rec = float32(float64(m_rec) + meandata), sigma_rec = sqrt(sigma2_rec), both replaced
// by the fill value where the pixel was never observed (meandata.mask).  With big_endian the
// 4 bytes of every value are reversed, so that the buffer is the byte image of NetCDF-3
// records and the host only has to write it out.
__global__ void __launch_bounds__(256) head_recon_records_kernel(
    const float4* __restrict__ xrec, int B, int HW, const double* __restrict__ meandata,
    const uint8_t* __restrict__ never, float fill, int mean_f32, int big_endian,
    uint32_t* __restrict__ records) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * HW) return;
  const int b = (int)(i / HW), p = (int)(i - (size_t)b * HW);
  const float4 r = __ldg(&xrec[i]);
  float m, s2, q, dq;
  head_eval(r.x, r.y, m, s2, q, dq);
  // numpy adds in the operands' common type: float64 for the usual (masked -> float64) mean,
  // float32 when the input had no mask at all and the mean stayed float32
  const double md = __ldg(&meandata[p]);
  float mean = mean_f32 ? __fadd_rn(m, (float)md) : (float)__dadd_rn((double)m, md);
  float sigma = __fsqrt_rn(s2);
  if (never && __ldg(&never[p])) mean = sigma = fill;
  uint32_t um = __float_as_uint(mean), us = __float_as_uint(sigma);
  if (big_endian) {
    um = __byte_perm(um, 0u, 0x0123);
    us = __byte_perm(us, 0u, 0x0123);
  }
  uint32_t* o = records + (size_t)b * 2 * HW + p;
  o[0] = um;
  o[HW] = us;
}

struct LossWork {
  unsigned int counter;
  unsigned int pad[3];
  double partial[1];   // [gridDim.x][4]
};

__global__ void __launch_bounds__(256) head_loss_kernel(
    const float4* __restrict__ xrec, const float2* __restrict__ xtrue, size_t n,
    int truth_uncertain, float neg_slope, const int32_t* __restrict__ n_noncloud,
    float4* __restrict__ g_out, double* __restrict__ sums, float* __restrict__ loss_out,
    LossWork* __restrict__ work, int gfmt) {
  pdl_trigger();
  pdl_wait();
  const bool normalise = n_noncloud != nullptr;
  const float nn = normalise ? (float)(*n_noncloud) : 1.0f;
  const float inv_n = 1.0f / nn;
  float s_log = 0.f, s_sq = 0.f, s_d2 = 0.f, s_w = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float4 r = __ldg(&xrec[i]);
    const float2 t = __ldg(&xtrue[i]);
    float m, s2, q, dq;
    head_eval(r.x, r.y, m, s2, q, dq);
    const float s2t = 1.0f / fmaxf(t.y, 1e-3f);
    const float mt = t.x * s2t;
    const float d = m - mt;
    const bool w = t.y != 0.f;
    float ga = 0.f, gb = 0.f;
    if (w) {
      const float d2 = d * d;
      if (truth_uncertain) {
        s_log += logf(s2 / s2t) + (s2t + d2) / s2;
        gb = dq * (-1.0f / q + (s2t + d2) - 2.0f * d * r.x / q);
      } else {
        s_log += logf(s2);
        s_sq += d2 / s2;
        gb = dq * (-1.0f / q + d2 - 2.0f * d * r.x / q);
      }
      ga = 2.0f * d;
      s_d2 += d2;
      s_w += 1.0f;
    }
    ga *= inv_n * (r.x > 0.f ? 1.0f : neg_slope);
    gb *= inv_n * (r.y > 0.f ? 1.0f : neg_slope);
    // gfmt 0: fp32, 1: fp32 rounded to TF32, 2: planar-8 bf16 (8 channels in the same 16 bytes)
    if (gfmt == 2)
      reinterpret_cast<uint4*>(g_out)[i] = make_uint4(bf2_pack(ga, gb), 0u, 0u, 0u);
    else
      g_out[i] = gfmt == 1 ? make_float4(to_tf32(ga), to_tf32(gb), 0.f, 0.f) : make_float4(ga, gb, 0.f, 0.f);
  }
  __shared__ double red[8][4];
  double v[4] = {(double)s_log, (double)s_sq, (double)s_d2, (double)s_w};
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = warp_sum(v[k]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int k = 0; k < 4; ++k) red[warp][k] = v[k];
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    for (int k = 0; k < 4; ++k) {
      double s = 0;
      for (int w8 = 0; w8 < 8; ++w8) s += red[w8][k];
      work->partial[(size_t)blockIdx.x * 4 + k] = s;
    }
    __threadfence();
    const unsigned done = atomicAdd(&work->counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    // per-block partials summed in a fixed order: warp k (k < 4) owns sum k, lane j takes
    // blocks j, j+32, ...
    __threadfence();
    if (warp < 4) {
      double s = 0;
      for (unsigned blk = lane; blk < gridDim.x; blk += 32)
        s += *((volatile double*)&work->partial[(size_t)blk * 4 + warp]);
      s = warp_sum(s);
      if (lane == 0) {
        sums[warp] = s;
        red[0][warp] = s;
      }
    }
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    const double dn = normalise ? (double)nn : red[0][3];
    loss_out[0] = (float)((red[0][0] + red[0][1]) / dn);
    loss_out[1] = (float)sqrt(red[0][2] / dn);
    work->counter = 0;
  }
}

static int loss_blocks(size_t n) {
  size_t b = (n + 1023) / 1024;
  if (b > 4 * 148) b = 4 * 148;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace dincae

using namespace dincae;

extern "C" int dincae_head_recon(const float* xrec, int B, int H, int W, float* m_rec,
                                 float* sigma2_rec, void* stream) {
  if (!xrec || !m_rec || !sigma2_rec || B <= 0 || H <= 0 || W <= 0) return DINCAE_EINVAL;
  const size_t n = (size_t)B * H * W;
  launch_k(head_recon_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, 
      reinterpret_cast<const float4*>(xrec), n, m_rec, sigma2_rec);
  return check_launch();
}

extern "C" int dincae_head_recon_records(const float* xrec, int B, int H, int W, const double* meandata,
                                         const uint8_t* never, float fill_value, int mean_is_f32,
                                         int big_endian, void* records, void* stream) {
  if (!xrec || !meandata || !records || B <= 0 || H <= 0 || W <= 0) return DINCAE_EINVAL;
  if ((long long)H * W >= (1ll << 30)) return DINCAE_EINVAL;
  const size_t n = (size_t)B * H * W;
  launch_k(head_recon_records_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream,
           reinterpret_cast<const float4*>(xrec), B, H * W, meandata, never, fill_value, mean_is_f32, big_endian,
           reinterpret_cast<uint32_t*>(records));
  return check_launch();
}

extern "C" size_t dincae_head_loss_workspace(int B, int H, int W) {
  const size_t n = (size_t)B * H * W;
  return align_up(16 + (size_t)loss_blocks(n) * 4 * sizeof(double), 256);
}

static int head_loss_launch(const float* xrec, const float* xtrue, int B, int H, int W,
                            int truth_uncertain, float neg_slope, const int32_t* n_noncloud,
                            float* g_out, double* sums, float* loss_out, void* workspace,
                            size_t workspace_bytes, void* stream, int gfmt) {
  if (!xrec || !xtrue || !g_out || !sums || !loss_out || !workspace || B <= 0 ||
      H <= 0 || W <= 0)
    return DINCAE_EINVAL;
  if (workspace_bytes < dincae_head_loss_workspace(B, H, W)) return DINCAE_ENOSPC;
  const size_t n = (size_t)B * H * W;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = cuda_rc(cudaMemsetAsync(workspace, 0, 16, st));
  if (rc) return rc;
  launch_k(head_loss_kernel, loss_blocks(n), 256, 0, st, 
      reinterpret_cast<const float4*>(xrec), reinterpret_cast<const float2*>(xtrue), n,
      truth_uncertain, neg_slope, n_noncloud, reinterpret_cast<float4*>(g_out), sums, loss_out,
      reinterpret_cast<LossWork*>(workspace), gfmt);
  return check_launch();
}

extern "C" int dincae_head_loss(const float* xrec, const float* xtrue, int B, int H, int W,
                                int truth_uncertain, float neg_slope, const int32_t* n_noncloud,
                                float* g_out, double* sums, float* loss_out, void* workspace,
                                size_t workspace_bytes, void* stream) {
  return head_loss_launch(xrec, xtrue, B, H, W, truth_uncertain, neg_slope, n_noncloud, g_out, sums, loss_out,
                          workspace, workspace_bytes, stream, g_conv_backend == 1 ? 1 : 0);
}

extern "C" int dincae_head_loss_bf16(const float* xrec, const float* xtrue, int B, int H, int W,
                                     int truth_uncertain, float neg_slope, const int32_t* n_noncloud,
                                     void* g_out, double* sums, float* loss_out, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  return head_loss_launch(xrec, xtrue, B, H, W, truth_uncertain, neg_slope, n_noncloud,
                          reinterpret_cast<float*>(g_out), sums, loss_out, workspace, workspace_bytes, stream, 2);
}
