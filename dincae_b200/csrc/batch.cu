// Input pipeline on the device: temporal-mean removal and the per-batch channel stacking,
// added-cloud masking and jitter of DINCAE/__init__.py:155-227.  HBM-bound gather kernels.
#include "common.cuh"
#include "bf16.cuh"
#include "philox.cuh"

namespace dincae {

// ---------------------------------------------------------------------------------------
// prepare_cube: numpy's MaskedArray.mean over axis 0 adds the 0-filled float32 slices in
// time order, then divides in float64; the anomaly and the division by obs_var are float64
// too and are rounded to float32 once when stored (DINCAE/__init__.py:168-169,180).
// Two kernels.  (1) cube_mean_kernel: the float32 sum of a pixel must run in time order, so a
// thread owns a pixel (coalesced across pixels) -- with only H*W threads the pass is bound by
// load latency, hence PC_UNROLL time steps are fetched into registers (all loads independent,
// in flight together) before the ordered adds.  (2) cube_anom_kernel: one thread per 4
// consecutive elements of the T*H*W cube, fully parallel.
// ---------------------------------------------------------------------------------------
constexpr int PC_UNROLL = 32;

__global__ void __launch_bounds__(64) cube_mean_kernel(const float* __restrict__ data,
                                                       const uint8_t* __restrict__ missing, int T, size_t HW,
                                                       int nomask, double* __restrict__ meandata,
                                                       uint8_t* __restrict__ never) {
  pdl_trigger();
  pdl_wait();
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const float* d = data + p;
  const uint8_t* m = missing + p;
  float sum = 0.f;
  int count = 0;
  int t = 0;
  for (; t + PC_UNROLL <= T; t += PC_UNROLL) {
    float v[PC_UNROLL];
    uint8_t k[PC_UNROLL];
#pragma unroll
    for (int u = 0; u < PC_UNROLL; ++u) {
      v[u] = __ldg(d + (size_t)(t + u) * HW);
      k[u] = __ldg(m + (size_t)(t + u) * HW);
    }
#pragma unroll
    for (int u = 0; u < PC_UNROLL; ++u) {
      sum = __fadd_rn(sum, k[u] ? 0.f : v[u]);
      count += k[u] ? 0 : 1;
    }
  }
  for (; t < T; ++t) {
    const bool miss = m[(size_t)t * HW] != 0;
    sum = __fadd_rn(sum, miss ? 0.f : d[(size_t)t * HW]);
    count += miss ? 0 : 1;
  }
  double mean;
  if (nomask) {
    mean = (double)__fdiv_rn(sum, (float)T);          // ndarray.mean of float32 stays float32
  } else {
    mean = count > 0 ? __ddiv_rn((double)sum * 1.0, (double)count) : 0.0;
  }
  meandata[p] = mean;
  never[p] = (count == 0) ? 1 : 0;
}

__device__ __forceinline__ float cube_anom1(float v, bool miss, double mean, bool nev, double obs_var) {
  if (miss || nev) return 0.f;
  return (float)__ddiv_rn(__dsub_rn((double)v, mean), obs_var);
}

// VEC = 4: H*W is a multiple of 4 and the buffers are 16-byte aligned (a quad never crosses a frame)
template <int VEC>
__global__ void __launch_bounds__(256) cube_anom_kernel(const float* __restrict__ data,
                                                        const uint8_t* __restrict__ missing, size_t n_items,
                                                        size_t HW, double obs_var,
                                                        const double* __restrict__ meandata,
                                                        const uint8_t* __restrict__ never,
                                                        float* __restrict__ anom) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  if (VEC == 4) {
    const size_t e = 4 * i;
    const size_t p = e % HW;
    const float4 v = __ldg(reinterpret_cast<const float4*>(data) + i);
    const uchar4 k = __ldg(reinterpret_cast<const uchar4*>(missing) + i);
    const double2 m01 = __ldg(reinterpret_cast<const double2*>(meandata + p));
    const double2 m23 = __ldg(reinterpret_cast<const double2*>(meandata + p) + 1);
    const uchar4 nv = __ldg(reinterpret_cast<const uchar4*>(never + p));
    float4 o;
    o.x = cube_anom1(v.x, k.x != 0, m01.x, nv.x != 0, obs_var);
    o.y = cube_anom1(v.y, k.y != 0, m01.y, nv.y != 0, obs_var);
    o.z = cube_anom1(v.z, k.z != 0, m23.x, nv.z != 0, obs_var);
    o.w = cube_anom1(v.w, k.w != 0, m23.y, nv.w != 0, obs_var);
    reinterpret_cast<float4*>(anom)[i] = o;
  } else {
    const size_t p = i % HW;
    anom[i] = cube_anom1(data[i], missing[i] != 0, meandata[p], never[p] != 0, obs_var);
  }
}

// Box-Muller on the MUFU units (lg2 / sqrt / sin / cos approximations, absolute error ~1e-6):
// the values are noise, and the IEEE logf / sqrtf / sincospif made the batch builder issue bound
// (~230 instructions per variable and pixel: 0.52 ms per 16-frame batch at configs[3]).
__device__ __forceinline__ float bm_radius(uint32_t a) {
  const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;   // (0,1]
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * __logf(u1)));
  return r;
}
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float r = bm_radius(a);
  const float th = ((float)b + 0.5f) * (6.283185307179586f * 2.3283064365386963e-10f);
  float s, c;
  __sincosf(th, &s, &c);
  return make_float2(r * c, r * s);
}
__device__ __forceinline__ float box_muller_cos(uint32_t a, uint32_t b) {
  const float th = ((float)b + 0.5f) * (6.283185307179586f * 2.3283064365386963e-10f);
  return bm_radius(a) * __cosf(th);
}

#define DINCAE_MAX_NDATA 8

struct BatchParams {
  const float* anom;
  const uint8_t* missing;
  const uint8_t* cloud_missing;
  const float* lon_scaled;
  const float* lat_scaled;
  const float* doy_cos;
  const float* doy_sin;
  const int32_t* frame_idx;
  const int32_t* cloud_idx;
  const double* noise;
  const int64_t* noise_seq;
  float* xin;              // P4 fp32 output (may be null when xin8 is set)
  uint4* xin8;             // P8 bf16 output [B][planes8][H][W][8] (bf16 network), or null
  float* xtrue;
  int32_t* n_noncloud;
  uint64_t seed;
  int ndata, T, H, W, ntime_win, B, nvar_planes, planes8;
  unsigned jitter_mask;    // bit j: jitter[j] != 0
  float jitter_f[DINCAE_MAX_NDATA];
  int round_tf32;          // network inputs rounded to TF32 (tensor-core conv back end)
  float inv_var[DINCAE_MAX_NDATA];
  double jitter[DINCAE_MAX_NDATA];
};

// One thread per (b, y, x).  All nvar channels of the pixel are produced plane by plane
// (float4 stores, coalesced along x).
// ND / WIN > 0: number of variables and time window known at compile time (<1,3> = the
// default SST configuration): the channel decode folds to constants, the plane loop unrolls and
// all loads of a pixel (missing / anomaly of the 3 frames, added-cloud mask) are independent
// and in flight together.  <0,0>: any configuration, run-time decode.
template <int ND, int WIN>
__global__ void __launch_bounds__(256) build_batch_kernel(const BatchParams P) {
  pdl_trigger();
  pdl_wait();
  const int HW = P.H * P.W;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = gid < (long long)P.B * HW;
  int cnt = 0;
  if (active) {
    const int b = (int)(gid / HW);
    const int p = (int)(gid - (long long)b * HW);
    const int y = p / P.W, x = p - y * P.W;
    const int i = P.frame_idx[b];
    const int nd = ND ? ND : P.ndata, nd2 = 2 * nd;
    const int win = WIN ? WIN : P.ntime_win;
    constexpr int NPL = (2 * ND * WIN + 4 + 3) / 4;
    const int n_planes = ND ? NPL : P.nvar_planes;
    const size_t THW = (size_t)P.T * HW;
    const bool train = P.cloud_idx != nullptr;
    const int ic = train ? P.cloud_idx[b] : 0;
    const int nvar = nd2 * win + 4;
    const int half = win / 2;

    // jitter slots per variable j: channels 2j, 2j+2nd+4, 2j+4nd+4 (reference quirk Q4)
    uint4 rnd = make_uint4(0, 0, 0, 0);
    float gauss[3] = {0.f, 0.f, 0.f};
    int gauss_j = -1;

    float lane[4];
    float pair[8];
#pragma unroll(ND ? NPL : 1)
    for (int plane = 0; plane < n_planes; ++plane) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = plane * 4 + k;
        float v = 0.f;
        if (c < nvar) {
          int j = -1, kind = 0, slot = -1, frame = i;
          if (c < nd2) {
            j = c >> 1; kind = c & 1; slot = 0;
          } else if (c < nd2 + 4) {
            const int s = c - nd2;
            v = s == 0 ? P.lon_scaled[x] : (s == 1 ? P.lat_scaled[y]
                                                   : (s == 2 ? P.doy_cos[i] : P.doy_sin[i]));
          } else {
            const int r = c - nd2 - 4;
            const int w = r / nd2;                 // neighbour slot 0 .. ntime_win-2
            const int rr = r - w * nd2;
            j = rr >> 1; kind = rr & 1; slot = w + 1;
            int nn = w - half;                     // offsets in order, centre skipped
            if (nn >= 0) nn += 1;
            frame = min(P.T - 1, max(0, i + nn));
          }
          if (j >= 0) {
            const size_t at = (size_t)j * THW + (size_t)frame * HW + p;
            const bool miss = P.missing[at] != 0;
            bool clouded = false;
            if (train && slot == 0)
              clouded = P.cloud_missing[(size_t)j * THW + (size_t)ic * HW + p] != 0;
            if (kind == 0) {
              v = (miss || clouded) ? 0.f : P.anom[at];
              if (train && slot <= 2) {
                if (P.noise != nullptr) {
                  const double z = P.noise[((size_t)b * 3 * nd + 3 * j + slot) * HW + p];
                  v = (float)((double)v + P.jitter[j] * z);
                } else if ((P.jitter_mask >> j) & 1u) {
                  if (gauss_j != j) {
                    const uint64_t seq = P.noise_seq ? (uint64_t)P.noise_seq[b] : (uint64_t)i;
                    rnd = philox4x32_10(
                        make_uint4((uint32_t)p, (uint32_t)j, (uint32_t)seq, (uint32_t)(seq >> 32)),
                        make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32)));
                    const float2 g0 = box_muller(rnd.x, rnd.y);
                    gauss[0] = g0.x; gauss[1] = g0.y; gauss[2] = box_muller_cos(rnd.z, rnd.w);
                    gauss_j = j;
                  }
                  v = fmaf(P.jitter_f[j], gauss[slot], v);
                }
              }
            } else {
              v = (miss || clouded) ? 0.f : P.inv_var[j];
            }
          }
        }
        lane[k] = v;
      }
      float4 out = make_float4(lane[0], lane[1], lane[2], lane[3]);
      if (P.round_tf32) out = to_tf32(out);
      if (P.xin) reinterpret_cast<float4*>(P.xin)[((size_t)(b * P.nvar_planes + plane) * P.H + y) * P.W + x] = out;
      if (P.xin8) {                                      // two P4 planes -> one P8 plane (same bits as p4_to_p8)
        const int h = plane & 1;
        pair[4 * h] = out.x; pair[4 * h + 1] = out.y; pair[4 * h + 2] = out.z; pair[4 * h + 3] = out.w;
        if (h == 1 || plane == n_planes - 1) {
          if (h == 0) { pair[4] = 0.f; pair[5] = 0.f; pair[6] = 0.f; pair[7] = 0.f; }
          P.xin8[((size_t)(b * P.planes8 + (plane >> 1)) * P.H + y) * P.W + x] = bf8_pack(pair);
        }
      }
    }
    if (P.xin8)                                          // planes beyond the stacked channels stay zero
      for (int q = (n_planes + 1) >> 1; q < P.planes8; ++q)
        P.xin8[((size_t)(b * P.planes8 + q) * P.H + y) * P.W + x] = make_uint4(0u, 0u, 0u, 0u);
    // target: un-perturbed channels 0 and 1 of the current frame (:227)
    {
      const size_t at = (size_t)i * HW + p;
      const bool miss = P.missing[at] != 0;
      const float t0 = miss ? 0.f : P.anom[at];
      const float t1 = miss ? 0.f : P.inv_var[0];
      reinterpret_cast<float2*>(P.xtrue)[(size_t)b * HW + p] = make_float2(t0, t1);
      cnt = (t1 != 0.f) ? 1 : 0;
    }
  }
  // count of observed target pixels: integer, hence order-independent
  unsigned m = __ballot_sync(0xffffffffu, cnt != 0);
  __shared__ int block_cnt;
  if (threadIdx.x == 0) block_cnt = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&block_cnt, __popc(m));
  __syncthreads();
  if (threadIdx.x == 0 && block_cnt) atomicAdd(P.n_noncloud, block_cnt);
}

}  // namespace dincae

using namespace dincae;

extern "C" int dincae_prepare_cube(const float* data, const uint8_t* missing, int T, int H, int W,
                                   double obs_var, int nomask, float* anom, double* meandata,
                                   uint8_t* never, void* stream) {
  if (!data || !missing || !anom || !meandata || !never || T <= 0 || H <= 0 || W <= 0)
    return DINCAE_EINVAL;
  const size_t HW = (size_t)H * W;
  const cudaStream_t st = (cudaStream_t)stream;
  launch_k(cube_mean_kernel, (unsigned)((HW + 63) / 64), 64, 0, st, data, missing, T, HW, nomask, meandata, never);
  const size_t n = HW * (size_t)T;
  const bool vec = (HW % 4 == 0) && (((uintptr_t)data | (uintptr_t)missing | (uintptr_t)anom |
                                      (uintptr_t)never) % 16 == 0) && ((uintptr_t)meandata % 16 == 0);
  const size_t items = vec ? n / 4 : n;
  if ((items + 255) / 256 > 0x7fffffffull) return DINCAE_EINVAL;
  if (vec)
    launch_k(cube_anom_kernel<4>, (unsigned)((items + 255) / 256), 256, 0, st, data, missing, items, HW, obs_var,
             (const double*)meandata, (const uint8_t*)never, anom);
  else
    launch_k(cube_anom_kernel<1>, (unsigned)((items + 255) / 256), 256, 0, st, data, missing, items, HW, obs_var,
             (const double*)meandata, (const uint8_t*)never, anom);
  return check_launch();
}

static int build_batch_impl(const float* anom, const uint8_t* missing,
                                  const uint8_t* cloud_missing,
                                  const float* lon_scaled, const float* lat_scaled,
                                  const float* doy_cos, const float* doy_sin,
                                  const double* inv_var_host, const double* jitter_std_host,
                                  int ndata, int T, int H, int W, int ntime_win,
                                  const int32_t* frame_idx, const int32_t* cloud_idx, int B,
                                  const double* noise, uint64_t seed, const int64_t* noise_seq,
                                  float* xin, int nvar_planes, void* xin8, int planes8, float* xtrue,
                                  int32_t* n_noncloud, void* stream) {
  if (!anom || !missing || !lon_scaled || !lat_scaled || !doy_cos || !doy_sin || !inv_var_host ||
      !frame_idx || (!xin && !xin8) || !xtrue || !n_noncloud)
    return DINCAE_EINVAL;
  if (xin8 && 2 * planes8 < nvar_planes) return DINCAE_EINVAL;
  if (ndata <= 0 || ndata > DINCAE_MAX_NDATA || T <= 0 || H <= 0 || W <= 0 || B <= 0 ||
      ntime_win < 1 || (ntime_win & 1) == 0)
    return DINCAE_EINVAL;
  const int nvar = 2 * ndata * ntime_win + 4;
  if (nvar_planes * 4 < nvar) return DINCAE_EINVAL;
  if (cloud_idx && ntime_win < 3) return DINCAE_EINVAL;   // reference raises IndexError (Q4)
  BatchParams P;
  P.anom = anom; P.missing = missing; P.cloud_missing = cloud_missing ? cloud_missing : missing; P.lon_scaled = lon_scaled; P.lat_scaled = lat_scaled;
  P.doy_cos = doy_cos; P.doy_sin = doy_sin; P.frame_idx = frame_idx; P.cloud_idx = cloud_idx;
  P.noise = noise; P.noise_seq = noise_seq; P.xin = xin; P.xtrue = xtrue;
  P.xin8 = reinterpret_cast<uint4*>(xin8); P.planes8 = planes8; P.jitter_mask = 0;
  P.n_noncloud = n_noncloud; P.seed = seed;
  P.ndata = ndata; P.T = T; P.H = H; P.W = W; P.ntime_win = ntime_win; P.B = B;
  P.nvar_planes = nvar_planes;
  P.round_tf32 = g_conv_backend == 1;
  for (int j = 0; j < DINCAE_MAX_NDATA; ++j) {
    P.inv_var[j] = j < ndata ? (float)inv_var_host[j] : 0.f;
    P.jitter[j] = (j < ndata && jitter_std_host) ? jitter_std_host[j] : 0.0;
    P.jitter_f[j] = (float)P.jitter[j];
    if (P.jitter[j] != 0.0) P.jitter_mask |= 1u << j;
  }
  const long long total = (long long)B * H * W;
  const int threads = 256;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  if (ndata == 1 && ntime_win == 3 && nvar_planes == 3)
    launch_k(build_batch_kernel<1, 3>, blocks, threads, 0, (cudaStream_t)stream, P);
  else if (ndata == 4 && ntime_win == 3 && nvar_planes == 7)       // BASELINE configs[3]: SST, chl, u, v
    launch_k(build_batch_kernel<4, 3>, blocks, threads, 0, (cudaStream_t)stream, P);
  else if (ndata == 3 && ntime_win == 3 && nvar_planes == 6)
    launch_k(build_batch_kernel<3, 3>, blocks, threads, 0, (cudaStream_t)stream, P);
  else if (ndata == 2 && ntime_win == 3 && nvar_planes == 4)
    launch_k(build_batch_kernel<2, 3>, blocks, threads, 0, (cudaStream_t)stream, P);
  else
    launch_k(build_batch_kernel<0, 0>, blocks, threads, 0, (cudaStream_t)stream, P);
  return check_launch();
}

extern "C" int dincae_build_batch(const float* anom, const uint8_t* missing,
                                  const uint8_t* cloud_missing,
                                  const float* lon_scaled, const float* lat_scaled,
                                  const float* doy_cos, const float* doy_sin,
                                  const double* inv_var_host, const double* jitter_std_host,
                                  int ndata, int T, int H, int W, int ntime_win,
                                  const int32_t* frame_idx, const int32_t* cloud_idx, int B,
                                  const double* noise, uint64_t seed, const int64_t* noise_seq,
                                  float* xin, int nvar_planes, float* xtrue,
                                  int32_t* n_noncloud, void* stream) {
  if (!xin) return DINCAE_EINVAL;
  return build_batch_impl(anom, missing, cloud_missing, lon_scaled, lat_scaled, doy_cos, doy_sin, inv_var_host,
                          jitter_std_host, ndata, T, H, W, ntime_win, frame_idx, cloud_idx, B, noise, seed,
                          noise_seq, xin, nvar_planes, nullptr, 0, xtrue, n_noncloud, stream);
}

// The same batch written as bf16 P8 planes [B][planes8][H][W][8] for the bf16 network (bit-identical
// to dincae_build_batch followed by dincae_p4_to_p8; `xin` may be null, then only the bf16 copy is
// produced -- saves the fp32 write and the conversion pass: 0.7 GB of traffic per 16 frames of
// configs[3]).
extern "C" int dincae_build_batch_p8(const float* anom, const uint8_t* missing,
                                     const uint8_t* cloud_missing,
                                     const float* lon_scaled, const float* lat_scaled,
                                     const float* doy_cos, const float* doy_sin,
                                     const double* inv_var_host, const double* jitter_std_host,
                                     int ndata, int T, int H, int W, int ntime_win,
                                     const int32_t* frame_idx, const int32_t* cloud_idx, int B,
                                     const double* noise, uint64_t seed, const int64_t* noise_seq,
                                     float* xin, int nvar_planes, void* xin8, int planes8, float* xtrue,
                                     int32_t* n_noncloud, void* stream) {
  if (!xin8 || planes8 <= 0) return DINCAE_EINVAL;
  return build_batch_impl(anom, missing, cloud_missing, lon_scaled, lat_scaled, doy_cos, doy_sin, inv_var_host,
                          jitter_std_host, ndata, T, H, W, ntime_win, frame_idx, cloud_idx, B, noise, seed,
                          noise_seq, xin, nvar_planes, xin8, planes8, xtrue, n_noncloud, stream);
}
