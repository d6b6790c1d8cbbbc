// Dropout for the dense bottleneck (tf.layers.dropout, DINCAE/__init__.py:483).  In the reference
// the layer is INERT (called without training=True: SURVEY.md Q2), so the default of this library is
// no dropout; these kernels implement what the call was meant to do, as an option:
//   forward : y <- y * keep(b, n) / (1 - rate), keep drawn from Philox4x32-10 keyed by
//             (seed, step, GLOBAL batch row, column) -- independent of batch slot and GPU count;
//             optionally the keep mask is written out (one byte per element) for tests / callers.
//   backward: dropout follows a ReLU, so the stored activation is > 0 exactly where the unit was
//             active AND kept; the ReLU gate the data-gradient kernels already apply therefore also
//             applies the mask, and only the 1 / (1 - rate) factor is left: dincae_scale_rows.
#include "common.cuh"
#include "philox.cuh"

namespace dincae {

// one thread per 4 consecutive columns (one Philox call = 4 draws)
__global__ void __launch_bounds__(256) dropout_fwd_kernel(float* __restrict__ y, int B, int N, int ld, float rate,
                                                          uint64_t seed, uint64_t step, int row0,
                                                          const int64_t* __restrict__ row_keys,
                                                          uint8_t* __restrict__ mask) {
  pdl_trigger();
  pdl_wait();
  const int nq = (N + 3) >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = i / nq, q = i - b * nq;
  const bool on = b < B;
  unsigned keep = 0;
  if (on) {
    // row key: the caller's per-frame sequence number (device array) or first_global_row + b
    const uint64_t rk = row_keys ? (uint64_t)row_keys[b] : (uint64_t)(row0 + b);
    const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)rk, (uint32_t)(rk >> 32), (uint32_t)step),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
    const float inv = 1.0f / (1.0f - rate);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = 4 * q + j;
      if (n < N) {
        const float u = ((float)(rr[j] >> 8) + 0.5f) * (1.0f / 16777216.0f);      // (0, 1)
        const bool k = u >= rate;
        keep |= (k ? 1u : 0u) << j;
        float* p = y + (size_t)b * ld + n;
        *p = k ? *p * inv : 0.f;
        if (mask) mask[(size_t)b * N + n] = k ? 1 : 0;
      }
    }
  }
  (void)keep;
}

__global__ void __launch_bounds__(256) scale_rows_kernel(float* __restrict__ x, int B, int N, int ld, float f) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * N) return;
  const int b = (int)(i / N), n = (int)(i - (size_t)b * N);
  x[(size_t)b * ld + n] *= f;
}

}  // namespace dincae

using namespace dincae;

extern "C" int dincae_dropout_fwd(float* y, int B, int N, int ld, float rate, uint64_t seed, uint64_t step,
                                  int first_global_row, const int64_t* row_keys, uint8_t* mask, void* stream) {
  if (!y || B <= 0 || N <= 0 || ld < N || rate < 0.f || rate >= 1.f) return DINCAE_EINVAL;
  const int nq = (N + 3) / 4;
  const long long total = (long long)B * nq;
  launch_k(dropout_fwd_kernel, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream, y, B, N, ld, rate, seed,
           step, first_global_row, row_keys, mask);
  return check_launch();
}

extern "C" int dincae_scale_rows(float* x, int B, int N, int ld, float factor, void* stream) {
  if (!x || B <= 0 || N <= 0 || ld < N) return DINCAE_EINVAL;
  const size_t total = (size_t)B * N;
  launch_k(scale_rows_kernel, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream, x, B, N, ld, factor);
  return check_launch();
}
