// Ensemble averaging of saved reconstructions (README.md:90 of the reference recommends
// averaging the data-*.nc files written every `save_each` epochs; the reference has no code
// for it).  Two HBM-bound element-wise kernels: accumulate one member into double-precision
// running sums (skipping fill values), then finalise mean and combined sigma.
#include "common.cuh"

namespace dincae {

__global__ void __launch_bounds__(256) ensemble_accumulate_kernel(
    const float* __restrict__ mean_rec, const float* __restrict__ sigma_rec, float fill, size_t n,
    double* __restrict__ acc, int32_t* __restrict__ count) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float m = mean_rec[i], s = sigma_rec[i];
    if (m == fill || s == fill) continue;            // masked in this member
    acc[i] += (double)m;
    acc[n + i] += (double)s * (double)s;
    acc[2 * n + i] += (double)m * (double)m;
    count[i] += 1;
  }
}

// mixture == 0: sigma = sqrt(mean of the members' variances)
// mixture != 0: sigma = sqrt(mean variance + variance of the members' means)  (law of total
//               variance for an equally weighted mixture)
__global__ void __launch_bounds__(256) ensemble_finalize_kernel(
    const double* __restrict__ acc, const int32_t* __restrict__ count, size_t n, int mixture,
    float fill, float* __restrict__ mean_out, float* __restrict__ sigma_out) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = count[i];
    if (c == 0) {
      mean_out[i] = fill;
      sigma_out[i] = fill;
      continue;
    }
    const double inv = 1.0 / (double)c;
    const double mu = acc[i] * inv;
    double var = acc[n + i] * inv;
    if (mixture) {
      const double spread = acc[2 * n + i] * inv - mu * mu;
      var += spread > 0.0 ? spread : 0.0;
    }
    mean_out[i] = (float)mu;
    sigma_out[i] = (float)sqrt(var);
  }
}

static unsigned ens_blocks(size_t n) {
  size_t b = (n + 255) / 256;
  if (b > 8 * 148) b = 8 * 148;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace dincae

using namespace dincae;

extern "C" int dincae_ensemble_accumulate(const float* mean_rec, const float* sigma_rec, float fill,
                                          size_t n, double* acc, int32_t* count, void* stream) {
  if (!mean_rec || !sigma_rec || !acc || !count || n == 0) return DINCAE_EINVAL;
  launch_k(ensemble_accumulate_kernel, ens_blocks(n), 256, 0, (cudaStream_t)stream, mean_rec, sigma_rec,
           fill, n, acc, count);
  return check_launch();
}

extern "C" int dincae_ensemble_finalize(const double* acc, const int32_t* count, size_t n, int mixture,
                                        float fill, float* mean_out, float* sigma_out, void* stream) {
  if (!acc || !count || !mean_out || !sigma_out || n == 0) return DINCAE_EINVAL;
  launch_k(ensemble_finalize_kernel, ens_blocks(n), 256, 0, (cudaStream_t)stream, acc, count, n, mixture,
           fill, mean_out, sigma_out);
  return check_launch();
}
