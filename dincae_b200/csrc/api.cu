// Library-level entry points of the C ABI (version, errors, device query).
#include "common.cuh"
#include <stdlib.h>

namespace dincae {
int g_last_cuda_error = 0;
long long g_launch_count = 0;
static int pdl_default() {
  // opt-in (DINCAE_PDL=1): results are bit-identical to the serialised run, but with the side
  // stream in place the early launches compete with it for SMs: 0.706 vs 0.689 ms per step
  const char* e = getenv("DINCAE_PDL");
  return (e && e[0] == '1') ? 1 : 0;
}
int g_use_pdl = pdl_default();
}

extern "C" const char* dincae_version(void) { return "dincae_b200 0.1.0 (sm_100a)"; }

extern "C" const char* dincae_error_string(int code) {
  switch (code) {
    case DINCAE_OK: return "ok";
    case DINCAE_EINVAL: return "invalid argument";
    case DINCAE_ECUDA: return "CUDA error";
    case DINCAE_ENOSPC: return "workspace too small";
    case DINCAE_EARCH: return "device is not sm_100";
    default: return "unknown error";
  }
}

extern "C" int dincae_last_cuda_error(void) { return dincae::g_last_cuda_error; }

extern "C" long long dincae_launch_count(int reset) {
  const long long n = dincae::g_launch_count;
  if (reset) dincae::g_launch_count = 0;
  return n;
}

extern "C" int dincae_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return dincae::cuda_rc(e);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return dincae::cuda_rc(e);
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return DINCAE_OK;
}
