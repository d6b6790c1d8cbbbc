// Parameter block shared by the FFMA and the tcgen05 convolution kernels.
#pragma once
#include "common.cuh"

namespace dincae {

struct ConvParams {
  ConvSrc src;
  const float* w;        // WP [9][w_cinp][cout_pad][4]
  const float* bias;     // [cout_pad] (forward) or null
  int w_cinp, cout_pad;
  int B, H, W;
  int out_planes;        // logical output planes to compute
  float slope;
  // forward outputs
  float4* out_full;
  float4* out_pool;
  uint16_t* out_sign;
  // data-gradient outputs
  int c_lo, c_hi;
  const float4* prev_act;
  float prev_slope;
  float4* g_prev;
  float4* dx_hi;
  const float4* dx_add;
  int round_out;         // round stored activations / gradients to TF32 (tensor-core back end)
};


// Tensor-core path (conv_tc.cu).  mode 0 forward, 1 data gradient.  Returns 0 on success,
// a negative DINCAE_E* code, or 1 when the shape does not fit and the caller should use
// the FFMA kernel.
int launch_conv_tc(int mode, const ConvParams& P, cudaStream_t st);

// Tensor-core weight gradient (conv_wgrad_tc.cu): writes `*nblk` partials of
// (9*cinp*cout_pad*4 + cout_pad) floats into `partial`; the caller reduces them.  Same return
// convention as launch_conv_tc (1 = shape not taken).
int wgrad_tc(const ConvSrc& u, const ConvSrc& g, int B, int H, int W, int cinp, int gp, int cout_pad,
             float* partial, size_t partial_bytes, int* nblk, cudaStream_t st);

}  // namespace dincae
