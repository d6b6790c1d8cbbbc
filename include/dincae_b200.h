/*
 * dincae_b200 -- C ABI of the B200 (sm_100a) kernels behind DINCAE's train + reconstruct
 * hot path.  Every entry point replaces a block of the TensorFlow-1.15 graph (or of the
 * numpy input pipeline) that the reference builds in DINCAE/__init__.py; the citation on
 * each declaration is the reference interface it stands in for.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - All data pointers are DEVICE pointers owned by the caller unless marked "host".
 *   - Nothing here allocates, synchronises the device or keeps global state; every call
 *     only enqueues work on `stream` (a cudaStream_t passed as void*), so sequences of
 *     calls can be captured into a CUDA graph.
 *   - Return value: 0 on success, a negative DINCAE_E* code otherwise
 *     (dincae_error_string() names it; dincae_last_cuda_error() gives the CUDA code).
 *
 * Activation layout ("planar-4", P4): a tensor with C channels at spatial size H x W is
 * stored as float [B][CP][H][W][4] with CP = ceil(C/4) channel planes; channel c lives in
 * plane c/4, lane c%4; lanes past C are zero.  A concatenation of two tensors is never
 * materialised: kernels take two sources.  A "half" source is stored at
 * (ceil(H/2), ceil(W/2)) and read at (y>>1, x>>1) == TF1 legacy nearest-neighbour resize
 * for the sizes this network produces (DINCAE/__init__.py:496-499).
 *
 * Conv weight layout ("WP"): float [9][CinP][CoutPad][4]: tap t = 3*ky+kx, input plane,
 * output channel (padded to CoutPad, a multiple of 4), 4 input lanes.  TF's HWIO kernel
 * [3,3,Cin,Cout] maps to it by dincae_b200.layout.pack_conv (host, init/export only).
 * Sign masks: uint16 [B][CP][H][ceil(W/4)], bit 4*(x%4)+lane set when the conv
 * pre-activation was > 0 (what LeakyReluGrad needs, instead of the full-res activation).
 */
#ifndef DINCAE_B200_H
#define DINCAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DINCAE_OK 0
#define DINCAE_EINVAL (-1)   /* bad argument (null pointer, non-positive size, ...) */
#define DINCAE_ECUDA (-2)    /* a CUDA runtime call / launch failed */
#define DINCAE_ENOSPC (-3)   /* caller-provided workspace too small */
#define DINCAE_EARCH (-4)    /* device is not sm_100 */

const char* dincae_version(void);
const char* dincae_error_string(int code);
int dincae_last_cuda_error(void);
/* Number of kernels this library has enqueued since the last reset (diagnostics for
 * bench.py's gpu_launches; the only process-global state besides the last CUDA error). */
long long dincae_launch_count(int reset);
/* sm_count, compute capability of the current device. */
int dincae_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- input pipeline ---------------------------------------------------------------- */

/* Temporal-mean removal, DINCAE/__init__.py:167-169,180 (data_generator_list pre-pass).
 * data[T,H,W] f32, missing[T,H,W] u8 (1 = masked).  Writes
 *   meandata[H,W] f64 = (f32 running sum over valid t) * 1. / count   (numpy MaskedArray.mean)
 *   never[H,W]    u8  = 1 where count == 0 (meandata.mask)
 *   anom[T,H,W]   f32 = f32(((f64)data - meandata) / obs_var), 0 where missing.
 * nomask != 0 selects numpy's ndarray.mean path for an input without a mask (f32 mean). */
int dincae_prepare_cube(const float* data, const uint8_t* missing, int T, int H, int W,
                        double obs_var, int nomask,
                        float* anom, double* meandata, uint8_t* never, void* stream);

/* One mini-batch of network inputs, DINCAE/__init__.py:196-227 (closure datagen) plus the
 * tf.data batching of :399-401.  For b < B, frame i = frame_idx[b]:
 *   xin  P4 [B][nvarP][H][W][4]: channels [ (anom_j, invvar_j) j<ndata | lon lat cos sin |
 *        (anom_j, invvar_j) at clamp(i+nn) for each nn != 0 of the time window ]
 *   xtrue [B][H][W][2] = channels 0,1 of the UNPERTURBED frame.
 * cloud_idx != NULL (train mode): where cloud_missing[j][cloud_idx[b]] is set (the `missing`
 *   argument of the reference generator; NULL = same array as `missing`), channels 2j,2j+1
 *   of the current frame are zeroed (:214-219); then jitter is added to channels 2j,
 *   2j+2nd+4, 2j+4nd+4 (:222-225): from noise[B][3*ndata][H][W] (f64 standard normals in the
 *   reference's draw order; x = f32((f64)x + jitter_std[j]*noise), bit-exact parity mode)
 *   or, if noise == NULL, from Philox4x32-10 keyed by (seed, noise_seq[b], pixel).
 * n_noncloud (int32, device) is incremented by the number of xtrue[...,1] != 0 (:541-544).
 * anom/missing: [ndata][T][H][W]; inv_var, jitter_std: host arrays of ndata doubles. */
int dincae_build_batch(const float* anom, const uint8_t* missing,
                       const uint8_t* cloud_missing,
                       const float* lon_scaled, const float* lat_scaled,
                       const float* doy_cos, const float* doy_sin,
                       const double* inv_var_host, const double* jitter_std_host,
                       int ndata, int T, int H, int W, int ntime_win,
                       const int32_t* frame_idx, const int32_t* cloud_idx, int B,
                       const double* noise, uint64_t seed, const int64_t* noise_seq,
                       float* xin, int nvar_planes, float* xtrue, int32_t* n_noncloud,
                       void* stream);

/* The same batch for the bf16 network: xin8 = bf16 P8 planes [B][planes8][H][W][8] of the stacked
 * channels (bit-identical to dincae_build_batch followed by dincae_p4_to_p8; planes beyond the
 * stacked channels are zero).  xin (fp32 P4) may be NULL -- then only the bf16 copy is written.
 * Replaces the same reference lines as dincae_build_batch. */
int dincae_build_batch_p8(const float* anom, const uint8_t* missing,
                          const uint8_t* cloud_missing,
                          const float* lon_scaled, const float* lat_scaled,
                          const float* doy_cos, const float* doy_sin,
                          const double* inv_var_host, const double* jitter_std_host,
                          int ndata, int T, int H, int W, int ntime_win,
                          const int32_t* frame_idx, const int32_t* cloud_idx, int B,
                          const double* noise, uint64_t seed, const int64_t* noise_seq,
                          float* xin, int nvar_planes, void* xin8, int planes8, float* xtrue,
                          int32_t* n_noncloud, void* stream);

/* ---- 3x3 convolutions -------------------------------------------------------------- */

/* Which kernels dincae_conv3x3_fwd / _dgrad / _wgrad launch: 1 (default) = tcgen05 kind::tf32
 * implicit GEMMs with TMEM accumulators and TMA loads (fp32 accumulate), 0 = fp32 FFMA kernels
 * (bit-for-bit fp32 arithmetic; also taken automatically for shapes the tensor-core path does
 * not cover).  Process-wide setting.
 * TF32 contract of back end 1: the tensor cores ignore the low 13 mantissa bits of an
 * operand.  Data staged through registers is rounded (nearest, ties away) by the kernels;
 * conv sources stored at the conv's own resolution are fetched by TMA untouched and must
 * therefore already be TF32-representable.  Every producer in this library rounds on store
 * while back end 1 is selected: the outputs of conv3x3_fwd / conv3x3_dgrad (both kernel
 * families), `xin` of dincae_build_batch (xtrue stays exact) and `g_out` of dincae_head_loss.
 * A caller feeding its own tensors should round them the same way (other values are
 * truncated, not rejected). */
int dincae_set_conv_backend(int backend);
int dincae_get_conv_backend(void);

/* tf.layers.conv2d(3x3, SAME) + leaky_relu [+ average_pooling2d(2,2,SAME)],
 * DINCAE/__init__.py:442-452 (encoder) and :496-513 (decoder: resize + concat + conv).
 * Input = concat(src0 [c0p planes; half-res if src0_half], src1 [c1p planes, may be NULL]).
 * Outputs (any may be NULL): out_full P4 [B][coutp][H][W][4] activated;
 * out_pool P4 at (ceil(H/2),ceil(W/2)) = SAME avg-pool of the activated output (padding
 * excluded from the divisor); out_sign = sign mask of the pre-activation. */
int dincae_conv3x3_fwd(const float* src0, int c0p, int src0_half,
                       const float* src1, int c1p,
                       const float* wpack, const float* bias, int cout_pad,
                       int B, int H, int W, int coutp, float neg_slope,
                       float* out_full, float* out_pool, uint16_t* out_sign,
                       void* stream);

/* Conv2DBackpropInput with the neighbouring pointwise gradients fused
 * (what optimizer.compute_gradients builds for :442-452 / :496-513).
 * Incoming gradient G w.r.t. the conv PRE-activation, gp planes at H x W, given either as
 *   g_full  P4 [B][gp][H][W][4], or
 *   g_pool  P4 at half res + sign: G = g_pool[y>>1][x>>1] / window_count * (sign ? 1 : neg_slope)
 *           (AvgPoolGrad followed by LeakyReluGrad of an encoder stage).
 * dU = G (*) W^T over the conv's cin planes, of which
 *   planes [0, c_lo)      (the half-res source): summed over each 2x2 block
 *                          (ResizeNearestNeighborGrad), multiplied by
 *                          (prev_act > 0 ? 1 : prev_slope) and written to g_prev
 *                          P4 [B][c_lo][ceil(H/2)][ceil(W/2)][4];  prev_act may be NULL (factor 1);
 *   planes [c_lo, c_lo+c_hi): written full-res to dx_hi P4 [B][c_hi][H][W][4], plus dx_add
 *                          (same shape) if not NULL.  c_hi == 0 skips them entirely. */
int dincae_conv3x3_dgrad(const float* g_full, const float* g_pool, const uint16_t* sign,
                         float neg_slope, int gp,
                         const float* wpack, int cinp, int cout_pad,
                         int B, int H, int W,
                         int c_lo, const float* prev_act, float prev_slope, float* g_prev,
                         int c_hi, float* dx_hi, const float* dx_add,
                         void* stream);

/* Conv2DBackpropFilter + BiasAddGrad.  Same input description as conv3x3_fwd and same
 * gradient description as conv3x3_dgrad.  dw in WP layout [9][c0p+c1p][cout_pad][4],
 * db [cout_pad].  Deterministic two-stage reduction through `workspace`
 * (dincae_conv3x3_wgrad_workspace() bytes). */
size_t dincae_conv3x3_wgrad_workspace(int c0p, int c1p, int cout_pad, int B, int H, int W);
/* Diagnostic: 1 when a weight gradient of this shape (cinp = c0p + c1p input planes, gp gradient
 * planes) runs on the tensor-core kernel under back end 1, 0 when it takes the FFMA kernel
 * (shape outside the tcgen05 plan).  No device work. */
int dincae_conv3x3_wgrad_on_tensor_cores(int cinp, int gp, int B, int H, int W);
int dincae_conv3x3_wgrad(const float* src0, int c0p, int src0_half,
                         const float* src1, int c1p,
                         const float* g_full, const float* g_pool, const uint16_t* sign,
                         float neg_slope, int gp, int cout_pad,
                         int B, int H, int W,
                         float* dw, float* db, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ---- bf16 convolution path (BASELINE configs[3]: 512x512, bf16 training) --------------------
 * Activations "planar-8" (P8): bf16 [B][CP8][H][W][8], CP8 = ceil(C/8) planes of 8 channels (16
 * bytes per pixel and plane, lanes past C zero).  Sign masks: uint8 [B][CP8][H][4*ceil(W/4)], bit j
 * of the byte of pixel x = lane j of the pre-activation was > 0.  Parameters stay fp32 masters in
 * the WP layout with every concatenated source padded to a multiple of 8 channels (an even number
 * of P4 planes); dincae_conv_weights_bf16 converts one layer's kernel into the two bf16 operand
 * copies the tensor-core kernels read (call it once per optimisation step):
 *   wf [ceil(cin8/2)][9][2][cout_pad][8]  forward (K = input channels), cout_pad % 16 == 0
 *   wd [ceil(co8/2)][9][2][nd][8]         data gradient (K = output channels, taps flipped) for
 *                                         the nd (a multiple of 16, <= 256) input channels from
 *                                         ci_first on; a layer with more input channels than one
 *                                         MMA is wide takes one wd and one conv3x3_dgrad_bf16
 *                                         call per channel range (e.g. c_lo and c_hi separately)
 * Either of wf / wd may be NULL (that copy is skipped).
 * dincae_conv_weights_bf16_elems returns the bf16 element count of wf (and of wd through
 * wd_elems).  All three convolution entry points mirror their fp32 counterparts above (same
 * fusions, same reference lines DINCAE/__init__.py:442-452, 496-513), with plane counts in units
 * of 8 channels; they run tcgen05.mma kind::f16 (bf16 operands, fp32 accumulation in TMEM).
 * conv3x3_fwd_bf16 can additionally write a planar-4 fp32 copy of the activated full-resolution
 * output (out_f32, out_f32_planes P4 planes): the last decoder layer hands the head / loss kernels
 * fp32 values.  conv3x3_wgrad_bf16 writes fp32 gradients in the WP layout (cinp4 = 2 * (c0p+c1p)). */
size_t dincae_conv_weights_bf16_elems(int cinp4, int cout_pad, int nd, size_t* wd_elems);
int dincae_conv_weights_bf16(const float* wpack, int cinp4, int cout_pad, int ci_first, int nd,
                             void* wf, void* wd, void* stream);
int dincae_p4_to_p8(const float* src, int c4, int c8, int B, int H, int W, void* dst, void* stream);
int dincae_p8_to_p4(const void* src, int c4, int c8, int B, int H, int W, float* dst, void* stream);
int dincae_conv3x3_fwd_bf16(const void* src0, int c0p, int src0_half, const void* src1, int c1p,
                            const void* wf, const float* bias, int cout_pad,
                            int B, int H, int W, int coutp, float neg_slope,
                            void* out_full, void* out_pool, uint8_t* out_sign,
                            float* out_f32, int out_f32_planes, void* stream);
int dincae_conv3x3_dgrad_bf16(const void* g_full, const void* g_pool, const uint8_t* sign,
                              float neg_slope, int gp, const void* wd, int nd,
                              int B, int H, int W,
                              int c_lo, const void* prev_act, float prev_slope, void* g_prev,
                              int c_hi, void* dx_hi, const void* dx_add, void* stream);
size_t dincae_conv3x3_wgrad_bf16_workspace(int cinp4, int cout_pad);
int dincae_conv3x3_wgrad_bf16(const void* src0, int c0p, int src0_half, const void* src1, int c1p,
                              const void* g_full, const void* g_pool, const uint8_t* sign,
                              float neg_slope, int gp, int cinp4, int cout_pad,
                              int B, int H, int W, float* dw, float* db,
                              void* workspace, size_t workspace_bytes, void* stream);
/* dincae_head_loss with the gradient written as planar-8 bf16 (lanes 0,1; the other 6 zero). */
int dincae_head_loss_bf16(const float* xrec, const float* xtrue, int B, int H, int W,
                          int truth_uncertain, float neg_slope, const int32_t* n_noncloud,
                          void* g_out, double* sums, float* loss_out,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- dense bottleneck -------------------------------------------------------------- */

/* tf.layers.dense(activation=relu), DINCAE/__init__.py:479-483 (dropout is inert there).
 * y[B,N] = act(x[B,K] w[K,N] + bias[N]); relu != 0 applies max(.,0). */
size_t dincae_dense_workspace(int B, int K, int N);
int dincae_dense_fwd(const float* x, const float* w, const float* bias, int B, int K, int N,
                     int relu, float* y, void* workspace, size_t workspace_bytes, void* stream);
/* dx[B,K] = (dz[B,N] w^T) * (act_in > 0) (act_in NULL: no mask)  -- MatMul grad + ReluGrad. */
int dincae_dense_bwd_data(const float* dz, const float* w, const float* act_in,
                          int B, int K, int N, float* dx,
                          void* workspace, size_t workspace_bytes, void* stream);
/* dw[K,N] = x^T dz, db[N] = sum_b dz. */
int dincae_dense_bwd_weight(const float* x, const float* dz, int B, int K, int N,
                            float* dw, float* db, void* stream);

/* ---- dense bottleneck on the tensor cores (bf16 path) ---------------------------------------------
 * tf.layers.dense forward and its data gradient (DINCAE/__init__.py:479-483) as tcgen05.mma
 * kind::f16 GEMMs over a TILED bf16 copy of the kernel, w8 [K/8][N/8][8 k][8 n] (K, N multiples of 8;
 * dincae_dense_tc_tile_weights makes it from the fp32 master [K][ldw]; the factored Adam kernel
 * can keep it up to date).  One copy serves both products.  x / dz / outputs stay fp32 row-major
 * (row pitches ldx / ldy / lddz / lddx); B <= 256.  fp32 accumulation; bias / ReLU / ReLU gate as in
 * dincae_dense_fwd / dincae_dense_bwd_data.  workspace: dincae_dense_tc_workspace() bytes. */
size_t dincae_dense_tc_weights_elems(int K, int N);
int dincae_dense_tc_tile_weights(const float* w, int ldw, int K, int N, void* w8, void* stream);
size_t dincae_dense_tc_workspace(int B, int K, int N);
int dincae_dense_tc_fwd(const float* x, int ldx, const void* w8, const float* bias, int B, int K, int N,
                        int relu, float* y, int ldy, void* workspace, size_t workspace_bytes,
                        void* stream);
int dincae_dense_tc_bwd_data(const float* dz, int lddz, const void* w8, const float* act_in, int B, int K,
                             int N, float* dx, int lddx, void* workspace, size_t workspace_bytes,
                             void* stream);

/* ---- dropout (optional; the reference's tf.layers.dropout is inert) -----------------
 * dincae_dropout_fwd: y[b][n] <- y * keep / (1 - rate) in place, keep drawn from Philox4x32-10 keyed by
 *   (seed, step, row key, n) with row key = row_keys[b] (device int64, e.g. the frame's sequence
 *   number) or first_global_row + b when row_keys is NULL: independent of batch slot and GPU count.
 *   `step` distinguishes the layers / calls that share a seed.  mask (may be NULL): keep mask, one
 *   byte per element [B][N].
 * Backward: dropout follows a ReLU, so the stored output is > 0 exactly where the unit was active
 *   AND kept -- the ReLU gate of dincae_dense_bwd_data / conv3x3_dgrad already applies the mask; the
 *   1 / (1 - rate) factor is dincae_scale_rows on that gradient. */
int dincae_dropout_fwd(float* y, int B, int N, int ld, float rate, uint64_t seed, uint64_t step,
                       int first_global_row, const int64_t* row_keys, uint8_t* mask, void* stream);
int dincae_scale_rows(float* x, int B, int N, int ld, float factor, void* stream);

/* ---- dense-bottleneck gradients in factored form (new capability) ------------------------------
 * The gradient of a dense kernel is dW = x^T dz (x [Bt,K]: the layer input, dz [Bt,N]: gradient of
 * its pre-activation, Bt frames of the GLOBAL batch).  These entry points let the optimiser work
 * from the factors instead of a materialised dW (DINCAE/__init__.py:479-483 MatMul gradients,
 * :598-603 clip_by_global_norm + ApplyAdam): no dW write / re-reads, and data-parallel ranks
 * all-gather Bt x (K + N) floats instead of all-reducing K x N.
 * Factor rows are addressed as `world` per-rank blocks: row bt lives at
 *   base + (bt / rows_per_rank) * rank_stride + (bt % rows_per_rank) * ld      (floats),
 * i.e. an all-gathered buffer of per-rank blobs is used in place (world = 1: a plain matrix).
 *   dense_factored_sumsq : *sumsq += ||x^T dz||_F^2 = sum_{b,b'} (x_b.x_b')(dz_b.dz_b') (fixed order).
 *   adam_update_factored : TF-Adam on w [K][ldw] (N columns) with g = x^T dz * grad_scale / *denom *
 *                          state[4] (clip scale), lr_t = state[3] -- the scalars adam_step_multi left;
 *                          w8 != NULL: also refreshes the tiled bf16 copy of dincae_dense_tc_*.
 *   dense_factored_expand: dw [K][ldw] = x^T dz (tests / callers that want the matrix).
 *   dense_bias_grad      : db[n] = sum_b dz[b][n] over the B rows (row pitch ld) of one rank.
 *   adam_step_multi      : global norm + TF-Adam over n parameters whose gradient is given as `world`
 *                          copies (summed on the fly in rank order), plus *extra_sumsq (device double,
 *                          may be NULL) = sum of squares of the factored gradients; writes state[] like
 *                          dincae_adam_step.  No L2 term on this path.
 *   sum_ranks_f64        : out[k] = sum_r in[r * rank_stride + k] (doubles; the per-rank loss sums). */
size_t dincae_dense_factored_workspace(int Bt, int K, int N);
int dincae_dense_factored_sumsq(const float* x, size_t x_rank_stride, int ldx,
                                const float* dz, size_t dz_rank_stride, int lddz,
                                int world, int rows_per_rank, int K, int N, double* sumsq,
                                void* workspace, size_t workspace_bytes, void* stream);
int dincae_adam_update_factored(float* w, float* m, float* v, int ldw,
                                const float* x, size_t x_rank_stride, int ldx,
                                const float* dz, size_t dz_rank_stride, int lddz,
                                int world, int rows_per_rank, int K, int N,
                                float grad_scale, const double* denom, float beta1, float beta2,
                                float eps, const float* state, void* w8, void* stream);
int dincae_dense_factored_expand(const float* x, size_t x_rank_stride, int ldx,
                                 const float* dz, size_t dz_rank_stride, int lddz,
                                 int world, int rows_per_rank, int K, int N, float* dw, int ldw,
                                 void* stream);
int dincae_dense_bias_grad(const float* dz, int B, int N, int ld, float* db, void* stream);
size_t dincae_adam_multi_workspace(size_t n);
int dincae_adam_step_multi(float* params, const float* grads, size_t grad_rank_stride, int world,
                           float* m, float* v, size_t n, float clip_norm, const float* lr,
                           float beta1, float beta2, float eps, float grad_scale, const double* denom,
                           const double* extra_sumsq, float* state,
                           void* workspace, size_t workspace_bytes, void* stream);
int dincae_sum_ranks_f64(const double* in, size_t rank_stride, int world, int count, double* out,
                         void* stream);
/* count (<= 32) doubles <-> (hi, lo) float pairs: lets ONE float32 all-reduce carry the gradient
 * buffer and, in its tail, the loss sums (pair sums recombine to ~1e-14 relative). */
int dincae_f64_to_hilo(const double* in, int count, float* out, void* stream);
int dincae_hilo_to_f64(const float* in, int count, double* out, void* stream);

/* ---- output head, loss ------------------------------------------------------------- */

/* DINCAE/__init__.py:520-523 + sinv :298-299: xrec P4 [B][1][H][W][4] (lanes 0,1) ->
 * m_rec, sigma2_rec [B][H][W]. */
int dincae_head_recon(const float* xrec, int B, int H, int W,
                      float* m_rec, float* sigma2_rec, void* stream);

/* Head fused with the per-batch arithmetic of `savesample` (DINCAE/__init__.py:237-248,
 * 288-291), emitted as the records of the output file: records[b] = [mean_rec[b][H][W] |
 * sigma_rec[b][H][W]] float32 with mean_rec = float32(float64(m_rec) + meandata[H][W])
 * (mean_is_f32 != 0: the float32 sum m_rec + float32(meandata), numpy's result type when the
 * input had no mask), sigma_rec = sqrt(sigma2_rec), both = fill_value where
 * never[H][W] != 0 (meandata.mask; NULL: nothing masked).  big_endian != 0 stores every value byte-reversed: the buffer is
 * then the byte image of NetCDF-3 records (time-major, the two record variables
 * interleaved) and goes to the file with one write.  Identity output transform only. */
int dincae_head_recon_records(const float* xrec, int B, int H, int W, const double* meandata,
                              const uint8_t* never, float fill_value, int mean_is_f32,
                              int big_endian, void* records, void* stream);

/* Masked Gaussian NLL (or its truth_uncertain KL variant), RMS and the gradient w.r.t.
 * the last conv's pre-activation, DINCAE/__init__.py:520-570 and their TF gradients.
 * n_noncloud: device int32 normaliser n of :544, or NULL.  With NULL the gradient is left
 *   UN-normalised (n := 1) so that data-parallel ranks can sum gradients and counts first;
 *   dincae_adam_step() then divides by the reduced count (its `denom` argument).
 * g_out P4 [B][1][H][W][4] = dcost/dxrec * (xrec > 0 ? 1 : neg_slope), lanes 2,3 zero.
 * sums (device, 4 doubles) receives [sum w log s2 (+KL terms), sum w d^2/s2, sum w d^2, sum w];
 * loss_out (device, 2 floats) = [cost, RMS] from these sums over n (n = sum w when
 *   n_noncloud is NULL).
 * workspace: dincae_head_loss_workspace() bytes. */
size_t dincae_head_loss_workspace(int B, int H, int W);
int dincae_head_loss(const float* xrec, const float* xtrue, int B, int H, int W,
                     int truth_uncertain, float neg_slope, const int32_t* n_noncloud,
                     float* g_out, double* sums, float* loss_out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- optimizer --------------------------------------------------------------------- */

/* tf.clip_by_global_norm + tf.train.AdamOptimizer.apply_gradients (+ the gradient of the
 * optional L2 term), DINCAE/__init__.py:563-567, 598-603.
 * params/grads/m/v: flat f32 arrays of n entries; the first n_decay entries are the
 * "kernel" variables that receive g += l2_beta * p before clipping.
 * state (device, 8 floats): [0] beta1^t, [1] beta2^t (initialise to beta1, beta2),
 * [2] global norm of this step, [3] lr_t, [4] clip scale; lr: device float.
 * grads are first multiplied by grad_scale and, if denom (device double) is not NULL, divided
 * by *denom (the all-reduced count of observed pixels, see dincae_head_loss). */
size_t dincae_adam_workspace(size_t n);
int dincae_adam_step(float* params, const float* grads, float* m, float* v,
                     size_t n, size_t n_decay, float l2_beta, float clip_norm,
                     const float* lr, float beta1, float beta2, float eps, float grad_scale,
                     const double* denom, float* state, void* workspace, size_t workspace_bytes, void* stream);

/* 0.5 * sum p^2 over the first n_decay params -> out (device float): value of the L2 term
 * of the reported cost (:565-567), multiplied by l2_beta by the caller. */
int dincae_l2_half_sumsq(const float* params, size_t n_decay, float* out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* dincae_conv3x3_wgrad in two steps, for callers that overlap the (small) reduction of one
 * layer with the main kernel of the next on another stream: `_partial` writes *n_partials
 * (host int, known when the call returns) partial gradients into `workspace`; `_reduce` sums
 * them in a fixed order into dw / db.  Same arguments and layouts as dincae_conv3x3_wgrad;
 * cinp = c0p + c1p. */
int dincae_conv3x3_wgrad_partial(const float* src0, int c0p, int src0_half, const float* src1, int c1p,
                                 const float* g_full, const float* g_pool, const uint16_t* sign,
                                 float neg_slope, int gp, int cout_pad, int B, int H, int W,
                                 void* workspace, size_t workspace_bytes, int* n_partials, void* stream);
int dincae_conv3x3_wgrad_reduce(const void* workspace, int n_partials, int cinp, int gp, int cout_pad,
                                float* dw, float* db, void* stream);

/* ---- ensemble averaging of saved reconstructions ------------------------------------- */

/* The reference's README (README.md:90) recommends averaging the reconstructions saved every
 * `save_each` epochs (DINCAE/__init__.py:671-689 writes them); it ships no code for it.
 * accumulate: add one member (mean_rec, sigma_rec: n floats each, `fill` marks masked
 *   entries, which are skipped) into acc[3][n] doubles = (sum m, sum sigma^2, sum m^2) and
 *   count[n] (both zero-initialised by the caller).
 * finalize: mean_out = sum m / count; sigma_out = sqrt(mean sigma^2) (mixture == 0) or
 *   sqrt(mean sigma^2 + variance of the member means) (mixture != 0); `fill` where count == 0. */
int dincae_ensemble_accumulate(const float* mean_rec, const float* sigma_rec, float fill, size_t n,
                               double* acc, int32_t* count, void* stream);
int dincae_ensemble_finalize(const double* acc, const int32_t* count, size_t n, int mixture,
                             float fill, float* mean_out, float* sigma_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DINCAE_B200_H */
