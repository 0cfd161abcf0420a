// Launch-side declarations of the EEGNet kernels (definitions in eegnet_fwd.cu /
// eegnet_bwd.cu / optim.cu).  Every launcher only enqueues work on `st`.
#pragma once
#include "eav_common.cuh"

namespace eav {

// ---- forward -------------------------------------------------------------------
// M1: temporal conv (EEGNet_tor.py:24,51): x -> y1 raw [N][F1][C][T] (+ BN1 partial sums)
// wt_scratch: tconv_fwd_tc_scratch_floats(d) floats for the packed tensor-core weight operand (may be null
// when that is 0)
int launch_tconv_fwd(const NetDims &d, const float *x, const int32_t *x_index, const float *params,
                     float *wt_scratch, float *y1, float *part, int *part_rows, cudaStream_t st);
bool tail_bwd_folds_bn3(const NetDims &d);   // eval mode, variant 0: dz3 leaves tail_bwd already scaled by BatchNorm-3's k
bool tc_path_enabled(const char *var);   // false when EAV_TC or <var> is "ffma" / "0"
// tensor-core (tcgen05) block-2 convolution, sepconv_tc.cu (mode 0 forward, 1 input gradient)
bool sepconv_use_tc(const NetDims &d);
size_t sepconv_tc_scratch_floats(const NetDims &d);
int sepconv_tc_rows_per_model(const NetDims &d);
int launch_sepconv_tc(const NetDims &d, int mode, const float *in, const float *params, float *wt_scratch, float *out,
                      float *part, int *part_rows, cudaStream_t st);
int sepconv_dw_tc_splits(const NetDims &d);       // partial slabs per model written by the tensor-core dW3 kernel
int launch_sepconv_dw_tc(const NetDims &d, const float *dy3, const float *d1, float *part, float *grads, int S,
                         cudaStream_t st);
// tensor-core (tcgen05) variant of the temporal convolution, tconv_tc.cu
bool tconv_fwd_use_tc(const NetDims &d);
size_t tconv_fwd_tc_scratch_floats(const NetDims &d);
int tconv_fwd_tc_rows_per_sample(const NetDims &d);
int launch_tconv_fwd_tc(const NetDims &d, const float *x, const int32_t *x_index, const float *params,
                        float *wt_scratch, float *y1, float *part, int *part_rows, cudaStream_t st);
// M2/M5/M7 statistics: partial sums -> {mean, invstd, scale, shift}; running-stat update
int launch_bn_finalize(const NetDims &d, int layer, const float *part, int rows_per_model,
                       double count, const double *sums, const float *params, float *bn_state, float4 *stats,
                       cudaStream_t st);
// eval mode: all three layers' {mean, invstd, scale, shift} from the running statistics in one launch
int launch_bn_eval_finalize_all(const NetDims &d, const float *params, const float *bn_state, float4 *s1, float4 *s2,
                                float4 *s3, cudaStream_t st);
// block-1 tail of the fused eval-mode backward: dW2 and dW1 partial sums -> grads, BatchNorm-1 gradients, in one launch
int launch_block1_bwd_finalize(const NetDims &d, const float *partw2, const float *partw1, const float *partbn, int S,
                               const float *params, const float4 *bnf1, const float4 *bnf2, float4 *bnb1, float *grads,
                               cudaStream_t st);
// partial rows -> per-(model, channel) float64 sums (the buffer a data-parallel run all-reduces)
int launch_bn_reduce(const NetDims &d, int layer, const float *part, int rows_per_model, double *sums,
                     cudaStream_t st);
// M2+M3+M4: BN1 (+ELU) + depthwise spatial conv -> y2 raw [N][G][T] (+ BN2 partial sums)
// bn2_pool / d1 non-null: eval-mode fusion of M5 (BN2 + ELU + AvgPool(1,4)) into the same kernel, see dw_fwd_fuses_pool()
int launch_dw_fwd(const NetDims &d, const float *y1, const float *params, const float4 *bn1,
                  float *y2, float *part, int *part_rows, const float4 *bn2_pool, float *d1, cudaStream_t st);
bool dw_fwd_fuses_pool(const NetDims &d);
// M5: BN2 + ELU + AvgPool(1,P1) + dropout -> d1 [N][G][T4]
int launch_pool1_fwd(const NetDims &d, const float *y2, const float4 *bn2, const uint8_t *mask1,
                     float *d1, cudaStream_t st);
// M6: (1,K2) 'same' conv over G channels (variant 0) -> y3 raw [N][F2][T4] (+ BN3 partials)
int launch_sepconv_fwd(const NetDims &d, const float *d1, const float *params, float *wt_scratch, float *y3,
                       float *part, int *part_rows, cudaStream_t st);
// variant 1 block 2: depthwise temporal conv + pointwise conv (CNN_EEG.py:35-37)
int launch_dwt_fwd(const NetDims &d, const float *d1, const float *params, float *y3d, cudaStream_t st);
int launch_pw_fwd(const NetDims &d, const float *y3d, const float *params, float *y3, float *part,
                  int *part_rows, cudaStream_t st);
// M7+M8: BN3 + ELU + AvgPool(1,P2) + dropout + flatten + dense (+ softmax) -> feat, out
int launch_tail_fwd(const NetDims &d, const float *y3, const float4 *bn3, const uint8_t *mask2,
                    const float *params, float *feat, float *out, float *probs_saved, cudaStream_t st);

// ---- backward ------------------------------------------------------------------
int launch_tail_bwd(const NetDims &d, const float *dout, const float *probs, const float *params,
                    const float *y3, const float4 *bn3, const uint8_t *mask2, float *dz,
                    float *dz3, float *part, cudaStream_t st);
int launch_dense_bwd_w(const NetDims &d, const float *feat, const float *dz, float *grads, cudaStream_t st);
int launch_bn_bwd_finalize(const NetDims &d, int layer, const float *part, int rows_per_model,
                           double count, const double *sums, const float *params, const float4 *bnf, float4 *bnb,
                           float *grads, cudaStream_t st);
int launch_bn_bwd_apply(const NetDims &d, float *dz3, const float *y3, const float4 *bnf3, const float4 *bnb3,
                        cudaStream_t st);
int launch_sepconv_bwd_dx(const NetDims &d, const float *dz3, const float *y3, const float4 *bnf3,
                          const float4 *bnb3, const float *params, float *wt_scratch, float *dd1, cudaStream_t st);
int launch_sepconv_bwd_dw(const NetDims &d, const float *dz3, const float *y3, const float4 *bnf3,
                          const float4 *bnb3, const float *d1, float *part, float *grads, cudaStream_t st);
int launch_pw_bwd(const NetDims &d, const float *dz3, const float *y3, const float4 *bnf3,
                  const float4 *bnb3, const float *y3d, const float *params, float *dy3d,
                  float *part, float *grads, cudaStream_t st);
int launch_dwt_bwd(const NetDims &d, const float *dy3d, const float *d1, const float *params,
                   float *dd1, float *part, float *grads, cudaStream_t st);
int launch_pool1_bwd(const NetDims &d, const float *dd1, const float *y2, const float4 *bnf2,
                     const uint8_t *mask1, float *dz2, float *part, cudaStream_t st);
int launch_dw_bwd(const NetDims &d, const float *dz2, const float *y2, const float4 *bnf2,
                  const float4 *bnb2, const float *y1, const float4 *bnf1, const float *params,
                  float *dz1, float *part_w, float *part_bn, float *grads, cudaStream_t st);
int launch_tconv_bwd_dw(const NetDims &d, const float *x, const int32_t *x_index, const float *dz1,
                        const float *y1, const float4 *bnf1, const float4 *bnb1, float *part,
                        float *grads, cudaStream_t st);

bool tconv_bwd_dw_use_tc(const NetDims &d);
int tconv_bwd_dw_tc_splits(const NetDims &d);     // partial slabs per model written by the tensor-core kernel
int launch_tconv_bwd_dw_tc(const NetDims &d, const float *x, const int32_t *x_index, const float *dz1,
                           const float *y1, const float4 *bnf1, const float4 *bnb1, float *part, int S,
                           cudaStream_t st);
// eval-mode fused backward of block 1 (tconv_tc.cu): dz1 regenerated from dz2 + y1 inside the dW1 kernel, which also
// produces dW2 and the BatchNorm-1 backward sums (replaces dw_bwd + tconv_bwd_dw; no dz1 round trip through HBM)
bool tconv_bwd_fused_ok(const NetDims &d);          // this launch takes the fused path (eval-mode BN only)
bool tconv_bwd_fused_shape_ok(const NetDims &d);    // ... could take it in eval mode (workspace sizing)
size_t tconv_bwd_fused_partw2_floats(const NetDims &d);
int launch_tconv_bwd_fused_tc(const NetDims &d, const float *x, const int32_t *x_index, const float *dz2,
                              const float *y1, const float4 *bnf1, const float4 *bnf2, const float *params,
                              float *part, float *partw2, float *partbn, float *grads, int S, cudaStream_t st);
int tconv_fwd_rows_per_sample(const NetDims &d);   // BN1 partial rows per sample written by tconv_fwd
int sepconv_fwd_rows_per_model(const NetDims &d);   // BN3 partial rows per model written by sepconv_fwd
int dw_fwd_tiles(const NetDims &d);   // time tiles per (sample, filter) of dw_fwd == BN2 partial rows per sample

// ---- small ops -------------------------------------------------------------------
int launch_renorm_rows(float *w, int64_t n_rows, int64_t row_len, int64_t row_stride,
                       int64_t rows_per_group, int64_t group_stride, float maxnorm, cudaStream_t st);
int launch_renorm_two(float *wa, int rows_a, int len_a, float *wb, int rows_b, int len_b, int M, int64_t pstride,
                      float maxnorm, cudaStream_t st);
int launch_reduce_partials(const float *part, int n_part, int64_t len, int n_models,
                           int64_t dst_stride, float *dst, cudaStream_t st);

}  // namespace eav
