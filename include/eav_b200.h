/*
 * eav_b200.h -- C ABI of libeav_b200.so: the B200 (sm_100a) implementation of the
 * EEG hot path of nubcico/EAV.
 *
 * The reference is pure Python and has NO plugin / operator / FFI layer
 * (SURVEY.md section 8b): its boundary is the Python class surface
 *     Dataload_eeg.DataLoadEEG            (Dataload_eeg.py:35-160)
 *     EAV_datasplit.EAVDataSplit          (EAV_datasplit.py:7-58)
 *     CNN_torch.EEGNet_tor.{EEGNet_tor,Trainer_uni}   (CNN_torch/EEGNet_tor.py:15-135)
 *     CNN_torch.CNN_EEG.{EEGNet,EEGNetTrainer}        (CNN_torch/CNN_EEG.py:7-162)
 * which eav_b200/ mirrors in Python.  This header is the native layer those
 * mirrors bind with ctypes; each entry point names the reference code it replaces.
 *
 * Conventions (all functions):
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - every pointer argument marked "device" is a CUDA device pointer owned by the
 *     caller; the library never allocates, frees or synchronises.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Calls only enqueue work, so they are CUDA-graph capturable.
 *   - scratch memory comes from the caller: size it with *_workspace_bytes().
 *   - return 0 on success, a negative EAV_ERR_* for argument errors, or a positive
 *     cudaError_t from the launch; eav_last_error_string() describes the last
 *     failure on the calling thread.
 */
#ifndef EAV_B200_H_
#define EAV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EAV_ABI_VERSION 1

#define EAV_ERR_BAD_ARG      (-1)
#define EAV_ERR_UNSUPPORTED  (-2)
#define EAV_ERR_WORKSPACE    (-3)

const char *eav_last_error_string(void);
int eav_abi_version(void);
/* Number of CUDA kernels this library has launched in this process so far. */
uint64_t eav_launch_count(void);
/* 0 when the current device is compute capability 10.x, else EAV_ERR_UNSUPPORTED. */
int eav_check_device(void);

/* ------------------------------------------------------------------------- */
/* Preprocessing: Dataload_eeg.py:85-152 (downsampling -> bandpass_filter ->
 * segment_and_select_classes) for a batch of subjects.                       */
/* ------------------------------------------------------------------------- */
typedef struct eav_preproc_cfg {
    int32_t n_subjects;    /* S                                                  */
    int32_t n_trials;      /* 200                                                */
    int32_t n_chans;       /* 30                                                 */
    int32_t trial_len;     /* 10000 raw samples per trial                        */
    int32_t down;          /* decimation factor fs_orig/fs_target (5)            */
    int32_t n_taps;        /* 2*10*down+1 = 101 (scipy.signal.resample_poly)     */
    int32_t n_sections;    /* biquads in the SOS cascade (5)                     */
    int32_t n_sub;         /* epochs per trial (4)                               */
    int32_t raw_is_f64;    /* 0: raw is float32, 1: raw is float64               */
    int32_t order;         /* EAV_PREPROC_ORDER_*                                */
} eav_preproc_cfg;

/* Dataload_eeg.py:85-121: resample_poly to fs_target, then band-pass at fs_target (the shipped order). */
#define EAV_PREPROC_ORDER_DECIMATE_FIRST 0
/* CNN_tensorflow/CNN_EEG_tf.py:64-75,182-189 (legacy): band-pass the raw recording at fs_orig (`sos` designed
 * for fs_orig), then resample_poly.  float32 recordings only; needs one more raw-sized buffer in the workspace. */
#define EAV_PREPROC_ORDER_BANDPASS_FIRST 1

size_t eav_preproc_workspace_bytes(const eav_preproc_cfg *cfg);

/*
 * raw      device  [S][n_trials][n_chans][trial_len]  f32 (or f64): the .mat memory
 *                  order of `seg` (Dataload_eeg.py:70-82).  Per (subject, channel)
 *                  the trials form ONE continuous sequence (Dataload_eeg.py:94).
 * taps     host    [n_taps] f64   firwin(2*10*down+1, 1/down, ('kaiser',5.0))
 * sos      host    [n_sections][6] f64   butter(5, band, 'bandpass', fs, 'sos')
 * epoch_slot device [S][n_trials] i32: output slot (in units of trials, i.e. the
 *                  first of its n_sub epochs is epoch n_sub*slot) of a kept trial,
 *                  or -1 for a trial that segment_and_select_classes drops.
 * epochs   device  [S][n_epochs_out][n_chans][trial_len/down/n_sub] f32
 *                  (the model's (N,1,Chans,Samples) layout), n_epochs_out per subject.
 * dec_out  device  optional (may be NULL) [S][n_chans][n_trials*trial_len/down] f32:
 *                  also returns the decimated sequence (the reference's self.seg
 *                  after downsampling(), Dataload_eeg.py:102).
 */
int eav_preproc_run(const eav_preproc_cfg *cfg, const void *raw, const double *taps,
                    const double *sos, const int32_t *epoch_slot, int32_t n_epochs_out,
                    float *epochs, float *dec_out, void *workspace, size_t workspace_bytes,
                    void *stream);

/* ------------------------------------------------------------------------- */
/* EEGNet: CNN_torch/EEGNet_tor.py:15-67 (variant 0) and CNN_torch/CNN_EEG.py:7-67
 * (variant 1).  `n_models` independent models (one per subject) are advanced by
 * the same launches; sample n = m*batch + b belongs to model m.               */
/* ------------------------------------------------------------------------- */
#define EAV_VARIANT_TOR 0   /* EEGNet_tor: ELU after BN1, full (1,K2) conv, softmax output */
#define EAV_VARIANT_CNN 1   /* CNN_EEG.EEGNet: no ELU after BN1, depthwise+pointwise, logits */

#define EAV_DROPOUT_NONE  0 /* eval mode or p == 0                                 */
#define EAV_DROPOUT_MASK  1 /* caller supplies keep-masks (uint8 1/0): parity mode */
#define EAV_DROPOUT_PHILOX 2/* on-device Philox4x32-10 keyed by (seed, element)   */
#define EAV_DROPOUT_PHILOX_2D 3 /* nn.Dropout2d (dropoutType != 'Dropout', EEGNet_tor.py:21): one on-device
                                   Philox draw per (sample, channel) row, the whole row kept or zeroed */

typedef struct eav_eegnet_cfg {
    int32_t n_models;      /* M                                                  */
    int32_t batch;         /* B samples per model                                */
    int32_t chans;         /* Chans   (30)                                       */
    int32_t samples;       /* Samples (500)                                      */
    int32_t kern_len;      /* kernLength (300)                                   */
    int32_t F1;            /* 8                                                  */
    int32_t D;             /* 8                                                  */
    int32_t F2;            /* 64                                                 */
    int32_t kern_len2;     /* 16                                                 */
    int32_t pool1;         /* 4                                                  */
    int32_t pool2;         /* 8                                                  */
    int32_t n_classes;     /* nb_classes                                         */
    int32_t variant;       /* EAV_VARIANT_*                                      */
    int32_t bn_train;      /* 1: batch statistics + running-stat update; 0: running stats */
    int32_t dropout_mode;  /* EAV_DROPOUT_*                                      */
    int32_t param_stride;  /* floats between consecutive models in params/grads/m/v (>= n_params) */
    int32_t bn_stride;     /* floats between consecutive models in bn_state (>= 2*(F1+F1*D+F2)) */
    float   dropout_p;     /* dropoutRate                                        */
    float   bn_eps;        /* 1e-5                                               */
    float   bn_momentum;   /* 0.1                                                */
    float   norm_rate;     /* max-norm of the forward hooks (variant 0); <= 0 disables */
    uint64_t seed;         /* Philox key (EAV_DROPOUT_PHILOX)                    */
    uint64_t step;         /* Philox stream position: change every step          */
    uint64_t step_device_ptr; /* if non-zero: device address of an int64 step counter that
                              replaces `step` (read at kernel run time, so a captured CUDA
                              graph can be replayed with a fresh dropout stream)            */
    int32_t dp_world;      /* data-parallel replicas sharing ONE model (large-batch mode, config 5):
                              `batch` is the per-rank share of a global batch of batch*dp_world.
                              0 or 1 = single device.  See eav_eegnet_stage_allreduce().         */
    int32_t reserved;
} eav_eegnet_cfg;

/* Number of parameters of one model and the offsets (in floats) of its tensors inside
 * a model's slice of the flat arena, in the reference's construction order
 * (EEGNet_tor.py:24-43 / CNN_EEG.py:20-55).  offsets must hold 12 entries:
 *   variant 0: W1 g1 b1 W2 g2 b2 W3 g3 b3 Wd bd (11 used)
 *   variant 1: W1 g1 b1 W2 g2 b2 W3dw W3pw g3 b3 Wc bc (12 used)
 * Returns n_params, or a negative error. */
int64_t eav_eegnet_param_layout(const eav_eegnet_cfg *cfg, int64_t *offsets);

size_t eav_eegnet_workspace_bytes(const eav_eegnet_cfg *cfg);

/* Inspection hook for tests/profilers: byte offsets inside the workspace of the saved
 * activations and backward scratch, in this order (16 entries):
 *   y1 y2 d1 y3d y3 feat probs dz  dz3 dd1 dy3d dz2 dz1  bnf1 bnf2 bnf3
 * (y1 [N][F1][C][T] raw conv output, y2 [N][G][T], d1 [N][G][T/4], y3 [N][F2][T/4],
 *  feat [N][F2*T/32]; dz* are the gradients w.r.t. the BatchNorm OUTPUTS). */
int eav_eegnet_workspace_offsets(const eav_eegnet_cfg *cfg, size_t *offsets16);

/*
 * Forward (EEGNet_tor.py:50-67 / CNN_EEG.py:57-67).
 * x        device [n_rows][chans][samples] f32 -- the resident dataset (or the batch)
 * x_index  device [M*B] i32 or NULL: row of x used by sample n (NULL = row n).
 * params   device [M][param_stride] f32.  NOT const: variant 0 applies the max-norm
 *          forward hooks (EEGNet_tor.py:33-34,47-48) after the layer used W_old.
 * bn_state device [M][bn_stride] f32: rm1 rv1 rm2 rv2 rm3 rv3 (updated when bn_train).
 * mask1    device [M*B][F1*D][samples/pool1] u8 keep-mask (EAV_DROPOUT_MASK) or NULL
 * mask2    device [M*B][F2][samples/pool1/pool2] u8 keep-mask or NULL
 * out      device [M*B][n_classes] f32: probabilities (variant 0) / logits (variant 1)
 * workspace: saved activations for eav_eegnet_backward (same cfg, same buffers).
 */
int eav_eegnet_forward(const eav_eegnet_cfg *cfg, const float *x, const int32_t *x_index,
                       float *params, float *bn_state, const uint8_t *mask1,
                       const uint8_t *mask2, float *out, void *workspace,
                       size_t workspace_bytes, void *stream);

/* The side effect a forward pass has on the weights -- the two max-norm hooks of variant 0 (EEGNet_tor.py:33-34,47-48:
 * rows of depthwiseConv.weight and dense.weight renormed to L2 <= norm_rate) -- without the forward pass.  No-op for
 * variant 1 or norm_rate <= 0. */
int eav_eegnet_apply_hooks(const eav_eegnet_cfg *cfg, float *params, void *stream);

/*
 * nn.CrossEntropyLoss()(out, targets) per model (EEGNet_tor.py:81,105) and its
 * gradient w.r.t. `out`.
 * targets  device [n_rows] i64 class indices, addressed through x_index like x.
 * loss     device [M] f32: mean over the model's batch.
 * dout     device [M*B][n_classes] f32 (may be NULL: loss only).
 * n_correct device [M] i32 (may be NULL): #samples whose argmax(out) == target.
 */
int eav_eegnet_loss(const eav_eegnet_cfg *cfg, const float *out, const int64_t *targets,
                    const int32_t *x_index, float *loss, float *dout, int32_t *n_correct,
                    void *stream);

/*
 * Backward of eav_eegnet_forward given d(loss)/d(out).
 * grads    device [M][param_stride] f32, every parameter's gradient is overwritten.
 */
int eav_eegnet_backward(const eav_eegnet_cfg *cfg, const float *x, const int32_t *x_index,
                        const float *params, const float *dout, const uint8_t *mask1,
                        const uint8_t *mask2, float *grads, void *workspace,
                        size_t workspace_bytes, void *stream);

/*
 * Profiling hook: forward and backward are fixed sequences of named kernel stages
 * (forward = [0, eav_eegnet_stage_forward_end()), backward = the rest).
 * eav_eegnet_run_stage launches exactly ONE stage on the buffers a previous full
 * forward/backward left in the workspace, so bench.py can time each kernel with CUDA
 * events on the launching stream.  Same argument meaning as forward/backward.
 */
int eav_eegnet_stage_count(void);
int eav_eegnet_stage_forward_end(void);
const char *eav_eegnet_stage_name(int stage);
int eav_eegnet_run_stage(const eav_eegnet_cfg *cfg, int stage, const float *x, const int32_t *x_index,
                         float *params, float *bn_state, const uint8_t *mask1, const uint8_t *mask2,
                         float *out, const float *dout, float *grads, void *workspace,
                         size_t workspace_bytes, void *stream);

/*
 * Large-batch data-parallel mode (dp_world > 1, BASELINE.json configs[4]): the caller drives
 * the stages one by one and, after a stage for which this returns n_doubles > 0, all-reduces
 * (sum, e.g. ncclAllReduce over NVLink) the n_doubles float64 values at workspace+offset_bytes
 * across the replicas before launching the next stage.  These are the BatchNorm statistics
 * (forward: sum, sum of squares; backward: sum dz, sum dz*xhat) that make train-mode BN equal to
 * the single-device result at the GLOBAL batch; eval-mode BN needs none.  The flat gradient
 * arena is all-reduced by the caller after the last stage (gradients are already scaled by
 * 1/(batch*dp_world) through eav_eegnet_loss).
 */
int eav_eegnet_stage_allreduce(const eav_eegnet_cfg *cfg, int stage, size_t *offset_bytes, size_t *n_doubles);

/*
 * torch.optim.Adam step (EEGNet_tor.py:82,110; betas/eps/no weight decay as there)
 * over a flat arena of n floats: m = b1*m+(1-b1)*g; v = b2*v+(1-b2)*g*g;
 * p -= (lr/(1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
 * step_count is t (1-based, already incremented by the caller).
 */
int eav_adam_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq,
                  int64_t n, int64_t step_count, float lr, float beta1, float beta2,
                  float eps, void *stream);

/* Same update with the step count t = *step_count_dev + 1 read on the device (bias
 * corrections computed in-kernel), then *step_count_dev is incremented by a trailing
 * 1-thread kernel: a captured CUDA graph of a whole training step replays correctly. */
int eav_adam_step_graph(float *params, const float *grads, float *exp_avg, float *exp_avg_sq,
                        int64_t n, int64_t *step_count_dev, float lr, float beta1, float beta2,
                        float eps, void *stream);

/* torch.renorm(p=2, dim=0, maxnorm) in place over `n_rows` rows of `row_len` floats
 * spaced `row_stride` apart (the max-norm hook body, EEGNet_tor.py:34,48). */
int eav_renorm_rows(float *w, int64_t n_rows, int64_t row_len, int64_t row_stride,
                    float maxnorm, void *stream);

/* ------------------------------------------------------------------------- */
/* Device-side epoch control: lets ONE CUDA graph hold a whole epoch of
 * Trainer_uni.train() (CNN_torch/EEGNet_tor.py:96-116: the shuffled batches of
 * DataLoader(shuffle=True), EEGNet_tor.py:92-93, incl. the ragged last batch) plus the
 * validation pass (EEGNet_tor.py:118-135), with no host work between steps.       */
/* ------------------------------------------------------------------------- */
/*
 * Fresh per-model permutation of the n_train training rows for the epoch *epoch_dev
 * (Philox4x32-10 keyed by (seed; subject id, epoch, element); NOT torch's CPU mt19937 stream --
 * parity runs keep the host DataLoader as the index source).
 * sched    device i32 [ceil(n_train/batch)][n_models*batch]: step s starts at s*n_models*batch and
 *          holds the x_index vector of eav_eegnet_forward for (n_models, B_s = min(batch, n_train - s*batch)),
 *          model-major; entries are absolute rows first_row + m*rows_per_model + perm_m[.].
 * subject_ids device i32 [n_models] or NULL (= 0..n_models-1): the random stream follows the subject,
 *          not the slot, so results do not depend on how subjects are sharded over GPUs.
 * epoch_dev device i64 epoch counter (read at run time; NULL = epoch 0).
 */
int eav_epoch_schedule(int32_t *sched, const int32_t *subject_ids, int32_t n_models, int32_t n_train,
                       int32_t batch, int64_t rows_per_model, int64_t first_row, uint64_t seed,
                       const int64_t *epoch_dev, void *stream);
/* acc device f64 [n_models][2]: acc[m][0] += loss[m]; acc[m][1] += n_correct[m] (n_correct may be NULL).
 * The running sums behind the per-epoch loss / accuracy the reference prints (EEGNet_tor.py:112-113,130-135). */
int eav_epoch_accumulate(const float *loss, const int32_t *n_correct, int32_t n_models, double *acc,
                         void *stream);
/* End of epoch e = *epoch_dev: history[e % max_epochs][m] = {mean train loss over n_train_steps, mean validation
 * loss over n_val_steps batches, validation accuracy = correct / n_val}; zeroes both accumulators; ++*epoch_dev.
 * history device f32 [max_epochs][n_models][3]. */
int eav_epoch_commit(double *train_acc, double *val_acc, int32_t n_models, int32_t n_train_steps,
                     int32_t n_val_steps, int32_t n_val, float *history, int32_t max_epochs,
                     int64_t *epoch_dev, void *stream);
/* When the validation pass of epoch e is pipelined with the training of epoch e+1 (it runs on a snapshot of the
 * parameters taken at the end of epoch e), eav_epoch_commit is called with val_acc == NULL and the validation columns
 * of history[(*epoch_dev + epoch_offset) % max_epochs] are filled one graph later by this call (epoch_offset = -1
 * after the next epoch's commit).  Zeroes val_acc; does not touch the epoch counter. */
int eav_epoch_commit_val(double *val_acc, int32_t n_models, int32_t n_val_steps, int32_t n_val, float *history,
                         int32_t max_epochs, const int64_t *epoch_dev, int32_t epoch_offset, void *stream);

/* ------------------------------------------------------------------------- */
/* ShallowConvNet of Transformer_torch/Transformer_EEG.py:107-148: Conv2d(1,40,(1,13)) -> 40
 * per-filter spatial Linear(30,1) -> n_layers single-head transformer layers (d = n_filters) -> BatchNorm2d ->
 * square -> AvgPool((1,pool), stride) -> log(clamp(., 1e-7, 1e4)) -> dropout -> Linear(F*U, n_classes, bias=False)
 * -> softmax.  One model, any batch; fp32.                                   */
/* ------------------------------------------------------------------------- */
typedef struct eav_shallow_cfg {
    int32_t batch, chans, samples;     /* B, 30, 500                                          */
    int32_t n_filters, kern;           /* 40 temporal filters of 13 taps (no padding)         */
    int32_t n_layers, ffn;             /* 12 transformer layers, hidden width 160             */
    int32_t pool, stride;              /* AvgPool window 35, stride 7                         */
    int32_t n_classes;
    int32_t bn_train;                  /* 1: batch statistics + running-stat update           */
    int32_t dropout_mode;              /* 0: none (eval); 1: caller supplies keep-masks       */
    float   dropout_p, bn_eps, bn_momentum, ln_eps;
} eav_shallow_cfg;

/* Parameters live in ONE flat fp32 array in the reference's named_parameters() order: conv.weight, bn.weight,
 * bn.bias, embedding.value_proj.{0..F-1}.weight, then per layer W_q W_k W_v ffn.net.0.{weight,bias}
 * ffn.net.3.{weight,bias} norm1.{weight,bias} norm2.{weight,bias}, then fc.weight.  offsets14 receives the offsets
 * of  conv bn.w bn.b emb | (layer 0) Wqkv W1 b1 W2 b2 g1 be1 g2 be2 | fc;  layer l adds l * *layer_stride.
 * Returns the parameter count. */
int64_t eav_shallow_param_layout(const eav_shallow_cfg *cfg, int64_t *offsets14, int64_t *layer_stride);
size_t eav_shallow_workspace_bytes(const eav_shallow_cfg *cfg);
/*
 * x        device [B][chans][samples] f32
 * bn_state device [2][n_filters] f32: running_mean, running_var (updated when bn_train)
 * masks    device u8 keep flags in forward call order (dropout_mode 1), else NULL:
 *          per layer  [B*T'][F] (after norm1), [B*T'][ffn] (inside the FFN), [B*T'][F] (after norm2);  then [B][F*U]
 * out      device [B][n_classes] probabilities;  workspace keeps the activations for eav_shallow_backward.
 */
int eav_shallow_forward(const eav_shallow_cfg *cfg, const float *x, const float *params, float *bn_state,
                        const uint8_t *masks, float *out, void *workspace, size_t workspace_bytes, void *stream);
/* Gradient of every parameter (grads: same layout as params, overwritten) given d(loss)/d(out). */
int eav_shallow_backward(const eav_shallow_cfg *cfg, const float *x, const float *params, const float *dout,
                         const uint8_t *masks, float *grads, void *workspace, size_t workspace_bytes, void *stream);

/* Measured-peak helper for bench.py: runs a register-resident FFMA loop on every SM
 * and returns the achieved fp32 TFLOP/s (host-synchronous; not part of the hot path). */
int eav_measure_fp32_peak(double *tflops, void *stream);
/* Same, for an 8x8 register outer product (three register operands per FFMA): the practical
 * ceiling of a register-blocked fp32 convolution/GEMM kernel on the CUDA cores. */
int eav_measure_fp32_peak_outer(double *tflops, void *stream);
/* mode 0: immediate form; 1: register outer product, scalar FFMA; 2: same with one operand in a uniform
 * register (constant bank); 3: same with Blackwell packed FFMA2 (fma.rn.f32x2). */
int eav_measure_fp32_peak_mode(int mode, double *tflops, void *stream);

/* ------------------------------------------------------------------------- */
/* Peer-memory all-reduce for the large-batch data-parallel driver (BASELINE configs[4]; replaces the
 * ncclAllReduce calls of the six BatchNorm statistic buffers, the loss and the flat gradient arena that
 * nn.DataParallel-style training of EEGNet_tor.py:86-88 would need).  One kernel per rank reads the peers' exchange
 * buffers over NVLink and sums them in rank order, so every rank gets bit-identical results.
 *   exchange buffer (per rank, mapped into every peer, ZERO-INITIALISED once): eav_peer_exchange_bytes(slot_bytes) bytes
 *   peer_bases_dev: device array [world] of the exchange buffers' addresses as seen from THIS rank
 *   call: 1, 2, 3, ... -- the same sequence on every rank; src/dst: n elements (float or double), may alias.
 *   call_base_dev (optional): device-resident step counter s; the effective call number is s * calls_per_step + call
 *   with call in 1..calls_per_step, so that a captured CUDA graph of a whole training step replays correctly.
 * All ranks must issue the same calls in the same order; a rank that never arrives makes the others trap after a
 * wall-time bound (120 s) instead of hanging. */
#define EAV_PEER_MAX_WORLD 16
#define EAV_PEER_MAX_CTAS 64
size_t eav_peer_exchange_bytes(size_t slot_bytes);
int eav_peer_allreduce(const void *src, void *dst, int64_t n, int is_f64, const uint64_t *peer_bases_dev,
                       int world, int rank, size_t slot_bytes, uint32_t call, const int64_t *call_base_dev,
                       uint32_t calls_per_step, void *stream);

/* Diagnostic for the tcgen05 path: runs reps x ksteps `tcgen05.mma.cta_group::1.kind::tf32` (M x N x 8 each) in one
 * CTA on a caller-supplied shared-memory image (image_floats fp32 words, <= 200 KB) with caller-supplied no-swizzle
 * operand descriptors {byte offset, LBO, SBO, major (0 = K, 1 = MN), byte advance per k-step}; a_bits / b_bits are
 * OR-ed into bits [32,64) of the A / B descriptor (swizzle mode in bits 61..63, base offset in 49..51; 0 = no swizzle).  Returns the fp32
 * accumulator tile d_out[128][N] (row = TMEM lane; of accumulator 0 when the MMAs rotate over n_acc > 1
 * independent TMEM column ranges) and the SM cycles from first issue to completion.  Used by
 * scripts/tc_probe.py and tests/test_gpu_tc.py to pin the operand address maps the tensor-core kernels rely on. */
int eav_tc_probe(const float *image_dev, int image_floats, int M, int N, int ksteps, int reps, int n_acc,
                 int a_off, int a_lbo, int a_sbo, int a_major, int a_step,
                 int b_off, int b_lbo, int b_sbo, int b_major, int b_step, int a_bits, int b_bits,
                 float *d_out_dev, long long *cycles_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EAV_B200_H_ */
