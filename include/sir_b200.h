/* sir_b200.h - C ABI of the B200-native speech-intent hot path (libsir_b200.so).
 *
 * The reference (avi2924/Speech-Intent-Recognizer) is pure Python: its "FFI" for this path is the set of
 * Python call sites that hand tensors to torchaudio / torch.nn.  Each entry point below names the reference
 * interface it stands in for (file:line under /root/reference).  The Python mirror of the reference classes
 * (speech-intent-recognizer_b200/scripts/*.py, models/models.py) binds these symbols with ctypes - see
 * INTEGRATION.md.
 *
 * Conventions
 *   - every pointer prefixed d_ is a DEVICE pointer owned by the caller (PyTorch allocates all tensors);
 *     the library owns only its handles (device constants, repacked weights, workspace);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it, nothing
 *     synchronises with the host except handle creation, weight upload and workspace growth;
 *   - return value 0 = success, negative = error; sir_last_error() returns a thread-local message;
 *   - a handle is bound to the device that was current when it was created and is not thread-safe on the host;
 *     it keeps its scratch per stream, so calls on up to 32 different streams may overlap on the device;
 *   - there is NO CPU fallback: every compute entry fails with SIR_ERR_CUDA if no CUDA device is usable.
 */
#ifndef SIR_B200_H
#define SIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIR_OK 0
#define SIR_ERR_INVALID (-1)   /* bad argument (message says which) */
#define SIR_ERR_CUDA (-2)      /* CUDA runtime error (message carries cudaGetErrorString) */
#define SIR_ERR_UNSUPPORTED (-3)

/* output selector of sir_frontend_forward */
#define SIR_OUT_MEL_POWER 0    /* MelSpectrogram only                    (mel_transform attribute)          */
#define SIR_OUT_MEL_DB 1       /* + AmplitudeToDB                        (amplitude_to_db(mel_transform(w))) */
#define SIR_OUT_LOGMEL_NORM 2  /* + per-utterance (x-mean)/(std+1e-5)    (extract_features)                  */
#define SIR_OUT_MFCC 3         /* dB with top_db floor + ortho DCT-II     (sir_frontend_mfcc only)           */

typedef struct sir_frontend sir_frontend;
typedef struct sir_model sir_model;
typedef struct sir_resampler sir_resampler;

const char* sir_last_error(void);
int sir_version(void);
/* Number of kernels this library has launched in the calling process (bench.py "gpu_launches"). */
int64_t sir_launch_count(void);

/* Per-stage device timing for bench.py / profiling (off by default): when enabled every kernel stage is
 * bracketed by CUDA events on the launching stream.  sir_profile_read synchronises the device, writes a
 * ';'-separated list of stage names plus total milliseconds and call counts per stage, clears the records and
 * returns the number of stages. */
void sir_profile_enable(int on);
int sir_profile_read(char* names, int names_cap, float* ms, int* calls, int max_stages);

/* ---- feature frontend ---------------------------------------------------------------------------------
 * sir_frontend_create  <->  AudioFeatureExtractor.__init__        scripts/precompute_features.py:21-36
 *                           (and the transform construction in    scripts/dataset.py:59-66)
 * Builds the periodic Hann window, FFT twiddles and the sparse HTK mel filterbank on the device.
 * Only n_fft = 1024, hop_length = 512 (every reference call site) and n_mels <= 128 are implemented;
 * anything else returns SIR_ERR_UNSUPPORTED.
 */
int sir_frontend_create(sir_frontend** out, int sample_rate, int n_mels, int n_fft, int hop_length);
void sir_frontend_destroy(sir_frontend* fe);

/* sir_frontend_forward  <->  the body of AudioFeatureExtractor.extract_features
 *                            scripts/precompute_features.py:59-73 (truncate, mel, dB, normalise),
 *                            its twins scripts/dataset.py:138-152 and scripts/test_model.py:78-94,
 *                            fused with SpecAugment masking (scripts/dataset.py:160-176 ==
 *                            scripts/augment.py:137-165) and pad/trim to a fixed frame count
 *                            (scripts/dataset.py:109-113, scripts/train.py:58-62), for a whole batch.
 *
 *   d_wave       [batch] rows of fp32 mono samples, row b starts at d_wave + b * wave_stride
 *   d_lengths    [batch] valid samples per row, or NULL (all rows have n_samples)
 *   n_samples    row capacity (>= every length)
 *   max_samples  truncate every utterance to this many samples first (int(max_duration*sr)); <= 0: none
 *   mode         SIR_OUT_*
 *   out_frames   time extent of the output rows.  Utterance b has T_b = 1 + L_b / 512 frames; frames
 *                beyond out_frames are dropped AFTER they took part in the normalisation statistics,
 *                missing frames are written as 0.0
 *   d_out        [batch, n_mels, out_frames] fp32
 *   d_masks      NULL or [batch, 4] int32 (t_start, t_end, f_start, f_end): bands of the UNPADDED feature
 *                map set to 0.0 after normalisation (only with SIR_OUT_LOGMEL_NORM)
 *   d_status     NULL or [batch] int32: 0 ok, 1 = utterance has <= n_fft/2 samples (reflect padding is
 *                undefined; the reference raises and returns None / zeros) -> row written as zeros
 */
int sir_frontend_forward(sir_frontend* fe, const float* d_wave, int64_t wave_stride, const int32_t* d_lengths,
                         int n_samples, int batch, int max_samples, int mode, int out_frames, float* d_out,
                         const int32_t* d_masks, int32_t* d_status, void* stream);

/* sir_frontend_mfcc / sir_preemphasis: the MFCC / pre-emphasis variant the north star names.  The reference itself
 * computes log-mel only; the semantics follow the library it builds on:
 * torchaudio.transforms.MFCC(sample_rate, n_mfcc, melkwargs={n_fft, hop_length, n_mels})
 *   = MelSpectrogram -> AmplitudeToDB("power", top_db) with the floor at the per-utterance maximum - top_db ->
 *     ortho DCT-II (create_dct)   TA:transforms/_transforms.py:634-718, TA:functional/functional.py:356-406,636-665
 * fused into the same kernel: dB values are staged, the maximum is merged across the cluster, every frame is clamped
 * and projected.  d_out [batch, n_mfcc, out_frames], zero beyond the utterance's frames; top_db <= 0 disables the floor.
 * sir_preemphasis: y[n] = x[n] - coeff x[n-1] per row (torchaudio.functional.preemphasis), out of place. */
int sir_frontend_mfcc(sir_frontend* fe, const float* d_wave, int64_t wave_stride, const int32_t* d_lengths, int n_samples,
                      int batch, int max_samples, int n_mfcc, float top_db, int out_frames, float* d_out, int32_t* d_status,
                      void* stream);
int sir_preemphasis(const float* d_in, float* d_out, int64_t stride, int n_samples, int batch, float coeff, void* stream);

/* sir_resampler_*  <->  torchaudio.transforms.Resample(orig_freq, new_freq) as the reference applies it before the
 *                       frontend to files that are not 16 kHz   scripts/precompute_features.py:54-56,
 *                       scripts/dataset.py:133-135, scripts/test_model.py:69-72
 * sinc_interp_hann polyphase FIR (lowpass_filter_width 6, rolloff 0.99 are torchaudio's defaults), taps built in
 * double precision.  Row b of d_out receives ceil(new * L_b / orig) samples (written to d_out_lengths[b] if given)
 * followed by zeros up to n_out. */
int sir_resampler_create(sir_resampler** out, int orig_freq, int new_freq, int lowpass_filter_width, double rolloff);
void sir_resampler_destroy(sir_resampler* r);
int64_t sir_resampler_output_length(const sir_resampler* r, int64_t n_in);
int sir_resampler_forward(sir_resampler* r, const float* d_in, int64_t in_stride, const int32_t* d_in_lengths, int n_in,
                          int batch, float* d_out, int64_t out_stride, int n_out, int32_t* d_out_lengths, void* stream);

/* sir_frontend_forward_pcm16: the same, ingesting 16-bit PCM rows (the payload of the WAV files the reference reads):
 * samples are scaled by 1/32768 on load - bit-identical to torchaudio.load's normalisation
 * (scripts/precompute_features.py:47, scripts/dataset.py:126) followed by sir_frontend_forward, with half the
 * HBM / PCIe bytes per utterance.  wave_stride and n_samples count samples. */
int sir_frontend_forward_pcm16(sir_frontend* fe, const int16_t* d_pcm, int64_t wave_stride, const int32_t* d_lengths,
                               int n_samples, int batch, int max_samples, int mode, int out_frames, float* d_out,
                               const int32_t* d_masks, int32_t* d_status, void* stream);

/* sir_amplitude_to_db  <->  AudioFeatureExtractor.amplitude_to_db  scripts/precompute_features.py:36,67
 * 10*log10(max(x, 1e-10)) elementwise over n values (in place allowed). */
int sir_amplitude_to_db(const float* d_in, float* d_out, int64_t n, void* stream);

/* sir_specaugment_sample  <->  the random draws of FSCIntentDataset.__getitem__/augment_features
 *                              scripts/dataset.py:105,166-171 and torchaudio mask_along_axis.
 * Counter-based Philox4x32-10 keyed on (seed, first_index + b): reproducible and independent of batch
 * split.  Writes [batch,4] int32 band parameters (empty bands where the gates say "no mask").
 *   d_frames   NULL or [batch] valid frame count per utterance (else n_frames for all)
 */
int sir_specaugment_sample(uint64_t seed, uint64_t first_index, int batch, int n_mels, int n_frames,
                           const int32_t* d_frames, float augment_prob, int time_mask_param,
                           int freq_mask_param, int32_t* d_masks, void* stream);

/* sir_features_finalize  <->  FSCIntentDataset.__getitem__ tail + collate_fn on CACHED features
 *                             scripts/dataset.py:105-113, scripts/train.py:49-70
 * Masks (optional) applied on the valid frames, then pad/trim from in_frames to out_frames. */
int sir_features_finalize(const float* d_in, int batch, int n_mels, int in_frames, const int32_t* d_frames,
                          const int32_t* d_masks, int out_frames, float* d_out, void* stream);

/* ---- classifier ---------------------------------------------------------------------------------------
 * sir_model_create  <->  CNNAudioGRU.__init__                      models/models.py:6-39
 * n_mels must be a multiple of 8 (GRU input = 128 * n_mels / 8; the reference hard-codes 1024 = 64 mels).
 */
int sir_model_create(sir_model** out, int num_classes, int n_mels);
void sir_model_destroy(sir_model* m);

/* sir_model_load_weights  <->  nn.Module.load_state_dict          scripts/evaluate.py:48, test_model.py:42
 * `weights` is ONE contiguous fp32 buffer (host or device pointer) holding the state_dict tensors in the
 * order of speech-intent-recognizer_b200/utils/synth.py:state_dict_spec (conv/bn 1..3, gru l0, l0_reverse,
 * l1, l1_reverse as weight_ih, weight_hh, bias_ih, bias_hh, attention, fc).  The buffer is copied into the
 * handle and repacked ON THE DEVICE (one kernel; BatchNorm folded for inference with eps = bn_eps), all ordered
 * on `stream`: a device source is never synchronised with the host. */
int sir_model_load_weights(sir_model* m, const float* weights, int64_t count, float bn_eps, void* stream);
int64_t sir_model_weight_count(const sir_model* m);

/* sir_model_forward  <->  CNNAudioGRU.forward (eval)              models/models.py:41-68
 *   d_features [batch, n_mels, n_frames] fp32 (the [B,1,n_mels,T] form has the same memory layout)
 *   d_logits   [batch, num_classes] fp32
 * n_frames >= 8. */
int sir_model_forward(sir_model* m, const float* d_features, int batch, int n_frames, float* d_logits,
                      void* stream);

/* sir_predict  <->  the evaluation head right after the logits:
 *                   scripts/test_model.py:121-156 (softmax, argmax, confidence, get_top_predictions k = 3) and
 *                   scripts/evaluate.py:79-98 (argmax, accuracy_score / confusion_matrix inputs)
 *   d_pred [batch] int32 = argmax (first maximum); d_conf [batch] = softmax probability of d_pred;
 *   d_topk_idx / d_topk_prob [batch, k] = argsort(probs)[::-1][:k] (ties: higher index first, as numpy does), k <= 8;
 *   with d_labels: d_correct[0] += #(pred == label), d_confusion[label * C + pred] += 1 (int64 counters, accumulated
 *   across calls; the caller zeroes them).  Any output pointer may be NULL. */
int sir_predict(const float* d_logits, int batch, int num_classes, int k, const int64_t* d_labels, int32_t* d_pred,
                float* d_conf, int32_t* d_topk_idx, float* d_topk_prob, int64_t* d_confusion, int64_t* d_correct, void* stream);

/* ---- training step --------------------------------------------------------------------------------------
 * The device work of scripts/train.py:80-116 (model.train() forward, CrossEntropyLoss, backward, Adam, GradScaler
 * unscale / inf-skip).  Parameters, gradients and Adam moments are FLAT fp32 device buffers owned by the caller,
 * in the same state_dict order as sir_model_load_weights (gradient slots of the BatchNorm running statistics are
 * written as zeros).  The data-parallel job all-reduces d_grads (one NCCL call) between sir_model_backward and
 * sir_adam_step.
 *
 * sir_model_train_forward  <->  CNNAudioGRU.forward in train mode      models/models.py:41-68 under model.train()
 *   BatchNorm uses batch statistics and updates running_mean / running_var inside d_params (momentum, unbiased
 *   variance; num_batches_tracked stays with the caller); the inter-layer GRU dropout (p = 0.5) keeps the elements
 *   of d_dropout_keep ([batch * n_frames/8 * 512] bytes, 1 = keep) or, when that is NULL, draws them from
 *   Philox4x32-10 keyed on (seed, offset).  Activations are kept inside the handle for sir_model_backward;
 *   d_features must stay alive until then.  n_frames % 8 == 0, n_mels == 64.
 */
int sir_model_train_forward(sir_model* m, float* d_params, const float* d_features, int batch, int n_frames,
                            const uint8_t* d_dropout_keep, uint64_t seed, uint64_t offset, float bn_momentum,
                            float bn_eps, float* d_logits, void* stream);

/* sir_model_backward  <->  loss.backward() through the model            scripts/train.py:99,107
 *   d_dlogits [batch, num_classes] -> d_grads [sir_model_weight_count] (overwritten, not accumulated). */
int sir_model_backward(sir_model* m, const float* d_params, const float* d_dlogits, float* d_grads, void* stream);

/* The same backward in two launches, so that a data-parallel job can all-reduce the first bucket of gradients while
 * the second is still being computed (what DistributedDataParallel's gradient buckets do behind scripts/train.py:107):
 *   part 1: attention + fc + both GRU layers.  Zeroes d_grads, then writes every gradient at or behind
 *           sir_model_gru_grad_offset() (96 % of the bytes); d_dlogits required.
 *   part 2: the three conv blocks.  Writes the gradients in front of that offset; d_dlogits may be NULL.
 * part 1 followed by part 2 on one stream is sir_model_backward. */
int sir_model_backward_part(sir_model* m, const float* d_params, const float* d_dlogits, float* d_grads, int part, void* stream);
int sir_model_gru_grad_offset(const sir_model* m, int64_t* offset);

/* sir_cross_entropy  <->  nn.CrossEntropyLoss()(output, label) + its backward    scripts/train.py:96,106,243
 *   d_loss[0] = mean_b( logsumexp(logits_b) - logits_b[label_b] );
 *   d_dlogits (may be NULL) = scale * (softmax(logits) - onehot(label)) / n_valid (scale = the GradScaler loss scale).
 *   Labels: -100 (torch's ignore_index) rows add neither loss nor gradient and leave the mean (n_valid counts the
 *   rest); any other label outside [0, num_classes) masks its row and returns d_loss[0] = NaN (torch asserts). */
int sir_cross_entropy(const float* d_logits, const int64_t* d_labels, int batch, int num_classes, float scale,
                      float* d_loss, float* d_dlogits, void* stream);

/* sir_grad_nonfinite: *d_flag = 1.0f if any of the `count` gradients is inf/nan (left untouched otherwise)
 *   <->  the inf check of GradScaler.unscale_/step   scripts/train.py:100 */
int sir_grad_nonfinite(const float* d_grads, int64_t count, float* d_flag, void* stream);

/* sir_adam_step  <->  scaler.step(optimizer) with optim.Adam(lr, weight_decay)   scripts/train.py:100,246-250
 *   torch.optim.Adam semantics (coupled L2: g = grad * inv_scale + weight_decay * p, bias-corrected moments, no
 *   amsgrad) over `n_segments` (<= 4) ranges of the flat buffers, `segments` = HOST array of (offset, count)
 *   pairs - the parameter ranges between the BatchNorm running statistics.  The whole update is skipped when
 *   *d_found_inf != 0 (d_found_inf may be NULL). */
int sir_adam_step(float* d_params, const float* d_grads, float* d_exp_avg, float* d_exp_avg_sq, const int64_t* segments,
                  int n_segments, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                  float inv_scale, const float* d_found_inf, void* stream);

/* Staged forward for host-resident batches (the public entry that overlaps the H2D copy of sub-batch i+1 with
 * the conv stack of sub-batch i): sir_model_forward_convs runs conv1..3 (+BN+ReLU+pool) for utterances
 * [first, first + count) of a batch of batch_total utterances, d_features pointing at utterance `first`;
 * sir_model_forward_head then runs the GRU layers, attention pooling and fc (models/models.py:60-67) over all
 * batch_total utterances.  Together they equal sir_model_forward; batch_total <= 336 (one workspace pass). */
int sir_model_forward_convs(sir_model* m, const float* d_features, int batch_total, int first, int count, int n_frames,
                            void* stream);
int sir_model_forward_head(sir_model* m, int batch_total, int n_frames, float* d_logits, void* stream);

/* sir_gemm_nt_split_f16: C[M,N] = A[M,K] W[N,K]^T + bias[N] on the 5th-gen tensor cores (tcgen05, TMEM
 * accumulators, TMA-staged operands) with the 3-pass fp16 hi/lo operand split that keeps fp32-level accuracy.
 * It is the contraction behind nn.GRU's input projection (models/models.py:60) and, in implicit-GEMM form,
 * conv2/conv3 (:51-52); exported stand-alone so that tests can check it against an fp64 matmul.
 * N % 128 == 0 and K % 64 == 0.  All pointers are device pointers to fp32. */
int sir_gemm_nt_split_f16(const float* d_a, const float* d_w, const float* d_bias, float* d_c, int M, int N, int K,
                          void* stream);

/* sir_gemm_tile_width: the activation-row tile width (UMMA N; one of 160, 176, 208, 256) the persistent kernel behind
 * sir_gemm_nt_split_f16 uses for M rows and N output columns on a device of `sms` SMs: the width whose 128 x width
 * tiles need the fewest rounds x width over the SMs.  Pure host arithmetic (no device call); exported so that the
 * choice can be tested without a GPU.  Returns 0 when the shape takes the non-persistent kernel (< 32 tiles). */
int sir_gemm_tile_width(int M, int N, int sms);

/* sir_conv3x3_nhwc_split_f16: 3x3 convolution (stride 1, zero padding 1, no bias) of channels-last fp32 tensors,
 * d_in [B,H,W,C_in] * d_w [9 taps][C_out][C_in] -> d_out [B,H,W,C_out], as the implicit GEMM on tcgen05 behind
 * conv2 / conv3 (models/models.py:12-15) and behind their data gradients; (C_in, C_out) in
 * {(32,64), (64,128), (128,64), (64,32)}.  Exported stand-alone so that tests can check it against conv2d. */
int sir_conv3x3_nhwc_split_f16(const float* d_in, const float* d_w, float* d_out, int B, int H, int W, int cin, int cout,
                               void* stream);

/* ---- device-resident step state: the training step as a replayable CUDA graph ------------------------------------
 * scripts/train.py:80-116 passes values that change every step through the host: the Adam step count (bias corrections),
 * the GradScaler scale, the dropout stream position.  With them in a 32-byte DEVICE struct (d_state: any 8-byte aligned
 * device allocation of SIR_TRAIN_STATE_BYTES) the whole step takes no per-step host argument and can be captured once
 * with cudaStreamBeginCapture / torch.cuda.graph and replayed:
 *     sir_train_state_begin      bias corrections of step + 1, inv_scale = 1 / (loss_scale * world)
 *     sir_model_train_forward    (dropout offset read from the state once sir_model_set_train_state was called)
 *     sir_cross_entropy_state    loss scale read from the state
 *     sir_model_backward, sir_grad_nonfinite, [gradient all-reduce]
 *     sir_adam_step_state        corrections / inv_scale read from the state; skipped when *d_found_inf != 0
 *     sir_train_state_end        step += (*d_found_inf == 0); dropout offset += increment
 * sir_train_state_set_scale is the host's GradScaler growth / back-off between replays (synchronises the stream). */
#define SIR_TRAIN_STATE_BYTES 32
int sir_train_state_init(void* d_state, float loss_scale, int step, uint64_t dropout_offset, void* stream);
int sir_train_state_set_scale(void* d_state, float loss_scale, void* stream);
int sir_train_state_begin(void* d_state, float beta1, float beta2, int world, void* stream);
int sir_train_state_end(void* d_state, const float* d_found_inf, uint64_t dropout_offset_increment, void* stream);
int sir_model_set_train_state(sir_model* m, const void* d_state);
int sir_cross_entropy_state(const float* d_logits, const int64_t* d_labels, int batch, int num_classes, const void* d_state,
                            float* d_loss, float* d_dlogits, void* stream);
int sir_adam_step_state(float* d_params, const float* d_grads, float* d_exp_avg, float* d_exp_avg_sq, const int64_t* segments,
                        int n_segments, float lr, float beta1, float beta2, float eps, float weight_decay, const void* d_state,
                        const float* d_found_inf, void* stream);

/* sir_pipeline_forward: frontend (SIR_OUT_LOGMEL_NORM, pad/trim to out_frames) + classifier in one call.
 * d_features [batch, n_mels, out_frames] is required (the classifier reads the features from it). */
int sir_pipeline_forward(sir_frontend* fe, sir_model* m, const float* d_wave, int64_t wave_stride,
                         const int32_t* d_lengths, int n_samples, int batch, int max_samples, int out_frames,
                         float* d_features, float* d_logits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIR_B200_H */
