/*
 * mixgan_b200.h — C ABI of the B200-native MixGAN-TTS diffusion-decoder reverse process.
 *
 * The reference (MaxMax2016/MixGAN-TTS) is pure Python and has no FFI layer; the
 * boundary it exposes for this path is two torch modules.  Every entry point below
 * names the reference method it replaces (paths relative to the reference root):
 *
 *   mgb_denoiser_forward   Denoiser.forward                      model/modules.py:420-446
 *                          (ResidualBlock.forward                model/blocks.py:1157-1176,
 *                           DiffusionEmbedding/Mish/LinearNorm   model/blocks.py:906-913,894-896,278-291)
 *   mgb_reverse_step       GaussianDiffusion.p_sample            model/diffusion.py:121-129
 *                          (+ q_posterior / q_posterior_sample   model/diffusion.py:104-119,
 *                             extract                            model/diffusion.py:26-29)
 *   mgb_sample             GaussianDiffusion.sampling + the tail of .forward (inference)
 *                                                                model/diffusion.py:155-165,200
 *   mgb_shallow_start      GaussianDiffusion.diffuse_fn * mask   model/diffusion.py:177-185,198-199
 *   mgb_denorm_mask        GaussianDiffusion.denorm_spec * mask  model/diffusion.py:231-232,164,200
 *   mgb_denoiser_train_forward / mgb_denoiser_backward
 *                          autograd through Denoiser.forward     model/modules.py:420-446 (train.py:126-184)
 *   mgb_length_regulate    LengthRegulator.LR / expand / pad     model/linguistic_encoder.py:383-416,
 *                          get_mask_from_lengths                 utils/tools.py:144-153,374-392
 *   mgb_durations_from_log duration rounding at inference        model/linguistic_encoder.py:310-314
 *   mgb_conv1d_forward / mgb_conv1d_backward / mgb_step_embedding
 *                          JCUDiscriminator.forward + autograd   model/mixgantts.py:186-288 (train.py:126-184)
 *   mgb_auxdec_forward     decoder + mel_linear + postnet        model/mixgantts.py:139-143
 *                          (Decoder transformer/Models.py:103-171, PostNet transformer/Layers.py:67-137)
 *   mgb_hifigan_forward    hifigan Generator.forward             hifigan/models.py:151-166 (utils/model.py:103-121)
 *
 * Conventions
 *   - Plain pointers and sizes only; no torch types.  All tensors are DEVICE pointers to
 *     densely packed row-major arrays unless a parameter says "host".
 *   - The caller owns every buffer (inputs, outputs, packed weights, workspace).  The library
 *     never allocates or frees caller-visible memory and never synchronises the device; all work
 *     is enqueued on `stream` (a cudaStream_t passed as void*).  Two training entry points
 *     (mgb_denoiser_backward in the bf16 mode, mgb_conv1d_backward) run the half of their work
 *     that nothing downstream waits for on a library-owned side stream (one non-blocking stream
 *     and a few events per device, created at first use): it is forked from and joined back to
 *     `stream` by events INSIDE the call, so the caller still sees plain stream semantics (and a
 *     CUDA-graph capture of `stream` records the fork / join as edges).
 *   - Return value: 0 = OK, negative = error (MGB_E_*); mgb_last_error() returns a thread-local
 *     message.  Nothing throws across the ABI.
 *   - There is no CPU fallback: on a device that is not sm_100 every compute entry point
 *     returns MGB_E_ARCH.
 *
 * Tensor layouts (fp32 unless noted)
 *   x_t, noise, x_prev, out : [B][n_mel][T]            (the reference's [B,1,M,T])
 *   cond                    : [B][T][d_encoder]        (as it arrives at GaussianDiffusion.forward)
 *   spk                     : [B][d_encoder] or NULL   (required when dims.multi_speaker)
 *   t                       : int64 [B]
 *   pad_mask                : uint8 [B][T], 1 = padding (reference convention at .forward) or NULL
 *   mel / coarse            : [B][T][n_mel]
 *   sched                   : float [3][K] = posterior_mean_coef1 | posterior_mean_coef2 |
 *                             sigma, with sigma[t] = (t == 0) ? 0 : exp(0.5*posterior_log_variance_clipped[t])
 */
#ifndef MIXGAN_B200_H_
#define MIXGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGB_ABI_VERSION 2

enum {
  MGB_OK = 0,
  MGB_E_ARG = -1,       /* bad argument (NULL, shape, unsupported dims) */
  MGB_E_ARCH = -2,      /* device is not sm_100 / library built without sm_100a code */
  MGB_E_WORKSPACE = -3, /* workspace or packed buffer too small */
  MGB_E_CUDA = -4,      /* a CUDA call failed; see mgb_last_error() */
  MGB_E_UNSUPPORTED = -5
};

/* Arithmetic of the convolution GEMMs. */
enum {
  MGB_PREC_FP32 = 0, /* fp32 operands and accumulation on the CUDA cores (parity mode)           */
  MGB_PREC_BF16 = 1, /* bf16 operands on tcgen05 tensor cores, fp32 accumulation in TMEM,
                        fp32 residual stream, fp32 posterior update (throughput mode)           */
  MGB_PREC_FP16 = 3  /* reference-precision mode ON THE TENSOR CORES: the same fused tcgen05 kernel with fp16
                        operands (11-bit significand = TF32's; kind::f16 runs fp16 and bf16 at one rate), fp32
                        accumulation, fp32 residual / skip streams and spills, biases as fp16 hi+lo pairs and an
                        ex2-based gate.  Meets the fp32 bar of 1e-3 relative L2 (inference entry points only). */
};
/* mgb_pack_weights / mgb_packed_bytes only: the MGB_PREC_FP32 buffer (same size and offsets) with ONLY the per-utterance
 * table weights filled (step MLP, diffusion / speaker projections, conditioner bias) in one launch — what the bf16
 * training forward needs as `packed`, re-packed every step. */
#define MGB_PACK_FP32_TABLES 2

typedef struct mgb_model_dims {
  int32_t n_mel;         /* 80  */
  int32_t channels;      /* 256 residual channels */
  int32_t d_encoder;     /* 256 conditioner width */
  int32_t layers;        /* 20  */
  int32_t multi_speaker; /* 0/1: per-block bias-free speaker projection present */
} mgb_model_dims;

int mgb_abi_version(void);
const char* mgb_last_error(void);

/* Instrumentation for bench.py.  mgb_launch_count: kernels this library has enqueued in this
 * process so far.  mgb_profile_enable(1) makes the library bracket every launch of the path's
 * dominant kernel with CUDA events on the launching stream; mgb_profile_collect synchronises
 * those events, returns their summed duration and count, and clears them. */
long long mgb_launch_count(void);
/* A host that replays a captured CUDA graph of this library's launches reports the replayed launches here. */
void mgb_note_launches(long long n);
void mgb_profile_enable(int on);
int mgb_profile_collect(float* total_ms, int* count);
/* mgb_profile_enable(2): IN-SITU timing instead — each of the first 256 launches of the fused tensor-core kernel after the
 * enable records min(start) / max(end) over its CTAs in %globaltimer nanoseconds, and the SM cycles / nanoseconds of its
 * first CTA, into a fixed region of the caller's workspace (two atomics per CTA, no extra launches, no events).
 * mgb_profile_read_stamps reads the region back from that workspace after the caller has synchronised the stream (this
 * instrumentation call does a blocking device-to-host copy): summed launch duration, number of launches recorded, their
 * min / max, and the SM clock the kernels really ran at (cycles / ns; each of the last three may be NULL). */
int mgb_profile_read_stamps(const void* workspace, float* total_ms, int* count, float* min_ms, float* max_ms,
                            float* sm_mhz);

/* Debug: synchronously read the watchdog word of the bf16 path from a workspace that was used with
 * the same (B, T).  0 = every kernel completed its barrier protocol; non-zero bits name the warp
 * role whose bounded wait expired (1 producer, 2 MMA issuer, 4 epilogue). */
int mgb_debug_status(const mgb_model_dims* dims, int precision, int B, int T, const void* workspace,
                     int* host_status);

/* 0 when `device` can run this library (compute capability 10.x), else MGB_E_ARCH. */
int mgb_device_check(int device);

/*
 * Weights.  `flat` is ONE fp32 device array holding the Denoiser's parameters in state_dict
 * layout ([out][in][k] conv weights, [out][in] linear weights) concatenated in this order:
 *   input_projection.0.conv.{weight,bias}, mlp.0.linear.weight, mlp.2.linear.weight,
 *   for each layer l: conv_layer.conv.{weight,bias}, diffusion_projection.linear.weight,
 *                     [speaker_projection.linear.weight,] conditioner_projection.conv.{weight,bias},
 *                     output_projection.conv.{weight,bias},
 *   skip_projection.conv.{weight,bias}, output_projection.conv.{weight,bias}.
 * mgb_pack_weights rewrites them once (per load_state_dict) into the kernels' layouts.
 */
size_t mgb_flat_weight_count(const mgb_model_dims* dims);
size_t mgb_packed_bytes(const mgb_model_dims* dims, int precision);
int mgb_pack_weights(const mgb_model_dims* dims, int precision, const float* flat,
                     void* packed, size_t packed_bytes, void* stream);

size_t mgb_workspace_bytes(const mgb_model_dims* dims, int precision, int B, int T, int K);

/* Denoiser.forward: out = x0 prediction, [B][n_mel][T]. */
int mgb_denoiser_forward(const mgb_model_dims* dims, int precision, const void* packed,
                         const float* x, const int64_t* t, const float* cond, const float* spk,
                         float* out, int B, int T, void* workspace, size_t workspace_bytes,
                         void* stream);

/* p_sample: x0 = Denoiser(x_t, t, cond, spk); clamp to [-1,1] if clip;
 * x_prev = coef1[t]*x0 + coef2[t]*x_t + sigma[t]*noise.  x0_out may be NULL. */
int mgb_reverse_step(const mgb_model_dims* dims, int precision, const void* packed,
                     const float* x_t, const int64_t* t, const float* cond, const float* spk,
                     const float* noise, const float* sched, int K, int clip,
                     float* x_prev, float* x0_out, int B, int T,
                     void* workspace, size_t workspace_bytes, void* stream);

/* sampling: K reverse steps from x_T with noises[K][B][n_mel][T] (noises[t] used at timestep t),
 * then mel_out[B][T][n_mel] = denorm_spec(x_0) * (1 - pad_mask).
 * states_out (optional): [K+1][B][T][n_mel] denormalised states, as sampling() returns them.
 * x0_norm_out (optional): [B][n_mel][T], the last step's normalised x_0 (== state K before denorm). */
int mgb_sample(const mgb_model_dims* dims, int precision, const void* packed,
               const float* x_T, const float* cond, const float* spk, const float* noises,
               const float* sched, int K, int clip, const float* spec_min, const float* spec_max,
               const uint8_t* pad_mask, float* states_out, float* mel_out, float* x0_norm_out,
               int B, int T, void* workspace, size_t workspace_bytes, void* stream);

/* The tensor-core paths read the conditioner as a 16-bit image on the batch row axis (8-channel chunks, zero rows between
 * utterances).  Every sampling / Denoiser call builds it itself; this entry runs only that HBM-bound conversion (fp32
 * [B][T][d_encoder] -> image inside `workspace`), so that a host can time it — bench.py reports its GB/s. */
int mgb_pack_cond(const mgb_model_dims* dims, int precision, const float* cond, int B, int T, void* workspace,
                  size_t workspace_bytes, void* stream);

/* x_T[B][n_mel][T] = (sqrt_acp * norm_spec(coarse)^T + sqrt_1m_acp * noise) * (1 - pad_mask). */
int mgb_shallow_start(const float* coarse, const float* noise, const float* spec_min,
                      const float* spec_max, float sqrt_acp, float sqrt_1m_acp,
                      const uint8_t* pad_mask, float* x_T, int B, int T, int n_mel, void* stream);

/* mel[B][T][n_mel] = denorm_spec(x[B][n_mel][T]^T) * (1 - pad_mask). */
int mgb_denorm_mask(const float* x, const float* spec_min, const float* spec_max,
                    const uint8_t* pad_mask, float* mel, int B, int T, int n_mel, void* stream);

/*
 * The elementwise steps of GaussianDiffusion.forward's training branch around the Denoiser call (model/diffusion.py:201-225),
 * fused (the torch composition is ~56 launches per call).  All states are fp32 [B][n_mel][T]; `t` int64 [B] in [0, K).
 *
 * mgb_train_diffuse             :206-207  x_t = diffuse_fn(mel, t) * valid and x_t_prev = diffuse_fn(mel, t - 1) * valid, with
 *                               diffuse_fn = norm_spec + transpose + q_sample and x_start itself where t - 1 < 0 (:177-185);
 *                               mel [B][T][n_mel]; sqrt_acp / sqrt_1m_acp = the two q_sample tables, float [K]
 * mgb_train_posterior           :210-212, :220 of the 'naive' model: x0_pred = clamp(denoiser_out * valid) (clamp if `clip`),
 *                               x_t_prev_pred = (coef1[t] x0_pred + coef2[t] x_t + sigma[t] noise) * valid; `sched` as in
 *                               mgb_reverse_step (float [3][K]: coef1 | coef2 | sigma)
 * mgb_train_posterior_backward  d/d denoiser_out of both outputs (either incoming gradient may be NULL = zero)
 */
int mgb_train_diffuse(const float* mel, const float* noise_t, const float* noise_prev, const float* spec_min,
                      const float* spec_max, const float* sqrt_acp, const float* sqrt_1m_acp, const int64_t* t,
                      const uint8_t* pad_mask, float* x_t, float* x_t_prev, int B, int T, int n_mel, int K, void* stream);
int mgb_train_posterior(const float* denoiser_out, const float* x_t, const float* noise, const float* sched, const int64_t* t,
                        const uint8_t* pad_mask, int clip, float* x0_pred, float* x_t_prev_pred, int B, int T, int n_mel,
                        int K, void* stream);
int mgb_train_posterior_backward(const float* grad_x0_pred, const float* grad_x_t_prev_pred, const float* denoiser_out,
                                 const float* sched, const int64_t* t, const uint8_t* pad_mask, int clip,
                                 float* grad_denoiser_out, int B, int T, int n_mel, int K, void* stream);

/*
 * Duration indexing of the front end (all integer results are bit-exact).
 *
 * mgb_durations_from_log   the word-level duration rounding of LinguisticEncoder.forward at inference
 *                          (model/linguistic_encoder.py:310-314):
 *                          dur[i] = (int64) max(round_half_even(exp(log_d[i]) - 1) * d_control, 0)   (.long() truncates)
 * mgb_length_regulate      LengthRegulator.LR / expand (linguistic_encoder.py:383-416) + pad (utils/tools.py:374-392):
 *                          x[B][S][D] fp32, dur int64 [B][S]; out[B][max_len][D]: each source row s repeated
 *                          max(dur[b][s], 0) times, zero padded; an utterance longer than max_len is CROPPED, exactly as
 *                          the reference (F.pad with a negative size crops); mel_len[b] = sum_s max(dur[b][s], 0) is the
 *                          TRUE length either way (int64).  mask_valid (optional, uint8 [B][max_len]) receives
 *                          get_mask_from_lengths(mel_len, max_len) (utils/tools.py:144-153: 1 = valid frame).
 *                          Needs B*(S+1) int64 of scratch in `workspace`; the scratch (the exclusive scan of the
 *                          durations) is what mgb_length_regulate_backward reads.
 * mgb_length_regulate_backward   d/dx of the expansion (the autograd of expand + cat + pad in the reference):
 *                          grad_x[b][s] = sum of grad_out[b][f] over the frames copied from row s, in frame order.
 * mgb_mask_from_lengths    get_mask_from_lengths for lengths that did not come out of mgb_length_regulate.
 */
int mgb_durations_from_log(const float* log_d, float d_control, int64_t* dur, int n, void* stream);
int mgb_length_regulate(const float* x, const int64_t* dur, float* out, int64_t* mel_len, uint8_t* mask_valid,
                        int B, int S, int D, int max_len, void* workspace, size_t workspace_bytes,
                        void* stream);
int mgb_length_regulate_backward(const float* grad_out, const void* workspace, float* grad_x, int B, int S, int D,
                                 int max_len, void* stream);
int mgb_mask_from_lengths(const int64_t* lengths, uint8_t* mask_valid, int B, int max_len, void* stream);

/*
 * Generic fp32 Conv1d on frames-major activations, forward and backward: the building block of the JCU discriminator
 * (reference model/mixgantts.py:186-288 - ConvNorm model/blocks.py:326-371 with stride 1 or 2 and
 * padding (k-1)/2, LinearNorm (k = 1), F.leaky_relu(., 0.2), Mish - driven 4x forward + 2x backward per training step by
 * train.py:126-184).  fp32 accuracy throughout — the reference's gradients are the parity target (1e-4): forward and data
 * gradient run on the tensor cores as 3 x TF32 (hi / lo operand split, fp32 accumulation; channel counts that are multiples of
 * 32), everything else, and the weight gradient, as exact-fp32 CUDA-core kernels.
 *   x [B][Tin][Cin], y / pre / grad_y [B][Tout][Cout] with Tout = mgb_conv1d_out_len(Tin, k, stride); one row per frame.
 *   w       : [Cout][Cin][k], the torch layout (a Linear weight [out][in] is k = 1); bias [Cout] or NULL
 *   rowbias : [B][Cin] or NULL - added to every EXISTING input row before the convolution (padding rows stay zero):
 *             the discriminator's "x + diffusion_step (+ speaker)" (mixgantts.py:275-276) fused into the operand gather
 *   act     : 0 none, 1 leaky_relu(0.2), 2 mish (needs `pre`, the pre-activation, for its backward), 3 relu
 *   backward: grad_x [B][Tin][Cin], grad_w [Cout][Cin][k], grad_bias [Cout], grad_rowbias [B][Cin]; each may be NULL
 *             (grad_rowbias needs grad_x).  Gradients are WRITTEN, not accumulated; the weight gradient is reduced in a
 *             fixed order (deterministic).
 *   workspace: mgb_conv1d_workspace_bytes(...) bytes of scratch, not kept between calls.
 * mgb_step_embedding: DiffusionEmbedding (model/blocks.py:899-913), emb [B][dim] = [sin(t f_i) | cos(t f_i)].
 */
int mgb_conv1d_out_len(int Tin, int k, int stride);
size_t mgb_conv1d_workspace_bytes(int B, int Tin, int Cin, int Cout, int k, int stride);
int mgb_conv1d_forward(const float* x, const float* w, const float* bias, const float* rowbias, float* y, float* pre,
                       int B, int Tin, int Cin, int Cout, int k, int stride, int act,
                       void* workspace, size_t workspace_bytes, void* stream);
int mgb_conv1d_backward(const float* x, const float* w, const float* rowbias, const float* y, const float* pre,
                        const float* grad_y, float* grad_x, float* grad_w, float* grad_bias, float* grad_rowbias,
                        int B, int Tin, int Cin, int Cout, int k, int stride, int act,
                        void* workspace, size_t workspace_bytes, void* stream);
int mgb_step_embedding(const int64_t* t, float* emb, int B, int dim, void* stream);

/*
 * Training (BASELINE configs[4]): Denoiser forward that keeps the activations its backward needs, and the
 * backward itself.  Replaces torch autograd through Denoiser.forward (model/modules.py:420-446,
 * model/blocks.py:1157-1176) as train.py:126-184 drives it.  MGB_PREC_FP32: fp32 GEMMs on the CUDA cores (parity mode);
 * MGB_PREC_BF16: every GEMM on tcgen05 (bf16 operands, fp32 accumulation, fp32 residual/skip/gradient streams).
 *   packed      : mgb_pack_weights(MGB_PREC_FP32) image of the CURRENT parameters (both precisions: the per-utterance
 *                 step MLP and projection tables are fp32)
 *   flat        : the same parameters in the canonical flat order (mgb_pack_weights' input)
 *   saved       : caller-owned activation stash, mgb_train_saved_bytes(dims, precision, B, T) bytes, written by the forward
 *   workspace   : mgb_train_workspace_bytes(dims, precision, B, T) bytes, shared by forward and backward
 *   grad_out    : d loss / d out, [B][n_mel][T]
 *   grad_flat   : d loss / d parameters in the canonical flat order (WRITTEN, not accumulated)
 *   grad_cond   : [B][T][d_encoder] or NULL;  grad_spk: [B][d_encoder] or NULL;  grad_x: [B][n_mel][T] or NULL
 * The backward is a sequence of mgb_train_segments(dims) = layers + 2 segments: 0 = output/skip projections,
 * 1..layers = residual blocks from the last to the first, layers + 1 = input projection + step MLP.  A call runs
 * segments [seg_begin, seg_end) and must be issued in order on one stream; after segment s the slice
 * [flat_begin, flat_end) that mgb_train_segment_range reports is final, so the host can start the NCCL
 * all-reduce of that gradient bucket while later segments run (grad_cond/grad_spk/grad_x are final after the last).
 */
size_t mgb_train_saved_bytes(const mgb_model_dims* dims, int precision, int B, int T);
size_t mgb_train_workspace_bytes(const mgb_model_dims* dims, int precision, int B, int T);
int mgb_train_segments(const mgb_model_dims* dims);
int mgb_train_segment_range(const mgb_model_dims* dims, int seg, size_t* flat_begin, size_t* flat_end);
int mgb_denoiser_train_forward(const mgb_model_dims* dims, int precision, const void* packed, const float* flat,
                               const float* x, const int64_t* t, const float* cond, const float* spk, float* out,
                               void* saved, size_t saved_bytes, int B, int T,
                               void* workspace, size_t workspace_bytes, void* stream);
int mgb_denoiser_backward(const mgb_model_dims* dims, int precision, const float* flat, const void* saved,
                          size_t saved_bytes, const int64_t* t, const float* cond, const float* spk,
                          const float* grad_out, float* grad_flat, float* grad_cond, float* grad_spk,
                          float* grad_x, int B, int T, int seg_begin, int seg_end,
                          void* workspace, size_t workspace_bytes, void* stream);
/* Debug: synchronously read the watchdog word of the bf16 training kernels from a workspace used with (B, T). */
int mgb_train_debug_status(const mgb_model_dims* dims, int B, int T, const void* workspace, int* host_status);

/*
 * Aux decoder (second half): the FastSpeech2 decoder + mel_linear + PostNet that produce the coarse
 * mel the shallow reverse diffusion starts from.  Replaces, at inference (model/mixgantts.py:139-143),
 *     coarse = mel_linear(decoder(output, mel_masks));  coarse = postnet(coarse) + coarse
 * i.e. Decoder.forward transformer/Models.py:137-171, FFTBlock transformer/Layers.py:11-31, MultiHeadAttention /
 * PositionwiseFeedForward transformer/SubLayers.py:8-97, ScaledDotProductAttention transformer/Modules.py:6-24,
 * PostNet transformer/Layers.py:67-137 (BatchNorm1d in eval mode).  Every matrix product runs on tcgen05 with fp16
 * operands and fp32 accumulation; residual sums, LayerNorm and softmax statistics are fp32.
 *   flat   : ONE fp32 device array of the parameters in this order (state_dict order without position_enc and
 *            num_batches_tracked):  for each FFT block: slf_attn.{w_qs,w_ks,w_vs}.{weight,bias}, slf_attn.layer_norm.{weight,
 *            bias}, slf_attn.fc.{weight,bias}, pos_ffn.w_1.{weight,bias}, pos_ffn.w_2.{weight,bias}, pos_ffn.layer_norm.{weight,
 *            bias};  mel_linear.{weight,bias};  for each PostNet layer: conv.{weight,bias}, batch norm {weight, bias,
 *            running_mean, running_var}.
 *   x      : [B][T][d_model] decoder input (the variance adaptor's output), pos : [T][d_model] position encoding rows
 *            (Decoder.position_enc[0, :T], or the sinusoid table for T > max_seq_len as Models.py:146-153)
 *   lens   : int32 [B] valid frames per utterance or NULL (= T): keys at or beyond it are masked, rows at or beyond it are
 *            zero-filled after every sub-layer (the reference's mel_masks)
 *   coarse : [B][T][n_mel] = postnet(mel) + mel;  dec_out [B][T][d_model] and mel_before [B][T][n_mel] are optional.
 */
typedef struct mgb_auxdec_dims {
  int32_t n_mel;          /* 80   */
  int32_t d_model;        /* 256  decoder_hidden */
  int32_t n_head;         /* 2    decoder_head   */
  int32_t d_inner;        /* 1024 conv_filter_size */
  int32_t ffn_kernel;     /* 9    conv_kernel_size[0] (the second FFN convolution is k = 1) */
  int32_t layers;         /* 6    decoder_layer */
  int32_t postnet_dim;    /* 512  */
  int32_t postnet_kernel; /* 5    */
  int32_t postnet_layers; /* 5    */
} mgb_auxdec_dims;
size_t mgb_auxdec_flat_count(const mgb_auxdec_dims* dims);
size_t mgb_auxdec_packed_bytes(const mgb_auxdec_dims* dims);
size_t mgb_auxdec_workspace_bytes(const mgb_auxdec_dims* dims, int B, int T);
int mgb_auxdec_pack(const mgb_auxdec_dims* dims, const float* flat, void* packed, size_t packed_bytes, void* stream);
int mgb_auxdec_forward(const mgb_auxdec_dims* dims, const void* packed, const float* x, const float* pos, const int32_t* lens,
                       float* coarse, float* dec_out, float* mel_before, int B, int T, void* workspace,
                       size_t workspace_bytes, void* stream);
int mgb_auxdec_debug_status(const mgb_auxdec_dims* dims, int B, int T, const void* workspace, int* host_status);

/*
 * HiFi-GAN generator: mel -> waveform after the diffusion decoder.  Replaces Generator.forward
 * hifigan/models.py:151-166 (ResBlock.forward :96-103) as utils/model.py:103-121 (vocoder_infer) calls it, weight norm
 * removed (utils/model.py:99).  tcgen05 implicit-GEMM convolutions, fp16 operands, fp32 accumulation and residual sums.
 *   flat : fp32 parameters in state_dict order after remove_weight_norm: conv_pre.{weight,bias}, ups.i.{weight,bias} for all
 *          i, resblocks.r.{convs1.0,convs1.1,convs1.2,convs2.0,convs2.1,convs2.2}.{weight,bias} for all r, conv_post.{weight,bias}
 *   mel  : [B][T][n_mel] frames-major (what GaussianDiffusion.forward returns; the reference transposes to [B][n_mel][T]
 *          before the call), wav : [B][T * mgb_hifigan_hop(dims)]
 */
typedef struct mgb_hifigan_dims {
  int32_t n_mel;                /* 80  */
  int32_t initial_channel;      /* 512 upsample_initial_channel */
  int32_t n_up;                 /* 4   */
  int32_t up_rates[8];          /* 8, 8, 2, 2 */
  int32_t up_kernels[8];        /* 16, 16, 4, 4 */
  int32_t n_res;                /* 3 resblocks per stage ("resblock": "1") */
  int32_t res_kernels[4];       /* 3, 7, 11 */
  int32_t res_dilations[4][3];  /* 1, 3, 5 each */
  int32_t split_mode;           /* operand precision: 0 = fp16 operands everywhere (2^-11 per operand);
                                   1 = fp16 hi + lo operand pairs (three MMAs per product, ~2^-21) in conv_pre, the
                                       transposed convolutions and conv_post (< 5 % of the FLOPs, ~40 % of the error): default;
                                   2 = hi + lo pairs in every layer: the parity mode (about 2.5x the time)                    */
} mgb_hifigan_dims;
size_t mgb_hifigan_flat_count(const mgb_hifigan_dims* dims);
size_t mgb_hifigan_packed_bytes(const mgb_hifigan_dims* dims);
int mgb_hifigan_hop(const mgb_hifigan_dims* dims);
size_t mgb_hifigan_workspace_bytes(const mgb_hifigan_dims* dims, int B, int T);
int mgb_hifigan_pack(const mgb_hifigan_dims* dims, const float* flat, void* packed, size_t packed_bytes, void* stream);
int mgb_hifigan_forward(const mgb_hifigan_dims* dims, const void* packed, const float* mel, float* wav, int B, int T,
                        void* workspace, size_t workspace_bytes, void* stream);
int mgb_hifigan_debug_status(const mgb_hifigan_dims* dims, int B, int T, const void* workspace, int* host_status);

#ifdef __cplusplus
}
#endif
#endif /* MIXGAN_B200_H_ */
