/*
 * after_b200.h -- C ABI of libafter_b200.so, the B200 (sm_100a) implementation of AFTER's
 * latent-sampling hot path.
 *
 * The reference (acids-ircam/AFTER) is pure Python/PyTorch and has no FFI of its own; each entry
 * point below names the reference *Python* call it replaces (file:line under /root/reference) --
 * that is the interface a binding (ctypes here, a TORCH_LIBRARY op or an nn_tilde backend
 * elsewhere) forwards to.  See INTEGRATION.md for the reference-side stubs.
 *
 * Conventions
 *   - plain C types only; tensors are raw pointers to contiguous fp32 (unless stated) with the
 *     reference's layouts: audio (B,1,S), latents (B,C,T) channel-first, cond (B,zt),
 *     time_cond (B,zs,T).
 *   - "dev" pointers are device pointers on the handle's GPU, "host" pointers are host memory.
 *   - every function returns 0 on success, a negative AFTER_E* code on failure;
 *     after_last_error(h) returns a human-readable message for the last failure on that handle
 *     (after_last_error(NULL): last failure of a call that had no handle).
 *   - a handle is NOT re-entrant: one handle per (device, stream); the caller serialises.
 *   - all work is enqueued on the cudaStream_t passed as `void* stream` (NULL = legacy default
 *     stream); functions do not synchronise unless stated.
 *   - the library owns its weight copies and workspace; it never retains or frees caller
 *     pointers.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef AFTER_B200_H_
#define AFTER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFTER_B200_ABI_VERSION 3

/* error codes */
#define AFTER_OK 0
#define AFTER_IGNORED 1          /* after_load_tensor: key is not a parameter this path uses */
#define AFTER_EINVAL (-1)        /* bad argument / unsupported configuration */
#define AFTER_ECUDA (-2)         /* CUDA runtime / driver error */
#define AFTER_ESTATE (-3)        /* call order violated (e.g. compute before finalize) */
#define AFTER_EMISSING (-4)      /* finalize: a required tensor was never loaded */
#define AFTER_ESHAPE (-5)        /* tensor shape does not match the configuration */
#define AFTER_ENOMEM (-6)

/* arithmetic modes (after_finalize_weights) */
#define AFTER_PRECISION_FP32 0   /* tcgen05 bf16x3 split products, fp32 accumulate: fp32-accurate */
#define AFTER_PRECISION_BF16 1   /* tcgen05 single bf16 product, fp32 accumulate */
#define AFTER_PRECISION_FP32_SIMT 2 /* fp32 FFMA everywhere (validation path, no tensor cores) */

/* tensor dtypes accepted by after_load_tensor */
#define AFTER_DTYPE_F32 0
#define AFTER_DTYPE_F64 1
#define AFTER_DTYPE_I64 2

/* which sub-model a tensor belongs to (state-dict key namespaces are per module in the reference) */
#define AFTER_MODULE_DENOISER 0          /* RectifiedFlow.net        : DenoiserV2 state_dict   */
#define AFTER_MODULE_AUTOENCODER 1       /* emb_model                : AutoEncoder state_dict  */
#define AFTER_MODULE_STRUCTURE_ENCODER 2 /* RectifiedFlow.encoder_time: Encoder1D state_dict   */
#define AFTER_MODULE_TIMBRE_ENCODER 3    /* RectifiedFlow.encoder    : ECAPATDNN state_dict    */
#define AFTER_MODULE_UNET 4              /* RectifiedFlow.net        : UNET1D state_dict (instead of DenoiserV2) */
#define AFTER_MODULE_LATENT_MAP 5        /* Streamer.project_model   : SmallAutoencoder state_dict (optional) */

/* classifier-free-guidance row layouts */
#define AFTER_CFG_AUDIO 0 /* (cond,tc)/(drop,tc)/(drop,drop), f = g_t/max(g_s,clamp)  model.py:730-759 */
#define AFTER_CFG_MIDI 1  /* (cond,tc)/(cond,drop)/(drop,drop), f = g_s/max(g_t,clamp) export_midi.py:322-360 */

#define AFTER_MAX_STAGES 8

/* Hyper-parameters; mirrors the gin bindings of the reference (base.gin:65-78, baseAE.gin:35-52).
 * A field set to 0 in the optional sub-models disables that sub-model. */
typedef struct after_config {
  int32_t abi_version; /* must be AFTER_B200_ABI_VERSION */
  /* --- DenoiserV2 (transformerv2.py:463-476) --- */
  int32_t n_channels;           /* latent channels (64) */
  int32_t seq_len;              /* maximum frames per sequence the workspace is sized for (256) */
  int32_t embed_dim;            /* 256 | 512 */
  int32_t cond_dim;             /* timbre dims (6) */
  int32_t noise_embed_dims;     /* fourier features (64) */
  int32_t n_layers;             /* 6 */
  int32_t mlp_multiplier;       /* 3 */
  int32_t tcond_dim;            /* structure dims (12, midi 128) */
  int32_t local_attention_size; /* 8 (midi 16) */
  int32_t attention_chunk_size; /* 4 */
  float drop_value;             /* -4.0 (base.gin:88) */
  int32_t max_batch;            /* largest B (streams) a call may pass; CFG rows = 3B */
  int32_t max_steps;            /* largest nb_steps for after_sample */
  /* --- AutoEncoder (SimpleNetsStream.py:834-849); ae_channels == 0 disables the codec --- */
  int32_t ae_in_channels;  /* 16 */
  int32_t ae_channels;     /* 64 */
  int32_t ae_z_channels;   /* 64 */
  int32_t ae_pqmf_bands;   /* 16 */
  int32_t ae_n_stages;     /* len(factors) */
  int32_t ae_multipliers[AFTER_MAX_STAGES + 1]; /* encoder multipliers */
  int32_t ae_dec_multipliers[AFTER_MAX_STAGES + 1]; /* int(m*decoder_ratio) of reversed multipliers */
  int32_t ae_factors[AFTER_MAX_STAGES];
  int32_t ae_dilations[AFTER_MAX_STAGES]; /* first ae_num_blocks entries used */
  int32_t ae_num_blocks;   /* 3 */
  int32_t ae_kernel_size;  /* 3 */
  int32_t ae_use_loudness; /* 1 */
  int64_t ae_max_samples;  /* longest audio (samples) per stream the codec workspace is sized for */
  /* --- Encoder1D structure encoder (encoder.py:116-237); se_n_blocks == 0 disables --- */
  int32_t se_in_size;
  int32_t se_n_blocks;                     /* len(channels) */
  int32_t se_channels[AFTER_MAX_STAGES];
  int32_t se_kernel_size;                  /* 5 */
  int32_t se_causal;                       /* 1: convs.get_padding.mode = 'causal' (base.gin:55) */
  int32_t se_use_tanh;
  /* --- ECAPATDNN timbre encoder (ecapa_encoder.py:458-566); te_n_blocks == 0 disables --- */
  int32_t te_in_size;
  int32_t te_n_blocks;                       /* len(channels) (4) */
  int32_t te_channels[AFTER_MAX_STAGES];     /* [512, 512, 512, 1024] */
  int32_t te_kernel_sizes[AFTER_MAX_STAGES]; /* [3, 3, 3, 3] */
  int32_t te_dilations[AFTER_MAX_STAGES];    /* [1, 1, 1, 1] */
  int32_t te_res2net_scale;                  /* 8 */
  int32_t te_se_channels;                    /* 128 */
  int32_t te_attention_channels;             /* 128 */
  int32_t te_out_dim;                        /* zt = 6 */
  int32_t te_global_context;                 /* 1 */
  int32_t te_use_tanh;
  /* --- streaming (transformerv2.py:119-155; after_scripts/export.py:74-79 binds it to LOCAL_ATTENTION_SIZE) --- */
  int32_t max_cache_size; /* frames of key/value history per (layer, diffusion step, sequence); 0 = offline only.
                             One history per cache_index in [0, max_steps) and per sequence in [0, 3*max_batch). */
  /* --- UNET1D conv denoiser (unet1d.py:255-268), the other `net` RectifiedFlow can bind; un_n_levels == 0 disables.
   *     seq_len / max_batch / max_steps / drop_value above size and parameterise it as they do DenoiserV2. --- */
  int32_t un_in_size;
  int32_t un_out_size;                    /* 0: same as un_in_size */
  int32_t un_n_levels;                    /* len(channels) */
  int32_t un_channels[AFTER_MAX_STAGES];
  int32_t un_ratios[AFTER_MAX_STAGES];    /* the reference's `ratios` (a leading 1 is prepended); un_n_levels - 1 entries used */
  int32_t un_kernel_size;                 /* odd, <= 7 */
  int32_t un_time_channels;               /* SPE dim */
  int32_t un_time_cond_in_channels;
  int32_t un_time_cond_channels;          /* 0: time_cond is concatenated to the input instead of embedded per scale */
  int32_t un_cond_channels;               /* 0: no global condition */
  int32_t un_n_attn_layers;
  int32_t un_use_res_last;
  /* --- streaming codec / structure encoder: what the exported models compute buffer by buffer
   *     (after_scripts/export_autoencoder.py:16-153, 305-319; after_scripts/export.py:14-17, 418-435).  stream_slots
   *     independent states (the exported Streamer holds two codec copies: structure and timbre); 0 = offline only. --- */
  int32_t stream_slots;
  int32_t stream_max_frames; /* latent frames per streaming call the state is sized for (0: 64) */
  int32_t stream_gn_frames;  /* CachedGroupNorm padding_size in latent frames = length of the export script's first call
                                (131072 samples = 64 frames, export_autoencoder.py:49-51; 0: 64) */
} after_config;

typedef struct after_ctx* after_handle;

/* Library / device discovery (no handle needed). */
int after_abi_version(void);
const char* after_build_info(void); /* arch, compiler, build flags */
int after_device_count(void);

/* Create a handle on CUDA device `device`.  Replaces constructing the reference modules
 * (RectifiedFlow(net=DenoiserV2(...), ...) model.py:20-50; AutoEncoder(...) SimpleNetsStream.py:834). */
int after_create(const after_config* cfg, int device, after_handle* out);
int after_destroy(after_handle h);
const char* after_last_error(after_handle h);

/* Feed one entry of a reference state_dict (torch.nn.Module.load_state_dict, model.py:221-247;
 * keys as listed in SURVEY.md appendix A.4).  `data` is HOST memory, C-contiguous.  Weight-norm
 * pairs (weight_g / weight_v), eval-mode BatchNorm statistics and GroupNorm affines are folded
 * inside the library at finalize.  Returns AFTER_IGNORED for keys the path does not need
 * (caches, position tables). */
int after_load_tensor(after_handle h, int module, const char* key, const void* data,
                      const int64_t* shape, int ndim, int dtype);

/* Fold / transpose / split the loaded weights, upload them, build TMA descriptors, allocate the
 * workspace.  Must be called once after the last after_load_tensor and before any compute call. */
int after_finalize_weights(after_handle h, int precision);

/* DenoiserV2.forward (transformerv2.py:517-543): out = net(x, time, cond, time_cond).
 * x, out: dev (N,C,T); time: dev (N,) ; cond: dev (N,zt); time_cond: dev (N,zs,T).  N <= 3*max_batch. */
int after_denoiser_forward(after_handle h, const float* x, const float* time, const float* cond,
                           const float* time_cond, float* out, int N, int T, void* stream);

/* UNET1D.forward (unet1d.py:376-429): out = net(x, time=, time_cond=, cond=) for a handle that carries UNET1D weights
 * (AFTER_MODULE_UNET).  x dev (N,in_size,T); time dev (N,); cond dev (N,cond_channels) or NULL when the net has no global
 * condition; time_cond dev (N,time_cond_in_channels,T) or NULL when it has none; out dev (N,out_size,T).
 * T must be a multiple of the product of the ratios. */
int after_unet_forward(after_handle h, const float* x, const float* time, const float* cond, const float* time_cond,
                       float* out, int N, int T, void* stream);

/* RectifiedFlow.model_forward (model.py:721-761): one CFG-combined velocity evaluation.  Runs over whichever `net` the
 * handle carries (DenoiserV2, or UNET1D when only AFTER_MODULE_UNET tensors were loaded); same for after_sample*.
 * x, out: dev (B,C,T); time: dev (B,); cond dev (B,zt); time_cond dev (B,zs,T). */
int after_model_forward(after_handle h, const float* x, const float* time, const float* cond,
                        const float* time_cond, float* out, int B, int T, float guidance_timbre,
                        float guidance_structure, int cfg_variant, float clamp, void* stream);

/* RectifiedFlow.sample (model.py:763-785): nb_steps Euler steps from x0.  x0,out dev (B,C,T). */
int after_sample(after_handle h, const float* x0, const float* cond, const float* time_cond,
                 float* out, int B, int T, int nb_steps, float guidance_timbre,
                 float guidance_structure, int cfg_variant, float clamp, void* stream);

/* Same as after_sample with HOST buffers (pageable or pinned): H2D of the inputs, the loop, D2H of
 * the result, and a stream synchronise before returning. */
int after_sample_host(after_handle h, const float* x0, const float* cond, const float* time_cond,
                      float* out, int B, int T, int nb_steps, float guidance_timbre,
                      float guidance_structure, int cfg_variant, float clamp, void* stream);

/* ---- streaming denoiser: per-diffusion-step rolling key/value caches (needs after_config.max_cache_size > 0) ----
 * after_denoiser_forward_cached: DenoiserV2.forward(..., cache_index) with MHAttention.max_cache_size > 0
 *   (transformerv2.py:190-236, 517-543): the block's keys/values are appended to the history of `cache_index`
 *   for attention (history cached un-rotated, queries rotated at offset = history length), and remembered as
 *   last_k / last_v.  The history starts as zeros, exactly like the reference's registered buffers.
 * after_model_forward_cached: RectifiedFlow.model_forward(..., cache_index) (model.py:721-761; exported
 *   Streamer.model_forward, after_scripts/export.py:356-396).
 * after_roll_cache: DenoiserV2.roll_cache(size, cache_index) (transformerv2.py:167-186, 433-435, 514-515).
 * after_reset_cache: zero every history (what re-instantiating the reference module does).
 * after_sample_stream: one audio block of the exported Streamer.sample (export.py:398-416): for step i,
 *   x += model_forward(x, t_i, ..., cache_index=i) / nb_steps, then roll_cache(T, i).  x_last,out dev (B,C,T). */
int after_denoiser_forward_cached(after_handle h, const float* x, const float* time, const float* cond,
                                  const float* time_cond, float* out, int N, int T, int cache_index, void* stream);
int after_model_forward_cached(after_handle h, const float* x, const float* time, const float* cond,
                               const float* time_cond, float* out, int B, int T, float guidance_timbre,
                               float guidance_structure, int cfg_variant, float clamp, int cache_index, void* stream);
int after_roll_cache(after_handle h, int roll_size, int cache_index, void* stream);
int after_reset_cache(after_handle h, void* stream);
int after_sample_stream(after_handle h, const float* x_last, const float* cond, const float* time_cond, float* out,
                        int B, int T, int nb_steps, float guidance_timbre, float guidance_structure, int cfg_variant,
                        float clamp, void* stream);

/* AutoEncoder.encode (SimpleNetsStream.py:918-941; z only, as export_autoencoder.py:251-258):
 * audio dev (B,1,S) -> z dev (B,Z,S/ratio).  S must be a multiple of the codec ratio. */
int after_ae_encode(after_handle h, const float* audio, float* z, int B, int64_t samples, void* stream);

/* AutoEncoder.decode (SimpleNetsStream.py:943-954): z dev (B,Z,T) -> audio dev (B,1,T*ratio). */
int after_ae_decode(after_handle h, const float* z, float* audio, int B, int T, void* stream);

/* ---- streaming codec and structure encoder (needs after_config.stream_slots > 0) ----
 * What `export_stream.ts` of a non-causal AutoEncoder computes per buffer (export_autoencoder.py AE_notcausal):
 * after_ae_encode_stream: offline PQMF of the buffer, then the encoder with cached convolutions (every conv keeps the last
 *   l + r + stride-delay frames of its input; residual branches delayed to match) and CachedGroupNorm(stream=True)
 *   (statistics over the previous stream_gn_frames + this buffer); the latents lag the offline encoder by its cumulative
 *   delay (8 frames for baseAE).  audio dev (B,1,S) -> z dev (B,Z,S/ratio).
 * after_ae_decode_stream: offline decoder over [z_buffer ; z] with CachedGroupNorm(stream=True), linear cross-fade of the
 *   first 4 latent frames with the tail kept from the previous call (export_autoencoder.py:128-153).  z dev (B,Z,T>=4) ->
 *   audio dev (B,1,T*ratio).
 * after_structure_encode_stream: Encoder1D.forward_stream (encoder.py:300-322) with cached convolutions.
 * after_stream_reset: state of a freshly constructed model (zero caches) for one slot.
 * cached_conv (acids-ircam/cached_conv >= 2.5.0) is an un-vendored dependency of the reference; its semantics are restated
 * in oracle/after_oracle_stream.py. */
int after_ae_encode_stream(after_handle h, int slot, const float* audio, float* z, int B, int64_t samples, void* stream);
int after_ae_decode_stream(after_handle h, int slot, const float* z, float* audio, int B, int T, void* stream);
int after_structure_encode_stream(after_handle h, int slot, const float* z, float* time_cond, int B, int T, void* stream);
int after_stream_reset(after_handle h, int slot, void* stream);

/* Encoder1D.forward (encoder.py:273-298): z dev (B,C,T) -> time_cond dev (B,zs,T). */
int after_structure_encode(after_handle h, const float* z, float* time_cond, int B, int T, void* stream);

/* ECAPATDNN.forward (ecapa_encoder.py:567-624), the timbre encoder: z dev (B,C,T) -> cond dev (B,zt). */
int after_timbre_encode(after_handle h, const float* z, float* cond, int B, int T, void* stream);

/* Streamer.latent2map / map2latent of the exported model (after_scripts/export.py:494-508): the input is averaged over
 * time, sent through the encoder (direction 0, latent2map) or decoder (direction 1, map2latent) half of the export-time
 * projection (after/diffusion/latent_plot.py:20-37, loaded as AFTER_MODULE_LATENT_MAP: encoder.{0,2,4}.{weight,bias},
 * decoder.{0,2,4}.{weight,bias}) and repeated over the buffer.  A handle without those tensors applies the identity the
 * reference exports with --nolatent_project (export.py:143).  x dev (B,C_in,T) -> out dev (B,*C_out,T); out must hold
 * B * 64 * T floats at most (layers are at most 64 wide); *C_out (may be NULL) receives the output channel count. */
int after_latent_map(after_handle h, int direction, const float* x, float* out, int B, int C_in, int T, int* C_out,
                     void* stream);

/* The whole audio-to-audio chain of the notebooks (notebooks/audio_to_audio_demo.ipynb cells 5/19):
 *   z_s = encode(audio_structure); z_t = encode(audio_timbre); time_cond = encoder_time(z_s); cond = encoder(z_t);
 *   x = sample(x0, cond, time_cond, nb_steps, g_t, g_s); audio_out = decode(x)
 * audio_* dev (B,1,S), x0 dev (B,C,S/ratio) (the prior noise is an INPUT so results are reproducible), out dev (B,1,S).
 * Needs denoiser, autoencoder, structure- and timbre-encoder weights on the handle.  after_generate_host is the same
 * with HOST buffers (H2D of the two audio batches and x0, D2H of the audio, stream synchronised before returning). */
int after_generate(after_handle h, const float* audio_structure, const float* audio_timbre, const float* x0,
                   float* audio_out, int B, int64_t samples, int nb_steps, float guidance_timbre,
                   float guidance_structure, void* stream);
int after_generate_host(after_handle h, const float* audio_structure, const float* audio_timbre, const float* x0,
                        float* audio_out, int B, int64_t samples, int nb_steps, float guidance_timbre,
                        float guidance_structure, void* stream);

/* Introspection for benchmarks: kernels launched by the library in this process (all handles; graph replays count
 * their captured kernels), bytes of device memory held by this handle, codec ratio (samples per latent frame). */
int64_t after_launch_count(after_handle h);
int64_t after_device_bytes(after_handle h);
int after_ae_ratio(after_handle h);

/* Per-kernel-class device timing for roofline reports (no reference counterpart: the reference has no
 * profiler).  While enabled every launch of the profiled classes is bracketed by CUDA
 * events on the library's work stream and CUDA-graph replay is bypassed, so the same kernels run one by one.
 * after_profile_read synchronises the device and returns, for one class, the number of launches since
 * after_profile_enable(h, 1), their summed duration (ms) and their summed algorithmic flops / bytes. */
#define AFTER_KERNEL_TAP_GEMM_TC 0   /* tcgen05 tap-GEMM (linears + convolutions), except the fused MLP */
#define AFTER_KERNEL_TAP_GEMM_SIMT 1 /* fp32 FFMA tap-GEMM */
#define AFTER_KERNEL_ATTENTION 2     /* banded attention + residual + AdaLN-c + LN3 */
#define AFTER_KERNEL_ROW_NORM 3      /* AdaLN-t + LN1 */
#define AFTER_KERNEL_ACT_OPERAND 4   /* GroupNorm/BatchNorm + Snake/SiLU operand pass */
#define AFTER_KERNEL_PQMF 5          /* PQMF analysis / synthesis */
#define AFTER_KERNEL_OTHER 6         /* unclassified */
#define AFTER_KERNEL_MLP_FUSED 7     /* fused MLP (up + down projection in one persistent tcgen05 launch) */
int after_profile_enable(after_handle h, int on);
int after_profile_read(after_handle h, int kernel_class, int64_t* launches, double* ms, double* flops,
                       double* bytes);

/* Unit-level entry point used by the parity tests of the tensor-core GEMM:
 * C[M,N] = A[M,K] * W[N,K]^T (+bias[N]) in the given precision; all dev fp32. */
int after_debug_gemm(after_handle h, const float* A, const float* W, const float* bias, float* C,
                     int M, int N, int K, int precision, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AFTER_B200_H_ */
