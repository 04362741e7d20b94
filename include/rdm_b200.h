/* librdm_b200 -- C ABI of the B200-native retrieval-augmented diffusion sampling hot path.
 *
 * The reference (CompVis/retrieval-augmented-diffusion-models @ 1017f2b) has no FFI of its own: its
 * boundary is YAML `target:` dotted paths resolved to Python classes.  Every
 * entry point below therefore cites the Python call site it replaces; the Python classes at the same
 * import paths (retrieval-augmented-diffusion-models_b200/rdm/...) bind these symbols via ctypes.
 *
 * Conventions: plain pointers and sizes only; all `*_dev` pointers are device pointers on the handle's
 * device, contiguous, 16-byte aligned, owned by the caller; work is stream-ordered on `stream`
 * (a cudaStream_t passed as void*; NULL = legacy default stream) with no hidden synchronisation unless
 * stated.  Every function returns 0 on success or a negative code and records a message retrievable
 * with rdm_last_error() (thread-local).  Handles are per-device and not thread-safe.
 */
#ifndef RDM_B200_H
#define RDM_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RDM_API __attribute__((visibility("default")))
#else
#define RDM_API
#endif

#define RDM_DTYPE_F32 0
#define RDM_DTYPE_F16 1

RDM_API const char* rdm_last_error(void);
/* ABI version of this header (bumped on any signature change). */
RDM_API int rdm_abi_version(void);
/* Number of kernel launches issued by this library so far in this process (bench.py `gpu_launches`). */
RDM_API unsigned long long rdm_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Retrieval: exact cosine top-k over the in-HBM CLIP database.
 * Replaces the ScaNN searcher the reference builds in
 *   rdm/data/retrieval_dataset/dsetbuilder.py:534-619 (train_searcher; DB rows L2-normalised at :574)
 * and queries at
 *   rdm/data/retrieval_dataset/dsetbuilder.py:490, rdm/models/diffusion/ddpm.py:298,906-908,
 *   rdm/models/autoregression/transformer.py:327-329, rdm/data/base.py:81-83
 *   (`searcher.search_batched(q_hat, final_num_neighbors=k) -> (indices, distances)`).
 * Result definition (bit-exact, batching/sharding independent): oracle/knn_ref.c.
 * ------------------------------------------------------------------------------------------------ */
typedef struct rdm_knn rdm_knn_t;

/* Create a searcher over `n` rows of dimension `d` (d in {256,512,768,1024}).  `db` is the RAW
 * (un-normalised) embedding matrix, row-major, dtype RDM_DTYPE_F16 or RDM_DTYPE_F32.
 * db_on_device != 0: `db` is a device pointer that the caller keeps alive for the handle's lifetime
 *                    (no copy; 20.9M x 512 fp16 = 21.4 GB stays resident once).
 * db_on_device == 0: `db` is a host pointer; the rows are copied to HBM in pinned-staged chunks.
 * `idx_base` is added to every reported index (row-sharded databases: rank r passes its first row).
 * Computes the per-row inverse L2 norms on the device (one pass over the DB) and synchronises. */
RDM_API int rdm_knn_create(rdm_knn_t** out, const void* db, int64_t n, int32_t d, int32_t dtype,
                   int32_t db_on_device, int64_t idx_base, int32_t device);
RDM_API void rdm_knn_destroy(rdm_knn_t* h);
RDM_API int64_t rdm_knn_size(const rdm_knn_t* h);
/* Copies the float32 inverse L2 norms [n] computed at create time into out_dev (device). */
RDM_API int rdm_knn_get_inv_norms(rdm_knn_t* h, float* out_dev, void* stream);

/* search_batched: q_hat_dev float32 [nq, d], ALREADY L2-normalised by the caller exactly as the
 * reference does (numpy fp32, ddpm.py:907).  1 <= k <= RDM_KNN_MAX_K.
 * Outputs (device): idx_out int64 [nq,k] (global indices, best first), dist_out float32 [nq,k]
 * (= (float)score), score_out float64 [nq,k] or NULL (exact scores, used by the shard merge). */
#define RDM_KNN_MAX_K 24
RDM_API int rdm_knn_search(rdm_knn_t* h, const float* q_hat_dev, int32_t nq, int32_t k,
                   int64_t* idx_out_dev, float* dist_out_dev, double* score_out_dev, void* stream);

/* The caller-side query normalisation of the reference, `q / np.linalg.norm(q, axis=1)[:, np.newaxis]` (ddpm.py:297,907;
 * dsetbuilder.py:487; base.py:82), for float32 rows [nq, d] on the device: bit-identical to NumPy's fp32 result (separately rounded
 * squares, NumPy's pairwise summation order, IEEE sqrt and division; restated in oracle/knn.py pairwise_sum_f32). */
RDM_API int rdm_knn_normalize(const float* q_dev, int32_t nq, int32_t d, float* q_hat_out_dev, int32_t device, void* stream);
/* rdm_knn_normalize + rdm_knn_search in one call: q_raw_dev float32 [nq, d] are RAW query embeddings (DB rows or CLIP outputs), i.e.
 * exactly what the reference holds before ddpm.py:907 -- no eager arithmetic is left between get_qids and the scan. */
RDM_API int rdm_knn_search_raw(rdm_knn_t* h, const float* q_raw_dev, int32_t nq, int32_t k,
                   int64_t* idx_out_dev, float* dist_out_dev, double* score_out_dev, void* stream);

/* Merge `parts` per-shard results (e.g. the all_gather of every rank's rdm_knn_search output):
 * idx_in int64 [parts, nq, k], score_in float64 [parts, nq, k] -> global top-k by (score desc, idx asc). */
RDM_API int rdm_knn_merge(const int64_t* idx_in_dev, const double* score_in_dev, int32_t parts, int32_t nq, int32_t k,
                  int64_t* idx_out_dev, float* dist_out_dev, double* score_out_dev, int32_t device, void* stream);

/* `data_pool['embedding'][nns]` then `.to(device).to(float)` (ddpm.py:921, dsetbuilder.py:493):
 * gathers RAW rows as float32.  idx_dev holds GLOBAL indices; rows outside this shard
 * ([idx_base, idx_base+n)) are written as zeros so shards can be summed. out_dev float32 [count, d]. */
RDM_API int rdm_knn_gather(rdm_knn_t* h, const int64_t* idx_dev, int64_t count, float* out_dev, void* stream);


/* ------------------------------------------------------------------------------------------------
 * U-Net eps-model: rdm/modules/diffusionmodules/openaimodel.py:66-317 (UNetModel.__init__),
 * :335-371 (forward) with rdm/modules/attention.py:122-196 SpatialTransformers cross-attending the
 * k retrieved CLIP vectors.  The struct mirrors the constructor arguments the shipped configs set
 * (models/rdm/imagenet/config.yaml:36-59).
 * ------------------------------------------------------------------------------------------------ */
typedef struct rdm_unet rdm_unet_t;
typedef struct rdm_unet_cfg {
    int32_t in_channels, model_channels, out_channels, num_res_blocks;
    int32_t n_attention_resolutions; int32_t attention_resolutions[8];   /* downsample rates with attention */
    int32_t n_channel_mult; int32_t channel_mult[8];
    int32_t num_head_channels;      /* d_head; must be 32 */
    int32_t num_heads;              /* used only when num_head_channels <= 0 */
    int32_t transformer_depth;      /* must be 1 */
    int32_t context_dim;            /* 512 (CLIP) */
} rdm_unet_cfg;

#define RDM_UNET_MODE_FP32 0        /* every contraction in fp32 on CUDA cores (strict parity mode) */
#define RDM_UNET_MODE_TC_BF16X3 1   /* tcgen05: operands split into bf16 hi+lo, 3 MMAs per product (fp32-grade, default for parity) */
#define RDM_UNET_MODE_TC_BF16 2     /* tcgen05: plain bf16 operands, fp32 accumulation (~6e-3 on DDIM-100 latents: outside the tolerance) */
#define RDM_UNET_MODE_TC_FP16X2 3   /* tcgen05: fp16 activations x (fp16 hi + lo) weights, 2 MMAs per product */
#define RDM_UNET_MODE_TC_FP16 4     /* tcgen05: plain fp16 operands (11-bit significand), 1 MMA per product */

RDM_API int rdm_unet_create(rdm_unet_t** out, const rdm_unet_cfg* cfg, int32_t device);
RDM_API void rdm_unet_destroy(rdm_unet_t* h);
/* State-dict interface (`model.load_state_dict`, scripts/rdm_sample.py:170): parameter names are the
 * reference's UNetModel keys, e.g. "input_blocks.4.1.transformer_blocks.0.attn2.to_k.weight". */
RDM_API int64_t rdm_unet_num_params(const rdm_unet_t* h);
RDM_API const char* rdm_unet_param_name(const rdm_unet_t* h, int64_t i);
RDM_API int64_t rdm_unet_param_numel(const rdm_unet_t* h, const char* name);     /* -1 if unknown */
/* host float32, PyTorch layout of that parameter (conv [Cout,Cin,kh,kw], linear [out,in]); re-packed on upload. */
RDM_API int rdm_unet_load(rdm_unet_t* h, const char* name, const float* host, int64_t numel);
RDM_API int64_t rdm_unet_missing(const rdm_unet_t* h);                           /* parameters not loaded yet */
RDM_API int rdm_unet_set_mode(rdm_unet_t* h, int32_t mode);
/* Debug aid: when on, every forward synchronises after each layer and records "<tag> <block> <layer> <mean> <absmean>"
 * lines (blocks in execution order: input_blocks, middle_block, output_blocks) retrievable as text. */
RDM_API int rdm_unet_set_debug(rdm_unet_t* h, int32_t on);
RDM_API const char* rdm_unet_debug_log(const rdm_unet_t* h);
/* context: float32 [B2, k, context_dim] (device).  Projects the step-invariant cross-attention K/V of all
 * SpatialTransformers once (attention.py:47-48); must precede rdm_unet_forward and be repeated when the
 * context, the batch or the weights change. */
RDM_API int rdm_unet_set_context(rdm_unet_t* h, const float* ctx_dev, int32_t B2, int32_t k, void* stream);
/* UNetModel.forward(x, timesteps, context): x float32 NCHW [Bx,C,H,W] with Bx == B2, or Bx == B2/2 to evaluate the
 * classifier-free-guidance doubling cat([x]*2) of ddim.py:233 without materialising it; t int64 [B2];
 * eps_out float32 NCHW [B2, out_channels, H, W]. */
RDM_API int rdm_unet_forward(rdm_unet_t* h, const float* x_dev, int32_t Bx, const int64_t* t_dev, int32_t B2,
                     int32_t H, int32_t W, float* eps_out_dev, void* stream);

/* Measurement aid: bit mask of kernel classes that are NOT launched by the following forwards (1 GroupNorm statistics, 2 GroupNorm apply,
 * 4 LayerNorm, 8 attention, 16 GEMMs with M >= 8192, 32 GEMMs with M < 8192, 64 un-fuse the cross-attention).  Outputs are garbage while a
 * mask is set; bench.py uses full-forward time minus GEMM-less forward time as the in-graph duration of the tcgen05 launches. */
RDM_API int rdm_unet_set_ablation(rdm_unet_t* h, int32_t mask);
/* Batch chains: the B2 rows of a forward / DDIM step are split into `chains` (1..8) contiguous sub-batches that run the whole layer
 * sequence concurrently on their own streams (fork / join by events, captured as branches of the step graph).  GroupNorm, LayerNorm and
 * attention are per-sample, so results do not depend on the split; small-resolution layers that cannot fill 148 SMs at once overlap.
 * Takes effect at the next forward. */
RDM_API int rdm_unet_set_chains(rdm_unet_t* h, int32_t chains);
/* CUDA-graph replay of the forward (default on).  Off: every kernel is launched eagerly. */
RDM_API int rdm_unet_set_graph(rdm_unet_t* h, int32_t on);
/* rdm_unet_forward, then the same forward replayed from a CUDA graph captured with external event-record nodes around every
 * GEMM launch (so the brackets hold the graph-replayed kernel only, no host launch gaps); synchronises.  out8 =
 * {tcgen05 GEMM ms, tcgen05 GEMM algorithmic flop, CUDA-core GEMM ms, CUDA-core GEMM flop, whole forward ms,
 *  #tcgen05 launches, #CUDA-core GEMM launches, 0}.  Used by bench.py for the live roofline numbers. */
RDM_API int rdm_unet_profile_forward(rdm_unet_t* h, const float* x_dev, int32_t Bx, const int64_t* t_dev, int32_t B2,
                             int32_t H, int32_t W, float* eps_out_dev, double* out8_host, void* stream);

/* Per-GEMM lines ("<engine> M= N= K= ks= HxW= act= ms= tflops=") of the last rdm_unet_profile_forward. */
RDM_API const char* rdm_unet_profile_text(const rdm_unet_t* h);

/* DDIMSampler.ddim_sampling (rdm/models/diffusion/ddim.py:143-215) for steps [first_step, first_step+num_steps) of a
 * schedule given as device tables IN SAMPLING ORDER: timesteps int64 [S] (= np.flip(ddim_timesteps)), coef float32
 * [S, 8] rows {sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma, 0, 0, 0}, optional noise float32
 * [S, B*C*H*W] (eta > 0).  x_dev float32 NCHW [B,C,H,W] is updated in place (x_T in, x_{t-1} of the last step out);
 * pred_x0_dev (optional) receives the last step's x0 prediction.  cfg_scale > 1 evaluates classifier-free guidance by
 * batch doubling (ddim.py:229-238): the context must have been set with B2 = 2B rows [cond | uncond]; otherwise B2 = B.
 * One CUDA graph {timestep fill, U-Net, fused CFG+DDIM update, step++} is captured once and replayed per step. */
RDM_API int rdm_ddim_sample(rdm_unet_t* h, float* x_dev, int32_t B, int32_t H, int32_t W, const int64_t* timesteps_dev,
                    const float* coef_dev, int32_t first_step, int32_t num_steps, float cfg_scale, const float* noise_dev,
                    float* pred_x0_dev, void* stream);

/* DDIMSampler.p_sample_ddim arithmetic after the model call (rdm/models/diffusion/ddim.py:236-238,253-267):
 * e = cfg ? e_u + scale*(e_c - e_u) : eps;  pred_x0 = (x - c0*e)/c1;  x_prev = c2*pred_x0 + c3*e (+ c4*noise),
 * coef_dev = {sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev-sigma^2), sigma} as float32 (device).
 * eps holds [cond | uncond] halves of n_per_half elements when cfg != 0.  Same rounding sequence as the
 * reference's float32 tensor expressions.  pred_x0/noise may be NULL. */
RDM_API int rdm_ddim_step(const float* x_dev, const float* eps_dev, int64_t n_per_half, int32_t cfg, float scale,
                  const float* coef_dev, const float* noise_dev, float* x_prev_dev, float* pred_x0_dev,
                  int32_t device, void* stream);


/* ------------------------------------------------------------------------------------------------
 * First-stage decode: `decode_first_stage` at rdm/models/diffusion/ddpm.py:840,981 ->
 * ldm VQModelInterface.decode = VectorQuantizer lookup -> post_quant_conv -> Decoder (conv_in, mid ResnetBlock / AttnBlock /
 * ResnetBlock, up levels with nearest-2x Upsample, GroupNorm-SiLU-conv_out); configuration = first_stage_config.params of
 * models/rdm/imagenet/config.yaml:60-80.  The decoder shares the executor of the U-Net: the handle is an rdm_unet_t and the
 * rdm_unet_num_params / param_name / param_numel / load / missing / set_mode / set_graph / destroy calls accept it.
 * Parameter names: latent-diffusion checkpoint keys below `first_stage_model.` ("decoder.mid.attn_1.q.weight",
 * "quantize.embedding.weight", "post_quant_conv.weight", ...).
 * ------------------------------------------------------------------------------------------------ */
typedef struct rdm_vqdec_cfg {
    int32_t embed_dim, n_embed;            /* codebook [n_embed, embed_dim]; embed_dim and z_channels both <= 4 (ldm VQ-f4/f8), or both
                                              multiples of 64 (taming VQGAN-f16 of the RARM models, models/rarm/imagenet/dogs/config.yaml:28-51:
                                              decoded from codebook entries, i.e. quantize = 0 only) */
    int32_t z_channels, resolution, out_ch, ch, num_res_blocks;
    int32_t n_ch_mult; int32_t ch_mult[8];
    int32_t n_attn_resolutions; int32_t attn_resolutions[8];
} rdm_vqdec_cfg;
RDM_API int rdm_vqdec_create(rdm_unet_t** out, const rdm_vqdec_cfg* cfg, int32_t device);
/* z float32 NCHW [B, embed_dim, h, w] (device) -> images float32 NCHW [B, out_ch, h * 2^(n_ch_mult-1), w * 2^(n_ch_mult-1)];
 * quantize == 0 is `force_not_quantize=True`.  fp16 tensor-core modes only (default fp16x2). */
RDM_API int rdm_vqdec_decode(rdm_unet_t* h, const float* z_dev, int32_t B, int32_t hh, int32_t ww, int32_t quantize, float* out_dev, void* stream);


/* ------------------------------------------------------------------------------------------------
 * CLIP encoders (retrieval queries): rdm/modules/custom_clip/model.py:201-235 (VisualTransformer), :166-198 (Transformer),
 * :304 (encode_image), :307-320 (encode_text); call sites rdm/modules/retrievers.py:83-95 (ClipImageRetriever),
 * scripts/rdm_sample.py:275-277 and scripts/rarm_sample.py:232-236 (clip.encode_text), dsetbuilder.py:461-473 (embed).
 * The struct mirrors the CLIP constructor (model.py:240-252); parameter names are the reference state-dict keys
 * ("visual.transformer.resblocks.0.attn.in_proj_weight", "text_projection", ...).
 * ------------------------------------------------------------------------------------------------ */
typedef struct rdm_clip rdm_clip_t;
typedef struct rdm_clip_cfg {
    int32_t embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size;
    int32_t context_length, vocab_size, transformer_width, transformer_heads, transformer_layers;
} rdm_clip_cfg;
RDM_API int rdm_clip_create(rdm_clip_t** out, const rdm_clip_cfg* cfg, int32_t device);
RDM_API void rdm_clip_destroy(rdm_clip_t* h);
RDM_API int64_t rdm_clip_num_params(const rdm_clip_t* h);
RDM_API const char* rdm_clip_param_name(const rdm_clip_t* h, int64_t i);
RDM_API int64_t rdm_clip_param_numel(const rdm_clip_t* h, const char* name);
RDM_API int rdm_clip_load(rdm_clip_t* h, const char* name, const float* host, int64_t numel);   /* host float32, reference layout */
RDM_API int64_t rdm_clip_missing(const rdm_clip_t* h);
RDM_API int rdm_clip_set_mode(rdm_clip_t* h, int32_t mode);                                       /* RDM_UNET_MODE_* (default bf16x3) */
/* CLIP.encode_text: tokens int64 [B, context_length] (device) -> float32 [B, embed_dim] */
RDM_API int rdm_clip_encode_text(rdm_clip_t* h, const int64_t* tokens_dev, int32_t B, float* out_dev, void* stream);
/* CLIP.encode_image: float32 NCHW [B, 3, R, R], already preprocessed -> float32 [B, embed_dim] */
RDM_API int rdm_clip_encode_image(rdm_clip_t* h, const float* image_dev, int32_t B, float* out_dev, void* stream);
/* ClipImageRetriever.preprocess (retrievers.py:83-95): x in [-1,1] NCHW [B,3,H,W] -> bicubic (align_corners=True) resize to
 * size x size, (x+1)/2, CLIP mean/std normalisation.  out float32 [B,3,size,size]. */
RDM_API int rdm_clip_preprocess(const float* image_dev, int32_t B, int32_t H, int32_t W, int32_t size, float* out_dev,
                        int32_t device, void* stream);


/* ------------------------------------------------------------------------------------------------
 * RARM decoder: rdm/modules/attention.py:199-272 (RetrievalPatchTransformer with `continuous: false`,
 * positional encodings, causal self-attention + cross-attention to the retrieved CLIP vectors, GEGLU feed-forward, Conv1d head;
 * models/rarm/imagenet/dogs/config.yaml:14-27) evaluated with per-layer key/value caches, and the sampling loop
 * rdm/models/autoregression/transformer.py:224-270 (LatentImageRETRO.sample; call sites :279-294, scripts/rarm_sample.py:253-266).
 * The struct mirrors `transformer_config.params`; parameter names are the reference state-dict keys below `transformer.`
 * ("proj_in.weight", "positional_encoding", "transformer_blocks.3.attn2.to_k.weight", "proj_out.weight" [out, C, 1], ...).
 * ------------------------------------------------------------------------------------------------ */
typedef struct rdm_rarm rdm_rarm_t;
typedef struct rdm_rarm_cfg {
    int32_t in_channels;        /* input vocabulary (16386 = codebook + mask + sos) */
    int32_t n_heads, d_head;    /* d_head must be 64 */
    int32_t depth, context_dim, sequence_length;
    int32_t out_channels;       /* output vocabulary (16384) */
} rdm_rarm_cfg;
RDM_API int rdm_rarm_create(rdm_rarm_t** out, const rdm_rarm_cfg* cfg, int32_t device);
RDM_API void rdm_rarm_destroy(rdm_rarm_t* h);
RDM_API int64_t rdm_rarm_num_params(const rdm_rarm_t* h);
RDM_API const char* rdm_rarm_param_name(const rdm_rarm_t* h, int64_t i);
RDM_API int64_t rdm_rarm_param_numel(const rdm_rarm_t* h, const char* name);
RDM_API int rdm_rarm_load(rdm_rarm_t* h, const char* name, const float* host, int64_t numel);   /* host float32, reference layout */
RDM_API int64_t rdm_rarm_missing(const rdm_rarm_t* h);
/* RDM_UNET_MODE_FP32: fp32 weights (strict parity);  RDM_UNET_MODE_TC_FP16 (default): fp16 weights, fp32 accumulation. */
RDM_API int rdm_rarm_set_mode(rdm_rarm_t* h, int32_t mode);
RDM_API int rdm_rarm_set_graph(rdm_rarm_t* h, int32_t on);
/* Retrieval context r: float32 [B2, k, context_dim] (device); under guidance B2 = 2B rows [r | zeros] (transformer.py:233-236).
 * Projects the step-invariant cross-attention keys/values of every layer once and restarts the sequences (empty caches). */
RDM_API int rdm_rarm_set_context(rdm_rarm_t* h, const float* ctx_dev, int32_t B2, int32_t k, void* stream);
/* `transformer(x, context=r)[:, pos]` for the token at position `pos` of every sequence, given that positions 0..pos-1 were fed
 * before (in order, since the last rdm_rarm_set_context): tokens_dev int64 [B] with B == B2 or B == B2/2 (guidance doubling:
 * row b + B reuses token b) -> logits_out_dev float32 [B2, out_channels]. */
RDM_API int rdm_rarm_forward_token(rdm_rarm_t* h, const int64_t* tokens_dev, int32_t B, int32_t pos, float* logits_out_dev, void* stream);
/* transformer.py:249-266 on given last-position logits [B or 2B, V] ([cond | uncond] when guided != 0; V <= 40960):
 * logits = (l_u + scale*(l_c - l_u)) / temperature -> top-k filter (top_k <= 0: none) -> softmax -> one token per sequence.
 * uniforms_dev float32 [B] in [0,1): token = first index whose cumulative probability exceeds u (the device-side definition of
 * torch.multinomial(probs, 1)); NULL: argmax (`sample=False`).  probs_out_dev (optional) float32 [B, V]. */
RDM_API int rdm_rarm_sample_step(rdm_rarm_t* h, const float* logits_dev, int32_t B, int32_t V, int32_t guided, float guidance_scale, float temperature,
                         int32_t top_k, const float* uniforms_dev, int64_t* token_out_dev, float* probs_out_dev, void* stream);
/* LatentImageRETRO.sample: tokens_dev int64 [B, n_prefix + steps]; columns [0, n_prefix) hold cat(c, x) (the sos token and any
 * start tokens, transformer.py:232), columns [n_prefix, n_prefix + steps) receive the sampled ids.  uniforms_dev float32
 * [steps, B] (NULL: greedy).  The context must have been set with B2 = B (guidance_scale <= 1) or 2B rows.  One CUDA graph
 * {embed, layers, head, guided top-k draw, position++} is captured once and replayed per position; no host synchronisation. */
RDM_API int rdm_rarm_sample(rdm_rarm_t* h, int64_t* tokens_dev, int32_t B, int32_t n_prefix, int32_t steps, float temperature, int32_t top_k,
                    float guidance_scale, const float* uniforms_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif
