/* msd_b200.h -- C ABI of libmsd_b200.so: the B200 (sm_100a) implementation of the DDPM
 * sampling hot path of magenta/music-spectrogram-diffusion.
 *
 * The reference has no FFI/plugin layer (it is pure Python on JAX/XLA); the drop-in boundary
 * is the Python protocol of `inference.InferenceModel` (music_spectrogram_diffusion/
 * inference.py:68-203).  These entry points are what a binding for that protocol binds; each
 * one names the reference code it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions: every function returns 0 on success, a negative code on failure
 * (-1 bad argument / unsupported configuration, -2 CUDA error, -3 missing weight);
 * `msd_last_error()` returns a thread-local message.  Pointers documented "device" are CUDA
 * device pointers owned by the caller; "host" are ordinary host pointers.  A context is bound
 * to one GPU and one stream at a time and is not thread-safe.  `stream` is a cudaStream_t
 * passed as void* (NULL = legacy default stream).
 */
#ifndef MSD_B200_H_
#define MSD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSD_B200_ABI_VERSION 4

typedef struct msd_ctx msd_ctx;

/* Static model + sampler description.
 * Mirrors network.T5Config (models/diffusion/network.py:54-72), the task feature lengths
 * (gin/tasks/mt3/context_mega.gin:5), diffusion_utils.DiffusionConfig/SamplerConfig
 * (models/diffusion/diffusion_utils.py:25-59) and the MelGAN codec constants
 * (audio_codecs.py:204-213). */
typedef struct msd_config {
  int32_t vocab_size;            /* T5Config.vocab_size (1536 for the mt3 vocabulary) */
  int32_t emb_dim;               /* multiple of 128, <= 1024 */
  int32_t num_heads;
  int32_t head_dim;              /* must be 64 */
  int32_t num_encoder_layers;
  int32_t num_decoder_layers;
  int32_t mlp_dim;               /* gated ('gelu','linear') MLP; multiple of 64 */
  int32_t inputs_length;         /* token positions, multiple of 128 */
  int32_t targets_length;        /* mel frames per segment, multiple of 128 */
  int32_t context_length;        /* context frames, multiple of 128 */
  int32_t n_dims;                /* mel bins, must be 128 */
  int32_t num_steps;             /* sampler schedule num_steps */
  int32_t max_batch;             /* segments per call (B) */
  int32_t sampler;               /* 0 = ddpm, 1 = ddim */
  int32_t logvar_type;           /* 0 = large, 1 = small, 2 = 'medium:<logvar_frac>' (ddpm only) */
  int32_t clip_x0;               /* SamplerConfig.clip_x0 */
  int32_t context_positions;     /* 0 = regular, 1 = terminal_relative */
  float max_decoder_noise_time;  /* T5Config.max_decoder_noise_time (2e4) */
  float eval_condition_weight;   /* classifier-free guidance weight; 1 disables the 2nd pass */
  float feature_min;             /* codec min_value (log 1e-5) */
  float feature_max;             /* codec max_value (4.0) */
  /* ABI 2: sampler variants of diffusion_utils.py (all 0 = the shipped gin defaults) */
  int32_t model_output;          /* DiffusionConfig.model_output: 0 eps, 1 x0, 2 v (288-321) */
  int32_t sampler_schedule;      /* sampler.schedule.name: 0 cosine, 1 linear (166-202) */
  int32_t train_schedule;        /* train_schedule.name: 0 cosine, 1 linear */
  int32_t train_num_steps;       /* train_schedule.num_steps (linear only) */
  float logvar_frac;             /* frac of 'medium:<frac>' (141-156) */
  float sampler_beta_start;      /* linear sampler schedule: beta range */
  float sampler_beta_stop;
  float train_beta_start;        /* linear train schedule: beta range */
  float train_beta_stop;
  int32_t cross_attend_style;    /* T5Config.decoder_cross_attend_style: 0 concat_encodings,
                                    1 sum_cross_attends (network.py:199-216) */
  int32_t rng_kind;              /* noise when msd_sample gets no init_z / noise: 0 = Philox4x32-10
                                    (library stream), 1 = jax.random threefry2x32 stream of
                                    PRNGKey(seed) / fold_in(key, i) (inference.py:203,
                                    diffusion_utils.py:389-390, 462) */
  /* ABI 3 */
  int32_t precision;             /* T5Config.dtype as executed: 0 = bf16 tensor-core operands with
                                    fp32 accumulation / residual / softmax (the fast path);
                                    1 = fp32-accurate: every dense layer as a 3 x bf16 split-
                                    precision tensor-core product (~2^-16 per product), fp32
                                    attention, exact tanh -- what the shipped gins ask for
                                    (gin/models/diffusion/context/t5_base.gin:72 dtype float32) */
} msd_config;

/* A named fp32 parameter in the reference's own layout (flax tree path joined by '/',
 * kernels [in, out]); see SURVEY.md App. B. */
typedef struct msd_tensor {
  const char* name;
  const float* data; /* host pointer */
  int32_t ndim;
  int64_t shape[4];
} msd_tensor;

const char* msd_last_error(void);
int msd_abi_version(void);

/* Replaces InferenceModel.__init__'s model construction (inference.py:71-111). */
int msd_create(const msd_config* cfg, int device, msd_ctx** out);
void msd_destroy(msd_ctx* ctx);

/* Replaces InferenceModel._restore_from_checkpoint (inference.py:159-176): takes the fp32
 * parameter tree, repacks it into the kernel layouts (bf16 [out, in], fused QKV / gated-MLP /
 * split-precision projections) and tabulates the timestep conditioning for all num_steps
 * (network.py:377-394 + layers.py:652-666: FiLM scale|bias for every (step, layer)). */
int msd_load_weights(msd_ctx* ctx, const msd_tensor* tensors, int32_t n);

/* Replaces ContextDiffusionModel.predict_batch_with_aux's scale_features + module.encode
 * (models/diffusion/models.py:361-371; network.py:537-559) and additionally projects the
 * concatenated encodings to every decoder layer's cross-attention K/V once (network.py:217-230
 * is loop-invariant).  tokens [B, inputs_length] int32, ctx_features [B, context_length,
 * n_dims] f32 in codec feature units, ctx_mask [B, context_length] int32: device pointers. */
int msd_encode(msd_ctx* ctx, const int32_t* tokens, const float* ctx_features,
               const int32_t* ctx_mask, int32_t batch, void* stream);

/* Replaces diffusion_utils.eval_scan + scale_to_features (diffusion_utils.py:456-476;
 * models.py:393-395) for the batch passed to the preceding msd_encode.
 * init_z  [B, targets_length, n_dims] f32 device, or NULL -> Philox N(0,1) from `seed`.
 * noise   [num_steps, B, targets_length, n_dims] f32 device (noise[i] is used at step i),
 *         or NULL -> Philox from `seed`.
 * mel_out [B, targets_length, n_dims] f32 device, codec feature units. */
int msd_sample(msd_ctx* ctx, const float* init_z, const float* noise, uint64_t seed,
               float* mel_out, void* stream);

/* ---- one song on two GPUs: classifier-free guidance split (BASELINE config 5) ----
 * The two decoder passes of a reverse step (diffusion_utils.py:415, 428-429) are independent until
 * the guidance combine (430-433).  With a peer attached, a context runs ONE of them (role 1: the
 * conditional pass incl. cross-attention, role 2: the unconditional one) and its sampler kernel
 * exchanges the 128 KB of predicted noise with the peer by direct NVLink stores plus a flag word
 * (no NCCL call, no host involvement inside the loop); both GPUs then apply the identical update,
 * so both hold the same z and the same final mel.  Protocol: each process calls msd_p2p_export,
 * the 64-byte handles are swapped by any host channel (torch.distributed in distributed.py), each
 * calls msd_p2p_attach with the OTHER rank's handle, and then both make the same msd_encode /
 * msd_sample calls.  The processes must live on one node with peer access between the GPUs. */
int msd_p2p_export(msd_ctx* ctx, void* handle_out /* 64 bytes */);
int msd_p2p_attach(msd_ctx* ctx, const void* peer_handle /* 64 bytes */, int32_t role);
int msd_p2p_detach(msd_ctx* ctx);

/* Test hook == module.decode (network.py:561-573) at diffusion step `step_i`
 * (time = (step_i + 1) / num_steps): z [B, targets_length, n_dims] f32 device ->
 * eps_out [B, ...] f32 device.  conditioned = 0 multiplies encodings and masks by 0
 * (models.py:376-377). */
int msd_decode_eps(msd_ctx* ctx, const float* z, int32_t step_i, int32_t conditioned,
                   float* eps_out, void* stream);

/* Test hook: copies the encoder outputs of the last msd_encode as bf16-rounded f32:
 * enc_out [B, inputs_length + context_length, emb_dim] device f32. */
int msd_get_encodings(msd_ctx* ctx, float* enc_out, void* stream);

/* Host copy of the per-step sampler scalars [num_steps][16]:
 * 0 x0_scale, 1 eps_scale (predict_x0_from_eps at the sampler's logsnr_t), 2 c_z, 3 c_x0,
 * 4 sigma (ddpm: mean = c_z z + c_x0 x0, std sigma; ddim: 2 = stdv_s, 3 = alpha_s),
 * 5 is_last, 6 logsnr_t, 7 logsnr_s, 8 p0, 9 p1 (eps = p0 z + p1 model_output, train schedule),
 * 10 q0, 11 q1 (x0 = q0 z + q1 model_output, train schedule), 12 e1, 13 e2
 * (predict_eps_from_x0 at logsnr_t: eps = e1 (z - x0 e2)), 14 logsnr_train, 15 reserved. */
int msd_get_step_table(msd_ctx* ctx, float* table_host);

/* Profiling hook: runs diffusion step `step_i` (1 <= step_i < num_steps) of the batch of the
 * last msd_encode `reps` times WITHOUT graph capture, bracketing every kernel launch with CUDA
 * events on the launching stream.  out[5][4] (doubles), one row per kernel class
 * {0 gemm, 1 attention, 2 rmsnorm/FiLM, 3 sampler, 4 other}:
 * {milliseconds per step, launches per step, algorithmic FLOPs per step, algorithmic bytes per
 * step}.  Leaves the sampler state (z) modified; call msd_sample afterwards as usual. */
int msd_profile_step(msd_ctx* ctx, int32_t step_i, int32_t reps, double* out);

/* Kernel launches issued so far by this library in this process (graph replays count their
 * kernel nodes). */
uint64_t msd_launch_count(void);

/* ---- operator-level entry points (unit parity against msd/layers.py) ------------------- */

/* DenseGeneral (layers.py:397-442): out[M,N] f32 = A[M,K] * W, with A given as bf16-rounded
 * f32 [M,K] device and W as f32 [K,N] host-layout device pointer (packed internally). */
int msd_op_dense(const float* a, const float* w, int32_t M, int32_t N, int32_t K, float* out,
                 void* stream);

/* Same, selecting the kernel: variant 0 = CTA-pair persistent kernel (default), 1 = single-CTA
 * kernel; block_n 0 = auto or one of 64/128/192/256 (192 only with variant 0). */
int msd_op_dense_variant(const float* a, const float* w, int32_t M, int32_t N, int32_t K,
                         float* out, int32_t variant, int32_t block_n, void* stream);

/* Micro-benchmark hook: average milliseconds of `iters` back-to-back launches of the bf16 GEMM
 * [M,K] x [N,K]^T with the given epilogue (0 bf16, 1 f32, 2 f32 + residual, 3 gated-GELU) on
 * zero-filled scratch buffers. */
int msd_bench_gemm(int32_t M, int32_t N, int32_t K, int32_t epilogue, int32_t variant,
                   int32_t block_n, int32_t iters, float* ms_out);

/* Micro-benchmark hook: average milliseconds of `iters` back-to-back launches of the bf16
 * attention kernel (+ its combine kernel when the split is not merged in-kernel) on scratch
 * buffers filled with small pseudo-random values; kv_static as in the decoder's cross-attention.
 * The instance / split is chosen as in production (or forced by MSD_ATTN_BKV / _SPLITS / _MERGE). */
int msd_bench_attention(int32_t nb, int32_t heads, int32_t Lq, int32_t Lk, int32_t iters,
                        float* ms_out);

/* dot_product_attention (layers.py:109-181) for head_dim 64 with a key-padding mask:
 * q [nb, Lq, heads*64], k/v [nb, Lk, heads*64] f32 device, key_mask [nb, Lk] int32 or NULL,
 * out [nb, Lq, heads*64] f32 device. */
int msd_op_attention(const float* q, const float* k, const float* v, const int32_t* key_mask,
                     int32_t nb, int32_t heads, int32_t Lq, int32_t Lk, float* out, void* stream);

/* Same with a debugging trace: `trace` (device, int64 [2][64][8], may be NULL) receives clock64
 * stamps of the softmax phases of CTA (0,0,0) for each key block (see attention_tcgen05.cu). */
int msd_op_attention_trace(const float* q, const float* k, const float* v,
                           const int32_t* key_mask, int32_t nb, int32_t heads, int32_t Lq,
                           int32_t Lk, float* out, int64_t* trace, void* stream);

/* DenseGeneral with each fused epilogue of the hot path (kernels.h GemmEpilogue), for unit parity:
 *   0 bf16 out                      out [M, N]            = bf16(a w)
 *   2 f32 + residual                out [M, N]            = a w + resid [M, N]
 *   3 gated GELU (layers.py:483-509) out [M, N]           = bf16(gelu_tanh(a w) * (a w1)), w / w1 [K, N]
 *   4 f32 + position rows           out [M (+ dup), N]    = a w + pos[(r % pos_rows - shift[r /
 *                                    pos_rows]) mod pos_rows] (network.py:327-334, 420-427), rows
 *                                    also stored at r + dup_rows when dup_rows > 0
 *   5 gated GELU, fp32-accurate     out [M, N]            = hi + lo of the [hi | lo | hi] output,
 *                                    computed from 3 x bf16 split operands (a, w, w1 used in full
 *                                    fp32 precision)
 * a [M, K], w [K, N] f32 device (bf16-rounded by the caller for epilogues 0-4); block_n 0 = auto.
 * Unused pointers may be NULL.  out is f32 device. */
int msd_op_dense_epilogue(const float* a, const float* w, const float* w1, int32_t M, int32_t N,
                          int32_t K, int32_t epilogue, int32_t block_n, const float* resid,
                          const float* pos, int32_t pos_rows, const int32_t* pos_shift,
                          int32_t dup_rows, float* out, void* stream);

/* The deferred-normalisation pair of the bf16 hot path: a residual projection
 * whose epilogue also prepares the next pre-norm, and the projection that consumes it.
 *   stage 1   x_out [M, d] = x + a w_out;  operand = bf16(x_out * g(row)), g = g_lo for rows <
 *             split_row, else g_hi;  row sums of squares of x_out kept per column tile
 *   stage 2   y [M, N2] = bf16(rsqrt(mean(x_out^2) + 1e-6)[row] * (operand w2) + bias)     (w2b NULL)
 *             y = bf16(gelu_tanh(u) * u1), u | u1 the same through w2 | w2b                (gated)
 * i.e. y == bf16((rmsnorm(x_out) * g) w2 + bias) up to operand rounding (layers.py:632-666 with
 * g = scale * (1 + film_scale), bias = film_bias w2).  a [M, K], w_out [K, d], x [M, d], g_* [d],
 * w2 / w2b [d, N2], bias [N2] (gated: [2 * N2] in accumulator column order: 32 of w2, 32 of w2b,
 * ...) or NULL; all f32 device, a / w_out / w2 / w2b bf16-rounded by the callee.  block_n1 /
 * block_n2: tile widths of the two GEMMs (0 = auto).  Outputs f32 device. */
int msd_op_dense_deferred_norm(const float* a, const float* w_out, const float* x, int32_t M, int32_t d,
                               int32_t K, const float* g_lo, const float* g_hi, int32_t split_row,
                               const float* w2, const float* w2b, int32_t N2, const float* bias,
                               int32_t block_n1, int32_t block_n2, float* x_out, float* y_out,
                               void* stream);

/* dot_product_attention of the fp32-accurate mode: as msd_op_attention, but q / k / v are used in
 * full fp32 and the result is returned as hi + lo of the kernel's [hi | lo | hi] output. */
int msd_op_attention_f32(const float* q, const float* k, const float* v, const int32_t* key_mask,
                         int32_t nb, int32_t heads, int32_t Lq, int32_t Lk, float* out,
                         void* stream);

/* LayerNorm (layers.py:632-649) followed by optional FiLM (layers.py:652-666) with explicit
 * scale|bias vector film [2*d] (NULL = none): out f32 (bf16-rounded) [rows, d]. */
int msd_op_rmsnorm_film(const float* x, const float* gamma, const float* film, int32_t rows,
                        int32_t d, float* out, void* stream);

/* jax.random.normal of the sampler's stream: step < 0 -> normal(PRNGKey(seed), [n]) (init_z,
 * diffusion_utils.py:462), else normal(fold_in(PRNGKey(seed), step), [n]) (389-390).
 * out: device f32 [n], n a multiple of 8.  Test hook for the rng_kind = 1 generator. */
int msd_op_jax_normal(uint64_t seed, int32_t step, int64_t n, float* out, void* stream);

/* The raw threefry2x32 words those normals are made from (jax.random.bits of the same key):
 * out device uint32 [n].  Integer work: the test compares bit-exactly. */
int msd_op_jax_bits(uint64_t seed, int32_t step, int64_t n, uint32_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSD_B200_H_ */
