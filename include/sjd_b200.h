/* sjd_b200.h — C ABI of the B200-native Speculative Jacobi Decoding hot path.
 *
 * The reference (tyshiwo1/Accelerating-T2I-AR-with-SJD) has no FFI: its "plugin API" is Python
 * class-swapping (scheduler/jacobi_iteration_lumina_mgpt.py:1340-1346).  This library sits UNDER the
 * Python mirror of that API (accelerating-t2i-ar-with-sjd_b200/hf_api.py): every entry point below replaces
 * a stretch of PyTorch-eager code of the reference's per-iteration hot loop, cited per function.
 *
 * Conventions: plain pointers and sizes; all device pointers are CUDA device memory of the current
 * device; `stream` is a cudaStream_t passed as void*; functions return 0 on success or a negative
 * SJD_E* code, never throw, and (except *_create) never allocate.  A context is thread-compatible,
 * not thread-safe.  bf16 buffers are raw uint16 storage.
 */
#ifndef SJD_B200_H_
#define SJD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SJD_OK 0
#define SJD_E_DRIVER (-1)   /* driver entry point / tensor-map encode failed */
#define SJD_E_TMAP (-2)
#define SJD_E_ARG (-3)      /* bad shape / argument */
#define SJD_E_SMEM (-4)
#define SJD_E_ATTR (-5)
#define SJD_E_LAUNCH (-6)   /* kernel launch failed; see sjd_last_error() */
#define SJD_E_ALLOC (-7)
#define SJD_E_STATE (-8)    /* weights missing, wrong call order */

#define SJD_MAX_ROWS 8      /* CFG rows per context (reference uses 2: cond + uncond) */
#define SJD_MAX_TOKENS 256  /* token rows per forward call (rows * window) */

/* Library identification / diagnostics. */
int sjd_version(void);
const char* sjd_last_error(void);
int sjd_device_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * Stand-alone weight-streaming GEMM  Y[M,N] = X[M,K] * W[N,K]^T  (bf16 in, fp32 accumulate), the kernel every
 * projection of the window forward runs on.  Replaces nn.Linear -> cuBLAS in the reference forward
 * (modeling_chameleon.py:527-529,579,193-195,1560; llamagen/llamagen.py:248,277,200,332).
 * `x` must have at least m_tile = round_up(M,16) rows (<= 256); K % 64 == 0.  `out` is [M,N] fp32 (out_f32=1,
 * optionally rounded through bf16) or bf16.  `ws` is sjd_gemm_workspace_bytes() of device memory that must be
 * ZERO before its first use (the kernel leaves its counters zero).  Split tiles are fixed up in-kernel in a fixed
 * order, so results are bit-reproducible.  grid_limit <= 0 means one CTA per SM.
 * ---------------------------------------------------------------------------------------------- */
size_t sjd_gemm_workspace_bytes(int N, int K, int m_tile, int grid_limit);
int sjd_gemm_bf16(const void* w, int N, int K, const void* x, int x_rows, int M, void* out, int out_f32,
                  int round_bf16, void* ws, int grid_limit, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Verify step (one call per Jacobi iteration).  Replaces sampling_logits2tokens
 * (scheduler/jacobi_iteration_lumina_mgpt.py:82-132), the 3-D logits processors
 * (scheduler/logit_processor_3dim.py:45-204; HF TopKLogitsWarper for LlamaGen/Emu3),
 * SpeculativeSampler.__call__ (:247-315) incl. residual resampling (:203-241), and
 * prefix_matching_next_tokens / find_first_misaligned_token_inds (:317-376).
 * ---------------------------------------------------------------------------------------------- */
typedef struct sjd_verify_args {
  const float* logits;    /* [(has_uncond ? 2 : 1) * W][V] fp32, cond rows first */
  int32_t W, V;
  int32_t has_uncond;     /* logits carry CFG-uncond rows */
  int32_t apply_cfg;      /* 1: g*(c-u)+u ; 0: cond only (check_is_force_no_cfg, :70-80) */
  float guidance;
  float temperature;
  int32_t allow_lo, allow_hi; /* grammar: ids outside [lo,hi) suppressed; off when hi <= lo */
  const int32_t* forced;  /* [W] forced id per window position (EOL/EOI/...), -1 = free; may be NULL */
  const int32_t* forced_resid; /* [W] forced id of the RESIDUAL distribution when the draft after position j is rejected
                           * (the reference re-runs its processors on a 1-token window there, which a position-
                           * and window-length-dependent grammar such as Emu3's answers differently); NULL = `forced`.
                           * A value <= -2 means "mask-forced" id f = -2 - value (Anole's processors fill every other id
                           * with finfo.min instead of writing a one-hot row): f is returned when its residual is finite,
                           * otherwise the draw is uniform over the other ids, as the reference's softmax makes it */
  int32_t top_k;          /* 0 = off; ties with the k-th largest are kept (scores < kth removed) */
  float top_p_thresh;     /* TopPLogitsWarper3d (logit_processor_3dim.py:406-419), applied after top-k: float32(1 - top_p);
                           * the smallest probabilities whose running sum stays <= this are removed; 0 = off (top_p = 1) */
  int32_t do_sample;      /* 0: argmax */
  int32_t scheme;         /* 0: 'speculative_jacobi', 1: 'jacobi' (:1032-1048) */
  const int32_t* draft;   /* [W] window ids, [0] = last accepted token */
  const int32_t* q_row;   /* [W] row of p_prev with the draft's distribution; -1: one-hot (fresh random draft) */
  const float* p_prev;    /* [>=W][V] probabilities written by the previous call */
  float* p_cur;           /* [>=W][V] out: probabilities of this call */
  const float* noise_e1;  /* [W][V] Exp(1) noise == torch.multinomial's (q = empty_like(p).exponential_()) */
  const float* noise_u;   /* [W] torch.rand([1,W,V])[0, i, draft[i]] */
  const float* noise_e2;  /* [V] Exp(1) noise of the residual multinomial (consumed only on rejection) */
  int32_t eoi_token;      /* accepted draft == eoi -> residual processed in text mode; -1 = never */
  int32_t text_top_k;
  float* resid;           /* [V] scratch */
  int32_t* next_tokens;   /* [W] scratch */
  int32_t* out_tokens;    /* [W] tokens after accept/resample: first `matched` are final, rest are next drafts */
  int32_t* out_info;      /* [4]: matched, rejected, first_reject, residual_text_mode */
  uint32_t* sync_ws;      /* one device word, ZERO before its first use (the kernel leaves it zero): with it the whole step
                           * is ONE launch — the CTA that finishes its window position last runs the accept scan.  NULL:
                           * two launches (rows, then accept).  Not to be shared by calls in flight on different streams. */
  /* Device-side noise.  rng_mode = 1: noise_e1 / noise_u / noise_e2 are ignored; the kernel computes, for exactly the
   * elements it consumes, the values the reference's torch.Generator(device='cuda') would have written into them
   * (jacobi_iteration_lumina_mgpt.py:118 multinomial -> exponential_ [W,V]; :260 rand [1,W,V]; :237-240 multinomial ->
   * exponential_ [1,V]): Philox4x32-10 keyed by rng_seed, counter from rng_off[k] (the generator's philox offset when
   * draw k starts, in 32-bit outputs) and torch's element -> (thread, iteration) mapping with grid-stride rng_span[k]
   * = 256 * min(SMs * maxThreadsPerSM / 256, ceil(numel / 256)).  The caller advances its offset exactly like torch:
   * by ((numel - 1) / (rng_span * 4) + 1) * 4 per draw. */
  int32_t rng_mode;
  uint64_t rng_seed;
  uint64_t rng_off[3];
  uint32_t rng_span[3];
  /* Candidate sets beyond one id range — what the 3-D Chameleon / Anole processors (scheduler/logit_processor_3dim.py:
   * 207-353, installed per multimodal_generation_mode by scheduler/jacobi_iteration_anhole.py:170-265) reduce to for one
   * call.  allow_mode 0 (a zero-initialised struct): [allow_lo, allow_hi) as above, ban[] not looked at; 1: the same
   * minus the ids ban[0], ban[1] (-1 = none); 2: every id EXCEPT [allow_lo, allow_hi) and ban[] (text between images:
   * image ids, end-of-image and, late, begin-of-image removed); 3: only the ids ban[0] and ban[1] (both >= 0) are kept;
   * 4: every id except ban[]. */
  int32_t allow_mode;
  int32_t ban[2];
  /* resid_set = 1: residual positions j >= resid_from whose forced_resid is -1 use THIS candidate set (same encoding)
   * instead of the window's — for grammars whose decision changes with the drafts accepted inside the window (Anole: the
   * positions after a forced begin-of-image are image ids, the positions after a forced end-of-image are text). */
  int32_t resid_set;
  int32_t resid_allow_mode, resid_allow_lo, resid_allow_hi, resid_ban[2], resid_from;
  /* Completion flag for a caller that reads the result without a stream synchronisation: when done_flag != NULL the kernel
   * stores done_seq to *done_flag after out_tokens / out_info are written, behind a system-scope fence.  Meant for
   * out_tokens / out_info / done_flag in mapped pinned HOST memory (device-accessible under unified addressing): the host
   * polls the flag and finds the result next to it — no device-to-host copy, no cudaStreamSynchronize on the per-iteration
   * path (the reference reads `.item()`s, i.e. synchronises, several times per iteration:
   * scheduler/jacobi_iteration_lumina_mgpt.py:335-376).  Needs sync_ws (the one-launch form). */
  int32_t* done_flag;
  int32_t done_seq;
} sjd_verify_args;

int sjd_verify(const sjd_verify_args* args, void* stream);
/* Test / developer entry: fills out[0..numel) with what rng_mode = 1 draws for a noise tensor of numel elements:
 * kind 0 = exponential_(1), kind 1 = rand.  Bit-compared against torch on the GPU by the parity tests. */
int sjd_debug_philox(float* out, uint64_t numel, uint64_t seed, uint64_t offset, uint32_t span, int kind, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Model context: the fused transformer stack of the draft-window forward with a static KV cache.
 * Replaces `outputs = self(**model_inputs)` (scheduler/jacobi_iteration_lumina_mgpt.py:1107) for the
 * Chameleon/Lumina (modeling_chameleon.py:1494-1591), LlamaGen (llamagen/llamagen.py:297-337 via
 * llamagen_solver.py:234-295) and Emu3 (emu3/mllm/modeling_emu3.py:1140-1278) decoder stacks, the mask
 * builders (_update_causal_mask, :1256-1336) and the KV roll-back (delete_false_key_value, :47-54 —
 * here roll-back is just a smaller kv_len on the next call).
 * ---------------------------------------------------------------------------------------------- */
typedef struct sjd_model_cfg {
  int32_t n_layers, d_model, n_heads, n_kv_heads, head_dim, d_ff, vocab;
  float rms_eps;
  int32_t qk_norm;           /* 1: per-head LayerNorm on q,k (Chameleon) */
  int32_t rope_interleaved;  /* 0: rotate-half (Chameleon, Emu3); 1: adjacent pairs (LlamaGen 2-D RoPE) */
  int32_t rows;              /* CFG rows held in the KV cache (1 or 2) */
  int32_t max_len;           /* KV cache capacity per row (tokens) */
  int32_t n_rope_pos;        /* rows of the rope tables */
  int32_t logits_round_bf16; /* 1: round logits through bf16 like a bf16 lm_head */
} sjd_model_cfg;

typedef struct sjd_layer_weights {   /* device pointers, bf16, nn.Linear layout [out, in]; COPIED by set_layer */
  const void* attn_norm;   /* [d] */
  const void* wqkv;        /* [(H + 2*Hkv) * Dh, d]  rows: q heads, k heads, v heads */
  const void* q_norm_w; const void* q_norm_b;  /* [H, Dh] or NULL */
  const void* k_norm_w; const void* k_norm_b;  /* [Hkv, Dh] or NULL */
  const void* wo;          /* [d, H*Dh] */
  const void* ffn_norm;    /* [d] */
  const void* w_gate_up;   /* [2*d_ff, d]  rows: gate then up */
  const void* w_down;      /* [d, d_ff] */
} sjd_layer_weights;

typedef struct sjd_ctx sjd_ctx;

int sjd_ctx_create(const sjd_model_cfg* cfg, sjd_ctx** out);
void sjd_ctx_destroy(sjd_ctx* ctx);
size_t sjd_ctx_device_bytes(const sjd_ctx* ctx);
/* The context re-lays the weights out once into its own memory (layers stacked per projection type, gate/up rows
 * interleaved 64/64 so SiLU*up fuses into the GEMM epilogue); the caller may free its tensors afterwards. */
int sjd_ctx_set_layer(sjd_ctx* ctx, int layer, const sjd_layer_weights* w);
/* embed: [vocab, d] bf16 (may be NULL if every forward passes embeddings); final_norm [d]; lm_head [vocab, d];
 * rope_cos/sin: fp32 [n_rope_pos, head_dim/2].  Copied. */
int sjd_ctx_set_globals(sjd_ctx* ctx, const void* embed, const void* final_norm, const void* lm_head,
                        const float* rope_cos, const float* rope_sin);

typedef struct sjd_forward_args {
  int32_t W;                 /* tokens per row in this call; rows*W <= SJD_MAX_TOKENS */
  const int32_t* ids;        /* device [rows*W] token ids (row-major), or NULL when embeds is given */
  const void* embeds;        /* device bf16 [rows*W, d] input embeddings, or NULL */
  const int32_t* rope_pos;   /* device [rows*W] index into the rope tables, < n_rope_pos (the call is rejected when
                              * kv_len + W - min(kv_lo) > n_rope_pos, i.e. when slot - first-visible-key positions would
                              * run past the table) */
  const int32_t* cache_pos;  /* device [rows*W] KV slot written by each token */
  int32_t kv_len;            /* keys already valid in the cache; this call's tokens sit at kv_len .. kv_len+W-1 */
  int32_t kv_lo[SJD_MAX_ROWS]; /* first visible key per row (CFG hidden prefix / left padding) */
  int32_t n_logit_tokens;    /* logits for the last n tokens of every row (0 < n <= W) */
  float* logits;             /* device out [rows * n_logit_tokens, vocab] fp32, row b first */
} sjd_forward_args;

int sjd_ctx_forward(sjd_ctx* ctx, const sjd_forward_args* a, void* stream);
/* Launches only the fused GEMMs of one window forward (same weights/buffers/epilogues/order, attention skipped):
 * lets bench.py time the dominant kernel (gemm_fused_kernel) in isolation with CUDA events. */
int sjd_ctx_gemm_only(sjd_ctx* ctx, int W, void* stream);
/* Developer timing: when device_buf != NULL every following GEMM launch i writes clock64 stamps of its epilogue
 * stages to device_buf[(i % n_launches)][cta < 256][16] (int64).  NULL switches it off. */
void sjd_debug_gemm_stamps(void* device_buf, int n_launches);
/* Developer timing of the tensor-core attention: CTA 0 of every following launch writes clock64 stamps of its pipeline
 * stages to device_buf[unit < 8][16] (int64; the last launch wins).  NULL switches it off. */
void sjd_debug_attn_stamps(void* device_buf);
/* ------------------------------------------------------------------------------------------------
 * Output side of the loop: image-token ids -> the latent feature map a VQGAN decoder starts from.
 * Replaces, in one launch, get_codebook_entry + post_quant_conv of the reference's decode_code:
 *   LlamaGen   llamagen/tokenizer/tokenizer_image/vq_model.py:52-55, :261-275 (L2-normalised codebook), :39, :47-49
 *   Chameleon  lumina_mgpt/model/chameleon_vae_ori/vqgan.py:594-597, :131-146, :589-592
 * codes: device int32 [n_pix] (n_pix = batch * hw, row-major latent grid); codebook: device fp32 [n_e, e_dim];
 * w / bias: post_quant_conv weight [z, e_dim] (its 1x1 kernel squeezed) and bias [z]; out: device fp32 [batch, z, hw]
 * (NCHW).  l2_norm = 1 normalises every codebook row like F.normalize (eps 1e-12).  Ids outside [0, n_e) are clamped.
 * ---------------------------------------------------------------------------------------------- */
int sjd_vq_lookup(const int32_t* codes, int n_pix, int hw, const float* codebook, int n_e, int e_dim, int l2_norm,
                  const float* w, const float* bias, int z, float* out, void* stream);

/* Developer / test (pure host arithmetic, no GPU needed): the work split attention_sw.cu would use for a window of W tokens
 * over kv_len cached keys — ub_out[150]: CTA c owns units [ub[c], ub[c+1]) (unit = (CFG row, kv head x row tile, 128-key tile),
 * key tiles fastest); info_out[8] = {grid, cluster size (0: segment form), accumulator columns, units, key tiles per run,
 * row tiles per kv head, heads per unit, V ring depth}.  grid_cap > 0 forces the segment form on at most that many CTAs. */
int sjd_debug_attn_sw_split(int W, int n_heads, int n_kv_heads, int rows, int kv_len, const int32_t* kv_lo, int sm_count,
                            int max_cluster, int grid_cap, uint16_t* ub_out, int32_t* info_out);

/* counts kernel launches issued by this library since load (bench.py's gpu_launches) */
uint64_t sjd_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SJD_B200_H_ */
