// C-ABI of the B200-native SJD hot path (include/sjd_b200.h) + the model context that strings the
// kernels of one Jacobi draft-window forward together.  Single translation unit: the kernels live in
// the .cu files included below and are only reachable through the extern "C" functions at the bottom.
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "../../include/sjd_b200.h"
#include "attention.cu"
#include "block_ops.cu"
#include "gemm_tcgen05.cu"
#include "verify.cu"

namespace sjd {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char* what) {
  cudaError_t ce = cudaGetLastError();
  snprintf(g_err, sizeof(g_err), "%s (code %d, cuda: %s)", what, code, cudaGetErrorString(ce));
  return code;
}

static inline int round16(int m) { return (m + 15) & ~15; }

__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int W,
                                   int n, int d) {
  // dst row r = b*n + t  <-  src row b*W + (W-n+t)
  const int r = blockIdx.x, b = r / n, t = r - b * n;
  const uint4* s = reinterpret_cast<const uint4*>(src + size_t(b * W + (W - n + t)) * d);
  uint4* o = reinterpret_cast<uint4*>(dst + size_t(r) * d);
  for (int i = threadIdx.x; i < d / 8; i += blockDim.x) o[i] = s[i];
}

struct WeightMap {
  CUtensorMap map;
  int N = 0, K = 0;
  bool ok = false;
};

struct Layer {
  sjd_layer_weights w;
  WeightMap qkv, o, gate_up, down;
  bool set = false;
};

}  // namespace sjd

struct sjd_ctx {
  sjd_model_cfg cfg;
  std::vector<sjd::Layer> layers;
  const __nv_bfloat16* embed = nullptr;
  const __nv_bfloat16* final_norm = nullptr;
  const float* rope_cos = nullptr;
  const float* rope_sin = nullptr;
  sjd::WeightMap lm_head;
  // device buffers
  __nv_bfloat16 *h = nullptr, *xn = nullptr, *q = nullptr, *attn = nullptr, *act = nullptr, *xl = nullptr;
  __nv_bfloat16 *kcache = nullptr, *vcache = nullptr;
  float *ws = nullptr, *part_o = nullptr, *part_ml = nullptr;
  size_t ws_floats = 0, bytes = 0;
  int max_chunks = 0;
  // activation tensor maps per m_tile (index m_tile/16), built lazily
  CUtensorMap xmap_xn[17], xmap_attn[17], xmap_act[17], xmap_xl[17];
  bool xmap_ok[17] = {false};
};

namespace sjd {

static int dmalloc(sjd_ctx* c, void** p, size_t bytes) {
  if (cudaMalloc(p, bytes) != cudaSuccess) return SJD_E_ALLOC;
  cudaMemset(*p, 0, bytes);
  c->bytes += bytes;
  return 0;
}

static int ensure_xmaps(sjd_ctx* c, int m_tile) {
  const int idx = m_tile / 16;
  if (c->xmap_ok[idx]) return 0;
  const sjd_model_cfg& g = c->cfg;
  const int hd = g.n_heads * g.head_dim;
  if (make_tmap_bf16_2d(&c->xmap_xn[idx], c->xn, SJD_MAX_TOKENS, g.d_model, m_tile)) return SJD_E_TMAP;
  if (make_tmap_bf16_2d(&c->xmap_attn[idx], c->attn, SJD_MAX_TOKENS, hd, m_tile)) return SJD_E_TMAP;
  if (make_tmap_bf16_2d(&c->xmap_act[idx], c->act, SJD_MAX_TOKENS, g.d_ff, m_tile)) return SJD_E_TMAP;
  if (make_tmap_bf16_2d(&c->xmap_xl[idx], c->xl, SJD_MAX_TOKENS, g.d_model, m_tile)) return SJD_E_TMAP;
  c->xmap_ok[idx] = true;
  return 0;
}

static int set_wmap(WeightMap* wm, const void* ptr, int N, int K) {
  if (!ptr || K % kBlockK) return SJD_E_ARG;
  if (make_tmap_bf16_2d(&wm->map, ptr, uint64_t(N), uint64_t(K), kBlockN)) return SJD_E_TMAP;
  wm->N = N;
  wm->K = K;
  wm->ok = true;
  return 0;
}

// one GEMM of the stack: weights map x activation map -> stream-K partials in c->ws
static int run_gemm(sjd_ctx* c, const WeightMap& wm, const CUtensorMap& xmap, int m_tile, StreamK* sk_out,
                    cudaStream_t s) {
  GemmLaunch g;
  g.tmap_w = wm.map;
  g.tmap_x = xmap;
  g.sk = gemm_partition(wm.N, wm.K, m_tile, 0);
  if (g.sk.ws_floats() > c->ws_floats) return SJD_E_ARG;
  const uint32_t stage_bytes = kATileBytes + uint32_t(m_tile) * kBlockK * 2;
  int stages = int((200u * 1024u) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  g.num_stages = stages;
  uint32_t cols = 32;
  while (cols < uint32_t(2 * m_tile)) cols <<= 1;
  g.tmem_cols = cols;
  g.smem_bytes = uint32_t(stages) * stage_bytes + 1024;
  *sk_out = g.sk;
  g_launches++;
  return gemm_launch(&g, c->ws, s);
}

static int gemm_attr_once() {
  static int rc = 1;
  if (rc == 1)
    rc = cudaFuncSetAttribute(gemm_streamk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) ==
                 cudaSuccess
             ? 0
             : SJD_E_ATTR;
  return rc;
}

}  // namespace sjd

using namespace sjd;

extern "C" {

int sjd_version(void) { return 100; }
const char* sjd_last_error(void) { return g_err; }
int sjd_device_sm_count(void) { return device_num_sms(); }
uint64_t sjd_launch_count(void) { return g_launches.load(); }

size_t sjd_gemm_workspace_bytes(int N, int K, int m_tile, int grid_limit) {
  return gemm_partition(N, K, m_tile, grid_limit).ws_floats() * sizeof(float);
}

int sjd_gemm_bf16(const void* w, int N, int K, const void* x, int x_rows, int m_tile, void* ws, int grid_limit,
                  void* stream) {
  if (gemm_attr_once()) return fail(SJD_E_ATTR, "cudaFuncSetAttribute(gemm)");
  GemmLaunch g;
  int rc = gemm_prepare(&g, w, N, K, x, x_rows, m_tile, grid_limit);
  if (rc) return fail(rc, "gemm_prepare");
  g_launches++;
  rc = gemm_launch(&g, static_cast<float*>(ws), static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "gemm_launch") : 0;
}

int sjd_gemm_reduce_bf16(const void* ws, int N, int K, int m_tile, int grid_limit, void* out, int M, void* stream) {
  StreamK sk = gemm_partition(N, K, m_tile, grid_limit);
  g_launches++;
  int rc = reduce_bf16(static_cast<const float*>(ws), sk, static_cast<__nv_bfloat16*>(out), M, N,
                       static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "reduce_bf16") : 0;
}

int sjd_gemm_reduce_f32(const void* ws, int N, int K, int m_tile, int grid_limit, float* out, int M, int round_bf16,
                        void* stream) {
  StreamK sk = gemm_partition(N, K, m_tile, grid_limit);
  g_launches++;
  int rc = logits_reduce(static_cast<const float*>(ws), sk, out, M, N, round_bf16, static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "logits_reduce") : 0;
}

int sjd_verify(const sjd_verify_args* a, void* stream) {
  if (!a || !a->logits || !a->p_cur || !a->draft || !a->out_tokens || !a->out_info || !a->next_tokens || !a->resid)
    return fail(SJD_E_ARG, "sjd_verify: null argument");
  if (a->W < 1 || a->W > SJD_MAX_TOKENS || a->V < 2) return fail(SJD_E_ARG, "sjd_verify: bad W/V");
  if (a->do_sample && !a->noise_e1) return fail(SJD_E_ARG, "sjd_verify: noise_e1 required when sampling");
  if (a->scheme == 0 && a->W > 1 && (!a->noise_u || !a->noise_e2 || !a->q_row))
    return fail(SJD_E_ARG, "sjd_verify: speculative scheme needs noise_u, noise_e2, q_row");
  VerifyParams p;
  p.logits = a->logits; p.W = a->W; p.V = a->V; p.has_uncond = a->has_uncond; p.apply_cfg = a->apply_cfg;
  p.guidance = a->guidance; p.temperature = a->temperature; p.allow_lo = a->allow_lo; p.allow_hi = a->allow_hi;
  p.forced = a->forced; p.top_k = a->top_k; p.do_sample = a->do_sample; p.scheme = a->scheme; p.draft = a->draft;
  p.q_row = a->q_row; p.p_prev = a->p_prev; p.p_cur = a->p_cur; p.noise_e1 = a->noise_e1; p.noise_u = a->noise_u;
  p.noise_e2 = a->noise_e2; p.eoi_token = a->eoi_token; p.text_top_k = a->text_top_k; p.resid = a->resid;
  p.next_tokens = a->next_tokens; p.out_tokens = a->out_tokens; p.out_info = a->out_info;
  g_launches += 2;
  int rc = verify_launch(p, static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "verify_launch") : 0;
}

int sjd_ctx_create(const sjd_model_cfg* cfg, sjd_ctx** out) {
  if (!cfg || !out) return fail(SJD_E_ARG, "sjd_ctx_create: null");
  const sjd_model_cfg& g = *cfg;
  if (g.rows < 1 || g.rows > SJD_MAX_ROWS || (g.head_dim != 64 && g.head_dim != 128) || g.d_model % 64 ||
      g.d_ff % 64 || (g.n_heads * g.head_dim) % 64 || g.n_heads % g.n_kv_heads || g.max_len < 1)
    return fail(SJD_E_ARG, "sjd_ctx_create: unsupported shape");
  if (gemm_attr_once()) return fail(SJD_E_ATTR, "cudaFuncSetAttribute(gemm)");
  sjd_ctx* c = new sjd_ctx();
  c->cfg = g;
  c->layers.resize(g.n_layers);
  const int hd = g.n_heads * g.head_dim;
  const size_t T = SJD_MAX_TOKENS;
  int rc = 0;
  rc |= dmalloc(c, (void**)&c->h, T * g.d_model * 2);
  rc |= dmalloc(c, (void**)&c->xn, T * g.d_model * 2);
  rc |= dmalloc(c, (void**)&c->xl, T * g.d_model * 2);
  rc |= dmalloc(c, (void**)&c->q, T * hd * 2);
  rc |= dmalloc(c, (void**)&c->attn, T * hd * 2);
  rc |= dmalloc(c, (void**)&c->act, T * size_t(g.d_ff) * 2);
  const size_t cache_elems = size_t(g.n_layers) * g.rows * g.n_kv_heads * size_t(g.max_len) * g.head_dim;
  rc |= dmalloc(c, (void**)&c->kcache, cache_elems * 2);
  rc |= dmalloc(c, (void**)&c->vcache, cache_elems * 2);
  // workspace: largest stream-K footprint over the GEMMs of the stack at the largest m_tile
  size_t wsf = 0;
  const int Ns[5] = {(g.n_heads + 2 * g.n_kv_heads) * g.head_dim, g.d_model, 2 * g.d_ff, g.d_model, g.vocab};
  const int Ks[5] = {g.d_model, hd, g.d_model, g.d_ff, g.d_model};
  for (int i = 0; i < 5; ++i) {
    size_t f = gemm_partition(Ns[i], Ks[i], SJD_MAX_TOKENS, 0).ws_floats();
    if (f > wsf) wsf = f;
  }
  c->ws_floats = wsf;
  rc |= dmalloc(c, (void**)&c->ws, wsf * sizeof(float));
  c->max_chunks = (g.max_len + kAttnChunk - 1) / kAttnChunk;
  rc |= dmalloc(c, (void**)&c->part_o, size_t(c->max_chunks) * T * hd * sizeof(float));
  rc |= dmalloc(c, (void**)&c->part_ml, size_t(c->max_chunks) * T * g.n_heads * 2 * sizeof(float));
  if (rc) {
    sjd_ctx_destroy(c);
    return fail(SJD_E_ALLOC, "sjd_ctx_create: cudaMalloc");
  }
  *out = c;
  return 0;
}

void sjd_ctx_destroy(sjd_ctx* c) {
  if (!c) return;
  void* ptrs[] = {c->h, c->xn, c->xl, c->q, c->attn, c->act, c->kcache, c->vcache, c->ws, c->part_o, c->part_ml};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete c;
}

size_t sjd_ctx_device_bytes(const sjd_ctx* c) { return c ? c->bytes : 0; }

int sjd_ctx_set_layer(sjd_ctx* c, int layer, const sjd_layer_weights* w) {
  if (!c || !w || layer < 0 || layer >= c->cfg.n_layers) return fail(SJD_E_ARG, "sjd_ctx_set_layer: bad args");
  const sjd_model_cfg& g = c->cfg;
  Layer& L = c->layers[layer];
  L.w = *w;
  const int hd = g.n_heads * g.head_dim;
  int rc = 0;
  rc |= set_wmap(&L.qkv, w->wqkv, (g.n_heads + 2 * g.n_kv_heads) * g.head_dim, g.d_model);
  rc |= set_wmap(&L.o, w->wo, g.d_model, hd);
  rc |= set_wmap(&L.gate_up, w->w_gate_up, 2 * g.d_ff, g.d_model);
  rc |= set_wmap(&L.down, w->w_down, g.d_model, g.d_ff);
  if (rc || !w->attn_norm || !w->ffn_norm) return fail(SJD_E_TMAP, "sjd_ctx_set_layer: tensor map / null weight");
  if (g.qk_norm && (!w->q_norm_w || !w->q_norm_b || !w->k_norm_w || !w->k_norm_b))
    return fail(SJD_E_ARG, "sjd_ctx_set_layer: qk_norm weights missing");
  L.set = true;
  return 0;
}

int sjd_ctx_set_globals(sjd_ctx* c, const void* embed, const void* final_norm, const void* lm_head,
                        const float* rope_cos, const float* rope_sin) {
  if (!c || !final_norm || !lm_head || !rope_cos || !rope_sin) return fail(SJD_E_ARG, "sjd_ctx_set_globals: null");
  c->embed = static_cast<const __nv_bfloat16*>(embed);
  c->final_norm = static_cast<const __nv_bfloat16*>(final_norm);
  c->rope_cos = rope_cos;
  c->rope_sin = rope_sin;
  if (set_wmap(&c->lm_head, lm_head, c->cfg.vocab, c->cfg.d_model)) return fail(SJD_E_TMAP, "lm_head tensor map");
  return 0;
}

int sjd_ctx_forward(sjd_ctx* c, const sjd_forward_args* a, void* stream) {
  if (!c || !a) return fail(SJD_E_ARG, "sjd_ctx_forward: null");
  const sjd_model_cfg& g = c->cfg;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int W = a->W, M = g.rows * W;
  if (W < 1 || M > SJD_MAX_TOKENS) return fail(SJD_E_ARG, "sjd_ctx_forward: rows*W out of range");
  if (a->kv_len < 0 || a->kv_len + W > g.max_len) return fail(SJD_E_ARG, "sjd_ctx_forward: KV cache overflow");
  if (a->n_logit_tokens < 1 || a->n_logit_tokens > W || !a->logits || !a->rope_pos || !a->cache_pos)
    return fail(SJD_E_ARG, "sjd_ctx_forward: bad logits/pos args");
  if (!c->lm_head.ok) return fail(SJD_E_STATE, "sjd_ctx_forward: globals not set");
  for (auto& L : c->layers)
    if (!L.set) return fail(SJD_E_STATE, "sjd_ctx_forward: layer weights not set");
  if (!a->ids && !a->embeds) return fail(SJD_E_ARG, "sjd_ctx_forward: ids or embeds required");
  if (a->ids && !c->embed) return fail(SJD_E_STATE, "sjd_ctx_forward: no embedding table");
  const int m_tile = round16(M), mi = m_tile / 16;
  if (ensure_xmaps(c, m_tile)) return fail(SJD_E_TMAP, "activation tensor maps");
  const int hd = g.n_heads * g.head_dim;
  int rc = 0;
  uint64_t launches = 0;

  if (a->ids) { rc |= embed_rows(a->ids, c->embed, c->h, M, g.d_model, s); launches++; }
  else if (cudaMemcpyAsync(c->h, a->embeds, size_t(M) * g.d_model * 2, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
    return fail(SJD_E_LAUNCH, "embeds copy");
  rc |= rmsnorm_rows(c->h, static_cast<const __nv_bfloat16*>(c->layers[0].w.attn_norm), c->xn, M, g.d_model,
                     g.rms_eps, s);
  launches++;

  AttnParams ap;
  ap.q = c->q; ap.part_o = c->part_o; ap.part_ml = c->part_ml; ap.out = c->attn;
  ap.rows = g.rows; ap.W = W; ap.H = g.n_heads; ap.Hkv = g.n_kv_heads; ap.Lmax = g.max_len; ap.kv_len = a->kv_len;
  for (int b = 0; b < kAttnMaxRows; ++b) ap.kv_lo[b] = b < g.rows ? a->kv_lo[b] : 0;
  ap.n_chunks = (a->kv_len + W + kAttnChunk - 1) / kAttnChunk;
  ap.scale_log2e = 1.4426950408889634f / sqrtf(float(g.head_dim));
  const size_t layer_cache = size_t(g.rows) * g.n_kv_heads * size_t(g.max_len) * g.head_dim;

  StreamK sk;
  for (int l = 0; l < g.n_layers && !rc; ++l) {
    Layer& L = c->layers[l];
    rc |= run_gemm(c, L.qkv, c->xmap_xn[mi], m_tile, &sk, s);
    QkvPostParams qp;
    qp.ws = c->ws; qp.sk = sk; qp.q_out = c->q;
    qp.k_cache = c->kcache + size_t(l) * layer_cache; qp.v_cache = c->vcache + size_t(l) * layer_cache;
    qp.rope_pos = a->rope_pos; qp.cache_pos = a->cache_pos; qp.rope_cos = c->rope_cos; qp.rope_sin = c->rope_sin;
    qp.q_norm_w = g.qk_norm ? static_cast<const __nv_bfloat16*>(L.w.q_norm_w) : nullptr;
    qp.q_norm_b = g.qk_norm ? static_cast<const __nv_bfloat16*>(L.w.q_norm_b) : nullptr;
    qp.k_norm_w = g.qk_norm ? static_cast<const __nv_bfloat16*>(L.w.k_norm_w) : nullptr;
    qp.k_norm_b = g.qk_norm ? static_cast<const __nv_bfloat16*>(L.w.k_norm_b) : nullptr;
    qp.M = M; qp.W = W; qp.H = g.n_heads; qp.Hkv = g.n_kv_heads; qp.Lmax = g.max_len;
    qp.rope_interleaved = g.rope_interleaved;
    rc |= qkv_post(qp, g.head_dim, s);
    ap.k = qp.k_cache; ap.v = qp.v_cache;
    rc |= attn_launch(ap, g.head_dim, s);
    rc |= run_gemm(c, L.o, c->xmap_attn[mi], m_tile, &sk, s);
    rc |= reduce_residual_rmsnorm(c->ws, sk, c->h, static_cast<const __nv_bfloat16*>(L.w.ffn_norm), c->xn, M,
                                  g.d_model, g.rms_eps, s);
    rc |= run_gemm(c, L.gate_up, c->xmap_xn[mi], m_tile, &sk, s);
    rc |= silu_mul(c->ws, sk, c->act, M, g.d_ff, s);
    rc |= run_gemm(c, L.down, c->xmap_act[mi], m_tile, &sk, s);
    const void* next_norm = (l + 1 < g.n_layers) ? c->layers[l + 1].w.attn_norm : c->final_norm;
    rc |= reduce_residual_rmsnorm(c->ws, sk, c->h, static_cast<const __nv_bfloat16*>(next_norm), c->xn, M,
                                  g.d_model, g.rms_eps, s);
    launches += 6;  // 4 GEMMs are counted in run_gemm
  }
  if (rc) return fail(SJD_E_LAUNCH, "sjd_ctx_forward: layer launch");

  const int n = a->n_logit_tokens, Ml = g.rows * n;
  const int ml_tile = round16(Ml), mli = ml_tile / 16;
  if (n == W) {
    rc |= run_gemm(c, c->lm_head, c->xmap_xn[mi], m_tile, &sk, s);
  } else {
    if (ensure_xmaps(c, ml_tile)) return fail(SJD_E_TMAP, "activation tensor maps");
    gather_rows_kernel<<<Ml, 128, 0, s>>>(c->xn, c->xl, W, n, g.d_model);
    launches++;
    rc |= run_gemm(c, c->lm_head, c->xmap_xl[mli], ml_tile, &sk, s);
  }
  rc |= logits_reduce(c->ws, sk, a->logits, Ml, g.vocab, g.logits_round_bf16, s);
  launches++;
  g_launches += launches;
  if (rc || cudaGetLastError() != cudaSuccess) return fail(SJD_E_LAUNCH, "sjd_ctx_forward: head launch");
  return 0;
}

// Launch only the GEMMs of one window forward (same weights, activation buffers and launch order as
// sjd_ctx_forward) — used by bench.py to time the dominant kernel in isolation for the roofline line.
int sjd_ctx_gemm_only(sjd_ctx* c, int W, void* stream) {
  if (!c || W < 1 || c->cfg.rows * W > SJD_MAX_TOKENS) return fail(SJD_E_ARG, "sjd_ctx_gemm_only: bad args");
  if (!c->lm_head.ok) return fail(SJD_E_STATE, "sjd_ctx_gemm_only: globals not set");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int m_tile = round16(c->cfg.rows * W), mi = m_tile / 16;
  if (ensure_xmaps(c, m_tile)) return fail(SJD_E_TMAP, "activation tensor maps");
  StreamK sk;
  int rc = 0;
  for (auto& L : c->layers) {
    if (!L.set) return fail(SJD_E_STATE, "sjd_ctx_gemm_only: layer weights not set");
    rc |= run_gemm(c, L.qkv, c->xmap_xn[mi], m_tile, &sk, s);
    rc |= run_gemm(c, L.o, c->xmap_attn[mi], m_tile, &sk, s);
    rc |= run_gemm(c, L.gate_up, c->xmap_xn[mi], m_tile, &sk, s);
    rc |= run_gemm(c, L.down, c->xmap_act[mi], m_tile, &sk, s);
  }
  rc |= run_gemm(c, c->lm_head, c->xmap_xn[mi], m_tile, &sk, s);
  return rc ? fail(SJD_E_LAUNCH, "sjd_ctx_gemm_only") : 0;
}

}  // extern "C"
