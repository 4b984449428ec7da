// C-ABI of the B200-native SJD hot path (include/sjd_b200.h) + the model context that strings the
// kernels of one Jacobi draft-window forward together.  Single translation unit: the kernels live in
// the .cu files included below and are only reachable through the extern "C" functions at the bottom.
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "../../include/sjd_b200.h"
#include "attention.cu"
#include "block_ops.cu"
#include "gemm_fused.cu"
#include "attention_tc.cu"
#include "attention_tct.cu"
#include "attention_sw.cu"
#include "verify.cu"
#include "vq_lookup.cu"

namespace sjd {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char* what) {
  cudaError_t ce = cudaGetLastError();
  snprintf(g_err, sizeof(g_err), "%s (code %d, cuda: %s)", what, code, cudaGetErrorString(ce));
  return code;
}

static inline int round16(int m) { return (m + 15) & ~15; }

struct WeightMap {
  CUtensorMap map;
  int N = 0, K = 0;   // per-layer rows / reduction length
  bool ok = false;
};

}  // namespace sjd

struct sjd_ctx {
  sjd_model_cfg cfg;
  std::vector<char> layer_set;
  bool globals_set = false, have_embed = false;
  // ---- context-owned, re-laid-out weights (one buffer per projection type, layers stacked along rows) ----
  __nv_bfloat16 *wqkv = nullptr, *wo = nullptr, *wgu = nullptr, *wdown = nullptr;   // [L*N, K]
  __nv_bfloat16 *attn_norm = nullptr, *ffn_norm = nullptr;                          // [L, d]
  __nv_bfloat16 *qn_w = nullptr, *qn_b = nullptr, *kn_w = nullptr, *kn_b = nullptr; // [L, H|Hkv, Dh]
  __nv_bfloat16 *embed = nullptr, *final_norm = nullptr, *lm_head = nullptr;
  float *rope_cos = nullptr, *rope_sin = nullptr;
  sjd::WeightMap m_qkv, m_o, m_gu, m_down, m_head;   // .map copies live in maps[mi].w[0..4]
  uint32_t* fin = nullptr;                             // chain counters [kMaxChainOps + 1]
  // ---- activations / caches / workspace ----
  __nv_bfloat16 *h = nullptr, *xn = nullptr, *q = nullptr, *attn = nullptr, *act = nullptr, *xl = nullptr;
  __nv_bfloat16 *kcache = nullptr, *vcache = nullptr;
  uint8_t* ws = nullptr;
  float *part_o = nullptr, *part_ml = nullptr;
  int32_t *pos_zero = nullptr, *pos_last = nullptr;   // [SJD_MAX_TOKENS] dummies for sjd_ctx_gemm_only
  size_t ws_bytes = 0, bytes = 0;
  int max_chunks = 0, arrive_cap = 0;
  std::vector<void*> allocs;
  // activation tensor maps per m_tile (index m_tile/16), built lazily
  sjd::TmapSet maps[17];   // w[]: qkv, o, gate_up, down, lm_head ; x[]: xn, attn, act, xl (box rows = m_tile)
  bool xmap_ok[17] = {false};
  // tensor-core attention (attention_tc.cu): K / V cache maps (box 128 keys), q maps per row-slot count Wp (index Wp/8)
  sjd::AttnTcMaps tcmaps[17];
  bool tcmap_ok[17] = {false};
  int attn_mode = 0;   // 0: pick per window (see forward_chain); SJD_ATTN = tc -> 1, mma -> 2, tct -> 3, sw -> 4 force one kernel
  int sw_grid = 0;       // developer/test knob SJD_ATTN_SW_GRID: cap on the kernel's CTAs (tiny shapes then walk runs, rings and segments)
  float sw_grow = 16.f;  // developer/test knob SJD_ATTN_SW_GROW: log2 growth over the reference maximum that ends a segment
  int sw_cluster = 4;    // SJD_ATTN_SW_CLUSTER: largest cluster (CTAs per run) of the in-kernel merge; 0 = partial slots + merge pre-op
  int sw_ncols = 0;      // developer/test knob SJD_ATTN_SW_NCOLS=64: always the 64-column instantiation
  bool sw_auto = true;   // whether mode 0 may pick the segment-accumulating small-window kernel (SJD_ATTN_SW_AUTO=0: round-1 rule)
  bool tct_auto = false;   // whether mode 0 may pick the transposed small-window kernel
  bool attn_tc_ok = true;
};

namespace sjd {

template <typename T>
static int dmalloc(sjd_ctx* c, T** p, size_t bytes) {
  void* v = nullptr;
  if (cudaMalloc(&v, bytes ? bytes : 16) != cudaSuccess) return SJD_E_ALLOC;
  cudaMemset(v, 0, bytes ? bytes : 16);
  c->bytes += bytes;
  c->allocs.push_back(v);
  *p = static_cast<T*>(v);
  return 0;
}

static int ensure_xmaps(sjd_ctx* c, int m_tile) {
  const int idx = m_tile / 16;
  if (c->xmap_ok[idx]) return 0;
  const sjd_model_cfg& g = c->cfg;
  const int hd = g.n_heads * g.head_dim;
  TmapSet& t = c->maps[idx];
  t.w[0] = c->m_qkv.map; t.w[1] = c->m_o.map; t.w[2] = c->m_gu.map; t.w[3] = c->m_down.map; t.w[4] = c->m_head.map;
  if (make_tmap_bf16_2d(&t.x[0], c->xn, SJD_MAX_TOKENS, g.d_model, m_tile)) return SJD_E_TMAP;
  if (make_tmap_bf16_2d(&t.x[1], c->attn, SJD_MAX_TOKENS, hd, m_tile)) return SJD_E_TMAP;
  if (make_tmap_bf16_2d(&t.x[2], c->act, SJD_MAX_TOKENS, g.d_ff, m_tile)) return SJD_E_TMAP;
  if (make_tmap_bf16_2d(&t.x[3], c->xl, SJD_MAX_TOKENS, g.d_model, m_tile)) return SJD_E_TMAP;
  c->xmap_ok[idx] = true;
  return 0;
}

static int ensure_tcmaps(sjd_ctx* c, int Wp) {
  const int idx = Wp / 8;
  if (c->tcmap_ok[idx]) return 0;
  const sjd_model_cfg& g = c->cfg;
  AttnTcMaps& t = c->tcmaps[idx];
  const uint64_t cache_rows = uint64_t(g.n_layers) * g.rows * g.n_kv_heads * uint64_t(g.max_len);
  if (make_tmap_bf16_2d(&t.k, c->kcache, cache_rows, g.head_dim, kTcKeys)) return SJD_E_TMAP;
  if (make_tmap_bf16_2d(&t.v, c->vcache, cache_rows, g.head_dim, kTcKeys)) return SJD_E_TMAP;
  if (make_tmap_bf16_2d(&t.q, c->q, SJD_MAX_TOKENS, uint64_t(g.n_heads) * g.head_dim, uint32_t(Wp))) return SJD_E_TMAP;
  c->tcmap_ok[idx] = true;
  return 0;
}

// weights are stored unit-packed: [layers][n_groups][K/64][tpu][128][64] (pack_tiles_kernel)
static int n_groups_of(int N) { return ((N + kBlockN - 1) / kBlockN + gemm_tpu() - 1) / gemm_tpu(); }
static size_t packed_elems(int N, int K) { return size_t(n_groups_of(N)) * gemm_tpu() * kBlockN * size_t(K); }
static int set_wmap(WeightMap* wm, const void* ptr, int layers, int N, int K) {
  if (!ptr || K % kBlockK) return SJD_E_ARG;
  const uint64_t units = uint64_t(layers) * uint64_t(n_groups_of(N)) * uint64_t(K / kBlockK);
  if (make_tmap_tiled(&wm->map, ptr, units)) return SJD_E_TMAP;
  wm->N = N;
  wm->K = K;
  wm->ok = true;
  return 0;
}

// workspace pointers of one GEMM launch inside the context's workspace
static void bind_ws(sjd_ctx* c, const StreamK& sk, GemmEpi* ep) {
  const GemmWorkspace w = gemm_workspace(sk, c->arrive_cap);
  ep->ws = reinterpret_cast<float*>(c->ws + w.slots_off);
  ep->ssq = reinterpret_cast<float*>(c->ws + w.ssq_off);
  ep->tile_arrive = reinterpret_cast<uint32_t*>(c->ws + w.arrive_off);
  ep->ctr = reinterpret_cast<uint32_t*>(c->ws + w.ctr_off);
}

// A chain under construction: GEMMs are appended and flushed as ONE persistent kernel (gemm_chain_kernel)
// whenever something else (attention, the lm_head row gather) has to run in between.
struct ChainBuilder {
  sjd_ctx* c;
  int m_tile;
  cudaStream_t s;
  Chain ch;
  int rc = 0;
  ChainBuilder(sjd_ctx* c_, int m_tile_, cudaStream_t s_) : c(c_), m_tile(m_tile_), s(s_) { reset(); }
  void reset() {
    memset(&ch, 0, sizeof(ch));
    ch.fin = c->fin;
  }
  enum { W_QKV = 0, W_O = 1, W_GU = 2, W_DOWN = 3, W_HEAD = 4, X_XN = 0, X_ATTN = 1, X_ACT = 2, X_XL = 3 };
  void add(const WeightMap& wm, int wmap, int layer, int xmap, GemmEpi ep) {
    if (ch.n_ops == kMaxChainOps) flush();
    GemmOp& op = ch.ops[ch.n_ops++];
    op.wmap = wmap;
    op.xmap = xmap;
    op.sk = gemm_partition(wm.N, wm.K, m_tile, 0);
    op.w_tiled = 1;
    op.w_row0 = layer * op.sk.n_groups * op.sk.kb;   // first unit of this layer's weights
    if (gemm_workspace(op.sk, c->arrive_cap).bytes > c->ws_bytes) rc |= SJD_E_ARG;
    ep.N = wm.N;
    bind_ws(c, op.sk, &ep);
    op.ep = ep;
  }
  void flush() {
    if (ch.n_ops == 0) return;
    g_launches++;
    if (!rc) rc |= chain_launch(c->maps[m_tile / 16], ch, s);
    reset();
  }
};

}  // namespace sjd

using namespace sjd;

extern "C" {

int sjd_version(void) { return 200; }
const char* sjd_last_error(void) { return g_err; }
int sjd_device_sm_count(void) { return device_num_sms(); }
uint64_t sjd_launch_count(void) { return g_launches.load(); }

static long long* g_attn_dbg = nullptr;
void sjd_debug_attn_stamps(void* device_buf) { g_attn_dbg = static_cast<long long*>(device_buf); }

void sjd_debug_gemm_stamps(void* device_buf, int n_launches) {
  g_dbg_buf = static_cast<long long*>(device_buf);
  g_dbg_cap = n_launches;
  g_dbg_idx = 0;
}

int sjd_vq_lookup(const int32_t* codes, int n_pix, int hw, const float* codebook, int n_e, int e_dim, int l2_norm,
                  const float* w, const float* bias, int z, float* out, void* stream) {
  if (!codes || !codebook || !w || !bias || !out || n_pix < 1 || hw < 1 || n_pix % hw || n_e < 1 || e_dim < 1 || z < 1)
    return fail(SJD_E_ARG, "sjd_vq_lookup: bad argument");
  const int rc = vq_lookup_launch(codes, n_pix, hw, codebook, n_e, e_dim, l2_norm, w, bias, z, out, static_cast<cudaStream_t>(stream));
  if (rc) return fail(rc == -3 ? SJD_E_ARG : SJD_E_LAUNCH, "sjd_vq_lookup: launch");
  g_launches++;
  return 0;
}

int sjd_debug_attn_sw_split(int W, int n_heads, int n_kv_heads, int rows, int kv_len, const int32_t* kv_lo, int sm_count,
                            int max_cluster, int grid_cap, uint16_t* ub_out, int32_t* info_out) {
  if (W < 1 || ((W + 7) & ~7) > kTctCols || n_heads < 1 || n_kv_heads < 1 || n_heads % n_kv_heads || rows < 1 ||
      rows > kAttnMaxRows || kv_len < 0 || !kv_lo || sm_count < 1 || !ub_out || !info_out)
    return fail(SJD_E_ARG, "sjd_debug_attn_sw_split: bad argument");
  AttnSwParams sp;
  memset(&sp, 0, sizeof(sp));
  AttnParams& a = sp.t.a;
  a.rows = rows; a.W = W; a.H = n_heads; a.Hkv = n_kv_heads; a.kv_len = kv_len;
  for (int b = 0; b < kAttnMaxRows; ++b) a.kv_lo[b] = b < rows ? kv_lo[b] : 0;
  attn_sw_plan(&sp, 0, max_cluster, sm_count);
  sp.grid_cap = grid_cap;
  const int n_units = a.n_chunks * a.Hkv * sp.t.mtiles * a.rows;
  if (n_units >= 65536) return fail(SJD_E_ARG, "sjd_debug_attn_sw_split: too many units");
  const int ng = attn_sw_grid(sp, sm_count);
  for (int c = 0; c < 150; ++c) ub_out[c] = sp.ub[c];
  info_out[0] = ng; info_out[1] = sp.cluster; info_out[2] = sp.ncols; info_out[3] = n_units;
  info_out[4] = a.n_chunks; info_out[5] = sp.t.mtiles; info_out[6] = sp.t.hpc; info_out[7] = sp.nv;
  return 0;
}

size_t sjd_gemm_workspace_bytes(int N, int K, int m_tile, int grid_limit) {
  return gemm_workspace(gemm_partition(N, K, m_tile, grid_limit)).bytes;
}

int sjd_gemm_bf16(const void* w, int N, int K, const void* x, int x_rows, int M, void* out, int out_f32,
                  int round_bf16, void* ws, int grid_limit, void* stream) {
  if (!w || !x || !out || !ws || M < 1 || M > SJD_MAX_TOKENS) return fail(SJD_E_ARG, "sjd_gemm_bf16: bad argument");
  const int m_tile = round16(M);
  if (K % kBlockK != 0 || x_rows < m_tile) return fail(SJD_E_ARG, "sjd_gemm_bf16: K % 64 != 0 or x has < m_tile rows");
  if (gemm_attr_once()) return fail(SJD_E_ATTR, "cudaFuncSetAttribute(gemm)");
  const StreamK sk = gemm_partition(N, K, m_tile, grid_limit);
  TmapSet maps;
  memset(&maps, 0, sizeof(maps));
  if (make_tmap_bf16_2d(&maps.w[0], w, uint64_t(N), uint64_t(K), uint32_t(kBlockN * (gemm_tpu() < 2 ? gemm_tpu() : 2))))
    return fail(SJD_E_TMAP, "weight tensor map");
  if (make_tmap_bf16_2d(&maps.x[0], x, uint64_t(x_rows), uint64_t(K), uint32_t(m_tile)))
    return fail(SJD_E_TMAP, "activation tensor map");
  Chain ch;
  memset(&ch, 0, sizeof(ch));
  GemmEpi& ep = ch.ops[0].ep;
  ep.mode = out_f32 ? EPI_F32 : EPI_BF16;
  ep.M = M;
  ep.N = N;
  ep.out = out;
  ep.ld_out = N;
  ep.round_bf16 = round_bf16;
  const GemmWorkspace wl = gemm_workspace(sk);
  uint8_t* b = static_cast<uint8_t*>(ws);
  ep.ws = reinterpret_cast<float*>(b + wl.slots_off);
  ep.ssq = reinterpret_cast<float*>(b + wl.ssq_off);
  ep.tile_arrive = reinterpret_cast<uint32_t*>(b + wl.arrive_off);
  ep.ctr = reinterpret_cast<uint32_t*>(b + wl.ctr_off);
  ch.fin = reinterpret_cast<uint32_t*>(b + wl.ctr_off) + 2 * kCtrStride;   // ctr region: 2 KB = 8 counters of 256 B (2 + 6 used)
  ch.n_ops = 1;
  ch.ops[0].sk = sk;
  g_launches++;
  const int rc = chain_launch(maps, ch, static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "chain_launch") : 0;
}

int sjd_verify(const sjd_verify_args* a, void* stream) {
  if (!a || !a->logits || !a->p_cur || !a->draft || !a->out_tokens || !a->out_info || !a->next_tokens || !a->resid)
    return fail(SJD_E_ARG, "sjd_verify: null argument");
  if (a->W < 1 || a->W > SJD_MAX_TOKENS || a->V < 2) return fail(SJD_E_ARG, "sjd_verify: bad W/V");
  if (!(a->top_p_thresh >= 0.f && a->top_p_thresh <= 1.f)) return fail(SJD_E_ARG, "sjd_verify: top_p_thresh outside [0, 1]");
  if (a->rng_mode != 0 && a->rng_mode != 1) return fail(SJD_E_ARG, "sjd_verify: rng_mode");
  if (a->rng_mode == 1 && (!a->rng_span[0] || !a->rng_span[1] || !a->rng_span[2] || (a->rng_off[0] & 3) ||
                           (a->rng_off[1] & 3) || (a->rng_off[2] & 3)))
    return fail(SJD_E_ARG, "sjd_verify: rng_span must be non-zero and rng_off multiples of 4");
  if (a->do_sample && !a->rng_mode && !a->noise_e1) return fail(SJD_E_ARG, "sjd_verify: noise_e1 required when sampling");
  if (a->scheme == 0 && a->W > 1 && (!a->q_row || (!a->rng_mode && (!a->noise_u || !a->noise_e2))))
    return fail(SJD_E_ARG, "sjd_verify: speculative scheme needs noise_u, noise_e2, q_row");
  VerifyParams p;
  p.logits = a->logits; p.W = a->W; p.V = a->V; p.has_uncond = a->has_uncond; p.apply_cfg = a->apply_cfg;
  p.guidance = a->guidance; p.temperature = a->temperature; p.allow_lo = a->allow_lo; p.allow_hi = a->allow_hi;
  p.forced = a->forced; p.forced_resid = a->forced_resid; p.top_k = a->top_k; p.top_p_thresh = a->top_p_thresh; p.do_sample = a->do_sample; p.scheme = a->scheme; p.draft = a->draft;
  p.q_row = a->q_row; p.p_prev = a->p_prev; p.p_cur = a->p_cur; p.noise_e1 = a->noise_e1; p.noise_u = a->noise_u;
  p.noise_e2 = a->noise_e2; p.eoi_token = a->eoi_token; p.text_top_k = a->text_top_k; p.resid = a->resid;
  p.next_tokens = a->next_tokens; p.out_tokens = a->out_tokens; p.out_info = a->out_info;
  p.sync_ws = a->sync_ws;
  if (a->allow_mode < 0 || a->allow_mode > 4 || (a->allow_mode == 3 && (a->ban[0] < 0 || a->ban[1] < 0)) ||
      (a->allow_mode == 2 && a->allow_hi < a->allow_lo))
    return fail(SJD_E_ARG, "sjd_verify: allow_mode / ban");
  p.allow_mode = a->allow_mode; p.ban[0] = a->ban[0]; p.ban[1] = a->ban[1];
  if (a->resid_set && (a->resid_allow_mode < 0 || a->resid_allow_mode > 4)) return fail(SJD_E_ARG, "sjd_verify: resid_allow_mode");
  p.resid_set = a->resid_set; p.resid_allow_mode = a->resid_allow_mode; p.resid_allow_lo = a->resid_allow_lo;
  p.resid_allow_hi = a->resid_allow_hi; p.resid_ban[0] = a->resid_ban[0]; p.resid_ban[1] = a->resid_ban[1];
  p.resid_from = a->resid_from;
  if (a->done_flag && !a->sync_ws) return fail(SJD_E_ARG, "sjd_verify: done_flag needs sync_ws (the one-launch form)");
  p.done_flag = a->done_flag; p.done_seq = a->done_seq;
  p.rng_mode = a->rng_mode; p.rng_seed = a->rng_seed;
  for (int k = 0; k < 3; ++k) { p.rng_off[k] = a->rng_off[k]; p.rng_span[k] = a->rng_span[k]; }
  g_launches += a->sync_ws ? 1 : 2;
  int rc = verify_launch(p, static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "verify_launch") : 0;
}

int sjd_debug_philox(float* out, uint64_t numel, uint64_t seed, uint64_t offset, uint32_t span, int kind, void* stream) {
  if (!out || !span || (offset & 3)) return fail(SJD_E_ARG, "sjd_debug_philox: bad argument");
  const int rc = philox_fill(out, numel, seed, offset, span, kind, static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "philox_fill") : 0;
}

int sjd_ctx_create(const sjd_model_cfg* cfg, sjd_ctx** out) {
  if (!cfg || !out) return fail(SJD_E_ARG, "sjd_ctx_create: null");
  const sjd_model_cfg& g = *cfg;
  const int hd = g.n_heads * g.head_dim;
  const int nqkv = (g.n_heads + 2 * g.n_kv_heads) * g.head_dim;
  if (g.rows < 1 || g.rows > SJD_MAX_ROWS || (g.head_dim != 64 && g.head_dim != 128) || g.d_model % 128 ||
      g.d_ff % 64 || hd % 64 || nqkv % 128 || g.n_heads % g.n_kv_heads || g.max_len < 1 || g.n_layers < 1 ||
      g.n_rope_pos < 1)
    return fail(SJD_E_ARG, "sjd_ctx_create: unsupported shape");
  if (!g.rope_interleaved && g.head_dim != 128)
    return fail(SJD_E_ARG, "sjd_ctx_create: rotate-half RoPE needs head_dim 128");
  if (g.qk_norm && (g.rope_interleaved || g.head_dim != 128))
    return fail(SJD_E_ARG, "sjd_ctx_create: qk_norm needs head_dim 128 with rotate-half RoPE");
  if (gemm_attr_once()) return fail(SJD_E_ATTR, "cudaFuncSetAttribute(gemm)");
  sjd_ctx* c = new sjd_ctx();
  c->cfg = g;
  c->layer_set.assign(g.n_layers, 0);
  const size_t T = SJD_MAX_TOKENS, L = g.n_layers, d = g.d_model;
  int rc = 0;
  rc |= dmalloc(c, &c->wqkv, L * packed_elems(nqkv, g.d_model) * 2);
  rc |= dmalloc(c, &c->wo, L * packed_elems(g.d_model, hd) * 2);
  rc |= dmalloc(c, &c->wgu, L * packed_elems(2 * g.d_ff, g.d_model) * 2);
  rc |= dmalloc(c, &c->wdown, L * packed_elems(g.d_model, g.d_ff) * 2);
  rc |= dmalloc(c, &c->attn_norm, L * d * 2);
  rc |= dmalloc(c, &c->ffn_norm, L * d * 2);
  if (g.qk_norm) {
    rc |= dmalloc(c, &c->qn_w, L * hd * 2);
    rc |= dmalloc(c, &c->qn_b, L * hd * 2);
    rc |= dmalloc(c, &c->kn_w, L * size_t(g.n_kv_heads) * g.head_dim * 2);
    rc |= dmalloc(c, &c->kn_b, L * size_t(g.n_kv_heads) * g.head_dim * 2);
  }
  rc |= dmalloc(c, &c->embed, size_t(g.vocab) * d * 2);
  rc |= dmalloc(c, &c->lm_head, packed_elems(g.vocab, g.d_model) * 2);
  rc |= dmalloc(c, &c->final_norm, d * 2);
  rc |= dmalloc(c, &c->rope_cos, size_t(g.n_rope_pos) * (g.head_dim / 2) * 4);
  rc |= dmalloc(c, &c->rope_sin, size_t(g.n_rope_pos) * (g.head_dim / 2) * 4);
  rc |= dmalloc(c, &c->h, T * d * 2);
  rc |= dmalloc(c, &c->xn, T * d * 2);
  rc |= dmalloc(c, &c->xl, T * d * 2);
  rc |= dmalloc(c, &c->q, T * hd * 2);
  rc |= dmalloc(c, &c->attn, T * hd * 2);
  rc |= dmalloc(c, &c->act, T * size_t(g.d_ff) * 2);
  const size_t cache_elems = L * g.rows * g.n_kv_heads * size_t(g.max_len) * g.head_dim;
  rc |= dmalloc(c, &c->kcache, cache_elems * 2);
  rc |= dmalloc(c, &c->vcache, cache_elems * 2);
  // workspace: largest footprint over the GEMMs of the stack at the largest m_tile (counters start at zero)
  size_t wsb = 0;
  const int Ns[5] = {nqkv, g.d_model, 2 * g.d_ff, g.d_model, g.vocab};
  const int Ks[5] = {g.d_model, hd, g.d_model, g.d_ff, g.d_model};
  for (int i = 0; i < 5; ++i) c->arrive_cap = std::max(c->arrive_cap, (Ns[i] + kBlockN - 1) / kBlockN);
  for (int i = 0; i < 5; ++i) {
    const size_t b = gemm_workspace(gemm_partition(Ns[i], Ks[i], SJD_MAX_TOKENS, 0), c->arrive_cap).bytes;
    if (b > wsb) wsb = b;
  }
  c->ws_bytes = wsb;
  rc |= dmalloc(c, &c->ws, wsb);
  c->max_chunks = (g.max_len + kAttnSub - 1) / kAttnSub;   // upper bound on key splits
  {
    const char* e = getenv("SJD_ATTN");   // "tc": tcgen05 attention (attention_tc.cu); "mma": mma.sync kernel (attention.cu)
    c->attn_mode = !e ? 0 : (strcmp(e, "tc") == 0 ? 1 : (strcmp(e, "mma") == 0 ? 2 : (strcmp(e, "tct") == 0 ? 3 : (strcmp(e, "sw") == 0 ? 4 : 0))));
    if (const char* e2 = getenv("SJD_ATTN_SW_AUTO")) c->sw_auto = atoi(e2) != 0;
    if (const char* e2 = getenv("SJD_ATTN_SW_GRID")) c->sw_grid = atoi(e2);
    if (const char* e2 = getenv("SJD_ATTN_SW_CLUSTER")) c->sw_cluster = atoi(e2);
    if (const char* e2 = getenv("SJD_ATTN_SW_NCOLS")) c->sw_ncols = atoi(e2);
    if (const char* e2 = getenv("SJD_ATTN_SW_GROW")) c->sw_grow = float(atof(e2));
    if (uint64_t(g.n_layers) * g.rows * g.n_kv_heads * uint64_t(g.max_len) >= (1ull << 31)) c->attn_tc_ok = false;
  }
  rc |= dmalloc(c, &c->part_o, size_t(c->max_chunks) * T * hd * sizeof(float));
  rc |= dmalloc(c, &c->part_ml, size_t(c->max_chunks) * T * g.n_heads * 2 * sizeof(float));
  rc |= dmalloc(c, &c->fin, (kMaxChainOps + 2) * kCtrStride * 4);
  rc |= dmalloc(c, &c->pos_zero, T * 4);
  rc |= dmalloc(c, &c->pos_last, T * 4);
  if (!rc) {
    std::vector<int32_t> last(T, g.max_len - 1);
    cudaMemcpy(c->pos_last, last.data(), T * 4, cudaMemcpyHostToDevice);
    rc |= set_wmap(&c->m_qkv, c->wqkv, g.n_layers, nqkv, g.d_model);
    rc |= set_wmap(&c->m_o, c->wo, g.n_layers, g.d_model, hd);
    rc |= set_wmap(&c->m_gu, c->wgu, g.n_layers, 2 * g.d_ff, g.d_model);
    rc |= set_wmap(&c->m_down, c->wdown, g.n_layers, g.d_model, g.d_ff);
    rc |= set_wmap(&c->m_head, c->lm_head, 1, g.vocab, g.d_model);
  }
  if (rc) {
    sjd_ctx_destroy(c);
    return fail(SJD_E_ALLOC, "sjd_ctx_create: cudaMalloc / tensor maps");
  }
  *out = c;
  return 0;
}

void sjd_ctx_destroy(sjd_ctx* c) {
  if (!c) return;
  for (void* p : c->allocs)
    if (p) cudaFree(p);
  delete c;
}

size_t sjd_ctx_device_bytes(const sjd_ctx* c) { return c ? c->bytes : 0; }

int sjd_ctx_set_layer(sjd_ctx* c, int layer, const sjd_layer_weights* w) {
  if (!c || !w || layer < 0 || layer >= c->cfg.n_layers) return fail(SJD_E_ARG, "sjd_ctx_set_layer: bad args");
  const sjd_model_cfg& g = c->cfg;
  if (!w->attn_norm || !w->ffn_norm || !w->wqkv || !w->wo || !w->w_gate_up || !w->w_down)
    return fail(SJD_E_ARG, "sjd_ctx_set_layer: null weight");
  if (g.qk_norm && (!w->q_norm_w || !w->q_norm_b || !w->k_norm_w || !w->k_norm_b))
    return fail(SJD_E_ARG, "sjd_ctx_set_layer: qk_norm weights missing");
  const size_t d = g.d_model, hd = size_t(g.n_heads) * g.head_dim, kvd = size_t(g.n_kv_heads) * g.head_dim;
  const size_t nqkv = hd + 2 * kvd, l = layer;
  cudaError_t e = cudaSuccess;
  auto cp = [&](void* dst, const void* src, size_t bytes) {
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, 0);
  };
  int prc = 0;
  auto bfp = [](const void* p) { return static_cast<const __nv_bfloat16*>(p); };
  const int tpu = gemm_tpu();
  prc |= pack_tiles(bfp(w->wqkv), c->wqkv + l * packed_elems(int(nqkv), g.d_model), int(nqkv), g.d_model, 0, tpu, 0);
  prc |= pack_tiles(bfp(w->wo), c->wo + l * packed_elems(g.d_model, int(hd)), g.d_model, int(hd), 0, tpu, 0);
  prc |= pack_tiles(bfp(w->w_down), c->wdown + l * packed_elems(g.d_model, g.d_ff), g.d_model, g.d_ff, 0, tpu, 0);
  prc |= pack_tiles(bfp(w->w_gate_up), c->wgu + l * packed_elems(2 * g.d_ff, g.d_model), 2 * g.d_ff, g.d_model, g.d_ff, tpu, 0);
  if (prc) return fail(SJD_E_LAUNCH, "sjd_ctx_set_layer: pack_tiles");
  cp(c->attn_norm + l * d, w->attn_norm, d * 2);
  cp(c->ffn_norm + l * d, w->ffn_norm, d * 2);
  if (g.qk_norm) {
    cp(c->qn_w + l * hd, w->q_norm_w, hd * 2);
    cp(c->qn_b + l * hd, w->q_norm_b, hd * 2);
    cp(c->kn_w + l * kvd, w->k_norm_w, kvd * 2);
    cp(c->kn_b + l * kvd, w->k_norm_b, kvd * 2);
  }
  if (e != cudaSuccess) return fail(SJD_E_LAUNCH, "sjd_ctx_set_layer: copy");
  if (cudaStreamSynchronize(0) != cudaSuccess) return fail(SJD_E_LAUNCH, "sjd_ctx_set_layer: sync");
  c->layer_set[layer] = 1;
  return 0;
}

int sjd_ctx_set_globals(sjd_ctx* c, const void* embed, const void* final_norm, const void* lm_head,
                        const float* rope_cos, const float* rope_sin) {
  if (!c || !final_norm || !lm_head || !rope_cos || !rope_sin) return fail(SJD_E_ARG, "sjd_ctx_set_globals: null");
  const sjd_model_cfg& g = c->cfg;
  const size_t vd = size_t(g.vocab) * g.d_model * 2, rb = size_t(g.n_rope_pos) * (g.head_dim / 2) * 4;
  cudaError_t e = cudaSuccess;
  if (embed) e = cudaMemcpy(c->embed, embed, vd, cudaMemcpyDeviceToDevice);
  if (e == cudaSuccess && pack_tiles(static_cast<const __nv_bfloat16*>(lm_head), c->lm_head, g.vocab, g.d_model, 0, gemm_tpu(), 0))
    e = cudaErrorUnknown;
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(c->final_norm, final_norm, size_t(g.d_model) * 2, cudaMemcpyDeviceToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(c->rope_cos, rope_cos, rb, cudaMemcpyDeviceToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(c->rope_sin, rope_sin, rb, cudaMemcpyDeviceToDevice);
  if (e != cudaSuccess) return fail(SJD_E_LAUNCH, "sjd_ctx_set_globals: copy");
  c->have_embed = embed != nullptr;
  c->globals_set = true;
  return 0;
}

static int ctx_ready(sjd_ctx* c) {
  if (!c->globals_set) return fail(SJD_E_STATE, "globals not set");
  for (char s : c->layer_set)
    if (!s) return fail(SJD_E_STATE, "layer weights not set");
  return 0;
}

// The layer stack of one window forward.  Per layer: [QKV] -> attention -> [O, GATE_UP, DOWN, next layer's QKV];
// the bracketed groups run as one persistent chain kernel each.  `head` (lm_head epilogue) joins the last chain
// when it can consume xn directly.  gemm_only skips attention (bench roofline leg).
static int forward_chain(sjd_ctx* c, int W, const sjd_forward_args* a, bool gemm_only, const GemmEpi* head,
                         cudaStream_t s) {
  const sjd_model_cfg& g = c->cfg;
  const int M = g.rows * W, m_tile = round16(M);
  const int hd = g.n_heads * g.head_dim;
  int rc = 0;
  AttnParams ap;
  memset(&ap, 0, sizeof(ap));
  if (!gemm_only) {
    ap.q = c->q; ap.part_o = c->part_o; ap.part_ml = c->part_ml; ap.out = c->attn;
    ap.rows = g.rows; ap.W = W; ap.H = g.n_heads; ap.Hkv = g.n_kv_heads; ap.Lmax = g.max_len; ap.kv_len = a->kv_len;
    for (int b = 0; b < kAttnMaxRows; ++b) ap.kv_lo[b] = b < g.rows ? a->kv_lo[b] : 0;
    attn_plan(&ap, device_num_sms());
    static const int l2pf = getenv("SJD_ATTN_L2PF") ? atoi(getenv("SJD_ATTN_L2PF")) : 1;
    ap.l2_prefetch = l2pf;
    ap.scale_log2e = 1.4426950408889634f / sqrtf(float(g.head_dim));
  }
  const size_t layer_cache = size_t(g.rows) * g.n_kv_heads * size_t(g.max_len) * g.head_dim;
  const size_t kvd = size_t(g.n_kv_heads) * g.head_dim;
  // Two attention kernels.  The tcgen05 one wins when the query rows that share a kv head fill
  // between half and all of one 128-row UMMA tile (measured: Lumina W=64, Emu3 W=32, profiles/r01g_config_sweep*); with
  // fewer rows its softmax threads idle or duplicate work, with more the K/V tile is streamed once per row tile.
  bool use_tc = false, use_tct = false, use_sw = false;
  if (!gemm_only && c->attn_tc_ok && W <= kTcRows) {
    const int rows_per_kv = (g.n_heads / g.n_kv_heads) * ((W + 7) & ~7);
    const bool tct_fits = g.head_dim == 128 && ((W + 7) & ~7) <= kTctCols;
    // round 2: the segment-accumulating small-window kernel (attention_sw.cu) takes every head-dim-128 window of up to
    // 64 tokens: Lumina / Chameleon (18.5 vs 22.7 us per layer at window 32, 24.4 vs 28.9 at 64) and the GQA shapes too,
    // although a kv head's query rows then need several row tiles that re-stream the K/V tile from L2 (Emu3 window 64
    // over 4 096 / 8 000 keys: 5.58 / 6.25 ms per forward vs 6.52 / 7.93; profiles/r02u_attn_sw.txt)
    use_sw = tct_fits && (c->attn_mode == 4 || (c->attn_mode == 0 && c->sw_auto));
    use_tct = !use_sw && tct_fits && (c->attn_mode == 3 || (c->attn_mode == 0 && rows_per_kv <= kTctCols && c->tct_auto));
    use_tc = !use_sw && !use_tct && (c->attn_mode == 1 || (c->attn_mode == 0 && rows_per_kv >= 64 && rows_per_kv <= kTcRows));
  }
  AttnSwParams swp;
  memset(&swp, 0, sizeof(swp));
  AttnTcParams& tp = swp.t;
  if (use_tc || use_tct || use_sw) {
    tp.a = ap;
    // the tcgen05 kernels' "whole span to L2 first" pass is neutral to slightly negative on the B200
    // (profiles/r02k_attn_tc_l2pf.txt): off unless asked for
    static const int tc_l2pf = getenv("SJD_ATTN_TC_L2PF") ? atoi(getenv("SJD_ATTN_TC_L2PF")) : 0;
    tp.a.l2_prefetch = tc_l2pf;
    if (use_sw) {
      attn_sw_plan(&swp, c->sw_ncols, c->sw_grid > 0 ? 0 : c->sw_cluster);
      swp.grid_cap = c->sw_grid;
      swp.grow = c->sw_grow;
    } else if (use_tct) attn_tct_plan(&tp);
    else attn_tc_plan(&tp, g.head_dim);
    if (tp.a.n_chunks > c->max_chunks || ensure_tcmaps(c, tp.Wp)) return SJD_E_TMAP;
  }
  GemmEpi base;
  memset(&base, 0, sizeof(base));
  base.M = M;
  ChainBuilder cb(c, m_tile, s);
  auto qkv_epi = [&](int l) {
    GemmEpi e = base;
    e.mode = EPI_QKV;
    e.q_out = c->q;
    e.k_cache = c->kcache + size_t(l) * layer_cache;
    e.v_cache = c->vcache + size_t(l) * layer_cache;
    e.rope_pos = a->rope_pos; e.cache_pos = a->cache_pos; e.rope_cos = c->rope_cos; e.rope_sin = c->rope_sin;
    if (g.qk_norm) {
      e.q_norm_w = c->qn_w + size_t(l) * hd; e.q_norm_b = c->qn_b + size_t(l) * hd;
      e.k_norm_w = c->kn_w + size_t(l) * kvd; e.k_norm_b = c->kn_b + size_t(l) * kvd;
    }
    e.W = W; e.H = g.n_heads; e.Hkv = g.n_kv_heads; e.Lmax = g.max_len; e.Dh = g.head_dim;
    e.rope_interleaved = g.rope_interleaved;
    return e;
  };
  cb.add(c->m_qkv, ChainBuilder::W_QKV, 0, ChainBuilder::X_XN, qkv_epi(0));
  // Sending the next attention's K/V span to L2 from the chain's producer (SJD_KV_PF=1) was measured on the B200 and is
  // OFF: attention + boundaries went from 22.9 to 40.4 us per layer (W = 32, 1 200 keys; profiles/r02c_chain_experiments.txt)
  // — 40 MB of bulk prefetches drain slower than the attention's own demand loads, which then queue behind them.
  static const int kv_pf = getenv("SJD_KV_PF") ? atoi(getenv("SJD_KV_PF")) : 0;
  for (int l = 0; l < g.n_layers && !rc; ++l) {
    if (!gemm_only) {
      if (kv_pf && a->kv_len > 0) {   // the chain about to be flushed ends right before layer l's attention
        KvPrefetch& kp = cb.ch.kvpf;
        kp.k = c->kcache + size_t(l) * layer_cache;
        kp.v = c->vcache + size_t(l) * layer_cache;
        kp.rows = g.rows; kp.Hkv = g.n_kv_heads; kp.Lmax = g.max_len; kp.Dh = g.head_dim; kp.kv_len = a->kv_len;
        for (int b = 0; b < 8; ++b) kp.lo[b] = b < g.rows ? a->kv_lo[b] : 0;
      }
      cb.flush();
      ap.k = c->kcache + size_t(l) * layer_cache;
      ap.v = c->vcache + size_t(l) * layer_cache;
      // developer timing (scripts/chain_time.py): SJD_DEBUG_ATTN=1 keeps the chain split here but launches no attention,
      // which separates the cost of the two kernel boundaries from the cost of the attention kernel itself
      static const int dbg_attn = getenv("SJD_DEBUG_ATTN") ? atoi(getenv("SJD_DEBUG_ATTN")) : 0;
      if (dbg_attn != 1) {
        if (use_tc || use_tct || use_sw) {
          tp.a.k = ap.k; tp.a.v = ap.v;
          tp.k_row0 = int(size_t(l) * g.rows * g.n_kv_heads * size_t(g.max_len));
          tp.dbg = g_attn_dbg;
          if (use_sw && (dbg_attn == 2 || dbg_attn == 4)) swp.grid_cap = -1;   // (developer timing: the kernel returns after its dependency wait)
          rc |= use_sw ? attn_sw_launch(c->tcmaps[tp.Wp / 8], swp, s)
                       : (use_tct ? attn_tct_launch(c->tcmaps[tp.Wp / 8], tp, s) : attn_tc_launch(c->tcmaps[tp.Wp / 8], tp, s));
          cb.ch.pre = attn_combine_desc(tp.a, g.head_dim);
          cb.ch.pre.sparse = use_sw ? 1 : 0;
          if (use_sw && swp.cluster > 0) cb.ch.pre.n_chunks = 0;   // merged and normalised by the attention kernel's cluster leaders
          if (dbg_attn == 3 || dbg_attn == 4) cb.ch.pre.n_chunks = 0;   // (developer timing: no split merge)
        } else {
          rc |= attn_launch(ap, g.head_dim, false, s);   // the split merge rides in the next chain kernel (pre-op)
          cb.ch.pre = attn_combine_desc(ap, g.head_dim);
        }
        g_launches += 1;
      }
    }
    GemmEpi e = base;
    e.mode = EPI_RESID_NORM;
    e.h = c->h; e.xn = c->xn; e.eps = g.rms_eps;
    e.norm_w = c->ffn_norm + size_t(l) * g.d_model;
    cb.add(c->m_o, ChainBuilder::W_O, l, ChainBuilder::X_ATTN, e);
    e = base;
    e.mode = EPI_SILU_MUL;
    e.out = c->act; e.ld_out = g.d_ff;
    cb.add(c->m_gu, ChainBuilder::W_GU, l, ChainBuilder::X_XN, e);
    e = base;
    e.mode = EPI_RESID_NORM;
    e.h = c->h; e.xn = c->xn; e.eps = g.rms_eps;
    e.norm_w = (l + 1 < g.n_layers) ? c->attn_norm + size_t(l + 1) * g.d_model : c->final_norm;
    cb.add(c->m_down, ChainBuilder::W_DOWN, l, ChainBuilder::X_ACT, e);
    if (l + 1 < g.n_layers) cb.add(c->m_qkv, ChainBuilder::W_QKV, l + 1, ChainBuilder::X_XN, qkv_epi(l + 1));
    else if (head) cb.add(c->m_head, ChainBuilder::W_HEAD, 0, ChainBuilder::X_XN, *head);
  }
  cb.flush();
  return rc | cb.rc;
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
  {
    // rope_pos lives on the device; under the engine's convention (position = cache slot - first visible key) the
    // largest index of this call is kv_len + W - 1 - min(kv_lo): reject calls that would read past the RoPE table
    // (the reference fails loudly on freqs_cis[input_pos], llamagen/llamagen.py:386)
    int lo = a->kv_lo[0];
    for (int b = 1; b < g.rows; ++b) lo = a->kv_lo[b] < lo ? a->kv_lo[b] : lo;
    if (lo < 0 || a->kv_len + W - lo > g.n_rope_pos) return fail(SJD_E_ARG, "sjd_ctx_forward: RoPE table too small for these positions");
  }
  if (ctx_ready(c)) return SJD_E_STATE;
  if (!a->ids && !a->embeds) return fail(SJD_E_ARG, "sjd_ctx_forward: ids or embeds required");
  if (a->ids && !c->have_embed) return fail(SJD_E_STATE, "sjd_ctx_forward: no embedding table");
  const int m_tile = round16(M);
  if (ensure_xmaps(c, m_tile)) return fail(SJD_E_TMAP, "activation tensor maps");
  int rc = embed_rmsnorm_rows(a->ids, c->embed, static_cast<const __nv_bfloat16*>(a->embeds), c->h, c->attn_norm,
                              c->xn, M, g.d_model, g.rms_eps, s);
  g_launches++;
  const int n = a->n_logit_tokens, Ml = g.rows * n;
  GemmEpi e;
  memset(&e, 0, sizeof(e));
  e.mode = EPI_F32;
  e.M = Ml;
  e.out = a->logits;
  e.ld_out = g.vocab;
  e.round_bf16 = g.logits_round_bf16;
  if (n == W) {
    rc |= forward_chain(c, W, a, false, &e, s);   // lm_head rides in the last chain
  } else {
    rc |= forward_chain(c, W, a, false, nullptr, s);
    const int ml_tile = round16(Ml);
    if (ensure_xmaps(c, ml_tile)) return fail(SJD_E_TMAP, "activation tensor maps");
    rc |= gather_rows(c->xn, c->xl, g.rows, W, n, g.d_model, s);
    g_launches++;
    ChainBuilder cb(c, ml_tile, s);
    cb.add(c->m_head, ChainBuilder::W_HEAD, 0, ChainBuilder::X_XL, e);
    cb.flush();
    rc |= cb.rc;
  }
  if (rc || cudaGetLastError() != cudaSuccess) return fail(SJD_E_LAUNCH, "sjd_ctx_forward: launch");
  return 0;
}

// Launches only the fused GEMMs of one window forward (same weights, buffers, epilogues and launch order as
// sjd_ctx_forward, attention skipped) — used by bench.py to time the dominant kernel in isolation.
int sjd_ctx_gemm_only(sjd_ctx* c, int W, void* stream) {
  if (!c || W < 1 || c->cfg.rows * W > SJD_MAX_TOKENS) return fail(SJD_E_ARG, "sjd_ctx_gemm_only: bad args");
  if (ctx_ready(c)) return SJD_E_STATE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const sjd_model_cfg& g = c->cfg;
  const int M = g.rows * W, m_tile = round16(M);
  if (ensure_xmaps(c, m_tile)) return fail(SJD_E_TMAP, "activation tensor maps");
  // qkv epilogue: rope position 0, k/v land in the last cache slot (never a live key: kv_len + W <= max_len - 1
  // is not guaranteed, so callers must not interleave this with a decode in flight)
  sjd_forward_args a;
  memset(&a, 0, sizeof(a));
  a.rope_pos = c->pos_zero;
  a.cache_pos = c->pos_last;
  GemmEpi e;
  memset(&e, 0, sizeof(e));
  e.mode = EPI_F32;
  e.M = M;
  e.out = c->part_o;   // scratch: rows*W*vocab floats must fit, else only one row is written
  e.ld_out = g.vocab;
  if (size_t(M) * g.vocab > size_t(c->max_chunks) * SJD_MAX_TOKENS * g.n_heads * g.head_dim) e.M = 1;
  const int rc = forward_chain(c, W, &a, true, &e, s);
  return rc ? fail(SJD_E_LAUNCH, "sjd_ctx_gemm_only") : 0;
}

}  // extern "C"
