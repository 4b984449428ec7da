// Weight-streaming "swap-AB" GEMM with fused, deterministic split-K fix-up and fused block epilogues.
//
//   Y[M,N] = X[M,K] * W[N,K]^T,   M = CFG rows * Jacobi window (1..256 token rows),  N,K in the thousands.
//
// The window is far below the B200 ridge (M FLOP/byte vs ~210), so this is an HBM streamer that happens to
// use tensor cores:
//   * the WEIGHT tile is the UMMA "A" operand (128 rows fill the M=128 slot of tcgen05.mma), the TOKEN rows are
//     the UMMA "N" dimension (16..256); accumulators [128 x m_tile] fp32 live in TMEM, double buffered;
//   * persistent grid of one CTA per SM; stream-K split of the flattened (tile, k-block) space so every SM
//     streams the same number of weight bytes whatever N is (streamk.cuh);
//   * warp 0 = TMA producer (SWIZZLE_128B, deep mbarrier ring), warp 1 = single-thread tcgen05.mma issuer,
//     warp 2 = activation TMA, warps 4..19 = epilogue;
//   * a launch runs a CHAIN of up to 4 GEMMs ([o_proj, gate_up, down, next qkv] | lm_head): the weight producer never
//     waits for data — it keeps the ring full across op boundaries and prefetches the next tiles into L2 while the
//     ring is blocked — and a separate activation producer waits for the previous op's "rows finalised" counter
//     (op 0: griddepcontrol.wait, the kernel is launched with programmatic dependent launch);
//   * split tiles are fixed up in-kernel, deterministically (partials are always added in CTA order):
//       ordinary epilogues — the CTA that owns a tile's head segment (which it computes last) is the finisher; the
//         other contributors have the tile as their FIRST segment and park fp32 partials early, so the finisher
//         never waits for a neighbour that is still streaming;
//       residual + RMSNorm — the statistic spans all tiles of a row, so every segment is parked, one grid-wide
//         counter tells when, and one CTA per token row sums the row's tiles, adds the residual, computes the
//         statistic locally and writes h and xn (one exchange, one pass);
//   * 16 epilogue warps: a scheduler issues one warp at a time, so the once-per-tile epilogue needs many warps.
// Epilogues (what the reference does in separate eager kernels after each nn.Linear):
//   EPI_F32        fp32 out (lm_head logits, modeling_chameleon.py:1560-1561), optional bf16 rounding
//   EPI_BF16       bf16 out (generic projection)
//   EPI_QKV        per-head QK-LayerNorm (:216-219) + RoPE (:153-177 rotate-half | llamagen.py:457-467 pairs)
//                  + KV-cache append (replaces DynamicCache.update's torch.cat, :547) + q store
//   EPI_RESID_NORM h += y ; xn = RMSNorm(h) * w  (residual + next norm, :643-659, :68-73), row-owner scheme above
//   EPI_SILU_MUL   act = silu(gate) * up on gate/up rows interleaved inside each tile (:193-195)
// Rounding points follow the bf16 reference: every nn.Linear output, norm output and residual sum is rounded to
// bf16 before the next op; reductions and norms are computed in fp32.
#include "common.cuh"
#include "streamk.cuh"
// (attention.cu is included before this file in the translation unit: AttnCombine, attn_combine_row)

namespace sjd {

constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kGemmThreads = 128 + kEpiThreads;  // warp 0 weight TMA, 1 MMA, 2 activation TMA, 3 idle, 4..19 epilogue
constexpr int kBlockN = 128;  // weight rows per tile (UMMA M)
constexpr int kBlockK = 64;   // bf16 K elements per stage = one 128-byte swizzle row
constexpr int kMaxStages = 12;
constexpr uint32_t kATileBytes = kBlockN * kBlockK * 2;  // 16 KB per weight tile; a unit carries tpu of them
constexpr int kMaxTpu = 4;
constexpr int kCtrStride = 64;                            // u32 between hot counters: one 256-byte region each
constexpr int kTileCtrStride = 8;                         // u32 per tile {arrived, done, pad..}: 32 bytes
constexpr int kLookahead = 0;                             // weight units (16 KB) prefetched into L2 beyond the ring while it is
                                                          // blocked.  Round 2 sweep on the B200 (profiles/r02c_chain_experiments.txt):
                                                          // 0 is best at <= 64 token rows (4354 GB/s vs 4278 at 32), 64 and the
                                                          // 'always ahead' mode are clearly worse: L2 prefetch requests queue
                                                          // ahead of the ring's demand loads
constexpr int kEpiChunk = 64;                             // token columns staged per epilogue pass
constexpr uint32_t kEpiStageBytes = kEpiChunk * kBlockN * 4;  // 16 KB

enum EpiMode { EPI_F32 = 0, EPI_BF16 = 1, EPI_QKV = 2, EPI_RESID_NORM = 3, EPI_SILU_MUL = 4 };

struct GemmEpi {
  int mode;
  int M;  // valid token rows (<= m_tile)
  int N;  // valid weight rows
  // EPI_F32 / EPI_BF16 / EPI_SILU_MUL
  void* out;
  int ld_out;
  int round_bf16;
  // EPI_RESID_NORM
  __nv_bfloat16* h;            // [rows][d] residual stream, updated in place
  const __nv_bfloat16* norm_w; // [d]
  __nv_bfloat16* xn;           // [rows][d]
  float* ssq;                  // [n_tiles][m_tile] per-tile sums of squares
  float eps;
  // EPI_QKV
  __nv_bfloat16* q_out;        // [M][H][Dh]
  __nv_bfloat16* k_cache;      // layer base [rows][Hkv][Lmax][Dh]
  __nv_bfloat16* v_cache;
  const int* rope_pos;         // [M]
  const int* cache_pos;        // [M]
  const float* rope_cos;       // [n_pos][Dh/2]
  const float* rope_sin;
  const __nv_bfloat16* q_norm_w;  // [H][Dh] or null
  const __nv_bfloat16* q_norm_b;
  const __nv_bfloat16* k_norm_w;  // [Hkv][Dh] or null
  const __nv_bfloat16* k_norm_b;
  int W, H, Hkv, Lmax, Dh, rope_interleaved;
  // split-K fix-up + grid-wide meeting point
  float* ws;                   // [2*grid][m_tile][128] fp32 partial slots (first / last segment of each CTA)
  uint32_t* tile_arrive;       // [n_tiles][kTileCtrStride] {arrived, done, pad}, zero between launches
  uint32_t* ctr;               // [0], [kCtrStride]: row-statistic meeting point {in, done}, zero between launches
  long long* dbg;              // optional [grid][16] clock64 stamps of the epilogue's stages (developer timing)
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- row epilogues ---------------------------------------------------------------------------------------
// A finished tile is consumed row by row (token row m; this lane holds tile columns 4*lane .. 4*lane+3).  Every
// global load an epilogue needs is issued up front for all rows a warp has in flight (EpiAux), so the L2 round
// trips overlap instead of forming one dependent chain per row.
struct EpiAux {
  uint2 h;       // EPI_RESID_NORM: residual row fragment (4 bf16)
  float4 c, s;   // EPI_QKV: rope cos / sin for this lane's columns
  int cpos;      // EPI_QKV: KV slot of the row
};
struct EpiTileConst {   // per (tile, lane) constants hoisted out of the row loop
  int slot, pos0, head;
  bool is_q, is_k, rot;
  float nw[4], nb[4];
  bool has_norm;
};

__device__ __forceinline__ void unpack4(const uint2 raw, float (&f)[4]) {
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
  f[0] = __low2float(a); f[1] = __high2float(a); f[2] = __low2float(b); f[3] = __high2float(b);
}
__device__ __forceinline__ uint2 pack4(const float (&f)[4]) {
  uint2 o;
  *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(f[0], f[1]);
  *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(f[2], f[3]);
  return o;
}

__device__ __forceinline__ EpiTileConst epi_tile_const(const GemmEpi& ep, int tile, int lane) {
  EpiTileConst tc;
  tc.slot = tc.pos0 = tc.head = 0;
  tc.is_q = tc.is_k = tc.rot = tc.has_norm = false;
  if (ep.mode != EPI_QKV) return tc;
  const int col = tile * kBlockN + 4 * lane;
  tc.slot = col / ep.Dh;   // head slot: [0,H) q, [H,H+Hkv) k, then v
  tc.pos0 = col % ep.Dh;   // place of this lane's first column inside its head
  tc.is_q = tc.slot < ep.H;
  tc.is_k = !tc.is_q && tc.slot < ep.H + ep.Hkv;
  tc.rot = tc.is_q || tc.is_k;
  tc.head = tc.is_q ? tc.slot : tc.slot - ep.H;
  const __nv_bfloat16* nw = tc.is_q ? ep.q_norm_w : ep.k_norm_w;
  const __nv_bfloat16* nb = tc.is_q ? ep.q_norm_b : ep.k_norm_b;
  tc.has_norm = tc.rot && nw != nullptr;
  if (tc.has_norm) {
    unpack4(*reinterpret_cast<const uint2*>(nw + tc.head * ep.Dh + tc.pos0), tc.nw);
    unpack4(*reinterpret_cast<const uint2*>(nb + tc.head * ep.Dh + tc.pos0), tc.nb);
  }
  return tc;
}

// s_pos: shared copy of {rope_pos[m], cache_pos[m]} (EPI_QKV)
__device__ __forceinline__ EpiAux epi_load_aux(const GemmEpi& ep, const EpiTileConst& tc, int tile, int m, int lane,
                                               const int2* s_pos) {
  EpiAux a;
  a.h = make_uint2(0u, 0u);
  a.c = a.s = make_float4(0.f, 0.f, 0.f, 0.f);
  a.cpos = 0;
  if (ep.mode == EPI_RESID_NORM) {
    a.h = *reinterpret_cast<const uint2*>(ep.h + size_t(m) * ep.N + tile * kBlockN + 4 * lane);
  } else if (ep.mode == EPI_QKV) {
    const int2 pp = s_pos[m];
    a.cpos = pp.y;
    if (tc.rot) {
      const float* cs = ep.rope_cos + size_t(pp.x) * (ep.Dh / 2);
      const float* sn = ep.rope_sin + size_t(pp.x) * (ep.Dh / 2);
      if (ep.rope_interleaved) {  // pairs (2i, 2i+1): this lane needs table entries pos0/2, pos0/2 + 1
        const float2 c2 = *reinterpret_cast<const float2*>(cs + (tc.pos0 >> 1));
        const float2 s2 = *reinterpret_cast<const float2*>(sn + (tc.pos0 >> 1));
        a.c = make_float4(c2.x, c2.y, 0.f, 0.f);
        a.s = make_float4(s2.x, s2.y, 0.f, 0.f);
      } else {                    // pairs (i, i + 64): entries (pos0 & 63) .. +3
        a.c = *reinterpret_cast<const float4*>(cs + (tc.pos0 & 63));
        a.s = *reinterpret_cast<const float4*>(sn + (tc.pos0 & 63));
      }
    }
  }
  return a;
}

__device__ __forceinline__ void epi_apply(const GemmEpi& ep, const EpiTileConst& tc, const EpiAux& aux, int tile,
                                          int m, int lane, float (&v)[4], int m_tile) {
  const int nb = tile * kBlockN, n0 = 4 * lane;
  if (ep.mode == EPI_F32) {
    float* o = static_cast<float*>(ep.out) + size_t(m) * ep.ld_out + nb + n0;
    if (ep.round_bf16) {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = bf16_round(v[e]);
    }
    if (nb + n0 + 3 < ep.N && (ep.ld_out & 3) == 0) {
      *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (nb + n0 + e < ep.N) o[e] = v[e];
    }
  } else if (ep.mode == EPI_BF16) {
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(ep.out) + size_t(m) * ep.ld_out + nb + n0;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (nb + n0 + e < ep.N) o[e] = __float2bfloat16_rn(v[e]);
  } else if (ep.mode == EPI_SILU_MUL) {
    // tile columns [4l, 4l+1] = gate rows (64*tile + 2l, +1), [4l+2, 4l+3] = the matching up rows (pack_gate_up)
    float a[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float g = bf16_round(v[e]), u = bf16_round(v[e + 2]);
      const float sg = bf16_round(g / (1.f + expf(-g)));
      a[e] = sg * u;
    }
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(ep.out) + size_t(m) * ep.ld_out + tile * 64 + 2 * lane;
    *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(a[0], a[1]);
  } else if (ep.mode == EPI_RESID_NORM) {
    float hv[4], x[4], ss = 0.f;
    unpack4(aux.h, hv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      x[e] = bf16_round(hv[e] + bf16_round(v[e]));
      ss += x[e] * x[e];
    }
    *reinterpret_cast<uint2*>(ep.h + size_t(m) * ep.N + nb + n0) = pack4(x);
    ss = warp_sum(ss);
    if (lane == 0) ep.ssq[size_t(tile) * m_tile + m] = ss;
  } else {  // EPI_QKV
    const int Dh = ep.Dh;
    float x[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) x[e] = bf16_round(v[e]);
    if (tc.has_norm) {  // per-head LayerNorm over Dh = 128 (one tile row = one head, warp-uniform), eps 1e-5
      const float mean = warp_sum(x[0] + x[1] + x[2] + x[3]) * (1.f / 128.f);
      float var = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float dl = x[e] - mean; var += dl * dl; }
      const float rstd = rsqrtf(warp_sum(var) * (1.f / 128.f) + 1e-5f);
#pragma unroll
      for (int e = 0; e < 4; ++e) x[e] = (x[e] - mean) * rstd * tc.nw[e] + tc.nb[e];
    }
    if (ep.rope_interleaved) {  // adjacent pairs (2i, 2i+1) inside the head
      if (tc.rot) {
        const float c0 = aux.c.x, s0 = aux.s.x, c1 = aux.c.y, s1 = aux.s.y;
        const float a0 = x[0], b0 = x[1], a1 = x[2], b1 = x[3];
        x[0] = a0 * c0 - b0 * s0; x[1] = b0 * c0 + a0 * s0;
        x[2] = a1 * c1 - b1 * s1; x[3] = b1 * c1 + a1 * s1;
      }
    } else {  // rotate-half pairs (i, i + 64), Dh = 128: the partner column lives in lane ^ 16 (all lanes shuffle)
      float xp[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) xp[e] = __shfl_xor_sync(0xffffffffu, x[e], 16);
      if (tc.rot) {
        const float cc[4] = {aux.c.x, aux.c.y, aux.c.z, aux.c.w}, sn[4] = {aux.s.x, aux.s.y, aux.s.z, aux.s.w};
        const bool hi = lane >= 16;
#pragma unroll
        for (int e = 0; e < 4; ++e) x[e] = hi ? (x[e] * cc[e] + xp[e] * sn[e]) : (x[e] * cc[e] - xp[e] * sn[e]);
      }
    }
    __nv_bfloat16* dst;
    if (tc.is_q) dst = ep.q_out + (size_t(m) * ep.H + tc.slot) * Dh + tc.pos0;
    else {
      const int hk = tc.is_k ? tc.slot - ep.H : tc.slot - ep.H - ep.Hkv;
      const int b = m / ep.W;
      __nv_bfloat16* base = tc.is_k ? ep.k_cache : ep.v_cache;
      dst = base + ((size_t(b) * ep.Hkv + hk) * ep.Lmax + aux.cpos) * Dh + tc.pos0;
    }
    *reinterpret_cast<uint2*>(dst) = pack4(x);
  }
}

// Sum of the parked partials of one (token row, tile) over its contributing CTAs c_first..c_last, in CTA order
// (bit-reproducible), C loads in flight: a tile of o_proj / down is split over 5-6 CTAs, so C = 8 fetches them in one
// L2 round trip where C = 4 needs two.
template <int C>
__device__ __forceinline__ void row_partial_sum(float (&v)[4], const float* __restrict__ pbase, int c_first, int c_last,
                                                bool first_uses_last_slot, int tpu, int hh, size_t slot_floats) {
#pragma unroll 1
  for (int cb = c_first; cb <= c_last; cb += C) {
    float4 pv[C];
#pragma unroll
    for (int cc = 0; cc < C; ++cc) {
      const int c = min(cb + cc, c_last);
      const int slot = (2 * c + ((c == c_first && first_uses_last_slot) ? 1 : 0)) * tpu + hh;
      pv[cc] = __ldcg(reinterpret_cast<const float4*>(pbase + size_t(slot) * slot_floats));
    }
#pragma unroll
    for (int cc = 0; cc < C; ++cc)   // CTA order: bit-reproducible
      if (cb + cc <= c_last) { v[0] += pv[cc].x; v[1] += pv[cc].y; v[2] += pv[cc].z; v[3] += pv[cc].w; }
  }
}

// ---- a chain of GEMMs executed by one persistent kernel ---------------------------------------------------
constexpr int kMaxChainOps = 4;
struct TmapSet {
  CUtensorMap w[5];   // weights: qkv, o, gate_up, down, lm_head (context) / [0] stand-alone
  CUtensorMap x[4];   // activations: xn, attn, act, xl (context) / [0] stand-alone
};
struct GemmOp {
  int wmap, xmap;
  int w_row0;     // plain [N,K] weights: first weight row of this op (layer * N)
  int w_tiled;    // 1: weights re-laid out as contiguous [unit][128][64] tiles; w_row0 is then the first UNIT

  StreamK sk;
  GemmEpi ep;
};
// K/V cache spans the attention kernel that FOLLOWS this chain will stream: sent to L2 by the weight producer once it
// has issued its last weight load, i.e. while the chain's last cross-CTA tail and the kernel boundary leave HBM idle.
struct KvPrefetch {
  const __nv_bfloat16* k;   // layer base [rows][Hkv][Lmax][Dh]; null = off
  const __nv_bfloat16* v;
  int rows, Hkv, Lmax, Dh;
  int kv_len;               // cached keys (the window's own keys are written by this chain's QKV epilogue)
  int lo[8];                // first visible key per row
};

struct Chain {
  int n_ops, num_stages;
  int lookahead;   // weight units prefetched into L2 beyond the ring
  int row_c8;      // 1: the row owners fetch 8 partials per L2 round trip (SJD_GEMM_ROWC8)
  int epi_rows2;   // 1: the finisher epilogue keeps two token rows per warp in flight (SJD_GEMM_EPI2)
  int pf_always;   // 1: keep the L2 prefetch frontier `lookahead` units ahead in steady state too (0: only while the ring is blocked)
  uint32_t tmem_cols;
  int nbuf;        // accumulator sets in TMEM
  uint32_t* fin;   // [(kMaxChainOps + 2) * kCtrStride] rows finalised per op, exit counter, pre-op counter; zero between launches
  // optional pre-op run by the (otherwise idle) epilogue warps before op 0: merge the attention's key-split partials
  // into the bf16 rows that op 0 (o_proj) reads; pre.n_chunks == 0 switches it off
  AttnCombine pre;
  KvPrefetch kvpf;
  GemmOp ops[kMaxChainOps];
};

// One persistent CTA per SM runs every GEMM of the chain back to back.  The WEIGHT producer never waits for
// data — it keeps the smem ring full across op boundaries — while the ACTIVATION producer of op i+1 waits for
// op i's "rows finalised" counter (op 0: for the previous kernel, griddepcontrol.wait).  So the HBM weight stream
// of the next projection overlaps the cross-CTA fix-up / normalisation tail of the current one.
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_chain_kernel(const __grid_constant__ TmapSet maps, const Chain ch) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 4];
  __shared__ uint32_t tmem_holder;
  __shared__ int2 s_pos[256];   // EPI_QKV: {rope_pos, cache_pos} per token row

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_stages = ch.num_stages;
  const int m_tile = ch.ops[0].sk.m_tile;   // common to the chain
  const int tpu = ch.ops[0].sk.tpu;         // weight tiles per unit, common to the chain
  const int nbuf = ch.nbuf;                 // accumulator sets in TMEM (2, or 1 when 2 * tpu * m_tile > 512 columns)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_tile_bytes = uint32_t(m_tile) * kBlockK * 2;
  const uint32_t a_bytes = uint32_t(tpu) * kATileBytes;
  const uint32_t stage_bytes = a_bytes + b_tile_bytes;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * kMaxStages + 2 + a); };
  // epilogue staging tile sits after the ring
  float* stage_tile = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) +
                                               size_t(num_stages) * stage_bytes);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < ch.n_ops; ++i) {
      tma_prefetch_desc(&maps.w[ch.ops[i].wmap]);
      tma_prefetch_desc(&maps.x[ch.ops[i].xmap]);
    }
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_holder), ch.tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_holder;
  pdl_launch_dependents();  // the next kernel may become resident as CTAs of this one retire

  const int cta = blockIdx.x;

  if (warp == 0) {
    // ===== WEIGHT producer: never waits for data, only for ring slots =====
    // While the ring is blocked (the consumer is waiting for activations at an op boundary) it keeps HBM busy
    // by prefetching the next weight tiles into L2, up to kLookahead units ahead.
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int pf_i = 0;                 // prefetch cursor (op, unit), never behind the issue cursor
      uint32_t pf_u = ch.ops[0].sk.begin(cta);
      int pf_ahead = 0;             // units the prefetch cursor is ahead of the issue cursor
      // a unit = tpu tiles of 128 weight rows: one TMA box of 128 * min(tpu, 2) rows, two boxes when tpu = 4
      const int box_rows = kBlockN * (tpu < 2 ? tpu : 2), n_box = (tpu + 1) / 2;
      auto w_coords = [&](const GemmOp& op, uint32_t u, int& c0, int& c1) {
        if (op.w_tiled) { c0 = 0; c1 = (op.w_row0 + int(u)) * kBlockN * tpu; }
        else { const uint32_t KB = uint32_t(op.sk.kb), grp = u / KB; c0 = int((u - grp * KB) * kBlockK); c1 = op.w_row0 + int(grp) * kBlockN * tpu; }
      };
      auto pf_normalise = [&]() {   // skip exhausted ops
        while (pf_i < ch.n_ops && pf_u >= ch.ops[pf_i].sk.begin(cta + 1)) {
          ++pf_i;
          if (pf_i < ch.n_ops) pf_u = ch.ops[pf_i].sk.begin(cta);
        }
      };
      for (int i = 0; i < ch.n_ops; ++i) {
        const GemmOp& op = ch.ops[i];
        const CUtensorMap* tw = &maps.w[op.wmap];
        const uint32_t u0 = op.sk.begin(cta), u1 = op.sk.begin(cta + 1);
        for (uint32_t u = u0; u < u1; ++u) {
          while (!mbar_try_wait(empty_bar(stage), phase ^ 1)) {
            if (pf_ahead == 0) { pf_i = i; pf_u = u; }
            pf_normalise();
            if (pf_ahead < ch.lookahead && pf_i < ch.n_ops) {
              int c0, c1;
              w_coords(ch.ops[pf_i], pf_u, c0, c1);
              for (int j = 0; j < n_box; ++j) tma_prefetch_l2_2d(&maps.w[ch.ops[pf_i].wmap], c0, c1 + j * box_rows);
              ++pf_u;
              ++pf_ahead;
            } else {
              __nanosleep(64);
            }
          }
          int c0, c1;
          w_coords(op, u, c0, c1);
          mbar_arrive_expect_tx(full_bar(stage), stage_bytes);
          for (int j = 0; j < n_box; ++j)
            tma_load_2d(smem_base + uint32_t(stage) * stage_bytes + uint32_t(j) * 2u * kATileBytes, tw, c0, c1 + j * box_rows,
                        full_bar(stage), kPolicyEvictFirst);
          if (pf_ahead > 0) --pf_ahead;
          if (ch.pf_always) {   // top the L2 frontier up: at most two prefetches per load issued
            if (pf_ahead == 0) { pf_i = i; pf_u = u + 1; }
            for (int k = 0; k < 2; ++k) {
              pf_normalise();
              if (pf_ahead >= ch.lookahead || pf_i >= ch.n_ops) break;
              int p0, p1;
              w_coords(ch.ops[pf_i], pf_u, p0, p1);
              for (int j = 0; j < n_box; ++j) tma_prefetch_l2_2d(&maps.w[ch.ops[pf_i].wmap], p0, p1 + j * box_rows);
              ++pf_u;
              ++pf_ahead;
            }
          }
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
      }
      if (ch.kvpf.k != nullptr) {
        // every weight load of the chain has been issued: HBM idles through the last op's tail and the kernel
        // boundary — send this CTA's share of the next attention's K/V span to L2 (32 KB pieces, round-robin)
        const KvPrefetch& kp = ch.kvpf;
        constexpr uint32_t kPiece = 32768;
        const int n_span = kp.rows * kp.Hkv * 2;
        for (int sp = 0; sp < n_span; ++sp) {
          const int bh = sp >> 1, b = bh / kp.Hkv;
          const int lo = (kp.lo[b] / 64) * 64;
          if (kp.kv_len <= lo) continue;
          const size_t bytes = size_t(kp.kv_len - lo) * kp.Dh * 2;
          const char* base = reinterpret_cast<const char*>((sp & 1) ? kp.v : kp.k) + (size_t(bh) * kp.Lmax + lo) * kp.Dh * 2;
          const uint32_t n_piece = uint32_t((bytes + kPiece - 1) / kPiece);
          for (uint32_t pc = uint32_t((cta + sp) % int(gridDim.x)); pc < n_piece; pc += gridDim.x) {
            const size_t off = size_t(pc) * kPiece;
            const uint32_t nb = uint32_t(bytes - off < kPiece ? bytes - off : kPiece);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + off), "r"(nb) : "memory");
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===== ACTIVATION producer: op i's input exists once op i-1 is finalised everywhere =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < ch.n_ops; ++i) {
        const GemmOp& op = ch.ops[i];
        const CUtensorMap* tx = &maps.x[op.xmap];
        const uint32_t u0 = op.sk.begin(cta), u1 = op.sk.begin(cta + 1), KB = uint32_t(op.sk.kb);
        if (u0 < u1) {
          if (i == 0) {
            pdl_wait();
            if (ch.pre.n_chunks > 0) {   // op 0 reads what the pre-op writes
              spin_until_ge(&ch.fin[(kMaxChainOps + 1) * kCtrStride], uint32_t(ch.pre.rows * ch.pre.H * ch.pre.W));
              fence_proxy_async_all();
            }
          } else {
            const GemmOp& pr = ch.ops[i - 1];
            const uint32_t target = uint32_t(pr.sk.n_tiles) * uint32_t(pr.ep.M);
            spin_until_ge(&ch.fin[(i - 1) * kCtrStride], target);
            fence_proxy_async_all();  // other CTAs' generic-proxy stores -> this thread's TMA reads
          }
        }
        for (uint32_t u = u0; u < u1; ++u) {
          const uint32_t kb = u % KB;
          mbar_wait(empty_bar(stage), phase ^ 1);
          tma_load_2d(smem_base + uint32_t(stage) * stage_bytes + a_bytes, tx, int(kb * kBlockK), 0,
                      full_bar(stage), kPolicyEvictLast);
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16_f32(kBlockN, uint32_t(m_tile));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int i = 0; i < ch.n_ops; ++i) {
        const GemmOp& op = ch.ops[i];
        const uint32_t u0 = op.sk.begin(cta), u1 = op.sk.begin(cta + 1), KB = uint32_t(op.sk.kb);
        uint32_t u = u0;
        while (u < u1) {
          const uint32_t grp = u / KB;
          const uint32_t seg_end = min(u1, (grp + 1) * KB);
          mbar_wait(tempty_bar(acc), acc_phase ^ 1);
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + uint32_t(acc * tpu) * uint32_t(m_tile);
          uint32_t accumulate = 0;
          for (; u < seg_end; ++u) {
            mbar_wait(full_bar(stage), phase);
            tcgen05_fence_after();
            const uint32_t sa = smem_base + uint32_t(stage) * stage_bytes;
            const uint64_t db = umma_desc_sw128_kmajor(sa + a_bytes);
#pragma unroll 1
            for (int h = 0; h < tpu; ++h) {   // the tiles of the unit share the activation tile (B operand)
              const uint64_t da = umma_desc_sw128_kmajor(sa + uint32_t(h) * kATileBytes);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
                umma_bf16_ss(d_tmem + uint32_t(h * m_tile), da + uint64_t(2 * k), db + uint64_t(2 * k), idesc,
                             (k == 0) ? accumulate : 1u);
            }
            accumulate = 1;
            umma_commit(empty_bar(stage));
            if (++stage == num_stages) { stage = 0; phase ^= 1; }
          }
          umma_commit(tfull_bar(acc));
          if (++acc == nbuf) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue warps (16): TMEM lane quarter = warp % 4, column group = (warp - 4) / 4 =====
    //  whole tile inside this CTA's range : TMEM -> smem staging -> fused epilogue, right away
    //  split tile, ordinary epilogues       : the CTA that owns the tile's head segment (which it computes LAST) is the
    //                                         tile's finisher; every other contributor has the tile as the FIRST segment
    //                                         of its range and parks its fp32 partial early.  The finisher adds the
    //                                         partials in CTA order (bit-reproducible) and runs the epilogue: nobody
    //                                         waits for a neighbour that is still streaming.
    //  EPI_RESID_NORM                       : the row statistic needs every tile of a row, so ALL segments are parked,
    //                                         one grid-wide counter tells when, and one CTA per token row then sums the
    //                                         row's tiles, adds the residual, computes the RMS statistic locally and
    //                                         writes h and xn — one exchange and one pass instead of two of each.
    // One scheduler runs one instruction stream at a time, so the epilogue's speed comes from having many warps
    // (one token row / one tile each), not from unrolling.
    pdl_wait();  // everything below touches buffers the previous kernel may still be writing
    const int ew = warp - 4;        // 0..15
    const int quarter = ew & 3;     // TMEM lane quarter this warp may access (== warp % 4)
    const int cgrp = ew >> 2;       // which 16-column blocks of the accumulator this warp drains
    const int nrow = quarter * 32 + lane;  // weight row of this thread inside the tile
    const int tid_e = threadIdx.x - 128;   // 0..511
    if (ch.pre.n_chunks > 0) {
      // ---- pre-op: split-KV combine of the attention that ran before this kernel, one (row, head, query) per warp ----
      const int total = ch.pre.rows * ch.pre.H * ch.pre.W;
      uint32_t mine = 0;
#pragma unroll 1
      for (int r = cta * kEpiWarps + ew; r < total; r += int(gridDim.x) * kEpiWarps) {
        if (ch.pre.head_dim == 128) attn_combine_row<128>(ch.pre, r, lane);
        else attn_combine_row<64>(ch.pre, r, lane);
        ++mine;
      }
      // rows per CTA: every warp counted its own; publish the CTA's total
      mine = __shfl_sync(0xffffffffu, mine, 0);
      __shared__ uint32_t s_pre_rows;
      if (tid_e == 0) s_pre_rows = 0;
      epi_bar();
      if (lane == 0 && mine) atomicAdd(&s_pre_rows, mine);
      epi_bar();
      if (tid_e == 0 && s_pre_rows) {
        __threadfence();
        atomicAdd(&ch.fin[(kMaxChainOps + 1) * kCtrStride], s_pre_rows);
      }
    }
    int acc = 0;
    uint32_t acc_phase = 0;
#pragma unroll 1
    for (int op_i = 0; op_i < ch.n_ops; ++op_i) {
      const GemmOp& op = ch.ops[op_i];
      const GemmEpi& ep = op.ep;
      const StreamK& sk = op.sk;
      const uint32_t u0 = sk.begin(cta), u1 = sk.begin(cta + 1), KB = uint32_t(sk.kb);
      const bool row_mode = (ep.mode == EPI_RESID_NORM);
      if (ep.mode == EPI_QKV) {
        // (for op_i > 0 the positions are kernel inputs, not produced by the chain: safe to read right away)
        for (int m = tid_e; m < ep.M; m += kEpiThreads) s_pos[m] = make_int2(ep.rope_pos[m], ep.cache_pos[m]);
        epi_bar();
      }
      const size_t slot_floats = sk.slot_floats();
      long long* dbg = (ep.dbg && tid_e == 0) ? ep.dbg + size_t(cta) * 16 : nullptr;
      if (dbg) dbg[0] = clock64();
      uint32_t u = u0;
      uint32_t my_rows = 0;   // (tile, token row) pairs this CTA made final
#pragma unroll 1
      while (u < u1) {
        const uint32_t grp = u / KB;
        const uint32_t seg_end = min(u1, (grp + 1) * KB);
        const bool head = (grp * KB >= u0);                     // this CTA owns the group's first k-block
        const bool whole = head && ((grp + 1) * KB <= u1);
        mbar_wait(tfull_bar(acc), acc_phase);
        tcgen05_fence_after();
        if (dbg) dbg[(seg_end == u1) ? 2 : 1] = clock64();   // accumulator of a (1) non-last / (2) last segment ready
        const uint32_t t_acc = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * tpu) * uint32_t(sk.m_tile);
        const int seg = (u == u0) ? 0 : 1;
        if (row_mode || !head) {
          // park the partials.  slots (2*cta + seg) * tpu + h: seg 0 = the CTA's first segment, 1 = its last (row mode
          // only, <= 2 segments); h = tile of the group
#pragma unroll 1
          for (int h = 0; h < tpu; ++h) {
            if (int(grp) * tpu + h >= sk.n_tiles) break;        // a ragged last group: nothing behind the missing tiles
            float* dst = ep.ws + size_t((2 * cta + seg) * tpu + h) * slot_floats + nrow;
            const uint32_t t_addr = t_acc + uint32_t(h * sk.m_tile);
#pragma unroll 1
            for (int m0 = 16 * cgrp; m0 < sk.m_tile; m0 += 64) {
              uint32_t v[16];
              tmem_ld_32x32b_x16(t_addr + uint32_t(m0), v);
              tmem_ld_wait();
              if (m0 < ep.M) {
#pragma unroll
                for (int j = 0; j < 16; ++j) __stcg(dst + size_t(m0 + j) * 128, __uint_as_float(v[j]));
              }
            }
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
          if (!row_mode) {
            epi_bar();   // every warp's partial stores happen-before thread 0's (cumulative) fence + flag
            if (tid_e == 0) {
              __threadfence();
              atomicAdd(&ep.tile_arrive[kTileCtrStride * grp], 1u);
            }
          }
        } else {
          // head segment (whole group, or the finisher's part of a split group): stage, add the parked partials, finish
          const int c_last = whole ? cta : sk.last_cta(int(grp));
          if (c_last > cta && tid_e == 0) {
            spin_until_ge(&ep.tile_arrive[kTileCtrStride * grp], uint32_t(c_last - cta));
            ep.tile_arrive[kTileCtrStride * grp] = 0;   // leave it zero for the next launch
          }
#pragma unroll 1
          for (int h = 0; h < tpu; ++h) {
            const int tile = int(grp) * tpu + h;
            const bool last_h = (h == tpu - 1) || (tile + 1 >= sk.n_tiles);
            if (tile >= sk.n_tiles) break;
            const uint32_t t_addr = t_acc + uint32_t(h * sk.m_tile);
            const EpiTileConst tc = epi_tile_const(ep, tile, lane);
#pragma unroll 1
            for (int c0m = 0; c0m < sk.m_tile; c0m += kEpiChunk) {
              const int cw = min(kEpiChunk, sk.m_tile - c0m);
#pragma unroll 1
              for (int j0 = 16 * cgrp; j0 < cw; j0 += 64) {
                uint32_t v[16];
                tmem_ld_32x32b_x16(t_addr + uint32_t(c0m + j0), v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) stage_tile[(j0 + j) * kBlockN + nrow] = __uint_as_float(v[j]);
              }
              if (last_h && c0m + kEpiChunk >= sk.m_tile) {  // every accumulator of the group fully drained
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(acc));
              }
              epi_bar();  // staging tile complete; also orders thread 0's acquire before everybody's partial reads
              if (ch.epi_rows2) {
                // two token rows per warp in flight: their aux / partial loads (L2 round trips) overlap
#pragma unroll 1
                for (int ml = ew; ml < cw; ml += 2 * kEpiWarps) {
                  const int mA = c0m + ml;
                  if (mA >= ep.M) break;
                  const bool hasB = (ml + kEpiWarps < cw) && (mA + kEpiWarps < ep.M);
                  const int mlB = hasB ? ml + kEpiWarps : ml, mB = c0m + mlB;
                  const EpiAux auxA = epi_load_aux(ep, tc, tile, mA, lane, s_pos);
                  const EpiAux auxB = epi_load_aux(ep, tc, tile, mB, lane, s_pos);
                  const float4 tA = *reinterpret_cast<const float4*>(stage_tile + ml * kBlockN + 4 * lane);
                  const float4 tB = *reinterpret_cast<const float4*>(stage_tile + mlB * kBlockN + 4 * lane);
                  float vA[4] = {tA.x, tA.y, tA.z, tA.w}, vB[4] = {tB.x, tB.y, tB.z, tB.w};
                  constexpr int C = 2;   // partials in flight per row
#pragma unroll 1
                  for (int cb = cta + 1; cb <= c_last; cb += C) {
                    float4 pa[C], pb[C];
#pragma unroll
                    for (int cc = 0; cc < C; ++cc) {   // unconditional (clamped) loads so that they batch
                      const int c = min(cb + cc, c_last);
                      const float* base = ep.ws + size_t(2 * c * tpu + h) * slot_floats + 4 * lane;
                      pa[cc] = __ldcg(reinterpret_cast<const float4*>(base + size_t(mA) * 128));
                      pb[cc] = __ldcg(reinterpret_cast<const float4*>(base + size_t(mB) * 128));
                    }
#pragma unroll
                    for (int cc = 0; cc < C; ++cc)   // CTA order: bit-reproducible
                      if (cb + cc <= c_last) {
                        vA[0] += pa[cc].x; vA[1] += pa[cc].y; vA[2] += pa[cc].z; vA[3] += pa[cc].w;
                        vB[0] += pb[cc].x; vB[1] += pb[cc].y; vB[2] += pb[cc].z; vB[3] += pb[cc].w;
                      }
                  }
                  epi_apply(ep, tc, auxA, tile, mA, lane, vA, sk.m_tile);
                  if (hasB) epi_apply(ep, tc, auxB, tile, mB, lane, vB, sk.m_tile);
                }
              } else {
#pragma unroll 1
              for (int ml = ew; ml < cw; ml += kEpiWarps) {   // one token row per warp at a time
                const int m = c0m + ml;
                if (m >= ep.M) break;
                const EpiAux aux = epi_load_aux(ep, tc, tile, m, lane, s_pos);
                const float4 t = *reinterpret_cast<const float4*>(stage_tile + ml * kBlockN + 4 * lane);
                float v[4] = {t.x, t.y, t.z, t.w};
                constexpr int C = 4;   // partials in flight
#pragma unroll 1
                for (int cb = cta + 1; cb <= c_last; cb += C) {
                  float4 pv[C];
#pragma unroll
                  for (int cc = 0; cc < C; ++cc) {   // unconditional (clamped) loads so that they batch
                    const int c = min(cb + cc, c_last);
                    pv[cc] = __ldcg(reinterpret_cast<const float4*>(ep.ws + size_t(2 * c * tpu + h) * slot_floats + size_t(m) * 128 + 4 * lane));
                  }
#pragma unroll
                  for (int cc = 0; cc < C; ++cc)   // CTA order: bit-reproducible
                    if (cb + cc <= c_last) { v[0] += pv[cc].x; v[1] += pv[cc].y; v[2] += pv[cc].z; v[3] += pv[cc].w; }
                }
                epi_apply(ep, tc, aux, tile, m, lane, v, sk.m_tile);
              }
              }
              epi_bar();  // rows done before the next chunk overwrites the staging tile
            }
            my_rows += uint32_t(ep.M);
          }
        }
        u = seg_end;
        if (++acc == nbuf) { acc = 0; acc_phase ^= 1; }
      }
      if (dbg) dbg[3] = clock64();   // all segments drained / parked
      // ---- EPI_RESID_NORM: one CTA per token row finishes the row across all tiles -----------------------------
      if (row_mode) {
        float* red = stage_tile;   // [kEpiWarps] partial sums of squares
        const uint32_t G = uint32_t(sk.grid);
        if (u1 > u0) {             // this CTA parked something: tell the row owners
          epi_bar();
          if (tid_e == 0) {
            __threadfence();
            atomicAdd(&ep.ctr[0], 1u);
          }
        }
        // rows m with (m * G) / M == cta
        const int m_lo = int((uint32_t(cta) * uint32_t(ep.M) + G - 1) / G);
        const int m_hi = uint32_t(cta) < G ? int((uint32_t(cta + 1) * uint32_t(ep.M) + G - 1) / G) : m_lo;
        if (m_hi > m_lo) {
          if (tid_e == 0) spin_until_ge(&ep.ctr[0], G);
          epi_bar();
          if (dbg) dbg[5] = clock64();   // every partial of the op is parked
          const float inv_d = 1.f / float(ep.N);
          constexpr int TPW = 4;         // tiles per warp: n_tiles <= 64 (checked on the host)
#pragma unroll 1
          for (int m = m_lo; m < m_hi; ++m) {
            float x[TPW][4];
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
              const int t = ew + kEpiWarps * i;
              if (t < sk.n_tiles) {
                const int grp = t / tpu, hh = t - grp * tpu;
                const int c_first = sk.first_cta(grp), c_last = sk.last_cta(grp);
                const bool first_uses_last_slot = (sk.begin(c_first) / KB) != uint32_t(grp);
                const size_t off = size_t(m) * ep.N + size_t(t) * kBlockN + 4 * lane;
                const uint2 hraw = *reinterpret_cast<const uint2*>(ep.h + off);
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                const float* pbase = ep.ws + size_t(m) * 128 + 4 * lane;
                if (ch.row_c8) row_partial_sum<8>(v, pbase, c_first, c_last, first_uses_last_slot, tpu, hh, slot_floats);
                else row_partial_sum<4>(v, pbase, c_first, c_last, first_uses_last_slot, tpu, hh, slot_floats);
                float hv[4];
                unpack4(hraw, hv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  x[i][e] = bf16_round(hv[e] + bf16_round(v[e]));
                  ss += x[i][e] * x[i][e];
                }
              }
            }
            ss = warp_sum(ss);
            if (lane == 0) red[ew] = ss;
            epi_bar();
            float tot = 0.f;
#pragma unroll
            for (int w2 = 0; w2 < kEpiWarps; ++w2) tot += red[w2];   // fixed order
            const float rinv = rsqrtf(tot * inv_d + ep.eps);
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
              const int t = ew + kEpiWarps * i;
              if (t < sk.n_tiles) {
                const size_t off = size_t(m) * ep.N + size_t(t) * kBlockN + 4 * lane;
                float wv[4], o[4];
                unpack4(*reinterpret_cast<const uint2*>(ep.norm_w + t * kBlockN + 4 * lane), wv);
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = wv[e] * bf16_round(x[i][e] * rinv);
                *reinterpret_cast<uint2*>(ep.h + off) = pack4(x[i]);
                *reinterpret_cast<uint2*>(ep.xn + off) = pack4(o);
              }
            }
            epi_bar();   // red[] is reused by the next row
          }
          my_rows += uint32_t(m_hi - m_lo) * uint32_t(sk.n_tiles);
          if (dbg) dbg[10] = clock64();   // rows written
          if (tid_e == 0) {
            // re-arm once every row owner has read the partials (M rows in total)
            const uint32_t done = atomicAdd(&ep.ctr[kCtrStride], uint32_t(m_hi - m_lo)) + uint32_t(m_hi - m_lo);
            if (done == uint32_t(ep.M)) {
              ep.ctr[0] = 0;
              ep.ctr[kCtrStride] = 0;
            }
          }
        }
      }
      if (dbg) dbg[4] = clock64();
      // ---- publish: this CTA's rows of op_i are final (the next op's activation producer waits for all) ----
      if (my_rows > 0) {
        epi_bar();
        if (tid_e == 0) {
          __threadfence();
          atomicAdd(&ch.fin[op_i * kCtrStride], my_rows);
        }
      }
      if (dbg) dbg[6] = clock64();   // epilogue warps done with this op
    }  // op loop
    // ---- last CTA out re-arms the chain counters for the next launch ----
    if (tid_e == 0) {
      __threadfence();
      const uint32_t out = atomicAdd(&ch.fin[kMaxChainOps * kCtrStride], 1u) + 1u;
      if (out == gridDim.x) {
        for (int i = 0; i <= kMaxChainOps + 1; ++i) ch.fin[i * kCtrStride] = 0;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, ch.tmem_cols);
  }
}

// ----------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols], 128-byte swizzle.
int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {kBlockK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

// weight tiles per stream-K unit (streamk.cuh).  Fixed per process: the context packs its weights for it.
// SJD_GEMM_TPU = 1 | 2 | 4 (developer A/B switch).  Default 1: measured on the B200 (profiles/r02b_chain_experiments.txt),
// sharing the activation tile between 2 or 4 weight tiles does NOT pay — with ~1 tile per CTA and op the wider stream-K
// groups are split over 2-4x more CTAs, and the extra fp32 partial traffic / fix-up fan-in costs more than the saved
// L2 -> SM activation re-reads (Lumina-7B, 64 token rows: 4256 GB/s at tpu 1, 3792 at tpu 2, 2296 at tpu 4).
int gemm_tpu() {
  static int tpu = 0;
  if (!tpu) {
    tpu = 1;
    if (const char* e = getenv("SJD_GEMM_TPU")) {
      const int v = atoi(e);
      if (v == 1 || v == 2 || v == 4) tpu = v;
    }
  }
  return tpu;
}
static inline uint32_t weight_box_rows() { return uint32_t(kBlockN * (gemm_tpu() < 2 ? gemm_tpu() : 2)); }

// tiled weights: [units * tpu * 128 rows, 64 cols] bf16; one unit = tpu contiguous 16 KB tiles = one or two boxes
int make_tmap_tiled(CUtensorMap* out, const void* ptr, uint64_t units) {
  return make_tmap_bf16_2d(out, ptr, units * uint64_t(kBlockN) * gemm_tpu(), kBlockK, weight_box_rows());
}

static int g_num_sms = 0;
int device_num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_num_sms;
}

StreamK gemm_partition(int N, int K, int m_tile, int grid_limit) {
  StreamK sk;
  sk.n_tiles = (N + kBlockN - 1) / kBlockN;
  sk.kb = K / kBlockK;
  sk.m_tile = m_tile;
  sk.tpu = gemm_tpu();
  sk.n_groups = (sk.n_tiles + sk.tpu - 1) / sk.tpu;
  int g = grid_limit > 0 ? grid_limit : device_num_sms();
  uint32_t U = uint32_t(sk.n_groups) * uint32_t(sk.kb);
  sk.grid = int(U < uint32_t(g) ? U : uint32_t(g));
  if (uint64_t(U) * uint64_t(sk.grid) >= (1ull << 31)) sk.grid = 0;   // 32-bit partition math would overflow: rejected by the caller
  return sk;
}

// workspace layout (counters FIRST, so their place does not move with m_tile / grid and they stay zero):
//   [2 KB of counters, one per 256 B: stats {in, done}, stand-alone chain counters][tile {arrive, done}: 32 B each][ssq: n_tiles * m_tile floats][2*grid*tpu partial slots]
struct GemmWorkspace {
  size_t ctr_off, arrive_off, ssq_off, slots_off, bytes;
};
// `arrive_cap` >= n_tiles reserves the counter region: a context sharing one workspace between GEMMs of different
// shapes passes the largest tile count so that no GEMM's data region ever overlaps another GEMM's counters.
GemmWorkspace gemm_workspace(const StreamK& sk, int arrive_cap = 0) {
  GemmWorkspace w;
  const int cap = arrive_cap > sk.n_tiles ? arrive_cap : sk.n_tiles;
  w.ctr_off = 0;
  w.arrive_off = 2048;
  w.ssq_off = w.arrive_off + size_t(cap) * kTileCtrStride * 4;
  w.slots_off = w.ssq_off + size_t(sk.n_tiles) * sk.m_tile * 4;
  w.bytes = w.slots_off + 2 * size_t(sk.grid) * size_t(sk.tpu) * sk.slot_floats() * 4;
  return w;
}

struct ChainShape {
  int num_stages;
  int nbuf;
  uint32_t tmem_cols;
  uint32_t smem_bytes;
};

static int g_smem_budget = 0;
int gemm_smem_budget() {
  if (!g_smem_budget) {
    g_smem_budget = 222 * 1024;
    if (const char* e = getenv("SJD_GEMM_SMEM_KB")) {
      const int kb = atoi(e);
      if (kb >= 64 && kb <= 222) g_smem_budget = kb * 1024;
    }
  }
  return g_smem_budget;
}

// ring depth / TMEM columns / dynamic smem for a given m_tile
int gemm_shape(ChainShape* g, int m_tile) {
  if (m_tile % 16 != 0 || m_tile < 16 || m_tile > 256) return -3;
  const int tpu = gemm_tpu();
  if (tpu * m_tile > 512) return -3;
  const uint32_t stage_bytes = uint32_t(tpu) * kATileBytes + uint32_t(m_tile) * kBlockK * 2;
  const uint32_t budget = uint32_t(gemm_smem_budget()) - 1024 - kEpiStageBytes;
  int stages = int(budget / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return -4;
  g->num_stages = stages;
  g->nbuf = (2 * tpu * m_tile <= 512) ? 2 : 1;   // double-buffered accumulators while they fit the 512 TMEM columns
  uint32_t cols = 32;
  while (cols < uint32_t(g->nbuf * tpu * m_tile)) cols <<= 1;
  g->tmem_cols = cols;
  g->smem_bytes = uint32_t(stages) * stage_bytes + kEpiStageBytes + 1024;
  return 0;
}

int gemm_attr_once() {
  static int rc = 1;
  if (rc == 1)
    rc = cudaFuncSetAttribute(gemm_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024) ==
                 cudaSuccess
             ? 0
             : -5;
  return rc;
}

long long* g_dbg_buf = nullptr;   // [g_dbg_cap ops][256 CTAs][8]
int g_dbg_cap = 0;
unsigned long long g_dbg_idx = 0;

// ops of one chain must share m_tile; the grid is the largest per-op grid (CTAs beyond an op's grid idle in it)
int chain_launch(const TmapSet& maps, Chain ch, cudaStream_t stream) {
  if (ch.n_ops < 1 || ch.n_ops > kMaxChainOps) return -3;
  ChainShape shp;
  if (gemm_shape(&shp, ch.ops[0].sk.m_tile)) return -4;
  ch.num_stages = shp.num_stages;
  ch.tmem_cols = shp.tmem_cols;
  ch.nbuf = shp.nbuf;
  static int lookahead = -1;
  if (lookahead < 0) {
    const char* e = getenv("SJD_GEMM_LOOKAHEAD");
    lookahead = e ? atoi(e) : kLookahead;
  }
  ch.lookahead = lookahead;
  static const int pf_always = getenv("SJD_GEMM_PF_ALWAYS") ? atoi(getenv("SJD_GEMM_PF_ALWAYS")) : 0;
  ch.pf_always = pf_always;
  // measured (profiles/r02g_epi2.txt): +1.4 % / +2.4 % chain throughput at 64 / 128 token rows, two runs each
  static const int epi2 = getenv("SJD_GEMM_EPI2") ? atoi(getenv("SJD_GEMM_EPI2")) : 1;
  ch.epi_rows2 = epi2;
  static const int rowc8 = getenv("SJD_GEMM_ROWC8") ? atoi(getenv("SJD_GEMM_ROWC8")) : 1;   // +0.3 % (profiles/r02h_rowc8.txt)
  ch.row_c8 = rowc8;
  int grid = 0;
  for (int i = 0; i < ch.n_ops; ++i) {
    if (ch.ops[i].sk.m_tile != ch.ops[0].sk.m_tile || ch.ops[i].sk.grid < 1) return -3;
    if (ch.ops[i].ep.mode == EPI_RESID_NORM) {   // row-owner scheme: <= 2 segments per CTA, <= 64 tiles per row
      const StreamK& k = ch.ops[i].sk;
      if (k.n_tiles > 64 || (k.units() + uint32_t(k.grid) - 1) / uint32_t(k.grid) > uint32_t(k.kb)) return -3;
      if (k.tpu != ch.ops[0].sk.tpu) return -3;
    }
    grid = ch.ops[i].sk.grid > grid ? ch.ops[i].sk.grid : grid;
    ch.ops[i].ep.dbg = g_dbg_buf ? g_dbg_buf + size_t(g_dbg_idx++ % uint64_t(g_dbg_cap)) * 256 * 16 : nullptr;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = shp.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, gemm_chain_kernel, maps, ch) == cudaSuccess ? 0 : -6;
}

}  // namespace sjd
