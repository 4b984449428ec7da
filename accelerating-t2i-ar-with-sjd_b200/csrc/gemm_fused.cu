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
//     warps 2..5 = epilogue;
//   * programmatic dependent launch: the producer fills the whole ring with WEIGHT tiles before
//     griddepcontrol.wait — weights never depend on the previous kernel, only the activation tiles do — so the
//     HBM stream of GEMM n+1 starts while GEMM n (or the attention kernel) is still draining;
//   * split tiles are fixed up in-kernel, reduce-scatter style: every contributor parks its fp32 partial, bumps a
//     per-tile counter, and — after its own last segment — sums ALL partials of the tile in CTA order
//     (bit-reproducible) for its share of the token rows and runs the epilogue on them.  The fix-up is spread
//     over all CTAs; no reduce kernels, no atomics on data.
// Epilogues (what the reference does in separate eager kernels after each nn.Linear):
//   EPI_F32        fp32 out (lm_head logits, modeling_chameleon.py:1560-1561), optional bf16 rounding
//   EPI_BF16       bf16 out (generic projection)
//   EPI_QKV        per-head QK-LayerNorm (:216-219) + RoPE (:153-177 rotate-half | llamagen.py:457-467 pairs)
//                  + KV-cache append (replaces DynamicCache.update's torch.cat, :547) + q store
//   EPI_RESID_NORM h += y ; xn = RMSNorm(h) * w  (residual + next norm, :643-659, :68-73) — the row statistic
//                  spans all tiles, so finishers meet at a grid-wide counter before writing xn
//   EPI_SILU_MUL   act = silu(gate) * up on gate/up rows interleaved 64/64 per tile (:193-195)
// Rounding points follow the bf16 reference: every nn.Linear output, norm output and residual sum is rounded to
// bf16 before the next op; reductions and norms are computed in fp32.
#include "common.cuh"
#include "streamk.cuh"

namespace sjd {

constexpr int kGemmThreads = 192;
constexpr int kBlockN = 128;  // weight rows per tile (UMMA M)
constexpr int kBlockK = 64;   // bf16 K elements per stage = one 128-byte swizzle row
constexpr int kMaxStages = 12;
constexpr uint32_t kATileBytes = kBlockN * kBlockK * 2;  // 16 KB
constexpr int kEpiChunk = 32;                             // token columns staged per epilogue pass
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
  uint32_t* tile_arrive;       // [n_tiles][2] {arrived, done}, zero between launches
  uint32_t* ctr;               // [2], zero between launches
  long long* dbg;              // optional [grid][8] clock64 stamps of the epilogue's stages (developer timing)
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

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

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_fused_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                  const int w_row0, const StreamK sk, const int num_stages, const uint32_t tmem_cols,
                  const GemmEpi ep) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 4];
  __shared__ uint32_t tmem_holder;
  __shared__ int2 s_pos[256];   // EPI_QKV: {rope_pos, cache_pos} per token row

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_tile_bytes = uint32_t(sk.m_tile) * kBlockK * 2;
  const uint32_t stage_bytes = kATileBytes + b_tile_bytes;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * kMaxStages + 2 + a); };
  // epilogue staging tile sits after the ring
  float* stage_tile = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) +
                                               size_t(num_stages) * stage_bytes);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_holder), tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_holder;
  pdl_launch_dependents();  // the next kernel may become resident (and prefetch its weights) as CTAs of this one retire

  const int cta = blockIdx.x;
  const uint32_t u0 = sk.begin(cta), u1 = sk.begin(cta + 1);
  const uint32_t KB = uint32_t(sk.kb);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      // (1) weights for the first ring-full of units: independent of the previous kernel
      const uint32_t n_pre = min(uint32_t(num_stages), u1 - u0);
      for (uint32_t i = 0; i < n_pre; ++i) {
        const uint32_t u = u0 + i, tile = u / KB, kb = u - tile * KB;
        mbar_arrive_expect_tx(full_bar(int(i)), stage_bytes);
        tma_load_2d(smem_base + i * stage_bytes, &tmap_w, int(kb * kBlockK), w_row0 + int(tile * kBlockN),
                    full_bar(int(i)), kPolicyEvictFirst);
      }
      // (2) activations exist only once the previous kernel has completed
      pdl_wait();
      for (uint32_t i = 0; i < n_pre; ++i) {
        const uint32_t u = u0 + i, kb = u % KB;
        tma_load_2d(smem_base + i * stage_bytes + kATileBytes, &tmap_x, int(kb * kBlockK), 0, full_bar(int(i)),
                    kPolicyEvictLast);
      }
      // (3) steady state
      int stage = int(n_pre % uint32_t(num_stages));
      uint32_t phase = (n_pre == uint32_t(num_stages)) ? 1u : 0u;
      for (uint32_t u = u0 + n_pre; u < u1; ++u) {
        const uint32_t tile = u / KB, kb = u - tile * KB;
        mbar_wait(empty_bar(stage), phase ^ 1);
        mbar_arrive_expect_tx(full_bar(stage), stage_bytes);
        const uint32_t sa = smem_base + uint32_t(stage) * stage_bytes;
        tma_load_2d(sa, &tmap_w, int(kb * kBlockK), w_row0 + int(tile * kBlockN), full_bar(stage), kPolicyEvictFirst);
        tma_load_2d(sa + kATileBytes, &tmap_x, int(kb * kBlockK), 0, full_bar(stage), kPolicyEvictLast);
        if (++stage == num_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16_f32(kBlockN, uint32_t(sk.m_tile));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      uint32_t u = u0;
      while (u < u1) {
        const uint32_t tile = u / KB;
        const uint32_t seg_end = min(u1, (tile + 1) * KB);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc) * uint32_t(sk.m_tile);
        uint32_t accumulate = 0;
        for (; u < seg_end; ++u) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + uint32_t(stage) * stage_bytes;
          const uint64_t da = umma_desc_sw128_kmajor(sa);
          const uint64_t db = umma_desc_sw128_kmajor(sa + kATileBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            umma_bf16_ss(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(empty_bar(stage));
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===== epilogue warps =====
    //  whole tile inside this CTA's range : TMEM -> smem staging -> fused epilogue, right away
    //  split tile (first / last segment)   : TMEM -> this CTA's fp32 partial slot; after the CTA's last segment the
    //                                        contributors of a tile share its rows (reduce-scatter): contributor j of
    //                                        k sums all k partials, in CTA order, for rows [jM/k, (j+1)M/k) and runs
    //                                        the fused epilogue on them — the fix-up is spread over all CTAs
    pdl_wait();  // everything below touches buffers the previous kernel may still be writing
    if (ep.mode == EPI_QKV) {
      for (int m = threadIdx.x - 64; m < ep.M; m += 128) s_pos[m] = make_int2(ep.rope_pos[m], ep.cache_pos[m]);
      epi_bar();
    }
    const int ew = warp - 2;       // 0..3
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int nrow = quarter * 32 + lane;  // weight row of this thread inside the tile
    const int tid_e = threadIdx.x - 64;    // 0..127
    const size_t slot_floats = sk.slot_floats();
    long long* dbg = (ep.dbg && tid_e == 0) ? ep.dbg + size_t(cta) * 8 : nullptr;
    if (dbg) dbg[0] = clock64();
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t u = u0;
    int n_whole = 0, first_whole = -1;   // whole tiles finished here (consecutive)
    int pend_tile[2] = {-1, -1};         // split tiles this CTA contributed to
    int n_pend = 0;
    while (u < u1) {
      const uint32_t tile = u / KB;
      const uint32_t seg_end = min(u1, (tile + 1) * KB);
      const bool whole = (tile * KB >= u0) && ((tile + 1) * KB <= u1);
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      if (dbg) dbg[(seg_end == u1) ? 2 : 1] = clock64();   // accumulator of a (1) non-last / (2) last segment ready
      const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc) * uint32_t(sk.m_tile);
      if (!whole) {
        // slot 2*cta: the CTA's first segment; 2*cta+1: its last segment (when that is a different, split tile)
        const int slot = 2 * cta + ((u == u0) ? 0 : 1);
        float* dst = ep.ws + size_t(slot) * slot_floats + nrow;
        for (int m0 = 0; m0 < sk.m_tile; m0 += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_addr + uint32_t(m0), v);
          tmem_ld_wait();
          if (m0 < ep.M) {
#pragma unroll
            for (int j = 0; j < 16; ++j) __stcg(dst + size_t(m0 + j) * 128, __uint_as_float(v[j]));
          }
        }
        tcgen05_fence_before();
        mbar_arrive(tempty_bar(acc));
        __threadfence();
        epi_bar();
        if (tid_e == 0) atomicAdd(&ep.tile_arrive[2 * tile], 1u);
        pend_tile[n_pend++] = int(tile);
      } else {
        const EpiTileConst tc = epi_tile_const(ep, int(tile), lane);
        for (int c0m = 0; c0m < sk.m_tile; c0m += kEpiChunk) {
          const int cw = min(kEpiChunk, sk.m_tile - c0m);
          for (int j0 = 0; j0 < cw; j0 += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_addr + uint32_t(c0m + j0), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) stage_tile[(j0 + j) * kBlockN + nrow] = __uint_as_float(v[j]);
          }
          if (c0m + kEpiChunk >= sk.m_tile) {  // accumulator fully drained
            tcgen05_fence_before();
            mbar_arrive(tempty_bar(acc));
          }
          epi_bar();  // staging tile complete
          {
            constexpr int R = kEpiChunk / 4;  // rows per warp per chunk, all in flight at once
            const int ml0 = ew * R, mrow0 = c0m + ml0;
            const int n_rows = min(min(R, cw - ml0), ep.M - mrow0);  // warp-uniform, may be <= 0
            EpiAux aux[R];
#pragma unroll
            for (int r = 0; r < R; ++r)
              if (r < n_rows) aux[r] = epi_load_aux(ep, tc, int(tile), mrow0 + r, lane, s_pos);
#pragma unroll
            for (int r = 0; r < R; ++r) {
              if (r < n_rows) {
                const float4 t = *reinterpret_cast<const float4*>(stage_tile + (ml0 + r) * kBlockN + 4 * lane);
                float v[4] = {t.x, t.y, t.z, t.w};
                epi_apply(ep, tc, aux[r], int(tile), mrow0 + r, lane, v, sk.m_tile);
              }
            }
          }
          epi_bar();  // rows done before the next chunk overwrites the staging tile
        }
        if (n_whole++ == 0) first_whole = int(tile);
      }
      u = seg_end;
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (dbg) dbg[3] = clock64();   // all segments drained / parked
    // ---- deferred fix-up of the split tiles: this CTA's share of the rows ------------------------------
    int share_lo[2] = {0, 0}, share_hi[2] = {0, 0};
    for (int pi = 0; pi < n_pend; ++pi) {
      const int tile = pend_tile[pi];
      const int c_first = sk.first_cta(tile), c_last = sk.last_cta(tile);
      const int k = c_last - c_first + 1, j = cta - c_first;
      if (tid_e == 0) {
        while (ld_acquire_u32(&ep.tile_arrive[2 * tile]) < uint32_t(k)) __nanosleep(32);
      }
      epi_bar();  // orders thread 0's acquire before everybody's partial reads
      const int r_lo = (j * ep.M) / k, r_hi = ((j + 1) * ep.M) / k;
      share_lo[pi] = r_lo;
      share_hi[pi] = r_hi;
      const EpiTileConst tc = epi_tile_const(ep, tile, lane);
      // the slot contributor c used for this tile: its first segment iff the tile holds the start of its range
      const int t_first_slot_cta = c_first;  // only c_first may have the tile as a non-first segment
      const bool first_uses_last_slot = (sk.begin(c_first) / KB) != uint32_t(tile);
      constexpr int R = 2, C = 4;
      for (int m = r_lo + R * ew; m < r_hi; m += 4 * R) {
        const int nr = min(R, r_hi - m);
        EpiAux aux[R];
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (r < nr) aux[r] = epi_load_aux(ep, tc, tile, m + r, lane, s_pos);
        float v[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r][0] = v[r][1] = v[r][2] = v[r][3] = 0.f;
        for (int cb = c_first; cb <= c_last; cb += C) {
          float4 pv[C][R];
#pragma unroll
          for (int cc = 0; cc < C; ++cc) {
            const int c = cb + cc;
            const int slot = 2 * c + ((c == t_first_slot_cta && first_uses_last_slot) ? 1 : 0);
            const float* pc = ep.ws + size_t(slot) * slot_floats + size_t(m) * 128 + 4 * lane;
#pragma unroll
            for (int r = 0; r < R; ++r)
              pv[cc][r] = (c <= c_last && r < nr) ? __ldcg(reinterpret_cast<const float4*>(pc + r * 128))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int cc = 0; cc < C; ++cc) {   // CTA order: bit-reproducible
            if (cb + cc <= c_last) {
#pragma unroll
              for (int r = 0; r < R; ++r) {
                v[r][0] += pv[cc][r].x; v[r][1] += pv[cc][r].y; v[r][2] += pv[cc][r].z; v[r][3] += pv[cc][r].w;
              }
            }
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (r < nr) epi_apply(ep, tc, aux[r], tile, m + r, lane, v[r], sk.m_tile);
      }
      // re-arm the tile's counters once every contributor has read the partials
      epi_bar();
      if (tid_e == 0) {
        const uint32_t done = atomicAdd(&ep.tile_arrive[2 * tile + 1], 1u) + 1u;
        if (done == uint32_t(k)) {
          ep.tile_arrive[2 * tile] = 0;
          ep.tile_arrive[2 * tile + 1] = 0;
        }
      }
    }
    if (dbg) dbg[4] = clock64();   // fix-up shares done
    // ---- EPI_RESID_NORM, second half: the row statistic needs every tile ------------------------------
    if (ep.mode == EPI_RESID_NORM) {
      const uint32_t my_rows = uint32_t(n_whole) * uint32_t(ep.M) + uint32_t(share_hi[0] - share_lo[0]) +
                               uint32_t(share_hi[1] - share_lo[1]);
      const uint32_t all_rows = uint32_t(sk.n_tiles) * uint32_t(ep.M);
      if (my_rows > 0) {
        __threadfence();
        epi_bar();
        if (tid_e == 0) {
          atomicAdd(&ep.ctr[0], my_rows);
          while (ld_acquire_u32(&ep.ctr[0]) < all_rows) __nanosleep(32);
          if (dbg) dbg[5] = clock64();   // every tile's statistic is in
        }
        epi_bar();
        const float inv_d = 1.f / float(ep.N);
        // pieces: [first_whole, first_whole + n_whole) x rows [0, M), then the two shares
        for (int piece = 0; piece < 3; ++piece) {
          int t_lo, t_n, r_lo, r_hi;
          if (piece == 0) { t_lo = first_whole; t_n = n_whole; r_lo = 0; r_hi = ep.M; }
          else { t_lo = pend_tile[piece - 1]; t_n = (piece - 1 < n_pend) ? 1 : 0; r_lo = share_lo[piece - 1]; r_hi = share_hi[piece - 1]; }
          if (t_n <= 0 || r_hi <= r_lo) continue;
          constexpr int R2 = 4;   // rows per warp in flight
          for (int mb = r_lo + ew * R2; mb < r_hi; mb += 4 * R2) {
            float rinv[R2];
#pragma unroll
            for (int r = 0; r < R2; ++r) {
              float sacc = 0.f;
              if (mb + r < r_hi)
                for (int t = lane; t < sk.n_tiles; t += 32) sacc += __ldcg(ep.ssq + size_t(t) * sk.m_tile + mb + r);
              rinv[r] = sacc;
            }
#pragma unroll
            for (int r = 0; r < R2; ++r) rinv[r] = rsqrtf(warp_sum(rinv[r]) * inv_d + ep.eps);
            for (int t = t_lo; t < t_lo + t_n; ++t) {
              float wv[4];
              unpack4(*reinterpret_cast<const uint2*>(ep.norm_w + t * kBlockN + 4 * lane), wv);
              uint2 hraw[R2];
#pragma unroll
              for (int r = 0; r < R2; ++r)
                if (mb + r < r_hi)
                  hraw[r] = *reinterpret_cast<const uint2*>(ep.h + size_t(mb + r) * ep.N + size_t(t) * kBlockN + 4 * lane);
#pragma unroll
              for (int r = 0; r < R2; ++r) {
                if (mb + r < r_hi) {
                  float hv[4], o[4];
                  unpack4(hraw[r], hv);
#pragma unroll
                  for (int e = 0; e < 4; ++e) o[e] = wv[e] * bf16_round(hv[e] * rinv[r]);
                  *reinterpret_cast<uint2*>(ep.xn + size_t(mb + r) * ep.N + size_t(t) * kBlockN + 4 * lane) = pack4(o);
                }
              }
            }
          }
        }
        epi_bar();
        if (tid_e == 0) {
          const uint32_t done = atomicAdd(&ep.ctr[1], my_rows) + my_rows;
          if (done == all_rows) {  // everybody is past the meeting point: re-arm for the next launch
            ep.ctr[0] = 0;
            ep.ctr[1] = 0;
          }
        }
      }
    }
    if (dbg) dbg[6] = clock64();   // epilogue warps done
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
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
  int g = grid_limit > 0 ? grid_limit : device_num_sms();
  uint32_t U = uint32_t(sk.n_tiles) * uint32_t(sk.kb);
  sk.grid = int(U < uint32_t(g) ? U : uint32_t(g));
  return sk;
}

// workspace layout (counters FIRST, so their place does not move with m_tile / grid and they stay zero):
//   [ctr: 4 u32][tile {arrive, done}: 2*cap u32, padded to 16 B][ssq: n_tiles * m_tile floats][2*grid partial slots]
struct GemmWorkspace {
  size_t ctr_off, arrive_off, ssq_off, slots_off, bytes;
};
// `arrive_cap` >= n_tiles reserves the counter region: a context sharing one workspace between GEMMs of different
// shapes passes the largest tile count so that no GEMM's data region ever overlaps another GEMM's counters.
GemmWorkspace gemm_workspace(const StreamK& sk, int arrive_cap = 0) {
  GemmWorkspace w;
  const int cap = arrive_cap > sk.n_tiles ? arrive_cap : sk.n_tiles;
  w.ctr_off = 0;
  w.arrive_off = 16;
  w.ssq_off = w.arrive_off + ((size_t(cap) * 8 + 15) & ~size_t(15));
  w.slots_off = w.ssq_off + size_t(sk.n_tiles) * sk.m_tile * 4;
  w.bytes = w.slots_off + 2 * size_t(sk.grid) * sk.slot_floats() * 4;
  return w;
}

struct GemmLaunch {
  CUtensorMap tmap_w, tmap_x;
  int w_row0;
  StreamK sk;
  int num_stages;
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
int gemm_shape(GemmLaunch* g, int m_tile) {
  if (m_tile % 16 != 0 || m_tile < 16 || m_tile > 256) return -3;
  const uint32_t stage_bytes = kATileBytes + uint32_t(m_tile) * kBlockK * 2;
  const uint32_t budget = uint32_t(gemm_smem_budget()) - 1024 - kEpiStageBytes;
  int stages = int(budget / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return -4;
  g->num_stages = stages;
  uint32_t cols = 32;
  while (cols < uint32_t(2 * m_tile)) cols <<= 1;
  g->tmem_cols = cols;
  g->smem_bytes = uint32_t(stages) * stage_bytes + kEpiStageBytes + 1024;
  return 0;
}

int gemm_attr_once() {
  static int rc = 1;
  if (rc == 1)
    rc = cudaFuncSetAttribute(gemm_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024) ==
                 cudaSuccess
             ? 0
             : -5;
  return rc;
}

long long* g_dbg_buf = nullptr;   // [g_dbg_cap launches][grid <= 256][8]
int g_dbg_cap = 0;
unsigned long long g_dbg_idx = 0;

int gemm_launch(const GemmLaunch* g, const GemmEpi& ep_in, cudaStream_t stream) {
  GemmEpi ep = ep_in;
  ep.dbg = g_dbg_buf ? g_dbg_buf + size_t(g_dbg_idx++ % uint64_t(g_dbg_cap)) * 256 * 8 : nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g->sk.grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = g->smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, gemm_fused_kernel, g->tmap_w, g->tmap_x, g->w_row0, g->sk, g->num_stages,
                            g->tmem_cols, ep) == cudaSuccess
             ? 0
             : -6;
}

}  // namespace sjd
