// Jacobi draft-window attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Same contract as attention.cu (key j visible to window query i of CFG row b iff kv_lo[b] <= j <= kv_len + i; one
// fp32 partial {sum p*v, max, sum p} per key split, merged by attn_combine_row in the next chain kernel), different
// machine mapping:
//   * work unit = (128-key tile, kv head [x query-row tile], CFG row); one persistent CTA per SM deals itself the
//     units cta, cta + grid, ... and runs them through a two-stage pipeline (TMA producer warp, one MMA-issuing thread,
//     two softmax groups of four warps — see the kernel's comment), so the K/V stream of the next tile, the tensor work
//     of this one and the softmax of two tiles overlap;
//   * the query rows of up to 128 / Wp heads that share the kv head are stacked along UMMA M (GQA: the K/V tile is
//     loaded once for all of them); rows beyond the stack are garbage that only ever produces garbage rows;
//   * S[128 q x 128 keys] = Q K^T :  A = Q tile, B = K tile, both K-major (head dim contiguous), SWIZZLE_128B boxes
//     straight from the q buffer / K cache by TMA;
//   * softmax on the TMEM accumulator with one thread per query row (row max and row sum are thread-local): two
//     passes of tcgen05.ld, probabilities rounded to bf16 like the reference's bf16 SDPA and written into the smem the Q
//     tile occupied (so the K tile is free for the next unit's K as soon as S exists), as the K-major A operand of the
//     second product;
//   * O[128 q x Dh] = P V :  B = the V tile exactly as TMA lands it ([key][head dim], head dim contiguous), consumed as
//     an MN-major operand — no transposed cache, no smem transpose;
//   * when the stacked rows fill only 1/rep of the 128 TMEM lanes (MHA, window 32: a quarter), Q is loaded rep times
//     and replica g takes columns [128 g / rep, 128 (g+1) / rep) of every row: all four softmax warps work, row max and
//     row sum meet in shared memory, every replica writes its slice of P into ALL replica rows, and replica g drains
//     its slice of the columns of O;
//   * a unit sees its whole key span at once, so there is no online rescaling: max, exponentials, one product.
// Reference semantics: SDPA over the additive window mask (modeling_chameleon.py:567-574, mask from
// scheduler/jacobi_iteration_lumina_mgpt.py:1256-1336; llamagen/llamagen.py:269-273).
#include "common.cuh"

namespace sjd {

constexpr int kTcKeys = 128;       // keys per unit (UMMA N of the first product, K of the second)
constexpr int kTcRows = 128;       // query-row slots per unit (UMMA M)

struct AttnTcParams {
  AttnParams a;          // geometry, partial buffers, kv_len / kv_lo (n_chunks = key tiles, span = kTcKeys)
  int head_dim;          // 64 | 128
  int Wp;                // row slots per head: W rounded up to 8 (swizzle atom) — the Q tensor map's box height
  int hpc;               // heads stacked per CTA
  int k_row0;            // first row of this layer in the [layers*rows*Hkv*Lmax, Dh] view of the caches
  long long* dbg;        // developer timing: CTA 0 writes clock64 stamps [unit < 8][16] (sjd_debug_attn_stamps)
  // exact division by multiplication for the small non-negative ints the unit decode needs: x / d == umulhi(x, m)
  uint32_t m_chunks, m_ny, m_mtiles, m_wp;
  int ny, mtiles;
};

__host__ __device__ __forceinline__ uint32_t tc_magic(uint32_t d) { return uint32_t((0x100000000ull + d - 1) / d); }
__device__ __forceinline__ int tc_div(int x, uint32_t magic, int d) {   // exact for 0 <= x < 2^16, 1 <= d < 2^16
  return d == 1 ? x : int(__umulhi(uint32_t(x), magic));
}

// UMMA descriptor of an MN-major bf16 operand laid out by TMA with SWIZZLE_128B: rows are K (keys), each row holds 64
// consecutive MN elements (128 bytes); 8-row groups are 1024 B apart (stride byte offset), the next 64 MN elements are
// `mn_chunk_bytes` away (leading byte offset).  Canonical form ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units.
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t mn_chunk_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((mn_chunk_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16 instruction descriptor with B taken MN-major (bit 16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32_bmn(uint32_t M, uint32_t N) {
  return umma_idesc_bf16_f32(M, N) | (1u << 16);
}

// 2^x for x <= 0 (softmax exponents): one MUFU op; results below the normal range flush to zero
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// mbarrier wait with back-off for the single-thread roles: a hot try_wait loop would steal issue slots from the
// softmax warp that shares its scheduler
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(32);
}

struct AttnTcMaps {
  CUtensorMap q, k, v;
};

// Work unit u -> (key tile, kv head x row tile, CFG row); key tiles run fastest so neighbouring CTAs share Q in L2.
struct TcUnit {
  int kt, hkv, h0, heads_here, b, key0, lo;
  bool hidden;
  // softmax-thread view (TMEM lane r = 32 * warp + lane)
  int R, Rr, rep, g, rr, hs, qi;
};

__device__ __forceinline__ TcUnit tc_unit(const AttnTcParams& p, int u, int r) {
  const AttnParams& a = p.a;
  TcUnit t;
  const int G = a.H / a.Hkv;
  const int q1 = tc_div(u, p.m_chunks, a.n_chunks);          // u / n_chunks
  t.kt = u - q1 * a.n_chunks;
  t.b = tc_div(q1, p.m_ny, p.ny);                            // (u / n_chunks) / ny
  const int y = q1 - t.b * p.ny;
  t.hkv = tc_div(y, p.m_mtiles, p.mtiles);
  const int mt = y - t.hkv * p.mtiles;
  t.h0 = t.hkv * G + mt * p.hpc;
  t.heads_here = min(p.hpc, G - mt * p.hpc);
  t.key0 = t.kt * kTcKeys;
  t.lo = a.kv_lo[t.b];
  t.hidden = t.key0 + kTcKeys <= t.lo;     // whole tile inside the hidden prefix: nothing to load
  t.R = t.heads_here * p.Wp;               // row slots in use
  t.Rr = (t.R + 31) & ~31;                 // replica pitch: whole warps (tcgen05.ld is warp-collective: uniform columns)
  t.rep = t.Rr <= 32 ? 4 : (t.Rr <= 64 ? 2 : 1);
  t.g = t.Rr == 32 ? (r >> 5) : (t.Rr == 64 ? (r >> 6) : (r >= t.Rr ? 1 : 0));
  t.rr = r - t.g * t.Rr;
  t.hs = tc_div(t.rr, p.m_wp, p.Wp);
  t.qi = t.rr - t.hs * p.Wp;
  return t;
}

// Persistent: one CTA per SM walks units u = cta, cta + grid, ... through a two-stage pipeline
//   warp 4 (TMA)  : K of unit n+2 as soon as S(n) exists; Q and V of unit n+2 once O(n) has read P (which sits on Q) and V
//   warp 5 (MMA)  : S(n) as soon as Q, K landed; then O(n-1) = P(n-1) V(n-1) once the softmax warps delivered P(n-1)
//   warps 0..3, 8..11 : two softmax groups; group k owns pipeline stage k, i.e. every other unit: softmax of S(n) into
//                   P(n), then the epilogue of its previous unit n-2 (whose O has long been ready)
// so the K/V stream of the next tile, the tensor work of this one and the softmax of two tiles overlap.  TMEM: stage s owns columns
// [256 s, 256 s + 128) for S and [256 s + 128, 256 s + 128 + Dh) for O.
constexpr int kTcThreads2 = 384;   // (attention_tct.cu) warps 0..3 and 8..11: the two softmax groups; 4: TMA; 5: MMA; 6, 7: idle
// attn_tc_kernel: each softmax group has EIGHT warps — two per TMEM lane quarter, each taking half of the row's columns
// (a lone warp per scheduler runs at ~0.2 IPC; the halves meet through the same shared-memory exchange as the replicas):
// warps 0..3 / 12..15 = group 0 halves 0 / 1, warps 8..11 / 16..19 = group 1 halves 0 / 1, 4 = TMA, 5 = MMA, 6, 7 idle
constexpr int kTcThreads3 = 640;

template <int DH>
__global__ void __launch_bounds__(kTcThreads3, 1)
attn_tc_kernel(const __grid_constant__ AttnTcMaps maps, const AttnTcParams p) {
  constexpr int NDA = DH / 64;                            // 64-wide head-dim atoms
  constexpr uint32_t kQBytes = NDA * kTcRows * 128;       // Q tile: NDA atoms of [128 rows][128 B]
  constexpr uint32_t kKBytes = NDA * kTcKeys * 128;       // K tile, same shape
  constexpr uint32_t kPBytes = 2 * kTcRows * 128;         // P tile: two 64-key atoms of [128 rows][128 B]
  // P is written where Q was (both are dead / born at "S done"), so the K tile is free the moment S has been computed
  // and the next unit's K can land while this unit's softmax runs; Q (from L2) and V follow once P V has drained
  constexpr uint32_t kQPBytes = kQBytes > kPBytes ? kQBytes : kPBytes;
  constexpr uint32_t kVBytes = NDA * kTcKeys * 128;       // V tile: NDA boxes of [128 keys][128 B]
  constexpr uint32_t kStage = kQPBytes + kKBytes + kVBytes;
  extern __shared__ uint8_t smem_raw[];
  // per stage: q landed, v landed, S done, P ready, O done, O drained, k landed
  __shared__ __align__(8) uint64_t bars[14];
  __shared__ uint32_t tmem_holder;
  __shared__ float xch_all[2 * 8 * kTcRows];              // per softmax group: {row max, row sum} x 2 column halves x 128 lanes, two units deep
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const AttnParams& a = p.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int which, int s) { return smem_u32(&bars[which * 2 + s]); };
  enum { B_Q = 0, B_V = 1, B_S = 2, B_P = 3, B_O = 4, B_E = 5, B_K = 6 };
  const int G = a.H / a.Hkv;
  const int n_units = a.n_chunks * a.Hkv * ((G + p.hpc - 1) / p.hpc) * a.rows;
  const int T = a.kv_len + a.W;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&maps.q);
      tma_prefetch_desc(&maps.k);
      tma_prefetch_desc(&maps.v);
      for (int s = 0; s < 2; ++s) {
        mbar_init(bar(B_Q, s), 1);
        mbar_init(bar(B_K, s), 1);
        mbar_init(bar(B_V, s), 1);
        mbar_init(bar(B_S, s), 1);
        mbar_init(bar(B_P, s), 8);
        mbar_init(bar(B_O, s), 1);
        mbar_init(bar(B_E, s), 4);
      }
      fence_barrier_init();
    }
  } else if (warp == 5) {
    tmem_alloc(smem_u32(&tmem_holder), 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_holder;
  // Keys below kv_len were cached by earlier forwards: the producer may stream the K/V tiles of its first two units
  // while the previous kernel is still finishing (everything else — q, the window's own K/V rows, the partial
  // buffers — belongs to the previous kernels until griddepcontrol.wait returns).
  int early = 0;   // units whose K/V were requested before the wait (producer warp only)
  if (warp == 4) {
    // Everything this CTA will stream goes to L2 now (fire and forget; L2 is the coherence point, so rows the previous
    // kernel is still writing are simply written into the prefetched lines): HBM works through the whole K/V span
    // from the first microsecond, the pipeline below then loads from L2 with a third of the latency.
    if (a.l2_prefetch) {
      int i = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const TcUnit t = tc_unit(p, u, 0);
        if (t.hidden) continue;
        const int krow = p.k_row0 + (t.b * a.Hkv + t.hkv) * a.Lmax + t.key0;
        for (int x = 0; x < 2 * NDA; ++x, ++i)
          if ((i & 31) == lane) tma_prefetch_l2_2d(x < NDA ? &maps.k : &maps.v, (x % NDA) * 64, krow);
      }
    }
    int n = 0;
    for (int u = blockIdx.x; u < n_units && n < 2; u += gridDim.x) {
      const TcUnit t = tc_unit(p, u, 0);
      if (t.hidden) continue;
      if (t.key0 + kTcKeys > a.kv_len) break;     // touches this window's keys: not before the wait
      const int s = n & 1;
      const uint32_t sK = base + uint32_t(s) * kStage + kQPBytes, sV = sK + kKBytes;
      if (lane == 0) {
        mbar_arrive_expect_tx(bar(B_K, s), kKBytes);
        mbar_arrive_expect_tx(bar(B_V, s), kVBytes);
      }
      __syncwarp();
      const int krow = p.k_row0 + (t.b * a.Hkv + t.hkv) * a.Lmax + t.key0;
      if (lane < NDA) tma_load_2d(sK + uint32_t(lane) * kTcKeys * 128, &maps.k, lane * 64, krow, bar(B_K, s), kPolicyEvictFirst);
      else if (lane < 2 * NDA)
        tma_load_2d(sV + uint32_t(lane - NDA) * kTcKeys * 128, &maps.v, (lane - NDA) * 64, krow, bar(B_V, s), kPolicyEvictFirst);
      ++n;
    }
    early = n;
  }
  // Wait first, THEN let the next kernel in: releasing the dependents before the wait lets the whole forward cascade
  // into residency (chain l+1 behind attention l+1 behind chain l ...), which was measured to deadlock.
  pdl_wait();
  pdl_launch_dependents();
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[15] = clock64();   // start of work

  if (warp == 4) {
    // ===== TMA producer: lane 0 arms the barriers, then every lane issues its share of the unit's box loads (a TMA
    // issue costs a thread ~100 ns: K, V and the replicated Q boxes of a unit are up to a dozen) =====
    int n = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const TcUnit t = tc_unit(p, u, 0);
      if (t.hidden) continue;
      const int s = n & 1, j = n >> 1;
      const uint32_t sQ = base + uint32_t(s) * kStage, sK = sQ + kQPBytes, sV = sK + kKBytes;
      const int nq = t.rep * t.heads_here * NDA;               // Q boxes
      const bool kv_done = n < early;                        // K/V of this unit are in flight already
      const int krow = p.k_row0 + (t.b * a.Hkv + t.hkv) * a.Lmax + t.key0;
      // K first: its smem is free as soon as the stage's previous S has been computed
      if (!kv_done) {
        if (lane == 0) {
          if (j >= 1) mbar_wait_backoff(bar(B_S, s), uint32_t(j - 1) & 1u);
          mbar_arrive_expect_tx(bar(B_K, s), kKBytes);
        }
        __syncwarp();
        if (lane < NDA)
          tma_load_2d(sK + uint32_t(lane) * kTcKeys * 128, &maps.k, lane * 64, krow, bar(B_K, s), kPolicyEvictFirst);
      }
      // Q (it lands where the previous P was) and V: once the stage's previous P V has read them
      if (lane == 0) {
        if (j >= 1) mbar_wait_backoff(bar(B_O, s), uint32_t(j - 1) & 1u);
        mbar_arrive_expect_tx(bar(B_Q, s), uint32_t(nq) * uint32_t(p.Wp) * 128u);
        if (!kv_done) mbar_arrive_expect_tx(bar(B_V, s), kVBytes);
      }
      __syncwarp();
      for (int l = lane; l < NDA + nq; l += 32) {
        if (l < nq) {
          const int d = l % NDA, gh = l / NDA;                  // NDA is 1 or 2
          const int gq = gh / t.heads_here, hs = gh - gq * t.heads_here;
          tma_load_2d(sQ + uint32_t(d) * kTcRows * 128 + uint32_t(gq * t.Rr + hs * p.Wp) * 128, &maps.q,
                      (t.h0 + hs) * DH + d * 64, t.b * a.W, bar(B_Q, s), kPolicyEvictLast);
        } else if (!kv_done) {
          const int d = l - nq;
          tma_load_2d(sV + uint32_t(d) * kTcKeys * 128, &maps.v, d * 64, krow, bar(B_V, s), kPolicyEvictFirst);
        }
      }
      if (p.dbg && blockIdx.x == 0 && lane == 0 && n < 8) p.dbg[n * 16 + 0] = clock64();
      ++n;
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16_f32(kTcRows, kTcKeys);
      const uint32_t idesc_o = umma_idesc_bf16_f32_bmn(kTcRows, DH);
      auto issue_pv = [&](int m) {   // O(m) = P(m) V(m)
        const int s = m & 1, j = m >> 1;
        const uint32_t sP = base + uint32_t(s) * kStage, sV = sP + kQPBytes + kKBytes;
        mbar_wait_backoff(bar(B_V, s), uint32_t(j) & 1u);
        mbar_wait_backoff(bar(B_P, s), uint32_t(j) & 1u);
        if (j >= 1) mbar_wait_backoff(bar(B_E, s), uint32_t(j - 1) & 1u);   // the previous O of this stage has been drained
        tcgen05_fence_after();
        if (p.dbg && blockIdx.x == 0 && m < 8) p.dbg[m * 16 + 3] = clock64();
        const uint32_t tO = tmem_base + uint32_t(s) * 256 + 128;
        uint32_t acc = 0;
#pragma unroll
        for (int ka = 0; ka < 2; ++ka) {
          const uint64_t da = umma_desc_sw128_kmajor(sP + uint32_t(ka) * kTcRows * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t db = umma_desc_sw128_mnmajor(sV + uint32_t(ka * 64 + k * 16) * 128, kTcKeys * 128);
            umma_bf16_ss(tO, da + uint64_t(2 * k), db, idesc_o, acc);
            acc = 1;
          }
        }
        umma_commit(bar(B_O, s));
        if (p.dbg && blockIdx.x == 0 && m < 8) p.dbg[m * 16 + 4] = clock64();
      };
      auto issue_s = [&](int m) {    // S(m) = Q(m) K(m)^T
        const int s = m & 1;
        const uint32_t sQ = base + uint32_t(s) * kStage, sK = sQ + kQPBytes;
        tcgen05_fence_after();
        const uint32_t tS = tmem_base + uint32_t(s) * 256;
        uint32_t acc = 0;
#pragma unroll
        for (int d = 0; d < NDA; ++d) {
          const uint64_t da = umma_desc_sw128_kmajor(sQ + uint32_t(d) * kTcRows * 128);
          const uint64_t db = umma_desc_sw128_kmajor(sK + uint32_t(d) * kTcKeys * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16_ss(tS, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc_s, acc);
            acc = 1;
          }
        }
        umma_commit(bar(B_S, s));
        if (p.dbg && blockIdx.x == 0 && m < 8) p.dbg[m * 16 + 2] = clock64();
      };
      int N = 0;   // units this CTA really runs
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) N += tc_unit(p, u, 0).hidden ? 0 : 1;
      // Issue whichever product has its inputs first: S(ns) needs Q, K of unit ns (and the S accumulator of its stage,
      // free once O(ns - 2) has been issued: that waited for P(ns - 2), i.e. for the last read of S); O(np) needs V and P
      // of unit np and the drained O accumulator.  In-order issue would park a ready O behind the next tile's loads.
      int ns = 0, np = 0;
      while (np < N) {
        bool did = false;
        if (np < ns) {
          const int s = np & 1, j = np >> 1;
          if (mbar_test_wait(bar(B_P, s), uint32_t(j) & 1u) && mbar_test_wait(bar(B_V, s), uint32_t(j) & 1u) &&
              (j < 1 || mbar_test_wait(bar(B_E, s), uint32_t(j - 1) & 1u))) {
            issue_pv(np);
            ++np;
            did = true;
          }
        }
        if (ns < N && ns - np < 2) {
          const int s = ns & 1, j = ns >> 1;
          if (mbar_test_wait(bar(B_K, s), uint32_t(j) & 1u) && mbar_test_wait(bar(B_Q, s), uint32_t(j) & 1u)) {
            if (p.dbg && blockIdx.x == 0 && ns < 8) p.dbg[ns * 16 + 1] = clock64();
            issue_s(ns);
            ++ns;
            did = true;
          }
        }
        if (!did) __nanosleep(20);
      }
    }
  } else if (warp < 4 || warp >= 8) {
    // ===== softmax + epilogue: one thread per query-row slot, straight from the TMEM accumulators.  Group `grp` takes
    // the units that run through pipeline stage `grp` =====
    const int widx = warp < 4 ? warp : warp - 4;   // 0..15
    const int qw = widx & 3, grp = (widx >> 2) & 1, half = widx >> 3;   // TMEM lane quarter = warp % 4
    const int r = qw * 32 + lane;
    const uint32_t t_row = (uint32_t(qw * 32) << 16);
    float* const xch = xch_all + grp * 8 * kTcRows;
    // state of the unit whose epilogue is still owed
    bool e_row_ok = false, e_first = false;
    size_t e_prow = 0;
    float e_mx = 0.f, e_l = 0.f;
    int e_c0 = 0, e_nc = 0, e_slot0 = 0, e_R = 0, e_h0 = 0;
    size_t e_base = 0;
    // per-warp transpose tile [32 rows][36 floats] behind the pipeline stages: TMEM hands a thread one ROW, global
    // memory wants a warp to write whole 128-byte lines
    float* const tile = reinterpret_cast<float*>(smem_raw + (base + 2 * kStage - smem_u32(smem_raw))) + (grp * 4 + qw) * (32 * 20);
    // O(m): every replica row holds the full product (P is written into all of them), so replica g drains columns
    // [g * Dh / rep, (g + 1) * Dh / rep) of its row: all four warps share the stores
    auto epilogue = [&](int m) {
      const int s = m & 1, j = m >> 1;
      mbar_wait(bar(B_O, s), uint32_t(j) & 1u);
      tcgen05_fence_after();
      if (p.dbg && blockIdx.x == 0 && qw == 0 && lane == 0 && m < 8) p.dbg[m * 16 + 8] = clock64();   // (half 0 only runs this)
      const uint32_t tO = tmem_base + uint32_t(s) * 256 + 128;
#pragma unroll 1
      for (int c = e_c0; c < e_c0 + e_nc; c += 16) {   // e_nc is a multiple of 16; warp-uniform
        uint32_t v[16];
        tmem_ld_32x32b_x16(tO + t_row + uint32_t(c), v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(tile + lane * 20 + 4 * q) =
              make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                          __uint_as_float(v[4 * q + 3]));
        __syncwarp();
        // lane -> (row of the tile, float4 of the row): 8 rows of 64 bytes per pass
        const int q = lane & 3;
#pragma unroll
        for (int r0 = 0; r0 < 32; r0 += 8) {
          const int rw = r0 + (lane >> 2);
          const int slot = e_slot0 + rw;                // row slot inside the replica
          const int hs2 = tc_div(slot, p.m_wp, p.Wp), qi2 = slot - hs2 * p.Wp;
          if (slot < e_R && qi2 < a.W) {
            const float4 x = *reinterpret_cast<const float4*>(tile + rw * 20 + 4 * q);
            *reinterpret_cast<float4*>(a.part_o + (e_base + size_t(e_h0 + hs2) * a.W + qi2) * DH + c + 4 * q) = x;
          }
        }
        __syncwarp();
      }
      if (e_row_ok && e_first) {
        a.part_ml[e_prow * 2] = e_mx;
        a.part_ml[e_prow * 2 + 1] = e_l;
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_E, s));
      if (p.dbg && blockIdx.x == 0 && qw == 0 && lane == 0 && m < 8) p.dbg[m * 16 + 9] = clock64();
    };
    int n = 0, owed = -1;   // owed: this group's unit whose epilogue has not run yet
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const TcUnit t = tc_unit(p, u, r);
      const bool row_ok = t.rr < t.R && t.qi < a.W && t.g < t.rep;   // this thread works on a real query row
      const bool first = t.g == 0 && half == 0;                      // replica 0, half 0 also owns {max, sum}
      const size_t prow = row_ok ? ((size_t(t.kt) * a.rows + t.b) * a.H + (t.h0 + t.hs)) * size_t(a.W) + t.qi : 0;
      if (t.hidden) {   // uniform per CTA
        if (grp == 0 && row_ok && first) {   // (first implies half 0)
          a.part_ml[prow * 2] = -INFINITY;
          a.part_ml[prow * 2 + 1] = 0.f;
        }
        continue;
      }
      if ((n & 1) != grp) {   // the other group's unit
        ++n;
        continue;
      }
      const int s = n & 1, j = n >> 1;
      const int lo = t.lo, key0 = t.key0, rep = t.rep, g = t.g, rr = t.rr;
      const int ncol = kTcKeys / rep / 2, col0 = g * (kTcKeys / rep) + half * ncol;   // this thread's slice of the row's 128 keys (16 | 32 | 64)
      uint8_t* const genP = smem_raw + (base + uint32_t(s) * kStage - smem_u32(smem_raw));   // over the Q tile
      const uint32_t tS = tmem_base + uint32_t(s) * 256;
      float* const xmax = xch + (j & 1) * 4 * kTcRows;        // exchange buffers alternate between units: no barrier
      float* const xsum = xmax + 2 * kTcRows;                 //   is needed to protect their reuse; [half][lane]
      // the group's previous unit first: its O has long been ready, and draining it now lets O(n) be issued the
      // moment P(n) is delivered
      if (half == 0 && owed >= 0) epilogue(owed);
      mbar_wait(bar(B_S, s), uint32_t(j) & 1u);
      tcgen05_fence_after();
      if (p.dbg && blockIdx.x == 0 && half == 0 && qw == 0 && lane == 0 && n < 8) p.dbg[n * 16 + 5] = clock64();
      const int j_hi = min(a.kv_len + t.qi, T - 1);           // last visible key of this row
      const float sc = a.scale_log2e;
      // ---- pass 1: row max over this thread's columns (four independent chains), 16 columns per step ----
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (g < rep) {
#pragma unroll 1
        for (int c = col0; c < col0 + ncol; c += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(tS + t_row + uint32_t(c), v);
          tmem_ld_wait();
          const int jk0 = key0 + c;
          const int e_lo = lo - jk0, e_hi = j_hi - jk0;   // visible columns of the step: [e_lo, e_hi]
#pragma unroll
          for (int e = 0; e < 16; ++e)
            m4[e & 3] = fmaxf(m4[e & 3], (e >= e_lo && e <= e_hi) ? __uint_as_float(v[e]) : -INFINITY);
        }
      }
      float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      {   // the row max is spread over the column halves and the replicas
        xmax[half * kTcRows + r] = mx;
        asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory");
        if (g < rep) {
          for (int g2 = 0; g2 < rep; ++g2)
            mx = fmaxf(mx, fmaxf(xmax[g2 * t.Rr + rr], xmax[kTcRows + g2 * t.Rr + rr]));
        }
      }
      mx *= sc;                                               // scale > 0: max commutes with it
      if (p.dbg && blockIdx.x == 0 && half == 0 && qw == 0 && lane == 0 && n < 8) p.dbg[n * 16 + 6] = clock64();
      const float ms = (mx == -INFINITY) ? 0.f : mx;
      // ---- pass 2: probabilities (bf16, like the reference's bf16 SDPA) into every replica's row of the P tile ----
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
      if (g < rep) {
#pragma unroll 1
        for (int c = col0; c < col0 + ncol; c += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(tS + t_row + uint32_t(c), v);
          tmem_ld_wait();
          const int jk0 = key0 + c;
          const int e_lo = lo - jk0, e_hi = j_hi - jk0;
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            float e0 = ex2_approx(fmaf(__uint_as_float(v[e]), sc, -ms));
            float e1 = ex2_approx(fmaf(__uint_as_float(v[e + 1]), sc, -ms));
            if (!(e >= e_lo && e <= e_hi)) e0 = 0.f;
            if (!(e + 1 >= e_lo && e + 1 <= e_hi)) e1 = 0.f;
            const __nv_bfloat162 pb = __floats2bfloat162_rn(e0, e1);   // probabilities enter P*V as bf16
            l4[(e >> 1) & 3] += __low2float(pb) + __high2float(pb);
            pk[e >> 1] = *reinterpret_cast<const uint32_t*>(&pb);
          }
          if (row_ok) {
            // K-major P tile: 64-key atom c / 64, two 16-byte chunks from (c % 64) / 8, 128B-swizzled by the row;
            // written into every replica's row
            const int ch = (c & 63) >> 3;
            for (int g2 = 0; g2 < rep; ++g2) {
              const int prw = g2 * t.Rr + rr;
              uint8_t* rowp = genP + (c >> 6) * (kTcRows * 128) + prw * 128;
              *reinterpret_cast<uint4*>(rowp + (((ch) ^ (prw & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ (prw & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
      }
      float lsum = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      tcgen05_fence_before();
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_P, s));
      if (p.dbg && blockIdx.x == 0 && half == 0 && qw == 0 && lane == 0 && n < 8) p.dbg[n * 16 + 7] = clock64();
      {
        xsum[half * kTcRows + r] = lsum;
        asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory");
        if (first) {
          lsum = 0.f;
          for (int g2 = 0; g2 < rep; ++g2) lsum += xsum[g2 * t.Rr + rr] + xsum[kTcRows + g2 * t.Rr + rr];   // fixed order
        }
      }
      if (half == 0) owed = n;
      e_row_ok = row_ok; e_first = first; e_prow = prow; e_mx = mx; e_l = lsum;
      e_nc = g < rep ? DH / rep : 0; e_c0 = g * e_nc;   // (a warp beyond the last replica drains nothing)
      e_slot0 = qw * 32 - g * t.Rr; e_R = t.R; e_h0 = t.h0;
      e_base = (size_t(t.kt) * a.rows + t.b) * a.H * size_t(a.W);
      ++n;
    }
    if (half == 0 && owed >= 0) epilogue(owed);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Geometry of the tensor-core attention for a window of W tokens per row.
void attn_tc_plan(AttnTcParams* p, int head_dim) {
  AttnParams& a = p->a;
  const int T = a.kv_len + a.W, G = a.H / a.Hkv;
  p->head_dim = head_dim;
  p->Wp = (a.W + 7) & ~7;
  int hpc = kTcRows / p->Wp;
  if (hpc > G) hpc = G;
  if (hpc < 1) hpc = 1;
  p->hpc = hpc;
  a.span = kTcKeys;
  a.n_chunks = (T + kTcKeys - 1) / kTcKeys;
  p->mtiles = (G + hpc - 1) / hpc;
  p->ny = a.Hkv * p->mtiles;
  p->m_chunks = tc_magic(uint32_t(a.n_chunks));
  p->m_ny = tc_magic(uint32_t(p->ny));
  p->m_mtiles = tc_magic(uint32_t(p->mtiles));
  p->m_wp = tc_magic(uint32_t(p->Wp));
}

constexpr int attn_tc_smem(int head_dim) {   // two pipeline stages of {Q | P, K, V} + eight 32 x 20 fp32 transpose tiles
  return 1024 + 2 * ((head_dim / 64) * kTcRows * 128 + 2 * kTcRows * 128 + (head_dim / 64) * kTcKeys * 128) + 8 * 32 * 20 * 4;
}

int attn_tc_launch(const AttnTcMaps& maps, const AttnTcParams& p, cudaStream_t stream) {
  const AttnParams& a = p.a;
  if (a.W > kTcRows || (p.head_dim != 64 && p.head_dim != 128)) return -3;
  const int G = a.H / a.Hkv, mtiles = (G + p.hpc - 1) / p.hpc;
  const int n_units = a.n_chunks * a.Hkv * mtiles * a.rows;
  const int sms = device_num_sms();
  dim3 grid(n_units < sms ? n_units : sms);
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(attn_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_tc_smem(128)) != cudaSuccess ||
        cudaFuncSetAttribute(attn_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_tc_smem(64)) != cudaSuccess)
      return -5;
    set = true;
  }
  if (p.head_dim == 128) return launch_pdl(attn_tc_kernel<128>, grid, dim3(kTcThreads3), attn_tc_smem(128), stream, maps, p);
  return launch_pdl(attn_tc_kernel<64>, grid, dim3(kTcThreads3), attn_tc_smem(64), stream, maps, p);
}

}  // namespace sjd
