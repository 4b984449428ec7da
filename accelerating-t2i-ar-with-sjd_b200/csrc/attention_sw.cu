// Small-window Jacobi attention, round-2 form: the transposed tcgen05 formulation of attention_tct.cu
//   S^T[128 keys x N q] = K Q^T,   O^T[128 d x N q] = V^T P^T,   L[128 x N q] = 1 P^T        (N = 32 | 64 query slots)
// behind a pipeline built for the HBM stream and for the latency of lone warps instead of for the tensor core
// (DESIGN.md 3.2a has the measurements behind every choice):
//   * every CTA owns a CONTIGUOUS range of (CFG row, kv head [x row tile], key tile) units, key tiles fastest, so it walks
//     along the keys of one head: Q is loaded once per head (two buffers), and the products of consecutive key tiles
//     ACCUMULATE in TMEM (O^T and L) under one reference maximum per 8-column group;
//   * CLUSTER MODE (default): a head belongs to one thread-block cluster of 1, 2 or 4
//     CTAs, each with ONE accumulator (a tile that outgrows the reference rescales it in place through tcgen05.ld / .st);
//     at the end the peers stage {O^T, L, m} in their own shared memory and the cluster leader merges them through
//     distributed shared memory and writes the NORMALISED bf16 attention rows — no fp32 partials in global memory, no
//     merge pre-op in the chain kernel that follows;
//   * SEGMENT FORM (runs > SMs, SJD_ATTN_SW_CLUSTER=0, test knobs): a tile that outgrows the reference by more than 2^grow, or a
//     new head, starts a new "segment" in the other accumulator; one fp32 partial {sum p v, m, sum p} per segment goes to
//     the partial slot of its first key tile (the other slots get the empty marker {-inf, 0}) and the next chain kernel's
//     pre-op merges them (attn_combine_row, sparse form);
//   * K and V tiles arrive through separate TMA rings (K: 2 x 32 KB, released by the commit of S^T; V: 2-3 x 32 KB,
//     released by the commit of O^T), P^T has its own two buffers: a K tile is free ~0.1 us after it landed, so the
//     next tiles stream while the softmax of this one runs; tiles of old keys are requested before griddepcontrol.wait;
//   * softmax: thread = key (TMEM lane), 16 warps (four per lane quarter, 8-16 columns each), a mask-free path for tiles
//     every query sees completely.  A probability only has to be scaled by a bound that is COMMON to the 128 keys of a
//     column and recorded with the result, not by the exact column maximum: each thread takes the max of 8 adjacent
//     columns (draft positions i..i+7 of one head), one redux.sync per group, the warps meet through 128 bytes of shared
//     memory and ONE named barrier, and the bound is rounded up to an INTEGER in the log2 domain — probabilities scaled
//     against different integers differ by exact powers of two, so the result does not depend on the reference (nor,
//     therefore, on the window a token shares);
//   * the MMA warp runs its issue loop warp-uniformly with only tcgen05.mma / commit under elect.sync, descriptors are
//     built once, barriers seen complete are not polled again, units are walked incrementally, every role decodes what
//     it needs before the dependency wait and waits inside its own branch.
// Mask, partial format, merge and reference semantics as in attention_tc.cu (SDPA over the additive window mask,
// modeling_chameleon.py:567-574, scheduler/jacobi_iteration_lumina_mgpt.py:1256-1336).  Head dim 128 only.
#include "common.cuh"

namespace sjd {

constexpr int kSwThreads = 576;                          // warps 0-15: softmax + epilogue (lane quarter = warp & 3, column quarter = warp >> 2); 16: TMA; 17: MMA
constexpr int kSwTmaWarp = 16, kSwMmaWarp = 17;
constexpr int kSwSoftThreads = 512;
constexpr uint32_t kSwTileBytes = 2 * kTcKeys * 128;     // a K or V tile: two 64-wide head-dim boxes of [128 keys][128 B]
constexpr uint32_t kSwPBytes = kTcKeys * 128;            // P^T: [128 keys][64 query slots] bf16

struct AttnSwParams {
  AttnTcParams t;   // geometry as attn_tct_plan leaves it
  int ncols;        // accumulator columns per unit: 32 | 64
  int nv;           // depth of the V ring
  int grid_cap;     // > 0: at most this many CTAs (test knob: small shapes then exercise runs, ring wrap-around, segments)
  float grow;       // a tile whose maximum exceeds the segment's reference by more than this (log2 units) starts a new segment
  // CTA c owns units [ub[c], ub[c + 1]).  Not an equal split: per-CTA stamps (profiles/r02aa_attn_sw_classes.txt) show a
  // CTA pays ~1.5 us per key tile (the HBM stream) and ~1.3 us more when its range crosses from one head to the next (the
  // old segment's epilogue, a new Q), so the host balances  units + run_cost * crossings  (attn_sw_split)
  uint16_t ub[150];
  // Cluster mode (cluster >= 1): every run (CFG row, kv head, row tile) belongs to ONE thread-block cluster of `cluster`
  // CTAs, each walking a contiguous slice of the run's key tiles into a single TMEM accumulator (a tile whose maximum
  // outgrows the reference rescales the accumulator in place instead of starting a segment).  At the end the CTAs stage
  // their {O^T, L, m} in their own shared memory, the cluster leader reads them through distributed shared memory, merges
  // and writes the NORMALISED bf16 attention rows: no fp32 partials in global memory, no merge pre-op and no grid-wide
  // wait for it in the chain kernel that follows.  0: the segment / partial-slot form above (runs > SMs, test knobs).
  int cluster;
};
constexpr int kSwMaxGrid = 148;

struct SwUnit {
  int kt, b, hkv, h0, heads, key0, lo, run, R;
  bool hidden;
};

__device__ __forceinline__ SwUnit sw_unit(const AttnTcParams& p, int u) {
  const AttnParams& a = p.a;
  SwUnit t;
  const int G = a.H / a.Hkv;
  t.run = tc_div(u, p.m_chunks, a.n_chunks);                 // (CFG row, kv head, row tile): what shares Q
  t.kt = u - t.run * a.n_chunks;
  t.b = tc_div(t.run, p.m_ny, p.ny);
  const int y = t.run - t.b * p.ny;
  t.hkv = tc_div(y, p.m_mtiles, p.mtiles);
  const int mt = y - t.hkv * p.mtiles;
  t.h0 = t.hkv * G + mt * p.hpc;
  t.heads = min(p.hpc, G - mt * p.hpc);
  t.key0 = t.kt * kTcKeys;
  t.lo = a.kv_lo[t.b];
  t.hidden = t.key0 + kTcKeys <= t.lo;                       // whole tile inside the hidden prefix
  t.R = t.heads * p.Wp;
  return t;
}

// Units of a CTA are consecutive: walking them costs an add and a compare (a full decode is ~40 dependent
// instructions, 0.1-0.3 us for a lone warp — measured on the producer's path right after the dependency wait).
__device__ __forceinline__ void sw_next(const AttnTcParams& p, SwUnit& t, int& u) {
  ++u;
  if (++t.kt == p.a.n_chunks) {
    t = sw_unit(p, u);
  } else {
    t.key0 += kTcKeys;
    t.hidden = t.key0 + kTcKeys <= t.lo;
  }
}
__device__ __forceinline__ void sw_skip_hidden(const AttnTcParams& p, SwUnit& t, int& u, int u1) {
  while (u < u1 && t.hidden) sw_next(p, t, u);
}

__device__ __forceinline__ int redux_max_s32(int v) {
  int r;
  asm volatile("redux.sync.max.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}

template <int NCOLS>
__global__ void __launch_bounds__(kSwThreads, 1)
attn_sw_kernel(const __grid_constant__ AttnTcMaps maps, const AttnSwParams sp) {
  constexpr int DH = 128;
  constexpr int NC = NCOLS / 4;     // columns per softmax thread (8 | 16)
  constexpr int NG = NC / 8;        // 8-column groups per softmax thread
  constexpr int NGT = NCOLS / 8;    // 8-column groups per unit
  constexpr uint32_t kQBytes = 2 * NCOLS * 128;              // Q rows of one run: two head-dim atoms of [NCOLS slots][128 B]
  const AttnTcParams& p = sp.t;
  const AttnParams& a = p.a;
  const int NV = sp.nv;
  extern __shared__ uint8_t smem_raw[];
  enum { FK = 0, EK = 2, FV = 4, EV = 8, FQ = 12, SD = 14, PR = 16, SG = 18, OD = 20, NBARS = 22 };
  __shared__ __align__(8) uint64_t bars[NBARS];
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) float xw[2][4][8];                // [unit parity][lane quarter][column group]: the warps' group maxima
  __shared__ int fresh_flag[2];                              // [unit parity] 1: this unit starts a segment (softmax -> MMA issuer)
  __shared__ float seg_m[2][8];                              // [segment parity][column group]: reference maxima of a segment awaiting its epilogue
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = base, sV = sK + 2 * kSwTileBytes, sP = sV + uint32_t(NV) * kSwTileBytes, sQ = sP + 2 * kSwPBytes,
                 sOnes = sQ + 2 * kQBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  const int u0 = sp.ub[blockIdx.x], u1 = sp.ub[blockIdx.x + 1];

  // the ones tile (A operand of the column-sum product): swizzling a constant is a no-op
  {
    uint4* ones = reinterpret_cast<uint4*>(smem_raw + (sOnes - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < int(kTcRows * 128 / 16); i += blockDim.x)
      ones[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async();
  }
  if (warp == kSwTmaWarp) {
    if (lane == 0) {
      tma_prefetch_desc(&maps.q);
      tma_prefetch_desc(&maps.k);
      tma_prefetch_desc(&maps.v);
      for (int i = 0; i < NBARS; ++i) mbar_init(bar(i), (i >= PR && i < SG) || i >= OD ? 16 : 1);
      fence_barrier_init();
    }
  } else if (warp == kSwMmaWarp) {
    tmem_alloc(smem_u32(&tmem_holder), 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_holder;   // columns: S^T 0 | 64, O^T 128 | 192, L 256 | 320

  // ---- producer helpers (warp 4, all lanes, warp-uniform arguments; the TMA instructions sit under elect.sync so that
  // their operands stay in uniform registers) ----
  auto issue_k = [&](const SwUnit& t, int n) {
    const int st = n & 1;
    const int krow = p.k_row0 + (t.b * a.Hkv + t.hkv) * a.Lmax + t.key0;
    const uint32_t dst = sK + uint32_t(st) * kSwTileBytes;
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(bar(FK + st), kSwTileBytes);
      tma_load_2d(dst, &maps.k, 0, krow, bar(FK + st), kPolicyEvictFirst);
      tma_load_2d(dst + kTcKeys * 128, &maps.k, 64, krow, bar(FK + st), kPolicyEvictFirst);
    }
    __syncwarp();
  };
  auto issue_v = [&](const SwUnit& t, int n) {
    const int st = n % NV;
    const int krow = p.k_row0 + (t.b * a.Hkv + t.hkv) * a.Lmax + t.key0;
    const uint32_t dst = sV + uint32_t(st) * kSwTileBytes;
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(bar(FV + st), kSwTileBytes);
      tma_load_2d(dst, &maps.v, 0, krow, bar(FV + st), kPolicyEvictFirst);
      tma_load_2d(dst + kTcKeys * 128, &maps.v, 64, krow, bar(FV + st), kPolicyEvictFirst);
    }
    __syncwarp();
  };
  auto issue_q = [&](int b, int h0, int heads, int r) {   // r: CTA-local run counter -> buffer r & 1
    const int qb = r & 1;
    const uint32_t dst = sQ + uint32_t(qb) * kQBytes;
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(bar(FQ + qb), uint32_t(heads) * 2u * uint32_t(p.Wp) * 128u);
      for (int hs = 0; hs < heads; ++hs) {
        tma_load_2d(dst + uint32_t(hs * p.Wp) * 128, &maps.q, (h0 + hs) * DH, b * a.W, bar(FQ + qb), kPolicyEvictLast);
        tma_load_2d(dst + NCOLS * 128 + uint32_t(hs * p.Wp) * 128, &maps.q, (h0 + hs) * DH + 64, b * a.W, bar(FQ + qb), kPolicyEvictLast);
      }
    }
    __syncwarp();
  };

  const int cl = sp.cluster;
  float fin_o[NC], fin_l[NC], fin_m[NG];   // cluster mode, softmax warps: this thread's columns of the CTA's O^T / L, their references
#pragma unroll
  for (int e = 0; e < NC; ++e) fin_o[e] = fin_l[e] = 0.f;
#pragma unroll
  for (int g = 0; g < NG; ++g) fin_m[g] = -INFINITY;
  if (sp.grid_cap < 0) {   // developer timing (SJD_DEBUG_ATTN=2|4): the cost of the kernel boundaries alone
    pdl_wait();
    pdl_launch_dependents();
    tcgen05_fence_before();
    __syncthreads();
    if (warp == kSwMmaWarp) tmem_dealloc(tmem_base, 512);
    return;
  }
  // Every role decodes what it needs BEFORE the dependency wait and then waits inside its own branch (the per-role
  // state must not be live across a common wait: registers).  Order: wait, THEN release the dependents
  // (attention_tc.cu explains why).
  auto dep_wait = [&]() {
    pdl_wait();
    pdl_launch_dependents();
  };

  if (warp == kSwTmaWarp) {
    // ===== TMA producer: two cursors (K ring, V ring), whichever has a free stage goes =====
    // Keys below kv_len were cached by earlier forwards: their tiles may stream while the previous kernel finishes
    // (q, the window's own K/V rows and the partial buffers are that kernel's until griddepcontrol.wait returns).
    // What the producer needs right after the wait — the first two runs' Q loads — is decoded before it.
    int nk = 0, nv = 0, uk = u0, uv = u0, k_runs = 0, k_last_run = -1;
    SwUnit tk = sw_unit(p, u0 < u1 ? u0 : 0), tv = tk;
    int q_have = 0, qa_b = 0, qa_h0 = 0, qa_heads = 0, qb_b = 0, qb_h0 = 0, qb_heads = 0;
    {
      SwUnit tq = tk;
      int uq = u0, last = -1;
      while (q_have < 2) {
        sw_skip_hidden(p, tq, uq, u1);
        if (uq >= u1) break;
        if (tq.run != last) {
          if (q_have == 0) { qa_b = tq.b; qa_h0 = tq.h0; qa_heads = tq.heads; }
          else { qb_b = tq.b; qb_h0 = tq.h0; qb_heads = tq.heads; }
          ++q_have;
          last = tq.run;
        }
        sw_next(p, tq, uq);
      }
    }
    sw_skip_hidden(p, tk, uk, u1);
    sw_skip_hidden(p, tv, uv, u1);
    // A range that STARTS with the tile holding this window's keys (the last tile of a head) must not keep the old
    // tiles behind it from streaming early (such CTAs ended 2.5-3 us after their peers): its ring slot 0 is filled
    // right after the wait, slots 1.. now
    const bool defer0 = uk < u1 && tk.key0 + kTcKeys > a.kv_len;
    const SwUnit t0 = tk;
    if (defer0) {
      k_runs = 1;
      k_last_run = tk.run;
      nk = 1;
      nv = 1;
      sw_next(p, tk, uk);
      sw_skip_hidden(p, tk, uk, u1);
      sw_next(p, tv, uv);
      sw_skip_hidden(p, tv, uv, u1);
    }
    while (nk < 2 && uk < u1 && tk.key0 + kTcKeys <= a.kv_len) {
      if (tk.run != k_last_run) {
        ++k_runs;
        k_last_run = tk.run;
      }
      issue_k(tk, nk);
      ++nk;
      sw_next(p, tk, uk);
      sw_skip_hidden(p, tk, uk, u1);
    }
    while (nv < NV && uv < u1 && tv.key0 + kTcKeys <= a.kv_len) {
      issue_v(tv, nv);
      ++nv;
      sw_next(p, tv, uv);
      sw_skip_hidden(p, tv, uv, u1);
    }
    dep_wait();
    // the head of the critical path: Q of the first two runs (both buffers are untouched)
    if (q_have > 0) issue_q(qa_b, qa_h0, qa_heads, 0);
    if (q_have > 1) issue_q(qb_b, qb_h0, qb_heads, 1);
    if (defer0) {
      issue_k(t0, 0);
      issue_v(t0, 0);
    }
    if (p.dbg && blockIdx.x == 0 && lane == 0) p.dbg[10] = clock64();
    while (uk < u1 || uv < u1) {
      bool did = false;
      if (uk < u1) {
        bool ok = true;
        if (nk >= 2) ok = __all_sync(0xffffffffu, mbar_test_wait(bar(EK + (nk & 1)), uint32_t((nk >> 1) - 1) & 1u)) != 0;
        if (ok) {
          // a third run's Q overwrites the buffer of run r-2: every S^T of that run precedes unit nk-2, whose commit
          // the stage test above has just seen
          if (tk.run != k_last_run) {
            ++k_runs;
            k_last_run = tk.run;
            if (k_runs > 2) issue_q(tk.b, tk.h0, tk.heads, k_runs - 1);
          }
          issue_k(tk, nk);
          if (p.dbg && blockIdx.x == 0 && lane == 0 && nk < 8) p.dbg[nk * 16 + 0] = clock64();
          ++nk;
          sw_next(p, tk, uk);
          sw_skip_hidden(p, tk, uk, u1);
          did = true;
        }
      }
      if (uv < u1) {
        bool ok = true;
        if (nv >= NV) ok = __all_sync(0xffffffffu, mbar_test_wait(bar(EV + nv % NV), uint32_t(nv / NV - 1) & 1u)) != 0;
        if (ok) {
          issue_v(tv, nv);
          ++nv;
          sw_next(p, tv, uv);
          sw_skip_hidden(p, tv, uv, u1);
          did = true;
        }
      }
      if (!did) __nanosleep(32);
    }
  } else if (warp == kSwMmaWarp) {
    // ===== MMA issuer.  Everything it does sits on the per-unit critical path.  The WHOLE warp runs the loop with
    // warp-uniform state and only the tcgen05 instructions sit under elect.sync: issued from a lane-divergent branch
    // (`if (lane == 0)`) every UTCHMMA is wrapped in a broadcast loop of ~15 instructions (0.6-1.0 us per 16-MMA product
    // in the first version's stamps).  Descriptors are built once (a k-step only adds to the start-address field) and a
    // barrier that has been seen complete is not polled again (a test_wait costs ~150 cycles) =====
    {
      const uint32_t idesc_s = umma_idesc_bf16_f32(kTcKeys, NCOLS);                 // A = K, B = Q: both K-major
      const uint32_t idesc_o = umma_idesc_bf16_f32_amn_bmn(DH, NCOLS);              // A = V (MN-major), B = P^T (MN-major)
      const uint32_t idesc_l = umma_idesc_bf16_f32_bmn(kTcRows, NCOLS);             // A = ones (K-major), B = P^T
      const uint64_t d_ones = umma_desc_sw128_kmajor(sOnes);
      const uint64_t d_k0 = umma_desc_sw128_kmajor(sK), d_q0 = umma_desc_sw128_kmajor(sQ);
      const uint64_t d_v0 = umma_desc_sw128_mnmajor(sV, kTcKeys * 128), d_p0 = umma_desc_sw128_mnmajor(sP, kTcKeys * 128);
      auto all_ok = [&](uint32_t b, uint32_t parity) { return __all_sync(0xffffffffu, mbar_test_wait(b, parity)) != 0; };
      int N = 0, mma_us = u0;       // real units of this CTA; the S cursor
      SwUnit mma_t = sw_unit(p, u0 < u1 ? u0 : 0);
      {
        SwUnit t = mma_t;
        for (int u = u0; u < u1;) {
          N += t.hidden ? 0 : 1;
          sw_next(p, t, u);
        }
      }
      sw_skip_hidden(p, mma_t, mma_us, u1);
      dep_wait();
      int ns = 0, np = 0, s_last_run = -1, s_runs = 0, seg = -1;
      bool fv_ok = false, fk_ok = false, fq_ok = false;   // sticky: landed V of unit np / K of unit ns / Q of unit ns' run
      int s_new_run = -1, s_r = 0;                        // decoded once per S unit
      while (np < N) {
        bool did = false;
        if (np < ns) {
          const int ts = np & 1, vst = np % NV;
          if (!fv_ok) fv_ok = all_ok(bar(FV + vst), uint32_t(np / NV) & 1u);
          if (fv_ok && all_ok(bar(PR + ts), uint32_t(np >> 1) & 1u)) {
            const bool fresh = __any_sync(0xffffffffu, *reinterpret_cast<volatile int*>(&fresh_flag[ts]) != 0) != 0;
            if (fresh) {
              if (seg >= 0 && elect_one_sync()) umma_commit(bar(SG + (seg & 1)));   // the segment that just ended: its products are all issued
              __syncwarp();
              ++seg;
              if (seg >= 2) mbar_wait_backoff(bar(OD + (seg & 1)), uint32_t((seg >> 1) - 1) & 1u);   // accumulators drained
            }
            tcgen05_fence_after();
            if (p.dbg && blockIdx.x == 0 && lane == 0 && np < 8) p.dbg[np * 16 + 3] = clock64();
            const uint32_t tO = tmem_base + 128 + uint32_t(seg & 1) * 64, tL = tmem_base + 256 + uint32_t(seg & 1) * 64;
            const uint64_t dv0 = d_v0 + uint64_t(uint32_t(vst) * (kSwTileBytes >> 4));
            const uint64_t dp0 = d_p0 + uint64_t(uint32_t(ts) * (kSwPBytes >> 4));
            ++np;
            if (elect_one_sync()) {
              uint32_t acc = fresh ? 0u : 1u;
#pragma unroll
              for (int k = 0; k < kTcKeys / 16; ++k) {   // 16 keys per step = two 8-row groups = 2 048 bytes of either tile
                umma_bf16_ss(tO, dv0 + uint64_t(k * 128), dp0 + uint64_t(k * 128), idesc_o, acc);
                umma_bf16_ss(tL, d_ones, dp0 + uint64_t(k * 128), idesc_l, acc);
                acc = 1;
              }
              umma_commit(bar(EV + vst));
              if (np == N) umma_commit(bar(SG + (seg & 1)));
            }
            __syncwarp();
            fv_ok = false;
            if (p.dbg && blockIdx.x == 0 && lane == 0 && np <= 8) p.dbg[(np - 1) * 16 + 4] = clock64();
            did = true;
          }
        }
        if (ns < N && ns - np < 2) {
          if (s_new_run < 0) {
            s_new_run = mma_t.run != s_last_run ? 1 : 0;
            s_last_run = mma_t.run;
            if (s_new_run) ++s_runs;
            s_r = s_runs - 1;
            fq_ok = !s_new_run;
          }
          const int qb = s_r & 1;
          if (!fk_ok) fk_ok = all_ok(bar(FK + (ns & 1)), uint32_t(ns >> 1) & 1u);
          if (fk_ok && !fq_ok) fq_ok = all_ok(bar(FQ + qb), uint32_t(s_r >> 1) & 1u);
          if (fk_ok && fq_ok) {
            if (p.dbg && blockIdx.x == 0 && lane == 0 && ns < 8) p.dbg[ns * 16 + 1] = clock64();
            tcgen05_fence_after();
            const uint32_t tS = tmem_base + uint32_t(ns & 1) * 64;
            const uint64_t dk = d_k0 + uint64_t(uint32_t(ns & 1) * (kSwTileBytes >> 4));
            const uint64_t dq = d_q0 + uint64_t(uint32_t(qb) * (kQBytes >> 4));
            if (elect_one_sync()) {
              uint32_t acc = 0;
#pragma unroll
              for (int d = 0; d < 2; ++d) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_bf16_ss(tS, dk + uint64_t(d * (kTcKeys * 128 >> 4) + 2 * k), dq + uint64_t(d * (NCOLS * 128 >> 4) + 2 * k), idesc_s, acc);
                  acc = 1;
                }
              }
              umma_commit(bar(EK + (ns & 1)));     // the K tile is free
              umma_commit(bar(SD + (ns & 1)));     // S^T is there
            }
            __syncwarp();
            if (p.dbg && blockIdx.x == 0 && lane == 0 && ns < 8) p.dbg[ns * 16 + 2] = clock64();
            ++ns;
            sw_next(p, mma_t, mma_us);
            sw_skip_hidden(p, mma_t, mma_us, u1);
            fk_ok = false;
            s_new_run = -1;
            did = true;
          }
        }
        if (!did) __nanosleep(20);
      }
    }
  } else if (warp < 16) {
    // ===== softmax + epilogue: 16 warps.  Thread = TMEM lane = KEY of the tile (softmax) / head-dim element
    // (epilogue); the four warps of a lane quarter take a quarter of the columns each.  These warps run one dependent
    // chain per unit (a lone warp issues an instruction every ~5 cycles), so the common case — a tile every query of
    // the window sees completely, no padded columns — takes a path without mask arithmetic =====
    const int wq = warp & 3, cq = warp >> 2;
    const int kl = wq * 32 + lane;
    const uint32_t t_row = uint32_t(wq * 32) << 16;
    const float sc = a.scale_log2e;
    const int cbase = cq * NC;                                        // first column of this thread
    const int hs_start = tc_div(cbase, p.m_wp, p.Wp), qi_start = cbase - hs_start * p.Wp;
    constexpr uint32_t kFull = (1u << NC) - 1u;
    const bool dense = p.Wp == a.W;                                   // no padded rows between the heads of a unit
    float mref[NGT];                                                  // the running segment's reference maxima (scaled log2 domain)
#pragma unroll
    for (int g = 0; g < NGT; ++g) mref[g] = -INFINITY;
    int seg = -1, last_run = -1, n = 0;
    int s_R = 0, s_h0 = 0;                                            // the running segment's geometry
    size_t s_base = 0;
    uint32_t colmask = 0;                                             // bit e: column cbase + e is a real query row of the run
    // drains segment `e_seg` (accumulators of buffer e_seg & 1; reference maxima in seg_m[e_seg & 1]) into its partial slot
    auto epilogue = [&](int e_seg, int e_R, int e_h0, size_t e_base) __attribute__((always_inline)) {
      const int ob = e_seg & 1;
      mbar_wait(bar(SG + ob), uint32_t(e_seg >> 1) & 1u);
      tcgen05_fence_after();
      if (p.dbg && blockIdx.x == 0 && warp == 0 && lane == 0 && e_seg < 8) p.dbg[e_seg * 16 + 8] = clock64();
      const uint32_t tO = tmem_base + 128 + uint32_t(ob) * 64, tL = tmem_base + 256 + uint32_t(ob) * 64;
      if (cbase < e_R) {   // warp-uniform
        uint32_t v[NC];
#pragma unroll
        for (int c0 = 0; c0 < NC; c0 += 8) tmem_ld_32x32b_x8(tO + t_row + uint32_t(cbase + c0), *reinterpret_cast<uint32_t(*)[8]>(&v[c0]));
        tmem_ld_wait();
        if (dense && cbase + NC <= e_R) {
          // rows of consecutive heads are consecutive: column c -> partial row e_h0 * W + c; a warp writes 128 contiguous bytes
          float* pcol = a.part_o + (e_base + size_t(e_h0) * a.W + cbase) * DH + kl;
#pragma unroll
          for (int e = 0; e < NC; ++e) pcol[e * DH] = __uint_as_float(v[e]);
        } else {
          // the next head starts W rows later: one pointer walked with adds
          float* pcol = a.part_o + (e_base + size_t(e_h0 + hs_start) * a.W + qi_start) * DH + kl;
          int qi = qi_start;
#pragma unroll
          for (int e = 0; e < NC; ++e) {
            if (qi < a.W && cbase + e < e_R) *pcol = __uint_as_float(v[e]);
            pcol += DH;
            if (++qi == p.Wp) {
              qi = 0;
              pcol -= (p.Wp - a.W) * DH;
            }
          }
        }
      }
      if (warp == 0) {   // {m, sum p}: every lane of L holds the sums; lane (c & 15) of the matching half-warp writes column c
#pragma unroll 1
        for (int c0 = 0; c0 < e_R; c0 += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(tL + uint32_t(c0), v);
          tmem_ld_wait();
          const int c = c0 + (lane & 15);
          if ((lane >> 4) == ((c0 >> 4) & 1) && c < e_R) {
            float l = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if ((lane & 15) == e) l = __uint_as_float(v[e]);
            const float m = seg_m[ob][c >> 3];
            const int hs2 = tc_div(c, p.m_wp, p.Wp), qi2 = c - hs2 * p.Wp;
            if (qi2 < a.W) {
              const size_t pr = e_base + size_t(e_h0 + hs2) * a.W + qi2;
              *reinterpret_cast<float2*>(a.part_ml + pr * 2) = make_float2(m, l);
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(OD + ob));
      if (p.dbg && blockIdx.x == 0 && warp == 0 && lane == 0 && e_seg < 8) p.dbg[e_seg * 16 + 9] = clock64();
    };
    auto mark_empty = [&](const SwUnit& t, size_t ubase) {   // slot (key tile, rows of this unit) contributes nothing
      if (warp == 0) {
        for (int c = lane; c < t.R; c += 32) {
          const int hs2 = tc_div(c, p.m_wp, p.Wp), qi2 = c - hs2 * p.Wp;
          if (qi2 < a.W)
            *reinterpret_cast<float2*>(a.part_ml + (ubase + size_t(t.h0 + hs2) * a.W + qi2) * 2) = make_float2(-INFINITY, 0.f);
        }
      }
    };

    SwUnit sm_t = sw_unit(p, u0 < u1 ? u0 : 0);
    dep_wait();
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[15] = clock64();
    if (p.dbg && threadIdx.x == 0) {   // per-CTA wall-clock (ns): [128 + 4 cta] = dependency wait returned, [+1] = end of work, [+2] = units
      long long tns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
      p.dbg[128 + 4 * blockIdx.x] = tns;
      p.dbg[128 + 4 * blockIdx.x + 2] = u1 - u0;
      p.dbg[128 + 4 * blockIdx.x + 3] = u0;
    }
    for (int u = u0; u < u1; ++u) {
      const SwUnit t = sm_t;
      if (u + 1 < u1) {
        int uu = u;
        sw_next(p, sm_t, uu);
      }
      const size_t ubase = (size_t(t.kt) * a.rows + t.b) * a.H * size_t(a.W);
      if (t.hidden) {   // uniform per CTA
        if (!cl) mark_empty(t, ubase);
        continue;
      }
      const bool new_run = t.run != last_run;
      last_run = t.run;
      if (new_run) {   // which of this thread's columns are query rows (R can shrink on the last row tile of a kv head)
        colmask = 0;
        int qi = qi_start;
#pragma unroll
        for (int e = 0; e < NC; ++e) {
          colmask |= uint32_t(qi < a.W && cbase + e < t.R) << e;
          if (++qi == p.Wp) qi = 0;
        }
      }
      const int ts = n & 1, j = n >> 1;
      const uint32_t tS = tmem_base + uint32_t(ts) * 64;
      mbar_wait(bar(SD + ts), uint32_t(j) & 1u);
      tcgen05_fence_after();
      if (p.dbg && blockIdx.x == 0 && warp == 0 && lane == 0 && n < 8) p.dbg[n * 16 + 5] = clock64();
      uint32_t v[NC];
#pragma unroll
      for (int c0 = 0; c0 < NC; c0 += 8) tmem_ld_32x32b_x8(tS + t_row + uint32_t(cbase + c0), *reinterpret_cast<uint32_t(*)[8]>(&v[c0]));
      tmem_ld_wait();
      // visibility of (this key, column e): bit e of okm.  Interior tile: every key is visible to every query row
      const bool interior = (t.key0 >= t.lo) && (t.key0 + kTcKeys - 1 <= a.kv_len);
      uint32_t okm = colmask;
      if (!interior) {
        const int jk = t.key0 + kl, dk = jk - a.kv_len;             // visible to query i iff lo <= jk and dk <= i
        uint32_t vis = 0;
        int qi = qi_start;
#pragma unroll
        for (int e = 0; e < NC; ++e) {
          vis |= uint32_t(dk <= qi) << e;
          if (++qi == p.Wp) qi = 0;
        }
        okm = jk >= t.lo ? (okm & vis) : 0u;
      }
      float gm[NG];
      if (okm == 0u) {   // none of this thread's (key, column) pairs is visible (padded columns of a narrow window, masked keys)
#pragma unroll
        for (int g = 0; g < NG; ++g) gm[g] = -INFINITY;
      } else if (okm == kFull) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const float m0 = fmaxf(fmaxf(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1])),
                                 fmaxf(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])));
          const float m1 = fmaxf(fmaxf(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])),
                                 fmaxf(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])));
          gm[g] = fmaxf(m0, m1);
        }
      } else {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          float m = -INFINITY;
#pragma unroll
          for (int e = 0; e < 8; ++e) m = fmaxf(m, (okm >> (8 * g + e)) & 1u ? __uint_as_float(v[8 * g + e]) : -INFINITY);
          gm[g] = m;
        }
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) gm[g] = ord2f(redux_max_s32(f2ord(gm[g]))) * sc;   // scale > 0: max commutes with it
      if (lane < NG) xw[ts][wq][cq * NG + lane] = (NG == 2 && lane == 1) ? gm[NG - 1] : gm[0];
      asm volatile("bar.sync 1, 512;" ::: "memory");
      float tmax[NGT];
#pragma unroll
      for (int g4 = 0; g4 < NGT / 4; ++g4) {
        float4 x = *reinterpret_cast<const float4*>(&xw[ts][0][4 * g4]);
#pragma unroll
        for (int q = 1; q < 4; ++q) {
          const float4 y = *reinterpret_cast<const float4*>(&xw[ts][q][4 * g4]);
          x.x = fmaxf(x.x, y.x); x.y = fmaxf(x.y, y.y); x.z = fmaxf(x.z, y.z); x.w = fmaxf(x.w, y.w);
        }
        // INTEGER references (log2 domain): probabilities scaled against two different integers differ by an exact
        // power of two, so their bf16 roundings, the fp32 products and the merge weights are the same numbers up to the
        // exponent — the result does not depend on which 8-column group, key tile or segment supplied the reference,
        // i.e. not on the window a token happens to share (tests: ..._logits_do_not_depend_on_the_window)
        tmax[4 * g4] = ceilf(x.x); tmax[4 * g4 + 1] = ceilf(x.y); tmax[4 * g4 + 2] = ceilf(x.z); tmax[4 * g4 + 3] = ceilf(x.w);
      }
      if (p.dbg && blockIdx.x == 0 && warp == 0 && lane == 0 && n < 8) p.dbg[n * 16 + 6] = clock64();
      // segment decision: identical in every thread (same inputs)
      bool grew = false;
#pragma unroll
      for (int g = 0; g < NGT; ++g) grew = grew || (tmax[g] > mref[g] + sp.grow);
      bool brk = cl ? seg < 0 : (new_run || grew);
      if (cl && seg >= 0 && grew) {
        // cluster mode keeps ONE accumulator per CTA: a tile that outgrows the reference rescales O^T and L in place.
        // Every product issued so far has completed once the commit behind PV(n-1) has arrived (its V stage's barrier).
        mbar_wait(bar(EV + (n - 1) % NV), uint32_t((n - 1) / NV) & 1u);
        tcgen05_fence_after();
        float nm[NGT];
#pragma unroll
        for (int g = 0; g < NGT; ++g) nm[g] = fmaxf(mref[g], tmax[g]);
        const uint32_t tO = tmem_base + 128 + t_row + uint32_t(cbase), tL = tmem_base + 256 + t_row + uint32_t(cbase);
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          float f = 1.f;
#pragma unroll
          for (int x = 0; x < 4; ++x)
            if (cq == x) f = nm[(x * NG + g) % NGT] == -INFINITY ? 1.f : exp2f(mref[(x * NG + g) % NGT] - nm[(x * NG + g) % NGT]);
          uint32_t vo[8], vl[8];
          tmem_ld_32x32b_x8(tO + uint32_t(8 * g), vo);
          tmem_ld_32x32b_x8(tL + uint32_t(8 * g), vl);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            vo[e] = __float_as_uint(__uint_as_float(vo[e]) * f);
            vl[e] = __float_as_uint(__uint_as_float(vl[e]) * f);
          }
          tmem_st_32x32b_x8(tO + uint32_t(8 * g), vo);
          tmem_st_32x32b_x8(tL + uint32_t(8 * g), vl);
        }
        tmem_st_wait();
#pragma unroll
        for (int g = 0; g < NGT; ++g) mref[g] = nm[g];
      }
      const int e_seg = seg, e_R = s_R, e_h0 = s_h0;
      const size_t e_base = s_base;
      if (brk) {
        if (threadIdx.x == 0 && seg >= 0) {   // the ended segment's maxima, for its epilogue (read by this warp only)
#pragma unroll
          for (int g = 0; g < NGT; ++g) seg_m[seg & 1][g] = mref[g];
        }
#pragma unroll
        for (int g = 0; g < NGT; ++g) mref[g] = tmax[g];
        ++seg;
        s_R = t.R; s_h0 = t.h0; s_base = ubase;
      } else if (!cl) {
        mark_empty(t, ubase);
      }
      // probabilities of this key for its columns (bf16, q contiguous) -> 16-byte chunks of its P^T row
      {
        uint8_t* const rowp = smem_raw + (sP + uint32_t(ts) * kSwPBytes - smem_u32(smem_raw)) + kl * 128;
        uint32_t pk[NC / 2];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          float ms = -INFINITY;
#pragma unroll
          for (int x = 0; x < 4; ++x)
            if (cq == x) ms = mref[(x * NG + g) % NGT];
          if (ms == -INFINITY) ms = 0.f;
          if (okm == 0u) {
#pragma unroll
            for (int e = 0; e < 8; e += 2) pk[(8 * g + e) >> 1] = 0u;
          } else if (okm == kFull) {
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              const int i = 8 * g + e;
              const __nv_bfloat162 pb = __floats2bfloat162_rn(ex2_approx(fmaf(__uint_as_float(v[i]), sc, -ms)),
                                                              ex2_approx(fmaf(__uint_as_float(v[i + 1]), sc, -ms)));
              pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&pb);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              const int i = 8 * g + e;
              const float p0 = (okm >> i) & 1u ? ex2_approx(fmaf(__uint_as_float(v[i]), sc, -ms)) : 0.f;
              const float p1 = (okm >> (i + 1)) & 1u ? ex2_approx(fmaf(__uint_as_float(v[i + 1]), sc, -ms)) : 0.f;
              const __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
              pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&pb);
            }
          }
        }
#pragma unroll
        for (int g = 0; g < NG; ++g)
          *reinterpret_cast<uint4*>(rowp + (((cq * NG + g) ^ (kl & 7)) << 4)) =
              make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
      }
      tcgen05_fence_before();
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
      if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(&fresh_flag[ts]) = brk ? 1 : 0;
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(PR + ts));
      if (p.dbg && blockIdx.x == 0 && warp == 0 && lane == 0 && n < 8) p.dbg[n * 16 + 7] = clock64();
      if (brk && e_seg >= 0) epilogue(e_seg, e_R, e_h0, e_base);
      ++n;
    }
    if (cl) {
      // this thread's columns of the CTA's accumulators (lane = head-dim element; every lane of L holds the sums)
      if (seg >= 0) {
        mbar_wait(bar(SG), 0u);
        tcgen05_fence_after();
        uint32_t vo[NC], vl[NC];
#pragma unroll
        for (int c0 = 0; c0 < NC; c0 += 8) {
          tmem_ld_32x32b_x8(tmem_base + 128 + t_row + uint32_t(cbase + c0), *reinterpret_cast<uint32_t(*)[8]>(&vo[c0]));
          tmem_ld_32x32b_x8(tmem_base + 256 + t_row + uint32_t(cbase + c0), *reinterpret_cast<uint32_t(*)[8]>(&vl[c0]));
        }
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < NC; ++e) {
          fin_o[e] = __uint_as_float(vo[e]);
          fin_l[e] = __uint_as_float(vl[e]);
        }
        tcgen05_fence_before();
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) {
#pragma unroll
        for (int x = 0; x < 4; ++x)
          if (cq == x) fin_m[g] = mref[(x * NG + g) % NGT];
      }
    } else if (seg >= 0) {
      if (threadIdx.x == 0) {
#pragma unroll
        for (int g = 0; g < NGT; ++g) seg_m[seg & 1][g] = mref[g];
      }
      __syncwarp();
      epilogue(seg, s_R, s_h0, s_base);
    }
  }
  if (cl) {
    // ===== cluster merge: peers stage {O^T [col][d], L [col], m [group]} in their own K ring (free by now), one cluster
    // barrier, the leader reads them through distributed shared memory, merges in rank order (fixed: reproducible) and
    // writes the normalised bf16 rows; a second barrier keeps the peers' shared memory alive until it has =====
    const uint32_t rank = cl > 1 ? cluster_ctarank() : 0u;
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[16 + 10] = clock64();   // (developer stamps: scripts/attn_sw_stamps.py)
    // staging layout: O^T as [d][NCOLS + 4] (a thread's columns are contiguous: 16-byte stores here, 16-byte distributed
    // shared memory loads in the leader), then L [NCOLS], then m [NGT]
    constexpr int kPitch = NCOLS + 4;
    float* const stO = reinterpret_cast<float*>(smem_raw + (sK - smem_u32(smem_raw)));   // [128][kPitch]
    float* const stL = stO + 128 * kPitch;                                                // [NCOLS]
    float* const stM = stL + NCOLS;                                                       // [NGT]
    const int wq = warp & 3, cq = warp >> 2, kl = wq * 32 + lane, cbase = cq * NC;
    const SwUnit tr = sw_unit(p, u0 < u1 ? u0 : 0);   // this CTA's run (decoded before the barrier: off the merge's path)
    if (cl > 1) {
      if (warp < 16 && rank != 0) {
#pragma unroll
        for (int e = 0; e < NC; e += 4)
          *reinterpret_cast<float4*>(stO + kl * kPitch + cbase + e) = make_float4(fin_o[e], fin_o[e + 1], fin_o[e + 2], fin_o[e + 3]);
        if (kl == 0) {
#pragma unroll
          for (int e = 0; e < NC; e += 4)
            *reinterpret_cast<float4*>(stL + cbase + e) = make_float4(fin_l[e], fin_l[e + 1], fin_l[e + 2], fin_l[e + 3]);
#pragma unroll
          for (int g = 0; g < NG; ++g) stM[cq * NG + g] = fin_m[g];
        }
      }
      cluster_sync_all();
    }
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[16 + 11] = clock64();
    const bool merger = warp < 16 && rank == 0 && u0 < u1;
    // second barrier, split: everybody except the leader's merging warps arrives at once; those arrive when their
    // distributed-shared-memory loads have returned; all wait at the very end (the peers' shared memory stays alive)
    if (cl > 1 && !merger) cluster_arrive_relaxed();
    if (merger) {
      // online merge in rank order (own accumulators first).  The references are integers in the log2 domain, so every
      // rescale is an exact power of two and the result equals attn_combine_row's batch form bit for bit; one DSMEM round
      // trip per peer (m, O^T and L requested together)
      float M[NG], num[NC], den[NC];
#pragma unroll
      for (int g = 0; g < NG; ++g) M[g] = fin_m[g];
#pragma unroll
      for (int e = 0; e < NC; ++e) {
        num[e] = fin_m[e >> 3] == -INFINITY ? 0.f : fin_o[e];
        den[e] = fin_m[e >> 3] == -INFINITY ? 0.f : fin_l[e];
      }
#pragma unroll 1
      for (int r = 1; r < cl; ++r) {
        const uint32_t bm = dsmem_addr(smem_u32(stM), uint32_t(r)), bo = dsmem_addr(smem_u32(stO), uint32_t(r)),
                       bl = dsmem_addr(smem_u32(stL), uint32_t(r));
        float mr[NG];
        float4 po[NC / 4], pl[NC / 4];
#pragma unroll
        for (int g = 0; g < NG; ++g) mr[g] = ld_dsmem_f32(bm + uint32_t(cq * NG + g) * 4u);
#pragma unroll
        for (int e4 = 0; e4 < NC / 4; ++e4) {
          po[e4] = ld_dsmem_v4(bo + uint32_t(kl * kPitch + cbase + 4 * e4) * 4u);
          pl[e4] = ld_dsmem_v4(bl + uint32_t(cbase + 4 * e4) * 4u);
        }
        float so[NG], sr[NG];   // scale of what has been merged so far / of peer r
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const float mn = fmaxf(M[g], mr[g]);
          so[g] = M[g] == -INFINITY ? 0.f : exp2f(M[g] - mn);
          sr[g] = mr[g] == -INFINITY ? 0.f : exp2f(mr[g] - mn);
          M[g] = mn;
        }
#pragma unroll
        for (int e4 = 0; e4 < NC / 4; ++e4) {
          const float a0 = so[(4 * e4) >> 3], a1 = sr[(4 * e4) >> 3];
          num[4 * e4] = num[4 * e4] * a0 + po[e4].x * a1; num[4 * e4 + 1] = num[4 * e4 + 1] * a0 + po[e4].y * a1;
          num[4 * e4 + 2] = num[4 * e4 + 2] * a0 + po[e4].z * a1; num[4 * e4 + 3] = num[4 * e4 + 3] * a0 + po[e4].w * a1;
          den[4 * e4] = den[4 * e4] * a0 + pl[e4].x * a1; den[4 * e4 + 1] = den[4 * e4 + 1] * a0 + pl[e4].y * a1;
          den[4 * e4 + 2] = den[4 * e4 + 2] * a0 + pl[e4].z * a1; den[4 * e4 + 3] = den[4 * e4 + 3] * a0 + pl[e4].w * a1;
        }
      }
      if (cl > 1) cluster_arrive_relaxed();   // the peers' shared memory has been read (the arithmetic above consumed every load)
      if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[16 + 12] = clock64();
      // same final arithmetic as attn_combine_row (reciprocal, then multiply): a token's attention row must not depend on
      // whether its window took this path or the partial-slot path
      if (tr.heads == 1 && p.Wp == a.W) {
        // one head, no padded rows: column c is query row c of head h0, a constant stride apart
        __nv_bfloat16* o = a.out + (size_t(tr.b * a.W + cbase) * a.H + tr.h0) * DH + kl;
        const size_t stride = size_t(a.H) * DH;
#pragma unroll
        for (int e = 0; e < NC; ++e) {
          const float inv = den[e] > 0.f ? 1.f / den[e] : 0.f;             // fully masked query (CFG hidden prefix) -> 0
          if (cbase + e < tr.R) o[e * stride] = __float2bfloat16_rn(num[e] * inv);
        }
      } else {
        int hs = tc_div(cbase, p.m_wp, p.Wp), qi = cbase - hs * p.Wp;
#pragma unroll
        for (int e = 0; e < NC; ++e) {
          const float inv = den[e] > 0.f ? 1.f / den[e] : 0.f;
          if (qi < a.W && cbase + e < tr.R)
            a.out[(size_t(tr.b * a.W + qi) * a.H + tr.h0 + hs) * DH + kl] = __float2bfloat16_rn(num[e] * inv);
          if (++qi == p.Wp) {
            qi = 0;
            ++hs;
          }
        }
      }
    }
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[16 + 13] = clock64();
    if (cl > 1) cluster_wait();
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[16 + 14] = clock64();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) {
    long long tns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
    p.dbg[128 + 4 * blockIdx.x + 1] = tns;
  }
  if (warp == kSwMmaWarp) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Geometry: up to 64 / Wp heads of a kv head per unit; 32 accumulator columns when they suffice.
void attn_sw_plan(AttnSwParams* sp, int force_ncols, int max_cluster, int sm_count = 0) {
  attn_tct_plan(&sp->t);
  const int R = sp->t.hpc * sp->t.Wp;
  sp->ncols = (R <= 32 && force_ncols != 64) ? 32 : 64;
  sp->nv = sp->ncols == 32 ? 3 : 2;
  // cluster mode: as many CTAs per run (1, 2 or 4) as fit one wave and have a key tile each
  const AttnParams& a = sp->t.a;
  if (sm_count <= 0) sm_count = device_num_sms();
  const int runs = a.Hkv * sp->t.mtiles * a.rows, sms = sm_count < kSwMaxGrid ? sm_count : kSwMaxGrid;
  int k = 0;
  // (measured, profiles/r02ba_attn_sw_cluster_tail.txt, r02bc_cluster_minr.txt: 17.1 -> 15.5 us per layer at window 32,
  // 24.3 -> 21.4 at 64, 19.0 -> 18.3 at 16, 19.2 -> 18.5 at 8; the first form of the cluster tail lost at window 16)
  static const int min_r = getenv("SJD_ATTN_SW_CLUSTER_MINR") ? atoi(getenv("SJD_ATTN_SW_CLUSTER_MINR")) : 8;
  if (max_cluster > 0 && runs <= sms && R >= min_r) {
    k = 1;
    while (k * 2 <= max_cluster && k * 2 <= 4 && runs * k * 2 <= sms && k * 2 <= a.n_chunks) k *= 2;
  }
  sp->cluster = k;
}

// Contiguous split of the units over `grid` CTAs that balances the modelled time  (real units) + run_cost * (head
// changes inside the range)  instead of the unit count: smallest common deadline by bisection over a greedy sweep.
// Hidden tiles (whole tile inside a row's hidden prefix) cost nothing.
void attn_sw_split(AttnSwParams* sp, int grid, float run_cost) {
  const AttnParams& a = sp->t.a;
  const int nc = a.n_chunks, n_runs = a.Hkv * sp->t.mtiles * a.rows, n_units = nc * n_runs;
  auto hidden = [&](int u) {
    const int run = u / nc, kt = u - run * nc, b = run / sp->t.ny;
    return (kt + 1) * kTcKeys <= a.kv_lo[b];
  };
  auto sweep = [&](float deadline, bool write) {
    int u = 0;
    for (int c = 0; c < grid; ++c) {
      if (write) sp->ub[c] = uint16_t(u);
      float cost = 0.f;
      int last_run = -1;
      while (u < n_units) {
        const int run = u / nc;
        float add = hidden(u) ? 0.f : 1.f;
        if (add > 0.f && last_run >= 0 && run != last_run) add += run_cost;
        if (cost + add > deadline && cost > 0.f) break;
        cost += add;
        if (add > 0.f) last_run = run;
        ++u;
      }
    }
    if (write) sp->ub[grid] = uint16_t(n_units);
    return u >= n_units;
  };
  float lo = 0.f, hi = float(n_units) + run_cost * n_runs + 1.f;
  for (int it = 0; it < 30; ++it) {
    const float mid = 0.5f * (lo + hi);
    if (sweep(mid, false)) hi = mid;
    else lo = mid;
  }
  sweep(hi, true);
}

constexpr int attn_sw_smem(int ncols, int nv) {   // K ring, V ring, two P^T buffers, two Q buffers, the ones tile
  return 1024 + 2 * int(kSwTileBytes) + nv * int(kSwTileBytes) + 2 * int(kSwPBytes) + 2 * (2 * ncols * 128) + kTcRows * 128;
}

// Grid size and the unit table sp.ub for the mode sp.cluster says (0: cost-balanced contiguous split; k >= 1: CTA r * k + j
// walks slice j of run r's key tiles).  Pure host arithmetic (tests/test_host_cpu.py checks it through sjd_debug_attn_sw_split).
int attn_sw_grid(AttnSwParams& sp, int sms) {
  const AttnTcParams& p = sp.t;
  const AttnParams& a = p.a;
  const int n_units = a.n_chunks * a.Hkv * p.mtiles * a.rows;
  int ng = n_units < sms ? n_units : sms;
  if (ng > kSwMaxGrid) ng = kSwMaxGrid;
  if (sp.grid_cap > 0) {
    sp.cluster = 0;
    if (sp.grid_cap < ng) ng = sp.grid_cap;
  }
  if (sp.cluster > 0) {
    const int k = sp.cluster, runs = a.Hkv * p.mtiles * a.rows, nc = a.n_chunks;
    ng = runs * k;
    if (sp.ub[ng] != uint16_t(n_units) || sp.ub[1] != uint16_t(nc / k))
      for (int c = 0; c <= ng; ++c) sp.ub[c] = uint16_t((c / k) * nc + ((c % k) * nc) / k);
  } else if (sp.ub[ng] != uint16_t(n_units)) {   // (the same split serves every layer of a forward)
    static const float run_cost = getenv("SJD_ATTN_SW_RUNCOST") ? float(atof(getenv("SJD_ATTN_SW_RUNCOST"))) : 0.5f;
    attn_sw_split(&sp, ng, run_cost);
  }
  return ng;
}

int attn_sw_launch(const AttnTcMaps& maps, AttnSwParams& sp, cudaStream_t stream) {
  const AttnTcParams& p = sp.t;
  const AttnParams& a = p.a;
  if (p.hpc * p.Wp > 64 || p.head_dim != 128) return -3;
  const int n_units = a.n_chunks * a.Hkv * p.mtiles * a.rows;
  if (n_units >= 65536) return -3;   // tc_div's exact range
  static bool cluster_refused = false;   // a cluster launch failed once on this device (e.g. a partitioned GPU): stay in the segment form
  if (cluster_refused) sp.cluster = 0;
  dim3 grid(attn_sw_grid(sp, device_num_sms()));
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(attn_sw_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_sw_smem(32, 3)) != cudaSuccess ||
        cudaFuncSetAttribute(attn_sw_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_sw_smem(64, 2)) != cudaSuccess)
      return -5;
    set = true;
  }
  const int k = sp.cluster > 1 ? sp.cluster : 1;
  const int rc = sp.ncols == 32
      ? launch_pdl_cluster(attn_sw_kernel<32>, grid, dim3(kSwThreads), attn_sw_smem(32, sp.nv), stream, k, maps, sp)
      : launch_pdl_cluster(attn_sw_kernel<64>, grid, dim3(kSwThreads), attn_sw_smem(64, sp.nv), stream, k, maps, sp);
  if (rc && k > 1 && !cluster_refused) {
    // the cluster could not be placed: fall back to the segment form (fp32 partial slots + merge pre-op in the next chain
    // kernel — the caller looks at sp.cluster after this call) instead of failing the forward
    cudaGetLastError();
    cluster_refused = true;
    sp.cluster = 0;
    sp.ub[0] = sp.ub[1] = 0;
    for (int c = 2; c < 150; ++c) sp.ub[c] = 0;
    return attn_sw_launch(maps, sp, stream);
  }
  return rc;
}

}  // namespace sjd
