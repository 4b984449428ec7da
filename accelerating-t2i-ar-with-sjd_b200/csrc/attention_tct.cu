// Small-window variant of the tcgen05 attention (attention_tc.cu): the products are TRANSPOSED so that the 128 TMEM
// lanes carry the 128 KEYS of the tile and the (few) query rows are the accumulator's columns.
//
//   S^T[128 keys x 64 q] = K Q^T      A = K tile (K-major, as TMA lands it), B = Q rows (K-major, N = 64 slots)
//   O^T[128 d   x 64 q] = V^T P^T     A = V tile as TMA lands it ([key][d]: M = d contiguous -> MN-major A),
//                                     B = P^T [key][q] written by the softmax threads (q contiguous -> MN-major B)
//   L  [128     x 64 q] = 1  P^T      row sums of P for free: A = a constant tile of ones, same B
//
// With 32 query rows (Lumina / Chameleon, window 32) attention_tc.cu fills only a quarter of its 128 accumulator
// lanes with real rows, loads Q four times to keep its softmax threads busy and does 4x redundant tensor work; here
// every lane is a real key, Q is loaded once (8 KB), the first product is 128x64 instead of 128x128, the epilogue's
// stores are coalesced without a transpose (lanes = consecutive head-dim elements of one output row), and the row
// sums come out of the tensor core.  The price: a softmax row now runs ACROSS threads, so the column max is taken
// with redux.sync per warp and met across the four warps in shared memory.
// Same pipeline, partial format and masking contract as attention_tc.cu.  Head dim 128 only (UMMA M = d).
#include "common.cuh"

namespace sjd {

constexpr int kTctCols = 64;   // query-row slots per unit (UMMA N)

__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32_amn_bmn(uint32_t M, uint32_t N) {
  return umma_idesc_bf16_f32(M, N) | (1u << 15) | (1u << 16);
}
// order-preserving float <-> int (so that redux.sync.max.s32 is a float max)
__device__ __forceinline__ int f2ord(float f) {
  const int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

__global__ void __launch_bounds__(kTcThreads2, 1)
attn_tct_kernel(const __grid_constant__ AttnTcMaps maps, const AttnTcParams p) {
  constexpr int DH = 128;
  constexpr uint32_t kKBytes = 2 * kTcKeys * 128;          // K tile: two head-dim atoms of [128 keys][128 B]; later P^T
  constexpr uint32_t kVBytes = 2 * kTcKeys * 128;          // V tile: two boxes of [128 keys][128 B]
  constexpr uint32_t kQBytes = 2 * kTctCols * 128;         // Q rows: two atoms of [64 slots][128 B]
  constexpr uint32_t kStage = kKBytes + kVBytes + kQBytes; // 80 KB
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[12];               // per stage: qk landed, v landed, S done, P ready, O done, O drained
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) float xmf_all[2 * kTctCols];                  // [group][column] tile-wide column max (scaled log2 domain)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sOnes = base + 2 * kStage;
  // per softmax group: a [128 keys][33] fp32 staging tile (+8 floats per 32 keys) for the column maxima
  constexpr uint32_t kStBytes = (kTcKeys * 33 + 32) * 4;
  const uint32_t sSt = sOnes + kTcRows * 128;
  const AttnParams& a = p.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int which, int s) { return smem_u32(&bars[which * 2 + s]); };
  enum { B_QK = 0, B_V = 1, B_S = 2, B_P = 3, B_O = 4, B_E = 5 };
  const int G = a.H / a.Hkv;
  const int n_units = a.n_chunks * a.Hkv * ((G + p.hpc - 1) / p.hpc) * a.rows;
  const int T = a.kv_len + a.W;

  // the ones tile (A operand of the row-sum product): 128 rows x 64 bf16 of 1.0 — swizzling a constant is a no-op
  {
    uint4* ones = reinterpret_cast<uint4*>(smem_raw + (sOnes - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < int(kTcRows * 128 / 16); i += blockDim.x)
      ones[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async();
  }
  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&maps.q);
      tma_prefetch_desc(&maps.k);
      tma_prefetch_desc(&maps.v);
      for (int s = 0; s < 2; ++s) {
        mbar_init(bar(B_QK, s), 1);
        mbar_init(bar(B_V, s), 1);
        mbar_init(bar(B_S, s), 1);
        mbar_init(bar(B_P, s), 4);
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
  int early = 0;
  if (warp == 4) {
    // whole K/V span of this CTA to L2 now, first two units' K/V into smem now (see attention_tc.cu)
    if (a.l2_prefetch) {
      int i = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const TcUnit t = tc_unit(p, u, 0);
        if (t.hidden) continue;
        const int krow = p.k_row0 + (t.b * a.Hkv + t.hkv) * a.Lmax + t.key0;
        for (int x = 0; x < 4; ++x, ++i)
          if ((i & 31) == lane) tma_prefetch_l2_2d(x < 2 ? &maps.k : &maps.v, (x & 1) * 64, krow);
      }
    }
    int n = 0;
    for (int u = blockIdx.x; u < n_units && n < 2; u += gridDim.x) {
      const TcUnit t = tc_unit(p, u, 0);
      if (t.hidden) continue;
      if (t.key0 + kTcKeys > a.kv_len) break;     // touches this window's keys: not before the wait
      const int s = n & 1;
      const uint32_t sK = base + uint32_t(s) * kStage, sV = sK + kKBytes;
      if (lane == 0) {
        mbar_arrive_expect_tx(bar(B_QK, s), uint32_t(t.heads_here) * 2u * uint32_t(p.Wp) * 128u + kKBytes);
        mbar_arrive_expect_tx(bar(B_V, s), kVBytes);
      }
      __syncwarp();
      const int krow = p.k_row0 + (t.b * a.Hkv + t.hkv) * a.Lmax + t.key0;
      if (lane < 2) tma_load_2d(sK + uint32_t(lane) * kTcKeys * 128, &maps.k, lane * 64, krow, bar(B_QK, s), kPolicyEvictFirst);
      else if (lane < 4)
        tma_load_2d(sV + uint32_t(lane - 2) * kTcKeys * 128, &maps.v, (lane - 2) * 64, krow, bar(B_V, s), kPolicyEvictFirst);
      ++n;
    }
    early = n;
  }
  pdl_wait();                 // wait, THEN release the dependents (attention_tc.cu explains why in this order)
  pdl_launch_dependents();
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[15] = clock64();

  if (warp == 4) {
    // ===== TMA producer =====
    int n = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const TcUnit t = tc_unit(p, u, 0);
      if (t.hidden) continue;
      const int s = n & 1, j = n >> 1;
      const uint32_t sK = base + uint32_t(s) * kStage, sV = sK + kKBytes, sQ = sV + kVBytes;
      const int nq = t.heads_here * 2;
      const bool kv_done = n < early;
      if (lane == 0) {
        if (j >= 1) mbar_wait_backoff(bar(B_O, s), uint32_t(j - 1) & 1u);   // the stage's previous products have read its smem
        if (!kv_done) {
          mbar_arrive_expect_tx(bar(B_QK, s), uint32_t(nq) * uint32_t(p.Wp) * 128u + kKBytes);
          mbar_arrive_expect_tx(bar(B_V, s), kVBytes);
        }
      }
      __syncwarp();
      const int krow = p.k_row0 + (t.b * a.Hkv + t.hkv) * a.Lmax + t.key0;
      for (int l = lane; l < 4 + nq; l += 32) {
        if (kv_done && l < 4) continue;
        if (l < 2) {
          tma_load_2d(sK + uint32_t(l) * kTcKeys * 128, &maps.k, l * 64, krow, bar(B_QK, s), kPolicyEvictFirst);
        } else if (l < 4) {
          tma_load_2d(sV + uint32_t(l - 2) * kTcKeys * 128, &maps.v, (l - 2) * 64, krow, bar(B_V, s), kPolicyEvictFirst);
        } else {
          const int x = l - 4, d = x & 1, hs = x >> 1;
          tma_load_2d(sQ + uint32_t(d) * kTctCols * 128 + uint32_t(hs * p.Wp) * 128, &maps.q, (t.h0 + hs) * DH + d * 64,
                      t.b * a.W, bar(B_QK, s), kPolicyEvictLast);
        }
      }
      if (p.dbg && blockIdx.x == 0 && lane == 0 && n < 8) p.dbg[n * 16 + 0] = clock64();
      ++n;
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16_f32(kTcKeys, kTctCols);              // A = K, B = Q: both K-major
      const uint32_t idesc_o = umma_idesc_bf16_f32_amn_bmn(DH, kTctCols);           // A = V (MN-major), B = P^T (MN-major)
      const uint32_t idesc_l = umma_idesc_bf16_f32_bmn(kTcRows, kTctCols);          // A = ones (K-major), B = P^T
      auto issue_s = [&](int m) {
        const int s = m & 1;
        const uint32_t sK = base + uint32_t(s) * kStage, sQ = sK + kKBytes + kVBytes;
        tcgen05_fence_after();
        const uint32_t tS = tmem_base + uint32_t(s) * 256;
        uint32_t acc = 0;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const uint64_t da = umma_desc_sw128_kmajor(sK + uint32_t(d) * kTcKeys * 128);
          const uint64_t db = umma_desc_sw128_kmajor(sQ + uint32_t(d) * kTctCols * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16_ss(tS, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc_s, acc);
            acc = 1;
          }
        }
        umma_commit(bar(B_S, s));
        if (p.dbg && blockIdx.x == 0 && m < 8) p.dbg[m * 16 + 2] = clock64();
      };
      auto issue_pv = [&](int m) {
        const int s = m & 1;
        const uint32_t sP = base + uint32_t(s) * kStage, sV = sP + kKBytes;   // P^T lives where K was
        tcgen05_fence_after();
        if (p.dbg && blockIdx.x == 0 && m < 8) p.dbg[m * 16 + 3] = clock64();
        const uint32_t tO = tmem_base + uint32_t(s) * 256 + 64, tL = tO + 64;
        const uint64_t d1 = umma_desc_sw128_kmajor(sOnes);
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < kTcKeys / 16; ++k) {   // 16 keys per step = two 8-row groups = 2 048 bytes of either tile
          const uint64_t dv = umma_desc_sw128_mnmajor(sV + uint32_t(k) * 2048, kTcKeys * 128);
          const uint64_t dp = umma_desc_sw128_mnmajor(sP + uint32_t(k) * 2048, kTcKeys * 128);
          umma_bf16_ss(tO, dv, dp, idesc_o, acc);
          umma_bf16_ss(tL, d1, dp, idesc_l, acc);
          acc = 1;
        }
        umma_commit(bar(B_O, s));
        if (p.dbg && blockIdx.x == 0 && m < 8) p.dbg[m * 16 + 4] = clock64();
      };
      int N = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) N += tc_unit(p, u, 0).hidden ? 0 : 1;
      int ns = 0, np = 0;   // next S^T / next O^T to issue: whichever has its inputs first (attention_tc.cu)
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
          if (mbar_test_wait(bar(B_QK, s), uint32_t(j) & 1u)) {
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
    // ===== softmax + epilogue.  Thread = TMEM lane = KEY of the tile (softmax) / head-dim element (epilogue) =====
    const int grp = warp >> 3, qw = warp & 3;
    const int kl = qw * 32 + lane;                          // lane of the accumulators
    const uint32_t t_row = (uint32_t(qw * 32) << 16);
    const float sc = a.scale_log2e;
    // deferred epilogue state (unit `owed`)
    int owed = -1, e_R = 0, e_h0 = 0;
    size_t e_base = 0;
    auto epilogue = [&](int m) {
      const int s = m & 1, j = m >> 1;
      mbar_wait(bar(B_O, s), uint32_t(j) & 1u);
      tcgen05_fence_after();
      if (p.dbg && blockIdx.x == 0 && grp == (m & 1) && qw == 0 && lane == 0 && m < 8) p.dbg[m * 16 + 8] = clock64();
      const uint32_t tO = tmem_base + uint32_t(s) * 256 + 64, tL = tO + 64;
      // slot c -> partial row: rows of one head are consecutive (512 bytes apart for this lane's element), the next
      // head starts W rows later: one pointer walked with adds, no per-element address arithmetic
      float* pcol = a.part_o + (e_base + size_t(e_h0) * a.W) * DH + kl;
      int qi = 0;
#pragma unroll 1
      for (int c0 = 0; c0 < e_R; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tO + t_row + uint32_t(c0), v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          if (qi < a.W && c0 + e < e_R) *pcol = __uint_as_float(v[e]);   // uniform: a warp writes 128 contiguous bytes
          pcol += DH;
          if (++qi == p.Wp) { qi = 0; pcol -= (p.Wp - a.W) * DH; }
        }
      }
      if (qw == 0) {   // row sums: every lane of L holds them; lane c (and 32 + c) of warp 0 writes column c's
#pragma unroll 1
        for (int c0 = 0; c0 < e_R; c0 += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(tL + t_row + uint32_t(c0), v);
          tmem_ld_wait();
          const int c = c0 + (lane & 15);
          if ((lane >> 4) == ((c0 >> 4) & 1) && c < e_R) {
            float l = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if ((lane & 15) == e) l = __uint_as_float(v[e]);
            const int hs2 = tc_div(c, p.m_wp, p.Wp), qi2 = c - hs2 * p.Wp;
            if (qi2 < a.W) a.part_ml[(e_base + size_t(e_h0 + hs2) * a.W + qi2) * 2 + 1] = l;
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_E, s));
      if (p.dbg && blockIdx.x == 0 && qw == 0 && lane == 0 && m < 8) p.dbg[m * 16 + 9] = clock64();
    };
    int n = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const TcUnit t = tc_unit(p, u, 0);
      const size_t ubase = (size_t(t.kt) * a.rows + t.b) * a.H * size_t(a.W);
      if (t.hidden) {   // uniform per CTA: nothing visible in this tile
        if (grp == 0 && qw == 0) {
          for (int c = lane; c < t.R; c += 32) {
            const int hs2 = tc_div(c, p.m_wp, p.Wp), qi2 = c - hs2 * p.Wp;
            if (qi2 < a.W) {
              const size_t pr = ubase + size_t(t.h0 + hs2) * a.W + qi2;
              a.part_ml[pr * 2] = -INFINITY;
              a.part_ml[pr * 2 + 1] = 0.f;
            }
          }
        }
        continue;
      }
      if ((n & 1) != grp) {
        ++n;
        continue;
      }
      const int s = n & 1, j = n >> 1;
      if (owed >= 0) epilogue(owed);
      uint8_t* const genP = smem_raw + (base + uint32_t(s) * kStage - smem_u32(smem_raw));
      const uint32_t tS = tmem_base + uint32_t(s) * 256;
      mbar_wait(bar(B_S, s), uint32_t(j) & 1u);
      tcgen05_fence_after();
      if (p.dbg && blockIdx.x == 0 && qw == 0 && lane == 0 && n < 8) p.dbg[n * 16 + 5] = clock64();
      const int R = t.R, lo = t.lo;
      const int jk = t.key0 + kl;                            // this thread's key
      const bool key_ok = jk >= lo && jk < T;
      // every (key, query) pair of the tile visible: no mask arithmetic
      const bool interior = (t.key0 >= lo) && (t.key0 + kTcKeys - 1 <= a.kv_len) && (t.key0 + kTcKeys <= T);
      float* const st = reinterpret_cast<float*>(smem_raw + (sSt + uint32_t(grp) * kStBytes - smem_u32(smem_raw)));
      float* const xmf = xmf_all + grp * kTctCols;
      uint8_t* rowp = genP + kl * 128;
      // 32 columns at a time: (a) every key thread parks its 32 scores in the staging tile; (b) thread (quarter q, column
      // c) takes the max over the 32 keys of quarter q — no cross-lane traffic, its mask is a key range — and two
      // shuffles merge the quarters; (c) the 32 column maxima go back through shared memory and every key thread turns
      // its scores into probabilities
#pragma unroll 1
      for (int ch0 = 0; ch0 < R; ch0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x16(tS + t_row + uint32_t(ch0), *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        tmem_ld_32x32b_x16(tS + t_row + uint32_t(ch0 + 16), *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
        tmem_ld_wait();
        {
          float* row = st + kl * 33 + 8 * qw;               // +8 floats per 32 keys: step (b) reads without bank conflicts
#pragma unroll
          for (int e = 0; e < 32; ++e) row[e] = __uint_as_float(v[e]);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        {
          const int q = lane >> 3, c = ch0 + 8 * qw + (lane & 7);       // this thread: keys [32 q, 32 q + 32) of column c
          const int hs2 = tc_div(c, p.m_wp, p.Wp), qi2 = c - hs2 * p.Wp;
          // visible keys of this column inside the quarter: [k_lo, k_hi]
          const int jq = t.key0 + 32 * q;
          int k_lo = 0, k_hi = 31;
          if (!interior) {
            k_lo = max(lo - jq, 0);
            k_hi = min(min(a.kv_len + qi2, T - 1) - jq, 31);
          }
          if (qi2 >= a.W || c >= R) k_hi = -1;
          const float* col = st + (32 * q) * 33 + 8 * q + (c - ch0);
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float x = col[k * 33];
            m4[k & 3] = fmaxf(m4[k & 3], (k >= k_lo && k <= k_hi) ? x : -INFINITY);
          }
          float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
          if (q == 0) {
            const float ms = mx * sc;                        // scale > 0: max commutes with it
            xmf[c] = ms;
            if (qi2 < a.W && c < R) a.part_ml[(ubase + size_t(t.h0 + hs2) * a.W + qi2) * 2] = ms;
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        if (p.dbg && blockIdx.x == 0 && qw == 0 && lane == 0 && n < 8) p.dbg[n * 16 + 6] = clock64();
        // (c) probabilities of this key for the 32 columns (bf16, q contiguous) -> four 16-byte chunks of its P^T row
        {
          float mcol[32];
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 m = *reinterpret_cast<const float4*>(xmf + ch0 + e);
            mcol[e] = m.x; mcol[e + 1] = m.y; mcol[e + 2] = m.z; mcol[e + 3] = m.w;
          }
          int qi = ch0 % p.Wp;                               // ch0 is 0 or 32; Wp divides 64 boundaries only when it divides 32
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float pe[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float ms = (mcol[e + h] == -INFINITY) ? 0.f : mcol[e + h];
              const bool ok = interior ? (qi < a.W) : (key_ok && jk <= a.kv_len + qi && qi < a.W);
              pe[h] = ok ? ex2_approx(fmaf(__uint_as_float(v[e + h]), sc, -ms)) : 0.f;
              if (++qi == p.Wp) qi = 0;
            }
            const __nv_bfloat162 pb = __floats2bfloat162_rn(pe[0], pe[1]);
            pk[e >> 1] = *reinterpret_cast<const uint32_t*>(&pb);
          }
          const int ch = ch0 >> 3;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            *reinterpret_cast<uint4*>(rowp + (((ch + q4) ^ (kl & 7)) << 4)) =
                make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
        }
      }
      {
        // slots beyond R feed accumulator columns nobody reads; zero them once per unit so they stay finite
        for (int c0 = (R + 31) & ~31; c0 < kTctCols; c0 += 8)
          *reinterpret_cast<uint4*>(rowp + (((c0 >> 3) ^ (kl & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
      }
      tcgen05_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_P, s));
      if (p.dbg && blockIdx.x == 0 && qw == 0 && lane == 0 && n < 8) p.dbg[n * 16 + 7] = clock64();
      owed = n;
      e_R = R; e_h0 = t.h0; e_base = ubase;
      ++n;
    }
    if (owed >= 0) epilogue(owed);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Geometry: up to 64 / Wp heads of a kv head per unit (64 accumulator columns).
void attn_tct_plan(AttnTcParams* p) {
  AttnParams& a = p->a;
  const int T = a.kv_len + a.W, G = a.H / a.Hkv;
  p->head_dim = 128;
  p->Wp = (a.W + 7) & ~7;
  int hpc = kTctCols / p->Wp;
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

constexpr int attn_tct_smem() {   // two stages of {K | P^T, V, Q}, the ones tile, two staging tiles for the column maxima
  return 1024 + 2 * (2 * kTcKeys * 128 + 2 * kTcKeys * 128 + 2 * kTctCols * 128) + kTcRows * 128 + 2 * (kTcKeys * 33 + 32) * 4;
}

int attn_tct_launch(const AttnTcMaps& maps, const AttnTcParams& p, cudaStream_t stream) {
  const AttnParams& a = p.a;
  if (p.Wp > kTctCols || p.head_dim != 128) return -3;
  const int G = a.H / a.Hkv, mtiles = (G + p.hpc - 1) / p.hpc;
  const int n_units = a.n_chunks * a.Hkv * mtiles * a.rows;
  const int sms = device_num_sms();
  dim3 grid(n_units < sms ? n_units : sms);
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(attn_tct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_tct_smem()) != cudaSuccess)
      return -5;
    set = true;
  }
  return launch_pdl(attn_tct_kernel, grid, dim3(kTcThreads2), attn_tct_smem(), stream, maps, p);
}

}  // namespace sjd
