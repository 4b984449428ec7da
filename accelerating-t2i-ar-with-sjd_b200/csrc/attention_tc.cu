// Jacobi draft-window attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Same contract as attention.cu (key j visible to window query i of CFG row b iff kv_lo[b] <= j <= kv_len + i; one
// fp32 partial {sum p*v, max, sum p} per key split, merged by attn_combine_row in the next chain kernel), different
// machine mapping:
//   * one CTA per (128-key tile, kv head [x query-row tile], CFG row); two CTAs share an SM (97 KB smem, 256 TMEM
//     columns each), so one CTA's TMA latency hides behind the other's math;
//   * the query rows of up to 128 / Wp heads that share the kv head are stacked along UMMA M (GQA: the K/V tile is
//     loaded once for all of them); rows beyond the stack are garbage that only ever produces garbage rows;
//   * S[128 q x 128 keys] = Q K^T :  A = Q tile, B = K tile, both K-major (head dim contiguous), SWIZZLE_128B boxes
//     straight from the q buffer / K cache by TMA;
//   * softmax on the TMEM accumulator with one thread per query row (row max and row sum are thread-local): two
//     passes of tcgen05.ld, probabilities rounded to bf16 like the reference's bf16 SDPA and written into the smem the K
//     tile occupied, as the K-major A operand of the second product;
//   * O[128 q x Dh] = P V :  B = the V tile exactly as TMA lands it ([key][head dim], head dim contiguous), consumed as
//     an MN-major operand — no transposed cache, no smem transpose;
//   * when the stacked rows fill only 1/rep of the 128 TMEM lanes (MHA, window 32: a quarter), Q is loaded rep times
//     and replica g takes columns [128 g / rep, 128 (g+1) / rep) of every row: all four softmax warps work, row max and
//     row sum meet in shared memory, and every replica writes its slice of P into replica 0's row;
//   * a CTA sees its whole key span at once, so there is no online rescaling: max, exponentials, one product.
// Reference semantics: SDPA over the additive window mask (modeling_chameleon.py:567-574, mask from
// scheduler/jacobi_iteration_lumina_mgpt.py:1256-1336; llamagen/llamagen.py:269-273).
#include "common.cuh"

namespace sjd {

constexpr int kTcKeys = 128;       // keys per CTA (UMMA N of the first product, K of the second)
constexpr int kTcRows = 128;       // query-row slots per CTA (UMMA M)
constexpr int kTcThreads = 160;    // warps 0..3: softmax / epilogue (TMEM lane quarter = warp), warp 4: TMA + MMA issue

struct AttnTcParams {
  AttnParams a;          // geometry, partial buffers, kv_len / kv_lo (n_chunks = key tiles, span = kTcKeys)
  int head_dim;          // 64 | 128
  int Wp;                // row slots per head: W rounded up to 8 (swizzle atom) — the Q tensor map's box height
  int hpc;               // heads stacked per CTA
  int k_row0;            // first row of this layer in the [layers*rows*Hkv*Lmax, Dh] view of the caches
};

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

struct AttnTcMaps {
  CUtensorMap q, k, v;
};

template <int DH>
__global__ void __launch_bounds__(kTcThreads, 2)
attn_tc_kernel(const __grid_constant__ AttnTcMaps maps, const AttnTcParams p) {
  constexpr int NDA = DH / 64;                            // 64-wide head-dim atoms
  constexpr uint32_t kQBytes = NDA * kTcRows * 128;       // Q tile: NDA atoms of [128 rows][128 B]
  constexpr uint32_t kKBytes = NDA * kTcKeys * 128;       // K tile, same shape
  constexpr uint32_t kPBytes = 2 * kTcRows * 128;         // P tile: two 64-key atoms of [128 rows][128 B]
  constexpr uint32_t kKPBytes = kKBytes > kPBytes ? kKBytes : kPBytes;
  constexpr uint32_t kVBytes = NDA * kTcKeys * 128;       // V tile: NDA boxes of [128 keys][128 B]
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ uint32_t tmem_holder;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = base + kQBytes, sV = sK + kKPBytes;
  uint8_t* const genP = smem_raw + (sK - smem_u32(smem_raw));
  const uint32_t bar_load = smem_u32(&bars[0]), bar_s = smem_u32(&bars[1]), bar_p = smem_u32(&bars[2]),
                 bar_o = smem_u32(&bars[3]);
  const AttnParams& a = p.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, b = blockIdx.z;
  const int G = a.H / a.Hkv;
  const int mtiles = (G + p.hpc - 1) / p.hpc;
  const int hkv = blockIdx.y / mtiles, mt = blockIdx.y - hkv * mtiles;
  const int h0 = hkv * G + mt * p.hpc;                      // first query head of this CTA
  const int heads_here = min(p.hpc, G - mt * p.hpc);
  const int T = a.kv_len + a.W, lo = a.kv_lo[b];
  const int key0 = kt * kTcKeys;
  const bool hidden = key0 + kTcKeys <= lo;                 // whole tile inside the hidden prefix: nothing to load

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&maps.q);
      tma_prefetch_desc(&maps.k);
      tma_prefetch_desc(&maps.v);
      mbar_init(bar_load, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 4);
      mbar_init(bar_o, 1);
      fence_barrier_init();
    }
    if (!hidden) {
      tmem_alloc(smem_u32(&tmem_holder), 256);
      tmem_relinquish();
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  // Wait first, THEN let the next kernel in: releasing the dependents before the wait lets the whole forward cascade
  // into residency (chain l+1 behind attention l+1 behind chain l ...), which was measured to deadlock.
  pdl_wait();   // q, this window's K/V rows and the partial buffers belong to the previous kernels until here
  pdl_launch_dependents();

  // softmax thread r = TMEM lane: (column replica g, head slot hs, window position qi)
  const int R = heads_here * p.Wp;                          // row slots in use
  const int Rr = (R + 31) & ~31;                            // replica pitch: whole warps, so a warp never straddles two
  const int rep = Rr <= 32 ? 4 : (Rr <= 64 ? 2 : 1);        //   replicas (tcgen05.ld is warp-collective: uniform columns)
  const int r = warp * 32 + lane;
  const int g = r / Rr, rr = r - g * Rr;
  const int hs = rr / p.Wp, qi = rr - hs * p.Wp;
  const bool row_ok = warp < 4 && rr < R && qi < a.W;       // this thread works on a real query row
  const bool valid = row_ok && g == 0;                      // ... and owns its output
  const size_t prow = valid ? ((size_t(kt) * a.rows + b) * a.H + (h0 + hs)) * size_t(a.W) + qi : 0;
  const int ncol = kTcKeys / rep, col0 = g * ncol;          // this thread's slice of the row's 128 keys
  __shared__ float xch[kTcRows];                            // row max, then row sum, across replicas

  if (hidden) {   // uniform per CTA
    if (valid) {
      a.part_ml[prow * 2] = -INFINITY;
      a.part_ml[prow * 2 + 1] = 0.f;
    }
    return;
  }
  const uint32_t tmem_base = tmem_holder;
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == 4) {
    if (lane == 0) {
      // ---- loads ----
      const uint32_t bytes = uint32_t(rep * heads_here) * NDA * uint32_t(p.Wp) * 128u + kKBytes + kVBytes;
      mbar_arrive_expect_tx(bar_load, bytes);
      const int krow = p.k_row0 + (b * a.Hkv + hkv) * a.Lmax + key0;
#pragma unroll
      for (int d = 0; d < NDA; ++d) {
        tma_load_2d(sK + uint32_t(d) * kTcKeys * 128, &maps.k, d * 64, krow, bar_load, kPolicyEvictFirst);
        tma_load_2d(sV + uint32_t(d) * kTcKeys * 128, &maps.v, d * 64, krow, bar_load, kPolicyEvictFirst);
      }
      for (int gq = 0; gq < rep; ++gq)
        for (int s = 0; s < heads_here; ++s)
#pragma unroll
          for (int d = 0; d < NDA; ++d)
            tma_load_2d(sQ + uint32_t(d) * kTcRows * 128 + uint32_t(gq * Rr + s * p.Wp) * 128, &maps.q,
                        (h0 + s) * DH + d * 64, b * a.W, bar_load, kPolicyEvictLast);
      mbar_wait(bar_load, 0);
      tcgen05_fence_after();
      // ---- S = Q K^T ----
      const uint32_t idesc_s = umma_idesc_bf16_f32(kTcRows, kTcKeys);
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
      umma_commit(bar_s);
      // ---- O = P V (P arrives in the smem the K tile occupied) ----
      mbar_wait(bar_p, 0);
      tcgen05_fence_after();
      const uint32_t idesc_o = umma_idesc_bf16_f32_bmn(kTcRows, DH);
      acc = 0;
#pragma unroll
      for (int ka = 0; ka < 2; ++ka) {
        const uint64_t da = umma_desc_sw128_kmajor(sK + uint32_t(ka) * kTcRows * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t db = umma_desc_sw128_mnmajor(sV + uint32_t(ka * 64 + k * 16) * 128, kTcKeys * 128);
          umma_bf16_ss(tO, da + uint64_t(2 * k), db, idesc_o, acc);
          acc = 1;
        }
      }
      umma_commit(bar_o);
    }
  } else {
    // ---- softmax: one thread per query row, straight from the TMEM accumulator ----
    mbar_wait(bar_s, 0);
    tcgen05_fence_after();
    const uint32_t t_row = (uint32_t(warp * 32) << 16);
    // keys every query of this CTA sees: no per-element mask arithmetic for them
    const bool interior = (key0 >= lo) && (key0 + kTcKeys - 1 <= a.kv_len) && (key0 + kTcKeys <= T);
    const int j_hi = min(a.kv_len + qi, T - 1);   // last visible key of this row
    float mx = -INFINITY;
    if (g < rep) {
#pragma unroll 1
      for (int c = col0; c < col0 + ncol; c += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tS + t_row + uint32_t(c), v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int j = key0 + c + e;
          const bool ok = interior || (j >= lo && j <= j_hi);
          if (ok) mx = fmaxf(mx, __uint_as_float(v[e]));
        }
      }
    }
    if (rep > 1) {   // uniform per CTA: the row max is spread over the replicas
      xch[r] = mx;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (g < rep) {
        for (int g2 = 0; g2 < rep; ++g2) mx = fmaxf(mx, xch[g2 * Rr + rr]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    mx *= a.scale_log2e;                                    // scale > 0: max commutes with it
    const float ms = (mx == -INFINITY) ? 0.f : mx;
    float lsum = 0.f;
    if (g < rep) {
#pragma unroll 1
      for (int c = col0; c < col0 + ncol; c += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tS + t_row + uint32_t(c), v);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const int j = key0 + c + e;
          const bool ok0 = interior || (j >= lo && j <= j_hi), ok1 = interior || (j + 1 >= lo && j + 1 <= j_hi);
          const float e0 = ok0 ? exp2f(__uint_as_float(v[e]) * a.scale_log2e - ms) : 0.f;
          const float e1 = ok1 ? exp2f(__uint_as_float(v[e + 1]) * a.scale_log2e - ms) : 0.f;
          const __nv_bfloat162 pb = __floats2bfloat162_rn(e0, e1);   // probabilities enter P*V as bf16
          lsum += __low2float(pb) + __high2float(pb);
          pk[e >> 1] = *reinterpret_cast<const uint32_t*>(&pb);
        }
        if (row_ok) {
          // row rr (replica 0's) of the K-major P tile: 64-key atom c / 64, 16-byte chunks (c % 64) / 8 and the next
          // one, 128B-swizzled
          uint8_t* rowp = genP + (c >> 6) * (kTcRows * 128) + rr * 128;
          const int ch = (c & 63) >> 3;
          *reinterpret_cast<uint4*>(rowp + (((ch) ^ (rr & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ (rr & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
    }
    if (rep > 1) {
      xch[r] = lsum;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (g == 0) {
        lsum = 0.f;
        for (int g2 = 0; g2 < rep; ++g2) lsum += xch[g2 * Rr + rr];   // fixed order
      }
    }
    tcgen05_fence_before();
    fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
    // ---- epilogue: unnormalised O row + {max, sum} as this split's partial ----
    mbar_wait(bar_o, 0);
    tcgen05_fence_after();
    float* po = a.part_o + prow * DH;
#pragma unroll 1
    for (int c = 0; c < DH; c += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tO + t_row + uint32_t(c), v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int e = 0; e < 16; e += 4)
          *reinterpret_cast<float4*>(po + c + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                               __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
      }
    }
    if (valid) {
      a.part_ml[prow * 2] = mx;
      a.part_ml[prow * 2 + 1] = lsum;
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 256);
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
}

constexpr int attn_tc_smem(int head_dim) {
  return 1024 + (head_dim / 64) * kTcRows * 128 + 2 * kTcRows * 128 + (head_dim / 64) * kTcKeys * 128;
}

int attn_tc_launch(const AttnTcMaps& maps, const AttnTcParams& p, cudaStream_t stream) {
  const AttnParams& a = p.a;
  if (a.W > kTcRows || (p.head_dim != 64 && p.head_dim != 128)) return -3;
  const int G = a.H / a.Hkv, mtiles = (G + p.hpc - 1) / p.hpc;
  dim3 grid(a.n_chunks, a.Hkv * mtiles, a.rows);
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(attn_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_tc_smem(128)) != cudaSuccess ||
        cudaFuncSetAttribute(attn_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_tc_smem(64)) != cudaSuccess)
      return -5;
    set = true;
  }
  if (p.head_dim == 128) return launch_pdl(attn_tc_kernel<128>, grid, dim3(kTcThreads), attn_tc_smem(128), stream, maps, p);
  return launch_pdl(attn_tc_kernel<64>, grid, dim3(kTcThreads), attn_tc_smem(64), stream, maps, p);
}

}  // namespace sjd
