// Jacobi draft-window attention over the static KV cache (decode-shaped, split-KV).
// For every CFG row b, kv head and key split one CTA streams its span of the K/V cache through shared memory in
// 64-key sub-chunks (cp.async double buffering, zero-filled past the end) while every warp runs flash-style online
// softmax for one 16-query tile of one query head (mma.sync m16n8k16 bf16, fp32 accumulate).  The Jacobi window mask of the
// reference — key j visible to window query i iff  kv_lo[b] <= j <= kv_len + i — is evaluated from
// indices; no mask tensor exists (reference builds [2B,1,W,T+W] additive masks:
// scheduler/jacobi_iteration_lumina_mgpt.py:1256-1336, consumed by SDPA at
// lumina_mgpt/model/chameleon/modeling_chameleon.py:567-574; llamagen/llamagen.py:269-273).
// A second kernel merges the per-chunk partials (log-sum-exp combine) into the bf16 attention output.
#include "common.cuh"

namespace sjd {

constexpr int kAttnSub = 64;      // keys per online-softmax step
constexpr int kAttnMaxRows = 8;   // CFG rows supported by the descriptor

struct AttnParams {
  const __nv_bfloat16* q;   // [rows*W][H][Dh]
  const __nv_bfloat16* k;   // layer base: [rows][Hkv][Lmax][Dh]
  const __nv_bfloat16* v;
  float* part_o;            // [chunks][rows][H][W][Dh]
  float* part_ml;           // [chunks][rows][H][W][2]
  __nv_bfloat16* out;       // [rows*W][H*Dh]
  int rows, W, H, Hkv, Lmax;
  int kv_len;               // keys cached before this window
  int kv_lo[kAttnMaxRows];  // first visible key per row
  int n_chunks;             // key splits (one partial each)
  int span;                 // keys per split (multiple of kAttnSub)
  int l2_prefetch;          // 1: every CTA sends its K/V span to L2 before the dependency wait
  float scale_log2e;        // softmax scale * log2(e)
};

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                                  uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// grid: (key splits, Hkv * tile groups, rows); block: TWO warps per 16-query tile (<= 4 tiles per CTA).
// Each CTA walks its span of keys in 64-key sub-chunks, double buffered with cp.async so that the K/V stream of
// sub-chunk s+1 is in flight while sub-chunk s is being multiplied.  The two warps of a tile take one 32-key half of
// every sub-chunk each (twice the warps in flight for the same shared memory) and merge their online-softmax states
// through shared memory at the end.  Only sub-chunks that touch the hidden prefix or the causal window are masked
// element by element; interior ones skip the mask arithmetic.
constexpr int kAttnHalf = kAttnSub / 2;   // keys per warp per sub-chunk
constexpr int kAttnMaxTiles = 4;          // query tiles per CTA (8 warps)

// STAGES = depth of the cp.async ring (2: the load of sub-chunk s+1 overlaps the math of s; 3: two sub-chunks in flight —
// the round-2 ncu capture shows the 2-stage form waiting on global memory, long_scoreboard being its top stall reason)
template <int DH, int NTHREADS, int MINB, int STAGES = 2>
__global__ void __launch_bounds__(NTHREADS, MINB) attn_window_kernel(AttnParams p) {
  constexpr int ROWB = DH * 2 + 16;  // padded smem row (bytes): conflict-free ldmatrix
  constexpr int STAGE = kAttnSub * ROWB;
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* sKb = smem;                    // [STAGES][kAttnSub][ROWB]
  uint8_t* sVb = smem + STAGES * STAGE;   // [STAGES][kAttnSub][ROWB]

  const int split = blockIdx.x, hkv = blockIdx.y % p.Hkv, tgroup = blockIdx.y / p.Hkv, b = blockIdx.z;
  if (p.l2_prefetch && threadIdx.x == 0) {
    // This CTA's key span is one contiguous piece of the K and of the V cache: send both to L2 while the previous
    // kernel is still finishing (L2 is the coherence point, so rows that kernel is still writing are simply written
    // into the prefetched lines); the cp.async stream below then pays L2 latency instead of HBM latency.
    const int kb = max(split * p.span, (p.kv_lo[b] / kAttnSub) * kAttnSub);
    const int ke = min(p.kv_len + p.W, split * p.span + p.span);
    if (ke > kb) {
      const size_t off = ((size_t(b) * p.Hkv + hkv) * size_t(p.Lmax) + kb) * DH;
      const uint32_t bytes = uint32_t(ke - kb) * DH * 2;   // multiple of 16
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.k + off), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.v + off), "r"(bytes) : "memory");
    }
  }
  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tiles_here = blockDim.x >> 6;
  const int T = p.kv_len + p.W;
  const int G = p.H / p.Hkv;
  const int tiles_per_head = (p.W + 15) >> 4;
  const int n_tiles = G * tiles_per_head;
  const int lo = p.kv_lo[b];
  const int tl = warp >> 1, kh = warp & 1;   // tile slot in this CTA, key half
  const int tile = tgroup * tiles_here + tl;
  const bool has_tile = tile < n_tiles;
  const int hq = hkv * G + (has_tile ? tile / tiles_per_head : 0);
  const int i0 = has_tile ? (tile % tiles_per_head) * 16 : 0;
  const int g = lane >> 2, t = lane & 3;
  const int r0 = i0 + g, r1 = i0 + g + 8;  // query indices of this thread's two rows

  // span of this split, trimmed to what any query of the window can see: keys [lo, T)
  int k_begin = split * p.span, k_end = min(T, k_begin + p.span);
  k_begin = max(k_begin, (lo / kAttnSub) * kAttnSub);
  const int n_sub = k_end > k_begin ? (k_end - k_begin + kAttnSub - 1) / kAttnSub : 0;
  const __nv_bfloat16* kg = p.k + (size_t(b) * p.Hkv + hkv) * size_t(p.Lmax) * DH;
  const __nv_bfloat16* vg = p.v + (size_t(b) * p.Hkv + hkv) * size_t(p.Lmax) * DH;
  auto load_sub = [&](int s) {
    constexpr int VEC_PER_ROW = DH / 8;
    const int key0 = k_begin + s * kAttnSub, nkeys = min(kAttnSub, k_end - key0);
    uint8_t* dK = sKb + (s % STAGES) * STAGE;
    uint8_t* dV = sVb + (s % STAGES) * STAGE;
    for (int idx = threadIdx.x; idx < kAttnSub * VEC_PER_ROW; idx += blockDim.x) {
      const int r = idx / VEC_PER_ROW, c = idx - r * VEC_PER_ROW;
      const int nb = r < nkeys ? 16 : 0;   // rows past the end are zero-filled
      const size_t goff = size_t(key0 + (r < nkeys ? r : 0)) * DH + c * 8;
      cp_async_16(smem_u32(dK + r * ROWB + c * 16), kg + goff, nb);
      cp_async_16(smem_u32(dV + r * ROWB + c * 16), vg + goff, nb);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float o[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  uint32_t qf[DH / 16][4];
  if (n_sub > 0) {
    load_sub(0);
    if (STAGES > 2 && n_sub > 1) load_sub(1);
    // Q fragments straight from global memory
    const __nv_bfloat16* q0 = p.q + (size_t(b * p.W + min(r0, p.W - 1)) * p.H + hq) * DH;
    const __nv_bfloat16* q1 = p.q + (size_t(b * p.W + min(r1, p.W - 1)) * p.H + hq) * DH;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
      const int c = kk * 16 + t * 2;
      qf[kk][0] = (has_tile && r0 < p.W) ? *reinterpret_cast<const uint32_t*>(q0 + c) : 0u;
      qf[kk][1] = (has_tile && r1 < p.W) ? *reinterpret_cast<const uint32_t*>(q1 + c) : 0u;
      qf[kk][2] = (has_tile && r0 < p.W) ? *reinterpret_cast<const uint32_t*>(q0 + c + 8) : 0u;
      qf[kk][3] = (has_tile && r1 < p.W) ? *reinterpret_cast<const uint32_t*>(q1 + c + 8) : 0u;
    }
  }
#pragma unroll 1
  for (int s = 0; s < n_sub; ++s) {
    if (STAGES == 2) {
      if (s + 1 < n_sub) {
        load_sub(s + 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
    } else {   // sub-chunk s + 2 goes into the stage that iteration s - 1 released at its closing barrier
      if (s + 2 < n_sub) {
        load_sub(s + 2);
        asm volatile("cp.async.wait_group 2;" ::: "memory");
      } else if (s + 1 < n_sub) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
    }
    __syncthreads();
    const int key0 = k_begin + s * kAttnSub + kh * kAttnHalf;   // first key of this warp's half
    // causal skip: every key of this half is beyond the last query of the tile, or past the end of the span
    if (has_tile && key0 < k_end && key0 <= p.kv_len + min(i0 + 15, p.W - 1)) {
      const uint8_t* sK = sKb + (s % STAGES) * STAGE + kh * kAttnHalf * ROWB;
      const uint8_t* sV = sVb + (s % STAGES) * STAGE + kh * kAttnHalf * ROWB;
      float sc[kAttnHalf / 8][4];
#pragma unroll
      for (int n = 0; n < kAttnHalf / 8; ++n) sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
#pragma unroll
        for (int np = 0; np < kAttnHalf / 16; ++np) {
          // matrices: (keys 0-7,d 0-7) (keys 0-7,d 8-15) (keys 8-15,d 0-7) (keys 8-15,d 8-15)
          const int key = np * 16 + (lane & 7) + ((lane >> 4) << 3);
          const int dof = kk * 16 + (((lane >> 3) & 1) << 3);
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(smem_u32(sK + key * ROWB + dof * 2), b0, b1, b2, b3);
          mma_bf16_16816(sc[np * 2], qf[kk], b0, b1);
          mma_bf16_16816(sc[np * 2 + 1], qf[kk], b2, b3);
        }
      }
      // scale into the log2 domain; mask only where the half touches the hidden prefix, the causal window or the end
      const bool interior = (key0 >= lo) && (key0 + kAttnHalf - 1 <= p.kv_len + i0) && (key0 + kAttnHalf <= k_end) &&
                            (i0 + 15 < p.W);
      float mx0 = -INFINITY, mx1 = -INFINITY;
      if (interior) {
#pragma unroll
        for (int n = 0; n < kAttnHalf / 8; ++n) {
#pragma unroll
          for (int e = 0; e < 4; ++e) sc[n][e] *= p.scale_log2e;
          mx0 = fmaxf(mx0, fmaxf(sc[n][0], sc[n][1]));
          mx1 = fmaxf(mx1, fmaxf(sc[n][2], sc[n][3]));
        }
      } else {
#pragma unroll
        for (int n = 0; n < kAttnHalf / 8; ++n) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = key0 + n * 8 + t * 2 + (e & 1);
            const int qi = (e < 2) ? r0 : r1;
            const bool ok = (j >= lo) && (j <= p.kv_len + qi) && (j < k_end) && (qi < p.W);
            const float val = ok ? sc[n][e] * p.scale_log2e : -INFINITY;
            sc[n][e] = val;
            if (e < 2) mx0 = fmaxf(mx0, val); else mx1 = fmaxf(mx1, val);
          }
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
      const float ms0 = (mn0 == -INFINITY) ? 0.f : mn0, ms1 = (mn1 == -INFINITY) ? 0.f : mn1;
      const float a0 = exp2f(m0 - ms0), a1 = exp2f(m1 - ms1);  // m == -inf -> 0
      m0 = mn0; m1 = mn1;
      l0 *= a0; l1 *= a1;
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) { o[n][0] *= a0; o[n][1] *= a0; o[n][2] *= a1; o[n][3] *= a1; }
      float ps0 = 0.f, ps1 = 0.f;
      uint32_t pf[kAttnHalf / 16][4];
#pragma unroll
      for (int n = 0; n < kAttnHalf / 8; ++n) {
        const float e0 = exp2f(sc[n][0] - ms0), e1 = exp2f(sc[n][1] - ms0);
        const float e2 = exp2f(sc[n][2] - ms1), e3 = exp2f(sc[n][3] - ms1);
        // probabilities enter P*V as bf16, like the reference's bf16 SDPA
        const uint32_t p01 = pack_bf16(e0, e1), p23 = pack_bf16(e2, e3);
        const __nv_bfloat162 v01 = *reinterpret_cast<const __nv_bfloat162*>(&p01);
        const __nv_bfloat162 v23 = *reinterpret_cast<const __nv_bfloat162*>(&p23);
        ps0 += __low2float(v01) + __high2float(v01);
        ps1 += __low2float(v23) + __high2float(v23);
        pf[n >> 1][(n & 1) * 2 + 0] = p01;
        pf[n >> 1][(n & 1) * 2 + 1] = p23;
      }
      l0 += ps0; l1 += ps1;
#pragma unroll
      for (int kk = 0; kk < kAttnHalf / 16; ++kk) {
#pragma unroll
        for (int dp = 0; dp < DH / 16; ++dp) {
          // trans matrices: (keys 0-7,d 0-7) (keys 8-15,d 0-7) (keys 0-7,d 8-15) (keys 8-15,d 8-15)
          const int key = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
          const int dof = dp * 16 + ((lane >> 4) << 3);
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4_trans(smem_u32(sV + key * ROWB + dof * 2), b0, b1, b2, b3);
          mma_bf16_16816(o[dp * 2], pf[kk], b0, b1);
          mma_bf16_16816(o[dp * 2 + 1], pf[kk], b2, b3);
        }
      }
    }
    __syncthreads();   // everybody is done with this stage before sub-chunk s+2 lands in it
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  // ---- merge the two key halves of each tile through shared memory (the K/V stages are free now) ----
  float* xch = reinterpret_cast<float*>(smem) + size_t(tl) * (16 * (DH + 2));   // [16 rows][DH + 2]
  if (kh == 1) {
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      *reinterpret_cast<float2*>(xch + g * (DH + 2) + n * 8 + t * 2) = make_float2(o[n][0], o[n][1]);
      *reinterpret_cast<float2*>(xch + (g + 8) * (DH + 2) + n * 8 + t * 2) = make_float2(o[n][2], o[n][3]);
    }
    if (t == 0) {
      xch[g * (DH + 2) + DH] = m0; xch[g * (DH + 2) + DH + 1] = l0;
      xch[(g + 8) * (DH + 2) + DH] = m1; xch[(g + 8) * (DH + 2) + DH + 1] = l1;
    }
  }
  __syncthreads();
  if (kh == 1 || !has_tile) return;
  {
    const float om0 = xch[g * (DH + 2) + DH], ol0 = xch[g * (DH + 2) + DH + 1];
    const float om1 = xch[(g + 8) * (DH + 2) + DH], ol1 = xch[(g + 8) * (DH + 2) + DH + 1];
    const float nm0 = fmaxf(m0, om0), nm1 = fmaxf(m1, om1);
    const float s0 = (nm0 == -INFINITY) ? 0.f : nm0, s1 = (nm1 == -INFINITY) ? 0.f : nm1;
    const float wa0 = exp2f(m0 - s0), wb0 = exp2f(om0 - s0), wa1 = exp2f(m1 - s1), wb1 = exp2f(om1 - s1);
    l0 = l0 * wa0 + ol0 * wb0;
    l1 = l1 * wa1 + ol1 * wb1;
    m0 = nm0; m1 = nm1;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      const float2 x0 = *reinterpret_cast<const float2*>(xch + g * (DH + 2) + n * 8 + t * 2);
      const float2 x1 = *reinterpret_cast<const float2*>(xch + (g + 8) * (DH + 2) + n * 8 + t * 2);
      o[n][0] = o[n][0] * wa0 + x0.x * wb0; o[n][1] = o[n][1] * wa0 + x0.y * wb0;
      o[n][2] = o[n][2] * wa1 + x1.x * wb1; o[n][3] = o[n][3] * wa1 + x1.y * wb1;
    }
  }
  // write partials
  const size_t base = ((size_t(split) * p.rows + b) * p.H + hq) * size_t(p.W);
  if (r0 < p.W) {
    float* po = p.part_o + (base + r0) * DH;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) *reinterpret_cast<float2*>(po + n * 8 + t * 2) = make_float2(o[n][0], o[n][1]);
    if (t == 0) { p.part_ml[(base + r0) * 2] = m0; p.part_ml[(base + r0) * 2 + 1] = l0; }
  }
  if (r1 < p.W) {
    float* po = p.part_o + (base + r1) * DH;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) *reinterpret_cast<float2*>(po + n * 8 + t * 2) = make_float2(o[n][2], o[n][3]);
    if (t == 0) { p.part_ml[(base + r1) * 2] = m1; p.part_ml[(base + r1) * 2 + 1] = l1; }
  }
}

// One warp per (row b, head, query i): merge chunk partials, normalise, write bf16.  Loads are issued in batches of
// kCombBatch chunks (clamped indices, surplus weights zero) so the L2 round trips overlap.
constexpr int kCombBatch = 8;

// what the combine needs, small enough to ride in the GEMM chain kernel's parameters
struct AttnCombine {
  const float* part_o;    // [splits][rows][H][W][Dh]
  const float* part_ml;   // [splits][rows][H][W][2]
  __nv_bfloat16* out;     // [rows*W][H*Dh]
  int rows, W, H, n_chunks, head_dim;
  int sparse;             // 1: most slots hold the empty marker {-inf, 0} (attention_sw.cu writes one partial per SEGMENT):
                          //    read the 8-byte {m, sum} pairs first and fetch only the live partial rows
};

// one warp: merge the split partials of row `widx` = ((b * H + h) * W + i), normalise, write bf16
template <int DH>
__device__ __forceinline__ void attn_combine_row(const AttnCombine& p, int widx, int lane) {
  const int i = widx % p.W, h = (widx / p.W) % p.H, b = widx / (p.W * p.H);
  const size_t stride = size_t(p.rows) * p.H * p.W;  // per chunk
  const size_t idx = (size_t(b) * p.H + h) * p.W + i;
  constexpr int PER = DH / 32;
  float acc[PER];
#pragma unroll
  for (int e = 0; e < PER; ++e) acc[e] = 0.f;
  float mrun = -INFINITY, lsum = 0.f;
  if (p.sparse) {
    constexpr int SB = 4;   // live partials merged per round trip
    for (int cb = 0; cb < p.n_chunks; cb += 32) {
      const int c = cb + lane;
      float2 mine = make_float2(-INFINITY, 0.f);
      if (c < p.n_chunks) mine = __ldcg(reinterpret_cast<const float2*>(p.part_ml + (c * stride + idx) * 2));
      unsigned live = __ballot_sync(0xffffffffu, mine.x != -INFINITY);   // ascending slot order = ascending key order
      while (live) {
        int src[SB];
        bool on[SB];
        const int first = __ffs(live) - 1;
#pragma unroll
        for (int k = 0; k < SB; ++k) {
          on[k] = live != 0;
          src[k] = on[k] ? __ffs(live) - 1 : first;
          if (on[k]) live &= live - 1;
        }
        float2 ml[SB];
        float po[SB][PER];
#pragma unroll
        for (int k = 0; k < SB; ++k) {
          ml[k].x = __shfl_sync(0xffffffffu, mine.x, src[k]);
          ml[k].y = __shfl_sync(0xffffffffu, mine.y, src[k]);
#pragma unroll
          for (int e = 0; e < PER; ++e) po[k][e] = __ldcg(p.part_o + ((cb + src[k]) * stride + idx) * DH + lane + 32 * e);
        }
        float mb = mrun;
#pragma unroll
        for (int k = 0; k < SB; ++k)
          if (on[k]) mb = fmaxf(mb, ml[k].x);
        const float rescale = exp2f(mrun - mb);   // mrun == -inf -> 0
        lsum *= rescale;
#pragma unroll
        for (int e = 0; e < PER; ++e) acc[e] *= rescale;
        mrun = mb;
#pragma unroll
        for (int k = 0; k < SB; ++k) {
          if (on[k]) {
            const float w = exp2f(ml[k].x - mb);
            lsum += w * ml[k].y;
#pragma unroll
            for (int e = 0; e < PER; ++e) acc[e] += w * po[k][e];
          }
        }
      }
    }
  } else
  for (int c0 = 0; c0 < p.n_chunks; c0 += kCombBatch) {
    float2 ml[kCombBatch];
    float po[kCombBatch][PER];
#pragma unroll
    for (int k = 0; k < kCombBatch; ++k) {
      const int c = min(c0 + k, p.n_chunks - 1);
      ml[k] = __ldcg(reinterpret_cast<const float2*>(p.part_ml + (c * stride + idx) * 2));
#pragma unroll
      for (int e = 0; e < PER; ++e) po[k][e] = __ldcg(p.part_o + (c * stride + idx) * DH + lane + 32 * e);
    }
    float mb = mrun;
#pragma unroll
    for (int k = 0; k < kCombBatch; ++k)
      if (c0 + k < p.n_chunks) mb = fmaxf(mb, ml[k].x);
    if (mb == -INFINITY) continue;
    const float rescale = exp2f(mrun - mb);   // mrun == -inf -> 0
    lsum *= rescale;
#pragma unroll
    for (int e = 0; e < PER; ++e) acc[e] *= rescale;
    mrun = mb;
#pragma unroll
    for (int k = 0; k < kCombBatch; ++k) {
      if (c0 + k < p.n_chunks && ml[k].x != -INFINITY) {
        const float w = exp2f(ml[k].x - mb);
        lsum += w * ml[k].y;
#pragma unroll
        for (int e = 0; e < PER; ++e) acc[e] += w * po[k][e];
      }
    }
  }
  const float inv = lsum > 0.f ? 1.f / lsum : 0.f;  // fully masked query (CFG hidden prefix) -> 0
  __nv_bfloat16* dst = p.out + (size_t(b * p.W + i) * p.H + h) * DH;
#pragma unroll
  for (int e = 0; e < PER; ++e) dst[lane + 32 * e] = __float2bfloat16_rn(acc[e] * inv);
}

// stand-alone combine kernel (used when no GEMM chain follows the attention)
template <int DH>
__global__ void __launch_bounds__(256) attn_combine_kernel(AttnCombine c) {
  pdl_wait();
  pdl_launch_dependents();
  const int widx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (widx >= c.rows * c.H * c.W) return;
  attn_combine_row<DH>(c, widx, threadIdx.x & 31);
}

// Chooses the key split: as many CTAs as fit in ONE wave of three per SM, spans in whole 64-key sub-chunks.
// (stages 3: two CTAs of 104 KB per SM instead of three of 70 KB)
int attn_stages() {
  static int st = 0;
  if (!st) {
    st = 3;   // measured (profiles/r02e_attn_stages.txt): 21.1 vs 22.4 us per layer at W = 32, 1 200 keys
    if (const char* e = getenv("SJD_ATTN_STAGES")) st = atoi(e) == 2 ? 2 : 3;
  }
  return st;
}

void attn_plan(AttnParams* p, int sm_count) {
  const int T = p->kv_len + p->W;
  const int G = p->H / p->Hkv;
  const int n_tiles = G * ((p->W + 15) / 16);
  const int tgroups = (n_tiles + kAttnMaxTiles - 1) / kAttnMaxTiles;
  const int base_ctas = p->rows * p->Hkv * tgroups;
  const int per_sm = (attn_stages() == 3 && n_tiles <= 2) ? 2 : 3;
  int want = (per_sm * sm_count) / base_ctas;   // splits wanted: never more CTAs than the resident slots (no second wave)
  const int n_sub = (T + kAttnSub - 1) / kAttnSub;
  if (want > n_sub) want = n_sub;
  if (want < 1) want = 1;
  const int sub_per_split = (n_sub + want - 1) / want;
  p->span = sub_per_split * kAttnSub;
  p->n_chunks = (T + p->span - 1) / p->span;
}

AttnCombine attn_combine_desc(const AttnParams& p, int head_dim) {
  AttnCombine c;
  c.part_o = p.part_o; c.part_ml = p.part_ml; c.out = p.out;
  c.rows = p.rows; c.W = p.W; c.H = p.H; c.n_chunks = p.n_chunks; c.head_dim = head_dim;
  c.sparse = 0;
  return c;
}

// with_combine = false: the caller folds the split merge into the GEMM chain kernel that follows (its epilogue
// warps are idle at that point) instead of paying a kernel for it
int attn_launch(const AttnParams& p, int head_dim, bool with_combine, cudaStream_t stream) {
  const int G = p.H / p.Hkv;
  const int n_tiles = G * ((p.W + 15) / 16);
  const int tiles_here = n_tiles < kAttnMaxTiles ? n_tiles : kAttnMaxTiles;
  const int nwarps = 2 * tiles_here;
  const int tgroups = (n_tiles + tiles_here - 1) / tiles_here;
  dim3 grid(p.n_chunks, p.Hkv * tgroups, p.rows);
  const int total = p.rows * p.H * p.W;
  dim3 cgrid((total + 7) / 8);
  int rc = 0;
  // thread count = 64 * tiles per CTA; the register cap is chosen so that three 64/128-thread CTAs share an SM
#define SJD_ATTN_LAUNCH(DH_, NT_, MINB_)                                                                          \
  do {                                                                                                            \
    constexpr int smem = 4 * kAttnSub * (DH_ * 2 + 16);                                                           \
    static bool set = false;                                                                                      \
    if (!set) {                                                                                                   \
      cudaFuncSetAttribute(attn_window_kernel<DH_, NT_, MINB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
      set = true;                                                                                                 \
    }                                                                                                             \
    rc |= launch_pdl(attn_window_kernel<DH_, NT_, MINB_>, grid, dim3(NT_), smem, stream, p);                      \
  } while (0)
  const int nt = nwarps * 32;
#define SJD_ATTN_LAUNCH3(DH_, NT_)                                                                                 \
  do {                                                                                                            \
    constexpr int smem = 6 * kAttnSub * (DH_ * 2 + 16);                                                           \
    static bool set = false;                                                                                      \
    if (!set) {                                                                                                   \
      cudaFuncSetAttribute(attn_window_kernel<DH_, NT_, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
      set = true;                                                                                                 \
    }                                                                                                             \
    rc |= launch_pdl(attn_window_kernel<DH_, NT_, 2, 3>, grid, dim3(NT_), smem, stream, p);                       \
  } while (0)
  if (attn_stages() == 3 && head_dim == 128 && n_tiles <= 2) {   // small windows of an MHA model: deeper ring, 2 CTAs / SM
    if (nt == 64) SJD_ATTN_LAUNCH3(128, 64);
    else SJD_ATTN_LAUNCH3(128, 128);
    if (with_combine) rc |= launch_pdl(attn_combine_kernel<128>, cgrid, dim3(256), 0, stream, attn_combine_desc(p, 128));
    return rc;
  }
  if (head_dim == 128) {
    if (nt == 64) SJD_ATTN_LAUNCH(128, 64, 3);
    else if (nt == 128) SJD_ATTN_LAUNCH(128, 128, 3);
    else if (nt == 192) SJD_ATTN_LAUNCH(128, 192, 1);
    else SJD_ATTN_LAUNCH(128, 256, 1);
    if (with_combine) rc |= launch_pdl(attn_combine_kernel<128>, cgrid, dim3(256), 0, stream, attn_combine_desc(p, 128));
  } else if (head_dim == 64) {
    if (nt == 64) SJD_ATTN_LAUNCH(64, 64, 3);
    else if (nt == 128) SJD_ATTN_LAUNCH(64, 128, 3);
    else if (nt == 192) SJD_ATTN_LAUNCH(64, 192, 1);
    else SJD_ATTN_LAUNCH(64, 256, 1);
    if (with_combine) rc |= launch_pdl(attn_combine_kernel<64>, cgrid, dim3(256), 0, stream, attn_combine_desc(p, 64));
  } else {
    return -3;
  }
#undef SJD_ATTN_LAUNCH
#undef SJD_ATTN_LAUNCH3
  return rc;
}

}  // namespace sjd
