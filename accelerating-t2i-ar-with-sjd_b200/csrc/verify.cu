// Speculative-Jacobi verify step: everything the reference does between "logits are back" and
// "how many draft tokens survived" — CFG mix, grammar mask, top-k, softmax, sampling, probabilistic
// accept/reject scan, residual resample, prefix match — as ONE launch with no host round trip and
// no Python loop (the CTA that finishes its window position last runs the accept scan).  Restates, on device:
//   sampling_logits2tokens                 scheduler/jacobi_iteration_lumina_mgpt.py:82-132
//   MultiTokensVLLogitsProcessor           scheduler/logit_processor_3dim.py:45-155   (as allow-range + forced ids)
//   MultiTokensInterleavedTopKLogitsWarper scheduler/logit_processor_3dim.py:158-204  (scores < k-th largest removed)
//   SpeculativeSampler.__call__            scheduler/jacobi_iteration_lumina_mgpt.py:247-315
//   reject_sampling_single_token           :209-241, get_reject_sampling_logits :203-207
//   find_first_misaligned_token_inds       :317-333 ('jacobi' scheme)
//   TopPLogitsWarper3d                     scheduler/logit_processor_3dim.py:355-419  (block_top_p, after top-k)
// torch.multinomial(p, 1) is argmax(p / Exp(1)) and torch.rand feeds the accept test; the noise tensors are
// produced by the caller with the same torch.Generator calls the reference makes, so token streams are
// reproducible against it.  The arithmetic deliberately avoids FMA contraction where the reference rounds
// between operations (CFG mix), and uses IEEE division like torch.
#include "common.cuh"

namespace sjd {

constexpr int kVerifyThreads = 1024;
constexpr int kHistBins = 2048;

struct VerifyParams {
  const float* logits;   // [(has_uncond ? 2 : 1) * W][V]; cond rows first
  int W, V;
  int has_uncond;        // logits carry the CFG-uncond rows
  int apply_cfg;         // mix g*(c-u)+u (else cond rows only: CFG disabled outside an image)
  float guidance;
  float temperature;     // scores / temperature when != 1
  int allow_lo, allow_hi;  // ids outside [lo,hi) -> -inf; disabled when hi <= lo
  int allow_mode;          // 0: [lo,hi) if hi > lo (ban ignored); 1: same minus ban[]; 2: complement of [lo,hi) minus ban[];
                           // 3: only the ids ban[0], ban[1]; 4: everything minus ban[]
  int ban[2];              // removed ids (-1 = none); mode 3: the kept ids
  int resid_set;           // 1: the residual (positions whose forced_resid is -1) uses the candidate set below instead
  int resid_allow_mode, resid_allow_lo, resid_allow_hi, resid_ban[2], resid_from;   // ... from reject position resid_from on
  int* done_flag;          // optional (mapped host memory): receives done_seq once the result is written (system-scope fence)
  int done_seq;
  const int* forced;     // [W] forced token id per window position, or -1
  const int* forced_resid;  // [W] forced id of the residual distribution at reject position j; null = forced
  int top_k;             // 0 = off
  float top_p_thresh;    // float32(1 - top_p): ascending running probability <= this is removed; 0 = off
  int do_sample;         // 0 = argmax
  int scheme;            // 0 = speculative_jacobi, 1 = jacobi
  const int* draft;      // [W] window ids ([0] = last accepted token)
  const int* q_row;      // [W] row of p_prev holding the draft distribution, -1 = one-hot at draft[i]
  const float* p_prev;   // [Wcap][V]
  float* p_cur;          // [Wcap][V] out: probabilities of this trip
  const float* noise_e1; // [W][V]  Exp(1) noise of torch.multinomial
  const float* noise_u;  // [W]     torch.rand values at (i, draft[i])
  const float* noise_e2; // [V]     Exp(1) noise of the residual multinomial
  int eoi_token;         // if an accepted draft equals this id the residual is processed in text mode
  int text_top_k;
  float* resid;          // [V] scratch
  int* next_tokens;      // [W] scratch: tokens sampled from p_cur
  int* out_tokens;       // [W] tokens after accept / resample
  int* out_info;         // [4] matched, rejected(0/1), first_reject, reserved
  unsigned int* sync_ws; // one zero-initialised word: lets the last row CTA run the accept scan in the same launch
  // device-side noise (rng_mode = 1): the kernel computes, for exactly the elements it consumes, the values torch's CUDA
  // generator would have written into the three noise tensors (Philox4x32-10, torch's element -> (thread, counter) map)
  int rng_mode;
  unsigned long long rng_seed;
  unsigned long long rng_off[3];   // philox offset at the start of the e1 / u / e2 draw
  unsigned int rng_span[3];        // blockDim * gridDim of torch's launch for that draw (its grid-stride)
};

// ---- torch-compatible Philox noise ------------------------------------------------------------------------------
// torch.Tensor.exponential_ / torch.rand on a CUDA generator run distribution_elementwise_grid_stride_kernel
// (ATen/native/cuda/DistributionTemplates.h): thread t = blockIdx * 256 + threadIdx calls curand_init(seed, t, offset)
// and, in grid-stride iteration k, one curand_uniform4 whose four values go to the elements t + span * (4k + j),
// j = 0..3, span = 256 * grid.  So element e sits at thread e % span, iteration (e / span) / 4, lane (e / span) % 4, and
// its value is Philox4x32-10(counter = {offset / 4 + iteration, subsequence = thread}, key = seed)[lane].
__device__ __noinline__ uint4 philox4x32_10(uint4 c, uint2 k) {   // one copy: the callers sit in unrolled loops
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// curand_uniform: (0, 1]
__device__ __forceinline__ float torch_uniform_raw(unsigned long long seed, unsigned long long offset, uint32_t span,
                                                   unsigned long long e) {
  const unsigned long long q = e / span;
  const uint32_t t = uint32_t(e - q * span), lane = uint32_t(q & 3ull);
  const unsigned long long ctr = (offset >> 2) + (q >> 2);
  const uint4 r = philox4x32_10(make_uint4(uint32_t(ctr), uint32_t(ctr >> 32), t, 0u),
                                make_uint2(uint32_t(seed), uint32_t(seed >> 32)));
  const uint32_t x = lane == 0 ? r.x : (lane == 1 ? r.y : (lane == 2 ? r.z : r.w));
  return __fmaf_rn(float(x), 2.3283064e-10f, 2.3283064e-10f / 2.0f);
}
// torch.rand: reverse_bounds maps 1.0 to 0.0 (uniform_and_transform, from = 0, to = 1)
__device__ __forceinline__ float torch_rand(unsigned long long seed, unsigned long long offset, uint32_t span,
                                            unsigned long long e) {
  const float u = torch_uniform_raw(seed, offset, span, e);
  return u == 1.0f ? 0.0f : u;
}
// exponential_(1): -log(u), with u >= 1 - eps/2 mapped to eps/2 (transformation::exponential, TransformationHelper.h).
// torch's CUDA build evaluates that log with the FAST intrinsic (__logf = lg2.approx * ln 2) — pinned on the B200 against
// torch 2.11: with logf 85 % of the elements differ in the last bit or two, with __logf none does
// (tests/test_gpu_parity.py::test_device_philox_matches_torch_generator, profiles/r02c notes).
__device__ __forceinline__ float torch_exponential(unsigned long long seed, unsigned long long offset, uint32_t span,
                                                   unsigned long long e) {
  const float u = torch_uniform_raw(seed, offset, span, e);
  const float lg = (u >= 1.0f - 1.1920929e-07f / 2.0f) ? -1.1920929e-07f / 2.0f : __logf(u);
  return -1.0f * lg;
}

// The Exp(1) noise of one multinomial row: either a tensor the caller filled, or computed on the fly.
struct NoiseRow {
  const float* ptr;               // rng_mode 0
  unsigned long long seed, off, base;   // rng_mode 1: element index = base + v
  uint32_t span;
  __device__ __forceinline__ float operator[](int v) const {
    return ptr ? ptr[v] : torch_exponential(seed, off, span, base + uint32_t(v));
  }
};
__device__ __forceinline__ NoiseRow noise_row_e1(const struct VerifyParams& p, int i);
__device__ __forceinline__ NoiseRow noise_row_e2(const struct VerifyParams& p);

__device__ __forceinline__ uint32_t f2key(float f) {
  if (f == 0.f) return 0x80000000u;  // -0 == +0 for "<"
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// The grammar's candidate set of a window (the Chameleon / Anole processors of logit_processor_3dim.py:207-353 reduce to
// one of these per call): [lo, hi) (mode 1, also Lumina / Emu3), its complement minus two ids (mode 2: text between
// images), a set of one or two ids (mode 3), or everything minus two ids (mode 0).
struct Cand {
  int kind;   // 0: everything, 1: [lo, hi), 2: complement of [lo, hi), 3: only b0 / b1; kinds 0-2 also drop b0, b1
  int lo, hi, b0, b1, span_lo, span_hi;
  __device__ __forceinline__ bool ok(int v) const {
    if (kind == 3) return v == b0 || v == b1;
    if (v == b0 || v == b1) return false;
    const bool in = v >= lo && v < hi;
    return kind == 1 ? in : (kind == 2 ? !in : true);
  }
};
__device__ __forceinline__ Cand make_cand_from(int m, int alo, int ahi, int b0, int b1, int V, bool text_mode) {
  Cand c;
  const bool ranged = !text_mode && (ahi > alo);
  if (text_mode) m = 0;
  c.kind = m == 2 ? 2 : (m == 3 ? 3 : ((m == 4 || !ranged) ? 0 : 1));
  c.lo = ranged ? max(alo, 0) : 0;
  c.hi = ranged ? min(ahi, V) : V;
  const bool bans = !text_mode && m >= 1;     // mode 0 = the zero-initialised legacy form: ban[] is not looked at
  c.b0 = bans ? b0 : -1;
  c.b1 = bans ? b1 : -1;
  c.span_lo = c.kind == 1 ? c.lo : 0;
  c.span_hi = c.kind == 1 ? c.hi : V;
  return c;
}
__device__ __forceinline__ Cand make_cand(const VerifyParams& p, bool text_mode) {
  return make_cand_from(p.allow_mode, p.allow_lo, p.allow_hi, p.ban[0], p.ban[1], p.V, text_mode);
}
// the residual's candidate set: the window's, unless the caller gave another one (a grammar whose decision changes with
// the drafts accepted inside the window, e.g. Anole right after a forced begin- / end-of-image)
__device__ __forceinline__ Cand make_cand_resid(const VerifyParams& p, bool text_mode, int j) {
  if (p.resid_set && !text_mode && j >= p.resid_from)
    return make_cand_from(p.resid_allow_mode, p.resid_allow_lo, p.resid_allow_hi, p.resid_ban[0], p.resid_ban[1], p.V, false);
  return make_cand(p, text_mode);
}

__device__ __forceinline__ NoiseRow noise_row_e1(const VerifyParams& p, int i) {
  NoiseRow n;
  n.ptr = p.rng_mode ? nullptr : p.noise_e1 + size_t(i) * p.V;
  n.seed = p.rng_seed; n.off = p.rng_off[0]; n.span = p.rng_span[0];
  n.base = (unsigned long long)(i) * (unsigned long long)(p.V);
  return n;
}
__device__ __forceinline__ NoiseRow noise_row_e2(const VerifyParams& p) {
  NoiseRow n;
  n.ptr = p.rng_mode ? nullptr : p.noise_e2;
  n.seed = p.rng_seed; n.off = p.rng_off[2]; n.span = p.rng_span[2];
  n.base = 0ull;
  return n;
}

struct BlockScratch {
  uint32_t hist[kHistBins];
  float redf[32];
  int redi[32];
  uint32_t sel_bin;
  uint32_t sel_above;
};


// Radix-select step shared by the two k-th-key routines: given hist[0..nb) (nb = 1024 or 2048 <= 2 * blockDim), find the
// LARGEST bin b whose suffix count sum_{i >= b} hist[i] reaches `remaining`; sel_bin = b, sel_above = sum_{i > b} hist[i].
// Every thread owns nb / blockDim consecutive bins in DESCENDING order; one block-wide exclusive scan of the per-thread
// sums (warp shuffles + 32 warp totals) replaces the serial walk a single warp used to do while 31 others waited at the
// barrier (30 % of the kernel's stall samples in the round-2 ncu capture, profiles/r02d_full_verify_summary.csv).
// Integer arithmetic: the result does not depend on the order of anything.
__device__ __forceinline__ void block_select_bin(uint32_t nb, uint32_t remaining, BlockScratch& sc) {
  const uint32_t per = (nb + blockDim.x - 1) / blockDim.x;        // 1 or 2 (kVerifyThreads = 1024)
  const uint32_t t = threadIdx.x, lane = t & 31u, wid = t >> 5;
  uint32_t local = 0;
  for (uint32_t j = 0; j < per; ++j) {
    const uint32_t d = t * per + j;                                // d-th bin from the top
    if (d < nb) local += sc.hist[nb - 1 - d];
  }
  uint32_t incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= uint32_t(o)) incl += v;
  }
  if (lane == 31) sc.redi[wid] = int(incl);
  __syncthreads();
  uint32_t wbase = 0;
  {
    const uint32_t nw = blockDim.x >> 5;
    uint32_t wt = lane < nw ? uint32_t(sc.redi[lane]) : 0u, wincl = wt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, wincl, o);
      if (lane >= uint32_t(o)) wincl += v;
    }
    wbase = __shfl_sync(0xffffffffu, wincl - wt, int(wid));        // exclusive total of the warps above this one
  }
  const uint32_t excl = wbase + incl - local;
  if (excl < remaining && excl + local >= remaining) {             // exactly one thread
    uint32_t run = excl;
    for (uint32_t j = 0; j < per; ++j) {
      const uint32_t d = t * per + j;
      const uint32_t c = d < nb ? sc.hist[nb - 1 - d] : 0u;
      if (run + c >= remaining) {
        sc.sel_bin = nb - 1 - d;
        sc.sel_above = run;
        break;
      }
      run += c;
    }
  }
  __syncthreads();
}

// k-th largest key among the finite entries of row[0..V); caller guarantees k <= #finite.
__device__ uint32_t block_kth_key(const float* __restrict__ row, int vb, int V, int k, BlockScratch& sc) {
  uint32_t prefix = 0, mask = 0;
  uint32_t remaining = uint32_t(k);
  const int shifts[3] = {21, 10, 0};
  const int nbits[3] = {11, 11, 10};
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = shifts[pass];
    const uint32_t nb = 1u << nbits[pass];
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) sc.hist[i] = 0;
    __syncthreads();
    for (int v = vb + threadIdx.x; v < V; v += blockDim.x) {
      const float f = row[v];
      if (f == -INFINITY) continue;
      const uint32_t key = f2key(f);
      if ((key & mask) == prefix) atomicAdd(&sc.hist[(key >> shift) & (nb - 1)], 1u);
    }
    __syncthreads();
    block_select_bin(nb, remaining, sc);
    prefix |= sc.sel_bin << shift;
    mask |= (nb - 1) << shift;
    remaining -= sc.sel_above;
    __syncthreads();
  }
  return prefix;
}

__device__ __forceinline__ float block_max(float v, BlockScratch& sc) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sc.redf[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x & 31) < (blockDim.x >> 5) ? sc.redf[threadIdx.x & 31] : -INFINITY;
  return warp_max(t);
}
__device__ __forceinline__ float block_sumf(float v, BlockScratch& sc) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sc.redf[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x & 31) < (blockDim.x >> 5) ? sc.redf[threadIdx.x & 31] : 0.f;
  return warp_sum(t);
}
__device__ __forceinline__ int block_sumi(int v, BlockScratch& sc) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sc.redi[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = (threadIdx.x & 31) < (blockDim.x >> 5) ? sc.redi[threadIdx.x & 31] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}
// argmax with lowest-index tie break
__device__ __forceinline__ int block_argmax(float v, int idx, BlockScratch& sc) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { sc.redf[threadIdx.x >> 5] = v; sc.redi[threadIdx.x >> 5] = idx; }
  __syncthreads();
  const bool in = (threadIdx.x & 31) < (blockDim.x >> 5);
  v = in ? sc.redf[threadIdx.x & 31] : -INFINITY;
  idx = in ? sc.redi[threadIdx.x & 31] : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  return idx;
}

// Shared tail: row[] holds processed scores (grammar applied). Applies top-k, softmax (in place -> probabilities)
// and draws the token.  Returns the token (valid in every thread).  Only ids [vb, V) are visited: the caller passes the
// grammar's candidate range (vb a multiple of blockDim, so that every thread walks the ids it would walk from 0, in
// the same order — the block reductions, and with them every result, do not depend on vb) and has made everything
// outside it -inf / probability 0.  `Vfull` is the vocabulary size (top-k is a no-op when k >= Vfull).
__device__ int block_topk_softmax_sample(float* __restrict__ row, int vb, int V, int Vfull, int top_k, int do_sample,
                                         const NoiseRow noise_e, BlockScratch& sc) {
  float mx = -INFINITY;
  int nfin = 0;
  for (int v = vb + threadIdx.x; v < V; v += blockDim.x) {
    const float f = row[v];
    mx = fmaxf(mx, f);
    nfin += (f != -INFINITY);
  }
  mx = block_max(mx, sc);
  nfin = block_sumi(nfin, sc);
  float thr = -INFINITY;  // scores < thr are removed
  if (top_k > 0 && top_k < Vfull && nfin > top_k) {
    const uint32_t key = block_kth_key(row, vb, V, top_k, sc);
    const uint32_t u = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
    thr = __uint_as_float(u);
  }
  float sum = 0.f;
  for (int v = vb + threadIdx.x; v < V; v += blockDim.x) {
    const float f = row[v];
    if (f >= thr && f != -INFINITY) sum += expf(f - mx);
  }
  sum = block_sumf(sum, sc);
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int v = vb + threadIdx.x; v < V; v += blockDim.x) {
    const float f = row[v];
    const bool keep = (f >= thr && f != -INFINITY);
    const float pr = keep ? expf(f - mx) / sum : 0.f;
    row[v] = pr;
    float val;
    if (do_sample) val = pr > 0.f ? pr / noise_e[v] : 0.f;   // 0 / noise = 0: no need to draw it
    else val = keep ? f : -INFINITY;
    if (val > best) { best = val; besti = v; }  // ascending v per thread keeps the lowest index on ties
  }
  return block_argmax(best, besti, sc);
}

// ---- register-resident variant --------------------------------------------------------------------------
// When the grammar (or the vocabulary) leaves at most kVerifyThreads * VPT candidate ids [lo, lo + n), each thread
// keeps its VPT processed scores in registers and every pass (max, radix select, sum, sample) runs on them: the
// logits are read once and the probabilities written once.  Element j of thread t is id v0 + j*blockDim + t with v0
// = lo rounded down to a multiple of blockDim, i.e. exactly the ids (in the same ascending order) the strided
// global-memory variant gives that thread, so the block reductions — and therefore every result — are bit-identical.
template <int VPT>
__device__ uint32_t block_kth_key_regs(const float (&s)[VPT], int k, BlockScratch& sc) {
  uint32_t prefix = 0, mask = 0;
  uint32_t remaining = uint32_t(k);
  const int shifts[3] = {21, 10, 0};
  const int nbits[3] = {11, 11, 10};
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = shifts[pass];
    const uint32_t nb = 1u << nbits[pass];
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) sc.hist[i] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const float f = s[j];
      if (f != -INFINITY) {
        const uint32_t key = f2key(f);
        if ((key & mask) == prefix) atomicAdd(&sc.hist[(key >> shift) & (nb - 1)], 1u);
      }
    }
    __syncthreads();
    block_select_bin(nb, remaining, sc);
    prefix |= sc.sel_bin << shift;
    mask |= (nb - 1) << shift;
    remaining -= sc.sel_above;
    __syncthreads();
  }
  return prefix;
}

// s[j]: processed score of id v0 + j*blockDim + tid (-inf when removed or outside [lo, hi)).  Writes the
// probabilities of ids [lo, hi) to p_out (the caller zeroes the rest of the row) and returns the token.
template <int VPT>
__device__ int block_topk_softmax_sample_regs(float (&s)[VPT], int v0, int lo, int hi, int top_k, int do_sample,
                                              const NoiseRow noise_e, float* __restrict__ p_out,
                                              int V, BlockScratch& sc) {
  float mx = -INFINITY;
  int nfin = 0;
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    mx = fmaxf(mx, s[j]);
    nfin += (s[j] != -INFINITY);
  }
  mx = block_max(mx, sc);
  nfin = block_sumi(nfin, sc);
  float thr = -INFINITY;  // scores < thr are removed
  if (top_k > 0 && top_k < V && nfin > top_k) {
    const uint32_t key = block_kth_key_regs<VPT>(s, top_k, sc);
    const uint32_t u = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
    thr = __uint_as_float(u);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < VPT; ++j)
    if (s[j] >= thr && s[j] != -INFINITY) sum += expf(s[j] - mx);
  sum = block_sumf(sum, sc);
  float best = -INFINITY;
  int besti = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    const int v = v0 + j * blockDim.x + threadIdx.x;
    if (v >= lo && v < hi) {
      const float f = s[j];
      const bool keep = (f >= thr && f != -INFINITY);
      const float pr = keep ? expf(f - mx) / sum : 0.f;
      if (p_out) p_out[v] = pr;
      float val;
      if (do_sample) val = pr > 0.f ? pr / noise_e[v] : 0.f;   // 0 / noise = 0: no need to draw it
      else val = keep ? f : -INFINITY;
      if (val > best) { best = val; besti = v; }  // ascending v per thread keeps the lowest index on ties
    }
  }
  // ids outside the candidate range have probability 0: in sampling mode 0 / noise = 0 can still win the argmax
  // when every candidate probability is 0 too — impossible here (the kept maximum has probability > 0)
  return block_argmax(best, besti, sc);
}


// ---- top-p (nucleus) ---------------------------------------------------------------------------------------
// TopPLogitsWarper3d (scheduler/logit_processor_3dim.py:406-419) sorts the scores ascending, takes the running sum of
// their softmax and removes every entry whose running sum is <= 1 - top_p; the largest entry always stays.  Here the
// row already holds that softmax (probabilities after top-k), so the removed set is "the smallest probabilities whose
// sum stays <= thresh": a 3-pass radix descent over the float bits of the probability with a histogram of SUMS
// finds the boundary value; sums are accumulated in 2^-44 fixed point (integer adds commute, so the result does
// not depend on the order threads arrive in).  Entries tied with the boundary value are removed lowest id first,
// as many as still fit (what a stable ascending sort gives).  The survivors are renormalised and the token redrawn.
constexpr float kFx = 17592186044416.f;   // 2^44

struct TopPScratch {
  unsigned long long hsum[kHistBins];
  unsigned long long below;   // fixed-point sum of everything under the current prefix
  uint32_t sel_bin;           // selected bin, or 0xffffffff: nothing crosses the threshold
  uint32_t warp_cnt[32];
  uint32_t base;
};

__device__ int block_top_p(float* __restrict__ row, int V, float thresh, int do_sample, int greedy_tok,
                           const NoiseRow noise_e, BlockScratch& sc, TopPScratch& tp) {
  const unsigned long long thresh_fx = __float2ull_rd(thresh * kFx);
  uint32_t prefix = 0, mask = 0;
  const int shifts[3] = {21, 10, 0};
  const int nbits[3] = {11, 11, 10};
  if (threadIdx.x == 0) { tp.below = 0; tp.sel_bin = 0; }
  bool none = false;
  for (int pass = 0; pass < 3 && !none; ++pass) {
    const int shift = shifts[pass];
    const uint32_t nb = 1u << nbits[pass];
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) tp.hsum[i] = 0ull;
    __syncthreads();
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      const float f = row[v];
      if (f > 0.f) {
        const uint32_t key = __float_as_uint(f);
        if ((key & mask) == prefix) atomicAdd(&tp.hsum[(key >> shift) & (nb - 1)], __float2ull_rn(f * kFx));
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      const uint32_t lane = threadIdx.x, per = nb / 32;
      const unsigned long long below = tp.below;
      unsigned long long csum = 0;
      for (uint32_t j = 0; j < per; ++j) csum += tp.hsum[lane * per + j];   // lane 0 = smallest keys
      unsigned long long incl = csum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= uint32_t(o)) incl += t;
      }
      const unsigned long long excl = incl - csum;
      const bool mine = (below + excl <= thresh_fx) && (below + incl > thresh_fx);
      const uint32_t any = __ballot_sync(0xffffffffu, mine);
      if (any == 0u) {
        if (lane == 0) tp.sel_bin = 0xffffffffu;
      } else if (mine) {
        unsigned long long run = below + excl;
        for (uint32_t j = 0; j < per; ++j) {
          const unsigned long long c = tp.hsum[lane * per + j];
          if (run + c > thresh_fx) {
            tp.sel_bin = lane * per + j;
            tp.below = run;
            break;
          }
          run += c;
        }
      }
    }
    __syncthreads();
    if (tp.sel_bin == 0xffffffffu) none = true;
    else {
      prefix |= tp.sel_bin << shift;
      mask |= (nb - 1) << shift;
    }
    __syncthreads();
  }
  float sum = 0.f;
  if (none) {
    // everything fits under the threshold (top_p ~ 0): only the last entry of the ascending order stays — the
    // largest probability, highest id among equals
    float mx = 0.f;
    for (int v = threadIdx.x; v < V; v += blockDim.x) mx = fmaxf(mx, row[v]);
    mx = block_max(mx, sc);
    int hi = -1;
    for (int v = threadIdx.x; v < V; v += blockDim.x)
      if (row[v] == mx) hi = v;
    __syncthreads();
    hi = -block_argmax(float(hi), -hi, sc);   // largest id: argmax of the id itself (exact in float for V < 2^24)
    for (int v = threadIdx.x; v < V; v += blockDim.x) row[v] = (v == hi) ? 1.f : 0.f;
    __syncthreads();
    return do_sample ? hi : greedy_tok;
  }
  // boundary value `prefix`: tp.hsum[sel_bin] = (#ties) * fx(value); remove the first n_rm ties (ascending id)
  const float bval = __uint_as_float(prefix);
  const unsigned long long fx_b = __float2ull_rn(bval * kFx);
  const unsigned long long room = thresh_fx - tp.below;
  const uint32_t n_rm = fx_b ? uint32_t(room / fx_b) : 0u;
  if (n_rm == 0u) {
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      const float f = row[v];
      const float keep = (f >= bval) ? f : 0.f;
      row[v] = keep;
      sum += keep;
    }
  } else {
    if (threadIdx.x == 0) tp.base = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int v0 = 0; v0 < V; v0 += blockDim.x) {   // ids in ascending order, one slab of blockDim ids at a time
      const int v = v0 + threadIdx.x;
      const float f = v < V ? row[v] : 0.f;
      const bool tie = (f == bval);
      const uint32_t bal = __ballot_sync(0xffffffffu, tie);
      if (lane == 0) tp.warp_cnt[wid] = __popc(bal);
      __syncthreads();
      uint32_t before = tp.base;
      for (uint32_t w2 = 0; w2 < wid; ++w2) before += tp.warp_cnt[w2];
      const uint32_t rank = before + __popc(bal & ((1u << lane) - 1u));
      if (v < V) {
        const float keep = (f > bval || (tie && rank >= n_rm)) ? f : 0.f;
        row[v] = keep;
        sum += keep;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (uint32_t w2 = 0; w2 < nw; ++w2) tot += tp.warp_cnt[w2];
        tp.base += tot;
      }
      __syncthreads();
    }
  }
  sum = block_sumf(sum, sc);
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float pr = row[v] / sum;
    row[v] = pr;
    if (do_sample) {
      const float val = pr > 0.f ? pr / noise_e[v] : 0.f;
      if (val > best) { best = val; besti = v; }
    }
  }
  const int tok = block_argmax(best, besti, sc);
  return do_sample ? tok : greedy_tok;
}

// row[a, b) = 0 with 16-byte stores where the alignment allows (V and the row base are multiples of 4 floats in every
// shipped vocabulary; the scalar head / tail covers the rest)
__device__ __forceinline__ void zero_fill(float* __restrict__ row, int a, int b) {
  if (b <= a) return;
  const uintptr_t addr = reinterpret_cast<uintptr_t>(row + a);
  int head = int(((16 - (addr & 15)) & 15) >> 2);
  if (head > b - a) head = b - a;
  for (int v = a + threadIdx.x; v < a + head; v += blockDim.x) row[v] = 0.f;
  const int a4 = a + head, n4 = (b - a4) >> 2;
  float4* p4 = reinterpret_cast<float4*>(row + a4);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int v = a4 + 4 * n4 + threadIdx.x; v < b; v += blockDim.x) row[v] = 0.f;
}

constexpr int kRegVPT = 16;   // ids per thread held in registers: candidate ranges spanning up to 16 384 ids

// One CTA per window position.
__device__ void verify_row(const VerifyParams& p, int i, BlockScratch& sc, TopPScratch& tp) {
  const int V = p.V;
  const float* c = p.logits + size_t(i) * V;
  const float* u = p.logits + size_t(p.W + i) * V;
  float* row = p.p_cur + size_t(i) * V;
  const int forced = p.forced ? p.forced[i] : -1;
  const bool mix = p.has_uncond && p.apply_cfg;
  if (forced >= 0) {
    // forced position: the distribution is one-hot whatever the logits are (0 -> exp(0)/1 = 1)
    for (int v = threadIdx.x; v < V; v += blockDim.x) row[v] = (v == forced) ? 1.f : 0.f;
    if (threadIdx.x == 0) p.next_tokens[i] = forced;
    return;
  }
  const Cand cd = make_cand(p, false);
  const int lo = cd.span_lo, hi = cd.span_hi;
  const int v0 = (lo / int(blockDim.x)) * int(blockDim.x);
  if (hi - v0 <= int(blockDim.x) * kRegVPT) {
    float s[kRegVPT];
#pragma unroll
    for (int j = 0; j < kRegVPT; ++j) {
      const int v = v0 + j * blockDim.x + threadIdx.x;
      float sv = -INFINITY;
      if (v >= lo && v < hi && cd.ok(v)) {
        sv = c[v];
        if (mix) {
          const float uu = u[v];
          sv = __fadd_rn(__fmul_rn(p.guidance, __fsub_rn(sv, uu)), uu);
        }
        if (p.temperature != 1.f) sv = sv / p.temperature;
      }
      s[j] = sv;
    }
    // everything outside the candidate range has probability zero
    zero_fill(row, 0, lo);
    zero_fill(row, hi, V);
    int tok = block_topk_softmax_sample_regs<kRegVPT>(s, v0, lo, hi, p.top_k, p.do_sample, noise_row_e1(p, i), row, V, sc);
    if (p.top_p_thresh > 0.f) {
      __syncthreads();   // the whole row of probabilities is in global memory
      tok = block_top_p(row, V, p.top_p_thresh, p.do_sample, tok, noise_row_e1(p, i), sc, tp);
    }
    if (threadIdx.x == 0) p.next_tokens[i] = tok;
    return;
  }
  // candidate range too wide for registers (Emu3: 32 768 visual ids of 184 622): same passes over the row in global
  // memory, but only over the blockDim-aligned span that covers [lo, hi); the rest of the row is probability 0
  const int ve = min(V, ((hi + int(blockDim.x) - 1) / int(blockDim.x)) * int(blockDim.x));
  zero_fill(row, 0, v0);
  zero_fill(row, ve, V);
  for (int v = v0 + threadIdx.x; v < ve; v += blockDim.x) {
    float s = -INFINITY;
    if (v >= lo && v < hi && cd.ok(v)) {
      s = c[v];
      if (mix) {
        const float uu = u[v];
        s = __fadd_rn(__fmul_rn(p.guidance, __fsub_rn(s, uu)), uu);
      }
      if (p.temperature != 1.f) s = s / p.temperature;
    }
    row[v] = s;
  }
  __syncthreads();
  int tok = block_topk_softmax_sample(row, v0, ve, V, p.top_k, p.do_sample, noise_row_e1(p, i), sc);
  if (p.top_p_thresh > 0.f) {
    __syncthreads();
    tok = block_top_p(row, V, p.top_p_thresh, p.do_sample, tok, noise_row_e1(p, i), sc, tp);
  }
  if (threadIdx.x == 0) p.next_tokens[i] = tok;
}

// Single CTA: accept scan over the window, prefix match, residual resample at the first rejection.
__device__ void verify_accept(const VerifyParams& p, BlockScratch& sc, TopPScratch& tp) {
  __shared__ int s_first;
  __shared__ int s_text_mode;
  const int W = p.W, V = p.V;
  if (threadIdx.x == 0) { s_first = W; s_text_mode = 0; }
  __syncthreads();
  const int i = threadIdx.x;
  if (i >= 1 && i < W) {
    const int x = p.draft[i];
    bool accept;
    if (p.scheme == 0) {
      const float px = p.p_cur[size_t(i - 1) * V + x];
      const int qr = p.q_row[i];
      const float qx = qr < 0 ? 1.f : p.p_prev[size_t(qr) * V + x];
      const float ui = p.rng_mode ? torch_rand(p.rng_seed, p.rng_off[1], p.rng_span[1],
                                               (unsigned long long)(i) * (unsigned long long)(V) + uint32_t(x))
                                  : p.noise_u[i];
      accept = ui < fminf(px / qx, 1.f);
    } else {
      accept = (x == p.next_tokens[i - 1]);
    }
    if (!accept) atomicMin(&s_first, i);
  }
  __syncthreads();
  const int first = s_first;
  const bool rejected = (p.scheme == 0) && (first < W);
  if (p.scheme == 0) {
    if (i < W) {
      int tok;
      if (i < first - 1) tok = p.draft[i + 1];          // accepted drafts
      else tok = p.next_tokens[i];                      // fresh samples (position first-1 is overwritten below)
      p.out_tokens[i] = tok;
      if (i >= 1 && i <= first - 1 && p.draft[i] == p.eoi_token) s_text_mode = 1;
    }
  } else if (i < W) {
    p.out_tokens[i] = p.next_tokens[i];
  }
  __syncthreads();
  if (rejected) {
    const int j = first - 1;  // window position being resampled
    const float* a = p.p_cur + size_t(j) * V;
    const int qr = p.q_row[first];
    const float* b = qr < 0 ? nullptr : p.p_prev + size_t(qr) * V;
    const int xd = p.draft[first];
    const bool text = s_text_mode != 0;
    const int* fr = p.forced_resid ? p.forced_resid : p.forced;
    // fr[j] >= 0: the grammar makes the residual one-hot at that id whatever its value (Lumina / Emu3 write 0 there
    // and -inf elsewhere); fr[j] <= -2: "mask-forced" id f = -2 - fr[j] — every OTHER id is filled with finfo.min (the
    // Anole processors, logit_processor_3dim.py:242-256): f wins if its residual is finite, else (p <= q there, i.e. f
    // was outside the window's support) softmax sees V-1 equal finfo.min entries and the draw is uniform over them.
    int forced = (!text && fr) ? fr[j] : -1;
    bool mask_forced = false;
    if (forced <= -2) { forced = -2 - forced; mask_forced = true; }
    const Cand cd = make_cand_resid(p, text, j);
    const int lo = cd.span_lo, hi = cd.span_hi;
    const int v0 = (lo / int(blockDim.x)) * int(blockDim.x);
    const int top_k = text ? p.text_top_k : p.top_k;
    int tok;
    if (forced >= 0 && mask_forced &&
        !(__fsub_rn(a[forced], b ? b[forced] : (forced == xd ? 1.f : 0.f)) > 0.f)) {
      const float pc = 1.f / float(V - 1);
      float best = -INFINITY;
      int besti = 0x7fffffff;
      for (int v = threadIdx.x; v < V; v += blockDim.x) {
        if (v == forced) continue;
        const float val = pc / noise_row_e2(p)[v];
        if (val > best) { best = val; besti = v; }
      }
      tok = block_argmax(best, besti, sc);
    } else if (forced >= 0) {
      tok = forced;   // one-hot residual support after the grammar: the multinomial can only return it
    } else if (hi - v0 <= int(blockDim.x) * kRegVPT && !(p.top_p_thresh > 0.f)) {
      float s[kRegVPT];
#pragma unroll
      for (int jj = 0; jj < kRegVPT; ++jj) {
        const int v = v0 + jj * blockDim.x + threadIdx.x;
        float sv = -INFINITY;
        if (v >= lo && v < hi && cd.ok(v)) {
          const float q = b ? b[v] : (v == xd ? 1.f : 0.f);
          sv = logf(fmaxf(__fsub_rn(a[v], q), 0.f));
          if (p.temperature != 1.f) sv = sv / p.temperature;
        }
        s[jj] = sv;
      }
      tok = block_topk_softmax_sample_regs<kRegVPT>(s, v0, lo, hi, top_k, 1, noise_row_e2(p), nullptr, V, sc);
    } else {
      const int ve = min(V, ((hi + int(blockDim.x) - 1) / int(blockDim.x)) * int(blockDim.x));
      const bool nucleus = p.top_p_thresh > 0.f;   // its pass walks the whole row: give it zeros outside the span
      if (nucleus) {
        for (int v = threadIdx.x; v < v0; v += blockDim.x) p.resid[v] = 0.f;
        for (int v = ve + threadIdx.x; v < V; v += blockDim.x) p.resid[v] = 0.f;
      }
      for (int v = v0 + threadIdx.x; v < ve; v += blockDim.x) {
        float s = -INFINITY;
        if (v >= lo && v < hi && cd.ok(v)) {
          const float q = b ? b[v] : (v == xd ? 1.f : 0.f);
          s = logf(fmaxf(__fsub_rn(a[v], q), 0.f));
          if (p.temperature != 1.f) s = s / p.temperature;
        }
        p.resid[v] = s;
      }
      __syncthreads();
      tok = block_topk_softmax_sample(p.resid, v0, ve, V, top_k, 1, noise_row_e2(p), sc);
      if (nucleus) {   // the residual goes through the same processors (reject_sampling_single_token)
        __syncthreads();
        tok = block_top_p(p.resid, V, p.top_p_thresh, 1, tok, noise_row_e2(p), sc, tp);
      }
    }
    if (threadIdx.x == 0) p.out_tokens[j] = tok;
  }
  if (threadIdx.x == 0) {
    p.out_info[0] = first;
    p.out_info[1] = rejected ? 1 : 0;
    p.out_info[2] = first;
    p.out_info[3] = s_text_mode;
    if (p.done_flag) {
      // every thread's out_tokens store precedes the last __syncthreads this thread passed: the fence is cumulative
      __threadfence_system();
      *reinterpret_cast<volatile int*>(p.done_flag) = p.done_seq;
    }
  }
}

// One launch for the whole verify step: every CTA processes its window position, the CTA that finishes LAST (a counter
// in sync_ws, which it leaves zero again) runs the accept scan — it then sees every row's probabilities and samples.
__global__ void __launch_bounds__(kVerifyThreads) verify_kernel(VerifyParams p) {
  __shared__ BlockScratch sc;
  __shared__ TopPScratch tp;
  __shared__ int s_last;
  verify_row(p, blockIdx.x, sc, tp);
  __syncthreads();                       // the whole row (p_cur, next_tokens) has been written by this CTA
  if (threadIdx.x == 0) {
    __threadfence();                     // ... and is visible device-wide before the count goes up
    s_last = atomicAdd(p.sync_ws, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) *p.sync_ws = 0u; // re-arm for the next launch
  __threadfence();                       // acquire side: the other CTAs' rows
  verify_accept(p, sc, tp);
}

// two-launch form for callers that pass no sync word
__global__ void __launch_bounds__(kVerifyThreads) verify_rows_kernel(VerifyParams p) {
  __shared__ BlockScratch sc;
  __shared__ TopPScratch tp;
  verify_row(p, blockIdx.x, sc, tp);
}
__global__ void __launch_bounds__(kVerifyThreads) verify_accept_kernel(VerifyParams p) {
  __shared__ BlockScratch sc;
  __shared__ TopPScratch tp;
  verify_accept(p, sc, tp);
}

// developer / test entry: what the kernel would draw for every element of a [numel] noise tensor
// kind: 0 = exponential_ as the verify kernel draws it, 1 = rand as the verify kernel draws it; developer variants used to
// pin the transform against torch on the GPU: 16 + v: exponential with log variant v (0 logf, 1 __logf, 2 __log2f * ln 2),
// +8: uniform conversion as separate multiply and add instead of one FMA
__global__ void philox_fill_kernel(float* out, unsigned long long numel, unsigned long long seed, unsigned long long off,
                                   uint32_t span, int kind) {
  for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < numel;
       e += (unsigned long long)gridDim.x * blockDim.x) {
    if (kind == 0) { out[e] = torch_exponential(seed, off, span, e); continue; }
    if (kind == 1) { out[e] = torch_rand(seed, off, span, e); continue; }
    float u = torch_uniform_raw(seed, off, span, e);
    if (kind & 8) {   // recompute the uniform without contraction
      const unsigned long long q = e / span;
      const uint32_t t = uint32_t(e - q * span), lane = uint32_t(q & 3ull);
      const unsigned long long ctr = (off >> 2) + (q >> 2);
      const uint4 r = philox4x32_10(make_uint4(uint32_t(ctr), uint32_t(ctr >> 32), t, 0u), make_uint2(uint32_t(seed), uint32_t(seed >> 32)));
      const uint32_t x = lane == 0 ? r.x : (lane == 1 ? r.y : (lane == 2 ? r.z : r.w));
      u = __fadd_rn(__fmul_rn(float(x), 2.3283064e-10f), 2.3283064e-10f / 2.0f);
    }
    const int v = kind & 7;
    if (kind & 32) { out[e] = (u == 1.0f) ? 0.0f : u; continue; }   // rand with the chosen uniform
    float lg;
    if (u >= 1.0f - 1.1920929e-07f / 2.0f) lg = -1.1920929e-07f / 2.0f;
    else lg = v == 0 ? logf(u) : (v == 1 ? __logf(u) : __log2f(u) * 0.6931471805599453f);
    out[e] = -1.0f * lg;
  }
}
int philox_fill(float* out, unsigned long long numel, unsigned long long seed, unsigned long long off, uint32_t span,
                int kind, cudaStream_t stream) {
  philox_fill_kernel<<<1024, 256, 0, stream>>>(out, numel, seed, off, span, kind);
  return cudaGetLastError() == cudaSuccess ? 0 : -6;
}

int verify_launch(const VerifyParams& p, cudaStream_t stream) {
  if (p.W < 1 || p.W > kVerifyThreads) return -3;
  if (p.sync_ws) {
    verify_kernel<<<p.W, kVerifyThreads, 0, stream>>>(p);
  } else {
    verify_rows_kernel<<<p.W, kVerifyThreads, 0, stream>>>(p);
    verify_accept_kernel<<<1, kVerifyThreads, 0, stream>>>(p);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -6;
}

}  // namespace sjd
