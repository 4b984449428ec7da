// Fused row-wise epilogues of the Jacobi window forward.  Every GEMM of the block leaves fp32
// stream-K partial tiles in a workspace (gemm_tcgen05.cu); the kernels here finish them and apply the
// element-wise / normalisation work of the reference transformer block in one pass each:
//   embed_rows                token gather (modeling_chameleon.py:1301, llamagen.py:314)
//   rmsnorm_rows              first RMSNorm of the stack (modeling_chameleon.py:68-73, llamagen.py:179-181)
//   reduce_residual_rmsnorm   o_proj/down_proj reduce + residual add + next RMSNorm (modeling_chameleon.py:643-659)
//   qkv_post                  qkv reduce + per-head QK-LayerNorm (:216-219) + RoPE (:153-177 rotate-half /
//                             llamagen.py:457-467 interleaved pairs) + KV-cache append (replaces
//                             DynamicCache.update torch.cat, :547, and llamagen.py:216-217)
//   silu_mul                  gate/up reduce + SiLU(gate)*up (modeling_chameleon.py:193-195)
//   logits_reduce             lm_head reduce -> fp32 logits (:1560-1561)
// Rounding points follow the bf16 reference: every nn.Linear output, norm output and residual sum is
// rounded to bf16 before the next op; reductions and norms are computed in fp32.
#include "common.cuh"
#include "streamk.cuh"

namespace sjd {

__global__ void embed_rows_kernel(const int* __restrict__ ids, const __nv_bfloat16* __restrict__ table,
                                  __nv_bfloat16* __restrict__ h, int d) {
  const int m = blockIdx.x;
  const uint4* src = reinterpret_cast<const uint4*>(table + size_t(ids[m]) * d);
  uint4* dst = reinterpret_cast<uint4*>(h + size_t(m) * d);
  for (int i = threadIdx.x; i < d / 8; i += blockDim.x) dst[i] = src[i];
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

// xn = bf16( w * bf16( x * rsqrt(mean(x^2) + eps) ) )
__global__ void rmsnorm_rows_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ w,
                                    __nv_bfloat16* __restrict__ xn, int d, float eps) {
  __shared__ float red[32];
  const int m = blockIdx.x;
  const __nv_bfloat16* row = h + size_t(m) * d;
  float ss = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const float x = __bfloat162float(row[i]);
    ss += x * x;
  }
  ss = block_sum(ss, red);
  const float r = rsqrtf(ss / float(d) + eps);
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const float x = __bfloat162float(row[i]);
    xn[size_t(m) * d + i] = __float2bfloat16_rn(__bfloat162float(w[i]) * bf16_round(x * r));
  }
}

// h <- bf16(h + bf16(sum partials));  xn <- rmsnorm(h) * w        (dynamic smem: d floats)
__global__ void reduce_residual_rmsnorm_kernel(const float* __restrict__ ws, StreamK sk,
                                               __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ w,
                                               __nv_bfloat16* __restrict__ xn, int d, float eps) {
  extern __shared__ float rowbuf[];
  __shared__ float red[32];
  const int m = blockIdx.x;
  __nv_bfloat16* row = h + size_t(m) * d;
  float ss = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const float lin = bf16_round(streamk_gather(ws, sk, m, i));
    const float x = bf16_round(__bfloat162float(row[i]) + lin);
    rowbuf[i] = x;
    ss += x * x;
  }
  ss = block_sum(ss, red);
  const float r = rsqrtf(ss / float(d) + eps);
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const float x = rowbuf[i];
    row[i] = __float2bfloat16_rn(x);
    xn[size_t(m) * d + i] = __float2bfloat16_rn(__bfloat162float(w[i]) * bf16_round(x * r));
  }
}

struct QkvPostParams {
  const float* ws;
  StreamK sk;
  __nv_bfloat16* q_out;          // [M][H][Dh]
  __nv_bfloat16* k_cache;        // layer base [rows][Hkv][Lmax][Dh]
  __nv_bfloat16* v_cache;
  const int* rope_pos;           // [M] index into the rope tables
  const int* cache_pos;          // [M] key slot to write
  const float* rope_cos;         // [n_pos][Dh/2]
  const float* rope_sin;
  const __nv_bfloat16* q_norm_w; // [H][Dh] or null
  const __nv_bfloat16* q_norm_b;
  const __nv_bfloat16* k_norm_w; // [Hkv][Dh] or null
  const __nv_bfloat16* k_norm_b;
  int M, W, H, Hkv, Lmax;
  int rope_interleaved;          // 0: rotate-half pairs (i, i+Dh/2); 1: pairs (2i, 2i+1)
};

// one warp per (token row m, head slot) ; slots: [0,H) = q, [H,H+Hkv) = k, [H+Hkv, H+2Hkv) = v
template <int DH>
__global__ void __launch_bounds__(256) qkv_post_kernel(QkvPostParams p) {
  constexpr int PER = DH / 32;  // 2 or 4 elements per lane
  const int slots = p.H + 2 * p.Hkv;
  const int widx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (widx >= p.M * slots) return;
  const int lane = threadIdx.x & 31;
  const int m = widx / slots, slot = widx - m * slots;
  const int n0 = slot * DH;
  // element index held in register e
  int ei[PER];
#pragma unroll
  for (int e = 0; e < PER; ++e) {
    if (p.rope_interleaved) ei[e] = (e >> 1) * 64 + lane * 2 + (e & 1);   // pairs (2l, 2l+1) [+64]
    else ei[e] = lane + 32 * e;                                            // pairs (e, e + PER/2)
  }
  float x[PER];
#pragma unroll
  for (int e = 0; e < PER; ++e) x[e] = bf16_round(streamk_gather(p.ws, p.sk, m, n0 + ei[e]));

  const int b = m / p.W;
  const bool is_q = slot < p.H, is_k = !is_q && slot < p.H + p.Hkv;
  if (is_q || is_k) {
    const int head = is_q ? slot : slot - p.H;
    const __nv_bfloat16* nw = is_q ? p.q_norm_w : p.k_norm_w;
    const __nv_bfloat16* nb = is_q ? p.q_norm_b : p.k_norm_b;
    if (nw) {  // per-head LayerNorm over Dh, eps 1e-5, then gamma/beta of this head
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < PER; ++e) s += x[e];
      const float mean = warp_sum(s) / float(DH);
      float v = 0.f;
#pragma unroll
      for (int e = 0; e < PER; ++e) { const float dlt = x[e] - mean; v += dlt * dlt; }
      const float rstd = rsqrtf(warp_sum(v) / float(DH) + 1e-5f);
#pragma unroll
      for (int e = 0; e < PER; ++e)
        x[e] = (x[e] - mean) * rstd * __bfloat162float(nw[head * DH + ei[e]]) + __bfloat162float(nb[head * DH + ei[e]]);
    }
    const float* cs = p.rope_cos + size_t(p.rope_pos[m]) * (DH / 2);
    const float* sn = p.rope_sin + size_t(p.rope_pos[m]) * (DH / 2);
    if (p.rope_interleaved) {
#pragma unroll
      for (int e = 0; e < PER; e += 2) {
        const int pi = ei[e] >> 1;
        const float c = cs[pi], s = sn[pi];
        const float a = x[e], bb = x[e + 1];
        x[e] = a * c - bb * s;
        x[e + 1] = bb * c + a * s;
      }
    } else {
#pragma unroll
      for (int e = 0; e < PER / 2; ++e) {
        const int pi = ei[e];  // < Dh/2
        const float c = cs[pi], s = sn[pi];
        const float a = x[e], bb = x[e + PER / 2];
        x[e] = a * c - bb * s;            // q*cos + rotate_half(q)*sin, first half: -x2*sin
        x[e + PER / 2] = bb * c + a * s;  // second half: +x1*sin
      }
    }
  }
  __nv_bfloat16* dst;
  if (is_q) dst = p.q_out + (size_t(m) * p.H + slot) * DH;
  else {
    const int hk = is_k ? slot - p.H : slot - p.H - p.Hkv;
    __nv_bfloat16* base = is_k ? p.k_cache : p.v_cache;
    dst = base + ((size_t(b) * p.Hkv + hk) * p.Lmax + p.cache_pos[m]) * DH;
  }
#pragma unroll
  for (int e = 0; e < PER; ++e) dst[ei[e]] = __float2bfloat16_rn(x[e]);
}

// act[m][n] = bf16( bf16(silu(bf16 gate)) * bf16 up ),  gate = cols [0,I), up = cols [I,2I)
__global__ void silu_mul_kernel(const float* __restrict__ ws, StreamK sk, __nv_bfloat16* __restrict__ act, int M,
                                int I) {
  const int m = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= I || m >= M) return;
  const float g = bf16_round(streamk_gather(ws, sk, m, n));
  const float u = bf16_round(streamk_gather(ws, sk, m, I + n));
  const float s = bf16_round(g / (1.f + expf(-g)));
  act[size_t(m) * I + n] = __float2bfloat16_rn(s * u);
}

__global__ void logits_reduce_kernel(const float* __restrict__ ws, StreamK sk, float* __restrict__ logits, int M,
                                     int V, int round_bf16) {
  const int m = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= V || m >= M) return;
  const float v = streamk_gather(ws, sk, m, n);
  logits[size_t(m) * V + n] = round_bf16 ? bf16_round(v) : v;
}

// generic: out_bf16[m][n] = bf16(sum partials)   (used by tests and by projections without an epilogue)
__global__ void reduce_bf16_kernel(const float* __restrict__ ws, StreamK sk, __nv_bfloat16* __restrict__ out, int M,
                                   int N) {
  const int m = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N || m >= M) return;
  out[size_t(m) * N + n] = __float2bfloat16_rn(streamk_gather(ws, sk, m, n));
}

// ---- launchers --------------------------------------------------------------------------
int embed_rows(const int* ids, const __nv_bfloat16* table, __nv_bfloat16* h, int M, int d, cudaStream_t s) {
  embed_rows_kernel<<<M, 128, 0, s>>>(ids, table, h, d);
  return cudaGetLastError() == cudaSuccess ? 0 : -6;
}
int rmsnorm_rows(const __nv_bfloat16* h, const __nv_bfloat16* w, __nv_bfloat16* xn, int M, int d, float eps,
                 cudaStream_t s) {
  rmsnorm_rows_kernel<<<M, 256, 0, s>>>(h, w, xn, d, eps);
  return cudaGetLastError() == cudaSuccess ? 0 : -6;
}
int reduce_residual_rmsnorm(const float* ws, const StreamK& sk, __nv_bfloat16* h, const __nv_bfloat16* w,
                            __nv_bfloat16* xn, int M, int d, float eps, cudaStream_t s) {
  reduce_residual_rmsnorm_kernel<<<M, 256, d * sizeof(float), s>>>(ws, sk, h, w, xn, d, eps);
  return cudaGetLastError() == cudaSuccess ? 0 : -6;
}
int qkv_post(const QkvPostParams& p, int head_dim, cudaStream_t s) {
  const int total = p.M * (p.H + 2 * p.Hkv);
  const int blocks = (total + 7) / 8;
  if (head_dim == 128) qkv_post_kernel<128><<<blocks, 256, 0, s>>>(p);
  else if (head_dim == 64) qkv_post_kernel<64><<<blocks, 256, 0, s>>>(p);
  else return -3;
  return cudaGetLastError() == cudaSuccess ? 0 : -6;
}
int silu_mul(const float* ws, const StreamK& sk, __nv_bfloat16* act, int M, int I, cudaStream_t s) {
  dim3 grid((I + 255) / 256, M);
  silu_mul_kernel<<<grid, 256, 0, s>>>(ws, sk, act, M, I);
  return cudaGetLastError() == cudaSuccess ? 0 : -6;
}
int logits_reduce(const float* ws, const StreamK& sk, float* logits, int M, int V, int round_bf16, cudaStream_t s) {
  dim3 grid((V + 255) / 256, M);
  logits_reduce_kernel<<<grid, 256, 0, s>>>(ws, sk, logits, M, V, round_bf16);
  return cudaGetLastError() == cudaSuccess ? 0 : -6;
}
int reduce_bf16(const float* ws, const StreamK& sk, __nv_bfloat16* out, int M, int N, cudaStream_t s) {
  dim3 grid((N + 255) / 256, M);
  reduce_bf16_kernel<<<grid, 256, 0, s>>>(ws, sk, out, M, N);
  return cudaGetLastError() == cudaSuccess ? 0 : -6;
}

}  // namespace sjd
