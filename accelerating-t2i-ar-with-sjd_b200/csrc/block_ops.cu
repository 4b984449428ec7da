// Small row-wise kernels around the fused GEMMs of the Jacobi window forward:
//   embed_rmsnorm_rows  token gather (modeling_chameleon.py:1301, llamagen.py:314) + the first RMSNorm of the
//                       stack (modeling_chameleon.py:68-73, llamagen.py:179-181)
//   gather_rows         keeps the last n tokens of every CFG row for the lm_head (the reference slices logits
//                       after computing all of them, jacobi_iteration_lumina_mgpt.py:97)
//   pack_tiles          one-time weight re-layout into contiguous 16 KB GEMM tiles (+ gate/up row interleave)
// Every other element-wise / normalisation op of the block is an epilogue of gemm_fused.cu.
#include "common.cuh"

namespace sjd {

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

// h[m] = table[ids[m]] (or the given embedding row);  xn[m] = bf16( w * bf16( h * rsqrt(mean(h^2) + eps) ) )
__global__ void __launch_bounds__(256)
embed_rmsnorm_rows_kernel(const int* __restrict__ ids, const __nv_bfloat16* __restrict__ table,
                          const __nv_bfloat16* __restrict__ embeds, __nv_bfloat16* __restrict__ h,
                          const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ xn, int d, float eps) {
  __shared__ float red[32];
  pdl_wait();
  pdl_launch_dependents();
  const int m = blockIdx.x;
  const __nv_bfloat16* src = embeds ? embeds + size_t(m) * d : table + size_t(ids[m]) * d;
  __nv_bfloat16* row = h + size_t(m) * d;
  float ss = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const __nv_bfloat16 t = src[i];
    row[i] = t;
    const float x = __bfloat162float(t);
    ss += x * x;
  }
  ss = block_sum(ss, red);
  const float r = rsqrtf(ss / float(d) + eps);
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const float x = __bfloat162float(src[i]);
    xn[size_t(m) * d + i] = __float2bfloat16_rn(__bfloat162float(w[i]) * bf16_round(x * r));
  }
}

// dst row r = b*n + t  <-  src row b*W + (W-n+t)
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int W,
                                   int n, int d) {
  pdl_wait();
  pdl_launch_dependents();
  const int r = blockIdx.x, b = r / n, t = r - b * n;
  const uint4* s = reinterpret_cast<const uint4*>(src + size_t(b * W + (W - n + t)) * d);
  uint4* o = reinterpret_cast<uint4*>(dst + size_t(r) * d);
  for (int i = threadIdx.x; i < d / 8; i += blockDim.x) o[i] = s[i];
}

// One-time weight re-layout into contiguous GEMM units: dst[(group*KB + kb)][h][128][64] <- src[row(tile, r)][kb*64 ..]
// with tile = group * tpu + h, so that every stream-K unit (tpu tiles against one activation tile, streamk.cuh) is one
// contiguous tpu * 16 KB block of HBM.  Rows beyond N (and the missing tiles of a ragged last group) are zero.
// gate_up != 0 additionally interleaves gate/up rows (src = [gate rows; up rows], ff rows each): tile row 4l+e is
// gate[64*tile + 2l + e] for e < 2 and up[64*tile + 2l + e - 2] otherwise, so that lane l of the GEMM epilogue holds
// gate and up of act columns 2l, 2l+1.
__global__ void pack_tiles_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int N, int K,
                                  int gate_up_ff, int tpu) {
  const int KB = K / 64;
  const int r = blockIdx.x & 127, tile = blockIdx.x >> 7;
  const int grp = tile / tpu, hh = tile - grp * tpu;
  int srow;
  if (gate_up_ff) {
    const int l = r >> 2, e = r & 3;
    srow = e < 2 ? tile * 64 + 2 * l + e : gate_up_ff + tile * 64 + 2 * l + (e - 2);
  } else {
    srow = tile * 128 + r;
  }
  for (int i = threadIdx.x; i < K / 8; i += blockDim.x) {   // 16-byte chunks along K
    const int kb = i >> 3, c = i & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (srow < N) v = *reinterpret_cast<const uint4*>(src + size_t(srow) * K + size_t(i) * 8);
    *reinterpret_cast<uint4*>(dst + (((size_t(grp) * KB + kb) * tpu + hh) * 128 + r) * 64 + c * 8) = v;
  }
}

// ---- launchers --------------------------------------------------------------------------
int embed_rmsnorm_rows(const int* ids, const __nv_bfloat16* table, const __nv_bfloat16* embeds, __nv_bfloat16* h,
                       const __nv_bfloat16* w, __nv_bfloat16* xn, int M, int d, float eps, cudaStream_t s) {
  return launch_pdl(embed_rmsnorm_rows_kernel, dim3(M), dim3(256), 0, s, ids, table, embeds, h, w, xn, d, eps);
}
int gather_rows(const __nv_bfloat16* src, __nv_bfloat16* dst, int rows, int W, int n, int d, cudaStream_t s) {
  return launch_pdl(gather_rows_kernel, dim3(rows * n), dim3(128), 0, s, src, dst, W, n, d);
}
int pack_tiles(const __nv_bfloat16* src, __nv_bfloat16* dst, int N, int K, int gate_up_ff, int tpu, cudaStream_t s) {
  const int n_tiles = ((N + 127) / 128 + tpu - 1) / tpu * tpu;   // whole groups: the tiles of a ragged last group are zero
  pack_tiles_kernel<<<n_tiles * 128, 128, 0, s>>>(src, dst, N, K, gate_up_ff, tpu);
  return cudaGetLastError() == cudaSuccess ? 0 : -6;
}

}  // namespace sjd
