// Output side of the decode loop: the image-token ids the SJD loop has emitted -> the latent feature
// map the VQGAN decoder starts from.  One kernel does what the reference spreads over an embedding gather, an optional
// L2 normalisation, a reshape / permute and a 1x1 convolution:
//   LlamaGen   vq_model.decode_code (llamagen/tokenizer/tokenizer_image/vq_model.py:52-55): get_codebook_entry (:261-275,
//              F.normalize(embedding)[indices] -> [B, C, h, w]) then post_quant_conv (:39, :47-49)
//   Chameleon  VQModel.decode_code (lumina_mgpt/model/chameleon_vae_ori/vqgan.py:594-597): embedding(indices)
//              (:131-146) then post_quant_conv (:589-592)
//   out[b][z][pix] = bias[z] + sum_e W[z][e] * cb[code[b][pix]][e] / (l2 ? max(||cb[code]||, 1e-12) : 1)
// fp32 like the reference's decoders.  The convolution stack that follows stays library code (cuDNN through torch).
#include "common.cuh"

namespace sjd {

constexpr int kVqPix = 32;   // pixels per block

__global__ void __launch_bounds__(256)
vq_lookup_postquant_kernel(const int* __restrict__ codes, int n_pix, int hw, const float* __restrict__ cb, int n_e, int e_dim,
                           int l2_norm, const float* __restrict__ w, const float* __restrict__ bias, int z,
                           float* __restrict__ out) {
  extern __shared__ float emb[];   // [kVqPix][e_dim + 1]
  __shared__ float inv_norm[kVqPix];
  const int p0 = blockIdx.x * kVqPix, ld = e_dim + 1;
  for (int i = threadIdx.x; i < kVqPix * e_dim; i += blockDim.x) {
    const int pp = i / e_dim, e = i - pp * e_dim, pix = p0 + pp;
    int code = pix < n_pix ? codes[pix] : 0;
    code = code < 0 ? 0 : (code >= n_e ? n_e - 1 : code);   // (the caller validates; never read outside the codebook)
    emb[pp * ld + e] = cb[size_t(code) * e_dim + e];
  }
  __syncthreads();
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int pp = warp; pp < kVqPix; pp += blockDim.x >> 5) {
      float ss = 0.f;
      for (int e = lane; e < e_dim; e += 32) ss += emb[pp * ld + e] * emb[pp * ld + e];
      ss = warp_sum(ss);
      if (lane == 0) inv_norm[pp] = l2_norm ? 1.f / fmaxf(sqrtf(ss), 1e-12f) : 1.f;
    }
  }
  __syncthreads();
  // thread -> (pixel of the tile, channel group): consecutive threads write consecutive pixels of one channel
  const int pp = threadIdx.x & (kVqPix - 1), pix = p0 + pp;
  if (pix >= n_pix) return;
  const int b = pix / hw, r = pix - b * hw;
  const float s = inv_norm[pp];
  for (int c = threadIdx.x / kVqPix; c < z; c += blockDim.x / kVqPix) {
    const float* wr = w + size_t(c) * e_dim;
    float acc = 0.f;
    for (int e = 0; e < e_dim; ++e) acc = fmaf(wr[e], emb[pp * ld + e] * s, acc);
    out[(size_t(b) * z + c) * hw + r] = acc + bias[c];
  }
}

int vq_lookup_launch(const int* codes, int n_pix, int hw, const float* cb, int n_e, int e_dim, int l2_norm, const float* w,
                     const float* bias, int z, float* out, cudaStream_t stream) {
  const size_t smem = size_t(kVqPix) * (e_dim + 1) * sizeof(float);
  if (smem > 200 * 1024) return -3;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    if (cudaFuncSetAttribute(vq_lookup_postquant_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
      return -5;
    attr = smem;
  }
  vq_lookup_postquant_kernel<<<(n_pix + kVqPix - 1) / kVqPix, 256, smem, stream>>>(codes, n_pix, hw, cb, n_e, e_dim, l2_norm, w,
                                                                                   bias, z, out);
  return cudaGetLastError() == cudaSuccess ? 0 : -6;
}

}  // namespace sjd
