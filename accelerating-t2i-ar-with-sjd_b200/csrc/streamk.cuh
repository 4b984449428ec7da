// Stream-K work partition shared by the weight-streaming GEMM (producer of partial tiles) and the
// reduce epilogues (consumers).  A GEMM  Y[M,N] = X[M,K] * W[N,K]^T  is cut into
//   units u = tile * KB + kb,  tile in [0, n_tiles) (128 weight rows), kb in [0, KB) (64 K columns)
// and CTA c of G owns the contiguous range [c*U/G, (c+1)*U/G).  Each maximal run of units of one tile
// inside one CTA is a *segment*; its fp32 partial tile goes to workspace slot (tile + c), which is
// unique per segment (tiles are non-decreasing in c).  The final value of tile t is the sum, in CTA
// order, over the CTAs [first_cta(t), last_cta(t)] — a fixed order, so results are deterministic.
#pragma once
#include <stdint.h>

namespace sjd {

struct StreamK {
  int n_tiles;   // ceil(N / 128)
  int kb;        // K / 64
  int m_tile;    // token rows held per partial tile (multiple of 16, <= 256)
  int grid;      // G

  __host__ __device__ __forceinline__ uint32_t units() const { return uint32_t(n_tiles) * uint32_t(kb); }
  __host__ __device__ __forceinline__ uint32_t begin(int c) const {
    return uint32_t((uint64_t(c) * units()) / uint64_t(grid));
  }
  // CTA that owns unit u: the largest c with begin(c) <= u.
  __host__ __device__ __forceinline__ int owner(uint32_t u) const {
    return int((uint64_t(u + 1) * uint64_t(grid) - 1) / uint64_t(units()));
  }
  __host__ __device__ __forceinline__ int first_cta(int tile) const { return owner(uint32_t(tile) * kb); }
  __host__ __device__ __forceinline__ int last_cta(int tile) const {
    return owner(uint32_t(tile + 1) * kb - 1);
  }
  __host__ __device__ __forceinline__ size_t slot_floats() const { return size_t(m_tile) * 128; }
  __host__ __device__ __forceinline__ size_t ws_floats() const {
    return size_t(n_tiles + grid) * slot_floats();
  }
};

// Sum of the partial tiles for output element (m, n); ws layout is [slot][m][128].
__device__ __forceinline__ float streamk_gather(const float* __restrict__ ws, const StreamK& sk, int m, int n) {
  const int tile = n >> 7, nl = n & 127;
  const int c0 = sk.first_cta(tile), c1 = sk.last_cta(tile);
  float acc = 0.f;
  for (int c = c0; c <= c1; ++c) acc += ws[size_t(tile + c) * sk.slot_floats() + size_t(m) * 128 + nl];
  return acc;
}

}  // namespace sjd
