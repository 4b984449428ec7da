// Stream-K work partition of the weight-streaming GEMM (gemm_fused.cu).  A GEMM  Y[M,N] = X[M,K] * W[N,K]^T  is cut into
//   units u = group * KB + kb,  group in [0, n_groups) (tpu tiles of 128 weight rows each), kb in [0, KB) (64 K columns)
// and CTA c of G owns the contiguous range [c*U/G, (c+1)*U/G).  One unit streams tpu weight tiles against ONE
// activation tile, so the activation bytes a CTA pulls from L2 per weight byte shrink by tpu (round 2: at 64+ token
// rows the kernel was bound by L2 -> SM traffic, not by HBM).  Each maximal run of units of one group inside one CTA
// is a *segment*.  Only the first and the last segment of a CTA can be part of a split group, so a CTA parks at most
// two segments = 2 * tpu fp32 partial tiles (workspace slots (2c + seg) * tpu + h).  The final value of a split tile
// is the sum of the partials of first_cta(group) .. last_cta(group) in CTA order — a fixed order, so results are
// bit-reproducible.
#pragma once
#include <stdint.h>

namespace sjd {

struct StreamK {
  int n_tiles;   // ceil(N / 128)
  int kb;        // K / 64
  int m_tile;    // token rows held per partial tile (multiple of 16, <= 256)
  int grid;      // G
  int tpu;       // weight tiles per unit (1, 2 or 4)
  int n_groups;  // ceil(n_tiles / tpu)

  __host__ __device__ __forceinline__ uint32_t units() const { return uint32_t(n_groups) * uint32_t(kb); }
  // CTAs beyond `grid` (a chain kernel may be launched wider than this op) get an empty range.
  __host__ __device__ __forceinline__ uint32_t begin(int c) const {
    const uint32_t cc = uint32_t(c < grid ? c : grid);
    return (cc * units()) / uint32_t(grid);   // 32-bit: the host checks units() * grid < 2^31
  }
  // CTA that owns unit u: the largest c with begin(c) <= u.
  __host__ __device__ __forceinline__ int owner(uint32_t u) const {
    return int(((u + 1) * uint32_t(grid) - 1) / units());
  }
  __host__ __device__ __forceinline__ int first_cta(int group) const { return owner(uint32_t(group) * kb); }
  __host__ __device__ __forceinline__ int last_cta(int group) const {
    return owner(uint32_t(group + 1) * kb - 1);
  }
  __host__ __device__ __forceinline__ size_t slot_floats() const { return size_t(m_tile) * 128; }
};

}  // namespace sjd
