// Weight-streaming "swap-AB" GEMM for the Jacobi draft window:  Y[M,N] = X[M,K] * W[N,K]^T with
// M = 2*B*W token rows (16..256) and N,K in the thousands.  The window is far below the B200 ridge
// (M FLOP/byte vs ~210), so the kernel is designed as an HBM streamer that happens to use tensor cores:
//   * the WEIGHT tile is the UMMA "A" operand (128 rows fill the M=128 slot of tcgen05.mma),
//     the TOKEN rows are the UMMA "N" dimension (16..256), accumulators [128 x M] fp32 live in TMEM;
//   * persistent grid of one CTA per SM, stream-K split of the flattened (tile, k-block) space so that
//     every SM streams the same number of weight bytes regardless of N (streamk.cuh);
//   * warp 0 = TMA producer (cp.async.bulk.tensor, SWIZZLE_128B, deep mbarrier ring),
//     warp 1 = single-thread tcgen05.mma issuer, warps 2..5 = TMEM -> fp32 partial-tile epilogue;
//   * TMEM accumulators are double buffered so the epilogue of a segment overlaps the next segment.
// Replaces the cuBLAS calls behind nn.Linear in the reference forward
// (lumina_mgpt/model/chameleon/modeling_chameleon.py:527-529,579,193-195,1560; llamagen/llamagen.py:248,277,200,332).
#include "common.cuh"
#include "streamk.cuh"

namespace sjd {

constexpr int kGemmThreads = 192;
constexpr int kBlockN = 128;  // weight rows per tile (UMMA M)
constexpr int kBlockK = 64;   // bf16 K elements per stage = one 128-byte swizzle row
constexpr int kMaxStages = 12;
constexpr uint32_t kATileBytes = kBlockN * kBlockK * 2;  // 16 KB

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_streamk_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                    float* __restrict__ ws, StreamK sk, int num_stages, uint32_t tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 4];
  __shared__ uint32_t tmem_holder;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_tile_bytes = uint32_t(sk.m_tile) * kBlockK * 2;
  const uint32_t stage_bytes = kATileBytes + b_tile_bytes;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * kMaxStages + 2 + a); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_holder), tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_holder;

  const int cta = blockIdx.x;
  const uint32_t u0 = sk.begin(cta), u1 = sk.begin(cta + 1);
  const uint32_t KB = uint32_t(sk.kb);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (uint32_t u = u0; u < u1; ++u) {
        const uint32_t tile = u / KB, kb = u - tile * KB;
        mbar_wait(empty_bar(stage), phase ^ 1);
        mbar_arrive_expect_tx(full_bar(stage), stage_bytes);
        const uint32_t sa = smem_base + uint32_t(stage) * stage_bytes;
        tma_load_2d(sa, &tmap_w, int(kb * kBlockK), int(tile * kBlockN), full_bar(stage), kPolicyEvictFirst);
        tma_load_2d(sa + kATileBytes, &tmap_x, int(kb * kBlockK), 0, full_bar(stage), kPolicyEvictLast);
        if (++stage == num_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16_f32(kBlockN, uint32_t(sk.m_tile));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      uint32_t u = u0;
      while (u < u1) {
        const uint32_t tile = u / KB;
        const uint32_t seg_end = min(u1, (tile + 1) * KB);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc) * uint32_t(sk.m_tile);
        uint32_t accumulate = 0;
        for (; u < seg_end; ++u) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + uint32_t(stage) * stage_bytes;
          const uint64_t da = umma_desc_sw128_kmajor(sa);
          const uint64_t db = umma_desc_sw128_kmajor(sa + kATileBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            umma_bf16_ss(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(empty_bar(stage));
          if (++stage == num_stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===== epilogue: TMEM -> fp32 partial tile in the stream-K workspace =====
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t u = u0;
    while (u < u1) {
      const uint32_t tile = u / KB;
      const uint32_t seg_end = min(u1, (tile + 1) * KB);
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      float* dst = ws + size_t(tile + cta) * sk.slot_floats() + row;
      const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc) * uint32_t(sk.m_tile);
      for (int m0 = 0; m0 < sk.m_tile; m0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_addr + uint32_t(m0), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[size_t(m0 + j) * 128] = __uint_as_float(v[j]);
      }
      tcgen05_fence_before();
      mbar_arrive(tempty_bar(acc));
      u = seg_end;
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ----------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols], 128-byte swizzle.
int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {kBlockK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

static int g_num_sms = 0;
int device_num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_num_sms;
}

StreamK gemm_partition(int N, int K, int m_tile, int grid_limit) {
  StreamK sk;
  sk.n_tiles = (N + kBlockN - 1) / kBlockN;
  sk.kb = K / kBlockK;
  sk.m_tile = m_tile;
  int g = grid_limit > 0 ? grid_limit : device_num_sms();
  uint32_t U = uint32_t(sk.n_tiles) * uint32_t(sk.kb);
  sk.grid = int(U < uint32_t(g) ? U : uint32_t(g));
  return sk;
}

struct GemmLaunch {
  CUtensorMap tmap_w, tmap_x;
  StreamK sk;
  int num_stages;
  uint32_t tmem_cols;
  uint32_t smem_bytes;
};

int gemm_prepare(GemmLaunch* g, const void* w, int N, int K, const void* x, int x_rows, int m_tile, int grid_limit) {
  if (K % kBlockK != 0 || m_tile % 16 != 0 || m_tile < 16 || m_tile > 256 || x_rows < m_tile) return -3;
  g->sk = gemm_partition(N, K, m_tile, grid_limit);
  if (make_tmap_bf16_2d(&g->tmap_w, w, uint64_t(N), uint64_t(K), kBlockN)) return -1;
  if (make_tmap_bf16_2d(&g->tmap_x, x, uint64_t(x_rows), uint64_t(K), uint32_t(m_tile))) return -1;
  const uint32_t stage_bytes = kATileBytes + uint32_t(m_tile) * kBlockK * 2;
  const uint32_t budget = 200 * 1024;
  int stages = int(budget / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return -4;
  g->num_stages = stages;
  uint32_t cols = 32;
  while (cols < uint32_t(2 * m_tile)) cols <<= 1;
  g->tmem_cols = cols;
  g->smem_bytes = uint32_t(stages) * stage_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_streamk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024) !=
        cudaSuccess)
      return -5;
    attr_set = true;
  }
  return 0;
}

int gemm_launch(const GemmLaunch* g, float* ws, cudaStream_t stream) {
  gemm_streamk_kernel<<<g->sk.grid, kGemmThreads, g->smem_bytes, stream>>>(g->tmap_w, g->tmap_x, ws, g->sk,
                                                                           g->num_stages, g->tmem_cols);
  return cudaGetLastError() == cudaSuccess ? 0 : -6;
}

}  // namespace sjd
