// K13 -- the GEGLU projection of the UNet's feed-forward blocks with its activation in the epilogue:
//   out[m, f] = (x[m, :] . Wv[f, :] + bv[f]) * gelu(x[m, :] . Wg[f, :] + bg[f])        bf16 in / out, fp32 accumulate
// Replaces `GEGLU.proj` (nn.Linear(C, 8 C), cuBLAS) followed by K6 `fd_geglu` inside diffusers' FeedForward of every
// BasicTransformerBlock, reached from the UNet call at pipeline/guide.py:56-58.  The [M, 8 C] projection (168 MB at 64 x 64,
// two samples) is never written: the value and gate halves of a 128-feature block are the two column halves of ONE
// accumulator tile, multiplied together on their way out of TMEM.
//
// Structure = K2 v4 (`k2_gemm2_kernel`): cta_group::2 pairs, one tcgen05.mma spans both CTAs (M = 256 token rows, N = 256),
// each CTA holds its own 128 rows of x and HALF of the B tile -- here rank 0's half is the VALUE weights of the feature
// block and rank 1's half the GATE weights, so no weight permutation is needed.  Persistent over (m pair, feature block)
// tiles, 5-stage TMA ring, accumulator double-buffered in TMEM (the epilogue of tile i runs under the mainloop of tile
// i + 1), 8 epilogue warps: `tcgen05.ld` value + gate columns -> + bias -> v * gelu(g) (erf GELU as F.gelu, evaluated by
// `gelu_fast` to 1.2e-5 relative) -> bf16 ->
// SWIZZLE_128B staging -> TMA tensor stores of 64-column chunks.
// Tile order: the feature block is the fast index, so the pairs running together share a few 256-row slabs of x (read from
// HBM once, then L2 hits) and all of W (<= 26 MB, L2-resident).  With m fastest every feature block re-streamed x from HBM
// while the 4x larger output was flushing the L2: ncu showed 730 MB of DRAM reads for an 84 MB x at C = 320.
#include <stdlib.h>

#include "fd_common.cuh"

namespace fd {
namespace {

constexpr int G_BM = 128;
constexpr int G_BF = 128;             // features per tile (value columns [0, 128), gate columns [128, 256) of the accumulator)
constexpr int G_BN = 2 * G_BF;
constexpr int G_BK = 64;
constexpr int G_STAGES = 5;
constexpr int G_THREADS = 320;        // TMA, MMA, 8 epilogue warps (two 64-feature halves x four lane quarters)
constexpr int G_A_BYTES = G_BM * G_BK * 2;
constexpr int G_B_BYTES = G_BF * G_BK * 2;          // this CTA's half of the B tile
constexpr int G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;  // 32768
constexpr int G_OUT_CHUNK = G_BM * 64 * 2;          // 128 rows x 64 bf16
constexpr int G_OUT_BUFS = 4;                        // two staging buffers per epilogue half
constexpr int G_SMEM = G_STAGES * G_STAGE_BYTES + G_OUT_BUFS * G_OUT_CHUNK + G_BN * 4 /*bias (fp32)*/ +
                       1024 /*align*/ + 256 /*bars*/;
constexpr int G_TMEM_COLS = 2 * G_BN;
static_assert(G_SMEM <= 227 * 1024, "K13 shared memory");
__device__ int g_k13_flag;  // first bounded wait that gave up (0 = none)

// gelu(g) = g * Phi(g) with Phi(-a) = 2^p(a), a = |g|: p is the degree-7 least-squares fit of log2(erfc(a / sqrt 2) / 2) on
// [0, 6.72] (profiles/r02/SUMMARY.md, K13): relative error of gelu <= 1.2e-5 for |g| <= 6.7 -- 1/300 of a bf16 ulp of the
// bf16 result -- and absolute error < 1e-15 beyond (p keeps falling; a is clamped at 11 so the Horner chain stays finite).
// One MUFU (ex2) and ~12 FP32 issue slots per value where erff() takes ~40: the epilogue, not the tensor pipe, bounds this
// kernel at K = 320 / 640.
__device__ __forceinline__ float gelu_fast(float g) {
  const float a = fminf(fabsf(g), 11.0f);
  float p = -1.7024729004333494e-06f;
  p = fmaf(p, a, 5.87926188018173e-05f);
  p = fmaf(p, a, -0.0009072798420675099f);
  p = fmaf(p, a, 0.008412986062467098f);
  p = fmaf(p, a, -0.053762633353471756f);
  p = fmaf(p, a, -0.45865824818611145f);
  p = fmaf(p, a, -1.1511805057525635f);
  p = fmaf(p, a, -0.9999994039535522f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p));
  return g * (g < 0.f ? e : 1.0f - e);
}

__global__ void __launch_bounds__(G_THREADS, 1)
k13_ff_geglu_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                    const __grid_constant__ CUtensorMap tm_out, const __nv_bfloat16* __restrict__ bias, int F, int K,
                    int m_pairs, int f_tiles, int f_run) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* out_stage = smem + G_STAGES * G_STAGE_BYTES;
  float* bias_s = reinterpret_cast<float*>(out_stage + G_OUT_BUFS * G_OUT_CHUNK);   // [half][64 value | 64 gate] fp32
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_s + G_BN);
  uint64_t* empty_bar = full_bar + G_STAGES;
  uint64_t* tmem_full = empty_bar + G_STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2] (the leader's copy counts: 2 CTAs x 8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = K / G_BK;
  // work item = (256-row block, run of f_run adjacent feature blocks); f_run > 1 only when K = G_STAGES k-blocks, so that
  // k-block kb always lands in ring stage kb and the x tile can STAY there for the whole run (only W travels)
  const int f_groups = f_tiles / f_run;
  const int total_tiles = m_pairs * f_groups;   // work items
  const bool a_resident = f_run > 1;
  const uint32_t crank = cluster_cta_rank();
  const bool leader = crank == 0;
  const int first_tile = static_cast<int>(blockIdx.x >> 1);
  const int tile_stride = static_cast<int>(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_out);
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 16);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, G_TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_barrier();  // both CTAs' barriers and TMEM exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
        const int m_blk = tile / f_groups;   // the feature run is the fast index: see the tile-order note at the top
        for (int j = 0; j < f_run; ++j) {
          const int f_blk = (tile % f_groups) * f_run + j;
          // B rows of this CTA: the value weights (rank 0) or the gate weights (rank 1) of feature block f_blk
          const int w_row = static_cast<int>(crank) * F + f_blk * G_BF;
          const bool load_a = !a_resident || j == 0;
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % G_STAGES;
            mbar_wait_bounded(&empty_bar[s], ((it / G_STAGES) & 1) ^ 1, &g_k13_flag, 1);
            if (leader) mbar_expect_tx(&full_bar[s], load_a ? 2 * G_STAGE_BYTES : 2 * G_B_BYTES);
            const uint32_t lead_bar = mapa_rank(smem_u32(&full_bar[s]), 0);
            if (load_a)
              tma_load_2d_2cta(smem + s * G_STAGE_BYTES, &tm_x, lead_bar, kb * G_BK,
                               m_blk * 2 * G_BM + static_cast<int>(crank) * G_BM);
            tma_load_2d_2cta(smem + s * G_STAGE_BYTES + G_A_BYTES, &tm_w, lead_bar, kb * G_BK, w_row);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc(UMMA_BF16, 2 * G_BM, G_BN, 0, 0);
      int it = 0, local = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride)
      for (int j = 0; j < f_run; ++j, ++local) {
        const int acc = local & 1;
        mbar_wait_bounded(&tmem_empty[acc], ((local >> 1) & 1) ^ 1, &g_k13_flag, 2);  // both epilogues drained it
        tc_fence_after();
        const uint32_t d = tmem_base + acc * G_BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % G_STAGES;
          mbar_wait_bounded(&full_bar[s], (it / G_STAGES) & 1, &g_k13_flag, 3);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem + s * G_STAGE_BYTES), 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + s * G_STAGE_BYTES + G_A_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < G_BK / 16; ++k)
            mma_f16_ss_2cta(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          tc_commit_2cta_mcast(&empty_bar[s], 0x3);  // both producers may refill the stage
        }
        tc_commit_2cta_mcast(&tmem_full[acc], 0x3);  // both epilogues may read their rows
      }
    }
  } else {
    // 8 epilogue warps: lane quarter = warp % 4 (TMEM access rule), feature half = (warp - 2) / 4: half h turns value columns
    // [64 h, 64 h + 64) and gate columns [128 + 64 h, ...) into 64 output features = one staging chunk, with its own pair of
    // staging buffers, named barrier and TMA-store issuer.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const int et = threadIdx.x - 64;                    // 0..255
    const bool issuer = et % 128 == 0;                  // first thread of each half
    uint8_t* my_stage = out_stage + half * 2 * G_OUT_CHUNK;
    int local = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride)
    for (int j = 0; j < f_run; ++j, ++local) {
      const int m_blk = tile / f_groups, f_blk = (tile % f_groups) * f_run + j;
      const int acc = local & 1;
      // this half's biases of the tile as fp32: 64 value + 64 gate, one per thread of the half.  Single-buffered: a thread
      // gets here only after the half's second named barrier of the previous tile, i.e. after every read of the old
      // values, and the new ones are read after the first named barrier below
      float* bs = bias_s + half * 128;
      {
        const int t = et - 128 * half;   // 0..127
        bs[t] = __bfloat162float(bias[(t < 64 ? 0 : F - 64) + f_blk * G_BF + 64 * half + t]);
      }
      mbar_wait_bounded(&tmem_full[acc], (local >> 1) & 1, &g_k13_flag, 4);
      tc_fence_after();
      const uint32_t tbase = tmem_base + acc * G_BN + (static_cast<uint32_t>(quarter * 32) << 16);
      uint32_t v[4][16], gt[4][16];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        tmem_ld_x16(tbase + 64 * half + 16 * g, v[g]);
        tmem_ld_x16(tbase + G_BF + 64 * half + 16 * g, gt[g]);
      }
      tmem_ld_wait();
      // the accumulator is in registers: hand the TMEM buffer back before the arithmetic
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa_rank(smem_u32(&tmem_empty[acc]), 0));  // on the LEADER's barrier
      uint8_t* buf = my_stage + (local & 1) * G_OUT_CHUNK;
      if (issuer) tma_store_wait_read<1>();
      named_bar_sync(1 + half, 128);   // staging buffer free, this tile's biases visible
      const uint32_t bv4 = smem_u32(bs), bg4 = bv4 + 256;   // 64 value biases, then 64 gate biases (fp32)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 bva = lds_f4(bv4 + 16 * (4 * g + 2 * h)), bvb = lds_f4(bv4 + 16 * (4 * g + 2 * h + 1));
          const float4 bga = lds_f4(bg4 + 16 * (4 * g + 2 * h)), bgb = lds_f4(bg4 + 16 * (4 * g + 2 * h + 1));
          const float bvv[8] = {bva.x, bva.y, bva.z, bva.w, bvb.x, bvb.y, bvb.z, bvb.w};
          const float bgg[8] = {bga.x, bga.y, bga.z, bga.w, bgb.x, bgb.y, bgb.z, bgb.w};
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = 8 * h + 2 * j;
            // diffusers GEGLU: hidden * F.gelu(gate).  The projection output is bf16 in the path this replaces (cuBLAS
            // writes bf16 before K6 reads it): round both halves to bf16 first so the two paths agree
            const __nv_bfloat162 v2 = __floats2bfloat162_rn(__uint_as_float(v[g][c]) + bvv[2 * j],
                                                            __uint_as_float(v[g][c + 1]) + bvv[2 * j + 1]);
            const __nv_bfloat162 g2 = __floats2bfloat162_rn(__uint_as_float(gt[g][c]) + bgg[2 * j],
                                                            __uint_as_float(gt[g][c + 1]) + bgg[2 * j + 1]);
            const uint32_t vu = *reinterpret_cast<const uint32_t*>(&v2), gu = *reinterpret_cast<const uint32_t*>(&g2);
            const float o0 = __uint_as_float(vu << 16) * gelu_fast(__uint_as_float(gu << 16));
            const float o1 = __uint_as_float(vu & 0xFFFF0000u) * gelu_fast(__uint_as_float(gu & 0xFFFF0000u));
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(o0, o1);
            pk[j] = *reinterpret_cast<const uint32_t*>(&b2);
          }
          *reinterpret_cast<uint4*>(buf + sw128_offset(row, 2 * g + h)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + half, 128);
      if (issuer) {
        tma_store_2d(&tm_out, buf, f_blk * G_BF + 64 * half, m_blk * 2 * G_BM + static_cast<int>(crank) * G_BM);
        tma_store_commit();
      }
    }
    if (issuer) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  cluster_barrier();  // no CTA exits (or frees TMEM) while its peer may still signal / multiply into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, G_TMEM_COLS);
  }
}

}  // namespace
}  // namespace fd

extern "C" int fd_ff_geglu(const void* x_bf16_dev, const void* w_bf16_dev, const void* bias_bf16_dev, void* out_bf16_dev,
                           int M, int F, int K, int64_t out_row_stride, void* stream) {
  using namespace fd;
  FD_REQUIRE(x_bf16_dev && w_bf16_dev && bias_bf16_dev && out_bf16_dev, "fd_ff_geglu: NULL pointer");
  FD_REQUIRE(M > 0 && F > 0 && K > 0, "fd_ff_geglu: non-positive shape %d %d %d", M, F, K);
  FD_REQUIRE(K % G_BK == 0, "fd_ff_geglu: K=%d must be a multiple of %d", K, G_BK);
  FD_REQUIRE(F % G_BF == 0, "fd_ff_geglu: F=%d (features per half) must be a multiple of %d", F, G_BF);
  if (out_row_stride <= 0) out_row_stride = F;
  FD_REQUIRE(out_row_stride >= F && out_row_stride % 8 == 0, "fd_ff_geglu: output row stride %lld must be a multiple of 8 and >= F",
             static_cast<long long>(out_row_stride));
  FD_REQUIRE(reinterpret_cast<uintptr_t>(x_bf16_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(w_bf16_dev) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(out_bf16_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(bias_bf16_dev) % 2 == 0,
             "fd_ff_geglu: x / w / out must be 16-byte aligned");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  CUtensorMap tm_x, tm_w, tm_out;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {G_BK, G_BM};
    rc = encode_tmap(&tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x_bf16_dev, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(2) * F};
    uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {G_BK, G_BF};
    rc = encode_tmap(&tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_bf16_dev, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(F), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(out_row_stride) * 2};
    uint32_t box[2] = {64, G_BM};
    rc = encode_tmap(&tm_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out_bf16_dev, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
  }
  const int sms = sm_count();
  if (sms < 2) return set_error(FD_ERR_CUDA, "fd_ff_geglu: needs at least two SMs");
  static thread_local int attr_device = -1;
  int dev = 0;
  FD_CUDA_OK(cudaGetDevice(&dev));
  if (attr_device != dev) {
    FD_CUDA_OK(cudaFuncSetAttribute(k13_ff_geglu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM));
    attr_device = dev;
  }
  const int m_pairs = (M + 2 * G_BM - 1) / (2 * G_BM), f_tiles = F / G_BF;
  const int max_pairs = sms / 2;
  // feature blocks per work item (x tile resident for the run): only when K is exactly G_STAGES k-blocks (C = 320, the sites
  // with the most rows, which are otherwise bound by the bytes crossing L2 -> SM: 128 flop per byte at K = 320).  Pick the
  // divisor of f_tiles with the fewest estimated tile-times (measured: a tile with its x resident is ~7 % shorter, 208 -> 194 us
  // at 32 samples; fewer, longer work items also remove a partial wave at small M: 19.1 -> 18.2 us at 2 samples).
  int f_run = 1;
  if (K == G_STAGES * G_BK) {
    static const int forced = [] { const char* e = getenv("FD_K13_FRUN"); return e ? atoi(e) : 0; }();
    double best = 1e30;
    for (int r = 1; r <= f_tiles; ++r) {
      if (f_tiles % r) continue;
      const long items = static_cast<long>(m_pairs) * (f_tiles / r);
      const double cost = static_cast<double>((items + max_pairs - 1) / max_pairs) * r * (1.0 - 0.07 * (1.0 - 1.0 / r));
      if (cost < best - 1e-9) { best = cost; f_run = r; }
    }
    if (forced > 0 && f_tiles % forced == 0) f_run = forced;
  }
  const int tiles = m_pairs * (f_tiles / f_run);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(2 * (tiles < max_pairs ? tiles : max_pairs)));
  cfg.blockDim = dim3(G_THREADS);
  cfg.dynamicSmemBytes = G_SMEM;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FD_CUDA_OK(cudaLaunchKernelEx(&cfg, k13_ff_geglu_kernel, tm_x, tm_w, tm_out, static_cast<const __nv_bfloat16*>(bias_bf16_dev), F,
                                K, m_pairs, f_tiles, f_run));
  return FD_OK;
}

// development aid: first bounded wait of K13 that gave up (0 = none); clears it.  Synchronises the device.
extern "C" int fd_debug_k13_flag(void) {
  int v = 0, z = 0;
  cudaMemcpyFromSymbol(&v, fd::g_k13_flag, sizeof(int));
  cudaMemcpyToSymbol(fd::g_k13_flag, &z, sizeof(int));
  return v;
}
