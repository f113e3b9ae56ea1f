// K2 -- cross-attention K/V projection of the fixed 77-token context, hoisted out of the
// denoising loop: ONE batched GEMM per prompt batch instead of 32 small Linears x 16 layers x
// every step (diffusers CrossAttention.forward `to_k(context)`, `to_v(context)`, reached from
// pipeline/guide.py:56-58).
//
//   out[M, N] = ctx[M, K] . w[N, K]^T      bf16 operands, fp32 accumulate, bf16 out
//   M = n_ctx * 80 (context rows, zero padded 77 -> 80), K = 768, N = 24960 (all to_k|to_v rows)
//
// v2: persistent, warp-specialised tcgen05 GEMM, one CTA per SM walking over 128 x 256 tiles
// (m fastest, so the CTAs running together share one weight tile in L2):
//   warp 0     : TMA producer, 4-stage mbarrier ring of {A 128x64, B 256x64} SWIZZLE_128B tiles
//   warp 1     : single-thread tcgen05.mma issuer (M128 N256 K16), accumulator in TMEM
//   warps 2-5  : epilogue (tcgen05.ld -> bf16 -> swizzled smem staging -> TMA store)
// The accumulator is double-buffered (2 x 256 TMEM columns), so the epilogue of tile i runs under
// the mainloop of tile i+1.  v3: when the number of m tiles is even the kernel runs as 2-CTA clusters
// whose CTAs share one weight tile through TMA multicast (see PAIR below).  v1 (one 128x128 tile per CTA, no overlap) was bound by L2->SM operand
// traffic at 64 FLOP/B per tile (profiles/r01/SUMMARY.md); 128x256 tiles need 25 % fewer bytes.
#include "fd_common.cuh"

namespace fd {
namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int B_STAGE_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int K2_THREADS = 192;
constexpr int OUT_CHUNK_BYTES = BM * 64 * 2;  // epilogue staging: 128 rows x 64 bf16, SWIZZLE_128B, two buffers
constexpr int K2_SMEM = STAGES * STAGE_BYTES + 2 * OUT_CHUNK_BYTES + 1024 /*align*/ + 256 /*bars*/;
constexpr int TMEM_COLS = 2 * BN;

// PAIR: launched as 2-CTA clusters; the two CTAs of a pair work on tiles (2q, 2q+1) = the same
// weight (B) tile and adjacent m tiles.  Each CTA TMA-loads its own A tile and HALF of the B tile,
// multicast into both CTAs' shared memory, which halves the L2 -> SM traffic of the dominant
// operand.  A stage may only be refilled once BOTH CTAs' MMAs have consumed it, so the MMA warps
// commit to the empty barrier of both CTAs (count 2).
template <bool PAIR>
__global__ void __launch_bounds__(K2_THREADS, 1)
k2_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
               const __grid_constant__ CUtensorMap tm_out, int M, int N, int K, int m_tiles, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* out_stage = smem + STAGES * STAGE_BYTES;  // 2 x OUT_CHUNK_BYTES
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_stage + 2 * OUT_CHUNK_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = K / BK;
  const int total_tiles = m_tiles * n_tiles;
  const uint32_t crank = PAIR ? cluster_cta_rank() : 0;
  // tile walk: single CTAs stride over tiles; pairs stride over tile pairs
  const int first_tile = PAIR ? 2 * static_cast<int>(blockIdx.x >> 1) + static_cast<int>(crank) : static_cast<int>(blockIdx.x);
  const int tile_stride = PAIR ? 2 * static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_out);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], PAIR ? 2 : 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_barrier();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;  // running k-block counter across tiles: stage = it % STAGES
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
        const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          tma_load_2d(smem + s * STAGE_BYTES, &tm_a, &full_bar[s], kb * BK, m_blk * BM);
          if (PAIR) {
            // my half of the weight tile (128 of its 256 rows), delivered to both CTAs
            tma_load_2d_mcast(smem + s * STAGE_BYTES + A_STAGE_BYTES + crank * (B_STAGE_BYTES / 2), &tm_b,
                              &full_bar[s], kb * BK, n_blk * BN + static_cast<int>(crank) * (BN / 2), 0x3);
          } else {
            tma_load_2d(smem + s * STAGE_BYTES + A_STAGE_BYTES, &tm_b, &full_bar[s], kb * BK, n_blk * BN);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(UMMA_BF16, BM, BN, 0, 0);
      int it = 0, local = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++local) {
        const int acc = local & 1;
        mbar_wait(&tmem_empty[acc], ((local >> 1) & 1) ^ 1);  // epilogue drained this buffer
        tc_fence_after();
        const uint32_t d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem + s * STAGE_BYTES), 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + s * STAGE_BYTES + A_STAGE_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)  // 16 bf16 = 32 B along K inside the swizzle row: +2
            mma_f16_ss(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          if (PAIR) tc_commit_mcast(&empty_bar[s], 0x3);  // both producers write into this stage
          else tc_commit(&empty_bar[s]);                  // smem slot reusable once these MMAs retire
        }
        tc_commit(&tmem_full[acc]);  // accumulator complete
      }
    }
  } else {
    // epilogue: warp w may only touch TMEM lanes [32*(w%4), +32).  Rows go to a SWIZZLE_128B staging
    // tile in shared memory and leave through TMA stores (full 128-byte lines, clipped at the
    // matrix edges) -- a thread-per-row store pattern writes half-filled 32-byte sectors.
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const bool issuer = threadIdx.x == 64;  // first epilogue thread issues the TMA stores
    int local = 0, chunk_no = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++local) {
      const int m_blk = tile % m_tiles, n_blk = tile / m_tiles;
      const int acc = local & 1;
      mbar_wait(&tmem_full[acc], (local >> 1) & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_base + acc * BN + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 64, ++chunk_no) {
        uint32_t v[4][16];
#pragma unroll
        for (int g = 0; g < 4; ++g) tmem_ld_x16(tbase + c0 + 16 * g, v[g]);
        tmem_ld_wait();
        uint8_t* buf = out_stage + (chunk_no & 1) * OUT_CHUNK_BYTES;
        // the store that last read this buffer (two chunks ago) must have finished reading it
        if (issuer) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(v[g][8 * h + 2 * j]),
                                                       __uint_as_float(v[g][8 * h + 2 * j + 1]));
              pk[j] = *reinterpret_cast<uint32_t*>(&b);
            }
            *reinterpret_cast<uint4*>(buf + sw128_offset(row, 2 * g + h)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (issuer) {
          tma_store_2d(&tm_out, buf, n_blk * BN + c0, m_blk * BM);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (issuer) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_barrier();  // no CTA exits while its peer may still multicast into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// v4: cta_group::2.  One tcgen05.mma spans a PAIR of CTAs: M = 256 (128 context rows in each CTA's shared
// memory and TMEM) x N = 256, and each CTA holds only HALF of the weight tile (128 of its 256 rows).  What
// bounds v3 is the bytes an SM has to take in per MMA cycle (A 16 KB + B 32 KB per 512 tensor cycles =
// 94 B / clk against ~50-64 B / clk an SM ingests, measured in profiles/r02/SUMMARY.md: multicast saves L2 reads,
// not SM ingest); with the pair sharing B every CTA takes in 16 + 16 KB per 512 cycles = 62 B / clk, and the
// 32 KB stages allow a 6-deep ring.  Leader CTA (rank 0): arms the stage barrier with both CTAs' bytes and
// issues the MMAs; both CTAs run their own TMA producer (their A rows + their half of B, signalled on the
// leader's barrier) and their own epilogue (their 128 rows of the accumulator).
constexpr int STAGES2 = 5;
constexpr int K2_THREADS2 = 320;       // TMA, MMA, 8 epilogue warps (two column halves x four lane quarters)
constexpr int OUT_BUFS2 = 4;           // two staging buffers per epilogue half
constexpr int B2_STAGE_BYTES = B_STAGE_BYTES / 2;
constexpr int STAGE2_BYTES = A_STAGE_BYTES + B2_STAGE_BYTES;  // 32768
constexpr int K2_SMEM2 = STAGES2 * STAGE2_BYTES + OUT_BUFS2 * OUT_CHUNK_BYTES + 1024 /*align*/ + 256 /*bars*/;
__device__ int g_k2_flag;  // first bounded wait that gave up (0 = none)

__global__ void __launch_bounds__(K2_THREADS2, 1)
k2_gemm2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_bh,
                const __grid_constant__ CUtensorMap tm_out, int M, int N, int K, int m_pairs, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* out_stage = smem + STAGES2 * STAGE2_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_stage + OUT_BUFS2 * OUT_CHUNK_BYTES);
  uint64_t* empty_bar = full_bar + STAGES2;
  uint64_t* tmem_full = empty_bar + STAGES2;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2] (leader's copy is the one that counts: 2 x 8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = K / BK;
  const int total_tiles = m_pairs * n_tiles;
  const uint32_t crank = cluster_cta_rank();
  const bool leader = crank == 0;
  const int first_tile = static_cast<int>(blockIdx.x >> 1);
  const int tile_stride = static_cast<int>(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_bh);
    tma_prefetch_desc(&tm_out);
    for (int s = 0; s < STAGES2; ++s) {
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
    tmem_alloc_2cta(tmem_slot, TMEM_COLS);
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
        const int m_blk = tile % m_pairs, n_blk = tile / m_pairs;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES2;
          mbar_wait_bounded(&empty_bar[s], ((it / STAGES2) & 1) ^ 1, &g_k2_flag, 1);
          if (leader) mbar_expect_tx(&full_bar[s], 2 * STAGE2_BYTES);
          const uint32_t lead_bar = mapa_rank(smem_u32(&full_bar[s]), 0);
          tma_load_2d_2cta(smem + s * STAGE2_BYTES, &tm_a, lead_bar, kb * BK, m_blk * 2 * BM + static_cast<int>(crank) * BM);
          tma_load_2d_2cta(smem + s * STAGE2_BYTES + A_STAGE_BYTES, &tm_bh, lead_bar, kb * BK,
                           n_blk * BN + static_cast<int>(crank) * (BN / 2));
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc(UMMA_BF16, 2 * BM, BN, 0, 0);
      int it = 0, local = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++local) {
        const int acc = local & 1;
        mbar_wait_bounded(&tmem_empty[acc], ((local >> 1) & 1) ^ 1, &g_k2_flag, 2);  // both epilogues drained it
        tc_fence_after();
        const uint32_t d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES2;
          mbar_wait_bounded(&full_bar[s], (it / STAGES2) & 1, &g_k2_flag, 3);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem + s * STAGE2_BYTES), 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + s * STAGE2_BYTES + A_STAGE_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            mma_f16_ss_2cta(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          tc_commit_2cta_mcast(&empty_bar[s], 0x3);  // both producers may refill the stage
        }
        tc_commit_2cta_mcast(&tmem_full[acc], 0x3);  // both epilogues may read their half
      }
    }
  } else {
    // 8 epilogue warps: lane quarter = warp % 4 (TMEM access rule), column half = (warp - 2) / 4.  Each half
    // drains its 128 columns in two 64-column chunks through its own pair of staging buffers, its own named
    // barrier and its own TMA-store issuer, so the two halves never wait for each other.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const bool issuer = (threadIdx.x - 64) % 128 == 0;  // first thread of each half
    uint8_t* my_stage = out_stage + half * 2 * OUT_CHUNK_BYTES;
    int local = 0, chunk_no = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++local) {
      const int m_blk = tile % m_pairs, n_blk = tile / m_pairs;
      const int acc = local & 1;
      mbar_wait_bounded(&tmem_full[acc], (local >> 1) & 1, &g_k2_flag, 4);
      tc_fence_after();
      const uint32_t tbase = tmem_base + acc * BN + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 64, ++chunk_no) {
        uint32_t v[4][16];
#pragma unroll
        for (int g = 0; g < 4; ++g) tmem_ld_x16(tbase + c0 + 16 * g, v[g]);
        tmem_ld_wait();
        uint8_t* buf = my_stage + (chunk_no & 1) * OUT_CHUNK_BYTES;
        if (issuer) tma_store_wait_read<1>();
        named_bar_sync(1 + half, 128);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(v[g][8 * h + 2 * j]),
                                                       __uint_as_float(v[g][8 * h + 2 * j + 1]));
              pk[j] = *reinterpret_cast<uint32_t*>(&b);
            }
            *reinterpret_cast<uint4*>(buf + sw128_offset(row, 2 * g + h)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + half, 128);
        if (issuer) {
          tma_store_2d(&tm_out, buf, n_blk * BN + c0, m_blk * 2 * BM + static_cast<int>(crank) * BM);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa_rank(smem_u32(&tmem_empty[acc]), 0));  // on the LEADER's barrier
    }
    if (issuer) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  cluster_barrier();  // no CTA exits (or frees TMEM) while its peer may still signal / multiply into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
}

// v5: clusters of SIX CTAs = three cta_group::2 pairs that work on the same 256-column weight tile and three adjacent
// 256-row context blocks (M = 720: the whole context).  v4 at N = 256 x M = 256 per pair pulls 223 MB through L2 for
// the 9-context shape (every weight tile three times, the context once per weight tile), which at the ~6300 B / clk
// the L2 slices deliver chip-wide (B300_MICROARCH "LTS cap"; K13's large shapes sit exactly on it) is 18 us of its 25.
// Here each CTA loads a THIRD of its pair-half of the weight tile (48 / 40 / 40 rows) and TMA-multicasts it to the same-rank
// CTAs of all three pairs, so a weight byte leaves L2 once: 153 MB.  A stage is refilled only when all three pairs have
// consumed it (every pair leader's tcgen05.commit is multicast to the six empty barriers, count 3).  The multicast's
// complete_tx lands on the barrier at the same offset in the destination CTA's PAIR LEADER (.cta_group::2 form, peer bit of the
// barrier address cleared).  Only 22 such clusters fit on the 148 SMs (132 SMs, profiles/microbench/cluster_occupancy.cu).
// MEASURED (profiles/k2_bench.py, warm L2): correct at the first run, 26.4 us vs v4 23.6 us at M = 720 (47.5 vs 43.8 at M = 1360):
// a tile runs 10 % faster (5.3 vs 5.9 us) but 98 tiles on 22 clusters are 5 waves instead of 4 -- the bytes each SM takes in
// per k-block (32 KB, ~35-44 B / clk achieved) are unchanged by multicast, and that, not the L2 read count, is the bound.
// Kept as an opt-in variant (fd_debug_set_k2_variant(3)) for the comparison; v4 stays the product path.
constexpr int K2_PAIRS3 = 3;
constexpr int K2_CLUSTER3 = 2 * K2_PAIRS3;
__device__ __forceinline__ void tma_load_2d_2cta_mcast(void* dst, const CUtensorMap* map, uint32_t bar_addr, int c0, int c1,
                                                       uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_addr), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}

__global__ void __launch_bounds__(K2_THREADS2, 1)
k2_gemm3_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b48,
                const __grid_constant__ CUtensorMap tm_b40, const __grid_constant__ CUtensorMap tm_out, int K, int m_groups,
                int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* out_stage = smem + STAGES2 * STAGE2_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_stage + OUT_BUFS2 * OUT_CHUNK_BYTES);
  uint64_t* empty_bar = full_bar + STAGES2;
  uint64_t* tmem_full = empty_bar + STAGES2;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2] (the pair leader's copy counts: 2 x 8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = K / BK;
  const int total_tiles = m_groups * n_tiles;
  const uint32_t crank = cluster_cta_rank();       // 0..5
  const uint32_t q = crank >> 1, r = crank & 1;    // pair within the cluster, rank within the pair
  const bool leader = r == 0;
  const int first_tile = static_cast<int>(blockIdx.x / K2_CLUSTER3);
  const int tile_stride = static_cast<int>(gridDim.x / K2_CLUSTER3);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b48);
    tma_prefetch_desc(&tm_b40);
    tma_prefetch_desc(&tm_out);
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], K2_PAIRS3);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 16);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_barrier();  // all six CTAs' barriers and TMEM exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // this CTA's third of the pair-half of the weight tile: rows [b_off, b_off + b_rows) of the 128
      const int b_off = q == 0 ? 0 : (q == 1 ? 48 : 88);
      const CUtensorMap* tm_b = q == 0 ? &tm_b48 : &tm_b40;
      const uint16_t mask = static_cast<uint16_t>(0x15u << r);  // the same-rank CTAs of the three pairs
      int it = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
        const int m_blk = (tile % m_groups) * K2_PAIRS3 + static_cast<int>(q), n_blk = tile / m_groups;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES2;
          mbar_wait_bounded(&empty_bar[s], ((it / STAGES2) & 1) ^ 1, &g_k2_flag, 11);  // all three pairs consumed it
          if (leader) mbar_expect_tx(&full_bar[s], 2 * STAGE2_BYTES);
          const uint32_t lead_bar = mapa_rank(smem_u32(&full_bar[s]), 2 * q);
          tma_load_2d_2cta(smem + s * STAGE2_BYTES, &tm_a, lead_bar, kb * BK, m_blk * 2 * BM + static_cast<int>(r) * BM);
          tma_load_2d_2cta_mcast(smem + s * STAGE2_BYTES + A_STAGE_BYTES + b_off * 128, tm_b,
                                 smem_u32(&full_bar[s]) & 0xFEFFFFFFu, kb * BK,
                                 n_blk * BN + static_cast<int>(r) * (BN / 2) + b_off, mask);
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc(UMMA_BF16, 2 * BM, BN, 0, 0);
      const uint16_t pair_mask = static_cast<uint16_t>(0x3u << (2 * q));
      int it = 0, local = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++local) {
        const int acc = local & 1;
        mbar_wait_bounded(&tmem_empty[acc], ((local >> 1) & 1) ^ 1, &g_k2_flag, 12);  // both epilogues drained it
        tc_fence_after();
        const uint32_t d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES2;
          mbar_wait_bounded(&full_bar[s], (it / STAGES2) & 1, &g_k2_flag, 13);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem + s * STAGE2_BYTES), 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + s * STAGE2_BYTES + A_STAGE_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            mma_f16_ss_2cta(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          tc_commit_2cta_mcast(&empty_bar[s], 0x3F);  // one of the three arrivals every CTA of the cluster waits for
        }
        tc_commit_2cta_mcast(&tmem_full[acc], pair_mask);  // both epilogues of this pair may read their half
      }
    }
  } else {
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const bool issuer = (threadIdx.x - 64) % 128 == 0;  // first thread of each half
    uint8_t* my_stage = out_stage + half * 2 * OUT_CHUNK_BYTES;
    int local = 0, chunk_no = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++local) {
      const int m_blk = (tile % m_groups) * K2_PAIRS3 + static_cast<int>(q), n_blk = tile / m_groups;
      const int acc = local & 1;
      mbar_wait_bounded(&tmem_full[acc], (local >> 1) & 1, &g_k2_flag, 14);
      tc_fence_after();
      const uint32_t tbase = tmem_base + acc * BN + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 64, ++chunk_no) {
        uint32_t v[4][16];
#pragma unroll
        for (int g = 0; g < 4; ++g) tmem_ld_x16(tbase + c0 + 16 * g, v[g]);
        tmem_ld_wait();
        uint8_t* buf = my_stage + (chunk_no & 1) * OUT_CHUNK_BYTES;
        if (issuer) tma_store_wait_read<1>();
        named_bar_sync(1 + half, 128);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(v[g][8 * h + 2 * j]),
                                                       __uint_as_float(v[g][8 * h + 2 * j + 1]));
              pk[j] = *reinterpret_cast<uint32_t*>(&b);
            }
            *reinterpret_cast<uint4*>(buf + sw128_offset(row, 2 * g + h)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + half, 128);
        if (issuer) {
          tma_store_2d(&tm_out, buf, n_blk * BN + c0, m_blk * 2 * BM + static_cast<int>(r) * BM);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa_rank(smem_u32(&tmem_empty[acc]), 2 * q));  // the pair LEADER's barrier
    }
    if (issuer) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  cluster_barrier();  // no CTA exits (or frees TMEM) while a CTA of the cluster may still signal / write / multiply into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
}

int g_k2_variant = 2;  // 2 = v4 cta_group::2 pairs when M > 128 (default), 1 = v3, 3 = v5 six-CTA clusters when M > 512 (measured slower)

}  // namespace
}  // namespace fd

// development aids: kernel variant, and the bounded-wait record of the cta_group::2 kernel (0 = clean)
extern "C" void fd_debug_set_k2_variant(int v) { fd::g_k2_variant = v; }
extern "C" int fd_debug_k2_flag(void) {
  int v = -1;
  cudaMemcpyFromSymbol(&v, fd::g_k2_flag, sizeof(int));
  int z = 0;
  cudaMemcpyToSymbol(fd::g_k2_flag, &z, sizeof(int));
  return v;
}

extern "C" int fd_kv_project(const void* ctx_bf16_dev, const void* w_bf16_dev, void* out_bf16_dev,
                             int M, int N, int K, void* stream) {
  using namespace fd;
  FD_REQUIRE(ctx_bf16_dev && w_bf16_dev && out_bf16_dev, "fd_kv_project: NULL pointer");
  FD_REQUIRE(M > 0 && N > 0 && K > 0, "fd_kv_project: non-positive shape %d %d %d", M, N, K);
  FD_REQUIRE(K % BK == 0, "fd_kv_project: K=%d must be a multiple of %d", K, BK);
  FD_REQUIRE(N % 8 == 0, "fd_kv_project: N=%d must be a multiple of 8", N);
  FD_REQUIRE(reinterpret_cast<uintptr_t>(ctx_bf16_dev) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(w_bf16_dev) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(out_bf16_dev) % 16 == 0,
             "fd_kv_project: pointers must be 16-byte aligned");
  int rc = check_device();
  if (rc != FD_OK) return rc;

  CUtensorMap tm_a, tm_b;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {BK, BM};
    rc = encode_tmap(&tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ctx_bf16_dev, dims, strides, box,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {BK, BN};
    rc = encode_tmap(&tm_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_bf16_dev, dims, strides, box,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
  }
  CUtensorMap tm_out;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(N) * 2};
    uint32_t box[2] = {64, BM};
    rc = encode_tmap(&tm_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out_bf16_dev, dims, strides, box,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
  }
  static thread_local int attr_device = -1;
  int dev = 0;
  FD_CUDA_OK(cudaGetDevice(&dev));
  if (attr_device != dev) {
    FD_CUDA_OK(cudaFuncSetAttribute(k2_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, K2_SMEM));
    FD_CUDA_OK(cudaFuncSetAttribute(k2_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K2_SMEM));
    attr_device = dev;
  }
  const int sms = sm_count();
  if (sms <= 0) return set_error(FD_ERR_CUDA, "fd_kv_project: cannot query SM count");
  const int m_tiles = (M + BM - 1) / BM, n_tiles = (N + BN - 1) / BN;
  const int total = m_tiles * n_tiles;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (g_k2_variant == 3 && m_tiles >= 5) {
    // v5: six-CTA clusters, three pairs x 256 rows; m groups of 768 rows (OOB rows: TMA zero-fills and clips)
    CUtensorMap tm_b48, tm_b40;
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box48[2] = {BK, 48}, box40[2] = {BK, 40};
    rc = encode_tmap(&tm_b48, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_bf16_dev, dims, strides, box48, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
    rc = encode_tmap(&tm_b40, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_bf16_dev, dims, strides, box40, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
    static thread_local int attr3_device = -1;
    static thread_local int max_clusters = 0;
    const int m_groups = (m_tiles + 2 * K2_PAIRS3 - 1) / (2 * K2_PAIRS3);
    const int tiles3 = m_groups * n_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(K2_CLUSTER3));
    cfg.blockDim = dim3(K2_THREADS2);
    cfg.dynamicSmemBytes = K2_SMEM2;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = K2_CLUSTER3;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (attr3_device != dev) {
      FD_CUDA_OK(cudaFuncSetAttribute(k2_gemm3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K2_SMEM2));
      FD_CUDA_OK(cudaOccupancyMaxActiveClusters(&max_clusters, k2_gemm3_kernel, &cfg));
      attr3_device = dev;
    }
    if (max_clusters >= 1) {
      cfg.gridDim = dim3(static_cast<unsigned>(K2_CLUSTER3 * (tiles3 < max_clusters ? tiles3 : max_clusters)));
      FD_CUDA_OK(cudaLaunchKernelEx(&cfg, k2_gemm3_kernel, tm_a, tm_b48, tm_b40, tm_out, K, m_groups, n_tiles));
      return FD_OK;
    }
  }
  if (g_k2_variant >= 2 && m_tiles >= 2 && sms >= 2) {
    // cta_group::2 path: pairs of CTAs own 256 rows (the last pair may be partly or, for an odd number of
    // m tiles, half empty: TMA zero-fills and clips)
    CUtensorMap tm_bh;
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {BK, BN / 2};
    rc = encode_tmap(&tm_bh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_bf16_dev, dims, strides, box,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
    static thread_local int attr2_device = -1;
    if (attr2_device != dev) {
      FD_CUDA_OK(cudaFuncSetAttribute(k2_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K2_SMEM2));
      attr2_device = dev;
    }
    const int m_pairs = (m_tiles + 1) / 2;
    const int tiles2 = m_pairs * n_tiles;
    const int max_pairs = sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(2 * (tiles2 < max_pairs ? tiles2 : max_pairs)));
    cfg.blockDim = dim3(K2_THREADS2);
    cfg.dynamicSmemBytes = K2_SMEM2;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FD_CUDA_OK(cudaLaunchKernelEx(&cfg, k2_gemm2_kernel, tm_a, tm_bh, tm_out, M, N, K, m_pairs, n_tiles));
    return FD_OK;
  }
  if (m_tiles % 2 == 0 && sms >= 2) {
    // paired path: tiles 2q and 2q+1 share their weight tile (m is the fast index and m_tiles is even)
    CUtensorMap tm_bh;  // half-height box of the weight tile: each CTA of a pair loads 128 of the 256 rows
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {BK, BN / 2};
    rc = encode_tmap(&tm_bh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_bf16_dev, dims, strides, box,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
    int pairs = total / 2, max_pairs = sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(2 * (pairs < max_pairs ? pairs : max_pairs)));
    cfg.blockDim = dim3(K2_THREADS);
    cfg.dynamicSmemBytes = K2_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FD_CUDA_OK(cudaLaunchKernelEx(&cfg, k2_gemm_kernel<true>, tm_a, tm_bh, tm_out, M, N, K, m_tiles, n_tiles));
    return FD_OK;
  }
  k2_gemm_kernel<false><<<total < sms ? total : sms, K2_THREADS, K2_SMEM, st>>>(tm_a, tm_b, tm_out, M, N, K, m_tiles,
                                                                               n_tiles);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}
