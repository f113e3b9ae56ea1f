// K3 -- cross-attention of the latent tokens over the CACHED K/V of the fixed 77-token context.
//
// Replaces diffusers' CrossAttention._attention (einsum QK^T * scale -> softmax -> . V, which
// materialises a [2B*8, N, 77] score tensor per layer), reached from pipeline/guide.py:56-58
// for each of the 16 attn2 layers of the SD-v1 UNet.  K and V are column
// slices of the K2 output, so nothing about the context is recomputed inside the loop.
//
// v3 (round 1; v1 was a serial TMA -> MMA -> softmax -> MMA -> store chain, v2 warp-specialised
// it with ONE softmax warpgroup, whose ~550 instructions per row at one warp per scheduler were
// the serial resource: ~1 us per tile per CTA against a 0.33 us MUFU floor).  A CTA owns one
// (sample, head) and walks over several 128-query tiles with TWO softmax warpgroups:
//   warp 0        : TMA producer -- K and V once, then the Q tiles through a QSTAGES ring
//   warp 1        : tcgen05.mma issuer -- S_i = Q_i K^T, then O_{i-1} = P_{i-1} V
//   warps 2..5    : warpgroup A, even tiles      } one TMEM lane = one query row: two passes over
//   warps 6..9    : warpgroup B, odd tiles       } the 80 logits (row max, then exp2 / sum / bf16 P
//                   written back over S as the A operand of P V), 16 columns at a time so the
//                   kernel stays under 96 registers and two CTAs (16 softmax warps) share an SM;
//                   packed FFMA2 / FADD2 / FMUL2 and FMNMX3 halve the non-MUFU issue slots.
//   S = Q K^T : M=128, N=80, K=d (k-steps of 16)   both operands K-major SWIZZLE_128B in smem
//   O = P V   : M=128, N=d (rounded to 16), K=80   A = P from TMEM, B = V MN-major in smem
// TMEM: S_A [0,80), S_B [80,160) fp32 (P = bf16x2 in the first 40 columns of its S buffer), then
// one O accumulator per warpgroup (d = 40: 256 columns, two CTAs per SM; d = 80 / 160: 512).
// Output: the epilogue writes O / rowsum as bf16 into the tile's own Q stage (same SWIZZLE_128B
// layout; the stage is idle once S is complete) and the producer warp sends it out with one TMA
// tensor store per 64-column chunk, clipped at n_q / d by the map.  (Per-thread 16-byte global
// stores -- 32 half-filled sectors per warp instruction -- backed the LSU up until the SM moved
// ~1 tile/us no matter how many warps it had.)  HBM-bound by design (AI = 77 FLOP/B).
#include "fd_common.cuh"

// tuning knobs (profiles/build_variants.py builds A/B libraries with -D overrides); the variants that
// lost their A/B -- FMA-pipe 2^x for part of the keys, pass-1 row max in two TMEM round trips, a
// column-split one-pass softmax, a flat tile scheduler -- are recorded in profiles/r01/SUMMARY.md
#ifndef K3_FIRST_TILES
#define K3_FIRST_TILES 1   // Q tiles requested up front; the rest of the ring follows once tile 0 has landed
                           // (8 x 4096 x 320: 4 -> 17.2 us, 2 -> 16.6, 1 -> 16.4; 8 x 1024 x 640: 8.1 / 7.6 / 7.2)
#endif
#ifndef K3_QSTAGES_SMALL
#define K3_QSTAGES_SMALL 4 // Q ring depth for d <= 128 (2: 19.6 us, 3: 17.6, 4: 17.0, 5: 17.5 at 8 x 4096 x 320)
#endif

namespace fd {
namespace {

constexpr int TQ = 128;     // query rows per tile
constexpr int TKV = 80;     // keys padded 77 -> 80
constexpr int K3_THREADS = 320;
constexpr int Q_CHUNK_BYTES = TQ * 128;
constexpr int KV_CHUNK_BYTES = TKV * 128;
constexpr int MAX_QSTAGES = 5;

template <int DH>
struct K3Cfg {
  static constexpr int NCHUNK = (DH + 63) / 64;        // 64-element d chunks (128 B swizzle rows)
  static constexpr int KSTEPS = (DH + 15) / 16;        // UMMA k-steps for Q K^T
  static constexpr int NPV = ((DH + 15) / 16) * 16;    // UMMA N for P V
  static constexpr int S_COLS = TKV;                   // one S / P buffer per warpgroup
  static constexpr int O_BASE = 2 * S_COLS;
  static constexpr int TMEM_COLS = O_BASE + 2 * NPV <= 256 ? 256 : 512;
  // a Q stage stays busy from its load until its tile's output has been read out by the TMA store
  static constexpr int QSTAGES = DH <= 128 ? K3_QSTAGES_SMALL : 3;
  static constexpr int Q_STAGE_BYTES = NCHUNK * Q_CHUNK_BYTES;
  static constexpr int SMEM = 1024 + QSTAGES * Q_STAGE_BYTES + 2 * NCHUNK * KV_CHUNK_BYTES + 256;
#ifdef K3_D40_ONE_CTA
  static constexpr int CTAS_PER_SM = 1;
#else
  static constexpr int CTAS_PER_SM = (TMEM_COLS == 256 && SMEM <= 112 * 1024) ? 2 : 1;
#endif
  // with the whole SM's registers for one CTA the 80 logits of a row stay in registers (one TMEM
  // round trip); two CTAs per SM cap the kernel at 96 registers and make two passes over TMEM
  static constexpr bool ONE_PASS = CTAS_PER_SM == 1;
  static_assert(SMEM <= 227 * 1024, "K3 shared memory");
};

// development aid: when set (fd_debug_set_k3_timing), %globaltimer (ns) stamps go to a device buffer:
// [0] start, [1] setup done, [7] end of CTA (0,0,0); [32 + 8 * tile + ev] per-tile events of the
// traced CTA (fd_debug_set_k3_trace_cta), ev = 0 MMA thread waits for Q, 1 S issued, 2 group waits
// for S, 3 S seen, 4 P written, 5 P V issued, 6 O seen, 7 epilogue done; [160 + 3 * cta + {0,1,2}]
// start, end and SM id of every CTA (profiles/k3_phase_timing.py prints them)
static long long* g_k3_timing = nullptr;
static int g_k3_timing_tile = 0;  // linear id of the CTA whose per-tile events are traced

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define K3_STAMP(slot)                                                                  \
  do {                                                                                  \
    if (a.timing && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) a.timing[slot] = gtime(); \
  } while (0)
// per-tile events of the traced CTA (linear id g_k3_timing_tile... see K3_EV_*): slot 32 + 8 * tile + ev
#define K3_EVENT(tile, ev)                                                              \
  do {                                                                                  \
    if (a.timing && cta_linear == a.timing_tile && (tile) < 16) a.timing[32 + 8 * (tile) + (ev)] = gtime(); \
  } while (0)

struct K3Args {
  long long* timing;
  const int32_t* ctx_index;
  __nv_bfloat16* out;
  int n_q, heads, t_valid, t_pad, n_tiles, timing_tile;
  float scale_log2e;
};

template <int DH>
__global__ void __launch_bounds__(K3_THREADS, K3Cfg<DH>::CTAS_PER_SM)
k3_cross_attn_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                     const K3Args a) {
  using Cfg = K3Cfg<DH>;
  constexpr int QS = Cfg::QSTAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sq = smem;                                   // QSTAGES x [NCHUNK x 128 x 128 B]
  uint8_t* sk = sq + QS * Cfg::Q_STAGE_BYTES;           // NCHUNK x 80 x 128 B
  uint8_t* sv = sk + Cfg::NCHUNK * KV_CHUNK_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sv + Cfg::NCHUNK * KV_CHUNK_BYTES);
  uint64_t* kv_full = bars;                             // K + V landed
  uint64_t* q_full = bars + 1;                          // [QS] Q stage landed
  uint64_t* out_full = q_full + MAX_QSTAGES;            // [QS] output tile staged in the Q stage
  uint64_t* s_full = out_full + MAX_QSTAGES;             // [2] S = Q K^T complete (per warpgroup)
  uint64_t* p_full = s_full + 2;                        // [2] P written by the 4 warps of the group
  uint64_t* o_full = p_full + 2;                        // [2] O = P V complete
  uint64_t* o_free = o_full + 2;                        // [2] epilogue has drained O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int sample = blockIdx.z;
  if (threadIdx.x == 0) K3_STAMP(0);
  int ctx_row = 0;
  if (threadIdx.x == 0) ctx_row = __ldg(a.ctx_index + sample) * a.t_pad;  // in flight during the barrier setup
  const int cta_linear = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (a.timing && threadIdx.x == 0) {  // per-CTA trace: start, end, SM id after the 16 phase slots
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    a.timing[160 + 3 * cta_linear] = gtime();
    a.timing[160 + 3 * cta_linear + 2] = smid;
  }
  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int my_tiles = (a.n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                       static_cast<int>(gridDim.x);

  // Q tiles requested up front: with every CTA asking for its whole ring at once the first tile of
  // the last CTAs queues behind ~18 MB of other CTAs' prefetches
  constexpr int FIRST = K3_FIRST_TILES < QS ? K3_FIRST_TILES : QS;
  const int n_first = my_tiles < FIRST ? my_tiles : FIRST;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_o);
    mbar_init(kv_full, 1);
    for (int s = 0; s < QS; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&out_full[s], 4);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);
      mbar_init(&o_full[s], 1);
      mbar_init(&o_free[s], 4);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  auto first_loads = [&]() {
    // Q first, K / V need the context row from global memory
    for (int i = 0; i < n_first; ++i) {
      mbar_expect_tx(&q_full[i], Cfg::Q_STAGE_BYTES);
      const int q0 = (static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x)) * TQ;
#pragma unroll
      for (int c = 0; c < Cfg::NCHUNK; ++c)
        tma_load_4d(sq + i * Cfg::Q_STAGE_BYTES + c * Q_CHUNK_BYTES, &tm_q, &q_full[i], c * 64, head, q0, sample);
    }
    mbar_expect_tx(kv_full, 2 * Cfg::NCHUNK * KV_CHUNK_BYTES);
#pragma unroll
    for (int c = 0; c < Cfg::NCHUNK; ++c) {
      tma_load_3d(sk + c * KV_CHUNK_BYTES, &tm_k, kv_full, c * 64, head, ctx_row);
      tma_load_3d(sv + c * KV_CHUNK_BYTES, &tm_v, kv_full, c * 64, head, ctx_row);
    }
  };
  // the first loads go out while warp 1 allocates TMEM and the CTA synchronises (2-3 % per launch)
  if (threadIdx.x == 0) first_loads();
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) K3_STAMP(1);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (threadIdx.x == 0) {
      auto store_tile = [&](int j) {  // output of tile j: staged by its warpgroup -> TMA store
        const int s = j % QS;
        mbar_wait(&out_full[s], (j / QS) & 1);
        const int q0 = (static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x)) * TQ;
#pragma unroll
        for (int c = 0; c < Cfg::NCHUNK; ++c)
          tma_store_4d(&tm_o, sq + s * Cfg::Q_STAGE_BYTES + c * Q_CHUNK_BYTES, c * 64, head, q0, sample);
        tma_store_commit();
      };
      if (n_first < my_tiles && n_first < QS) mbar_wait(&q_full[0], 0);  // tile 0 is in: fill the ring
      for (int i = n_first; i < my_tiles; ++i) {
        const int s = i % QS;
        if (i >= QS) {
          store_tile(i - QS);
          tma_store_wait_read<0>();  // the stage may be overwritten once the store has read it
        }
        mbar_expect_tx(&q_full[s], Cfg::Q_STAGE_BYTES);
        const int q0 = (static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x)) * TQ;
#pragma unroll
        for (int c = 0; c < Cfg::NCHUNK; ++c)
          tma_load_4d(sq + s * Cfg::Q_STAGE_BYTES + c * Q_CHUNK_BYTES, &tm_q, &q_full[s], c * 64, head, q0,
                      sample);
      }
      for (int j = my_tiles > QS ? my_tiles - QS : 0; j < my_tiles; ++j) store_tile(j);
      tma_store_wait_all();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc(UMMA_BF16, TQ, TKV, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc(UMMA_BF16, TQ, Cfg::NPV, 0, 1);  // B (= V) MN-major
      auto issue_pv = [&](int j) {
        const int b = j & 1;
        mbar_wait(&p_full[b], (j >> 1) & 1);
        // the previous user of this O accumulator (tile j-2) must have been drained by its epilogue
        if (j >= 2) mbar_wait(&o_free[b], ((j >> 1) - 1) & 1);
        tc_fence_after();
        K3_EVENT(j, 5);  // P V issue
        const uint32_t pbuf = tmem_base + b * Cfg::S_COLS;
        const uint32_t obuf = tmem_base + Cfg::O_BASE + b * Cfg::NPV;
#pragma unroll
        for (int k = 0; k < TKV / 16; ++k) {
          // 16 keys = two 8-row groups (SBO = 1024 B); next 64-wide d chunk LBO = KV_CHUNK_BYTES away
          const uint64_t vd = umma_desc_sw128(smem_u32(sv + k * 2048), KV_CHUNK_BYTES, 1024);
          mma_f16_ts(obuf, pbuf + 8 * k, vd, idesc_o, k != 0);
        }
        tc_commit(&o_full[b]);
      };
      mbar_wait(kv_full, 0);
      for (int i = 0; i < my_tiles; ++i) {
        const int b = i & 1;
        const int st = i % QS;
        K3_EVENT(i, 0);  // MMA thread starts waiting for Q
        mbar_wait(&q_full[st], (i / QS) & 1);
        K3_EVENT(i, 1);  // Q landed, S issue
        // S buffer b was last read (as P) by P V of tile i-2, issued earlier by this thread: the
        // tensor pipe executes in issue order, so no extra wait is needed before overwriting it
        tc_fence_after();
        const uint32_t buf = tmem_base + b * Cfg::S_COLS;
        const uint8_t* qs = sq + st * Cfg::Q_STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
          const int c = ks >> 2, kk = ks & 3;
          const uint64_t qd = umma_desc_sw128(smem_u32(qs + c * Q_CHUNK_BYTES), 16, 1024) + 2 * kk;
          const uint64_t kd = umma_desc_sw128(smem_u32(sk + c * KV_CHUNK_BYTES), 16, 1024) + 2 * kk;
          mma_f16_ss(buf, qd, kd, idesc_s, ks != 0);
        }
        tc_commit(&s_full[b]);
        if (i >= 1) issue_pv(i - 1);
      }
      if (my_tiles >= 1) issue_pv(my_tiles - 1);
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue warpgroups
    const int wg = (warp - 2) >> 2;  // 0: even tiles, 1: odd tiles
    const int quarter = warp & 3;    // TMEM lanes [32*quarter, +32)
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t sbuf = tmem_base + wg * Cfg::S_COLS + lane_addr;
    const uint32_t obuf = tmem_base + Cfg::O_BASE + wg * Cfg::NPV + lane_addr;
    const bool stamp = lane == 0 && quarter == 0;

    auto epilogue = [&](int j, float sum) {
      mbar_wait(&o_full[wg], (j >> 1) & 1);
      tc_fence_after();
      if (stamp) K3_EVENT(j, 6);  // O seen
      const int st = j % QS;
      uint8_t* stage = sq + st * Cfg::Q_STAGE_BYTES;  // Q of this tile was consumed by S(j): reuse it
      const float inv = rcp_approx(sum);
      const uint64_t inv2 = f2_pack(inv, inv);
      // EPI_CH 16-column loads per tcgen05.wait::ld (a wait drains every outstanding load, so the
      // number of waits is the number of TMEM round trips): 48 / 64 columns per round trip
      constexpr int NCH = Cfg::NPV / 16;
      constexpr int EPI_CH = Cfg::ONE_PASS ? 4 : 3;
#pragma unroll
      for (int c0 = 0; c0 < NCH; c0 += EPI_CH) {
        uint32_t v[EPI_CH][16];
#pragma unroll
        for (int g = 0; g < EPI_CH; ++g)
          if (c0 + g < NCH) tmem_ld_x16(obuf + 16 * (c0 + g), v[g]);
        tmem_ld_wait();
        if (c0 + EPI_CH >= NCH) {
          // O is in registers: hand the accumulator back before the smem / store part
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_free[wg]);
        }
#pragma unroll
        for (int g = 0; g < EPI_CH; ++g) {
          const int c = c0 + g;
          if (c < NCH) {
            uint32_t pk[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float lo, hi;
              f2_unpack(f2_mul(f2_pack(__uint_as_float(v[g][2 * k]), __uint_as_float(v[g][2 * k + 1])), inv2), lo, hi);
              pk[k] = pack_bf16x2(lo, hi);
            }
            // columns [16c, 16c+16) = 16-byte units 2(c%4), 2(c%4)+1 of 64-column chunk c/4 (SWIZZLE_128B)
            uint8_t* chunk = stage + (c >> 2) * Q_CHUNK_BYTES;
            *reinterpret_cast<uint4*>(chunk + sw128_offset(row, 2 * (c & 3))) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(chunk + sw128_offset(row, 2 * (c & 3) + 1)) =
                make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&out_full[st]);
      if (stamp) K3_EVENT(j, 7);  // epilogue done
    };

    const uint64_t scale2 = f2_pack(a.scale_log2e, a.scale_log2e);
    for (int i = wg; i < my_tiles; i += 2) {
      if (stamp) K3_EVENT(i, 2);  // group starts waiting for S
      mbar_wait(&s_full[wg], (i >> 1) & 1);
      tc_fence_after();
      if (stamp) K3_EVENT(i, 3);  // S seen
      uint64_t acc0 = f2_pack(0.f, 0.f), acc1 = acc0;
      if constexpr (Cfg::ONE_PASS) {
        // ---- all 80 logits in registers: one TMEM round trip
        uint32_t v[TKV / 16][16];
#pragma unroll
        for (int c = 0; c < TKV / 16; ++c) tmem_ld_x16(sbuf + 16 * c, v[c]);
        tmem_ld_wait();
        // keys >= t_valid only exist in the last 16 columns (t_valid > 64 is checked on the host)
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (TKV - 16 + j >= a.t_valid) v[TKV / 16 - 1][j] = 0xff800000u;  // -inf
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int c = 0; c < TKV / 16; ++c)
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            m0 = fmax3(m0, __uint_as_float(v[c][j]), __uint_as_float(v[c][j + 1]));
            m1 = fmax3(m1, __uint_as_float(v[c][j + 2]), __uint_as_float(v[c][j + 3]));
          }
        const float nmx = -fmaxf(m0, m1) * a.scale_log2e;
        const uint64_t nmx2 = f2_pack(nmx, nmx);
#pragma unroll
        for (int c = 0; c < TKV / 16; ++c) {
          uint32_t pk[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float t0, t1;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(v[c][2 * k]), __uint_as_float(v[c][2 * k + 1])), scale2, nmx2),
                      t0, t1);
            const float e0 = ex2_approx(t0), e1 = ex2_approx(t1);
            if (k & 1) acc1 = f2_add(acc1, f2_pack(e0, e1));
            else acc0 = f2_add(acc0, f2_pack(e0, e1));
            pk[k] = pack_bf16x2(e0, e1);
          }
          tmem_st_x8(sbuf + 8 * c, pk);
        }
      } else {
        // ---- pass 1: row max over the 80 logits, 16 columns per step, next load in flight
        float mx;
        {
          uint32_t v[2][16];
          tmem_ld_x16(sbuf, v[0]);
          float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
          for (int c = 0; c < TKV / 16; ++c) {
            tmem_ld_wait();
            if (c + 1 < TKV / 16) tmem_ld_x16(sbuf + 16 * (c + 1), v[(c + 1) & 1]);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float x0 = __uint_as_float(v[c & 1][j]), x1 = __uint_as_float(v[c & 1][j + 1]);
              float x2 = __uint_as_float(v[c & 1][j + 2]), x3 = __uint_as_float(v[c & 1][j + 3]);
              if (c == TKV / 16 - 1) {
                if (16 * c + j >= a.t_valid) x0 = -INFINITY;
                if (16 * c + j + 1 >= a.t_valid) x1 = -INFINITY;
                if (16 * c + j + 2 >= a.t_valid) x2 = -INFINITY;
                if (16 * c + j + 3 >= a.t_valid) x3 = -INFINITY;
              }
              m0 = fmax3(m0, x0, x1);
              m1 = fmax3(m1, x2, x3);
            }
          }
          mx = fmaxf(m0, m1);
        }
        // ---- pass 2: exp2(s * scale*log2e - max * scale*log2e) (one FFMA2 per two keys + one MUFU per
        // key; scale > 0), row sum, bf16 P stored over S; 16 columns per step, next load in flight
        const float nmx = -mx * a.scale_log2e;
        const uint64_t nmx2 = f2_pack(nmx, nmx);
        {
          uint32_t v[2][16];
          tmem_ld_x16(sbuf, v[0]);
#pragma unroll
          for (int c = 0; c < TKV / 16; ++c) {
            tmem_ld_wait();
            if (c + 1 < TKV / 16) tmem_ld_x16(sbuf + 16 * (c + 1), v[(c + 1) & 1]);
            uint32_t pk[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float x0 = __uint_as_float(v[c & 1][2 * k]), x1 = __uint_as_float(v[c & 1][2 * k + 1]);
              if (c == TKV / 16 - 1) {
                if (16 * c + 2 * k >= a.t_valid) x0 = -INFINITY;
                if (16 * c + 2 * k + 1 >= a.t_valid) x1 = -INFINITY;
              }
              float t0, t1;
              f2_unpack(f2_fma(f2_pack(x0, x1), scale2, nmx2), t0, t1);
              const float e0 = ex2_approx(t0), e1 = ex2_approx(t1);
              if (k & 1) acc1 = f2_add(acc1, f2_pack(e0, e1));
              else acc0 = f2_add(acc0, f2_pack(e0, e1));
              pk[k] = pack_bf16x2(e0, e1);
            }
            tmem_st_x8(sbuf + 8 * c, pk);
          }
        }
      }
      float sa, sb;
      f2_unpack(f2_add(acc0, acc1), sa, sb);
      const float sum = sa + sb;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[wg]);
      if (stamp) K3_EVENT(i, 4);  // P written
      // epilogue right behind this tile's P V: the group's O accumulator is free again before its
      // next softmax ends, so P V(i) never waits and S(i+2) is issued straight after it (a deferred
      // epilogue put  softmax(i) -> epilogue(i-2) -> P V(i) -> S(i+2)  on one serial chain)
      epilogue(i, sum);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
  if (threadIdx.x == 0) K3_STAMP(7);
  if (a.timing && threadIdx.x == 0) a.timing[160 + 3 * cta_linear + 1] = gtime();
}

template <int DH>
int launch_k3(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
              const K3Args& a, dim3 grid,
              cudaStream_t st) {
  using Cfg = K3Cfg<DH>;
  static thread_local int attr_device = -1;  // the attribute is per device; set it once per thread/device
  int dev = 0;
  FD_CUDA_OK(cudaGetDevice(&dev));
  if (attr_device != dev) {
    FD_CUDA_OK(cudaFuncSetAttribute(k3_cross_attn_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_device = dev;
  }
  k3_cross_attn_kernel<DH><<<grid, K3_THREADS, Cfg::SMEM, st>>>(tq, tk, tv, to, a);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}

}  // namespace
}  // namespace fd

// development aid (not part of the product ABI): device buffer of >= 16 + 3 * #CTAs int64 for phase timestamps and the per-CTA trace
extern "C" void fd_debug_set_k3_timing(void* buf_dev) { fd::g_k3_timing = static_cast<long long*>(buf_dev); }
extern "C" void fd_debug_set_k3_trace_cta(int cta_linear) { fd::g_k3_timing_tile = cta_linear; }

extern "C" int fd_cross_attn(const void* q_bf16_dev, const void* kv_bf16_dev, int64_t kv_rows, int64_t kv_row_stride,
                             int k_col_off, int v_col_off, const int32_t* ctx_index_dev, int n_samples, int n_q,
                             int heads, int d_head, int t_valid, int t_pad, float scale, void* out_bf16_dev,
                             void* stream) {
  using namespace fd;
  FD_REQUIRE(q_bf16_dev && kv_bf16_dev && ctx_index_dev && out_bf16_dev, "fd_cross_attn: NULL pointer");
  FD_REQUIRE(n_samples > 0 && n_q > 0 && heads > 0, "fd_cross_attn: non-positive shape");
  FD_REQUIRE(d_head == 40 || d_head == 80 || d_head == 160, "fd_cross_attn: d_head=%d not in {40, 80, 160}", d_head);
  FD_REQUIRE(t_pad == TKV && t_valid > TKV - 16 && t_valid <= t_pad,
             "fd_cross_attn: need t_pad == %d and %d < t_valid <= t_pad", TKV, TKV - 16);
  FD_REQUIRE(scale > 0.f, "fd_cross_attn: scale must be positive");
  FD_REQUIRE(k_col_off % 8 == 0 && v_col_off % 8 == 0 && kv_row_stride % 8 == 0,
             "fd_cross_attn: column offsets / row stride must be multiples of 8 elements");
  FD_REQUIRE(k_col_off + heads * d_head <= kv_row_stride && v_col_off + heads * d_head <= kv_row_stride,
             "fd_cross_attn: K/V slice exceeds the cache row");
  FD_REQUIRE(kv_rows % t_pad == 0, "fd_cross_attn: kv_rows=%lld not a multiple of t_pad", (long long)kv_rows);
  FD_REQUIRE(reinterpret_cast<uintptr_t>(q_bf16_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(kv_bf16_dev) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(out_bf16_dev) % 16 == 0,
             "fd_cross_attn: pointers must be 16-byte aligned");
  FD_REQUIRE(heads <= 65535 && n_samples <= 65535, "fd_cross_attn: heads / n_samples exceed grid limits");
  int rc = check_device();
  if (rc != FD_OK) return rc;

  const uint64_t C = static_cast<uint64_t>(heads) * d_head;
  CUtensorMap tq, tk, tv, to;
  for (int which = 0; which < 2; ++which) {
    uint64_t dims[4] = {static_cast<uint64_t>(d_head), static_cast<uint64_t>(heads), static_cast<uint64_t>(n_q),
                        static_cast<uint64_t>(n_samples)};
    uint64_t strides[3] = {static_cast<uint64_t>(d_head) * 2, C * 2, static_cast<uint64_t>(n_q) * C * 2};
    uint32_t box[4] = {64, 1, TQ, 1};
    rc = encode_tmap(which == 0 ? &tq : &to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                     which == 0 ? q_bf16_dev : static_cast<const void*>(out_bf16_dev), dims, strides, box,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
  }
  for (int which = 0; which < 2; ++which) {
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(kv_bf16_dev) + (which == 0 ? k_col_off : v_col_off);
    uint64_t dims[3] = {static_cast<uint64_t>(d_head), static_cast<uint64_t>(heads), static_cast<uint64_t>(kv_rows)};
    uint64_t strides[2] = {static_cast<uint64_t>(d_head) * 2, static_cast<uint64_t>(kv_row_stride) * 2};
    uint32_t box[3] = {64, 1, TKV};
    rc = encode_tmap(which == 0 ? &tk : &tv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
  }
  const int sms = sm_count();
  if (sms <= 0) return set_error(FD_ERR_CUDA, "fd_cross_attn: cannot query SM count");
  K3Args a;
  a.timing = g_k3_timing;
  a.timing_tile = g_k3_timing_tile;
  a.ctx_index = ctx_index_dev;
  a.out = static_cast<__nv_bfloat16*>(out_bf16_dev);
  a.n_q = n_q;
  a.heads = heads;
  a.t_valid = t_valid;
  a.t_pad = t_pad;
  a.n_tiles = (n_q + TQ - 1) / TQ;
  a.scale_log2e = scale * 1.4426950408889634f;
  // CTAs per (sample, head): fill the resident capacity once (no second wave), each CTA walking
  // over >= 1 query tile
  const int pairs = heads * n_samples;
  const int per_sm = d_head == 160 ? K3Cfg<160>::CTAS_PER_SM : (d_head == 80 ? K3Cfg<80>::CTAS_PER_SM : K3Cfg<40>::CTAS_PER_SM);
  int per_pair = (per_sm * sms) / pairs;
  if (per_pair > a.n_tiles) per_pair = a.n_tiles;
  if (per_pair < 1) per_pair = 1;
  dim3 grid(per_pair, heads, n_samples);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d_head == 40) return launch_k3<40>(tq, tk, tv, to, a, grid, st);
  if (d_head == 80) return launch_k3<80>(tq, tk, tv, to, a, grid, st);
  return launch_k3<160>(tq, tk, tv, to, a, grid, st);
}
