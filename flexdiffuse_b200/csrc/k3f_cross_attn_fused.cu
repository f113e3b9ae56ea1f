// K3F -- the WHOLE attn2 layer in one launch:  out = to_out( softmax( to_q(x) K^T * scale ) V ) + bias
//
// Replaces, per cross-attention site, the three launches K3 left behind (cuBLAS to_q, K3, cuBLAS
// to_out) of diffusers' CrossAttention.forward, reached from pipeline/guide.py:56-58 for each of the
// 16 attn2 layers of the SD-v1 UNet (SURVEY 2.3 K3, 8d "fused variant": bare 77-key attention is
// AI = 77 FLOP/B and HBM-bound by construction; with both projections inside, the kernel reads the
// hidden states once, writes them once and is tensor-bound).  K and V still come from the K2 cache.
//
// Decomposition.  A CTA owns (sample, 128-query tile, 320-channel head group g): 8 heads of d = 40,
// 4 of d = 80 or 2 of d = 160.  The G = C / 320 CTAs of one query tile form a thread-block CLUSTER
// (G = 1, 2, 4), because to_out contracts over all C channels:
//   P1  Q_g[128 x 320]   = X[128 x C] . Wq[g]^T     TMA ring {X 128x64, Wq 320x64} -> tcgen05.mma
//                                                    (2 x N=160 per k-step), fp32 in TMEM [0, 320)
//   P2  Q -> bf16, compacted IN PLACE in TMEM (the A operand of S comes from TMEM, like P in K3)
//   P3  per head: S = Q_h K_h^T (A from TMEM), one-pass softmax with one TMEM lane per query row,
//       O_h = P V_h (P from TMEM, V MN-major from the K2 cache layout), O_h / rowsum -> bf16 ->
//       swizzled staging -> TMA store into the [S, N, C] attention-output buffer (two softmax
//       warpgroups alternate heads, exactly K3 v3's pipeline with heads in the role of tiles)
//   P4  (G > 1) the producers of the cluster rendezvous on a remote mbarrier: every head group of
//       this query tile is in global memory (L2)
//   P5  out_g[128 x 320] = O[128 x C] . Wo[g]^T      same ring, same MMAs, accumulator [0, 320)
//   P6  + bias -> bf16 -> swizzled staging -> TMA store (clipped at n_q)
// Shared memory is ONE 224 KB arena used three ways: P1 / P5 the 4-stage operand ring, P3 the K / V
// tiles of the group's heads + two output staging buffers, P6 the [128 x 320] output staging.
// TMEM (512 columns): Q fp32 [0,320) -> Q bf16 [0, HG*KQ/2) + O accumulators behind it; S / P
// buffers [320, 480); the P5 accumulator reuses [0, 320).
//
// Every mbarrier wait is bounded (watchdog on %globaltimer): a protocol error ends the kernel with
// a code in fd_debug_k3f_status() instead of hanging the device.
#include "fd_common.cuh"

namespace fd {
namespace {

constexpr int F_TQ = 128;                 // query rows per tile
constexpr int F_TKV = 80;                 // keys padded 77 -> 80
constexpr int F_NQ = 320;                 // channels (= heads x d) per CTA
constexpr int F_NB = 160;                 // UMMA N (two MMAs per k-step)
constexpr int F_THREADS = 320;
constexpr int F_STAGES = 4;
constexpr int F_A_BYTES = F_TQ * 128;     // 128 rows x 64 bf16, SWIZZLE_128B
constexpr int F_B_BYTES = F_NQ * 128;     // 320 rows x 64 bf16
constexpr int F_STAGE_BYTES = F_A_BYTES + F_B_BYTES;
constexpr int F_ARENA = F_STAGES * F_STAGE_BYTES;   // 229376
constexpr int F_KV_CHUNK = F_TKV * 128;
constexpr int F_Q_CHUNK = F_TQ * 128;
constexpr int F_TAIL = 2048;              // barriers + TMEM slot + bias
constexpr int F_SMEM = 1024 + F_ARENA + F_TAIL;
constexpr int F_TMEM_COLS = 512;
constexpr int F_S_BASE = 320;
constexpr int F_ACC_BASE = 192;            // to_out accumulator: TMEM [192, 512)
static_assert(F_SMEM <= 227 * 1024, "K3F shared memory");

template <int DH>
struct FCfg {
  static constexpr int HG = F_NQ / DH;                   // heads per CTA
  static constexpr int NCHUNK = (DH + 63) / 64;          // 64-wide d chunks of K / V / O staging
  static constexpr int KSTEPS = (DH + 15) / 16;          // UMMA k-steps of Q K^T
  static constexpr int KQ2 = KSTEPS * 8;                 // TMEM columns of one head's bf16 Q
  static constexpr int NPV = KSTEPS * 16;                // UMMA N of P V
  static constexpr int QB_END = HG * KQ2;
  static constexpr int NOBUF = (QB_END + 2 * NPV <= F_NQ) ? 2 : 1;   // O accumulators
  static constexpr int O_BASE = QB_END;
  static constexpr int KV_HEAD_BYTES = 2 * NCHUNK * F_KV_CHUNK;
  static constexpr int KV_BYTES = HG * KV_HEAD_BYTES;
  static constexpr int OST_BYTES = NCHUNK * F_Q_CHUNK;   // one warpgroup's output staging
  static_assert(HG * DH == F_NQ, "d_head must divide 320");
  static_assert(KQ2 <= DH, "in-place bf16 compaction needs KQ/2 <= d");
  static_assert(QB_END + NOBUF * NPV <= F_NQ, "TMEM layout");
  static_assert(KV_BYTES + 2 * OST_BYTES <= F_ARENA, "P3 arena");
  static_assert(KV_BYTES < (1 << 20), "mbarrier tx count");
};

// [0] first failing wait code, [1] block id, [2] extra
__device__ int g_k3f_status[4];
// development aid: %globaltimer stamps of CTA (0,0,0) when FArgs::trace is set (fd_debug_k3f_trace):
// [0] start [1] setup done [2] q_done seen [3] conversion done [4] K/V landed (MMA thread)
// [8+4j..] head j: S seen, P written, O seen, staged   [48] stores complete [49] rendezvous done
// [50] out_full seen [51] end
__device__ long long g_k3f_trace[64];

struct FArgs {
  const int32_t* ctx_index;
  const __nv_bfloat16* bias;
  int n_q, t_valid, t_pad, kchunks, G, phases, trace, store_attn;
  float scale_log2e;
};

__device__ __forceinline__ long long f_gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// bounded wait: returns when the phase completes, when another thread of the CTA has given up, or
// after ~0.25 s (recording `code`)
template <bool CLUSTER_SCOPE = false>
__device__ __forceinline__ void wd_wait(uint64_t* bar, uint32_t parity, int code, volatile int* abort_s) {
  uint32_t done = 0;
  long long t0 = 0;
  uint32_t spins = 0;
  while (true) {
    if (CLUSTER_SCOPE) {
      asm volatile(
          "{\n\t.reg .pred P;\n\t"
          "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, P;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(bar)), "r"(parity)
          : "memory");
    } else {
      asm volatile(
          "{\n\t.reg .pred P;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, P;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(bar)), "r"(parity)
          : "memory");
    }
    if (done) return;
    if ((++spins & 255u) == 0) {
      if (*abort_s) return;
      const long long t = f_gtime();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 250000000LL) {
        *abort_s = 1;
        if (atomicCAS(&g_k3f_status[0], 0, code) == 0) {
          g_k3f_status[1] = static_cast<int>(blockIdx.x | (blockIdx.y << 4) | (blockIdx.z << 20));
          g_k3f_status[2] = static_cast<int>(parity);
        }
        return;
      }
    }
  }
}

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void remote_mbar_arrive(uint64_t* bar, uint32_t cta_rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta_rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

#define F_STAMP(slot)                                                                          \
  do {                                                                                         \
    if (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_k3f_trace[slot] = f_gtime(); \
  } while (0)

enum : int {
  W_EMPTY = 1, W_FULL = 2, W_QDONE_P = 3, W_QDONE_W = 4, W_KV = 5, W_QC = 6, W_SFULL = 7, W_PFULL = 8,
  W_OFULL = 9, W_OFREE = 10, W_OSTFULL = 11, W_OSTFREE = 12, W_OREADY = 13, W_OUTFULL = 14,
  W_EMPTY5 = 15, W_FULL5 = 16, W_OBF = 17
};

// P6 helper: NG 16-column groups of the fp32 accumulator (from group c_first) + bias -> bf16 -> the
// SWIZZLE_128B staging chunks at the head of the arena (one TMEM round trip for the NG groups)
template <int NG>
__device__ __forceinline__ void p6_groups(uint32_t tacc, int c_first, const float* bias_s, uint8_t* arena, int row) {
  uint32_t v[NG][16];
#pragma unroll
  for (int q = 0; q < NG; ++q) tmem_ld_x16(tacc + 16 * (c_first + q), v[q]);
  tmem_ld_wait();
#pragma unroll
  for (int q = 0; q < NG; ++q) {
    const int c = c_first + q;
    uint32_t pk[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      pk[k] = pack_bf16x2(__uint_as_float(v[q][2 * k]) + bias_s[16 * c + 2 * k],
                          __uint_as_float(v[q][2 * k + 1]) + bias_s[16 * c + 2 * k + 1]);
    uint8_t* chunk = arena + (c >> 2) * F_Q_CHUNK;
    *reinterpret_cast<uint4*>(chunk + sw128_offset(row, 2 * (c & 3))) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(chunk + sw128_offset(row, 2 * (c & 3) + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

template <int DH>
__global__ void __launch_bounds__(F_THREADS, 1)
k3f_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_wq,
           const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
           const __grid_constant__ CUtensorMap tm_os, const __grid_constant__ CUtensorMap tm_oa,
           const __grid_constant__ CUtensorMap tm_wo, const __grid_constant__ CUtensorMap tm_out,
           const FArgs a) {
  using Cfg = FCfg<DH>;
  constexpr int HG = Cfg::HG;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* arena = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~static_cast<uintptr_t>(1023));
  uint8_t* kv_s = arena;                                   // P3: HG x {K, V} x NCHUNK x [80 x 128 B]
  uint8_t* ost_s = arena + Cfg::KV_BYTES;                  // P3: 2 x NCHUNK x [128 x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(arena + F_ARENA);
  uint64_t* full = bars;                 // [4]
  uint64_t* empty = full + F_STAGES;     // [4]
  uint64_t* q_done = empty + F_STAGES;   // Q accumulator complete
  uint64_t* k_full = q_done + 1;         // [8] K tile of head j landed
  uint64_t* v_full = k_full + 8;         // [8] V tile of head j landed
  uint64_t* qc_done = v_full + 8;        // bf16 Q in place (8 warps)
  uint64_t* s_full = qc_done + 1;        // [2]
  uint64_t* p_full = s_full + 2;         // [2] (4 warps)
  uint64_t* o_full = p_full + 2;         // [2]
  uint64_t* o_free = o_full + 2;         // [2] (4 warps)
  uint64_t* ost_full = o_free + 2;       // [2] (4 warps)
  uint64_t* ost_free = ost_full + 2;     // [2] (producer)
  uint64_t* o_ready = ost_free + 2;      // every CTA of the cluster has stored its heads (G arrivals)
  uint64_t* out_full = o_ready + 1;
  uint64_t* obf_done = out_full + 1;     // every head's bf16 output is in TMEM (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(obf_done + 1);
  volatile int* abort_s = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* bias_s = reinterpret_cast<float*>(arena + F_ARENA + 512);   // [320]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x;              // head group == rank in the cluster
  const int row0 = blockIdx.y * F_TQ;
  const int sample = blockIdx.z;
  const int KC = a.kchunks;
  if (threadIdx.x == 0) F_STAMP(0);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_wq);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_os);
    tma_prefetch_desc(&tm_oa);
    tma_prefetch_desc(&tm_wo);
    tma_prefetch_desc(&tm_out);
    for (int s = 0; s < F_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(q_done, 1);
    for (int j = 0; j < 8; ++j) {
      mbar_init(&k_full[j], 1);
      mbar_init(&v_full[j], 1);
    }
    mbar_init(qc_done, 8);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);
      mbar_init(&o_full[s], 1);
      mbar_init(&o_free[s], 4);
      mbar_init(&ost_full[s], 4);
      mbar_init(&ost_free[s], 1);
    }
    mbar_init(o_ready, static_cast<uint32_t>(a.G));
    mbar_init(out_full, 1);
    mbar_init(obf_done, 8);
    *abort_s = 0;
    fence_mbar_init();
    fence_proxy_async_smem();
    // the first F_STAGES k-chunks of P1 go out now, under the TMEM allocation and the CTA barrier
    // (their barriers were initialised by this very thread; nothing else touches the ring yet)
    for (int kc = 0; kc < F_STAGES && kc < KC; ++kc) {
      uint8_t* st = arena + kc * F_STAGE_BYTES;
      mbar_expect_tx(&full[kc], F_STAGE_BYTES);
      tma_load_3d(st, &tm_x, &full[kc], kc * 64, row0, sample);
      tma_load_2d(st + F_A_BYTES, &tm_wq, &full[kc], kc * 64, g * F_NQ);
      tma_load_2d(st + F_A_BYTES + F_NB * 128, &tm_wq, &full[kc], kc * 64, g * F_NQ + F_NB);
    }
  }
  if (threadIdx.x >= 64) {  // this group's to_out bias as fp32
    const int i = threadIdx.x - 64;
    bias_s[i] = __bfloat162float(a.bias[g * F_NQ + i]);
    if (i < F_NQ - 256) bias_s[256 + i] = __bfloat162float(a.bias[g * F_NQ + 256 + i]);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, F_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (a.G > 1) cluster_barrier();  // the peers' o_ready barriers exist before anybody signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) F_STAMP(1);

  if (warp == 0) {
    // ================================================================== TMA producer
    if (threadIdx.x == 0) {
      int it = 0;
      auto gemm_loads = [&](const CUtensorMap* ta, const CUtensorMap* tb, int wcode) {
        for (int kc = 0; kc < KC; ++kc, ++it) {
          if (it < F_STAGES) continue;  // issued before the CTA barrier
          const int s = it % F_STAGES;
          wd_wait(&empty[s], ((it / F_STAGES) & 1) ^ 1, wcode, abort_s);
          uint8_t* st = arena + s * F_STAGE_BYTES;
          mbar_expect_tx(&full[s], F_STAGE_BYTES);
          tma_load_3d(st, ta, &full[s], kc * 64, row0, sample);
          tma_load_2d(st + F_A_BYTES, tb, &full[s], kc * 64, g * F_NQ);
          tma_load_2d(st + F_A_BYTES + F_NB * 128, tb, &full[s], kc * 64, g * F_NQ + F_NB);
        }
      };
      // ---- P1
      gemm_loads(&tm_x, &tm_wq, W_EMPTY);
      if (a.phases & 2) {
        // ---- K / V of this group's heads: the ring is dead once the Q accumulator is complete
        wd_wait(q_done, 0, W_QDONE_P, abort_s);
        const int ctx_row = __ldg(a.ctx_index + sample) * a.t_pad;
        // one barrier per tile, in the order the MMA thread needs them (K_0, V_0, K_1, ...): the
        // first head starts as soon as ITS 2 x NCHUNK x 10 KB are in, the rest lands under the softmax
#pragma unroll 1
        for (int j = 0; j < HG; ++j) {
          uint8_t* kj = kv_s + j * Cfg::KV_HEAD_BYTES;
          uint8_t* vj = kj + Cfg::NCHUNK * F_KV_CHUNK;
          mbar_expect_tx(&k_full[j], Cfg::NCHUNK * F_KV_CHUNK);
#pragma unroll
          for (int c = 0; c < Cfg::NCHUNK; ++c)
            tma_load_3d(kj + c * F_KV_CHUNK, &tm_k, &k_full[j], c * 64, g * HG + j, ctx_row);
          mbar_expect_tx(&v_full[j], Cfg::NCHUNK * F_KV_CHUNK);
#pragma unroll
          for (int c = 0; c < Cfg::NCHUNK; ++c)
            tma_load_3d(vj + c * F_KV_CHUNK, &tm_v, &v_full[j], c * 64, g * HG + j, ctx_row);
        }
        // ---- P3: ship each head's output tile as its warpgroup stages it (only when somebody reads
        // it from global memory: the other head groups of the cluster, or the caller)
        if (a.store_attn) {
#pragma unroll 1
          for (int j = 0; j < HG; ++j) {
            const int w = j & 1;
            wd_wait(&ost_full[w], (j >> 1) & 1, W_OSTFULL, abort_s);
#pragma unroll
            for (int c = 0; c < Cfg::NCHUNK; ++c)
              tma_store_4d(&tm_os, ost_s + w * Cfg::OST_BYTES + c * F_Q_CHUNK, c * 64, g * HG + j, row0, sample);
            tma_store_commit();
            tma_store_wait_read<0>();
            mbar_arrive(&ost_free[w]);
          }
          tma_store_wait_all();  // writes performed, not just read
          fence_proxy_async_all();
        }
        F_STAMP(48);
      }
      if (a.phases & 4) {
        // ---- P4 / P5.  Chunk order: this group's own 5 k-chunks first -- their A operand is the bf16
        // attention output sitting in TMEM, so only Wo travels -- then the other groups' chunks,
        // whose A operand comes back from global memory once the whole cluster has stored its heads.
        const bool own_tmem = (a.phases & 2) != 0;
        if (a.G > 1) {
          __threadfence();
          for (int p = 0; p < a.G; ++p) remote_mbar_arrive(o_ready, static_cast<uint32_t>(p));
        }
        if (own_tmem) wd_wait(obf_done, 0, W_OBF, abort_s);  // K / V and staging are dead: the ring may refill
        for (int i = 0; i < KC; ++i, ++it) {
          const int o = i - 5;
          const int kc = i < 5 ? g * 5 + i : (o < g * 5 ? o : o + 5);
          const bool a_from_tmem = own_tmem && i < 5;
          if (i == 5) {
            wd_wait<true>(o_ready, 0, W_OREADY, abort_s);
            fence_proxy_async_all();
            F_STAMP(49);
          }
          const int s = it % F_STAGES;
          wd_wait(&empty[s], ((it / F_STAGES) & 1) ^ 1, W_EMPTY5, abort_s);
          uint8_t* st = arena + s * F_STAGE_BYTES;
          mbar_expect_tx(&full[s], a_from_tmem ? F_B_BYTES : F_STAGE_BYTES);
          if (!a_from_tmem) tma_load_3d(st, &tm_oa, &full[s], kc * 64, row0, sample);
          tma_load_2d(st + F_A_BYTES, &tm_wo, &full[s], kc * 64, g * F_NQ);
          tma_load_2d(st + F_A_BYTES + F_NB * 128, &tm_wo, &full[s], kc * 64, g * F_NQ + F_NB);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_g = umma_idesc(UMMA_BF16, F_TQ, F_NB, 0, 0);
      constexpr uint32_t idesc_s = umma_idesc(UMMA_BF16, F_TQ, F_TKV, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc(UMMA_BF16, F_TQ, Cfg::NPV, 0, 1);  // B (= V) MN-major
      int it = 0;
      auto gemm_mmas = [&](int wcode) {
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int s = it % F_STAGES;
          wd_wait(&full[s], (it / F_STAGES) & 1, wcode, abort_s);
          tc_fence_after();
          const uint32_t st = smem_u32(arena + s * F_STAGE_BYTES);
          const uint64_t ad = umma_desc_sw128(st, 16, 1024);
          const uint64_t b0 = umma_desc_sw128(st + F_A_BYTES, 16, 1024);
          const uint64_t b1 = umma_desc_sw128(st + F_A_BYTES + F_NB * 128, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            mma_f16_ss(tmem_base, ad + 2 * k, b0 + 2 * k, idesc_g, (kc | k) != 0);
            mma_f16_ss(tmem_base + F_NB, ad + 2 * k, b1 + 2 * k, idesc_g, (kc | k) != 0);
          }
          tc_commit(&empty[s]);
        }
      };
      // ---- P1
      gemm_mmas(W_FULL);
      tc_commit(q_done);
      if (a.phases & 2) {
        // ---- P3
        auto issue_pv = [&](int j) {
          const int b = j & 1;
          const int ob = Cfg::NOBUF == 2 ? b : 0;
          wd_wait(&v_full[j], 0, W_KV, abort_s);
          wd_wait(&p_full[b], (j >> 1) & 1, W_PFULL, abort_s);
          if (j >= Cfg::NOBUF) wd_wait(&o_free[ob], ((j / Cfg::NOBUF) - 1) & 1, W_OFREE, abort_s);
          tc_fence_after();
          const uint32_t pbuf = tmem_base + F_S_BASE + b * F_TKV;
          const uint32_t obuf = tmem_base + Cfg::O_BASE + ob * Cfg::NPV;
          const uint32_t sv = smem_u32(kv_s + j * Cfg::KV_HEAD_BYTES + Cfg::NCHUNK * F_KV_CHUNK);
#pragma unroll
          for (int k = 0; k < F_TKV / 16; ++k) {
            const uint64_t vd = umma_desc_sw128(sv + k * 2048, F_KV_CHUNK, 1024);
            mma_f16_ts(obuf, pbuf + 8 * k, vd, idesc_o, k != 0);
          }
          tc_commit(&o_full[b]);  // one barrier per warpgroup even when they share the accumulator
        };
        wd_wait(qc_done, 0, W_QC, abort_s);
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < HG; ++j) {
          const int b = j & 1;
          wd_wait(&k_full[j], 0, W_KV, abort_s);
          if (j == 0) F_STAMP(4);
          const uint32_t sbuf = tmem_base + F_S_BASE + b * F_TKV;
          const uint32_t qb = tmem_base + j * Cfg::KQ2;
          const uint32_t sk = smem_u32(kv_s + j * Cfg::KV_HEAD_BYTES);
#pragma unroll
          for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
            const int c = ks >> 2, kk = ks & 3;
            const uint64_t kd = umma_desc_sw128(sk + c * F_KV_CHUNK, 16, 1024) + 2 * kk;
            mma_f16_ts(sbuf, qb + 8 * ks, kd, idesc_s, ks != 0);
          }
          tc_commit(&s_full[b]);
          if (j >= 1) issue_pv(j - 1);
        }
        issue_pv(HG - 1);
      }
      if (a.phases & 4) {
        // ---- P5: out_g = O . Wo[g]^T into TMEM [192, 512) (S / O accumulators are dead by then; the
        // bf16 attention output of this group occupies [0, 160) as the A operand of its own chunks)
        const bool own_tmem = (a.phases & 2) != 0;
        const uint32_t acc = tmem_base + F_ACC_BASE;
        if (own_tmem) {
          wd_wait(obf_done, 0, W_OBF, abort_s);
          tc_fence_after();
        }
        for (int i = 0; i < KC; ++i, ++it) {
          const bool a_from_tmem = own_tmem && i < 5;
          const int s = it % F_STAGES;
          wd_wait(&full[s], (it / F_STAGES) & 1, W_FULL5, abort_s);
          tc_fence_after();
          const uint32_t st = smem_u32(arena + s * F_STAGE_BYTES);
          const uint64_t ad = umma_desc_sw128(st, 16, 1024);
          const uint64_t b0 = umma_desc_sw128(st + F_A_BYTES, 16, 1024);
          const uint64_t b1 = umma_desc_sw128(st + F_A_BYTES + F_NB * 128, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (a_from_tmem) {
              const uint32_t at = tmem_base + 32 * i + 8 * k;  // 64 channels = 32 packed columns per chunk
              mma_f16_ts(acc, at, b0 + 2 * k, idesc_g, (i | k) != 0);
              mma_f16_ts(acc + F_NB, at, b1 + 2 * k, idesc_g, (i | k) != 0);
            } else {
              mma_f16_ss(acc, ad + 2 * k, b0 + 2 * k, idesc_g, (i | k) != 0);
              mma_f16_ss(acc + F_NB, ad + 2 * k, b1 + 2 * k, idesc_g, (i | k) != 0);
            }
          }
          tc_commit(&empty[s]);
        }
        tc_commit(out_full);
      }
    }
  } else {
    // ================================================================== softmax / epilogue warpgroups
    const int wg = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tq = tmem_base + lane_addr;

    if (a.phases & 2) {
      // ---- P2: fp32 Q -> bf16 pairs, compacted in place.  80 source columns per round, split between
      // the two warpgroups; "all read, barrier, all write" because a round's destination overlaps
      // the other warpgroup's source.  Destinations never reach the next round's source columns.
      wd_wait(q_done, 0, W_QDONE_W, abort_s);
      tc_fence_after();
      const bool stamp = quarter == 0 && lane == 0;
      if (stamp && wg == 0) F_STAMP(2);
#pragma unroll 1
      for (int t = 0; t < F_NQ / 80; ++t) {
        uint32_t v[3][16];
        uint32_t src, dst;
        int n_src;  // 32-bit source columns of this thread in this round
        if (DH == 40) {
          const int h = 2 * t + wg;
          src = 40 * h;
          dst = 24 * h;
          n_src = 40;
        } else {
          src = 80 * t + (wg ? 48 : 0);
          dst = 40 * t + (wg ? 24 : 0);
          n_src = wg ? 32 : 48;
        }
        tmem_ld_x16(tq + src, v[0]);
        tmem_ld_x16(tq + src + 16, v[1]);
        if (DH == 40) {
          uint32_t t8[8];
          tmem_ld_x8(tq + src + 32, t8);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) v[2][i] = t8[i];
#pragma unroll
          for (int i = 8; i < 16; ++i) v[2][i] = 0u;  // d 40 -> 48: zero k-padding
        } else {
          if (n_src == 48) tmem_ld_x16(tq + src + 32, v[2]);
          tmem_ld_wait();
        }
        named_bar_sync(3, 256);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          if (DH != 40 && q == 2 && n_src != 48) break;
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            pk[i] = pack_bf16x2(__uint_as_float(v[q][2 * i]), __uint_as_float(v[q][2 * i + 1]));
          tmem_st_x8(tq + dst + 8 * q, pk);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(qc_done);
      if (stamp && wg == 0) F_STAMP(3);

      // ---- P3: heads wg, wg + 2, ...
      const uint32_t sbuf = tq + F_S_BASE + wg * F_TKV;
      const int ob = Cfg::NOBUF == 2 ? wg : 0;
      const uint32_t obuf = tq + Cfg::O_BASE + ob * Cfg::NPV;
      uint8_t* stage = ost_s + wg * Cfg::OST_BYTES;
      const uint64_t scale2 = f2_pack(a.scale_log2e, a.scale_log2e);
#pragma unroll 1
      for (int j = wg; j < HG; j += 2) {
        wd_wait(&s_full[wg], (j >> 1) & 1, W_SFULL, abort_s);
        tc_fence_after();
        if (stamp) F_STAMP(8 + 4 * j);
        uint64_t acc0 = f2_pack(0.f, 0.f), acc1 = acc0;
        {
          uint32_t v[F_TKV / 16][16];
#pragma unroll
          for (int c = 0; c < F_TKV / 16; ++c) tmem_ld_x16(sbuf + 16 * c, v[c]);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (F_TKV - 16 + i >= a.t_valid) v[F_TKV / 16 - 1][i] = 0xff800000u;  // -inf
          float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
          for (int c = 0; c < F_TKV / 16; ++c)
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              m0 = fmax3(m0, __uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1]));
              m1 = fmax3(m1, __uint_as_float(v[c][i + 2]), __uint_as_float(v[c][i + 3]));
            }
          const float nmx = -fmaxf(m0, m1) * a.scale_log2e;
          const uint64_t nmx2 = f2_pack(nmx, nmx);
#pragma unroll
          for (int c = 0; c < F_TKV / 16; ++c) {
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
        }
        float sa, sb;
        f2_unpack(f2_add(acc0, acc1), sa, sb);
        const float inv = rcp_approx(sa + sb);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[wg]);
        if (stamp) F_STAMP(9 + 4 * j);

        // epilogue of this head: O / rowsum -> bf16 -> staging (SWIZZLE_128B) -> TMA store by warp 0
        wd_wait(&o_full[wg], (j >> 1) & 1, W_OFULL, abort_s);
        tc_fence_after();
        if (stamp) F_STAMP(10 + 4 * j);
        if (a.store_attn && j >= 2) wd_wait(&ost_free[wg], ((j >> 1) - 1) & 1, W_OSTFREE, abort_s);
        const uint64_t inv2 = f2_pack(inv, inv);
        const uint32_t obf = tq + j * (DH / 2);  // this head's bf16 output: over its own (dead) bf16 Q
        constexpr int NCH = Cfg::NPV / 16;
        constexpr int EPI_CH = 4;
#pragma unroll
        for (int c0 = 0; c0 < NCH; c0 += EPI_CH) {
          uint32_t v[EPI_CH][16];
#pragma unroll
          for (int q = 0; q < EPI_CH; ++q)
            if (c0 + q < NCH) tmem_ld_x16(obuf + 16 * (c0 + q), v[q]);
          tmem_ld_wait();
          if (c0 + EPI_CH >= NCH) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[ob]);
          }
#pragma unroll
          for (int q = 0; q < EPI_CH; ++q) {
            const int c = c0 + q;
            if (c < NCH) {
              uint32_t pk[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                float lo, hi;
                f2_unpack(f2_mul(f2_pack(__uint_as_float(v[q][2 * k]), __uint_as_float(v[q][2 * k + 1])), inv2), lo,
                          hi);
                pk[k] = pack_bf16x2(lo, hi);
              }
              // A operand of this group's own to_out chunks (16 c + 16 <= d, or the 8-column tail of d = 40)
              if (16 * c + 16 <= DH) {
                tmem_st_x8(obf + 8 * c, pk);
              } else {
                const uint32_t p4[4] = {pk[0], pk[1], pk[2], pk[3]};
                tmem_st_x4(obf + 8 * c, p4);
              }
              if (a.store_attn) {
                uint8_t* chunk = stage + (c >> 2) * F_Q_CHUNK;
                *reinterpret_cast<uint4*>(chunk + sw128_offset(row, 2 * (c & 3))) =
                    make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(chunk + sw128_offset(row, 2 * (c & 3) + 1)) =
                    make_uint4(pk[4], pk[5], pk[6], pk[7]);
              }
            }
          }
        }
        if (a.store_attn) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ost_full[wg]);
        }
        if (stamp) F_STAMP(11 + 4 * j);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(obf_done);
    }

    if (a.phases & 4) {
      // ---- P6: out accumulator + bias -> bf16 -> [128 x 320] staging at the head of the arena.
      // warpgroup w takes columns [160 w, 160 w + 160)
      wd_wait(out_full, 0, W_OUTFULL, abort_s);
      tc_fence_after();
      if (threadIdx.x == 64) F_STAMP(50);
      // 20 column groups of 16 = 5 staging chunks of 64 columns.  Warpgroup 0 takes groups 0-9, warpgroup 1
      // groups 10-19; every chunk leaves through its own TMA store as soon as its four groups are staged,
      // so the stores read shared memory while the later chunks are still being converted.
      const uint32_t tacc = tq + F_ACC_BASE;
      if (wg == 0) {
        p6_groups<4>(tacc, 0, bias_s, arena, row);
        fence_proxy_async_smem();
        named_bar_sync(4, 128);
        if (threadIdx.x == 64) {
          tma_store_4d(&tm_out, arena, 0, g, row0, sample);
          tma_store_commit();
        }
        p6_groups<4>(tacc, 4, bias_s, arena, row);
        fence_proxy_async_smem();
        named_bar_sync(4, 128);
        if (threadIdx.x == 64) {
          tma_store_4d(&tm_out, arena + F_Q_CHUNK, 64, g, row0, sample);
          tma_store_commit();
        }
        p6_groups<2>(tacc, 8, bias_s, arena, row);
        fence_proxy_async_smem();
        named_bar_sync(2, 256);  // chunk 2 is shared with warpgroup 1
        if (threadIdx.x == 64) {
          tma_store_4d(&tm_out, arena + 2 * F_Q_CHUNK, 128, g, row0, sample);
          tma_store_commit();
          tma_store_wait_read<0>();  // shared memory may go away; the writes complete with the grid
        }
      } else {
        p6_groups<2>(tacc, 10, bias_s, arena, row);
        fence_proxy_async_smem();
        named_bar_sync(2, 256);
        p6_groups<4>(tacc, 12, bias_s, arena, row);
        fence_proxy_async_smem();
        named_bar_sync(5, 128);
        if (threadIdx.x == 192) {
          tma_store_4d(&tm_out, arena + 3 * F_Q_CHUNK, 192, g, row0, sample);
          tma_store_commit();
        }
        p6_groups<4>(tacc, 16, bias_s, arena, row);
        fence_proxy_async_smem();
        named_bar_sync(5, 128);
        if (threadIdx.x == 192) {
          tma_store_4d(&tm_out, arena + 4 * F_Q_CHUNK, 256, g, row0, sample);
          tma_store_commit();
          tma_store_wait_read<0>();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (a.G > 1) cluster_barrier();  // no CTA exits while a peer may still signal its o_ready
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, F_TMEM_COLS);
  }
  if (threadIdx.x == 0) F_STAMP(51);
}

template <int DH>
int launch_k3f(const CUtensorMap* tm, const FArgs& a, int n_tiles, int n_samples, cudaStream_t st) {
  static thread_local int attr_device = -1;
  int dev = 0;
  FD_CUDA_OK(cudaGetDevice(&dev));
  if (attr_device != dev) {
    FD_CUDA_OK(cudaFuncSetAttribute(k3f_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
    attr_device = dev;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(a.G), static_cast<unsigned>(n_tiles), static_cast<unsigned>(n_samples));
  cfg.blockDim = dim3(F_THREADS);
  cfg.dynamicSmemBytes = F_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(a.G);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FD_CUDA_OK(cudaLaunchKernelEx(&cfg, k3f_kernel<DH>, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], tm[7], a));
  return FD_OK;
}

int g_k3f_phases = 7;
int g_k3f_trace_on = 0;
int g_k3f_keep_attn = 1;   // G == 1: also write the attention output when the caller passes a buffer

}  // namespace
}  // namespace fd

// development aids (not part of the product ABI): which phases run (1 = to_q, 2 = attention, 4 = to_out)
// and the watchdog record {first failing wait, block id, parity, 0}
extern "C" void fd_debug_set_k3f_phases(int mask) { fd::g_k3f_phases = mask; }
extern "C" int fd_debug_k3f_trace(long long* out64, int enable) {
  fd::g_k3f_trace_on = enable;
  if (out64 && cudaMemcpyFromSymbol(out64, fd::g_k3f_trace, 64 * sizeof(long long)) != cudaSuccess) return FD_ERR_CUDA;
  return FD_OK;
}
extern "C" int fd_debug_k3f_status(int* out4, int reset) {
  if (cudaMemcpyFromSymbol(out4, fd::g_k3f_status, 4 * sizeof(int)) != cudaSuccess) return FD_ERR_CUDA;
  if (reset) {
    int z[4] = {0, 0, 0, 0};
    if (cudaMemcpyToSymbol(fd::g_k3f_status, z, sizeof(z)) != cudaSuccess) return FD_ERR_CUDA;
  }
  return FD_OK;
}

extern "C" int fd_cross_attn_fused(const void* x_bf16_dev, const void* wq_bf16_dev, const void* kv_bf16_dev,
                                   int64_t kv_rows, int64_t kv_row_stride, int k_col_off, int v_col_off,
                                   const int32_t* ctx_index_dev, const void* wo_bf16_dev, const void* bo_bf16_dev,
                                   int n_samples, int n_q, int heads, int d_head, int t_valid, int t_pad,
                                   float scale, void* attn_bf16_dev, void* out_bf16_dev, void* stream) {
  using namespace fd;
  FD_REQUIRE(x_bf16_dev && wq_bf16_dev && kv_bf16_dev && ctx_index_dev && wo_bf16_dev && bo_bf16_dev && out_bf16_dev,
             "fd_cross_attn_fused: NULL pointer");
  FD_REQUIRE(n_samples > 0 && n_q > 0 && heads > 0, "fd_cross_attn_fused: non-positive shape");
  FD_REQUIRE(d_head == 40 || d_head == 80 || d_head == 160,
             "fd_cross_attn_fused: d_head=%d not in {40, 80, 160}", d_head);
  const int C = heads * d_head;
  FD_REQUIRE(C % F_NQ == 0 && C / F_NQ >= 1 && C / F_NQ <= 8,
             "fd_cross_attn_fused: heads*d_head=%d must be a multiple of %d (at most %d)", C, F_NQ, 8 * F_NQ);
  FD_REQUIRE(t_pad == F_TKV && t_valid > F_TKV - 16 && t_valid <= t_pad,
             "fd_cross_attn_fused: need t_pad == %d and %d < t_valid <= t_pad", F_TKV, F_TKV - 16);
  FD_REQUIRE(scale > 0.f, "fd_cross_attn_fused: scale must be positive");
  FD_REQUIRE(k_col_off % 8 == 0 && v_col_off % 8 == 0 && kv_row_stride % 8 == 0,
             "fd_cross_attn_fused: column offsets / row stride must be multiples of 8 elements");
  FD_REQUIRE(k_col_off + C <= kv_row_stride && v_col_off + C <= kv_row_stride,
             "fd_cross_attn_fused: K/V slice exceeds the cache row");
  FD_REQUIRE(kv_rows % t_pad == 0, "fd_cross_attn_fused: kv_rows=%lld not a multiple of t_pad", (long long)kv_rows);
  FD_REQUIRE(attn_bf16_dev || C == F_NQ,
             "fd_cross_attn_fused: attn_bf16_dev may only be NULL when heads*d_head == %d (the head groups of wider "
             "layers exchange their attention output through it)", F_NQ);
  const void* ptrs[7] = {x_bf16_dev, wq_bf16_dev, kv_bf16_dev, wo_bf16_dev, bo_bf16_dev,
                         attn_bf16_dev ? attn_bf16_dev : out_bf16_dev, out_bf16_dev};
  for (int i = 0; i < 7; ++i)
    FD_REQUIRE(reinterpret_cast<uintptr_t>(ptrs[i]) % 16 == 0, "fd_cross_attn_fused: pointers must be 16-byte aligned");
  FD_REQUIRE(n_samples <= 65535 && (n_q + F_TQ - 1) / F_TQ <= 65535, "fd_cross_attn_fused: grid limits");
  int rc = check_device();
  if (rc != FD_OK) return rc;

  const uint64_t uC = static_cast<uint64_t>(C), uN = static_cast<uint64_t>(n_q), uS = static_cast<uint64_t>(n_samples);
  const int G = C / F_NQ;
  CUtensorMap tm[8];
  // [0] x, [5] attention output as the A operand of to_out: {C, n_q, samples}, box 64 x 128 rows
  for (int which = 0; which < 2; ++which) {
    uint64_t dims[3] = {uC, uN, uS};
    uint64_t strides[2] = {uC * 2, uN * uC * 2};
    uint32_t box[3] = {64, F_TQ, 1};
    rc = encode_tmap(&tm[which == 0 ? 0 : 5], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                     which == 0 ? x_bf16_dev : (attn_bf16_dev ? attn_bf16_dev : out_bf16_dev), dims, strides, box,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
  }
  // [1] Wq, [6] Wo: {k = C, n = C}, box 64 x 160 rows
  for (int which = 0; which < 2; ++which) {
    uint64_t dims[2] = {uC, uC};
    uint64_t strides[1] = {uC * 2};
    uint32_t box[2] = {64, F_NB};
    rc = encode_tmap(&tm[which == 0 ? 1 : 6], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     which == 0 ? wq_bf16_dev : wo_bf16_dev, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
  }
  // [2] K, [3] V: column slices of the K2 cache, {d, heads, rows}, box 64 x 1 x 80 (zero-filled past d)
  for (int which = 0; which < 2; ++which) {
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(kv_bf16_dev) + (which == 0 ? k_col_off : v_col_off);
    uint64_t dims[3] = {static_cast<uint64_t>(d_head), static_cast<uint64_t>(heads), static_cast<uint64_t>(kv_rows)};
    uint64_t strides[2] = {static_cast<uint64_t>(d_head) * 2, static_cast<uint64_t>(kv_row_stride) * 2};
    uint32_t box[3] = {64, 1, F_TKV};
    rc = encode_tmap(&tm[2 + which], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
  }
  // [4] attention output per head: {d, heads, n_q, samples}, box 64 x 1 x 128 x 1 (clipped at d / n_q)
  {
    uint64_t dims[4] = {static_cast<uint64_t>(d_head), static_cast<uint64_t>(heads), uN, uS};
    uint64_t strides[3] = {static_cast<uint64_t>(d_head) * 2, uC * 2, uN * uC * 2};
    uint32_t box[4] = {64, 1, F_TQ, 1};
    rc = encode_tmap(&tm[4], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, attn_bf16_dev ? attn_bf16_dev : out_bf16_dev, dims,
                     strides, box,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
  }
  // [7] layer output per head group: {320, G, n_q, samples}, box 64 x 1 x 128 x 1
  {
    uint64_t dims[4] = {static_cast<uint64_t>(F_NQ), static_cast<uint64_t>(G), uN, uS};
    uint64_t strides[3] = {static_cast<uint64_t>(F_NQ) * 2, uC * 2, uN * uC * 2};
    uint32_t box[4] = {64, 1, F_TQ, 1};
    rc = encode_tmap(&tm[7], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out_bf16_dev, dims, strides, box,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
  }
  FArgs a;
  a.ctx_index = ctx_index_dev;
  a.bias = static_cast<const __nv_bfloat16*>(bo_bf16_dev);
  a.n_q = n_q;
  a.t_valid = t_valid;
  a.t_pad = t_pad;
  a.kchunks = C / 64;
  a.G = G;
  a.phases = g_k3f_phases | 1;
  a.trace = g_k3f_trace_on;
  a.store_attn = (attn_bf16_dev != nullptr && (G > 1 || g_k3f_keep_attn || (a.phases & 6) != 6)) ? 1 : 0;
  a.scale_log2e = scale * 1.4426950408889634f;
  const int n_tiles = (n_q + F_TQ - 1) / F_TQ;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d_head == 40) return launch_k3f<40>(tm, a, n_tiles, n_samples, st);
  if (d_head == 80) return launch_k3f<80>(tm, a, n_tiles, n_samples, st);
  return launch_k3f<160>(tm, a, n_tiles, n_samples, st);
}
