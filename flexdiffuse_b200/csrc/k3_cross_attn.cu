// K3 -- cross-attention of the latent tokens over the CACHED K/V of the fixed 77-token context.
//
// Replaces diffusers' CrossAttention._attention (einsum QK^T * scale -> softmax -> . V, which
// materialises a [2B*8, N, 77] score tensor per layer), reached from pipeline/guide.py:56-58
// for each of the 16 attn2 layers of the SD-v1 UNet.  K and V are column
// slices of the K2 output, so nothing about the context is recomputed inside the loop.
//
// v2 (round 1, after the first ncu pass showed v1 latency-bound at <=20 warps/SM with a serial
// TMA -> MMA -> softmax -> MMA -> store chain per CTA): warp-specialised and software-pipelined.
// A CTA owns one (sample, head) and walks over several 128-query tiles:
//   warp 0        : TMA producer -- K and V once, then the Q tiles through a 2-stage ring
//   warp 1        : tcgen05.mma issuer -- S_i = Q_i K^T is issued while the softmax warps still
//                   work on tile i-1; O_i = P_i V follows as soon as P_i is in TMEM
//   warps 2..5    : one TMEM lane = one query row -> softmax over the 77 keys in registers,
//                   P written back to TMEM as bf16 (A operand of the second MMA); the epilogue
//                   of tile i-1 (O / rowsum -> bf16 -> global) runs after the softmax of tile i
//                   so the P V latency is hidden.
//   S = Q K^T : M=128, N=80, K=d (k-steps of 16)   both operands K-major SWIZZLE_128B in smem
//   O = P V   : M=128, N=d (rounded to 16), K=80   A = P from TMEM, B = V MN-major in smem
// TMEM: S0 [0,80) and S1 [80,160) fp32 (P = bf16x2 in the first 40 columns of its S buffer),
// one O accumulator [160, 160+N).  S is double-buffered so S_{i+1} is computed under the softmax
// of tile i; O needs only one buffer because the epilogue of tile i runs before P V of tile i+1
// is issued.  HBM-bound by design (AI = 77 FLOP/B).
#include "fd_common.cuh"

namespace fd {
namespace {

constexpr int TQ = 128;     // query rows per tile
constexpr int TKV = 80;     // keys padded 77 -> 80
constexpr int K3_THREADS = 192;
constexpr int Q_CHUNK_BYTES = TQ * 128;
constexpr int KV_CHUNK_BYTES = TKV * 128;
constexpr int QSTAGES = 2;

template <int DH>
struct K3Cfg {
  static constexpr int NCHUNK = (DH + 63) / 64;        // 64-element d chunks (128 B swizzle rows)
  static constexpr int KSTEPS = (DH + 15) / 16;        // UMMA k-steps for Q K^T
  static constexpr int NPV = ((DH + 15) / 16) * 16;    // UMMA N for P V
  static constexpr int SMEM_EST = 1024 + 2 * ((DH + 63) / 64) * (TQ * 128) + 2 * ((DH + 63) / 64) * (TKV * 128) + 256;
  static constexpr int S_COLS = TKV;                   // one S / P buffer
  static constexpr int O_BASE = 2 * S_COLS;            // O accumulator after the two S buffers
  static constexpr int TMEM_COLS = O_BASE + NPV <= 256 ? 256 : 512;
  static constexpr int CTAS_PER_SM = (O_BASE + NPV <= 256 && SMEM_EST <= 110 * 1024) ? 2 : 1;
  static constexpr int Q_STAGE_BYTES = NCHUNK * Q_CHUNK_BYTES;
  static constexpr int SMEM = 1024 + QSTAGES * Q_STAGE_BYTES + 2 * NCHUNK * KV_CHUNK_BYTES + 256;
};

// development aid: when set (fd_debug_set_k3_timing), CTA (0,0,0) records %globaltimer (ns) at its
// phase boundaries: [0] start, [1] setup done, [2] K/V+Q0 landed, [3] S0 ready, [4] P0 written,
// [5] O0 ready, [6] epilogue 0 done, [7] kernel end
static long long* g_k3_timing = nullptr;

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define K3_STAMP(slot)                                                                  \
  do {                                                                                  \
    if (a.timing && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) a.timing[slot] = gtime(); \
  } while (0)

struct K3Args {
  long long* timing;
  const int32_t* ctx_index;
  __nv_bfloat16* out;
  int n_q, heads, t_valid, t_pad, n_tiles;
  float scale_log2e;
};

template <int DH>
__global__ void __launch_bounds__(K3_THREADS)
k3_cross_attn_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const K3Args a) {
  using Cfg = K3Cfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sq = smem;                                   // QSTAGES x [NCHUNK x 128 x 128 B]
  uint8_t* sk = sq + QSTAGES * Cfg::Q_STAGE_BYTES;      // NCHUNK x 80 x 128 B
  uint8_t* sv = sk + Cfg::NCHUNK * KV_CHUNK_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sv + Cfg::NCHUNK * KV_CHUNK_BYTES);
  uint64_t* kv_full = bars;          // K + V landed
  uint64_t* q_full = bars + 1;       // [2] Q stage landed
  uint64_t* q_empty = bars + 3;      // [2] Q stage consumed by the MMA
  uint64_t* s_full = bars + 5;       // [2] S = Q K^T complete
  uint64_t* p_full = bars + 7;       // [2] P written by the 4 softmax warps
  uint64_t* o_full = bars + 9;       // O = P V complete
  uint64_t* o_free = bars + 10;      // epilogue has drained O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int sample = blockIdx.z;
  if (threadIdx.x == 0) K3_STAMP(0);
  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int my_tiles = (a.n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                       static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);
    }
    mbar_init(o_full, 1);
    mbar_init(o_free, 4);
    fence_mbar_init();
  }
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
    if (elect_one()) {
      const int ctx_row = a.ctx_index[sample] * a.t_pad;
      mbar_expect_tx(kv_full, 2 * Cfg::NCHUNK * KV_CHUNK_BYTES);
#pragma unroll
      for (int c = 0; c < Cfg::NCHUNK; ++c) {
        tma_load_3d(sk + c * KV_CHUNK_BYTES, &tm_k, kv_full, c * 64, head, ctx_row);
        tma_load_3d(sv + c * KV_CHUNK_BYTES, &tm_v, kv_full, c * 64, head, ctx_row);
      }
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        mbar_wait_backoff(&q_empty[s], ph ^ 1);
        mbar_expect_tx(&q_full[s], Cfg::Q_STAGE_BYTES);
        const int q0 = (static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x)) * TQ;
#pragma unroll
        for (int c = 0; c < Cfg::NCHUNK; ++c)
          tma_load_4d(sq + s * Cfg::Q_STAGE_BYTES + c * Q_CHUNK_BYTES, &tm_q, &q_full[s], c * 64, head, q0,
                      sample);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc(UMMA_BF16, TQ, TKV, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc(UMMA_BF16, TQ, Cfg::NPV, 0, 1);  // B (= V) MN-major
      auto issue_pv = [&](int j) {
        const int b = j & 1;
        mbar_wait_backoff(&p_full[b], (j >> 1) & 1);
        if (j >= 1) mbar_wait_backoff(o_free, (j - 1) & 1);  // epilogue of tile j-1 drained O
        tc_fence_after();
        const uint32_t pbuf = tmem_base + b * Cfg::S_COLS;
#pragma unroll
        for (int k = 0; k < TKV / 16; ++k) {
          // 16 keys = two 8-row groups (SBO = 1024 B); next 64-wide d chunk LBO = KV_CHUNK_BYTES away
          const uint64_t vd = umma_desc_sw128(smem_u32(sv + k * 2048), KV_CHUNK_BYTES, 1024);
          mma_f16_ts(tmem_base + Cfg::O_BASE, pbuf + 8 * k, vd, idesc_o, k != 0);
        }
        tc_commit(o_full);
      };
      mbar_wait_backoff(kv_full, 0);
      for (int i = 0; i < my_tiles; ++i) {
        const int b = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        mbar_wait_backoff(&q_full[b], ph);
        if (i == 0) K3_STAMP(2);
        // S buffer b was last read by P V of tile i-2, issued earlier by this thread: the tensor
        // pipe executes in issue order, so no extra wait is needed before overwriting it
        tc_fence_after();
        const uint32_t buf = tmem_base + b * Cfg::S_COLS;
        const uint8_t* qs = sq + b * Cfg::Q_STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
          const int c = ks >> 2, kk = ks & 3;
          const uint64_t qd = umma_desc_sw128(smem_u32(qs + c * Q_CHUNK_BYTES), 16, 1024) + 2 * kk;
          const uint64_t kd = umma_desc_sw128(smem_u32(sk + c * KV_CHUNK_BYTES), 16, 1024) + 2 * kk;
          mma_f16_ss(buf, qd, kd, idesc_s, ks != 0);
        }
        tc_commit(&s_full[b]);
        tc_commit(&q_empty[b]);
        if (i >= 1) issue_pv(i - 1);
      }
      if (my_tiles >= 1) issue_pv(my_tiles - 1);
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue warps
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32)
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    float sum_prev = 1.f;

    auto epilogue = [&](int j, float sum) {
      mbar_wait(o_full, j & 1);
      tc_fence_after();
      if (j == 0 && threadIdx.x == 64) K3_STAMP(5);
      const uint32_t buf = tmem_base + lane_addr + Cfg::O_BASE;
      const int q = (static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x)) * TQ + row;
      const float inv = 1.0f / sum;
      __nv_bfloat16* dst =
          a.out + (static_cast<size_t>(sample) * a.n_q + q) * (static_cast<size_t>(a.heads) * DH) + head * DH;
      constexpr int GROUP = 3;  // 16-column loads in flight per wait (48 registers)
#pragma unroll
      for (int c0 = 0; c0 < Cfg::NPV; c0 += 16 * GROUP) {
        uint32_t v[GROUP][16];
#pragma unroll
        for (int g = 0; g < GROUP; ++g)
          if (c0 + 16 * g < Cfg::NPV) tmem_ld_x16(buf + c0 + 16 * g, v[g]);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < GROUP; ++g) {
          const int c = c0 + 16 * g;
          if (c < Cfg::NPV && q < a.n_q) {
            uint32_t pk[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              __nv_bfloat162 o = __floats2bfloat162_rn(__uint_as_float(v[g][2 * k]) * inv,
                                                       __uint_as_float(v[g][2 * k + 1]) * inv);
              pk[k] = *reinterpret_cast<uint32_t*>(&o);
            }
            if (c + 8 <= DH) *reinterpret_cast<uint4*>(dst + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            if (c + 16 <= DH) *reinterpret_cast<uint4*>(dst + c + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      if (j == 0 && threadIdx.x == 64) K3_STAMP(6);
    };

    for (int i = 0; i < my_tiles; ++i) {
      const int b = i & 1;
      mbar_wait(&s_full[b], (i >> 1) & 1);
      tc_fence_after();
      if (i == 0 && threadIdx.x == 64) K3_STAMP(3);
      const uint32_t buf = tmem_base + b * Cfg::S_COLS + lane_addr;
      float p[TKV];
      {
        // all five 16-column loads in flight, one wait
        uint32_t v[TKV / 16][16];
#pragma unroll
        for (int c = 0; c < TKV / 16; ++c) tmem_ld_x16(buf + 16 * c, v[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < TKV / 16; ++c)
#pragma unroll
          for (int q = 0; q < 16; ++q) p[16 * c + q] = __uint_as_float(v[c][q]);
      }
      // keys >= t_valid only exist in the last 16 columns (t_valid > 64 is checked on the host)
#pragma unroll
      for (int j = TKV - 16; j < TKV; ++j)
        if (j >= a.t_valid) p[j] = -INFINITY;
      float m0 = p[0], m1 = p[1];
#pragma unroll
      for (int j = 2; j < TKV; j += 2) {
        m0 = fmaxf(m0, p[j]);
        m1 = fmaxf(m1, p[j + 1]);
      }
      // exp2(s * scale*log2e - max * scale*log2e): one FFMA + one MUFU per key (scale > 0)
      const float nmx = -fmaxf(m0, m1) * a.scale_log2e;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int j = 0; j < TKV; j += 4) {
        p[j] = ex2_approx(fmaf(p[j], a.scale_log2e, nmx));
        p[j + 1] = ex2_approx(fmaf(p[j + 1], a.scale_log2e, nmx));
        p[j + 2] = ex2_approx(fmaf(p[j + 2], a.scale_log2e, nmx));
        p[j + 3] = ex2_approx(fmaf(p[j + 3], a.scale_log2e, nmx));
        s0 += p[j];
        s1 += p[j + 1];
        s2 += p[j + 2];
        s3 += p[j + 3];
      }
      const float sum = (s0 + s1) + (s2 + s3);
#pragma unroll
      for (int c = 0; c < TKV / 2; c += 8) {
        uint32_t pk[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          __nv_bfloat162 o = __floats2bfloat162_rn(p[2 * (c + q)], p[2 * (c + q) + 1]);
          pk[q] = *reinterpret_cast<uint32_t*>(&o);
        }
        tmem_st_x8(buf + c, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[b]);
      if (i == 0 && threadIdx.x == 64) K3_STAMP(4);
      // epilogue of the previous tile while this tile's P V runs
      if (i >= 1) epilogue(i - 1, sum_prev);
      sum_prev = sum;
    }
    if (my_tiles >= 1) epilogue(my_tiles - 1, sum_prev);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
  if (threadIdx.x == 0) K3_STAMP(7);
}

template <int DH>
int launch_k3(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const K3Args& a, dim3 grid,
              cudaStream_t st) {
  using Cfg = K3Cfg<DH>;
  static thread_local int attr_device = -1;  // the attribute is per device; set it once per thread/device
  int dev = 0;
  FD_CUDA_OK(cudaGetDevice(&dev));
  if (attr_device != dev) {
    FD_CUDA_OK(cudaFuncSetAttribute(k3_cross_attn_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_device = dev;
  }
  k3_cross_attn_kernel<DH><<<grid, K3_THREADS, Cfg::SMEM, st>>>(tq, tk, tv, a);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}

}  // namespace
}  // namespace fd

// development aid (not part of the product ABI): device buffer of >= 8 int64 for phase timestamps
extern "C" void fd_debug_set_k3_timing(void* buf_dev) { fd::g_k3_timing = static_cast<long long*>(buf_dev); }

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
  CUtensorMap tq, tk, tv;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(d_head), static_cast<uint64_t>(heads), static_cast<uint64_t>(n_q),
                        static_cast<uint64_t>(n_samples)};
    uint64_t strides[3] = {static_cast<uint64_t>(d_head) * 2, C * 2, static_cast<uint64_t>(n_q) * C * 2};
    uint32_t box[4] = {64, 1, TQ, 1};
    rc = encode_tmap(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, q_bf16_dev, dims, strides, box,
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
  if (d_head == 40) return launch_k3<40>(tq, tk, tv, a, grid, st);
  if (d_head == 80) return launch_k3<80>(tq, tk, tv, a, grid, st);
  return launch_k3<160>(tq, tk, tv, a, grid, st);
}
