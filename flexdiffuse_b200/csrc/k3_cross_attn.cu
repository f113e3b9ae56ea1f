// K3 -- cross-attention of the latent tokens over the CACHED K/V of the fixed 77-token context.
//
// Replaces diffusers' CrossAttention._attention (einsum QK^T * scale -> softmax -> . V, which
// materialises a [2B*8, N, 77] score tensor per layer), reached from pipeline/guide.py:56-58
// for each of the 16 attn2 layers of the SD-v1 UNet.  K and V are column
// slices of the K2 output, so nothing about the context is recomputed inside the loop.
//
// One CTA = one (sample, head, 128-query tile).  4 warps:
//   thread 0       : TMA loads (Q tile, K, V; SWIZZLE_128B, zero-filled to 64-wide d chunks and
//                    to 80 keys) and the two tcgen05.mma chains
//   all 128 threads: one TMEM lane = one query row -> softmax over the 77 keys entirely in
//                    registers, P written back to TMEM as bf16 (A operand of the second MMA),
//                    epilogue O / rowsum -> bf16 -> global.
//   S = Q K^T : M=128, N=80, K=d (k-steps of 16)   both operands K-major in smem
//   O = P V   : M=128, N=d (rounded to 16), K=80   A = P from TMEM, B = V MN-major in smem
// TMEM columns: S [0,80) fp32, P [0,40) bf16x2 aliasing S, O [40, 40+N) (S is dead by then).
// HBM-bound by design (AI = 77 FLOP/B): the point is to read Q once, write O once.
#include "fd_common.cuh"

namespace fd {
namespace {

constexpr int TQ = 128;     // query rows per CTA
constexpr int TKV = 80;     // keys padded 77 -> 80
constexpr int K3_THREADS = 128;
constexpr int Q_CHUNK_BYTES = TQ * 128;
constexpr int KV_CHUNK_BYTES = TKV * 128;
constexpr int O_COL = 40;

template <int DH>
struct K3Cfg {
  static constexpr int NCHUNK = (DH + 63) / 64;        // 64-element d chunks (128 B swizzle rows)
  static constexpr int KSTEPS = (DH + 15) / 16;        // UMMA k-steps for Q K^T
  static constexpr int NPV = ((DH + 15) / 16) * 16;    // UMMA N for P V
  static constexpr int TMEM_COLS = (O_COL + NPV) <= 128 ? 128 : 256;
  static constexpr int SMEM = 1024 + NCHUNK * (Q_CHUNK_BYTES + 2 * KV_CHUNK_BYTES) + 64;
};

struct K3Args {
  const int32_t* ctx_index;
  __nv_bfloat16* out;
  int n_q, heads, t_valid, t_pad;
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
  uint8_t* sq = smem;
  uint8_t* sk = sq + Cfg::NCHUNK * Q_CHUNK_BYTES;
  uint8_t* sv = sk + Cfg::NCHUNK * KV_CHUNK_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sv + Cfg::NCHUNK * KV_CHUNK_BYTES);
  uint64_t* qk_bar = bars;      // Q + K landed
  uint64_t* v_bar = bars + 1;   // V landed
  uint64_t* s_bar = bars + 2;   // S = Q K^T complete
  uint64_t* o_bar = bars + 3;   // O = P V complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int q0 = blockIdx.x * TQ;
  const int head = blockIdx.y;
  const int sample = blockIdx.z;

  if (tid == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    mbar_init(qk_bar, 1);
    mbar_init(v_bar, 1);
    mbar_init(s_bar, 1);
    mbar_init(o_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (tid == 0) {
    const int ctx_row = a.ctx_index[sample] * a.t_pad;
    mbar_expect_tx(qk_bar, Cfg::NCHUNK * (Q_CHUNK_BYTES + KV_CHUNK_BYTES));
#pragma unroll
    for (int c = 0; c < Cfg::NCHUNK; ++c) {
      tma_load_4d(sq + c * Q_CHUNK_BYTES, &tm_q, qk_bar, c * 64, head, q0, sample);
      tma_load_3d(sk + c * KV_CHUNK_BYTES, &tm_k, qk_bar, c * 64, head, ctx_row);
    }
    mbar_expect_tx(v_bar, Cfg::NCHUNK * KV_CHUNK_BYTES);
#pragma unroll
    for (int c = 0; c < Cfg::NCHUNK; ++c)
      tma_load_3d(sv + c * KV_CHUNK_BYTES, &tm_v, v_bar, c * 64, head, ctx_row);

    // S = Q K^T
    mbar_wait(qk_bar, 0);
    tc_fence_after();
    constexpr uint32_t idesc_s = umma_idesc(UMMA_BF16, TQ, TKV, 0, 0);
#pragma unroll
    for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
      const int c = ks >> 2, kk = ks & 3;
      const uint64_t qd = umma_desc_sw128(smem_u32(sq + c * Q_CHUNK_BYTES), 16, 1024) + 2 * kk;
      const uint64_t kd = umma_desc_sw128(smem_u32(sk + c * KV_CHUNK_BYTES), 16, 1024) + 2 * kk;
      mma_f16_ss(tmem_base, qd, kd, idesc_s, ks != 0);
    }
    tc_commit(s_bar);
  }
  __syncwarp();

  // ---- softmax over the keys: one TMEM lane per thread
  mbar_wait(s_bar, 0);
  tc_fence_after();
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  float p[TKV];
#pragma unroll
  for (int c = 0; c < TKV; c += 16) {
    uint32_t v[16];
    tmem_ld_x16(tmem_base + lane_addr + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 16; ++q) p[c + q] = __uint_as_float(v[q]);
  }
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < TKV; ++j) {
    p[j] = (j < a.t_valid) ? p[j] * a.scale_log2e : -INFINITY;
    mx = fmaxf(mx, p[j]);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < TKV; ++j) {
    p[j] = exp2f(p[j] - mx);  // masked keys: exp2(-inf) = 0
    sum += p[j];
  }
#pragma unroll
  for (int c = 0; c < TKV / 2; c += 8) {
    uint32_t pk[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      __nv_bfloat162 b = __floats2bfloat162_rn(p[2 * (c + q)], p[2 * (c + q) + 1]);
      pk[q] = *reinterpret_cast<uint32_t*>(&b);
    }
    tmem_st_x8(tmem_base + lane_addr + c, pk);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();

  // ---- O = P V
  if (tid == 0) {
    tc_fence_after();
    mbar_wait(v_bar, 0);
    tc_fence_after();
    constexpr uint32_t idesc_o = umma_idesc(UMMA_BF16, TQ, Cfg::NPV, 0, 1);  // B (= V) is MN-major
#pragma unroll
    for (int j = 0; j < TKV / 16; ++j) {
      // 16 keys = two 8-row groups (SBO = 1024 B); next 64-wide d chunk LBO = KV_CHUNK_BYTES away
      const uint64_t vd = umma_desc_sw128(smem_u32(sv + j * 2048), KV_CHUNK_BYTES, 1024);
      mma_f16_ts(tmem_base + O_COL, tmem_base + 8 * j, vd, idesc_o, j != 0);
    }
    tc_commit(o_bar);
  }
  __syncwarp();

  mbar_wait(o_bar, 0);
  tc_fence_after();
  {
    const int row = warp * 32 + lane;
    const int q = q0 + row;
    const float inv = 1.0f / sum;
    __nv_bfloat16* dst =
        a.out + (static_cast<size_t>(sample) * a.n_q + q) * (static_cast<size_t>(a.heads) * DH) + head * DH;
#pragma unroll
    for (int c = 0; c < Cfg::NPV; c += 16) {
      uint32_t v[16];
      tmem_ld_x16(tmem_base + lane_addr + O_COL + c, v);
      tmem_ld_wait();
      if (q < a.n_q) {
        uint32_t pk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          __nv_bfloat162 b =
              __floats2bfloat162_rn(__uint_as_float(v[2 * k]) * inv, __uint_as_float(v[2 * k + 1]) * inv);
          pk[k] = *reinterpret_cast<uint32_t*>(&b);
        }
        if (c + 8 <= DH) *reinterpret_cast<uint4*>(dst + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        if (c + 16 <= DH) *reinterpret_cast<uint4*>(dst + c + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int DH>
int launch_k3(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const K3Args& a, dim3 grid,
              cudaStream_t st) {
  using Cfg = K3Cfg<DH>;
  FD_CUDA_OK(cudaFuncSetAttribute(k3_cross_attn_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  k3_cross_attn_kernel<DH><<<grid, K3_THREADS, Cfg::SMEM, st>>>(tq, tk, tv, a);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}

}  // namespace
}  // namespace fd

extern "C" int fd_cross_attn(const void* q_bf16_dev, const void* kv_bf16_dev, int64_t kv_rows, int64_t kv_row_stride,
                             int k_col_off, int v_col_off, const int32_t* ctx_index_dev, int n_samples, int n_q,
                             int heads, int d_head, int t_valid, int t_pad, float scale, void* out_bf16_dev,
                             void* stream) {
  using namespace fd;
  FD_REQUIRE(q_bf16_dev && kv_bf16_dev && ctx_index_dev && out_bf16_dev, "fd_cross_attn: NULL pointer");
  FD_REQUIRE(n_samples > 0 && n_q > 0 && heads > 0, "fd_cross_attn: non-positive shape");
  FD_REQUIRE(d_head == 40 || d_head == 80 || d_head == 160, "fd_cross_attn: d_head=%d not in {40, 80, 160}", d_head);
  FD_REQUIRE(t_pad == TKV && t_valid >= 1 && t_valid <= t_pad, "fd_cross_attn: need t_pad == %d and 1 <= t_valid <= t_pad",
             TKV);
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
  K3Args a;
  a.ctx_index = ctx_index_dev;
  a.out = static_cast<__nv_bfloat16*>(out_bf16_dev);
  a.n_q = n_q;
  a.heads = heads;
  a.t_valid = t_valid;
  a.t_pad = t_pad;
  a.scale_log2e = scale * 1.4426950408889634f;
  dim3 grid((n_q + TQ - 1) / TQ, heads, n_samples);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d_head == 40) return launch_k3<40>(tq, tk, tv, a, grid, st);
  if (d_head == 80) return launch_k3<80>(tq, tk, tv, a, grid, st);
  return launch_k3<160>(tq, tk, tv, a, grid, st);
}
