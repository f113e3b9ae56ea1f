// K11 -- fp32-accurate Linear on tcgen05 for the CLIP towers:
//   out[M, N] = act( x[M, K] . W[N, K]^T + bias[N] )
// Replaces the fp32 `nn.Linear`s (q / k / v / out_proj, fc1, fc2) of the ViT-L/14 vision tower and the
// text tower that `CLIPEncoder.image` / `.prompt` run (reference encode/clip.py:57-65, 86-100 through
// transformers' CLIPModel).  The reference computes them in fp32 and K1's decisions hang on near-ties
// of the resulting embeddings, so a single bf16 / tf32 pass is not acceptable (3e-4 relative error on the
// tower output, measured); in fp32 they are CUDA-core GEMMs: 155 GFLOP per image, 7.6 ms on a B200.
//
// Same arithmetic as K1 / K1P: every operand row is scaled by a power of two and split into two fp16
// terms, x 2^e = h1 + h2 (11 + 11 mantissa bits); the three exact products h2.w1 + h1.w2 + h1.w1 are
// accumulated in fp32 in TMEM, which carries the result to fp32 accuracy (error <= ~4e-7 |x| |w| per row
// pair; tests/test_linear_x3_gpu.py).  Unlike K1P the scale is chosen PER ROW from the row's maximum
// (max |x| 2^e in [2^13, 2^14)), so no activation range has to be assumed inside the towers.
//   k11_split_kernel<V, LN> fp32 rows -> h1 / h2 planes + the row's inverse scale (weights once per version, activations
//                           once per call; one 128-thread CTA per row, the row in registers); LN = true: the pre-LN
//                           LayerNorm of the towers' blocks fused in front of the split
//   k11_gemm_kernel<BN>     both operands by TMA (SWIZZLE_128B, K chunks of 64, multi-stage ring),
//                           warps 0 / 2 producers, warp 1 MMA issuer (per k-step h1 x [w1; w2] as one M128 N=2BN MMA, then h2 x w1),
//                           warps 2-5 epilogue: TMEM -> x inv_a[row] x inv_w[col] (+ bias, activation, residual) ->
//                           fp32 global; with split-K the last CTA of a tile to finish adds the slices in ascending order
#include <cuda_fp16.h>

#include "fd_common.cuh"

namespace fd {
namespace {

constexpr int L_THREADS = 192;
constexpr int L_KC = 64;
constexpr int L_A_PLANE = 128 * 128;   // 128 rows x 128 B

__device__ int g_k11_flag;  // 1: non-finite operand; >= 16: a bounded wait gave up (development aid)

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == FD_LINEAR_ACT_QUICK_GELU) return v / (1.0f + __expf(-1.702f * v));
  if (act == FD_LINEAR_ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
  return v;
}

// Operand preparation: ONE CTA of 128 threads per row (a warp per row left 33 CTAs for the towers' 257 rows and
// made K = 4096 a 14 us launch).  The row lives in registers (V float4 per thread, K = 512 V), the row maximum picks the
// power of two, then the two-term split.  LN: y = (x - mean) rstd gamma + beta first, never written to memory.
constexpr int SP_THREADS = 128;

__device__ __forceinline__ float block_reduce_128(float v, bool is_max, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, u) : v + u;
  }
  const int w = threadIdx.x >> 5;
  __syncthreads();   // red[] may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  // fixed order over the 4 warps: bit-reproducible
  return is_max ? fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])) : (red[0] + red[1]) + (red[2] + red[3]);
}

template <int V, bool LN>
__global__ void __launch_bounds__(SP_THREADS) k11_split_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float eps,
                                                               __half* __restrict__ h1p, __half* __restrict__ h2p,
                                                               float* __restrict__ inv_scale, int K) {
  __shared__ float red[4];
  const int row = blockIdx.x, tid = threadIdx.x;
  const int k4 = K >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * K);
  float4 v[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = tid + SP_THREADS * i;
    v[i] = c < k4 ? __ldg(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  bool bad = false;
  if (LN) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = block_reduce_128(sum, false, red) / K;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
      if (tid + SP_THREADS * i < k4) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
      }
    sq = block_reduce_128(sq, false, red);
    const float rstd = rsqrtf(sq / K + eps);
    bad = !(sq <= 3.0e38f);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = tid + SP_THREADS * i;
      if (c < k4) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c);
        v[i].x = (v[i].x - mean) * rstd * g.x + bt.x;
        v[i].y = (v[i].y - mean) * rstd * g.y + bt.y;
        v[i].z = (v[i].z - mean) * rstd * g.z + bt.z;
        v[i].w = (v[i].w - mean) * rstd * g.w + bt.w;
      }
    }
  }
  float mx = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
    bad |= !(fabsf(v[i].x) <= 3.0e38f) | !(fabsf(v[i].y) <= 3.0e38f) | !(fabsf(v[i].z) <= 3.0e38f) | !(fabsf(v[i].w) <= 3.0e38f);
  }
  mx = block_reduce_128(mx, true, red);
  int ex = 0;
  if (mx > 0.f) frexpf(mx, &ex);            // mx = m 2^ex, m in [0.5, 1)
  ex = max(-100, min(100, ex));
  const float scale = exp2f(static_cast<float>(14 - ex));   // mx * scale in [2^13, 2^14)
  if (bad) g_k11_flag = 1;
  if (tid == 0) inv_scale[row] = exp2f(static_cast<float>(ex - 14));
  uint2* o1 = reinterpret_cast<uint2*>(h1p + static_cast<size_t>(row) * K);
  uint2* o2 = reinterpret_cast<uint2*>(h2p + static_cast<size_t>(row) * K);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = tid + SP_THREADS * i;
    if (c < k4) {
      const float a = v[i].x * scale, b = v[i].y * scale, cc = v[i].z * scale, d = v[i].w * scale;
      const __half2 p0 = __floats2half2_rn(a, b), p1 = __floats2half2_rn(cc, d);
      const float2 f0 = __half22float2(p0), f1 = __half22float2(p1);
      const __half2 q0 = __floats2half2_rn(a - f0.x, b - f0.y), q1 = __floats2half2_rn(cc - f1.x, d - f1.y);
      o1[c] = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
      o2[c] = make_uint2(*reinterpret_cast<const uint32_t*>(&q0), *reinterpret_cast<const uint32_t*>(&q1));
    }
  }
}

template <bool LN>
int launch_split(const float* x, const float* gamma, const float* beta, float eps, __half* h1, __half* h2, float* inv,
                 int rows, int K, cudaStream_t st) {
  const int v = (K / 4 + SP_THREADS - 1) / SP_THREADS;   // float4 per thread
  switch (v) {
#define FD_SP_CASE(V) case V: k11_split_kernel<V, LN><<<rows, SP_THREADS, 0, st>>>(x, gamma, beta, eps, h1, h2, inv, K); break;
    FD_SP_CASE(1) FD_SP_CASE(2) FD_SP_CASE(3) FD_SP_CASE(4) FD_SP_CASE(5) FD_SP_CASE(6) FD_SP_CASE(7) FD_SP_CASE(8)
    FD_SP_CASE(9) FD_SP_CASE(10) FD_SP_CASE(11) FD_SP_CASE(12) FD_SP_CASE(13) FD_SP_CASE(14) FD_SP_CASE(15) FD_SP_CASE(16)
#undef FD_SP_CASE
    default: return set_error(FD_ERR_ARG, "fd_linear_x3_split: K=%d exceeds 8192", K);
  }
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}

struct LArgs {
  const float* inv_a;   // [M]
  const float* inv_w;   // [N]
  const float* bias;      // [N] or nullptr
  const float* residual;  // [M, N] or nullptr: out = residual + act(...)
  float* out;             // [M, N]
  int M, N, K, act, split_k, kc_per_split;
};

template <int BN>
__global__ void __launch_bounds__(L_THREADS, 1)
k11_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const LArgs a) {
  constexpr int W_PLANE = BN * 128;
  constexpr int STAGE = 2 * L_A_PLANE + 2 * W_PLANE;
  constexpr int STAGES = (200 * 1024) / STAGE;   // BN 128: 3 stages of 64 KB, BN 64: 4 of 48 KB
  extern __shared__ uint8_t smem_raw[];
  uint8_t* stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(stage + STAGES * STAGE);
  uint64_t* empty = full + STAGES;
  uint64_t* done = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * BN, z = blockIdx.z;
  const int kc0 = z * a.kc_per_split;
  const int num_kc = min(a.kc_per_split, a.K / L_KC - kc0);
  if (tid == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 2);   // one arrival (+ its bytes) from each of the two producer threads
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Two producer threads, ONE box each per K chunk: a thread issues at most one TMA box per ~340 cycles whatever its size
  // (profiles/microbench/tma_rate_all.cu), so both planes of an operand travel as one 3-D box {64, rows, 2 planes} and the
  // activation / weight boxes come from different warps (warp 0 and, while it has nothing else to do, epilogue warp 2).
  if (warp == 0 || warp == 2) {
    if (elect_one()) {
      const bool is_w = warp == 2;
      for (int i = 0; i < num_kc; ++i) {
        const int s = i % STAGES;
        mbar_wait_bounded(&empty[s], ((i / STAGES) & 1) ^ 1, &g_k11_flag, 16);
        uint8_t* st = stage + s * STAGE;
        const int k = (kc0 + i) * L_KC;
        if (is_w) {
          mbar_expect_tx(&full[s], 2 * W_PLANE);
          tma_load_3d(st + 2 * L_A_PLANE, &tm_w, &full[s], k, n0, 0);
        } else {
          mbar_expect_tx(&full[s], 2 * L_A_PLANE);
          tma_load_3d(st, &tm_a, &full[s], k, m0, 0);
        }
      }
    }
    __syncwarp();
  }
  if (warp == 0) {
  } else if (warp == 1) {
    if (elect_one()) {
      // two MMAs per k-step instead of three: h1 against BOTH weight planes at once (their tiles are adjacent in shared
      // memory: one B operand of 2 BN rows -> accumulators D1 | D2 side by side), then h2 against w1 into D1.  The A tile is
      // read twice instead of three times -- the kernel is bound by shared-memory bandwidth (SS-mode MMAs at N <= 128 read
      // as many bytes as the tensor pipe can take) and by the single-thread issue rate at N = 64.
      constexpr uint32_t idesc2 = umma_idesc(UMMA_F16, 128, 2 * BN, 0, 0);
      constexpr uint32_t idesc1 = umma_idesc(UMMA_F16, 128, BN, 0, 0);
      for (int i = 0; i < num_kc; ++i) {
        const int s = i % STAGES;
        mbar_wait_bounded(&full[s], (i / STAGES) & 1, &g_k11_flag, 17);
        tc_fence_after();
        const uint32_t base = smem_u32(stage + s * STAGE);
        const uint64_t a1 = umma_desc_sw128(base, 16, 1024), a2 = umma_desc_sw128(base + L_A_PLANE, 16, 1024);
        const uint64_t w12 = umma_desc_sw128(base + 2 * L_A_PLANE, 16, 1024);   // w1 rows, then w2 rows
#pragma unroll
        for (int ks = 0; ks < L_KC / 16; ++ks) {
          mma_f16_ss(tmem_base, a1 + 2 * ks, w12 + 2 * ks, idesc2, (i | ks) != 0);   // D1 += h1.w1, D2 += h1.w2
          mma_f16_ss(tmem_base, a2 + 2 * ks, w12 + 2 * ks, idesc1, 1);               // D1 += h2.w1
        }
        tc_commit(&empty[s]);
      }
      tc_commit(done);
    }
  } else {
    // epilogue: warp w owns TMEM lanes 32 (w % 4) .. + 31 = rows m0 + 32 (w % 4) + lane
    const int quarter = warp & 3;
    const int row = m0 + quarter * 32 + lane;
    const bool row_ok = row < a.M;
    const float ia = row_ok ? a.inv_a[row] : 0.f;
    mbar_wait_bounded(done, 0, &g_k11_flag, 18);
    tc_fence_after();
    const size_t rofs = static_cast<size_t>(row_ok ? row : 0) * a.N;
    auto finish = [&](float4 o, int n) {   // + bias, activation, + residual -> out
      if (a.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + n));
        o.x += b.x;
        o.y += b.y;
        o.z += b.z;
        o.w += b.w;
      }
      o.x = act_apply(o.x, a.act);
      o.y = act_apply(o.y, a.act);
      o.z = act_apply(o.z, a.act);
      o.w = act_apply(o.w, a.act);
      if (a.residual) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(a.residual + rofs + n));
        o.x += r.x;
        o.y += r.y;
        o.z += r.z;
        o.w += r.w;
      }
      *reinterpret_cast<float4*>(a.out + rofs + n) = o;
    };
    // split-K: the K slices of one output tile are the CTAs of a cluster (rank = slice).  Each parks its scaled fp32
    // tile in its own shared memory (the operand ring is dead by now); after a cluster barrier rank r adds rows
    // [128 r / split_k, 128 (r + 1) / split_k) of all slices in ascending slice order over DSMEM and finishes them
    // with coalesced stores.  No partials in global memory, no second launch, bit-reproducible.
    constexpr int PITCH = BN + 4;   // floats per parked row: row-per-lane 16-byte writes without extra bank conflicts
    float* park = reinterpret_cast<float*>(stage);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[2][16], u[2][16];   // D1 (h1.w1 + h2.w1) and D2 (h1.w2) of the same 32 columns
      const uint32_t tl = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c0;
      tmem_ld_x16(tl, v[0]);
      tmem_ld_x16(tl + 16, v[1]);
      tmem_ld_x16(tl + BN, u[0]);
      tmem_ld_x16(tl + BN + 16, u[1]);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int q = 0; q < 16; q += 4) {
          const int cl = c0 + 16 * g + q, n = n0 + cl;
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && n < a.N) {  // N % 4 == 0
            const float4 iw = __ldg(reinterpret_cast<const float4*>(a.inv_w + n));
            o = make_float4((__uint_as_float(v[g][q]) + __uint_as_float(u[g][q])) * (ia * iw.x),
                            (__uint_as_float(v[g][q + 1]) + __uint_as_float(u[g][q + 1])) * (ia * iw.y),
                            (__uint_as_float(v[g][q + 2]) + __uint_as_float(u[g][q + 2])) * (ia * iw.z),
                            (__uint_as_float(v[g][q + 3]) + __uint_as_float(u[g][q + 3])) * (ia * iw.w));
            if (a.split_k == 1) finish(o, n);
          }
          if (a.split_k > 1) *reinterpret_cast<float4*>(park + (quarter * 32 + lane) * PITCH + cl) = o;
        }
    }
  }
  if (a.split_k > 1) {
    // (all 192 threads of every CTA of the cluster take part in the barriers)
    cluster_barrier();   // every slice is parked (release / acquire at cluster scope)
    if (warp >= 2) {
      constexpr int PITCH = BN + 4;
      float* park = reinterpret_cast<float*>(stage);
      const int rows_per = 128 / a.split_k;          // split_k is a power of two <= 16
      const int r_lo = z * rows_per;
      const int t = tid - 64;                        // 0..127
      constexpr int C4 = BN / 4;
      const uint32_t park_s = smem_u32(park);
      for (int idx = t; idx < rows_per * C4; idx += 128) {
        const int rl = idx / C4, c4 = idx - rl * C4;
        const int rr = r_lo + rl, n = n0 + 4 * c4;
        if (m0 + rr < a.M && n < a.N) {
          const uint32_t off = park_s + static_cast<uint32_t>((rr * PITCH + 4 * c4) * 4);
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int zz = 0; zz < a.split_k; ++zz) {
            uint32_t remote;
            float4 p;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(off), "r"(zz));
            asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(p.x), "=f"(p.y), "=f"(p.z), "=f"(p.w)
                         : "r"(remote));
            if (zz == 0) acc = p;
            else {
              acc.x += p.x;
              acc.y += p.y;
              acc.z += p.z;
              acc.w += p.w;
            }
          }
          // (finish() addresses rows through rofs of THIS thread's epilogue row: use explicit offsets here)
          const size_t ro = static_cast<size_t>(m0 + rr) * a.N;
          if (a.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + n));
            acc.x += b.x;
            acc.y += b.y;
            acc.z += b.z;
            acc.w += b.w;
          }
          acc.x = act_apply(acc.x, a.act);
          acc.y = act_apply(acc.y, a.act);
          acc.z = act_apply(acc.z, a.act);
          acc.w = act_apply(acc.w, a.act);
          if (a.residual) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(a.residual + ro + n));
            acc.x += r.x;
            acc.y += r.y;
            acc.z += r.z;
            acc.w += r.w;
          }
          *reinterpret_cast<float4*>(a.out + ro + n) = acc;
        }
      }
    }
    cluster_barrier();   // nobody leaves while a peer may still read its parked tile
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int BN>
int launch_k11(const CUtensorMap* tm, const LArgs& a, cudaStream_t st) {
  constexpr int STAGE = 2 * L_A_PLANE + 2 * BN * 128;
  constexpr int STAGES = (200 * 1024) / STAGE;
  constexpr int SMEM = 1024 + STAGES * STAGE + 256;
  static_assert(SMEM <= 227 * 1024, "K11 shared memory");
  static_assert(128 * (BN + 4) * 4 <= STAGES * STAGE, "the parked split-K tile fits the operand ring");
  FD_CUDA_OK(cudaFuncSetAttribute(k11_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((a.M + 127) / 128, (a.N + BN - 1) / BN, a.split_k);
  cfg.blockDim = dim3(L_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = static_cast<unsigned>(a.split_k);
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FD_CUDA_OK(cudaLaunchKernelEx(&cfg, k11_gemm_kernel<BN>, tm[0], tm[1], a));
  return FD_OK;
}

}  // namespace
}  // namespace fd

extern "C" int64_t fd_linear_x3_operand_bytes(int rows, int K) {
  // [h1 plane][h2 plane][inverse row scales], each 256-byte aligned
  const int64_t plane = (static_cast<int64_t>(rows) * K * 2 + 255) / 256 * 256;
  return 2 * plane + (static_cast<int64_t>(rows) * 4 + 255) / 256 * 256;
}

extern "C" int fd_linear_x3_split(const float* x_dev, int rows, int K, void* operand_dev, int64_t operand_bytes,
                                  void* stream) {
  using namespace fd;
  FD_REQUIRE(x_dev && operand_dev, "fd_linear_x3_split: NULL pointer");
  FD_REQUIRE(rows > 0 && K > 0 && K % L_KC == 0 && K <= 8192,
             "fd_linear_x3_split: need rows > 0 and K a multiple of 64 up to 8192 (rows=%d, K=%d)", rows, K);
  FD_REQUIRE(operand_bytes >= fd_linear_x3_operand_bytes(rows, K), "fd_linear_x3_split: operand buffer too small");
  FD_REQUIRE(reinterpret_cast<uintptr_t>(x_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(operand_dev) % 256 == 0,
             "fd_linear_x3_split: x must be 16-byte and the operand buffer 256-byte aligned");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  const int64_t plane = (static_cast<int64_t>(rows) * K * 2 + 255) / 256 * 256;
  uint8_t* base = static_cast<uint8_t*>(operand_dev);
  return launch_split<false>(x_dev, nullptr, nullptr, 0.f, reinterpret_cast<__half*>(base), reinterpret_cast<__half*>(base + plane),
                             reinterpret_cast<float*>(base + 2 * plane), rows, K, static_cast<cudaStream_t>(stream));
}

extern "C" int fd_linear_x3_split_ln(const float* x_dev, int rows, int K, const float* gamma_dev, const float* beta_dev,
                                     float eps, void* operand_dev, int64_t operand_bytes, void* stream) {
  using namespace fd;
  FD_REQUIRE(x_dev && gamma_dev && beta_dev && operand_dev, "fd_linear_x3_split_ln: NULL pointer");
  FD_REQUIRE(rows > 0 && K > 0 && K % L_KC == 0 && K <= 8192,
             "fd_linear_x3_split_ln: need rows > 0 and K a multiple of 64 up to 8192 (rows=%d, K=%d)", rows, K);
  FD_REQUIRE(operand_bytes >= fd_linear_x3_operand_bytes(rows, K), "fd_linear_x3_split_ln: operand buffer too small");
  FD_REQUIRE(reinterpret_cast<uintptr_t>(x_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(gamma_dev) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(beta_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(operand_dev) % 256 == 0,
             "fd_linear_x3_split_ln: x / gamma / beta must be 16-byte and the operand buffer 256-byte aligned");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  const int64_t plane = (static_cast<int64_t>(rows) * K * 2 + 255) / 256 * 256;
  uint8_t* base = static_cast<uint8_t*>(operand_dev);
  __half* h1 = reinterpret_cast<__half*>(base);
  __half* h2 = reinterpret_cast<__half*>(base + plane);
  float* inv = reinterpret_cast<float*>(base + 2 * plane);
  return launch_split<true>(x_dev, gamma_dev, beta_dev, eps, h1, h2, inv, rows, K, static_cast<cudaStream_t>(stream));
}

extern "C" int fd_linear_x3(const void* act_operand_dev, int M, const void* weight_operand_dev, int N, int K,
                            const float* bias_dev, int act, const float* residual_dev, float* out_dev, int split_k,
                            void* stream) {
  using namespace fd;
  FD_REQUIRE(act_operand_dev && weight_operand_dev && out_dev, "fd_linear_x3: NULL pointer");
  FD_REQUIRE(M > 0 && N > 0 && K > 0, "fd_linear_x3: non-positive shape");
  FD_REQUIRE(K % L_KC == 0 && N % 4 == 0, "fd_linear_x3: need K %% 64 == 0 and N %% 4 == 0 (K=%d, N=%d)", K, N);
  FD_REQUIRE(act >= 0 && act <= FD_LINEAR_ACT_GELU, "fd_linear_x3: unknown activation %d", act);
  FD_REQUIRE(split_k == 1 || split_k == 2 || split_k == 4 || split_k == 8, "fd_linear_x3: split_k=%d not in {1, 2, 4, 8}",
             split_k);
  FD_REQUIRE(reinterpret_cast<uintptr_t>(out_dev) % 16 == 0 && (!bias_dev || reinterpret_cast<uintptr_t>(bias_dev) % 16 == 0) &&
                 (!residual_dev || reinterpret_cast<uintptr_t>(residual_dev) % 16 == 0),
             "fd_linear_x3: out / bias / residual must be 16-byte aligned");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int num_kc = K / L_KC;
  while (split_k > num_kc) split_k >>= 1;
  const int bn = N >= 2048 || (N % 128 == 0 && static_cast<int64_t>((M + 127) / 128) * (N / 128) * split_k >= 96) ? 128 : 64;
  const uint8_t* ab = static_cast<const uint8_t*>(act_operand_dev);
  const uint8_t* wb = static_cast<const uint8_t*>(weight_operand_dev);
  const int64_t a_plane = (static_cast<int64_t>(M) * K * 2 + 255) / 256 * 256;
  const int64_t w_plane = (static_cast<int64_t>(N) * K * 2 + 255) / 256 * 256;
  CUtensorMap tm[2];
  for (int which = 0; which < 2; ++which) {
    // {K, rows, 2 planes}: one box {64, tile rows, 2} brings the h1 and h2 tiles of a K chunk, back to back in shared memory
    const bool is_w = which == 1;
    uint64_t dims[3] = {static_cast<uint64_t>(K), static_cast<uint64_t>(is_w ? N : M), 2};
    uint64_t strides[2] = {static_cast<uint64_t>(K) * 2, static_cast<uint64_t>(is_w ? w_plane : a_plane)};
    uint32_t box[3] = {L_KC, static_cast<uint32_t>(is_w ? bn : 128), 2};
    rc = encode_tmap(&tm[which], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, is_w ? static_cast<const void*>(wb) : ab, dims, strides, box,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FD_OK) return rc;
  }
  LArgs a;
  a.inv_a = reinterpret_cast<const float*>(ab + 2 * a_plane);
  a.inv_w = reinterpret_cast<const float*>(wb + 2 * w_plane);
  a.bias = bias_dev;
  a.residual = residual_dev;
  a.out = out_dev;
  a.M = M;
  a.N = N;
  a.K = K;
  a.act = act;
  a.split_k = split_k;
  a.kc_per_split = (num_kc + split_k - 1) / split_k;
  return bn == 128 ? launch_k11<128>(tm, a, st) : launch_k11<64>(tm, a, st);
}

// development aid / health check: 1 = a non-finite operand was split since the last call, >= 16 = a pipeline wait
// gave up; clears the flag.  Synchronises the device.
extern "C" int fd_linear_x3_flag(void) {
  int v = 0, z = 0;
  if (cudaMemcpyFromSymbol(&v, fd::g_k11_flag, sizeof(int)) != cudaSuccess) return FD_ERR_CUDA;
  if (v) cudaMemcpyToSymbol(fd::g_k11_flag, &z, sizeof(int));
  return v;
}
