// K5 / K6 -- normalisation + activation glue of the UNet forward that the denoising loop
// (pipeline/guide.py:56-58 -> unet(...)) spends most of its non-GEMM time in.
//
// The round-1 ncu launch list (profiles/r01/SUMMARY.md) showed that at B=1 the ATen GroupNorm path
// (NHWC->NCHW copy, RowwiseMoments, ComputeFusedParams, apply, separate SiLU, separate broadcast add
// of the time embedding) and the GEGLU gelu/mul pair on strided views are ~45 % of a denoising step.
// These are pure streaming kernels over L2-resident activations:
//
//   K5  y = act( GroupNorm( x + bias[n,c] ) * gamma[c] + beta[c] )        NHWC bf16 in / out
//       diffusers ResnetBlock2D: norm1 -> SiLU, (+ time_emb_proj) -> norm2 -> SiLU;
//       SpatialTransformer.norm, conv_norm_out.  Small (latency-bound) activations: ONE launch, a
//       thread-block cluster per (sample, group) with the slab resident in shared memory and the
//       statistics exchanged over DSMEM.  Large activations (batches, the VAE): three streaming
//       launches (stats / finalize / apply) with 16-byte row-coalesced accesses.  Fixed summation
//       order everywhere (bit-reproducible).
//   K7  y = x + h + bias[c]: the resnet residual add with conv2's bias folded in (cuDNN would
//       otherwise add every convolution bias with a separate broadcast kernel).
//   K6  out[m, f] = in[m, f] * gelu(in[m, F + f])                          diffusers GEGLU
#include <math.h>
#include <stdlib.h>

#include "fd_common.cuh"

namespace fd {
namespace {

constexpr int GN_THREADS = 256;   // 8 warps: warp w takes rows w, w+8, ... of the slab
constexpr int GN_MAX_GROUPS = 32;
constexpr int GN_VEC_MAX_SLABS = 512;               // streaming path: slabs per sample
constexpr int64_t GN_CLUSTER_MAX_BYTES = 20 << 20;  // larger activations take the streaming path

// x * sigmoid(x) with the approximate division (MUFU.RCP + FMUL, 2 ulp) instead of the IEEE one (~10 more instructions per
// element: the apply passes were instruction-bound, not memory-bound, with it); exp(-x) = inf gives x * 0.
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__device__ __forceinline__ float2 bf2_to_f2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
}

// ---- streaming path for large activations (batches > 1, the VAE): three launches, every global
// access a 16-byte vector over full rows.  A thread owns a fixed column of 8 channels (so its
// gamma / beta / bias live in registers) and walks down the rows of its slab.
//   stats    : grid (slabs, N), block (C/8, ty): per-slab, per-group (sum, sumsq) partials
//   finalize : grid (G, N): partials -> (mean, rstd), fixed order
//   apply    : same mapping as stats
struct GnVecArgs {
  const __nv_bfloat16* x;
  const __nv_bfloat16* bias;
  int64_t bias_stride;
  const __nv_bfloat16* gamma;
  const __nv_bfloat16* beta;
  float2* partial;  // [N, GN_VEC_MAX_SLABS, G]
  float2* stats;    // [N, G]
  __nv_bfloat16* y;
  int HW, C, G, slabs, rows_per_slab;
  float eps;
  int act_silu;
};

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[2 * k] = __uint_as_float(w[k] << 16);
    f[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
  }
}

__global__ void __launch_bounds__(512) k5_gn_vec_stats_kernel(const GnVecArgs a) {
  extern __shared__ float2 s_pair[];  // [ty][C/2] pair sums
  const int c8 = threadIdx.x, ty = threadIdx.y, TY = blockDim.y;
  const int n = blockIdx.y, slab = blockIdx.x;
  const int vec_per_row = a.C >> 3, pairs = a.C >> 1;
  const int r0 = slab * a.rows_per_slab, r1 = min(a.HW, r0 + a.rows_per_slab);
  const uint4* xb = reinterpret_cast<const uint4*>(a.x) + static_cast<size_t>(n) * a.HW * vec_per_row + c8;
  float b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (a.bias) unpack8(reinterpret_cast<const uint4*>(a.bias + static_cast<size_t>(n) * a.bias_stride)[c8], b);
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  int r = r0 + ty;
  for (; r + 3 * TY < r1; r += 4 * TY) {  // four rows in flight
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = xb[static_cast<size_t>(r + u * TY) * vec_per_row];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float x0 = f[2 * k] + b[2 * k], x1 = f[2 * k + 1] + b[2 * k + 1];
        s[k] += x0 + x1;
        q[k] += x0 * x0 + x1 * x1;
      }
    }
  }
  for (; r < r1; r += TY) {
    float f[8];
    unpack8(xb[static_cast<size_t>(r) * vec_per_row], f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float x0 = f[2 * k] + b[2 * k], x1 = f[2 * k + 1] + b[2 * k + 1];
      s[k] += x0 + x1;
      q[k] += x0 * x0 + x1 * x1;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) s_pair[ty * pairs + c8 * 4 + k] = make_float2(s[k], q[k]);
  __syncthreads();
  // group g <- its C/G/2 pairs over all ty, fixed order
  const int tid = ty * blockDim.x + c8, cg2 = (a.C / a.G) >> 1;
  if (tid < a.G) {
    float ss = 0.f, qq = 0.f;
    for (int t = 0; t < TY; ++t)
      for (int k = 0; k < cg2; ++k) {
        const float2 v = s_pair[t * pairs + tid * cg2 + k];
        ss += v.x;
        qq += v.y;
      }
    a.partial[(static_cast<size_t>(n) * GN_VEC_MAX_SLABS + slab) * a.G + tid] = make_float2(ss, qq);
  }
}

__global__ void __launch_bounds__(32) k5_gn_vec_finalize_kernel(const GnVecArgs a) {
  const int g = blockIdx.x, n = blockIdx.y, lane = threadIdx.x;
  float ss = 0.f, qq = 0.f;
  for (int sl = lane; sl < a.slabs; sl += 32) {
    const float2 v = a.partial[(static_cast<size_t>(n) * GN_VEC_MAX_SLABS + sl) * a.G + g];
    ss += v.x;
    qq += v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    qq += __shfl_xor_sync(0xffffffffu, qq, o);
  }
  if (lane == 0) {
    const float cnt = static_cast<float>(a.HW) * (a.C / a.G);
    const float mean = ss / cnt;
    a.stats[static_cast<size_t>(n) * a.G + g] = make_float2(mean, rsqrtf(fmaxf(qq / cnt - mean * mean, 0.f) + a.eps));
  }
}

__global__ void __launch_bounds__(512) k5_gn_vec_apply_kernel(const GnVecArgs a) {
  const int c8 = threadIdx.x, ty = threadIdx.y, TY = blockDim.y;
  const int n = blockIdx.y, slab = blockIdx.x;
  const int vec_per_row = a.C >> 3, cg = a.C / a.G;
  const int r0 = slab * a.rows_per_slab, r1 = min(a.HW, r0 + a.rows_per_slab);
  const size_t base = static_cast<size_t>(n) * a.HW * vec_per_row + c8;
  const uint4* xb = reinterpret_cast<const uint4*>(a.x) + base;
  uint4* yb = reinterpret_cast<uint4*>(a.y) + base;
  float sc[8], sh[8];
  {
    float gam[8], bet[8], b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    unpack8(reinterpret_cast<const uint4*>(a.gamma)[c8], gam);
    unpack8(reinterpret_cast<const uint4*>(a.beta)[c8], bet);
    if (a.bias) unpack8(reinterpret_cast<const uint4*>(a.bias + static_cast<size_t>(n) * a.bias_stride)[c8], b);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float2 st = a.stats[static_cast<size_t>(n) * a.G + (c8 * 8 + k) / cg];
      sc[k] = st.y * gam[k];
      sh[k] = (b[k] - st.x) * sc[k] + bet[k];
    }
  }
  auto emit = [&](const uint4& v, int row) {
    float f[8];
    unpack8(v, f);
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float y0 = f[2 * k] * sc[2 * k] + sh[2 * k], y1 = f[2 * k + 1] * sc[2 * k + 1] + sh[2 * k + 1];
      if (a.act_silu) {
        y0 = silu_fast(y0);
        y1 = silu_fast(y1);
      }
      __nv_bfloat162 p = __floats2bfloat162_rn(y0, y1);
      o[k] = *reinterpret_cast<uint32_t*>(&p);
    }
    yb[static_cast<size_t>(row) * vec_per_row] = make_uint4(o[0], o[1], o[2], o[3]);
  };
  int r = r0 + ty;
  for (; r + 3 * TY < r1; r += 4 * TY) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = xb[static_cast<size_t>(r + u * TY) * vec_per_row];
#pragma unroll
    for (int u = 0; u < 4; ++u) emit(v[u], r + u * TY);
  }
  for (; r < r1; r += TY) emit(xb[static_cast<size_t>(r) * vec_per_row], r);
}

// K5 fast path: one thread-block CLUSTER per (sample, set of `gset` adjacent groups).  gset is the
// smallest group count whose channel strip is a multiple of 16 bytes (C=320: 4 groups = 80 B per
// pixel), so every global access is a 16-byte vector (the first version took one group per cluster
// and read 20-byte strips with 4-byte accesses: 62 % sector efficiency, ~0.8 TB/s on L2-resident
// activations).  Each CTA of the cluster pulls its rows of the strip into shared memory once, the
// per-group (sum, sumsq) partials are exchanged through distributed shared memory, and the slab is
// normalised straight out of shared memory: x is read from L2/HBM exactly once, one launch, no
// workspace, fixed reduction order.
// Thread mapping: sv = strip bytes / 16 vectors per pixel; thread t owns vector column t % sv (so
// the groups, gamma, beta and bias of its 8 channels are fixed) and rows t / sv, + RP, ...
struct GnClusterArgs {
  const __nv_bfloat16* x;
  const __nv_bfloat16* h;        // optional residual branch: the kernel normalises s = bf16(x + (h + rbias[c])), K7's sum,
  const __nv_bfloat16* rbias;    // and writes s to sum_out (fd_add_groupnorm_act: K7 folded into the GroupNorm that follows it)
  __nv_bfloat16* sum_out;
  const __nv_bfloat16* bias;
  int64_t bias_stride;
  const __nv_bfloat16* gamma;
  const __nv_bfloat16* beta;
  __nv_bfloat16* y;
  int HW, C, G, rows_per_cta, gset, sv;
  float eps;
  int act_silu;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// split barrier used only to know that every CTA of the cluster has STARTED (its shared memory may be
// written remotely after that): arrive at kernel entry, wait right before the first remote store
__device__ __forceinline__ void cluster_arrive_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store into the same shared-memory variable of CTA `rank` of this cluster
__device__ __forceinline__ void dsmem_st_f2(float2* local_ptr, uint32_t rank, float2 v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(v.x), "f"(v.y) : "memory");
}

constexpr int GN_MAX_GSET = 8;    // groups per cluster
constexpr int GN_MAX_SV = 32;     // 16-byte vectors per pixel strip (<= 512 B)

template <bool RES>   // RES: the residual branch (a.h) is summed in; rows go through in batches of 4 + 4 loads instead of 8
__global__ void __launch_bounds__(GN_THREADS) k5_gn_cluster_kernel(const GnClusterArgs a) {
  extern __shared__ uint4 slab[];                      // [rows_per_cta][sv]
  __shared__ float2 s_col[GN_THREADS * 4];             // [RP][sv * 4] pair-column partials, tree-reduced over RP
  __shared__ float2 all_partial[8][GN_MAX_GSET];       // [cluster rank][group]: every CTA's partials, pushed by the peers
  __shared__ float2 s_stats[GN_MAX_GSET];
  const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
  cluster_arrive_relaxed();
  const int sets = a.G / a.gset;
  const int set_id = blockIdx.x / csize;  // (n, group set) flattened
  const int n = set_id / sets, gs = set_id - n * sets;
  const int sv = a.sv, RP = GN_THREADS / sv, cg2 = (a.C / a.G) >> 1, cg = a.C / a.G;
  const int vec_per_row = a.C >> 3;
  const int r0 = crank * a.rows_per_cta;
  const int rows = max(0, min(a.HW, r0 + a.rows_per_cta) - r0);
  const int tid = threadIdx.x;
  const bool active = tid < RP * sv;
  const int j = tid % sv, rr = tid / sv;
  const int vec0 = gs * sv + j;  // this thread's 16-byte column within the pixel row
  const size_t gbase = (static_cast<size_t>(n) * a.HW + r0) * vec_per_row + vec0;
  const uint4* xb = reinterpret_cast<const uint4*>(a.x) + gbase;

  float b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float rb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const uint4* hb = RES ? reinterpret_cast<const uint4*>(a.h) + gbase : nullptr;
  uint4* sb = RES ? reinterpret_cast<uint4*>(a.sum_out) + gbase : nullptr;
  uint4 gam_raw = make_uint4(0, 0, 0, 0), bet_raw = gam_raw;
  if (active) {
    if (a.bias) unpack8(reinterpret_cast<const uint4*>(a.bias + static_cast<size_t>(n) * a.bias_stride)[vec0], b);
    if (RES) unpack8(reinterpret_cast<const uint4*>(a.rbias)[vec0], rb);
    gam_raw = reinterpret_cast<const uint4*>(a.gamma)[vec0];  // needed after the statistics
    bet_raw = reinterpret_cast<const uint4*>(a.beta)[vec0];
  }
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
    constexpr int U = RES ? 4 : 8;
    for (int r = rr; r < rows; r += U * RP) {  // up to eight loads in flight
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (r + u * RP < rows) v[u] = xb[static_cast<size_t>(r + u * RP) * vec_per_row];
      if constexpr (RES) {   // s = bf16(x + (h + rbias)), exactly K7's arithmetic; s is what gets stored, normalised and kept in the slab
        uint4 hv[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (r + u * RP < rows) hv[u] = hb[static_cast<size_t>(r + u * RP) * vec_per_row];
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (r + u * RP < rows) {
            float fx[8], fh[8];
            unpack8(v[u], fx);
            unpack8(hv[u], fh);
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              __nv_bfloat162 p2 = __floats2bfloat162_rn(fx[2 * k] + (fh[2 * k] + rb[2 * k]), fx[2 * k + 1] + (fh[2 * k + 1] + rb[2 * k + 1]));
              o[k] = *reinterpret_cast<uint32_t*>(&p2);
            }
            v[u] = make_uint4(o[0], o[1], o[2], o[3]);
            sb[static_cast<size_t>(r + u * RP) * vec_per_row] = v[u];
          }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (r + u * RP < rows) {
          slab[(r + u * RP) * sv + j] = v[u];
          float f[8];
          unpack8(v[u], f);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float x0 = f[2 * k] + b[2 * k], x1 = f[2 * k + 1] + b[2 * k + 1];
            s[k] += x0 + x1;
            q[k] += x0 * x0 + x1 * x1;
          }
        }
    }
  }
  // ---- CTA reduction in a fixed order: tree over the RP row-threads of every pair column, then
  // each group adds its C/G/2 pair columns
  const int ncol = sv * 4;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (active) s_col[rr * ncol + j * 4 + k] = make_float2(s[k], q[k]);
  __syncthreads();
  int span = 1;
  while (span < RP) span <<= 1;
  for (int st = span >> 1; st > 0; st >>= 1) {
    if (active && rr < st && rr + st < RP) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 u = s_col[rr * ncol + j * 4 + k];
        const float2 w = s_col[(rr + st) * ncol + j * 4 + k];
        u.x += w.x;
        u.y += w.y;
        s_col[rr * ncol + j * 4 + k] = u;
      }
    }
    __syncthreads();
  }
  cluster_wait();  // every peer is running (it arrived at entry); normally long satisfied by now
  if (tid < a.gset) {
    float ss = 0.f, qq = 0.f;
    for (int k = 0; k < cg2; ++k) {
      const float2 v = s_col[tid * cg2 + k];
      ss += v.x;
      qq += v.y;
    }
    // push model: write this CTA's partials into every peer's shared memory (and its own), so that
    // after the barrier each CTA only reads locally and nobody has to wait for its peers before
    // exiting (a pull model needs a second cluster barrier at the end of the kernel)
    const float2 mine = make_float2(ss, qq);
    for (uint32_t r = 0; r < csize; ++r) dsmem_st_f2(&all_partial[crank][tid], r, mine);
  }
  cluster_sync_all();  // release / acquire: every CTA's partials have landed everywhere
  if (tid < a.gset) {
    float ts = 0.f, tq = 0.f;
    for (uint32_t r = 0; r < csize; ++r) {  // same order on every CTA => identical statistics
      ts += all_partial[r][tid].x;
      tq += all_partial[r][tid].y;
    }
    const float cnt = static_cast<float>(a.HW) * cg;
    const float mean = ts / cnt;
    s_stats[tid] = make_float2(mean, rsqrtf(fmaxf(tq / cnt - mean * mean, 0.f) + a.eps));
  }
  __syncthreads();

  if (active) {
    float sc[8], sh[8];
    {
      float gam[8], bet[8];
      unpack8(gam_raw, gam);
      unpack8(bet_raw, bet);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float2 st = s_stats[(j * 8 + k) / cg];
        sc[k] = st.y * gam[k];
        sh[k] = (b[k] - st.x) * sc[k] + bet[k];
      }
    }
    uint4* yb = reinterpret_cast<uint4*>(a.y) + gbase;
    for (int r = rr; r < rows; r += RP) {
      float f[8];
      unpack8(slab[r * sv + j], f);
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float y0 = f[2 * k] * sc[2 * k] + sh[2 * k], y1 = f[2 * k + 1] * sc[2 * k + 1] + sh[2 * k + 1];
        if (a.act_silu) {
          y0 = silu_fast(y0);
          y1 = silu_fast(y1);
        }
        __nv_bfloat162 p2 = __floats2bfloat162_rn(y0, y1);
        o[k] = *reinterpret_cast<uint32_t*>(&p2);
      }
      yb[static_cast<size_t>(r) * vec_per_row] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// K7: y = x + h + bias[c]   (residual add with the convolution bias folded in), NHWC bf16.  h == nullptr: y = x + bias[c] --
// the bias of conv_in and of the down / upsample convolutions, which cuDNN would otherwise add with ATen's broadcasting
// (non-vectorised) elementwise kernel: 7.9 us per launch at B = 1 against ~3 us here.  y may alias x.
__global__ void __launch_bounds__(256) k7_add_bias_residual_kernel(const uint4* x,
                                                                   const uint4* __restrict__ h,
                                                                   const uint4* __restrict__ bias, uint4* y,
                                                                   int64_t total8, int c8) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 xv = x[i], hv = h ? h[i] : make_uint4(0u, 0u, 0u, 0u), bv = bias[i % c8];
    const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w}, hs[4] = {hv.x, hv.y, hv.z, hv.w},
                   bs[4] = {bv.x, bv.y, bv.z, bv.w};
    uint32_t os[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = bf2_to_f2(xs[k]), b = bf2_to_f2(hs[k]), c = bf2_to_f2(bs[k]);
      __nv_bfloat162 o = __floats2bfloat162_rn(a.x + (b.x + c.x), a.y + (b.y + c.y));
      os[k] = *reinterpret_cast<uint32_t*>(&o);
    }
    y[i] = make_uint4(os[0], os[1], os[2], os[3]);
  }
}

// K8: s = x + y (optional), n = LayerNorm(s) * gamma + beta.  One warp per row of C channels
// (C/64 bf16 pairs per lane, kept in registers: exact two-pass mean / variance).  Writes the
// residual sum and the normalised row in the same pass (BasicTransformerBlock's
// `x = attn(norm(x)) + x` followed by the next `norm(x)`).
template <int PAIRS_PER_LANE>
__global__ void __launch_bounds__(256) k8_add_layernorm_kernel(const uint32_t* __restrict__ x,
                                                               const uint32_t* __restrict__ y,
                                                               const uint32_t* __restrict__ gamma,
                                                               const uint32_t* __restrict__ beta,
                                                               uint32_t* __restrict__ sum_out,
                                                               uint32_t* __restrict__ norm_out, int64_t M, float eps,
                                                               int64_t sum_stride_pairs) {
  constexpr int PAIRS = PAIRS_PER_LANE * 32;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const uint32_t* xr = x + row * PAIRS;
  uint32_t xv[PAIRS_PER_LANE], yv[PAIRS_PER_LANE], gv[PAIRS_PER_LANE], bv[PAIRS_PER_LANE];
#pragma unroll
  for (int i = 0; i < PAIRS_PER_LANE; ++i) xv[i] = xr[lane + 32 * i];
  // gamma / beta are only needed after both reductions: ask for them now so their latency hides there
  // (few-row launches, C >= 640: 4.2 -> 2.9 us; at C = 320 with 8192 rows the later load is 0.2 us faster)
  constexpr bool HOIST = PAIRS_PER_LANE >= 10;
  if constexpr (HOIST) {
#pragma unroll
    for (int i = 0; i < PAIRS_PER_LANE; ++i) {
      gv[i] = gamma[lane + 32 * i];
      bv[i] = beta[lane + 32 * i];
    }
  }
  if (y) {
    const uint32_t* yr = y + row * PAIRS;
#pragma unroll
    for (int i = 0; i < PAIRS_PER_LANE; ++i) yv[i] = yr[lane + 32 * i];
  }
  float2 v[PAIRS_PER_LANE];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PAIRS_PER_LANE; ++i) {
    v[i] = bf2_to_f2(xv[i]);
    if (y) {
      const float2 t = bf2_to_f2(yv[i]);
      // the residual stream is bf16: normalise exactly what is stored
      __nv_bfloat162 r = __floats2bfloat162_rn(v[i].x + t.x, v[i].y + t.y);
      const uint32_t rb = *reinterpret_cast<uint32_t*>(&r);
      if (sum_out) sum_out[row * sum_stride_pairs + lane + 32 * i] = rb;
      v[i] = bf2_to_f2(rb);
    }
    s += v[i].x + v[i].y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (2.0f * PAIRS);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PAIRS_PER_LANE; ++i) {
    const float dx = v[i].x - mean, dy = v[i].y - mean;
    q += dx * dx + dy * dy;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (2.0f * PAIRS) + eps);
#pragma unroll
  for (int i = 0; i < PAIRS_PER_LANE; ++i) {
    const float2 g = bf2_to_f2(HOIST ? gv[i] : gamma[lane + 32 * i]), b = bf2_to_f2(HOIST ? bv[i] : beta[lane + 32 * i]);
    __nv_bfloat162 o = __floats2bfloat162_rn((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
    norm_out[row * PAIRS + lane + 32 * i] = *reinterpret_cast<uint32_t*>(&o);
  }
}

// K6: GEGLU.  in [M, 2F] bf16 -> out [M, F] bf16; 8 elements (16 B) per thread.
__global__ void __launch_bounds__(256) k6_geglu_kernel(const __nv_bfloat16* __restrict__ in,
                                                       __nv_bfloat16* __restrict__ out, int64_t M, int F) {
  const int f8 = F >> 3;
  const int64_t total = M * f8;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t m = i / f8;
    const int c = static_cast<int>(i - m * f8) << 3;
    const uint4 xv = *reinterpret_cast<const uint4*>(in + m * 2 * F + c);
    const uint4 gv = *reinterpret_cast<const uint4*>(in + m * 2 * F + F + c);
    const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w};
    const uint32_t gs[4] = {gv.x, gv.y, gv.z, gv.w};
    uint32_t os[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 x = bf2_to_f2(xs[k]), g = bf2_to_f2(gs[k]);
      // exact (erf) GELU, as torch.nn.functional.gelu's default
      const float gx = 0.5f * g.x * (1.0f + erff(g.x * 0.70710678118654752f));
      const float gy = 0.5f * g.y * (1.0f + erff(g.y * 0.70710678118654752f));
      __nv_bfloat162 o = __floats2bfloat162_rn(x.x * gx, x.y * gy);
      os[k] = *reinterpret_cast<uint32_t*>(&o);
    }
    *reinterpret_cast<uint4*>(out + m * F + c) = make_uint4(os[0], os[1], os[2], os[3]);
  }
}

}  // namespace

}  // namespace fd

extern "C" int64_t fd_groupnorm_act_workspace_bytes(int N, int HW, int C, int G) {
  // streaming path: per-slab per-group partials + per-group statistics (float2 each)
  (void)HW;
  (void)C;
  return (static_cast<int64_t>(N) * fd::GN_VEC_MAX_SLABS * G + static_cast<int64_t>(N) * G) * 8;
}

extern "C" int fd_add_bias_residual(const void* x_bf16_dev, const void* h_bf16_dev, const void* bias_bf16_dev,
                                    void* y_bf16_dev, int64_t n_elem, int C, void* stream);

static int fd_groupnorm_impl(const void* x_bf16_dev, const void* bias_bf16_dev, const void* gamma_bf16_dev,
                             const void* beta_bf16_dev, void* workspace_dev, void* y_bf16_dev, int N, int HW,
                             int C, int G, float eps, int act_silu, int64_t bias_row_stride, void* stream,
                             const void* h_bf16_dev, const void* rbias_bf16_dev, void* sum_out_bf16_dev) {
  using namespace fd;
  FD_REQUIRE(x_bf16_dev && gamma_bf16_dev && beta_bf16_dev && workspace_dev && y_bf16_dev,
             "fd_groupnorm_act: NULL pointer");
  int rc = FD_OK;
  FD_REQUIRE(N > 0 && HW > 0 && C > 0, "fd_groupnorm_act: non-positive shape");
  FD_REQUIRE(G > 0 && G <= GN_MAX_GROUPS && C % G == 0 && (C / G) % 2 == 0 && C % 64 == 0 && C <= 4096,
             "fd_groupnorm_act: need G <= %d, C %% G == 0, even channels per group, C %% 64 == 0, C <= 4096 (C=%d G=%d)",
             GN_MAX_GROUPS, C, G);
  FD_REQUIRE(N <= 65535, "fd_groupnorm_act: N exceeds grid limits");
  auto mis16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 != 0; };
  FD_REQUIRE(!mis16(x_bf16_dev) && !mis16(y_bf16_dev) && !mis16(gamma_bf16_dev) && !mis16(beta_bf16_dev) &&
                 reinterpret_cast<uintptr_t>(workspace_dev) % 8 == 0,
             "fd_groupnorm_act: x / y / gamma / beta must be 16-byte aligned, workspace 8-byte aligned");
  FD_REQUIRE(!bias_bf16_dev || (bias_row_stride >= C && bias_row_stride % 8 == 0 && !mis16(bias_bf16_dev)),
             "fd_groupnorm_act: bias row stride must be a multiple of 8 and >= C, pointer 16-byte aligned");
  rc = check_device();
  if (rc != FD_OK) return rc;
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x_bf16_dev);
  const __nv_bfloat16* bp = static_cast<const __nv_bfloat16*>(bias_bf16_dev);
  const __nv_bfloat16* gp = static_cast<const __nv_bfloat16*>(gamma_bf16_dev);
  const __nv_bfloat16* tp = static_cast<const __nv_bfloat16*>(beta_bf16_dev);
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y_bf16_dev);
  const int64_t total_bytes = static_cast<int64_t>(N) * HW * C * 2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  constexpr int64_t gn_min_cta_bytes = 4096;  // measured: 0 / 4 / 8 KB within 0.1 us, 20 KB and more slower
  static const int64_t cluster_max = [] {  // development override of the path switch
    const char* e = getenv("FD_GN_CLUSTER_MAX_BYTES");
    return e ? static_cast<int64_t>(atoll(e)) : GN_CLUSTER_MAX_BYTES;
  }();
  if (total_bytes <= cluster_max) {
    // small / medium activations: a set of groups fits the shared memory of a cluster of <= 8 CTAs
    const int cgc = C / G;
    int gset = 1;
    while (gset <= GN_MAX_GSET && (gset * cgc * 2) % 16 != 0) gset *= 2;
    const int sv = gset * cgc * 2 / 16;
    if (gset <= GN_MAX_GSET && G % gset == 0 && sv >= 1 && sv <= GN_MAX_SV) {
      const int64_t set_bytes = static_cast<int64_t>(HW) * sv * 16;
      const int64_t n_sets = static_cast<int64_t>(N) * (G / gset);
      int cl = 1;
      // do not split a strip below 4 KB per CTA just to fill the SMs: a wider cluster barrier costs more
      while (cl < 8 && (set_bytes / cl > 64 * 1024 || (n_sets * cl < sm_count() && set_bytes / cl > gn_min_cta_bytes))) cl *= 2;
      while (cl > 1 && HW < cl * 8) cl /= 2;
      static const int force_cl = [] {  // development override (A/B of the cluster width)
        const char* e = getenv("FD_GN_FORCE_CL");
        return e ? atoi(e) : 0;
      }();
      if (force_cl > 0 && set_bytes / force_cl <= 96 * 1024 && HW >= force_cl * 8) cl = force_cl;
      const int rows_per_cta = (HW + cl - 1) / cl;
      const int64_t smem = static_cast<int64_t>(rows_per_cta) * sv * 16;
      // > 96 KB per CTA (C=960 at 64x64: 123 KB) measured 2.4x slower than the streaming path
      if (smem <= 96 * 1024 && n_sets * cl <= 0x7fffffff) {
        GnClusterArgs c;
        c.x = xp; c.bias = bp; c.bias_stride = bias_row_stride; c.gamma = gp; c.beta = tp; c.y = yp;
        c.h = static_cast<const __nv_bfloat16*>(h_bf16_dev);
        c.rbias = static_cast<const __nv_bfloat16*>(rbias_bf16_dev);
        c.sum_out = static_cast<__nv_bfloat16*>(sum_out_bf16_dev);
        c.HW = HW; c.C = C; c.G = G; c.rows_per_cta = rows_per_cta; c.gset = gset; c.sv = sv; c.eps = eps;
        c.act_silu = act_silu;
        static thread_local int attr_device = -1;
        int dev = 0;
        FD_CUDA_OK(cudaGetDevice(&dev));
        if (attr_device != dev) {
          FD_CUDA_OK(cudaFuncSetAttribute(k5_gn_cluster_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
          FD_CUDA_OK(cudaFuncSetAttribute(k5_gn_cluster_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
          attr_device = dev;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(static_cast<unsigned>(n_sets * cl));
        cfg.blockDim = dim3(GN_THREADS);
        cfg.dynamicSmemBytes = static_cast<size_t>(smem);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cl;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (c.h) FD_CUDA_OK(cudaLaunchKernelEx(&cfg, k5_gn_cluster_kernel<true>, c));
        else FD_CUDA_OK(cudaLaunchKernelEx(&cfg, k5_gn_cluster_kernel<false>, c));
        return FD_OK;
      }
    }
  }
  // streaming path (a residual branch is summed by K7 first: three more launches follow anyway)
  if (h_bf16_dev) {
    rc = fd_add_bias_residual(x_bf16_dev, h_bf16_dev, rbias_bf16_dev, sum_out_bf16_dev, static_cast<int64_t>(N) * HW * C, C, stream);
    if (rc != FD_OK) return rc;
    xp = static_cast<const __nv_bfloat16*>(sum_out_bf16_dev);
  }
  GnVecArgs v;
  v.x = xp; v.bias = bp; v.bias_stride = bias_row_stride; v.gamma = gp; v.beta = tp; v.y = yp;
  v.HW = HW; v.C = C; v.G = G; v.eps = eps; v.act_silu = act_silu;
  const int tx = C / 8;
  int ty = 256 / tx;
  if (ty < 1) ty = 1;
  // ~64 KB of input per CTA, at most GN_VEC_MAX_SLABS slabs per sample
  int rows = (64 * 1024) / (C * 2);
  if (rows < 4 * ty) rows = 4 * ty;
  int slabs = (HW + rows - 1) / rows;
  if (slabs > GN_VEC_MAX_SLABS) slabs = GN_VEC_MAX_SLABS;
  v.rows_per_slab = (HW + slabs - 1) / slabs;
  v.slabs = (HW + v.rows_per_slab - 1) / v.rows_per_slab;
  v.partial = static_cast<float2*>(workspace_dev);
  v.stats = v.partial + static_cast<size_t>(N) * GN_VEC_MAX_SLABS * G;
  dim3 grid(v.slabs, N), block(tx, ty);
  const size_t smem = static_cast<size_t>(ty) * (C / 2) * sizeof(float2);
  k5_gn_vec_stats_kernel<<<grid, block, smem, st>>>(v);
  k5_gn_vec_finalize_kernel<<<dim3(G, N), 32, 0, st>>>(v);
  k5_gn_vec_apply_kernel<<<grid, block, 0, st>>>(v);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}

extern "C" int fd_groupnorm_act(const void* x_bf16_dev, const void* bias_bf16_dev, const void* gamma_bf16_dev,
                                const void* beta_bf16_dev, void* workspace_dev, void* y_bf16_dev, int N, int HW,
                                int C, int G, float eps, int act_silu, int64_t bias_row_stride, void* stream) {
  return fd_groupnorm_impl(x_bf16_dev, bias_bf16_dev, gamma_bf16_dev, beta_bf16_dev, workspace_dev, y_bf16_dev, N, HW, C, G, eps,
                           act_silu, bias_row_stride, stream, nullptr, nullptr, nullptr);
}

// s = x + h + rbias[c] (K7) and y = act(GroupNorm(s)) in one launch where the cluster kernel applies (else K7 + the
// streaming GroupNorm): a resnet's residual add folded into the GroupNorm of the block that consumes its output.
extern "C" int fd_add_groupnorm_act(const void* x_bf16_dev, const void* h_bf16_dev, const void* rbias_bf16_dev,
                                    void* sum_out_bf16_dev, const void* gamma_bf16_dev, const void* beta_bf16_dev,
                                    void* workspace_dev, void* y_bf16_dev, int N, int HW, int C, int G, float eps,
                                    int act_silu, void* stream) {
  using namespace fd;
  FD_REQUIRE(h_bf16_dev && rbias_bf16_dev && sum_out_bf16_dev, "fd_add_groupnorm_act: NULL pointer");
  auto mis16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 != 0; };
  FD_REQUIRE(!mis16(h_bf16_dev) && !mis16(rbias_bf16_dev) && !mis16(sum_out_bf16_dev),
             "fd_add_groupnorm_act: h / rbias / sum_out must be 16-byte aligned");
  return fd_groupnorm_impl(x_bf16_dev, nullptr, gamma_bf16_dev, beta_bf16_dev, workspace_dev, y_bf16_dev, N, HW, C, G, eps,
                           act_silu, 0, stream, h_bf16_dev, rbias_bf16_dev, sum_out_bf16_dev);
}

extern "C" int fd_add_bias_residual(const void* x_bf16_dev, const void* h_bf16_dev, const void* bias_bf16_dev,
                                    void* y_bf16_dev, int64_t n_elem, int C, void* stream) {
  using namespace fd;
  FD_REQUIRE(x_bf16_dev && bias_bf16_dev && y_bf16_dev, "fd_add_bias_residual: NULL pointer");
  FD_REQUIRE(C > 0 && C % 8 == 0 && n_elem > 0 && n_elem % C == 0,
             "fd_add_bias_residual: need C %% 8 == 0 and n_elem a multiple of C");
  auto mis = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 != 0; };
  FD_REQUIRE(!mis(x_bf16_dev) && !mis(h_bf16_dev) && !mis(bias_bf16_dev) && !mis(y_bf16_dev),
             "fd_add_bias_residual: pointers must be 16-byte aligned");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  const int sms = sm_count();
  if (sms <= 0) return set_error(FD_ERR_CUDA, "fd_add_bias_residual: cannot query SM count");
  const int64_t total8 = n_elem / 8;
  int64_t want = (total8 + 255) / 256, cap = static_cast<int64_t>(sms) * 8;
  k7_add_bias_residual_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0,
                                static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x_bf16_dev), static_cast<const uint4*>(h_bf16_dev),
      static_cast<const uint4*>(bias_bf16_dev), static_cast<uint4*>(y_bf16_dev), total8, C / 8);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}

extern "C" int fd_add_layernorm(const void* x_bf16_dev, const void* y_bf16_dev, const void* gamma_bf16_dev,
                                const void* beta_bf16_dev, void* sum_out_bf16_dev, void* norm_out_bf16_dev,
                                int64_t M, int C, float eps, int64_t sum_out_row_stride, void* stream) {
  using namespace fd;
  FD_REQUIRE(x_bf16_dev && gamma_bf16_dev && beta_bf16_dev && norm_out_bf16_dev, "fd_add_layernorm: NULL pointer");
  FD_REQUIRE(M > 0 && (C == 320 || C == 640 || C == 1280), "fd_add_layernorm: need M > 0 and C in {320, 640, 1280}");
  FD_REQUIRE(!sum_out_bf16_dev || y_bf16_dev, "fd_add_layernorm: sum_out without y");
  if (sum_out_row_stride <= 0) sum_out_row_stride = C;
  FD_REQUIRE(sum_out_row_stride >= C && sum_out_row_stride % 2 == 0, "fd_add_layernorm: sum_out row stride %lld must be even and >= C",
             static_cast<long long>(sum_out_row_stride));
  const int64_t ssp = sum_out_row_stride / 2;
  int rc = check_device();
  if (rc != FD_OK) return rc;
  const unsigned grid = static_cast<unsigned>((M + 7) / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto X = static_cast<const uint32_t*>(x_bf16_dev), Y = static_cast<const uint32_t*>(y_bf16_dev);
  auto Gm = static_cast<const uint32_t*>(gamma_bf16_dev), Bt = static_cast<const uint32_t*>(beta_bf16_dev);
  auto So = static_cast<uint32_t*>(sum_out_bf16_dev), No = static_cast<uint32_t*>(norm_out_bf16_dev);
  if (C == 320) k8_add_layernorm_kernel<5><<<grid, 256, 0, st>>>(X, Y, Gm, Bt, So, No, M, eps, ssp);
  else if (C == 640) k8_add_layernorm_kernel<10><<<grid, 256, 0, st>>>(X, Y, Gm, Bt, So, No, M, eps, ssp);
  else k8_add_layernorm_kernel<20><<<grid, 256, 0, st>>>(X, Y, Gm, Bt, So, No, M, eps, ssp);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}

extern "C" int fd_geglu(const void* in_bf16_dev, void* out_bf16_dev, int64_t M, int F, void* stream) {
  using namespace fd;
  FD_REQUIRE(in_bf16_dev && out_bf16_dev, "fd_geglu: NULL pointer");
  FD_REQUIRE(M > 0 && F > 0 && F % 8 == 0, "fd_geglu: need M > 0 and F a positive multiple of 8");
  FD_REQUIRE(reinterpret_cast<uintptr_t>(in_bf16_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(out_bf16_dev) % 16 == 0,
             "fd_geglu: pointers must be 16-byte aligned");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  const int sms = sm_count();
  if (sms <= 0) return set_error(FD_ERR_CUDA, "fd_geglu: cannot query SM count");
  const int64_t total = M * (F >> 3);
  int64_t want = (total + 255) / 256, cap = static_cast<int64_t>(sms) * 8;
  k6_geglu_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in_bf16_dev), static_cast<__nv_bfloat16*>(out_bf16_dev), M, F);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}
