// K1 -- guide-token x text-token similarity map, Linear / Clustered / Threshold re-weighting and
// the lerp blend, in one kernel per (prompt, parameter-chunk).
//
// Replaces (reference, pure Python, ~0.1-0.3 s per blend on 8 CPU cores):
//   guidance.py:23-85    _map_emb             257 row-matmuls, 19.5k .item() syncs, Python sort
//   guidance.py:88-172   _traverse_a_to_b / _clustered_guidance
//   guidance.py:175-193  _blend_weights
//   guidance.py:215-272  Tweener.tween
// Algorithm spec: SURVEY.md 3.6 (verified bit-level against the reference by the oracle tests).
//
// Structure of one CTA (384 threads, 1 CTA / SM):
//   1. GEMM  D[i,j] = <guide_i, text_j>  (A<=384 x T<=80 x D) on tcgen05, kind::tf32 with the
//      3-pass hi/lo split (hi*hi + lo*hi + hi*lo) so the logits are fp32-equivalent
//      (a single bf16/tf32 pass flips arg-max / threshold decisions).
//      Operands go global -> registers (split + sum of squares for the L2 norms) -> shared memory
//      in the K-major SWIZZLE_128B UMMA layout; next chunk's global loads overlap the MMAs.
//      Accumulators: up to 3 tiles of 128 lanes x 80 columns in TMEM.
//   2. Softmax over the text tokens: one thread per TMEM lane (= guide token), no shuffles.
//      P^T is parked in shared memory (aliasing the operand staging area).
//   3. Column arg-max / greedy no-reuse assignment / direct mapping: warp-shuffle reductions.
//   4. Weight heuristics for <=96 tokens in one warp (ballot bitmasks for peaks / valleys).
//   5. 3-way select / lerp of the [T, D] rows, 128-bit coalesced.
#include <math.h>

#include "fd_common.cuh"

namespace fd {
namespace {

constexpr int K1_THREADS = 384;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int NPAD = 80;          // UMMA N (text tokens padded)
constexpr int MAX_TILES = 3;      // guide tokens padded to <= 3 x 128
constexpr int KC = 32;            // fp32 per K chunk (one 128 B swizzle row)
constexpr int A_TILE_BYTES = 128 * 128;
constexpr int B_TILE_BYTES = NPAD * 128;
constexpr int STAGE_BYTES = 2 * MAX_TILES * A_TILE_BYTES + 2 * B_TILE_BYTES;  // 118,784
constexpr int PT_STRIDE = MAX_TILES * 128;                                   // floats per P^T row
constexpr int MAXT = 80;
constexpr int A_ITEMS = (MAX_TILES * 128 * 8 + K1_THREADS - 1) / K1_THREADS;  // 8
constexpr int B_ITEMS = (NPAD * 8 + K1_THREADS - 1) / K1_THREADS;            // 2

static_assert((MAXT - 1) * PT_STRIDE * 4 <= STAGE_BYTES + 8192, "P^T must fit the staging area");

struct K1Args {
  const float* text;     // [n_text, T, D]
  const float* guide;    // [guide_batch, A, D]
  int n_text, guide_batch, T, A, D;
  const fd_tween_params* params;  // device [n_params]
  const float* lin_w;             // device [n_params, T]
  int n_params, params_per_cta;
  float* out;
  float* map_s;
  int32_t* map_idx;
  float* weights;
  int32_t* status;
  float* sim;
};

struct K1Smem {
  // staging / P^T area comes first (1024-aligned), then this struct
  float inv_norm_a[MAX_TILES * 128];
  float inv_norm_b[NPAD];
  float map_s[MAXT + 16];
  int map_idx[MAXT + 16];
  float col_s[MAXT + 16];   // per-column best over unused rows (no-reuse greedy)
  int col_i[MAXT + 16];
  unsigned char assigned[MAXT + 16];
  unsigned int used[MAX_TILES * 128 / 32];  // bitmask of guide tokens already consumed
  float iw[MAXT + 16];
  int sel[MAXT + 16];
  int pick_r, pick_i, flag;
  uint64_t mma_bar;
  uint32_t tmem_slot;
};

constexpr int K1_SMEM_BYTES = 1024 + STAGE_BYTES + 8192 + sizeof(K1Smem);

__device__ __forceinline__ void argmax_combine(float& s, int& i, float os, int oi) {
  // larger s wins, ties -> lower index; index < 0 means "nothing"
  if (oi >= 0 && (i < 0 || os > s || (os == s && oi < i))) {
    s = os;
    i = oi;
  }
}

// arg-max of column r of P (stored transposed: pt[r * PT_STRIDE + i]) over guide tokens i < A that
// are not in `used` (used == nullptr: all allowed).  Whole warp cooperates; result in all lanes.
__device__ __forceinline__ void warp_col_argmax(const float* pt_row, int A, const unsigned int* used,
                                                int lane, float& best_s, int& best_i) {
  float s = -1.0f;
  int idx = -1;
  for (int i = lane; i < A; i += 32) {
    if (used && ((used[i >> 5] >> (i & 31)) & 1u)) continue;
    const float v = pt_row[i];
    if (idx < 0 || v > s) {  // ascending i within a lane: strict > keeps the lowest index
      s = v;
      idx = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float os = __shfl_xor_sync(0xffffffffu, s, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    argmax_combine(s, idx, os, oi);
  }
  best_s = s;
  best_i = idx;
}

// highest unused guide index (< A), -1 if none; and mark every guide token used.  One warp.
__device__ __forceinline__ int warp_consume_all_unused(unsigned int* used, int A, int lane) {
  int hi = -1;
  const int words = (A + 31) / 32;
  for (int w = lane; w < words; w += 32) {
    unsigned int valid = (w == words - 1 && (A & 31)) ? ((1u << (A & 31)) - 1u) : 0xffffffffu;
    unsigned int freeb = ~used[w] & valid;
    if (freeb) hi = max(hi, w * 32 + 31 - __clz(freeb));
    used[w] |= valid;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  return hi;
}

__global__ void __launch_bounds__(K1_THREADS, 1) k1_sim_blend_kernel(const K1Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~static_cast<uintptr_t>(1023));
  uint8_t* a_hi = stage;
  uint8_t* a_lo = stage + MAX_TILES * A_TILE_BYTES;
  uint8_t* b_hi = stage + 2 * MAX_TILES * A_TILE_BYTES;
  uint8_t* b_lo = b_hi + B_TILE_BYTES;
  float* pt = reinterpret_cast<float*>(stage);  // aliases the staging area after the GEMM
  K1Smem& sm = *reinterpret_cast<K1Smem*>(stage + STAGE_BYTES + 8192);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int T = a.T, A = a.A, D = a.D;
  const int n_tiles = (A + 127) / 128;
  const int b_idx = blockIdx.x;
  const float* text = a.text + static_cast<size_t>(b_idx) * T * D;
  const float* guide = a.guide + (a.guide_batch == 1 ? 0 : static_cast<size_t>(b_idx) * A * D);
  const uint32_t tmem_cols = 256;  // 3 x 80 = 240 -> next power of two

  if (tid == 0) {
    mbar_init(&sm.mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&sm.tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_slot;

  // ------------------------------------------------------------------ 1. GEMM
  const int a_items_total = A * 8;  // float4 items per chunk (row, 16 B column)
  const int b_items_total = T * 8;
  float4 ra[A_ITEMS], rb[B_ITEMS];
  float ssa[A_ITEMS], ssb[B_ITEMS];
#pragma unroll
  for (int j = 0; j < A_ITEMS; ++j) ssa[j] = 0.f;
#pragma unroll
  for (int j = 0; j < B_ITEMS; ++j) ssb[j] = 0.f;

  auto load_chunk = [&](int kc) {
#pragma unroll
    for (int j = 0; j < A_ITEMS; ++j) {
      const int f = tid + j * K1_THREADS;
      if (f < a_items_total)
        ra[j] = __ldg(reinterpret_cast<const float4*>(guide + static_cast<size_t>(f >> 3) * D + kc * KC) + (f & 7));
      else
        ra[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < B_ITEMS; ++j) {
      const int f = tid + j * K1_THREADS;
      if (f < b_items_total)
        rb[j] = __ldg(reinterpret_cast<const float4*>(text + static_cast<size_t>(f >> 3) * D + kc * KC) + (f & 7));
      else
        rb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto split_store = [&](const float4& v, uint8_t* hi_base, uint8_t* lo_base, int f) {
    const uint32_t row = f >> 3, c16 = f & 7;
    const uint32_t tile = row >> 7, r = row & 127;
    const uint32_t off = tile * A_TILE_BYTES + sw128_offset(r, c16);
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    l.x = v.x - h.x;
    l.y = v.y - h.y;
    l.z = v.z - h.z;
    l.w = v.w - h.w;
    *reinterpret_cast<float4*>(hi_base + off) = h;
    *reinterpret_cast<float4*>(lo_base + off) = l;
  };

  const int num_kc = D / KC;
  const uint32_t idesc = umma_idesc(UMMA_TF32, 128, NPAD, 0, 0);
  load_chunk(0);
  for (int kc = 0; kc < num_kc; ++kc) {
    if (kc > 0) mbar_wait(&sm.mma_bar, (kc - 1) & 1);  // MMAs of the previous chunk retired
#pragma unroll
    for (int j = 0; j < A_ITEMS; ++j) {
      const int f = tid + j * K1_THREADS;
      if (f < n_tiles * 128 * 8) {  // pad rows are written as zeros
        split_store(ra[j], a_hi, a_lo, f);
        ssa[j] += ra[j].x * ra[j].x + ra[j].y * ra[j].y + ra[j].z * ra[j].z + ra[j].w * ra[j].w;
      }
    }
#pragma unroll
    for (int j = 0; j < B_ITEMS; ++j) {
      const int f = tid + j * K1_THREADS;
      if (f < NPAD * 8) {
        split_store(rb[j], b_hi, b_lo, f);  // tile index is always 0 for f < 128*8
        ssb[j] += rb[j].x * rb[j].x + rb[j].y * rb[j].y + rb[j].z * rb[j].z + rb[j].w * rb[j].w;
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      for (int t = 0; t < n_tiles; ++t) {
        const uint64_t ah = umma_desc_sw128(smem_u32(a_hi + t * A_TILE_BYTES), 16, 1024);
        const uint64_t al = umma_desc_sw128(smem_u32(a_lo + t * A_TILE_BYTES), 16, 1024);
        const uint64_t bh = umma_desc_sw128(smem_u32(b_hi), 16, 1024);
        const uint64_t bl = umma_desc_sw128(smem_u32(b_lo), 16, 1024);
        const uint32_t d = tmem_base + t * NPAD;
#pragma unroll
        for (int ks = 0; ks < KC / 8; ++ks) {  // UMMA_K = 8 for tf32 = 32 B = +2 in the desc
          mma_tf32_ss(d, al + 2 * ks, bh + 2 * ks, idesc, (kc | ks) != 0);  // small terms first
          mma_tf32_ss(d, ah + 2 * ks, bl + 2 * ks, idesc, 1);
          mma_tf32_ss(d, ah + 2 * ks, bh + 2 * ks, idesc, 1);
        }
      }
      tc_commit(&sm.mma_bar);
    }
    if (kc + 1 < num_kc) load_chunk(kc + 1);  // overlaps the MMAs just issued
  }
  // L2 norms: the 8 lanes that share a row sit in one aligned group of 8 lanes
#pragma unroll
  for (int j = 0; j < A_ITEMS; ++j) {
    float s = ssa[j];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    const int f = tid + j * K1_THREADS;
    if ((f & 7) == 0 && (f >> 3) < MAX_TILES * 128) sm.inv_norm_a[f >> 3] = 1.0f / sqrtf(s);
  }
#pragma unroll
  for (int j = 0; j < B_ITEMS; ++j) {
    float s = ssb[j];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    const int f = tid + j * K1_THREADS;
    if ((f & 7) == 0 && (f >> 3) < NPAD) sm.inv_norm_b[f >> 3] = 1.0f / sqrtf(s);
  }
  mbar_wait(&sm.mma_bar, (num_kc - 1) & 1);
  tc_fence_after();
  __syncthreads();  // norms visible; staging area free for P^T

  // ------------------------------------------------------------------ 2. softmax per lane
  {
    const int tile = warp >> 2, quarter = warp & 3;
    if (tile < n_tiles) {
      const int i = tile * 128 + quarter * 32 + lane;
      float l[NPAD];
#pragma unroll
      for (int c = 0; c < NPAD; c += 16) {
        uint32_t v[16];
        tmem_ld_x16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + tile * NPAD + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q) l[c + q] = __uint_as_float(v[q]);
      }
      if (i < A) {
        const float ia = sm.inv_norm_a[i];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < NPAD; ++j) {
          // 100 * cos(guide_i, text_j)   (guidance.py:43-50)
          l[j] = (j < T) ? 100.0f * (l[j] * ia * sm.inv_norm_b[j]) : -INFINITY;
          mx = fmaxf(mx, l[j]);
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < NPAD; ++j) {
          l[j] = (j < T) ? expf(l[j] - mx) : 0.f;
          sum += l[j];
        }
        float* simrow = (a.sim && blockIdx.y == 0) ? a.sim + (static_cast<size_t>(b_idx) * A + i) * T : nullptr;
#pragma unroll
        for (int j = 0; j < NPAD; ++j) {
          if (j < T) {
            const float p = l[j] / sum;
            if (j >= 1) pt[(j - 1) * PT_STRIDE + i] = p;  // header column dropped (guidance.py:55)
            if (simrow) simrow[j] = p;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();

  // ------------------------------------------------------------------ 3..5 per parameter set
  const int p_begin = blockIdx.y * a.params_per_cta;
  const int p_end = min(a.n_params, p_begin + a.params_per_cta);
  const int ncol = T - 1;  // mapped rows 0..T-2 <-> text tokens 1..T-1
  int have_mode = -1, have_reuse = -1;
  for (int p = p_begin; p < p_end; ++p) {
    const fd_tween_params prm = a.params[p];
    const size_t bp = static_cast<size_t>(b_idx) * a.n_params + p;

    // ---------------- 3. mapping (recomputed only when mode / reuse change)
    if (prm.align_mode != have_mode || (prm.mapping_reuse != 0) != have_reuse) {
      have_mode = prm.align_mode;
      have_reuse = prm.mapping_reuse != 0;
      for (int r = tid; r < MAXT + 16; r += K1_THREADS) {
        sm.map_s[r] = 0.f;
        sm.map_idx[r] = 0;
        sm.assigned[r] = 0;
      }
      for (int w = tid; w < MAX_TILES * 128 / 32; w += K1_THREADS) sm.used[w] = 0u;
      __syncthreads();
      if (have_mode == FD_GUIDE_ORDER_DIRECT) {
        // guidance.py:60-69: map[r] = (r, P[r, r+1]) where both exist
        for (int r = tid; r < ncol && r < A; r += K1_THREADS) {
          sm.map_idx[r] = r;
          sm.map_s[r] = pt[r * PT_STRIDE + r];
        }
      } else if (have_reuse) {
        // guidance.py:57-59,70-84 with reuse: per column arg-max, ties -> lowest guide index;
        // a column whose maximum is exactly 0.0 keeps being overwritten (Q4) -> last guide token
        for (int r = warp; r < ncol; r += K1_WARPS) {
          float s;
          int i;
          warp_col_argmax(pt + r * PT_STRIDE, A, nullptr, lane, s, i);
          if (lane == 0) {
            if (s > 0.f) {
              sm.map_s[r] = s;
              sm.map_idx[r] = i;
            } else {
              sm.map_s[r] = 0.f;
              sm.map_idx[r] = A - 1;
            }
          }
        }
      } else if (have_mode == FD_GUIDE_ORDER_TEXT) {
        // text order, no reuse: columns in order, each takes its best still-unused guide token
        if (warp == 0) {
          for (int r = 0; r < ncol; ++r) {
            float s;
            int i;
            warp_col_argmax(pt + r * PT_STRIDE, A, sm.used, lane, s, i);
            __syncwarp();
            if (i < 0) break;  // every guide token consumed: the rest stay (0, 0)
            if (s > 0.f) {
              if (lane == 0) {
                sm.map_s[r] = s;
                sm.map_idx[r] = i;
                sm.used[i >> 5] |= 1u << (i & 31);
              }
            } else {
              // Q4: a zero similarity never marks the row assigned, so every remaining tuple of
              // this column is taken in turn and each consumes its guide token
              const int hi = warp_consume_all_unused(sm.used, A, lane);
              if (lane == 0) {
                sm.map_s[r] = 0.f;
                sm.map_idx[r] = hi;
              }
            }
            __syncwarp();
          }
        }
      } else {
        // alignment order, no reuse: global greedy on (-s, text, guide)
        for (int r = warp; r < ncol; r += K1_WARPS) {
          float s;
          int i;
          warp_col_argmax(pt + r * PT_STRIDE, A, nullptr, lane, s, i);
          if (lane == 0) {
            sm.col_s[r] = s;
            sm.col_i[r] = i;
          }
        }
        __syncthreads();
        for (int it = 0; it < ncol; ++it) {
          if (warp == 0) {
            float s = -1.f;
            int r_best = -1;
            for (int r = lane; r < ncol; r += 32) {
              if (sm.assigned[r] || sm.col_i[r] < 0) continue;
              const float v = sm.col_s[r];
              if (r_best < 0 || v > s) {
                s = v;
                r_best = r;
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float os = __shfl_xor_sync(0xffffffffu, s, o);
              const int orr = __shfl_xor_sync(0xffffffffu, r_best, o);
              argmax_combine(s, r_best, os, orr);
            }
            if (r_best >= 0 && s > 0.f) {
              if (lane == 0) {
                const int i = sm.col_i[r_best];
                sm.map_s[r_best] = s;
                sm.map_idx[r_best] = i;
                sm.assigned[r_best] = 1;
                sm.used[i >> 5] |= 1u << (i & 31);
                sm.pick_r = r_best;
                sm.pick_i = i;
                sm.flag = 1;
              }
            } else {
              // only zero similarities remain: the first unassigned column swallows every unused
              // guide token (Q4), later columns find nothing
              if (r_best >= 0) {
                int r_first = 1 << 30;
                for (int r = lane; r < ncol; r += 32)
                  if (!sm.assigned[r] && sm.col_i[r] >= 0) r_first = min(r_first, r);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) r_first = min(r_first, __shfl_xor_sync(0xffffffffu, r_first, o));
                const int hi = warp_consume_all_unused(sm.used, A, lane);
                if (lane == 0 && hi >= 0) {
                  sm.map_s[r_first] = 0.f;
                  sm.map_idx[r_first] = hi;
                }
              }
              if (lane == 0) sm.flag = 0;
            }
          }
          __syncthreads();
          if (!sm.flag) break;
          const int pi = sm.pick_i;
          for (int r = warp; r < ncol; r += K1_WARPS) {
            if (sm.assigned[r] || sm.col_i[r] != pi) continue;  // warp-uniform
            float s;
            int i;
            warp_col_argmax(pt + r * PT_STRIDE, A, sm.used, lane, s, i);
            if (lane == 0) {
              sm.col_s[r] = s;
              sm.col_i[r] = i;
            }
          }
          __syncthreads();
        }
      }
      __syncthreads();
    }

    // ---------------- 4. weights: one warp, token r = lane + 32 k
    if (warp == 0) {
      constexpr int KR = 3;  // 96 >= MAXT tokens
      double s_d[KR];
      float w[KR];
      for (int k = 0; k < KR; ++k) {
        const int r = lane + 32 * k;
        s_d[k] = (r < T) ? static_cast<double>(sm.map_s[r]) : 0.0;
        w[k] = (r < T) ? a.lin_w[static_cast<size_t>(p) * T + r] : 0.f;
      }
      // avg_similarity = mapped_tokens[:, 1].mean()  (guidance.py:219): numpy float64 pairwise sum
      double avg;
      {
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        double res = 0.0;
        if (lane == 0) {
          if (T < 8) {
            for (int r = 0; r < T; ++r) res += static_cast<double>(sm.map_s[r]);
          } else {
            for (int q = 0; q < 8; ++q) acc[q] = static_cast<double>(sm.map_s[q]);
            int r = 8;
            for (; r < T - (T % 8); r += 8)
              for (int q = 0; q < 8; ++q) acc[q] += static_cast<double>(sm.map_s[r + q]);
            res = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
            for (; r < T; ++r) res += static_cast<double>(sm.map_s[r]);
          }
          res = res / static_cast<double>(T);
        }
        avg = __shfl_sync(0xffffffffu, res, 0);
      }
      int status = FD_BLEND_OK;
      auto blend = [&](float (&aw)[KR], const float (&bw)[KR]) {
        // _blend_weights (guidance.py:175-193): global-sign switch
        float amax = -INFINITY, bmax = -INFINITY;
        for (int k = 0; k < KR; ++k)
          if (lane + 32 * k < T) {
            amax = fmaxf(amax, aw[k]);
            bmax = fmaxf(bmax, bw[k]);
          }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
          bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
        }
        for (int k = 0; k < KR; ++k) {
          if (amax >= 0.f) {
            if (bmax >= 0.f) aw[k] = fmaxf(aw[k], bw[k]);
            else aw[k] = __fadd_rn(aw[k], bw[k]);
          } else {
            aw[k] = fminf(aw[k], bw[k]);
          }
        }
      };

      if (prm.clustered != 0.0) {
        // _clustered_guidance (guidance.py:135-172)
        unsigned int peak_mask[KR];
        for (int k = 0; k < KR; ++k) {
          const int r = lane + 32 * k;
          bool pk = false;
          if (r >= 1 && r <= T - 2) {
            const double s = s_d[k];
            const double sl = static_cast<double>(sm.map_s[r - 1]);
            const double sr = static_cast<double>(sm.map_s[r + 1]);
            pk = !(s < avg) && (sl <= s) && (s >= sr);
          }
          peak_mask[k] = __ballot_sync(0xffffffffu, pk);
        }
        const bool any_peak = (peak_mask[0] | peak_mask[1] | peak_mask[2]) != 0u;
        if (any_peak) {
          // adjacent peaks => valley lands on the next peak => d = 0 => ZeroDivisionError (Q6)
          bool adj = false;
          for (int k = 0; k < KR; ++k) {
            if (peak_mask[k] & (peak_mask[k] >> 1)) adj = true;
            if (k + 1 < KR && (peak_mask[k] >> 31) && (peak_mask[k + 1] & 1u)) adj = true;
          }
          if (adj) status = FD_BLEND_ZERO_DIVISION;
          auto prev_peak = [&](int r) {  // largest peak index <= r, -1 if none
            for (int k = r >> 5; k >= 0; --k) {
              unsigned int m = peak_mask[k];
              if (k == (r >> 5)) m &= (r & 31) == 31 ? 0xffffffffu : ((1u << ((r & 31) + 1)) - 1u);
              if (m) return k * 32 + 31 - __clz(m);
            }
            return -1;
          };
          auto next_peak = [&](int r) {  // smallest peak index >= r, -1 if none
            for (int k = r >> 5; k < KR; ++k) {
              unsigned int m = peak_mask[k];
              if (k == (r >> 5)) m &= ~((1u << (r & 31)) - 1u);
              if (m) return k * 32 + __ffs(m) - 1;
            }
            return -1;
          };
          float cw[KR];
          for (int k = 0; k < KR; ++k) {
            const int r = lane + 32 * k;
            float c = 1.0f;
            if (r < T && !adj) {
              const int pl = prev_peak(r), pr = next_peak(r);
              if (r == 0) {
                c = 0.0f;  // weights[0] -= slope (guidance.py:116-118); bl[0] is always 0
              } else if (pl == r) {
                c = 1.0f;
              } else if (pl >= 0) {
                const int vr = (pr >= 0) ? pl + (pr - pl + 1) / 2 : T - 1;  // p1 + ceil(d / 2)
                if (r <= vr) {
                  const double g = 1.0 / static_cast<double>(vr - pl);       // traverse_right
                  c = __fsub_rn(1.0f, static_cast<float>(g * static_cast<double>(r - pl)));
                } else {
                  const double g = 1.0 / static_cast<double>(pr - vr);       // traverse_left
                  c = __fsub_rn(1.0f, static_cast<float>(g * static_cast<double>(pr - r)));
                }
              } else {
                const double g = 1.0 / static_cast<double>(pr);              // left of first peak
                c = __fsub_rn(1.0f, static_cast<float>(g * static_cast<double>(pr - r)));
              }
            }
            cw[k] = __fmul_rn(c, static_cast<float>(prm.clustered));
          }
          blend(w, cw);
        }
      }
      if (prm.threshold_mult != 0.0) {
        // guidance.py:241-246
        float th[KR];
        for (int k = 0; k < KR; ++k) th[k] = (s_d[k] < prm.threshold_floor) ? 0.f : static_cast<float>(prm.threshold_mult);
        blend(w, th);
      }
      if (prm.header_max < 1.0 && lane == 0) {
        // guidance.py:249-254
        const double hw = static_cast<double>(w[0]);
        w[0] = (hw >= 0.0) ? static_cast<float>(fmin(hw, prm.header_max)) : static_cast<float>(fmax(hw, -prm.header_max));
      }
      for (int k = 0; k < KR; ++k) {
        const int r = lane + 32 * k;
        if (r < T) {
          // guidance.py:259-271
          const double iw = fmin(static_cast<double>(w[k]), prm.max_guidance);
          const double sd = 1.0 - s_d[k];
          int sel = 2;
          if (iw == 0.0) sel = 0;
          else if (fabs(iw) >= sd) sel = 1;
          sm.iw[r] = static_cast<float>(iw);
          sm.sel[r] = sel;
          if (a.weights) a.weights[bp * T + r] = w[k];
          if (a.map_s) a.map_s[bp * T + r] = sm.map_s[r];
          if (a.map_idx) a.map_idx[bp * T + r] = sm.map_idx[r];
        }
      }
      if (lane == 0 && a.status) a.status[bp] = status;
    }
    __syncthreads();

    // ---------------- 5. select / lerp, 2 rows of 192 float4 per pass
    {
      const int d4 = D / 4;
      float* outp = a.out + bp * T * D;
      for (int idx = tid; idx < T * d4; idx += K1_THREADS) {
        const int r = idx / d4, c = idx - r * d4;
        const int sel = sm.sel[r];
        const float4 bv = __ldg(reinterpret_cast<const float4*>(text + static_cast<size_t>(r) * D) + c);
        float4 o = bv;
        if (sel != 0) {
          const float4 av = __ldg(reinterpret_cast<const float4*>(guide + static_cast<size_t>(sm.map_idx[r]) * D) + c);
          if (sel == 1) {
            o = av;
          } else {
            const float w = sm.iw[r];
            // base + (alt - base) * iw, every op rounded separately like the torch expression
            o.x = __fadd_rn(bv.x, __fmul_rn(__fsub_rn(av.x, bv.x), w));
            o.y = __fadd_rn(bv.y, __fmul_rn(__fsub_rn(av.y, bv.y), w));
            o.z = __fadd_rn(bv.z, __fmul_rn(__fsub_rn(av.z, bv.z), w));
            o.w = __fadd_rn(bv.w, __fmul_rn(__fsub_rn(av.w, bv.w), w));
          }
        }
        __stcs(reinterpret_cast<float4*>(outp + static_cast<size_t>(r) * D) + c, o);
      }
    }
    __syncthreads();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace
}  // namespace fd

extern "C" int fd_sim_blend(const float* text_dev, const float* guide_dev, int n_text, int guide_batch, int T, int A,
                            int D, const fd_tween_params* params_dev, const float* linear_weights_dev, int n_params,
                            float* out_dev, float* map_s_dev, int32_t* map_idx_dev, float* weights_dev,
                            int32_t* status_dev, float* sim_dev, void* stream) {
  using namespace fd;
  FD_REQUIRE(text_dev && guide_dev && params_dev && linear_weights_dev && out_dev, "fd_sim_blend: NULL pointer");
  FD_REQUIRE(n_text > 0 && n_params > 0, "fd_sim_blend: n_text=%d n_params=%d must be positive", n_text, n_params);
  FD_REQUIRE(guide_batch == 1 || guide_batch == n_text, "fd_sim_blend: guide_batch=%d must be 1 or n_text=%d",
             guide_batch, n_text);
  FD_REQUIRE(T >= 2 && T <= MAXT, "fd_sim_blend: T=%d outside [2, %d]", T, MAXT);
  FD_REQUIRE(A >= 1 && A <= MAX_TILES * 128, "fd_sim_blend: A=%d outside [1, %d]", A, MAX_TILES * 128);
  FD_REQUIRE(D >= KC && D % KC == 0, "fd_sim_blend: D=%d must be a positive multiple of %d", D, KC);
  FD_REQUIRE(reinterpret_cast<uintptr_t>(text_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(guide_dev) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(out_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(params_dev) % 8 == 0,
             "fd_sim_blend: text/guide/out must be 16-byte aligned");
  int rc = check_device();
  if (rc != FD_OK) return rc;

  K1Args a;
  a.text = text_dev;
  a.guide = guide_dev;
  a.n_text = n_text;
  a.guide_batch = guide_batch;
  a.T = T;
  a.A = A;
  a.D = D;
  a.params = params_dev;
  a.lin_w = linear_weights_dev;
  a.n_params = n_params;
  a.out = out_dev;
  a.map_s = map_s_dev;
  a.map_idx = map_idx_dev;
  a.weights = weights_dev;
  a.status = status_dev;
  a.sim = sim_dev;
  // enough CTAs to fill the machine: split the parameter sets of one prompt over several CTAs
  // (each redoes the cheap similarity GEMM) when there are few prompts
  const int sms = sm_count();
  if (sms <= 0) return set_error(FD_ERR_CUDA, "fd_sim_blend: cannot query SM count");
  int chunks = 1;
  if (n_text < sms) chunks = (sms + n_text - 1) / n_text;
  if (chunks > n_params) chunks = n_params;
  a.params_per_cta = (n_params + chunks - 1) / chunks;
  chunks = (n_params + a.params_per_cta - 1) / a.params_per_cta;
  FD_REQUIRE(chunks <= 65535, "fd_sim_blend: too many parameter chunks");
  FD_CUDA_OK(cudaFuncSetAttribute(k1_sim_blend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BYTES));
  dim3 grid(n_text, chunks);
  k1_sim_blend_kernel<<<grid, K1_THREADS, K1_SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(a);
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}
