// K12 -- exact-fp32 attention core for the CLIP towers' short sequences (257 vision tokens, 77 text tokens):
//   out[b, t, h, :] = softmax_j( scale <q[b,t,h,:], k[b,j,h,:]> (+ causal mask) ) . v[b, :, h, :]
// Replaces the fp32 attention inside transformers' CLIPAttention that `encode/clip.py:57-65, 86-100` reach through
// `clip.text_model` / `clip.vision_model`.  Everything around it (q / k / v / out_proj) is K11; the towers are
// fp32 in the reference, so this stays plain fp32 FMAs (CUDA cores): 2 x 2 x H T^2 d = 0.27 GFLOP per vision
// layer is nothing, the job is latency -- torch's fp32 SDPA needs 30 us for [1, 16, 257, 64] on a B200.
// One CTA per (batch x head, block of 32 queries): K^T and V of the head in shared memory, each of the 8 warps owns 4
// query rows and shares every K / V shared-memory load between them; scores and probabilities live in registers
// (key j = lane + 32 i), the probabilities go through a per-warp shared-memory strip for the P.V product.
#include "fd_common.cuh"

namespace fd {
namespace {

constexpr int AT_THREADS = 256;
constexpr int AT_WARPS = AT_THREADS / 32;
constexpr int AT_ROWS = 4;                    // query rows per warp
constexpr int AT_QB = AT_WARPS * AT_ROWS;     // 32 queries per CTA

struct AtArgs {
  const float* q;
  const float* k;
  const float* v;
  float* out;
  int64_t row_stride;   // floats between consecutive tokens of q / k / v
  int B, T, H, d, causal;
  float scale;
};

// NI = keys per lane (32 NI >= T), NM = output columns per lane (32 NM >= d), DFULL: d == 32 NM (no column guards)
template <int NI, int NM, bool DFULL>
__global__ void __launch_bounds__(AT_THREADS, 1) k12_attn_f32_kernel(const AtArgs a) {
  extern __shared__ float at_smem[];
  const int T = a.T, d = a.d;
  constexpr int Tp = 32 * NI + 1;             // odd pitch (conflict-free transposed stores) covering every lane's keys
  float* Kt = at_smem;                        // [d][Tp]
  float* Vs = Kt + static_cast<size_t>(d) * Tp;          // [T][d]
  float* Qs = Vs + static_cast<size_t>(T) * d;           // [AT_WARPS][d][AT_ROWS]
  float* Ps = Qs + AT_WARPS * d * AT_ROWS;               // [AT_WARPS][32 NI][AT_ROWS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y, b = bh / a.H, h = bh - b * a.H;
  const int q0 = blockIdx.x * AT_QB;
  const float* kb = a.k + (static_cast<size_t>(b) * T) * a.row_stride + h * d;
  const float* vb = a.v + (static_cast<size_t>(b) * T) * a.row_stride + h * d;
  const float* qb = a.q + (static_cast<size_t>(b) * T) * a.row_stride + h * d;
  // keys this CTA needs: all of them, or (causal) only those up to its last query
  const int t_keys = a.causal ? min(T, q0 + AT_QB) : T;
  const int d4 = d >> 2;
  // K and V of the head: batches of 4 items (8 x 16-byte loads) per thread in flight before the first shared-memory store
  {
    constexpr int U = 4;
    const int n_items = t_keys * d4;
    for (int base = tid; base < n_items; base += AT_THREADS * U) {
      float4 kv[U], vv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = base + u * AT_THREADS;
        if (idx < n_items) {
          const int j = idx / d4, c = (idx - j * d4) << 2;
          kv[u] = __ldg(reinterpret_cast<const float4*>(kb + static_cast<size_t>(j) * a.row_stride + c));
          vv[u] = __ldg(reinterpret_cast<const float4*>(vb + static_cast<size_t>(j) * a.row_stride + c));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = base + u * AT_THREADS;
        if (idx < n_items) {
          const int j = idx / d4, c = (idx - j * d4) << 2;
          Kt[(c + 0) * Tp + j] = kv[u].x;
          Kt[(c + 1) * Tp + j] = kv[u].y;
          Kt[(c + 2) * Tp + j] = kv[u].z;
          Kt[(c + 3) * Tp + j] = kv[u].w;
          *reinterpret_cast<float4*>(Vs + static_cast<size_t>(j) * d + c) = vv[u];
        }
      }
    }
  }
  // keys [t_keys, 32 NI) do not exist (or are masked for every query of this CTA): zero columns, so the score loop
  // runs without guards (their scores are dropped by the softmax mask)
  for (int idx = tid; idx < d * (32 * NI - t_keys); idx += AT_THREADS) {
    const int dd = idx / (32 * NI - t_keys), j = t_keys + idx - dd * (32 * NI - t_keys);
    Kt[dd * Tp + j] = 0.f;
  }
  // this warp's 4 query rows, interleaved [dd][row] so one float4 load serves the 4 rows
  float* qw = Qs + warp * d * AT_ROWS;
  const int row0 = q0 + warp * AT_ROWS;
  for (int idx = lane; idx < d * AT_ROWS; idx += 32) {
    const int r = idx / d, dd = idx - r * d;
    const int t = row0 + r;
    qw[dd * AT_ROWS + r] = t < T ? __ldg(qb + static_cast<size_t>(t) * a.row_stride + dd) * a.scale : 0.f;
  }
  __syncthreads();
  if (row0 >= T) return;

  // ---- scores: s[r][i] = <q_r, k_(lane + 32 i)>
  float s[AT_ROWS][NI];
#pragma unroll
  for (int r = 0; r < AT_ROWS; ++r)
#pragma unroll
    for (int i = 0; i < NI; ++i) s[r][i] = 0.f;
#pragma unroll 4
  for (int dd = 0; dd < d; ++dd) {
    const float4 qv = *reinterpret_cast<const float4*>(qw + dd * AT_ROWS);
    const float* kr = Kt + dd * Tp + lane;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const float kv = kr[32 * i];
      s[0][i] = fmaf(qv.x, kv, s[0][i]);
      s[1][i] = fmaf(qv.y, kv, s[1][i]);
      s[2][i] = fmaf(qv.z, kv, s[2][i]);
      s[3][i] = fmaf(qv.w, kv, s[3][i]);
    }
  }
  // ---- softmax per row (keys beyond T, or beyond the query when causal, are masked)
  float* pw = Ps + warp * (32 * NI) * AT_ROWS;
#pragma unroll
  for (int r = 0; r < AT_ROWS; ++r) {
    const int t = row0 + r;
    const int lim = a.causal ? min(t + 1, T) : T;   // keys j < lim
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < NI; ++i)
      if (lane + 32 * i < lim) mx = fmaxf(mx, s[r][i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const float e = (lane + 32 * i < lim) ? __expf(s[r][i] - mx) : 0.f;
      s[r][i] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int i = 0; i < NI; ++i) pw[(lane + 32 * i) * AT_ROWS + r] = s[r][i] * inv;
  }
  __syncwarp();
  // ---- out[r][lane + 32 m] = sum_j p[r][j] v[j][lane + 32 m]   (d <= 128: up to 4 columns per lane)
  float o[AT_ROWS][NM];
#pragma unroll
  for (int r = 0; r < AT_ROWS; ++r)
#pragma unroll
    for (int m = 0; m < NM; ++m) o[r][m] = 0.f;
  const int jmax = a.causal ? min(T, row0 + AT_ROWS) : T;
#pragma unroll 2
  for (int j = 0; j < jmax; ++j) {
    const float4 pv = *reinterpret_cast<const float4*>(pw + j * AT_ROWS);
    const float* vr = Vs + static_cast<size_t>(j) * d + lane;
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      if (DFULL || lane + 32 * m < d) {
        const float vv = vr[32 * m];
        o[0][m] = fmaf(pv.x, vv, o[0][m]);
        o[1][m] = fmaf(pv.y, vv, o[1][m]);
        o[2][m] = fmaf(pv.z, vv, o[2][m]);
        o[3][m] = fmaf(pv.w, vv, o[3][m]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < AT_ROWS; ++r) {
    const int t = row0 + r;
    if (t < T) {
      float* orow = a.out + (static_cast<size_t>(b) * T + t) * (static_cast<size_t>(a.H) * d) + h * d;
#pragma unroll
      for (int m = 0; m < NM; ++m)
        if (DFULL || lane + 32 * m < d) orow[lane + 32 * m] = o[r][m];
    }
  }
}

}  // namespace
}  // namespace fd

extern "C" int fd_attention_f32(const float* q_dev, const float* k_dev, const float* v_dev, int64_t row_stride,
                                float* out_dev, int B, int T, int H, int d, float scale, int causal, void* stream) {
  using namespace fd;
  FD_REQUIRE(q_dev && k_dev && v_dev && out_dev, "fd_attention_f32: NULL pointer");
  FD_REQUIRE(B > 0 && T > 0 && H > 0 && d > 0, "fd_attention_f32: non-positive shape");
  FD_REQUIRE(d % 4 == 0 && d <= 128, "fd_attention_f32: d_head=%d must be a multiple of 4, at most 128", d);
  FD_REQUIRE(T <= 288, "fd_attention_f32: T=%d exceeds 288 (the CLIP towers need 77 and 257)", T);
  FD_REQUIRE(row_stride >= static_cast<int64_t>(H) * d && row_stride % 4 == 0,
             "fd_attention_f32: row_stride=%lld must cover H * d and be a multiple of 4", (long long)row_stride);
  FD_REQUIRE(reinterpret_cast<uintptr_t>(q_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(k_dev) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(v_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(out_dev) % 16 == 0,
             "fd_attention_f32: pointers must be 16-byte aligned");
  FD_REQUIRE(static_cast<int64_t>(B) * H <= 65535, "fd_attention_f32: too many (batch, head) pairs");
  int rc = check_device();
  if (rc != FD_OK) return rc;
  const int ni = T <= 96 ? 3 : 9;
  const int Tp = 32 * ni + 1;
  const size_t smem = (static_cast<size_t>(d) * Tp + static_cast<size_t>(T) * d + AT_WARPS * d * AT_ROWS +
                       static_cast<size_t>(AT_WARPS) * 32 * ni * AT_ROWS) * sizeof(float);
  FD_REQUIRE(smem <= 227 * 1024, "fd_attention_f32: T=%d, d=%d need %zu bytes of shared memory", T, d, smem);
  AtArgs a;
  a.q = q_dev;
  a.k = k_dev;
  a.v = v_dev;
  a.out = out_dev;
  a.row_stride = row_stride;
  a.B = B;
  a.T = T;
  a.H = H;
  a.d = d;
  a.causal = causal ? 1 : 0;
  a.scale = scale;
  dim3 grid((T + AT_QB - 1) / AT_QB, B * H);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nm = (d + 31) / 32;
  const bool dfull = d % 32 == 0;
#define FD_AT_LAUNCH(NI_, NM_, DF_)                                                                                         \
  do {                                                                                                                      \
    FD_CUDA_OK(cudaFuncSetAttribute(k12_attn_f32_kernel<NI_, NM_, DF_>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                    static_cast<int>(smem)));                                                               \
    k12_attn_f32_kernel<NI_, NM_, DF_><<<grid, AT_THREADS, smem, st>>>(a);                                                  \
  } while (0)
#define FD_AT_NM(NI_)                                                          \
  do {                                                                         \
    if (nm == 1) { if (dfull) FD_AT_LAUNCH(NI_, 1, true); else FD_AT_LAUNCH(NI_, 1, false); }      \
    else if (nm == 2) { if (dfull) FD_AT_LAUNCH(NI_, 2, true); else FD_AT_LAUNCH(NI_, 2, false); } \
    else if (nm == 3) { if (dfull) FD_AT_LAUNCH(NI_, 3, true); else FD_AT_LAUNCH(NI_, 3, false); } \
    else { if (dfull) FD_AT_LAUNCH(NI_, 4, true); else FD_AT_LAUNCH(NI_, 4, false); }              \
  } while (0)
  if (ni == 3) FD_AT_NM(3);
  else FD_AT_NM(9);
#undef FD_AT_NM
#undef FD_AT_LAUNCH
  FD_CUDA_OK(cudaGetLastError());
  return FD_OK;
}
