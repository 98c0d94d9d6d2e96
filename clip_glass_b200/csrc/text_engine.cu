// img2txt fitness path (BASELINE config 5): GPT-2 greedy decode with a KV cache and the CLIP text
// tower, behind the C ABI in include/clipglass_b200.h (glass_text_*).
//
//   models.py:45-62       GPT2.generate: cat(latent tokens, init tokens) -> sample_sequence (30 steps, top-1)
//   gpt2/model.py:45-175  12 x (LN -> c_attn -> masked attention over the cache -> c_proj -> +res -> LN -> c_fc -> GELU
//                         -> c_proj -> +res) -> ln_f -> tied LM head
//   generator.py:53-59    clip tokens -> CLIP.encode_text (clip/model.py:292-320) -> cosine vs cached image features
//
// Integer output must equal the reference's, which runs GPT-2 in fp32: every GPT-2 GEMM runs on the tensor cores as a
// SPLIT-fp16 product.  An fp32 value x is carried as hi = fp16(x) and lo = fp16((x - hi) * 2^11) (22 significant bits);
// x.w ~= hi_x.hi_w + (lo_x.hi_w + hi_x.lo_w) * 2^-11 with the two groups accumulated in fp32 in separate TMEM columns
// (so the correction terms stay in fp16's normal range) and combined in the epilogue; the dropped lo.lo term is
// 2^-22 relative.  LayerNorm, attention over the cache, the residual stream and the logits stay fp32.
// The CLIP text tower runs fp16 "as built" (clip/model.py:339-360) through the same GEMM kernel in plain mode with the
// reference's rounding points (linear output rounded to fp16 before the activation / residual add).
#include <cudaTypedefs.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/clipglass_b200.h"
#include "common.cuh"
#include "tcgen05.cuh"

using namespace glass;

namespace glass {
namespace {

thread_local std::string t_last_error;
int tfail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_last_error = buf;
  return code;
}
#define TCUDA_OK(expr)                                                                                      \
  do {                                                                                                      \
    cudaError_t err__ = (expr);                                                                             \
    if (err__ != cudaSuccess)                                                                               \
      return tfail(GLASS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
  } while (0)
#define TRC(expr)                      \
  do {                                 \
    int rc__ = (expr);                 \
    if (rc__ != GLASS_OK) return rc__; \
  } while (0)

constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;

__device__ __forceinline__ float warp_sum_t(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_t(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float rh_t(float v) { return __half2float(__float2half_rn(v)); }
__device__ __forceinline__ void split_store(float v, __half* hi, __half* lo) {
  const __half h = __float2half_rn(v);
  *hi = h;
  *lo = __float2half_rn((v - __half2float(h)) * kLoScale);
}

// ---------------------------------------------------------------------------
// GEMM on the tensor cores: C[M,N] = A[M,K] . W[N,K]^T (+ bias, activation, residual)
// ---------------------------------------------------------------------------
enum { kTActNone = 0, kTActGeluTanh = 1, kTActQuickGelu = 2 };
struct GemmParams {
  int M, N, K;
  int a_rows;                 // rows of the A box (128, or 64 when M <= 64: the upper half of the tile is never read back)
  int splits;                 // split-K: blockIdx.z handles K / splits; the raw fp32 partial sums go to
                              // out_f32 + z*M*N (no bias / activation / residual: the consumer reduces them)
  const float* bias;          // [N] or null
  int act;
  int round_fp16;             // plain mode: round acc + bias to fp16 before the activation (the reference's op boundary)
  const float* res_f32;       // [M,N] residual added after the activation (GPT-2 residual stream), or null
  const __half* res_f16;      // [M,N] fp16 residual (CLIP: x + f(x), added in fp32, rounded once), or null
  float* out_f32;             // [M,N] or null
  __half* out_hi;             // [M,N] fp16 (plain output, or the hi part of a split output), or null
  __half* out_lo;             // [M,N] lo part of a split output, or null
};

// kARows: rows of the A tile that are actually loaded (128, or 64 for the decode steps with M <= 64: the MMA still
// reads 128 rows, the upper half being whatever follows in shared memory; those accumulator rows are never stored).
// A decode GEMM is a latency-bound stream of small TMA boxes: the pipeline is as deep as shared memory allows
// (8 stages of 24 KB at kARows = 64, BN = 32 against a ~1.5 us TMA round trip).
template <int BN, bool kSplit, int kARows>
struct GemmCfg {
  static constexpr int kABytes = kARows * 64 * 2, kWBytes = BN * 64 * 2;
  static constexpr int kStageBytes = (kSplit ? 2 : 1) * (kABytes + kWBytes);
  // (full-height tiles, i.e. the prefill and the CLIP text tower with thousands of rows: four stages, so that two CTAs
  // fit on an SM; measured 4.3 ms against 5.1 ms for the text tower with eight)
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStagesCap = kARows == 64 ? 8 : 4;
  static constexpr int kStages = kStagesRaw > kStagesCap ? kStagesCap : kStagesRaw;
  static constexpr int kSlack = kARows == 64 ? 16 * 1024 : 0;       // what the M=128 MMA may read behind a half-height A tile
  static constexpr int kSmemBytes = kStages * kStageBytes + kSlack + 1024 + 256;
  static constexpr int kAccCols = (kSplit ? 2 : 1) * BN;
  static constexpr int kTmemCols = kAccCols < 32 ? 32 : kAccCols;
  static constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
};

template <int BN, bool kSplit, int kARows>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
               const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl,
               const GemmParams p) {
  using C = GemmCfg<BN, kSplit, kARows>;
  constexpr int kGemmStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kGemmStages * C::kStageBytes + C::kSlack);
  uint64_t* empty_bar = full_bar + kGemmStages;
  uint64_t* tmem_full = empty_bar + kGemmStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y;
  const int kiters = p.K / 64 / (p.splits > 1 ? p.splits : 1);
  const int kit0 = (p.splits > 1 ? (int)blockIdx.z : 0) * kiters;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_ah);
    prefetch_tmap(&map_wh);
    if (kSplit) { prefetch_tmap(&map_al); prefetch_tmap(&map_wl); }
    for (int s = 0; s < kGemmStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(C::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kit = 0; kit < kiters; ++kit) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * C::kStageBytes;
        mbar_expect_tx(&full_bar[stage], C::kStageBytes);
        const int kc = (kit0 + kit) * 64;
        tma_load_2d(&map_ah, sa, &full_bar[stage], kc, m_tile * 128);
        tma_load_2d(&map_wh, sa + C::kABytes, &full_bar[stage], kc, n_tile * BN);
        if (kSplit) {
          tma_load_2d(&map_al, sa + C::kABytes + C::kWBytes, &full_bar[stage], kc, m_tile * 128);
          tma_load_2d(&map_wl, sa + 2 * C::kABytes + C::kWBytes, &full_bar[stage], kc, n_tile * BN);
        }
        if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kit = 0; kit < kiters; ++kit) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
        const uint64_t dah = make_smem_desc<64>(sa);
        const uint64_t dwh = make_smem_desc<64>(sa + C::kABytes);
        const uint64_t dal = make_smem_desc<64>(sa + C::kABytes + C::kWBytes);
        const uint64_t dwl = make_smem_desc<64>(sa + 2 * C::kABytes + C::kWBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t o = (uint64_t)(k * 2);
          tc_mma_f16(tmem_base, dah + o, dwh + o, C::kIdesc, (uint32_t)(kit | k));
          if (kSplit) {
            tc_mma_f16(tmem_base + BN, dal + o, dwh + o, C::kIdesc, (uint32_t)(kit | k));
            tc_mma_f16(tmem_base + BN, dah + o, dwl + o, C::kIdesc, 1u);
          }
        }
        tc_commit(&empty_bar[stage]);
        if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
      }
      tc_commit(tmem_full);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int m = m_tile * 128 + row;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN / 16; ++c) {
      float v[16];
      tc_ld16(taddr + c * 16, v);
      if (kSplit) {
        float w[16];
        tc_ld16(taddr + BN + c * 16, w);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaf(w[j], kLoInv, v[j]);
      }
      if (m < p.M && p.splits > 1) {
        float* dst = p.out_f32 + (size_t)blockIdx.z * p.M * p.N + (size_t)m * p.N + n_tile * BN + c * 16;
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else if (m < p.M) {
        const int n0 = n_tile * BN + c * 16;
        const size_t o = (size_t)m * p.N + n0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float t = v[j] + (p.bias != nullptr ? __ldg(p.bias + n0 + j) : 0.f);
          if (p.round_fp16) t = rh_t(t);
          if (p.act == kTActGeluTanh) {
            t = 0.5f * t * (1.f + tanhf(0.7978845608028654f * (t + 0.044715f * t * t * t)));
          } else if (p.act == kTActQuickGelu) {
            t = t / (1.f + __expf(-1.702f * t));
            if (p.round_fp16) t = rh_t(t);
          }
          if (p.res_f32 != nullptr) t += p.res_f32[o + j];
          if (p.res_f16 != nullptr) t += __half2float(p.res_f16[o + j]);
          v[j] = t;
        }
        if (p.out_f32 != nullptr) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(p.out_f32 + o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (p.out_lo != nullptr) {
#pragma unroll
          for (int j = 0; j < 16; ++j) split_store(v[j], p.out_hi + o + j, p.out_lo + o + j);
        } else if (p.out_hi != nullptr) {
          uint4 w0, w1;
          __half2* h0 = reinterpret_cast<__half2*>(&w0);
          __half2* h1 = reinterpret_cast<__half2*>(&w1);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            h0[j] = f2h2_sat(v[2 * j], v[2 * j + 1]);
            h1[j] = f2h2_sat(v[8 + 2 * j], v[8 + 2 * j + 1]);
          }
          *reinterpret_cast<uint4*>(p.out_hi + o) = w0;
          *reinterpret_cast<uint4*>(p.out_hi + o + 8) = w1;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::kTmemCols));
  }
}

// ---------------------------------------------------------------------------
// GPT-2 helpers (fp32)
// ---------------------------------------------------------------------------
// h[b*Tn + t] = wte[token] + wpe[pos]  (gpt2/model.py:148-156); tokens int32 [P][Ttot], columns col0 .. col0+Tn-1
__global__ void gpt2_embed_kernel(const int* __restrict__ tokens, int Ttot, int col0, int Tn, const float* __restrict__ wte,
                                  const float* __restrict__ wpe, float* __restrict__ h, int E) {
  const int row = blockIdx.x;                  // b*Tn + t
  const int b = row / Tn, t = row - b * Tn;
  const int tok = tokens[(size_t)b * Ttot + col0 + t];
  const float4* we = reinterpret_cast<const float4*>(wte + (size_t)tok * E);
  const float4* pe = reinterpret_cast<const float4*>(wpe + (size_t)(col0 + t) * E);
  float4* o = reinterpret_cast<float4*>(h + (size_t)row * E);
  for (int i = threadIdx.x; i < E / 4; i += blockDim.x) {
    const float4 a = we[i], c = pe[i];
    o[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
  }
}

// TF-style LayerNorm (gpt2/model.py:16-29) of rows in_row(r) = r*row_mul + row_add of x -> split fp16 output row r.
// One 128-thread block per row (a decode step has only P rows: one warp per row left 8 blocks on 148 SMs, 7 us).
__global__ void __launch_bounds__(128) gpt2_layernorm_split_kernel(const float* __restrict__ x, int row_mul, int row_add,
                                                                   const float* __restrict__ w, const float* __restrict__ b,
                                                                   float eps, __half* __restrict__ hi, __half* __restrict__ lo,
                                                                   int rows, int E) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  if (r >= rows) return;
  const float* xr = x + (size_t)(r * row_mul + row_add) * E;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  float v[8];                                   // E <= 1024
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = threadIdx.x + k * 128;
    v[k] = i < E ? xr[i] : 0.f;
    s += v[k];
  }
  s = warp_sum_t(s);
  if (lane == 0) red[wp] = s;
  __syncthreads();
  const float u = (red[0] + red[1] + red[2] + red[3]) / (float)E;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = threadIdx.x + k * 128;
    const float d = i < E ? v[k] - u : 0.f;
    q += d * d;
  }
  q = warp_sum_t(q);
  if (lane == 0) red[4 + wp] = q;
  __syncthreads();
  const float inv = 1.f / sqrtf((red[4] + red[5] + red[6] + red[7]) / (float)E + eps);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = threadIdx.x + k * 128;
    if (i < E) split_store(w[i] * ((v[k] - u) * inv) + b[i], hi + (size_t)r * E + i, lo + (size_t)r * E + i);
  }
}

// Decode-step consumers of split-K partial sums.  x = sum_s part[s] + bias, reduced in the fixed order s = 0, 1, ...
// (deterministic: the arg-max of the logits must not depend on scheduling).
// (a) residual update + LayerNorm: h[r] += x[r]; then the TF-style LayerNorm of h[r] -> split fp16 (the LayerNorm that
//     follows attn.c_proj / mlp.c_proj in the block structure: ln_2, the next layer's ln_1, or ln_f)
__global__ void __launch_bounds__(128) gpt2_reduce_residual_ln_kernel(const float* __restrict__ part, int S, size_t part_stride,
                                                                      const float* __restrict__ bias, float* __restrict__ h,
                                                                      const float* __restrict__ w, const float* __restrict__ b,
                                                                      float eps, __half* __restrict__ hi, __half* __restrict__ lo,
                                                                      int E) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  float v[8];
  float s = 0.f;
  {
    // the S partial sums of an element are added in the fixed order z = 0, 1, ...; the loads of four splits x eight
    // elements are issued together (a serial load-add chain per element made this kernel 22 us: 144 dependent L2 hits)
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const float* pr = part + (size_t)r * E + threadIdx.x;
    for (int z = 0; z < S; z += 4) {
      float t[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          t[u][k] = (z + u < S && threadIdx.x + k * 128 < E) ? pr[(size_t)(z + u) * part_stride + k * 128] : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += t[u][k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = threadIdx.x + k * 128;
      v[k] = 0.f;
      if (i < E) {
        v[k] = h[(size_t)r * E + i] + (acc[k] + bias[i]);
        h[(size_t)r * E + i] = v[k];
      }
      s += v[k];
    }
  }
  s = warp_sum_t(s);
  if (lane == 0) red[wp] = s;
  __syncthreads();
  const float u = (red[0] + red[1] + red[2] + red[3]) / (float)E;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = threadIdx.x + k * 128;
    const float d = i < E ? v[k] - u : 0.f;
    q += d * d;
  }
  q = warp_sum_t(q);
  if (lane == 0) red[4 + wp] = q;
  __syncthreads();
  const float inv = 1.f / sqrtf((red[4] + red[5] + red[6] + red[7]) / (float)E + eps);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = threadIdx.x + k * 128;
    if (i < E) split_store(w[i] * ((v[k] - u) * inv) + b[i], hi + (size_t)r * E + i, lo + (size_t)r * E + i);
  }
}
// (b) c_fc: tanh-GELU(x) -> split fp16
__global__ void gpt2_reduce_gelu_split_kernel(const float* __restrict__ part, int S, size_t part_stride,
                                              const float* __restrict__ bias, __half* __restrict__ hi, __half* __restrict__ lo,
                                              int M, int N) {
  const size_t n = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int z = 0; z < S; ++z) acc += part[(size_t)z * part_stride + i];
    float t = acc + bias[i % N];
    t = 0.5f * t * (1.f + tanhf(0.7978845608028654f * (t + 0.044715f * t * t * t)));
    split_store(t, hi + i, lo + i);
  }
}
// (c) c_attn: qkv[b][3E] = x (plain fp32, read by the attention kernel)
__global__ void gpt2_reduce_bias_kernel(const float* __restrict__ part, int S, size_t part_stride, const float* __restrict__ bias,
                                        float* __restrict__ out, int M, int N) {
  const size_t n = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int z = 0; z < S; ++z) acc += part[(size_t)z * part_stride + i];
    out[i] = acc + bias[i % N];
  }
}

// Masked attention over the KV cache (gpt2/model.py:59-95).  One block per (candidate, head); head dim 64.
// qkv fp32 [P*Tn][3E]; cache K/V fp32 [P][H][Tmax][64] (this layer); the Tn new positions past .. past+Tn-1 are
// appended, then every new query attends to keys 0 .. its own position.  Output split fp16 [P*Tn][E].
__global__ void gpt2_attention_kernel(const float* __restrict__ qkv, float* __restrict__ kc, float* __restrict__ vc,
                                      __half* __restrict__ out_hi, __half* __restrict__ out_lo, int Tn, int past, int Tmax,
                                      int H, int E) {
  extern __shared__ float sm[];
  const int b = blockIdx.x / H, hd = blockIdx.x - b * H;
  const int ns = past + Tn;
  float* ks = sm;                              // [ns][65]
  float* vs = ks + Tmax * 65;
  float* qs = vs + Tmax * 65;                  // [Tn][65]
  float* sc = qs + Tn * 65;                    // [Tn][Tmax + 1]
  float* kcb = kc + ((size_t)(b * H + hd) * Tmax) * 64;
  float* vcb = vc + ((size_t)(b * H + hd) * Tmax) * 64;
  for (int i = threadIdx.x; i < Tn * 64; i += blockDim.x) {
    const int t = i >> 6, d = i & 63;
    const float* r = qkv + (size_t)(b * Tn + t) * 3 * E + hd * 64 + d;
    qs[t * 65 + d] = r[0];
    kcb[(size_t)(past + t) * 64 + d] = r[E];
    vcb[(size_t)(past + t) * 64 + d] = r[2 * E];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ns * 64; i += blockDim.x) {
    const int t = i >> 6, d = i & 63;
    ks[t * 65 + d] = kcb[(size_t)t * 64 + d];
    vs[t * 65 + d] = vcb[(size_t)t * 64 + d];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Tn * ns; i += blockDim.x) {
    const int a = i / ns, c = i - a * ns;
    float acc = 0.f;
#pragma unroll 16
    for (int d = 0; d < 64; ++d) acc = fmaf(qs[a * 65 + d], ks[c * 65 + d], acc);
    sc[a * (Tmax + 1) + c] = acc * 0.125f;                       // / sqrt(64)  (scale=True)
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int a = warp; a < Tn; a += nw) {
    const int lim = past + a + 1;                                // keys 0 .. own position (masked ones weigh exactly 0)
    float m = -INFINITY;
    for (int c = lane; c < lim; c += 32) m = fmaxf(m, sc[a * (Tmax + 1) + c]);
    m = warp_max_t(m);
    float s = 0.f;
    for (int c = lane; c < lim; c += 32) {
      const float e = expf(sc[a * (Tmax + 1) + c] - m);
      sc[a * (Tmax + 1) + c] = e;
      s += e;
    }
    s = warp_sum_t(s);
    const float inv = 1.f / s;
    for (int c = lane; c < ns; c += 32) sc[a * (Tmax + 1) + c] = c < lim ? sc[a * (Tmax + 1) + c] * inv : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Tn * 64; i += blockDim.x) {
    const int a = i >> 6, d = i & 63;
    float acc = 0.f;
    for (int c = 0; c < past + a + 1; ++c) acc = fmaf(sc[a * (Tmax + 1) + c], vs[c * 65 + d], acc);
    const size_t o = (size_t)(b * Tn + a) * E + hd * 64 + d;
    split_store(acc, out_hi + o, out_lo + o);
  }
}

// Single-query form of the kernel above for the decode steps (Tn = 1): one block of 64 threads per (candidate, head).
// The new key / value row is appended to the cache; scores: lane pairs... each of the 64 threads owns ONE key position
// (ns <= 64 per pass) and reads its 256-byte cache row; the output dimension d is then owned by thread d.  No
// shared-memory staging of the whole cache (that copy was most of the 16.6 us of the general kernel per step).
__global__ void __launch_bounds__(64) gpt2_attention_decode_kernel(const float* __restrict__ qkv, float* __restrict__ kc,
                                                                  float* __restrict__ vc, __half* __restrict__ out_hi,
                                                                  __half* __restrict__ out_lo, int past, int Tmax, int H,
                                                                  int E) {
  __shared__ float qs[64];
  __shared__ float pr[192];
  __shared__ float red[4];
  const int b = blockIdx.x / H, hd = blockIdx.x - b * H;
  const int ns = past + 1;
  const int d = threadIdx.x;
  float* kcb = kc + ((size_t)(b * H + hd) * Tmax) * 64;
  float* vcb = vc + ((size_t)(b * H + hd) * Tmax) * 64;
  const float* r = qkv + (size_t)b * 3 * E + hd * 64 + d;
  qs[d] = r[0];
  const float knew = r[E], vnew = r[2 * E];
  kcb[(size_t)past * 64 + d] = knew;
  vcb[(size_t)past * 64 + d] = vnew;
  __syncthreads();                                   // the appended row is read back below by other threads
  float m = -INFINITY;
  for (int c = threadIdx.x; c < ns; c += 64) {
    const float4* kr = reinterpret_cast<const float4*>(kcb + (size_t)c * 64);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 kv = kr[i];
      acc = fmaf(qs[4 * i], kv.x, acc); acc = fmaf(qs[4 * i + 1], kv.y, acc);
      acc = fmaf(qs[4 * i + 2], kv.z, acc); acc = fmaf(qs[4 * i + 3], kv.w, acc);
    }
    acc *= 0.125f;
    pr[c] = acc;
    m = fmaxf(m, acc);
  }
  m = warp_max_t(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = fmaxf(red[0], red[1]);
  float ssum = 0.f;
  for (int c = threadIdx.x; c < ns; c += 64) {
    const float e = expf(pr[c] - m);
    pr[c] = e;
    ssum += e;
  }
  ssum = warp_sum_t(ssum);
  if ((threadIdx.x & 31) == 0) red[2 + (threadIdx.x >> 5)] = ssum;
  __syncthreads();
  const float inv = 1.f / (red[2] + red[3]);
  float acc = 0.f;
  for (int c = 0; c < ns; ++c) acc = fmaf(pr[c] * inv, vcb[(size_t)c * 64 + d], acc);
  const size_t o = (size_t)b * E + hd * 64 + d;
  split_store(acc, out_hi + o, out_lo + o);
}

// next token = arg-max of the logits over [0, vocab) (first maximum, like torch.topk k=1 on distinct values);
// written to tokens[b][col] (gpt2/sample.py:31-35 with sample=False).  1024 threads per row, 16-byte loads.
__global__ void __launch_bounds__(1024) gpt2_argmax_kernel(const float* __restrict__ logits, int Npad, int vocab,
                                                           int* __restrict__ tokens, int Ttot, int col) {
  __shared__ float bv[32];
  __shared__ int bi[32];
  const int b = blockIdx.x;
  const float4* l4 = reinterpret_cast<const float4*>(logits + (size_t)b * Npad);
  float best = -INFINITY;
  int idx = 0x7fffffff;
  auto take = [&](float v, int i) {
    if (i < vocab && (v > best || (v == best && i < idx))) { best = v; idx = i; }
  };
  for (int i4 = threadIdx.x; i4 < Npad / 4; i4 += blockDim.x) {
    const float4 v = __ldg(l4 + i4);
    take(v.x, 4 * i4); take(v.y, 4 * i4 + 1); take(v.z, 4 * i4 + 2); take(v.w, 4 * i4 + 3);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { bv[w] = best; bi[w] = idx; }
  __syncthreads();
  if (w == 0) {
    best = lane < (int)(blockDim.x >> 5) ? bv[lane] : -INFINITY;
    idx = lane < (int)(blockDim.x >> 5) ? bi[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
    }
    if (lane == 0) tokens[(size_t)b * Ttot + col] = idx;
  }
}

__global__ void tokens_from_i64_kernel(const long long* __restrict__ src, int n_src_cols, int* __restrict__ dst, int Ttot,
                                       int P, const int* __restrict__ init, int n_init) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int ctx = n_src_cols + n_init;
  if (i >= P * ctx) return;
  const int b = i / ctx, t = i - b * ctx;
  dst[(size_t)b * Ttot + t] = t < n_src_cols ? (int)src[(size_t)b * n_src_cols + t] : init[t - n_src_cols];
}
__global__ void tokens_to_i64_kernel(const int* __restrict__ src, long long* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// ---------------------------------------------------------------------------
// CLIP text tower helpers (fp16 as built)
// ---------------------------------------------------------------------------
// x = fp16(fp16(token_embedding[tok]) + fp16(positional_embedding[t]))  (clip/model.py:308-310); also eot[b] = argmax_t tok
__global__ void text_embed_kernel(const long long* __restrict__ tokens, const __half* __restrict__ emb,
                                  const __half* __restrict__ pos, __half* __restrict__ x, int* __restrict__ eot, int T, int W) {
  const int row = blockIdx.x;
  const int b = row / T, t = row - b * T;
  const long long tok = tokens[row];
  for (int i = threadIdx.x; i < W; i += blockDim.x)
    x[(size_t)row * W + i] = __float2half_rn(__half2float(emb[(size_t)tok * W + i]) + __half2float(pos[(size_t)t * W + i]));
  if (t == 0 && threadIdx.x == 0) {
    long long best = tokens[row];
    int bi = 0;
    for (int k = 1; k < T; ++k)
      if (tokens[row + k] > best) { best = tokens[row + k]; bi = k; }   // first maximum (torch.argmax)
    eot[b] = bi;
  }
}

// clip/model.py:152-158 LayerNorm (fp32 arithmetic, eps 1e-5) on fp16 rows -> fp16
__global__ void text_layernorm_kernel(const __half* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bb,
                                      __half* __restrict__ out, int rows, int W) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const __half* xr = x + (size_t)r * W;
  float s = 0.f;
  for (int i = lane; i < W; i += 32) s += __half2float(xr[i]);
  const float mean = warp_sum_t(s) / (float)W;
  float q = 0.f;
  for (int i = lane; i < W; i += 32) { const float d = __half2float(xr[i]) - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum_t(q) / (float)W + 1e-5f);
  for (int i = lane; i < W; i += 32)
    out[(size_t)r * W + i] = __float2half_rn((__half2float(xr[i]) - mean) * rstd * w[i] + bb[i]);
}

// nn.MultiheadAttention core with the causal mask of clip/model.py:292-298: one block per (sequence, head), head dim 64
__global__ void text_attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int T, int W) {
  extern __shared__ float sm[];
  float* q = sm;
  float* k = q + T * 65;
  float* v = k + T * 65;
  float* sc = v + T * 65;                      // [T][T+1]
  const int heads = W / 64;
  const int b = blockIdx.x / heads, hd = blockIdx.x - b * heads;
  const __half* base = qkv + (size_t)b * T * 3 * W + hd * 64;
  for (int i = threadIdx.x; i < T * 64; i += blockDim.x) {
    const int t = i >> 6, d = i & 63;
    const __half* r = base + (size_t)t * 3 * W + d;
    q[t * 65 + d] = __half2float(r[0]) * 0.125f;
    k[t * 65 + d] = __half2float(r[W]);
    v[t * 65 + d] = __half2float(r[2 * W]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    const int a = i / T, c = i - a * T;
    float acc = 0.f;
    if (c <= a) {
#pragma unroll 16
      for (int d = 0; d < 64; ++d) acc = fmaf(q[a * 65 + d], k[c * 65 + d], acc);
    }
    sc[a * (T + 1) + c] = acc;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int a = warp; a < T; a += nw) {
    float m = -INFINITY;
    for (int c = lane; c <= a; c += 32) m = fmaxf(m, sc[a * (T + 1) + c]);
    m = warp_max_t(m);
    float s = 0.f;
    for (int c = lane; c <= a; c += 32) {
      const float e = __expf(sc[a * (T + 1) + c] - m);
      sc[a * (T + 1) + c] = e;
      s += e;
    }
    s = warp_sum_t(s);
    const float inv = 1.f / s;
    for (int c = lane; c <= a; c += 32) sc[a * (T + 1) + c] *= inv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * 64; i += blockDim.x) {
    const int a = i >> 6, d = i & 63;
    float acc = 0.f;
    for (int c = 0; c <= a; ++c) acc = fmaf(sc[a * (T + 1) + c], v[c * 65 + d], acc);
    out[((size_t)b * T + a) * W + hd * 64 + d] = __float2half_rn(acc);
  }
}

// ln_final of the EOT row, @ text_projection, cosine vs the cached image features (clip/model.py:314-320, generator.py:59)
__global__ void text_final_kernel(const __half* __restrict__ x, const int* __restrict__ eot, const float* lw, const float* lb,
                                  const float* __restrict__ proj, const float* __restrict__ image, float* features,
                                  float* sim, int T, int W, int E) {
  extern __shared__ float sm[];
  float* c = sm;
  __shared__ float red[32];
  const int b = blockIdx.x;
  const __half* xr = x + ((size_t)b * T + eot[b]) * W;
  auto block_sum = [&](float v) {
    v = warp_sum_t(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
    if (w == 0) {
      t = warp_sum_t(t);
      if (l == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
  };
  float s = 0.f;
  for (int i = threadIdx.x; i < W; i += blockDim.x) { c[i] = __half2float(xr[i]); s += c[i]; }
  const float mean = block_sum(s) / (float)W;
  float qv = 0.f;
  for (int i = threadIdx.x; i < W; i += blockDim.x) { const float d = c[i] - mean; qv += d * d; }
  const float rstd = rsqrtf(block_sum(qv) / (float)W + 1e-5f);
  for (int i = threadIdx.x; i < W; i += blockDim.x) c[i] = rh_t((c[i] - mean) * rstd * lw[i] + lb[i]);
  __syncthreads();
  float dot = 0.f, nf = 0.f, nt = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    int i = 0;
    for (; i + 4 <= W; i += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) a4[u] = fmaf(c[i + u], __ldg(proj + (size_t)(i + u) * E + e), a4[u]);
    }
    for (; i < W; ++i) a4[0] = fmaf(c[i], __ldg(proj + (size_t)i * E + e), a4[0]);
    const float acc = rh_t((a4[0] + a4[1]) + (a4[2] + a4[3]));
    if (features != nullptr) features[(size_t)b * E + e] = acc;
    const float t = image[e];
    dot += acc * t; nf += acc * acc; nt += t * t;
  }
  dot = block_sum(dot);
  nf = block_sum(nf);
  nt = block_sum(nt);
  if (threadIdx.x == 0) sim[b] = dot / fmaxf(sqrtf(nf) * sqrtf(nt), 1e-8f);
}

struct TTensor { void* ptr = nullptr; size_t bytes = 0; };

}  // namespace
}  // namespace glass

// ===========================================================================
// engine
// ===========================================================================
struct glass_text_engine {
  glass_text_config cfg{};
  int num_sms = 148;
  std::map<std::string, TTensor> tensors;
  bool finalized = false, have_image = false;
  PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  std::map<std::tuple<const void*, int, int, int>, CUtensorMap> maps;     // (ptr, rows, K, box rows)
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  int64_t launches = 0;
  // CUDA graphs of one whole decode (all ~2700 launches of glass_text_generate) and of one text-tower pass, per
  // population size; built on the second call of a size (the first runs eagerly and configures the kernels)
  cudaStream_t cap_stream = nullptr;
  std::map<int, cudaGraphExec_t> gen_graphs, sim_graphs;
  std::map<int, int> gen_calls, sim_calls;
  std::map<int, int64_t> gen_graph_launches, sim_graph_launches;
  // optional CUDA-event timing of the GEMM launches (bench.py roofline)
  bool timing = false;
  std::vector<cudaEvent_t> ev;
  size_t ev_used = 0;
  double gemm_bytes = 0;      // algorithmic bytes of the timed launches: operands read once + outputs written once
  // GPT-2 workspace
  int Tctx = 0, Ttot = 0, Npad = 0;
  int* tokens = nullptr;
  int* init_tokens = nullptr;
  long long* tok64 = nullptr;
  long long* zin64 = nullptr;
  float *h = nullptr, *qkv = nullptr, *logits = nullptr, *kcache = nullptr, *vcache = nullptr;
  float* partials = nullptr;   // split-K partial sums of the decode GEMMs: [splits][64][N]
  __half *a_hi = nullptr, *a_lo = nullptr, *g_hi = nullptr, *g_lo = nullptr, *f_hi = nullptr, *f_lo = nullptr;
  // CLIP text workspace
  long long* ctok = nullptr;
  int* eot = nullptr;
  __half *tx = nullptr, *th = nullptr, *tqkv = nullptr, *tatt = nullptr, *tfc = nullptr;
  float *tfeat = nullptr, *tsim = nullptr, *image = nullptr;
};

namespace glass {
namespace {

template <class T>
T* tt(glass_text_engine* e, const std::string& name) {
  auto it = e->tensors.find(name);
  return it == e->tensors.end() ? nullptr : reinterpret_cast<T*>(it->second.ptr);
}
int tcheck(glass_text_engine* e, const std::string& name, size_t bytes) {
  auto it = e->tensors.find(name);
  if (it == e->tensors.end()) return tfail(GLASS_ERR_STATE, "weight tensor '%s' was not set", name.c_str());
  if (it->second.bytes != bytes)
    return tfail(GLASS_ERR_ARG, "weight tensor '%s' has %zu bytes, expected %zu", name.c_str(), it->second.bytes, bytes);
  return GLASS_OK;
}

// fp16 row-major [rows][K] tensor as a 2-D TMA map with a (64 x box_rows) box, 128-byte swizzle; rows beyond `rows`
// are zero-filled by the TMA unit
int get_map(glass_text_engine* e, const void* ptr, int rows, int K, int box_rows, const CUtensorMap** out) {
  auto key = std::make_tuple(ptr, rows, K, box_rows);
  auto it = e->maps.find(key);
  if (it == e->maps.end()) {
    CUtensorMap m;
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = e->encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return tfail(GLASS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a [%d x %d] fp16 tensor", (int)r, rows, K);
    it = e->maps.emplace(key, m).first;
  }
  *out = &it->second;
  return GLASS_OK;
}

template <int BN, bool kSplit, int kARows>
int launch_gemm_t(glass_text_engine* e, const __half* a_hi, const __half* a_lo, const __half* w_hi, const __half* w_lo,
                  const GemmParams& p_in, cudaStream_t s) {
  using C = GemmCfg<BN, kSplit, kARows>;
  static bool configured = false;
  if (!configured) {
    TCUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, kSplit, kARows>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    configured = true;
  }
  const CUtensorMap *mah, *mal, *mwh, *mwl;
  GemmParams p = p_in;
  p.a_rows = kARows;
  TRC(get_map(e, a_hi, p.M, p.K, p.a_rows, &mah));
  TRC(get_map(e, w_hi, p.N, p.K, BN, &mwh));
  mal = mah;
  mwl = mwh;
  if (kSplit) {
    TRC(get_map(e, a_lo, p.M, p.K, p.a_rows, &mal));
    TRC(get_map(e, w_lo, p.N, p.K, BN, &mwl));
  }
  dim3 grid(p.N / BN, (p.M + 127) / 128, p.splits > 1 ? p.splits : 1);
  const bool timed = e->timing && e->ev_used + 2 <= e->ev.size();
  if (timed) cudaEventRecord(e->ev[e->ev_used], s);
  gemm_tc_kernel<BN, kSplit, kARows><<<grid, 192, C::kSmemBytes, s>>>(*mah, *mal, *mwh, *mwl, p);
  if (timed) {
    cudaEventRecord(e->ev[e->ev_used + 1], s);
    e->ev_used += 2;
    const double in_b = ((double)p.M * p.K + (double)p.N * p.K) * 2.0 * (kSplit ? 2 : 1);
    const double out_b = (double)p.M * p.N * ((p.out_f32 ? 4 : 0) + (p.out_hi ? 2 : 0) + (p.out_lo ? 2 : 0) +
                                              (p.res_f32 ? 4 : 0) + (p.res_f16 ? 2 : 0));
    e->gemm_bytes += in_b + out_b;
  }
  e->launches++;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return tfail(GLASS_ERR_CUDA, "gemm_tc_kernel launch failed: %s", cudaGetErrorString(err));
  return GLASS_OK;
}

// N tile: 64 columns, 32 when that is what it takes to put at least ~64 CTAs on the machine (decode steps have one
// 128-row m-tile: the GEMM is a weight-streaming pass, parallelism comes from N)
template <bool kSplit>
int launch_gemm(glass_text_engine* e, const __half* a_hi, const __half* a_lo, const __half* w_hi, const __half* w_lo,
                const GemmParams& p, cudaStream_t s) {
  if (p.K % 64 != 0 || p.N % 32 != 0 || p.M <= 0) return tfail(GLASS_ERR_ARG, "unsupported GEMM shape %dx%dx%d", p.M, p.N, p.K);
  const int m_tiles = (p.M + 127) / 128;
  if (p.M <= 64) {                                   // decode steps: half-height A tiles, deep pipeline
    if (p.N % 64 == 0 && p.N / 64 >= 96) return launch_gemm_t<64, kSplit, 64>(e, a_hi, a_lo, w_hi, w_lo, p, s);
    return launch_gemm_t<32, kSplit, 64>(e, a_hi, a_lo, w_hi, w_lo, p, s);
  }
  if (p.N % 64 == 0 && (p.N / 64) * m_tiles >= 96) return launch_gemm_t<64, kSplit, 128>(e, a_hi, a_lo, w_hi, w_lo, p, s);
  return launch_gemm_t<32, kSplit, 128>(e, a_hi, a_lo, w_hi, w_lo, p, s);
}

// Decode-step GEMM (M <= 64 rows): 128-column tiles and split-K so that n_tiles * splits CTAs fill the machine;
// writes `splits` fp32 partial sums [splits][M][N] that the consumer kernel reduces in a fixed order.
int launch_gemm_decode(glass_text_engine* e, const __half* a_hi, const __half* a_lo, const __half* w_hi, const __half* w_lo,
                       int M, int N, int K, float* partials, int* splits_out, cudaStream_t s) {
  if (K % 64 != 0 || N % 128 != 0 || M <= 0 || M > 64) return tfail(GLASS_ERR_ARG, "unsupported decode GEMM %dx%dx%d", M, N, K);
  const int kiters = K / 64, n_tiles = N / 128;
  int S = 1;
  for (int d = 1; d <= kiters; ++d)
    if (kiters % d == 0 && n_tiles * d <= e->num_sms) S = d;
  GemmParams g{};
  g.M = M; g.N = N; g.K = K; g.out_f32 = partials; g.splits = S;
  *splits_out = S;
  return launch_gemm_t<128, true, 64>(e, a_hi, a_lo, w_hi, w_lo, g, s);
}

#define TLAUNCH(expr)                                                                              \
  do {                                                                                             \
    expr;                                                                                          \
    e->launches++;                                                                                 \
    cudaError_t err__ = cudaGetLastError();                                                        \
    if (err__ != cudaSuccess) return tfail(GLASS_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(err__)); \
  } while (0)

int validate_text(glass_text_engine* e) {
  const glass_text_config& c = e->cfg;
  char nm[64];
  const size_t E = c.gpt2_embd;
  if (c.gpt2_layers > 0) {
    TRC(tcheck(e, "g2.wte", (size_t)c.gpt2_vocab * E * 4));
    TRC(tcheck(e, "g2.wte.hi", (size_t)e->Npad * E * 2));
    TRC(tcheck(e, "g2.wte.lo", (size_t)e->Npad * E * 2));
    TRC(tcheck(e, "g2.wpe", (size_t)c.gpt2_positions * E * 4));
    TRC(tcheck(e, "g2.init", (size_t)c.n_init * 4));
    for (const char* n : {"g2.lnf.w", "g2.lnf.b"}) TRC(tcheck(e, n, E * 4));
    for (int l = 0; l < c.gpt2_layers; ++l) {
      auto f = [&](const char* sfx) { snprintf(nm, sizeof nm, "g2.l%d.%s", l, sfx); return std::string(nm); };
      for (const char* n : {"ln1.w", "ln1.b", "ln2.w", "ln2.b", "proj.b", "proj2.b"}) TRC(tcheck(e, f(n), E * 4));
      TRC(tcheck(e, f("attn.b"), 3 * E * 4));
      TRC(tcheck(e, f("fc.b"), 4 * E * 4));
      for (const char* hl : {"hi", "lo"}) {
        TRC(tcheck(e, f((std::string("attn.w.") + hl).c_str()), 3 * E * E * 2));
        TRC(tcheck(e, f((std::string("proj.w.") + hl).c_str()), E * E * 2));
        TRC(tcheck(e, f((std::string("fc.w.") + hl).c_str()), 4 * E * E * 2));
        TRC(tcheck(e, f((std::string("proj2.w.") + hl).c_str()), 4 * E * E * 2));
      }
    }
  }
  const size_t W = c.text_width;
  if (c.text_layers > 0) {
    TRC(tcheck(e, "t.tok", (size_t)c.text_vocab * W * 2));
    TRC(tcheck(e, "t.pos", (size_t)c.text_context * W * 2));
    for (const char* n : {"t.lnf.w", "t.lnf.b"}) TRC(tcheck(e, n, W * 4));
    TRC(tcheck(e, "t.proj", W * c.text_embed_dim * 4));
    for (int l = 0; l < c.text_layers; ++l) {
      auto f = [&](const char* sfx) { snprintf(nm, sizeof nm, "t.l%d.%s", l, sfx); return std::string(nm); };
      for (const char* n : {"ln1.w", "ln1.b", "ln2.w", "ln2.b", "out.b", "proj.b"}) TRC(tcheck(e, f(n), W * 4));
      TRC(tcheck(e, f("qkv.w"), 3 * W * W * 2));
      TRC(tcheck(e, f("qkv.b"), 3 * W * 4));
      TRC(tcheck(e, f("out.w"), W * W * 2));
      TRC(tcheck(e, f("fc.w"), 4 * W * W * 2));
      TRC(tcheck(e, f("fc.b"), 4 * W * 4));
      TRC(tcheck(e, f("proj.w"), 4 * W * W * 2));
    }
  }
  return GLASS_OK;
}

struct Carver {
  uint8_t* base = nullptr;
  size_t off = 0;
  template <class T>
  T* take(size_t n) {
    off = (off + 1023) & ~size_t(1023);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

void layout_text(glass_text_engine* e, Carver& a) {
  const glass_text_config& c = e->cfg;
  const size_t P = c.max_population;
  if (c.gpt2_layers > 0) {
    const size_t E = c.gpt2_embd, M = P * e->Tctx, H = c.gpt2_heads;
    e->tokens = a.take<int>(P * e->Ttot);
    e->init_tokens = a.take<int>(16);
    e->tok64 = a.take<long long>(P * e->Ttot);
    e->zin64 = a.take<long long>(P * c.dim_z);
    e->h = a.take<float>(M * E);
    e->qkv = a.take<float>(M * 3 * E);
    e->logits = a.take<float>(P * e->Npad);
    e->partials = a.take<float>((size_t)e->num_sms * 64 * 128 + 1024);   // n_tiles * splits <= num_sms tiles of 64 x 128
    e->kcache = a.take<float>((size_t)c.gpt2_layers * P * H * e->Ttot * 64);
    e->vcache = a.take<float>((size_t)c.gpt2_layers * P * H * e->Ttot * 64);
    e->a_hi = a.take<__half>(M * E);
    e->a_lo = a.take<__half>(M * E);
    e->g_hi = a.take<__half>(M * 4 * E);
    e->g_lo = a.take<__half>(M * 4 * E);
    e->f_hi = a.take<__half>(P * E);
    e->f_lo = a.take<__half>(P * E);
  }
  if (c.text_layers > 0) {
    const size_t W = c.text_width, M = P * c.text_context;
    e->ctok = a.take<long long>(M);
    e->eot = a.take<int>(P);
    e->tx = a.take<__half>(M * W);
    e->th = a.take<__half>(M * W);
    e->tqkv = a.take<__half>(M * 3 * W);
    e->tatt = a.take<__half>(M * W);
    e->tfc = a.take<__half>(M * 4 * W);
    e->tfeat = a.take<float>(P * c.text_embed_dim);
    e->tsim = a.take<float>(P);
    e->image = a.take<float>(c.text_embed_dim);
  }
}

// one GPT2LMHeadModel.forward over Tn new positions starting at column col0 (gpt2/model.py:126-175, 196-210); the
// arg-max of the last position's logits lands in tokens[:, col0 + Tn]
// Decode step (one new position per candidate, M = P <= 64 rows): every GEMM is a weight-streaming pass with one
// 128-row m-tile, so it runs split-K over (N / 128) x splits ~ 148 CTAs and leaves fp32 partial sums; the reduction is
// fused into the kernel that consumes the result (residual update + LayerNorm, bias + GELU, bias).
int gpt2_forward_decode(glass_text_engine* e, int P, int col0, cudaStream_t s) {
  const glass_text_config& c = e->cfg;
  const int E = c.gpt2_embd, H = c.gpt2_heads, M = P;
  char nm[64];
  int S = 1;
  const size_t ps = (size_t)M;             // partial stride factor: M * N floats per split
  TLAUNCH((gpt2_embed_kernel<<<M, 128, 0, s>>>(e->tokens, e->Ttot, col0, 1, tt<float>(e, "g2.wte"), tt<float>(e, "g2.wpe"),
                                               e->h, E)));
  for (int l = 0; l < c.gpt2_layers; ++l) {
    auto f = [&](const char* sfx) { snprintf(nm, sizeof nm, "g2.l%d.%s", l, sfx); return std::string(nm); };
    auto fp = [&](const char* sfx) { snprintf(nm, sizeof nm, "g2.l%d.%s", l - 1, sfx); return std::string(nm); };
    if (l == 0) {
      TLAUNCH((gpt2_layernorm_split_kernel<<<M, 128, 0, s>>>(e->h, 1, 0, tt<float>(e, f("ln1.w")), tt<float>(e, f("ln1.b")),
                                                             c.gpt2_eps, e->a_hi, e->a_lo, M, E)));
    } else {      // h += mlp.c_proj of the previous layer (pending partial sums), then this layer's ln_1
      const float* pb = tt<float>(e, fp("proj2.b"));
      TLAUNCH((gpt2_reduce_residual_ln_kernel<<<M, 128, 0, s>>>(e->partials, S, ps * E, pb, e->h, tt<float>(e, f("ln1.w")),
                                                                tt<float>(e, f("ln1.b")), c.gpt2_eps, e->a_hi, e->a_lo, E)));
    }
    TRC(launch_gemm_decode(e, e->a_hi, e->a_lo, tt<__half>(e, f("attn.w.hi")), tt<__half>(e, f("attn.w.lo")), M, 3 * E, E,
                           e->partials, &S, s));
    TLAUNCH((gpt2_reduce_bias_kernel<<<(M * 3 * E + 255) / 256, 256, 0, s>>>(e->partials, S, ps * 3 * E, tt<float>(e, f("attn.b")),
                                                                             e->qkv, M, 3 * E)));
    const size_t coff = (size_t)l * P * H * e->Ttot * 64;
    TLAUNCH((gpt2_attention_decode_kernel<<<P * H, 64, 0, s>>>(e->qkv, e->kcache + coff, e->vcache + coff, e->a_hi, e->a_lo,
                                                               col0, e->Ttot, H, E)));
    TRC(launch_gemm_decode(e, e->a_hi, e->a_lo, tt<__half>(e, f("proj.w.hi")), tt<__half>(e, f("proj.w.lo")), M, E, E,
                           e->partials, &S, s));
    TLAUNCH((gpt2_reduce_residual_ln_kernel<<<M, 128, 0, s>>>(e->partials, S, ps * E, tt<float>(e, f("proj.b")), e->h,
                                                              tt<float>(e, f("ln2.w")), tt<float>(e, f("ln2.b")), c.gpt2_eps,
                                                              e->a_hi, e->a_lo, E)));
    TRC(launch_gemm_decode(e, e->a_hi, e->a_lo, tt<__half>(e, f("fc.w.hi")), tt<__half>(e, f("fc.w.lo")), M, 4 * E, E,
                           e->partials, &S, s));
    TLAUNCH((gpt2_reduce_gelu_split_kernel<<<(M * 4 * E + 255) / 256, 256, 0, s>>>(e->partials, S, ps * 4 * E,
                                                                                   tt<float>(e, f("fc.b")), e->g_hi, e->g_lo,
                                                                                   M, 4 * E)));
    TRC(launch_gemm_decode(e, e->g_hi, e->g_lo, tt<__half>(e, f("proj2.w.hi")), tt<__half>(e, f("proj2.w.lo")), M, E, 4 * E,
                           e->partials, &S, s));
  }
  snprintf(nm, sizeof nm, "g2.l%d.proj2.b", c.gpt2_layers - 1);
  TLAUNCH((gpt2_reduce_residual_ln_kernel<<<M, 128, 0, s>>>(e->partials, S, ps * E, tt<float>(e, nm), e->h,
                                                            tt<float>(e, "g2.lnf.w"), tt<float>(e, "g2.lnf.b"), c.gpt2_eps,
                                                            e->f_hi, e->f_lo, E)));
  GemmParams g{};
  g.M = P; g.N = e->Npad; g.K = E; g.out_f32 = e->logits;
  TRC(launch_gemm<true>(e, e->f_hi, e->f_lo, tt<__half>(e, "g2.wte.hi"), tt<__half>(e, "g2.wte.lo"), g, s));
  TLAUNCH((gpt2_argmax_kernel<<<P, 1024, 0, s>>>(e->logits, e->Npad, c.gpt2_vocab, e->tokens, e->Ttot, col0 + 1)));
  return GLASS_OK;
}

int gpt2_forward(glass_text_engine* e, int P, int col0, int Tn, cudaStream_t s) {
  const glass_text_config& c = e->cfg;
  const int E = c.gpt2_embd, H = c.gpt2_heads, M = P * Tn;
  if (Tn == 1 && P <= 64 && E % 128 == 0 && col0 + 1 <= 192 && !(c.flags & GLASS_TEXT_FLAG_NO_SPLIT_K))
    return gpt2_forward_decode(e, P, col0, s);
  char nm[64];
  TLAUNCH((gpt2_embed_kernel<<<M, 128, 0, s>>>(e->tokens, e->Ttot, col0, Tn, tt<float>(e, "g2.wte"), tt<float>(e, "g2.wpe"),
                                               e->h, E)));
  const size_t attn_smem = sizeof(float) * ((size_t)2 * e->Ttot * 65 + (size_t)Tn * 65 + (size_t)Tn * (e->Ttot + 1));
  for (int l = 0; l < c.gpt2_layers; ++l) {
    auto f = [&](const char* sfx) { snprintf(nm, sizeof nm, "g2.l%d.%s", l, sfx); return std::string(nm); };
    TLAUNCH((gpt2_layernorm_split_kernel<<<M, 128, 0, s>>>(e->h, 1, 0, tt<float>(e, f("ln1.w")),
                                                                      tt<float>(e, f("ln1.b")), c.gpt2_eps, e->a_hi, e->a_lo, M, E)));
    GemmParams g{};
    g.M = M; g.N = 3 * E; g.K = E; g.bias = tt<float>(e, f("attn.b")); g.out_f32 = e->qkv;
    TRC(launch_gemm<true>(e, e->a_hi, e->a_lo, tt<__half>(e, f("attn.w.hi")), tt<__half>(e, f("attn.w.lo")), g, s));
    const size_t coff = (size_t)l * P * H * e->Ttot * 64;
    if (Tn == 1 && col0 + 1 <= 192)
      TLAUNCH((gpt2_attention_decode_kernel<<<P * H, 64, 0, s>>>(e->qkv, e->kcache + coff, e->vcache + coff, e->a_hi,
                                                                 e->a_lo, col0, e->Ttot, H, E)));
    else
      TLAUNCH((gpt2_attention_kernel<<<P * H, 128, attn_smem, s>>>(e->qkv, e->kcache + coff, e->vcache + coff, e->a_hi,
                                                                    e->a_lo, Tn, col0, e->Ttot, H, E)));
    g = GemmParams{};
    g.M = M; g.N = E; g.K = E; g.bias = tt<float>(e, f("proj.b")); g.res_f32 = e->h; g.out_f32 = e->h;
    TRC(launch_gemm<true>(e, e->a_hi, e->a_lo, tt<__half>(e, f("proj.w.hi")), tt<__half>(e, f("proj.w.lo")), g, s));
    TLAUNCH((gpt2_layernorm_split_kernel<<<M, 128, 0, s>>>(e->h, 1, 0, tt<float>(e, f("ln2.w")),
                                                                      tt<float>(e, f("ln2.b")), c.gpt2_eps, e->a_hi, e->a_lo, M, E)));
    g = GemmParams{};
    g.M = M; g.N = 4 * E; g.K = E; g.bias = tt<float>(e, f("fc.b")); g.act = kTActGeluTanh; g.out_hi = e->g_hi; g.out_lo = e->g_lo;
    TRC(launch_gemm<true>(e, e->a_hi, e->a_lo, tt<__half>(e, f("fc.w.hi")), tt<__half>(e, f("fc.w.lo")), g, s));
    g = GemmParams{};
    g.M = M; g.N = E; g.K = 4 * E; g.bias = tt<float>(e, f("proj2.b")); g.res_f32 = e->h; g.out_f32 = e->h;
    TRC(launch_gemm<true>(e, e->g_hi, e->g_lo, tt<__half>(e, f("proj2.w.hi")), tt<__half>(e, f("proj2.w.lo")), g, s));
  }
  // ln_f and the LM head on the last position only (gpt2/sample.py:29: logits[:, -1, :])
  TLAUNCH((gpt2_layernorm_split_kernel<<<P, 128, 0, s>>>(e->h, Tn, Tn - 1, tt<float>(e, "g2.lnf.w"),
                                                                    tt<float>(e, "g2.lnf.b"), c.gpt2_eps, e->f_hi, e->f_lo, P, E)));
  GemmParams g{};
  g.M = P; g.N = e->Npad; g.K = E; g.out_f32 = e->logits;
  TRC(launch_gemm<true>(e, e->f_hi, e->f_lo, tt<__half>(e, "g2.wte.hi"), tt<__half>(e, "g2.wte.lo"), g, s));
  TLAUNCH((gpt2_argmax_kernel<<<P, 1024, 0, s>>>(e->logits, e->Npad, c.gpt2_vocab, e->tokens, e->Ttot, col0 + Tn)));
  return GLASS_OK;
}


// the launch sequence of glass_text_generate between the H2D of z and the D2H of the tokens
int generate_body(glass_text_engine* e, int pop, cudaStream_t s) {
  const glass_text_config& c = e->cfg;
  TLAUNCH((tokens_from_i64_kernel<<<(pop * e->Tctx + 255) / 256, 256, 0, s>>>(e->zin64, c.dim_z, e->tokens, e->Ttot, pop,
                                                                              e->init_tokens, c.n_init)));
  TRC(gpt2_forward(e, pop, 0, e->Tctx, s));                                   // models.py:47-48: the whole context
  for (int step = 1; step < c.max_tokens_len; ++step)                         // gpt2/sample.py:26-35
    TRC(gpt2_forward(e, pop, e->Tctx + step - 1, 1, s));
  TLAUNCH((tokens_to_i64_kernel<<<(pop * e->Ttot + 255) / 256, 256, 0, s>>>(e->tokens, e->tok64, pop * e->Ttot)));
  return GLASS_OK;
}

// the launch sequence of glass_text_similarity between the H2D of the tokens and the D2H of the scores
int similarity_body(glass_text_engine* e, int pop, cudaStream_t s) {
  const glass_text_config& c = e->cfg;
  const int T = c.text_context, W = c.text_width, M = pop * T;
  char nm[64];
  TLAUNCH((text_embed_kernel<<<M, 128, 0, s>>>(e->ctok, tt<__half>(e, "t.tok"), tt<__half>(e, "t.pos"), e->tx, e->eot, T, W)));
  const size_t attn_smem = sizeof(float) * ((size_t)3 * T * 65 + (size_t)T * (T + 1));
  for (int l = 0; l < c.text_layers; ++l) {
    auto f = [&](const char* sfx) { snprintf(nm, sizeof nm, "t.l%d.%s", l, sfx); return std::string(nm); };
    TLAUNCH((text_layernorm_kernel<<<(M + 7) / 8, 256, 0, s>>>(e->tx, tt<float>(e, f("ln1.w")), tt<float>(e, f("ln1.b")), e->th, M, W)));
    GemmParams g{};
    g.M = M; g.N = 3 * W; g.K = W; g.bias = tt<float>(e, f("qkv.b")); g.out_hi = e->tqkv;
    TRC(launch_gemm<false>(e, e->th, nullptr, tt<__half>(e, f("qkv.w")), nullptr, g, s));
    TLAUNCH((text_attention_kernel<<<pop * c.text_heads, 256, attn_smem, s>>>(e->tqkv, e->tatt, T, W)));
    g = GemmParams{};
    g.M = M; g.N = W; g.K = W; g.bias = tt<float>(e, f("out.b")); g.round_fp16 = 1; g.res_f16 = e->tx; g.out_hi = e->tx;
    TRC(launch_gemm<false>(e, e->tatt, nullptr, tt<__half>(e, f("out.w")), nullptr, g, s));
    TLAUNCH((text_layernorm_kernel<<<(M + 7) / 8, 256, 0, s>>>(e->tx, tt<float>(e, f("ln2.w")), tt<float>(e, f("ln2.b")), e->th, M, W)));
    g = GemmParams{};
    g.M = M; g.N = 4 * W; g.K = W; g.bias = tt<float>(e, f("fc.b")); g.round_fp16 = 1; g.act = kTActQuickGelu; g.out_hi = e->tfc;
    TRC(launch_gemm<false>(e, e->th, nullptr, tt<__half>(e, f("fc.w")), nullptr, g, s));
    g = GemmParams{};
    g.M = M; g.N = W; g.K = 4 * W; g.bias = tt<float>(e, f("proj.b")); g.round_fp16 = 1; g.res_f16 = e->tx; g.out_hi = e->tx;
    TRC(launch_gemm<false>(e, e->tfc, nullptr, tt<__half>(e, f("proj.w")), nullptr, g, s));
  }
  TLAUNCH((text_final_kernel<<<pop, 256, W * sizeof(float), s>>>(e->tx, e->eot, tt<float>(e, "t.lnf.w"), tt<float>(e, "t.lnf.b"),
                                                                  tt<float>(e, "t.proj"), e->image, e->tfeat, e->tsim, T, W, c.text_embed_dim)));
  return GLASS_OK;
}

// First call of a population size: eager launches (configures kernel attributes, fills the tensor-map cache).  Second
// call: capture the same sequence into a CUDA graph (all pointers are engine-owned and fixed).  Later calls: replay.
// Timing mode always launches eagerly.
int run_or_replay(glass_text_engine* e, int pop, cudaStream_t s, std::map<int, cudaGraphExec_t>& graphs,
                  std::map<int, int>& calls, std::map<int, int64_t>& graph_launches,
                  int (*body)(glass_text_engine*, int, cudaStream_t)) {
  const bool use_graph = !e->timing && !(e->cfg.flags & GLASS_FLAG_NO_GRAPH);
  if (!use_graph || calls[pop]++ == 0) return body(e, pop, s);
  auto it = graphs.find(pop);
  if (it == graphs.end()) {
    const int64_t before = e->launches;
    cudaGraph_t graph = nullptr;
    TCUDA_OK(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = body(e, pop, e->cap_stream);
    cudaError_t err = cudaStreamEndCapture(e->cap_stream, &graph);
    graph_launches[pop] = e->launches - before;
    e->launches = before;
    if (rc != GLASS_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (err != cudaSuccess) return tfail(GLASS_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(err));
    cudaGraphExec_t exec = nullptr;
    err = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (err != cudaSuccess) return tfail(GLASS_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(err));
    it = graphs.emplace(pop, exec).first;
  }
  TCUDA_OK(cudaGraphLaunch(it->second, s));
  e->launches += graph_launches[pop];
  return GLASS_OK;
}

}  // namespace
}  // namespace glass

extern "C" {

const char* glass_text_last_error(void) { return t_last_error.c_str(); }

int glass_text_create(const glass_text_config* cfg, glass_text_engine** out) {
  if (!cfg || !out) return tfail(GLASS_ERR_ARG, "null argument");
  if (cfg->max_population <= 0) return tfail(GLASS_ERR_ARG, "max_population must be positive");
  if (cfg->gpt2_layers > 0) {
    if (cfg->gpt2_embd % 64 != 0 || cfg->gpt2_heads <= 0 || cfg->gpt2_embd / cfg->gpt2_heads != 64)
      return tfail(GLASS_ERR_ARG, "GPT-2: n_embd must be a multiple of 64 with 64-wide heads");
    if (cfg->dim_z <= 0 || cfg->n_init < 0 || cfg->n_init > 16 || cfg->max_tokens_len <= 0 ||
        cfg->dim_z + cfg->n_init + cfg->max_tokens_len > cfg->gpt2_positions)
      return tfail(GLASS_ERR_ARG, "GPT-2: dim_z + init tokens + max_tokens_len must fit n_positions");
    if (cfg->dim_z + cfg->n_init + cfg->max_tokens_len > 160) return tfail(GLASS_ERR_ARG, "GPT-2: sequences longer than 160 are not supported");
  }
  if (cfg->text_layers > 0) {
    if (cfg->text_width % 64 != 0 || cfg->text_heads * 64 != cfg->text_width || cfg->text_context > 96)
      return tfail(GLASS_ERR_ARG, "CLIP text tower: width must be heads x 64, context <= 96");
  }
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0 || cfg->device >= ndev)
    return tfail(GLASS_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                 err == cudaSuccess ? "device ordinal out of range" : cudaGetErrorString(err));
  TCUDA_OK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  TCUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return tfail(GLASS_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (err != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess)
    return tfail(GLASS_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
  glass_text_engine* e = new glass_text_engine();
  e->cfg = *cfg;
  e->num_sms = prop.multiProcessorCount;
  e->encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  e->Tctx = cfg->dim_z + cfg->n_init;
  e->Ttot = e->Tctx + cfg->max_tokens_len;
  e->Npad = (cfg->gpt2_vocab + 63) / 64 * 64;
  *out = e;
  return GLASS_OK;
}

int glass_text_set_tensor(glass_text_engine* e, const char* name, const void* host_data, size_t nbytes) {
  if (!e || !name || !host_data || nbytes == 0) return tfail(GLASS_ERR_ARG, "null argument");
  TCUDA_OK(cudaSetDevice(e->cfg.device));
  TTensor& t = e->tensors[name];
  if (t.ptr && t.bytes != nbytes) { cudaFree(t.ptr); t.ptr = nullptr; }
  if (!t.ptr) TCUDA_OK(cudaMalloc(&t.ptr, nbytes));
  t.bytes = nbytes;
  TCUDA_OK(cudaMemcpy(t.ptr, host_data, nbytes, cudaMemcpyHostToDevice));
  e->maps.clear();
  return GLASS_OK;
}

int glass_text_finalize(glass_text_engine* e) {
  if (!e) return tfail(GLASS_ERR_ARG, "null engine");
  TCUDA_OK(cudaSetDevice(e->cfg.device));
  TRC(validate_text(e));
  Carver probe;
  layout_text(e, probe);
  const size_t need = probe.off + 4096;
  void* base = nullptr;
  cudaError_t err = cudaMalloc(&base, need);
  if (err != cudaSuccess)
    return tfail(GLASS_ERR_NOMEM, "workspace of %.2f GB for max_population=%d: %s", need / 1e9, e->cfg.max_population, cudaGetErrorString(err));
  TCUDA_OK(cudaMemset(base, 0, need));
  e->arena = (uint8_t*)base;
  e->arena_bytes = need;
  Carver real;
  real.base = e->arena;
  layout_text(e, real);
  if (e->cfg.gpt2_layers > 0)
    TCUDA_OK(cudaMemcpy(e->init_tokens, tt<int>(e, "g2.init"), (size_t)e->cfg.n_init * 4, cudaMemcpyDeviceToDevice));
  TCUDA_OK(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
  TCUDA_OK(cudaFuncSetAttribute(gpt2_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  TCUDA_OK(cudaFuncSetAttribute(text_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
  e->finalized = true;
  return GLASS_OK;
}

int glass_text_set_image_features(glass_text_engine* e, const float* host_image, int32_t n) {
  if (!e || !e->finalized || e->cfg.text_layers <= 0) return tfail(GLASS_ERR_STATE, "engine has no finalized CLIP text tower");
  if (n != e->cfg.text_embed_dim) return tfail(GLASS_ERR_ARG, "image feature length %d != embed dim %d", n, e->cfg.text_embed_dim);
  TCUDA_OK(cudaSetDevice(e->cfg.device));
  TCUDA_OK(cudaMemcpy(e->image, host_image, (size_t)n * 4, cudaMemcpyHostToDevice));
  e->have_image = true;
  return GLASS_OK;
}

int glass_text_generate(glass_text_engine* e, const int64_t* z_host, int32_t pop, int64_t* tokens_host, void* stream) {
  if (!e || !e->finalized || e->cfg.gpt2_layers <= 0) return tfail(GLASS_ERR_STATE, "engine has no finalized GPT-2");
  if (!z_host || !tokens_host) return tfail(GLASS_ERR_ARG, "null argument");
  if (pop <= 0 || pop > e->cfg.max_population) return tfail(GLASS_ERR_ARG, "population %d outside (0, %d]", pop, e->cfg.max_population);
  const glass_text_config& c = e->cfg;
  for (size_t i = 0; i < (size_t)pop * c.dim_z; ++i)
    if (z_host[i] < 0 || z_host[i] >= c.gpt2_vocab)
      return tfail(GLASS_ERR_ARG, "latent token %lld outside [0, %d) (the reference's embedding lookup raises IndexError)",
                   (long long)z_host[i], c.gpt2_vocab);
  TCUDA_OK(cudaSetDevice(c.device));
  cudaStream_t s = (cudaStream_t)stream;
  TCUDA_OK(cudaMemcpyAsync(e->zin64, z_host, (size_t)pop * c.dim_z * 8, cudaMemcpyHostToDevice, s));
  TRC(run_or_replay(e, pop, s, e->gen_graphs, e->gen_calls, e->gen_graph_launches, generate_body));
  TCUDA_OK(cudaMemcpyAsync(tokens_host, e->tok64, (size_t)pop * e->Ttot * 8, cudaMemcpyDeviceToHost, s));
  TCUDA_OK(cudaStreamSynchronize(s));
  return GLASS_OK;
}

int glass_text_similarity(glass_text_engine* e, const int64_t* clip_tokens_host, int32_t pop, float* sim_host,
                          float* features_host, void* stream) {
  if (!e || !e->finalized || e->cfg.text_layers <= 0) return tfail(GLASS_ERR_STATE, "engine has no finalized CLIP text tower");
  if (!e->have_image) return tfail(GLASS_ERR_STATE, "glass_text_set_image_features was not called");
  if (!clip_tokens_host || !sim_host) return tfail(GLASS_ERR_ARG, "null argument");
  if (pop <= 0 || pop > e->cfg.max_population) return tfail(GLASS_ERR_ARG, "population %d outside (0, %d]", pop, e->cfg.max_population);
  const glass_text_config& c = e->cfg;
  const int T = c.text_context, W = c.text_width, M = pop * T;
  for (size_t i = 0; i < (size_t)M; ++i)
    if (clip_tokens_host[i] < 0 || clip_tokens_host[i] >= c.text_vocab)
      return tfail(GLASS_ERR_ARG, "CLIP token %lld outside [0, %d)", (long long)clip_tokens_host[i], c.text_vocab);
  TCUDA_OK(cudaSetDevice(c.device));
  cudaStream_t s = (cudaStream_t)stream;
  char nm[64];
  TCUDA_OK(cudaMemcpyAsync(e->ctok, clip_tokens_host, (size_t)M * 8, cudaMemcpyHostToDevice, s));
  TRC(run_or_replay(e, pop, s, e->sim_graphs, e->sim_calls, e->sim_graph_launches, similarity_body));
  TCUDA_OK(cudaMemcpyAsync(sim_host, e->tsim, (size_t)pop * 4, cudaMemcpyDeviceToHost, s));
  if (features_host != nullptr)
    TCUDA_OK(cudaMemcpyAsync(features_host, e->tfeat, (size_t)pop * c.text_embed_dim * 4, cudaMemcpyDeviceToHost, s));
  TCUDA_OK(cudaStreamSynchronize(s));
  return GLASS_OK;
}

int64_t glass_text_launch_count(const glass_text_engine* e) { return e ? e->launches : 0; }


int glass_text_set_timing(glass_text_engine* e, int32_t enable) {
  if (!e) return tfail(GLASS_ERR_ARG, "null engine");
  TCUDA_OK(cudaSetDevice(e->cfg.device));
  if (enable && e->ev.empty()) {
    e->ev.resize(8192);
    for (auto& ev : e->ev) TCUDA_OK(cudaEventCreate(&ev));
  }
  e->timing = enable != 0;
  e->ev_used = 0;
  e->gemm_bytes = 0;
  return GLASS_OK;
}

int glass_text_gemm_time(glass_text_engine* e, float* ms, int32_t* launches, double* bytes) {
  if (!e) return tfail(GLASS_ERR_ARG, "null engine");
  TCUDA_OK(cudaSetDevice(e->cfg.device));
  TCUDA_OK(cudaDeviceSynchronize());
  float total = 0.f;
  for (size_t i = 0; i + 1 < e->ev_used; i += 2) {
    float t = 0.f;
    TCUDA_OK(cudaEventElapsedTime(&t, e->ev[i], e->ev[i + 1]));
    total += t;
  }
  if (ms) *ms = total;
  if (launches) *launches = (int32_t)(e->ev_used / 2);
  if (bytes) *bytes = e->gemm_bytes;
  e->ev_used = 0;
  e->gemm_bytes = 0;
  return GLASS_OK;
}

int glass_text_destroy(glass_text_engine* e) {
  if (!e) return GLASS_OK;
  cudaSetDevice(e->cfg.device);
  for (auto& kv : e->tensors) cudaFree(kv.second.ptr);
  if (e->arena) cudaFree(e->arena);
  for (auto& kv : e->gen_graphs) cudaGraphExecDestroy(kv.second);
  for (auto& kv : e->sim_graphs) cudaGraphExecDestroy(kv.second);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  for (auto& ev : e->ev) cudaEventDestroy(ev);
  delete e;
  return GLASS_OK;
}

}  // extern "C"
