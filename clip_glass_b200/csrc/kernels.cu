// HBM-bound helper kernels of the fitness path: coalesced, vectorised where it
// matters, warp-shuffle reductions.  Everything GEMM-shaped lives in conv_tc.cu.
#include "kernels.cuh"

namespace glass {

namespace {

constexpr int kThreads = 256;
inline int blocks_for(size_t n, int threads = kThreads, int cap = 148 * 16) {
  size_t b = (n + threads - 1) / threads;
  if (b > (size_t)cap) b = cap;
  if (b == 0) b = 1;
  return (int)b;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ float rh(float v) { return __half2float(__float2half_rn(v)); }

// ---------------------------------------------------------------------------
__global__ void latents_to_f32_kernel(const double* x, float* z, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    z[i] = (float)x[i];
}

__global__ void pixelnorm_kernel(const float* z, float* out, int L) {
  __shared__ float red[32];
  const float* zi = z + (size_t)blockIdx.x * L;
  float s = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) s += zi[i] * zi[i];
  s = block_sum(s, red);
  const float r = rsqrtf(s / (float)L + 1e-8f);
  for (int i = threadIdx.x; i < L; i += blockDim.x) out[(size_t)blockIdx.x * L + i] = zi[i] * r;
}

__global__ void vecmat_kernel(const float* __restrict__ in, int in_stride, const float* __restrict__ Wt,
                              const float* __restrict__ bias, float* __restrict__ out, int out_stride, int K, int N,
                              int mode) {
  extern __shared__ float row[];
  const int b = blockIdx.y;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float v = in[(size_t)b * in_stride + k];
    row[k] = (mode == 2) ? v * v : v;
  }
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  // four independent accumulators and 16 loads in flight per thread: the loop is L2-latency-bound otherwise
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int k = 0;
  for (; k + 16 <= K; k += 16) {
    float w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = __ldg(Wt + (size_t)(k + j) * N + n);
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      a0 = fmaf(row[k + j], w[j], a0);
      a1 = fmaf(row[k + j + 1], w[j + 1], a1);
      a2 = fmaf(row[k + j + 2], w[j + 2], a2);
      a3 = fmaf(row[k + j + 3], w[j + 3], a3);
    }
  }
  for (; k < K; ++k) a0 = fmaf(row[k], __ldg(Wt + (size_t)k * N + n), a0);
  float acc = (a0 + a1) + (a2 + a3);
  if (bias != nullptr) acc += bias[n];
  if (mode == 1) acc = (acc > 0.f ? acc : 0.2f * acc) * kSqrt2;
  if (mode == 2) acc = rsqrtf(acc + 1e-8f);
  out[(size_t)b * out_stride + n] = acc;
}

__global__ void const_input_kernel(const float* cst, const float* styles, int stride, __half* out, int P, int C) {
  const size_t n = (size_t)P * 16 * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int pix = (int)((i / C) % 16);
    const int b = (int)(i / ((size_t)16 * C));
    out[i] = __float2half_rn(cst[pix * C + c] * styles[(size_t)b * stride + c]);
  }
}

__global__ void rgb_weights_kernel(const float* W, const float* styles, int stride, float* out, int P, int C) {
  const size_t n = (size_t)P * 3 * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % C);
    const int c = (int)((i / C) % 3);
    const int b = (int)(i / ((size_t)3 * C));
    out[i] = W[c * C + o] * styles[(size_t)b * stride + o];
  }
}

__global__ void rgb_combine_kernel(const float4* __restrict__ yprev, const float4* __restrict__ slabs, int n_slabs,
                                   const float* __restrict__ bias, float4* __restrict__ yout, float* __restrict__ image,
                                   int P, int H, int W) {
  const size_t n = (size_t)P * H * W;
  const float b0 = bias[0], b1 = bias[1], b2 = bias[2];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % W);
    const int Y = (int)((i / W) % H);
    const int b = (int)(i / ((size_t)W * H));
    float r = b0, g = b1, bl = b2;
    for (int s = 0; s < n_slabs; ++s) {
      const float4 t = __ldg(slabs + (size_t)s * n + i);
      r += t.x; g += t.y; bl += t.z;
    }
    if (yprev != nullptr) {
      // v[2z] = .75 x[z-1] + .25 x[z];  v[2z+1] = .25 x[z-1] + .75 x[z];  x[-1] = 0  (per axis)
      const int Hp = H >> 1, Wp = W >> 1;
      const int zy = Y >> 1, zx = X >> 1;
      const float wy0 = (Y & 1) ? 0.25f : 0.75f, wy1 = 1.f - wy0;   // weights of rows zy-1, zy
      const float wx0 = (X & 1) ? 0.25f : 0.75f, wx1 = 1.f - wx0;
      const float4* base = yprev + (size_t)b * Hp * Wp;
      float4 a = make_float4(0, 0, 0, 0), c = a, d = a;
      const float4 e = __ldg(base + (size_t)zy * Wp + zx);
      if (zy > 0 && zx > 0) a = __ldg(base + (size_t)(zy - 1) * Wp + zx - 1);
      if (zy > 0) c = __ldg(base + (size_t)(zy - 1) * Wp + zx);
      if (zx > 0) d = __ldg(base + (size_t)zy * Wp + zx - 1);
      r += wy0 * (wx0 * a.x + wx1 * c.x) + wy1 * (wx0 * d.x + wx1 * e.x);
      g += wy0 * (wx0 * a.y + wx1 * c.y) + wy1 * (wx0 * d.y + wx1 * e.y);
      bl += wy0 * (wx0 * a.z + wx1 * c.z) + wy1 * (wx0 * d.z + wx1 * e.z);
    }
    if (yout != nullptr) yout[i] = make_float4(r, g, bl, 0.f);
    if (image != nullptr) {
      const size_t plane = (size_t)H * W;
      float* ip = image + (size_t)b * 3 * plane + (size_t)Y * W + X;
      ip[0] = fminf(fmaxf((r + 1.f) * 0.5f, 0.f), 1.f);
      ip[plane] = fminf(fmaxf((g + 1.f) * 0.5f, 0.f), 1.f);
      ip[2 * plane] = fminf(fmaxf((bl + 1.f) * 0.5f, 0.f), 1.f);
    }
  }
}

// Philox4x32-10
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__global__ void noise_kernel(float* out, size_t n, uint64_t seed, uint64_t offset) {
  const size_t quads = (n + 3) / 4;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (size_t)gridDim.x * blockDim.x) {
    const uint64_t ctr = offset + q;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x9E3779B9u, 0u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    const float u0 = ((float)c[0] + 0.5f) * 2.3283064365386963e-10f;
    const float u1 = ((float)c[1] + 0.5f) * 2.3283064365386963e-10f;
    const float u2 = ((float)c[2] + 0.5f) * 2.3283064365386963e-10f;
    const float u3 = ((float)c[3] + 0.5f) * 2.3283064365386963e-10f;
    const float r0 = sqrtf(-2.f * __logf(u0)), r1 = sqrtf(-2.f * __logf(u2));
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * u1, &s0, &c0);
    __sincosf(6.283185307179586f * u3, &s1, &c1);
    const float v[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
    for (int j = 0; j < 4; ++j)
      if (q * 4 + j < n) out[q * 4 + j] = v[j];
  }
}

// ---------------------------------------------------------------------------
// CLIP
// ---------------------------------------------------------------------------
__global__ void resize_patches_kernel(const float* __restrict__ images, __half* __restrict__ patches, int P, int Rin,
                                      int Rout, int patch) {
  const int g = Rout / patch;
  const int kdim = 3 * patch * patch;
  const float scale = (float)Rin / (float)Rout;
  const size_t n = (size_t)P * 3 * Rout * Rout;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Rout);
    const int oy = (int)((i / Rout) % Rout);
    const int c = (int)((i / ((size_t)Rout * Rout)) % 3);
    const int b = (int)(i / ((size_t)3 * Rout * Rout));
    float sy = scale * (oy + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
    float sx = scale * (ox + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
    const int y0 = (int)sy, x0 = (int)sx;
    const int yp = (y0 < Rin - 1) ? 1 : 0, xp = (x0 < Rin - 1) ? 1 : 0;
    const float ly = sy - y0, lx = sx - x0;
    const float* src = images + ((size_t)b * 3 + c) * Rin * Rin;
    const float v00 = __ldg(src + (size_t)y0 * Rin + x0), v01 = __ldg(src + (size_t)y0 * Rin + x0 + xp);
    const float v10 = __ldg(src + (size_t)(y0 + yp) * Rin + x0), v11 = __ldg(src + (size_t)(y0 + yp) * Rin + x0 + xp);
    const float v = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
    const int gy = oy / patch, py = oy - gy * patch, gx = ox / patch, px = ox - gx * patch;
    patches[((size_t)b * g * g + gy * g + gx) * kdim + c * patch * patch + py * patch + px] = __float2half_rn(v);
  }
}

constexpr int kMaxPerLane = 32;   // rows up to 1024 wide

__device__ __forceinline__ void warp_layernorm_store(float (&v)[kMaxPerLane], int W, const float* lw, const float* lb,
                                                     __half* out) {
  const int lane = threadIdx.x & 31;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxPerLane; ++j)
    if (lane + 32 * j < W) s += v[j];
  const float mean = warp_sum(s) / (float)W;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxPerLane; ++j)
    if (lane + 32 * j < W) { const float d = v[j] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / (float)W + 1e-5f);
#pragma unroll
  for (int j = 0; j < kMaxPerLane; ++j) {
    const int i = lane + 32 * j;
    if (i < W) out[i] = __float2half_rn((v[j] - mean) * rstd * lw[i] + lb[i]);
  }
}

__global__ void embed_lnpre_kernel(const __half* __restrict__ emb, const float* cls, const float* pos, const float* lw,
                                   const float* lb, __half* __restrict__ tokens, int P, int T, int W) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= P * T) return;
  const int lane = threadIdx.x & 31;
  const int b = row / T, t = row - b * T;
  float v[kMaxPerLane];
#pragma unroll
  for (int j = 0; j < kMaxPerLane; ++j) {
    const int i = lane + 32 * j;
    v[j] = 0.f;
    if (i < W) {
      const float base = (t == 0) ? rh(cls[i]) : __half2float(emb[((size_t)b * (T - 1) + t - 1) * W + i]);
      v[j] = rh(base + rh(pos[(size_t)t * W + i]));
    }
  }
  warp_layernorm_store(v, W, lw, lb, tokens + (size_t)row * W);
}

__global__ void layernorm_kernel(const __half* __restrict__ x, const float* lw, const float* lb,
                                 __half* __restrict__ out, int M, int W) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  float v[kMaxPerLane];
#pragma unroll
  for (int j = 0; j < kMaxPerLane; ++j) {
    const int i = lane + 32 * j;
    v[j] = (i < W) ? __half2float(x[(size_t)row * W + i]) : 0.f;
  }
  warp_layernorm_store(v, W, lw, lb, out + (size_t)row * W);
}

// one block per (image, head); T <= 64 tokens, head dim 64
constexpr int kHd = 64;
__global__ void attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int T, int W) {
  extern __shared__ float sm[];
  float* q = sm;                         // [T][65]
  float* k = q + T * (kHd + 1);
  float* v = k + T * (kHd + 1);
  float* sc = v + T * (kHd + 1);         // [T][T+1]
  const int heads = W / kHd;
  const int b = blockIdx.x / heads, hd = blockIdx.x - b * heads;
  const __half* base = qkv + (size_t)b * T * 3 * W + hd * kHd;
  for (int i = threadIdx.x; i < T * kHd; i += blockDim.x) {
    const int t = i / kHd, d = i - t * kHd;
    const __half* r = base + (size_t)t * 3 * W + d;
    q[t * (kHd + 1) + d] = __half2float(r[0]) * 0.125f;
    k[t * (kHd + 1) + d] = __half2float(r[W]);
    v[t * (kHd + 1) + d] = __half2float(r[2 * W]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    const int a = i / T, c = i - a * T;
    float acc = 0.f;
#pragma unroll 16
    for (int d = 0; d < kHd; ++d) acc = fmaf(q[a * (kHd + 1) + d], k[c * (kHd + 1) + d], acc);
    sc[a * (T + 1) + c] = acc;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int a = warp; a < T; a += nw) {
    float m = -INFINITY;
    for (int c = lane; c < T; c += 32) m = fmaxf(m, sc[a * (T + 1) + c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < T; c += 32) {
      const float e = __expf(sc[a * (T + 1) + c] - m);
      sc[a * (T + 1) + c] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.f / s;
    for (int c = lane; c < T; c += 32) sc[a * (T + 1) + c] *= inv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * kHd; i += blockDim.x) {
    const int a = i / kHd, d = i - a * kHd;
    float acc = 0.f;
    for (int c = 0; c < T; ++c) acc = fmaf(sc[a * (T + 1) + c], v[c * (kHd + 1) + d], acc);
    out[((size_t)b * T + a) * W + hd * kHd + d] = __float2half_rn(acc);
  }
}

__global__ void final_cosine_kernel(const __half* __restrict__ tokens, const float* lw, const float* lb,
                                    const float* __restrict__ proj, const float* __restrict__ text, float* features,
                                    float* sim, float* neg_sim, int T, int W, int E) {
  extern __shared__ float sm[];
  float* c = sm;          // [W] ln_post(cls), fp16-rounded
  __shared__ float red[32];
  const int b = blockIdx.x;
  const __half* x = tokens + (size_t)b * T * W;   // token 0 = class token
  float s = 0.f;
  for (int i = threadIdx.x; i < W; i += blockDim.x) { c[i] = __half2float(x[i]); s += c[i]; }
  const float mean = block_sum(s, red) / (float)W;
  float qv = 0.f;
  for (int i = threadIdx.x; i < W; i += blockDim.x) { const float d = c[i] - mean; qv += d * d; }
  const float rstd = rsqrtf(block_sum(qv, red) / (float)W + 1e-5f);
  for (int i = threadIdx.x; i < W; i += blockDim.x) c[i] = rh((c[i] - mean) * rstd * lw[i] + lb[i]);
  __syncthreads();
  float dot = 0.f, nf = 0.f, nt = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < W; ++i) acc = fmaf(c[i], __ldg(proj + (size_t)i * E + e), acc);
    acc = rh(acc);
    if (features != nullptr) features[(size_t)b * E + e] = acc;
    const float t = text[e];
    dot += acc * t; nf += acc * acc; nt += t * t;
  }
  dot = block_sum(dot, red);
  nf = block_sum(nf, red);
  nt = block_sum(nt, red);
  if (threadIdx.x == 0) {
    const float v = dot / fmaxf(sqrtf(nf) * sqrtf(nt), 1e-8f);
    if (sim != nullptr) sim[b] = v;
    if (neg_sim != nullptr) neg_sim[b] = -v;
  }
}

// ---------------------------------------------------------------------------
// discriminator
// ---------------------------------------------------------------------------
template <int C>
__global__ void from_rgb_kernel(const float* __restrict__ images, const float* __restrict__ Wt,
                                const float* __restrict__ bias, __half* __restrict__ out, int P, int R) {
  __shared__ float w[3 * C + C];
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) w[i] = Wt[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) w[3 * C + i] = bias[i];
  __syncthreads();
  constexpr int G = C / 8;                       // 8-channel groups per pixel; consecutive threads -> consecutive 16 B
  const size_t plane = (size_t)R * R;
  const size_t n = (size_t)P * plane * G;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    const size_t pixi = i / G;
    const size_t b = pixi / plane, pix = pixi - b * plane;
    const float* ip = images + b * 3 * plane + pix;
    const float r = __ldg(ip) * 2.f - 1.f, gg = __ldg(ip + plane) * 2.f - 1.f, bl = __ldg(ip + 2 * plane) * 2.f - 1.f;
    uint4 pk;
    __half2* h2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = g * 8 + 2 * j + u;
        const float t = fmaf(r, w[c], fmaf(gg, w[C + c], fmaf(bl, w[2 * C + c], w[3 * C + c])));
        a[u] = fmaxf(t, 0.2f * t) * kSqrt2;
      }
      h2[j] = __floats2half2_rn(a[0], a[1]);
    }
    *reinterpret_cast<uint4*>(out + pixi * C + g * 8) = pk;
  }
}

// FIR (pad 1) sampled at stride 2.  One block = 8x16 outputs x 32 channels: the (18 x 34)-pixel input patch is
// staged once in shared memory (coalesced 16-byte loads, pixel pitch padded to 80 B against bank conflicts), so
// HBM/L2 see every input byte once and the 16 taps per output come from shared memory.
constexpr int kFdTH = 8, kFdTW = 16, kFdC = 32, kFdPitch = 40;   // pitch in halfs (80 bytes)
__global__ void __launch_bounds__(256) fir_down_kernel(const __half* __restrict__ x, __half* __restrict__ out, int N,
                                                       int H, int W, int C) {
  __shared__ __align__(16) __half tile[(2 * kFdTH + 2) * (2 * kFdTW + 2) * kFdPitch];
  const int Ho = H >> 1, Wo = W >> 1;
  const int tiles_x = (Wo + kFdTW - 1) / kFdTW, tiles_y = (Ho + kFdTH - 1) / kFdTH;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int c0 = blockIdx.y * kFdC;
  const int iy0 = 2 * ty * kFdTH - 1, ix0 = 2 * tx * kFdTW - 1;
  constexpr int IW = 2 * kFdTW + 2, IH = 2 * kFdTH + 2;
  for (int i = threadIdx.x; i < IH * IW * 4; i += blockDim.x) {
    const int g = i & 3, pix = i >> 2;
    const int py = pix / IW, px = pix - py * IW;
    const int yy = iy0 + py, xx = ix0 + px;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W)
      v = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)b * H + yy) * W + xx) * C + c0 + g * 8));
    *reinterpret_cast<uint4*>(tile + pix * kFdPitch + g * 8) = v;
  }
  __syncthreads();
  const float f[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  const int g = threadIdx.x & 3;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int op = (threadIdx.x >> 2) + 64 * pass;
    const int oy = op / kFdTW, ox = op - oy * kFdTW;
    const int zy = ty * kFdTH + oy, zx = tx * kFdTW + ox;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int jy = 0; jy < 4; ++jy) {
#pragma unroll
      for (int jx = 0; jx < 4; ++jx) {
        const uint4 v = *reinterpret_cast<const uint4*>(tile + ((2 * oy + jy) * IW + 2 * ox + jx) * kFdPitch + g * 8);
        const __half2* h2 = reinterpret_cast<const __half2*>(&v);
        const float wgt = f[jy] * f[jx];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 tt = __half22float2(h2[j]);
          acc[2 * j] = fmaf(wgt, tt.x, acc[2 * j]);
          acc[2 * j + 1] = fmaf(wgt, tt.y, acc[2 * j + 1]);
        }
      }
    }
    if (zy < Ho && zx < Wo) {
      uint4 o;
      __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
      *reinterpret_cast<uint4*>(out + (((size_t)b * Ho + zy) * Wo + zx) * C + c0 + g * 8) = o;
    }
  }
}

// ---------------------------------------------------------------------------
// exact polyphase helpers (streaming, HBM-bound)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void ld8(const __half* p, float (&v)[8]) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h2 = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __half22float2(h2[j]);
    v[2 * j] = t.x;
    v[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ void st8(__half* p, const float (&v)[8]) {
  uint4 o;
  __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
  *reinterpret_cast<uint4*>(p) = o;
}

// fp16 pair sums (one rounding of <= 1/2 ulp each), everything after that in fp32: ~30 % fewer instructions than
// converting every tap, without letting fp16 accumulate the filter.
__device__ __forceinline__ void pair_sum8(const uint4& p, const uint4& q, float (&v)[8]) {
  const __half2* a = reinterpret_cast<const __half2*>(&p);
  const __half2* b = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __half22float2(__hadd2(a[j], b[j]));
    v[2 * j] = t.x;
    v[2 * j + 1] = t.y;
  }
}

// One thread: 8 channels x kUpRows consecutive output rows of one output column.  Input rows stream through a
// sliding window of four row accumulators: row Y = Z0+r-1 is filtered horizontally once,
//   h = 1/4 (u[X-1] + u[X+2]) + 3/4 (u[X] + u[X+1]),
// and added to the (up to four) outputs Z = Y-2 .. Y+1 with the vertical taps; a finished output gets noise, bias,
// lrelu*sqrt2 and the next layer's style and is stored as fp16.
constexpr int kUpRows = 8;
__global__ void __launch_bounds__(256) upfir_kernel(
    const __half* __restrict__ u, __half* __restrict__ out, const float* __restrict__ noise, size_t noise_group_stride,
    int noise_group_div, const float* __restrict__ noise_strength, const float* __restrict__ bias,
    const float* __restrict__ out_scale, int out_scale_stride, int P, int Hout, int Wout, int C) {
  const int C8 = C >> 3, Hq = Hout / kUpRows;
  const int Hu = Hout + 2, Wu = Wout + 2;
  const size_t n = (size_t)P * Hq * Wout * C8;
  const float f[4] = {0.25f, 0.75f, 0.75f, 0.25f};
  const float nstr = noise != nullptr ? __ldg(noise_strength) : 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int X = (int)((i / C8) % Wout);
    const int zq = (int)((i / ((size_t)C8 * Wout)) % Hq);
    const int b = (int)(i / ((size_t)C8 * Wout * Hq));
    const int Z0 = zq * kUpRows;
    float bs[8], sc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bs[j] = __ldg(bias + c8 * 8 + j);
      sc[j] = __ldg(out_scale + (size_t)b * out_scale_stride + c8 * 8 + j);
    }
    const float* nz = noise != nullptr ? noise + (size_t)(b / noise_group_div) * noise_group_stride : nullptr;
    const __half* ub = u + ((size_t)b * Hu * Wu) * C + c8 * 8;
    // X-1 >= 0 always holds except at X == 0; X+2 <= Wout+1 = Wu-1 always holds
    const bool left = X > 0;
    float acc[kUpRows][8];
#pragma unroll
    for (int k = 0; k < kUpRows; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
#pragma unroll
    for (int r = 0; r < kUpRows + 3; ++r) {
      const int Y = Z0 + r - 1;
      if (Y >= 0 && Y < Hu) {
        const __half* rp = ub + ((size_t)Y * Wu + X) * C;
        const uint4 zero = make_uint4(0, 0, 0, 0);
        const uint4 ua = left ? __ldg(reinterpret_cast<const uint4*>(rp - C)) : zero;
        const uint4 ubv = __ldg(reinterpret_cast<const uint4*>(rp));
        const uint4 uc = __ldg(reinterpret_cast<const uint4*>(rp + C));
        const uint4 ud = __ldg(reinterpret_cast<const uint4*>(rp + 2 * C));
        float so[8], si[8];
        pair_sum8(ua, ud, so);
        pair_sum8(ubv, uc, si);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float hv = fmaf(0.75f, si[j], 0.25f * so[j]);
          // row r feeds output k = r - jv with vertical tap jv (v[Z] = sum_jv f[jv] u[Z + jv - 1])
#pragma unroll
          for (int jv = 0; jv < 4; ++jv) {
            const int k = r - jv;
            if (k >= 0 && k < kUpRows) acc[k][j] = fmaf(f[jv], hv, acc[k][j]);
          }
        }
      }
      const int kdone = r - 3;       // output row whose last input row was just added
      if (kdone >= 0) {
        const int Z = Z0 + kdone;
        const float nv = nz != nullptr ? nstr * __ldg(nz + (size_t)Z * Wout + X) : 0.f;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = acc[kdone][j] + nv + bs[j];
          t = fmaxf(t, 0.2f * t) * kSqrt2;
          v[j] = t * sc[j];
        }
        st8(out + (((size_t)b * Hout + Z) * Wout + X) * C + c8 * 8, v);
      }
    }
  }
}

// One thread: 8 channels of two vertically adjacent space-to-depth cells (8 outputs) from a 7x5 input patch.
// u[Y][X] = sum_{jy,jx} f[jy] f[jx] a[Y+jy-2][X+jx-2];  cell (z,w) holds u[2z+py][2w+px];  f = [1,3,3,1]/8.
__global__ void __launch_bounds__(256) blur_s2d_kernel(const __half* __restrict__ a, __half* __restrict__ out, int P,
                                                       int H, int W, int C) {
  const int C8 = C >> 3, Hs = (H >> 1) + 1, Ws = (W >> 1) + 1, Hp = (Hs + 1) >> 1;
  const size_t n = (size_t)P * Hp * Ws * C8;
  const float f[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int w = (int)((i / C8) % Ws);
    const int zp = (int)((i / ((size_t)C8 * Ws)) % Hp);
    const int b = (int)(i / ((size_t)C8 * Ws * Hp));
    const int z0 = 2 * zp;                       // cells z0 and z0+1: output rows Y = 2*z0 .. 2*z0+3
    const __half* ab = a + ((size_t)b * H * W) * C + c8 * 8;
    float acc[4][2][8];                          // [output row 0..3][px][channel]
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int px = 0; px < 2; ++px)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[k][px][j] = 0.f;
    const int x0 = 2 * w - 2;                    // input columns x0 .. x0+4
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      const int yy = 2 * z0 + r - 2;             // input row; feeds output row k = r - jy
      if (yy < 0 || yy >= H) continue;
      const __half* rp = ab + ((size_t)yy * W) * C;
      uint4 c[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const int xx = x0 + q;
        c[q] = (xx >= 0 && xx < W) ? __ldg(reinterpret_cast<const uint4*>(rp + (size_t)xx * C)) : make_uint4(0, 0, 0, 0);
      }
      // px = 0: taps on columns 0..3; px = 1: taps on columns 1..4
      float o0[8], i0[8], o1[8], i1[8];
      pair_sum8(c[0], c[3], o0);
      pair_sum8(c[1], c[2], i0);
      pair_sum8(c[1], c[4], o1);
      pair_sum8(c[2], c[3], i1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float h0 = fmaf(0.375f, i0[j], 0.125f * o0[j]);
        const float h1 = fmaf(0.375f, i1[j], 0.125f * o1[j]);
#pragma unroll
        for (int jy = 0; jy < 4; ++jy) {
          const int k = r - jy;
          if (k >= 0 && k < 4) {
            acc[k][0][j] = fmaf(f[jy], h0, acc[k][0][j]);
            acc[k][1][j] = fmaf(f[jy], h1, acc[k][1][j]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int z = z0 + (k >> 1), py = k & 1;
      if (z >= Hs) continue;
      __half* op = out + (((size_t)b * Hs + z) * Ws + w) * (4 * C) + c8 * 8;
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        const bool inside = (2 * z + py <= H) && (2 * w + px <= W);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = inside ? acc[k][px][j] : 0.f;
        st8(op + (py * 2 + px) * C, v);
      }
    }
  }
}

// one block per (minibatch, member m in [0, batch/group)); samples mb*batch + gi*(batch/group) + m
__global__ void mbstd_kernel(const __half* __restrict__ x, __half* __restrict__ out, int batch, int group, int C,
                             int Cpad) {
  __shared__ float red[32];
  const int per = batch / group;
  const int mb = blockIdx.x / per, m = blockIdx.x - mb * per;
  const int n = 16 * C;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int pix = i / C, c = i - pix * C;
    float vals[8];
    float mean = 0.f;
    for (int gi = 0; gi < group; ++gi) {
      const size_t s = (size_t)mb * batch + (size_t)gi * per + m;
      vals[gi] = __half2float(x[(s * 16 + pix) * C + c]);
      mean += vals[gi];
    }
    mean /= (float)group;
    float var = 0.f;
    for (int gi = 0; gi < group; ++gi) {
      const size_t s = (size_t)mb * batch + (size_t)gi * per + m;
      const float d = vals[gi] - mean;
      var += d * d;
      out[(s * 16 + pix) * Cpad + c] = __float2half_rn(d);
    }
    acc += sqrtf(var / (float)group + 1e-8f);
  }
  const float feat = block_sum(acc, red) / (float)n;
  for (int i = threadIdx.x; i < group * 16 * (Cpad - C); i += blockDim.x) {
    const int cc = i % (Cpad - C);
    const int pix = (i / (Cpad - C)) % 16;
    const int gi = i / ((Cpad - C) * 16);
    const size_t s = (size_t)mb * batch + (size_t)gi * per + m;
    out[(s * 16 + pix) * Cpad + C + cc] = __float2half_rn(cc == 0 ? feat : 0.f);
  }
}

__global__ void dense1_hinge_kernel(const __half* __restrict__ x, const float* w, const float* b, float* logits,
                                    float* hinge, int P, int C) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= P) return;
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int i = lane; i < C; i += 32) acc = fmaf(__half2float(x[(size_t)row * C + i]), w[i], acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    const float l = acc + b[0];
    if (logits != nullptr) logits[row] = l;
    if (hinge != nullptr) hinge[row] = fmaxf(1.f - l, 0.f);
  }
}

}  // namespace

#define GLASS_RET() return cudaGetLastError()

cudaError_t k_latents_to_f32(const double* x, float* z, size_t n, cudaStream_t s) {
  latents_to_f32_kernel<<<blocks_for(n), kThreads, 0, s>>>(x, z, n);
  GLASS_RET();
}
cudaError_t k_pixelnorm(const float* z, float* out, int P, int L, cudaStream_t s) {
  pixelnorm_kernel<<<P, 256, 0, s>>>(z, out, L);
  GLASS_RET();
}
cudaError_t k_vecmat(const float* in, int in_stride, const float* Wt, const float* bias, float* out, int out_stride,
                     int P, int K, int N, int mode, cudaStream_t s) {
  dim3 grid((N + 127) / 128, P);
  vecmat_kernel<<<grid, 128, K * sizeof(float), s>>>(in, in_stride, Wt, bias, out, out_stride, K, N, mode);
  GLASS_RET();
}
cudaError_t k_const_input(const float* cst, const float* styles, int stride, __half* out, int P, int C,
                          cudaStream_t s) {
  const_input_kernel<<<blocks_for((size_t)P * 16 * C), kThreads, 0, s>>>(cst, styles, stride, out, P, C);
  GLASS_RET();
}
cudaError_t k_rgb_weights(const float* W, const float* styles, int stride, float* out, int P, int C, cudaStream_t s) {
  rgb_weights_kernel<<<blocks_for((size_t)P * 3 * C), kThreads, 0, s>>>(W, styles, stride, out, P, C);
  GLASS_RET();
}
cudaError_t k_rgb_combine(const float4* yprev, const float4* slabs, int n_slabs, const float* bias, float4* yout,
                          float* image, int P, int H, int W, cudaStream_t s) {
  rgb_combine_kernel<<<blocks_for((size_t)P * H * W), kThreads, 0, s>>>(yprev, slabs, n_slabs, bias, yout, image, P, H,
                                                                       W);
  GLASS_RET();
}
cudaError_t k_noise(float* out, size_t n, uint64_t seed, uint64_t offset, cudaStream_t s) {
  noise_kernel<<<blocks_for((n + 3) / 4), kThreads, 0, s>>>(out, n, seed, offset);
  GLASS_RET();
}
cudaError_t k_resize_patches(const float* images, __half* patches, int P, int Rin, int Rout, int patch,
                             cudaStream_t s) {
  resize_patches_kernel<<<blocks_for((size_t)P * 3 * Rout * Rout), kThreads, 0, s>>>(images, patches, P, Rin, Rout,
                                                                                    patch);
  GLASS_RET();
}
cudaError_t k_embed_lnpre(const __half* patch_emb, const float* cls, const float* pos, const float* lw, const float* lb,
                          __half* tokens, int P, int T, int W, cudaStream_t s) {
  if (W > 32 * kMaxPerLane) return cudaErrorInvalidValue;
  const int rows = P * T;
  embed_lnpre_kernel<<<(rows + 7) / 8, 256, 0, s>>>(patch_emb, cls, pos, lw, lb, tokens, P, T, W);
  GLASS_RET();
}
cudaError_t k_layernorm(const __half* x, const float* w, const float* b, __half* out, int M, int W, cudaStream_t s) {
  if (W > 32 * kMaxPerLane) return cudaErrorInvalidValue;
  layernorm_kernel<<<(M + 7) / 8, 256, 0, s>>>(x, w, b, out, M, W);
  GLASS_RET();
}
cudaError_t k_attention(const __half* qkv, __half* out, int P, int T, int W, cudaStream_t s) {
  if (T > 64 || W % kHd != 0) return cudaErrorInvalidValue;
  const size_t smem = sizeof(float) * ((size_t)3 * T * (kHd + 1) + (size_t)T * (T + 1));
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (err != cudaSuccess) return err;
    configured = true;
  }
  attention_kernel<<<P * (W / kHd), 128, smem, s>>>(qkv, out, T, W);
  GLASS_RET();
}
cudaError_t k_final_cosine(const __half* tokens, const float* lw, const float* lb, const float* proj,
                           const float* text, float* features, float* sim, float* neg_sim, int P, int T, int W, int E,
                           cudaStream_t s) {
  final_cosine_kernel<<<P, 256, W * sizeof(float), s>>>(tokens, lw, lb, proj, text, features, sim, neg_sim, T, W, E);
  GLASS_RET();
}
cudaError_t k_from_rgb(const float* images, const float* Wt, const float* bias, __half* out, int P, int R, int C,
                       cudaStream_t s) {
  const int blocks = blocks_for((size_t)P * R * R * (C / 8), kThreads, 148 * 32);
  if (C == 32) from_rgb_kernel<32><<<blocks, kThreads, 0, s>>>(images, Wt, bias, out, P, R);
  else if (C == 64) from_rgb_kernel<64><<<blocks, kThreads, 0, s>>>(images, Wt, bias, out, P, R);
  else if (C == 128) from_rgb_kernel<128><<<blocks, kThreads, 0, s>>>(images, Wt, bias, out, P, R);
  else return cudaErrorInvalidValue;
  GLASS_RET();
}
cudaError_t k_fir_down(const __half* x, __half* out, int N, int H, int W, int C, cudaStream_t s) {
  if (C % kFdC != 0) return cudaErrorInvalidValue;
  const int Ho = H / 2, Wo = W / 2;
  const int tiles = ((Wo + kFdTW - 1) / kFdTW) * ((Ho + kFdTH - 1) / kFdTH) * N;
  dim3 grid(tiles, C / kFdC);
  fir_down_kernel<<<grid, 256, 0, s>>>(x, out, N, H, W, C);
  GLASS_RET();
}
cudaError_t k_upfir(const __half* u, __half* out, const float* noise, size_t noise_group_stride, int noise_group_div,
                    const float* noise_strength, const float* bias, const float* out_scale, int out_scale_stride,
                    int P, int Hout, int Wout, int C, cudaStream_t s) {
  if (C % 8 != 0 || Hout % kUpRows != 0) return cudaErrorInvalidValue;
  const size_t n = (size_t)P * (Hout / kUpRows) * Wout * (C / 8);
  upfir_kernel<<<blocks_for(n, kThreads, 148 * 32), kThreads, 0, s>>>(u, out, noise, noise_group_stride, noise_group_div,
                                                                     noise_strength, bias, out_scale, out_scale_stride,
                                                                     P, Hout, Wout, C);
  GLASS_RET();
}
cudaError_t k_blur_s2d(const __half* a, __half* out, int P, int H, int W, int C, cudaStream_t s) {
  if (C % 8 != 0 || (H & 1) || (W & 1)) return cudaErrorInvalidValue;
  const size_t n = (size_t)P * ((H / 2 + 2) / 2) * (W / 2 + 1) * (C / 8);
  blur_s2d_kernel<<<blocks_for(n, kThreads, 148 * 32), kThreads, 0, s>>>(a, out, P, H, W, C);
  GLASS_RET();
}
cudaError_t k_mbstd(const __half* x, __half* out, int P, int batch, int group, int C, int Cpad, cudaStream_t s) {
  if (group > 8 || batch % group != 0 || P % batch != 0) return cudaErrorInvalidValue;
  const int blocks = (P / batch) * (batch / group);
  mbstd_kernel<<<blocks, 256, 0, s>>>(x, out, batch, group, C, Cpad);
  GLASS_RET();
}
cudaError_t k_dense1_hinge(const __half* x, const float* w, const float* b, float* logits, float* hinge, int P, int C,
                           cudaStream_t s) {
  dense1_hinge_kernel<<<(P + 7) / 8, 256, 0, s>>>(x, w, b, logits, hinge, P, C);
  GLASS_RET();
}

}  // namespace glass
