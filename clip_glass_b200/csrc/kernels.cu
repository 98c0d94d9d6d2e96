// HBM-bound helper kernels of the fitness path: coalesced, vectorised where it
// matters, warp-shuffle reductions.  Everything GEMM-shaped lives in conv_tc.cu.
#include <type_traits>
#include "kernels.cuh"
#include "fir_tile.cuh"

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

// The same product for a whole population tile per block: out[b][n] for 32 candidates x 16 output columns, with the
// candidates' input rows (32 x K fp32) and the 16-column weight slice (K x 16) staged in shared memory ONCE.
// vecmat_kernel reads every weight through L2 once per candidate in 32 dependent batches (34 us per 512x512 mapping
// layer at P = 64, pure latency); here a weight is read once per 32 candidates.  Thread = (column n, 2 candidates); the
// four partial sums per output run over k = 0,4,8,.. / 1,5,9,.. / .. in ascending order and are combined as
// (a0 + a1) + (a2 + a3): exactly vecmat_kernel's summation order, so the results are bit-identical (K % 16 == 0).
constexpr int kVtCand = 32, kVtCols = 16, kVtPer = 2;       // 256 threads = 16 columns x 16 candidate pairs
__global__ void __launch_bounds__(256) vecmat_tile_kernel(const float* __restrict__ in, int in_stride,
                                                          const float* __restrict__ Wt, const float* __restrict__ bias,
                                                          float* __restrict__ out, int out_stride, int P, int K, int N,
                                                          int mode) {
  extern __shared__ float4 vt_smem[];
  float* xs = reinterpret_cast<float*>(vt_smem);            // [kVtCand][K]
  float* ws = xs + (size_t)kVtCand * K;                     // [K][kVtCols]
  const int n0 = blockIdx.x * kVtCols, b0 = blockIdx.y * kVtCand;
  const int nb = min(kVtCand, P - b0);
  // staging: eight 16-byte loads in flight per thread before the first store (issued one by one, every load exposed a
  // full L2 round trip: 32 dependent round trips per thread made the first version slower than vecmat_kernel)
  const int K4 = K >> 2;
  const bool wvec = (N & 3) == 0 && n0 + kVtCols <= N;      // whole 16-column rows, 16-byte aligned
  for (int i0 = threadIdx.x; i0 < K * (kVtCols / 4); i0 += 8 * blockDim.x) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x, k = i / (kVtCols / 4), c4 = i - k * (kVtCols / 4);
      v[u] = make_float4(0, 0, 0, 0);
      if (i < K * (kVtCols / 4)) {
        const float* wp = Wt + (size_t)k * N + n0 + 4 * c4;
        if (wvec) {
          v[u] = __ldg(reinterpret_cast<const float4*>(wp));
        } else {
          if (n0 + 4 * c4 + 0 < N) v[u].x = __ldg(wp + 0);
          if (n0 + 4 * c4 + 1 < N) v[u].y = __ldg(wp + 1);
          if (n0 + 4 * c4 + 2 < N) v[u].z = __ldg(wp + 2);
          if (n0 + 4 * c4 + 3 < N) v[u].w = __ldg(wp + 3);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < K * (kVtCols / 4)) reinterpret_cast<float4*>(ws)[i] = v[u];
    }
  }
  for (int i0 = threadIdx.x; i0 < kVtCand * K4; i0 += 8 * blockDim.x) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x, r = i / K4, k4 = i - r * K4;
      v[u] = (i < kVtCand * K4 && r < nb) ? __ldg(reinterpret_cast<const float4*>(in + (size_t)(b0 + r) * in_stride) + k4)
                                          : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < kVtCand * K4) reinterpret_cast<float4*>(xs)[i] = v[u];
    }
  }
  __syncthreads();
  const int c = threadIdx.x & (kVtCols - 1), cg = threadIdx.x / kVtCols;     // candidates kVtPer cg .. + kVtPer - 1
  float acc[kVtPer][4];
#pragma unroll
  for (int u = 0; u < kVtPer; ++u) acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f;
  const float4* x4 = reinterpret_cast<const float4*>(xs) + (size_t)(kVtPer * cg) * K4;
#pragma unroll 4
  for (int k4 = 0; k4 < K4; ++k4) {
    const float w0 = ws[(4 * k4 + 0) * kVtCols + c], w1 = ws[(4 * k4 + 1) * kVtCols + c];
    const float w2 = ws[(4 * k4 + 2) * kVtCols + c], w3 = ws[(4 * k4 + 3) * kVtCols + c];
#pragma unroll
    for (int u = 0; u < kVtPer; ++u) {
      const float4 x = x4[(size_t)u * K4 + k4];
      acc[u][0] = fmaf(x.x, w0, acc[u][0]);
      acc[u][1] = fmaf(x.y, w1, acc[u][1]);
      acc[u][2] = fmaf(x.z, w2, acc[u][2]);
      acc[u][3] = fmaf(x.w, w3, acc[u][3]);
    }
  }
  const int n = n0 + c;
  if (n >= N) return;
  const float bv = bias != nullptr ? bias[n] : 0.f;
#pragma unroll
  for (int u = 0; u < kVtPer; ++u) {
    const int b = b0 + kVtPer * cg + u;
    if (b >= P) break;
    float a = (acc[u][0] + acc[u][1]) + (acc[u][2] + acc[u][3]);
    if (bias != nullptr) a += bv;
    if (mode == 1) a = (a > 0.f ? a : 0.2f * a) * kSqrt2;
    out[(size_t)b * out_stride + n] = a;
  }
}

// Several independent vecmat jobs (the 17 demodulation GEMVs of one generator pass) in ONE launch: blockIdx.z = job.
__global__ void vecmat_batched_kernel(const VecmatBatch jobs, int in_stride, int mode) {
  extern __shared__ float row[];
  const VecmatJob& jb = jobs.job[blockIdx.z];
  const int K = jb.K, N = jb.N;
  if ((int)(blockIdx.x * blockDim.x) >= N) return;
  const int b = blockIdx.y;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float v = jb.in[(size_t)b * in_stride + k];
    row[k] = (mode == 2) ? v * v : v;
  }
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* __restrict__ Wt = jb.Wt;
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
  if (mode == 1) acc = (acc > 0.f ? acc : 0.2f * acc) * kSqrt2;
  if (mode == 2) acc = rsqrtf(acc + 1e-8f);
  if (jb.post_mul != nullptr) acc *= jb.post_mul[(size_t)b * jb.post_stride];
  jb.out[(size_t)b * N + n] = acc;
}

// one block per (layer, sample): m = 2^ceil(log2 max|s|) over the layer's slice; styles_n = s / m
__global__ void style_norm_kernel(const float* __restrict__ styles, float* __restrict__ styles_n,
                                  float* __restrict__ mscale, int S, const StyleSlices sl) {
  __shared__ float red[32];
  const int layer = blockIdx.x, b = blockIdx.y;
  const int off = sl.off[layer], cin = sl.cin[layer];
  const float* src = styles + (size_t)b * S + off;
  float mx = 0.f;
  for (int i = threadIdx.x; i < cin; i += blockDim.x) {
    const float a = fabsf(src[i]);
    mx = (a > mx && a <= 3.0e38f) ? a : mx;          // ignores NaN / inf (they propagate through s/m anyway)
  }
  mx = warp_max(mx);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (l == 0) red[w] = mx;
  __syncthreads();
  mx = 0.f;
  for (int i = 0; i < nw; ++i) mx = fmaxf(mx, red[i]);
  int ex = 0;
  float m = 1.f;
  if (mx > 0.f) {
    frexpf(mx, &ex);                                   // mx = f * 2^ex, f in [0.5, 1)  =>  |s| / 2^ex < 1
    m = ldexpf(1.f, ex);
  }
  const float inv = 1.f / m;                           // exact
  for (int i = threadIdx.x; i < cin; i += blockDim.x) styles_n[(size_t)b * S + off + i] = src[i] * inv;
  if (threadIdx.x == 0) mscale[(size_t)b * sl.n + layer] = m;
}

__global__ void range_scan_kernel(const __half* __restrict__ x, size_t n, unsigned long long* ctr) {
  unsigned long long bad = 0, sat = 0;
  float mx = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = fabsf(__half2float(x[i]));
    if (!(v <= 3.0e38f)) ++bad;                        // inf or NaN
    else {
      if (v >= 65504.f) ++sat;
      mx = fmaxf(mx, v);
    }
  }
  if (bad) atomicAdd(ctr + 0, bad);
  if (sat) atomicAdd(ctr + 1, sat);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(ctr + 2, (unsigned long long)__float_as_uint(mx));
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
      r = __fadd_rn(r, t.x); g = __fadd_rn(g, t.y); bl = __fadd_rn(bl, t.z);
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
      r = __fadd_rn(r, skip_up2(wy0, wy1, wx0, wx1, a.x, c.x, d.x, e.x));
      g = __fadd_rn(g, skip_up2(wy0, wy1, wx0, wx1, a.y, c.y, d.y, e.y));
      bl = __fadd_rn(bl, skip_up2(wy0, wy1, wx0, wx1, a.z, c.z, d.z, e.z));
    }
    if (yout != nullptr) yout[i] = make_float4(r, g, bl, 0.f);
    if (image != nullptr) {
      const size_t plane = (size_t)H * W;
      float* ip = image + (size_t)b * 3 * plane + (size_t)Y * W + X;
      ip[0] = image_value(r);
      ip[plane] = image_value(g);
      ip[2 * plane] = image_value(bl);
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
  for (int j = 0; j < 4; ++j) oh[j] = f2h2_sat(v[2 * j], v[2 * j + 1]);
  *reinterpret_cast<uint4*>(p) = o;
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

// One warp per row; a lane owns 8 contiguous channels of every 256-channel chunk (16-byte loads and stores; the
// per-element mapping of warp_layernorm_store costs 24 two-byte loads per lane at W = 768 and was latency-bound).
__global__ void layernorm_kernel(const __half* __restrict__ x, const float* __restrict__ lw, const float* __restrict__ lb,
                                 __half* __restrict__ out, int M, int W) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  constexpr int kChunks = kMaxPerLane / 8;       // rows up to 1024 wide
  float v[kChunks][8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kChunks; ++j) {
    const int i0 = j * 256 + lane * 8;
    if (i0 < W) {
      ld8(x + (size_t)row * W + i0, v[j]);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[j][u];
    }
  }
  const float mean = warp_sum(s) / (float)W;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < kChunks; ++j)
    if (j * 256 + lane * 8 < W) {
#pragma unroll
      for (int u = 0; u < 8; ++u) { const float d = v[j][u] - mean; q += d * d; }
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)W + 1e-5f);
#pragma unroll
  for (int j = 0; j < kChunks; ++j) {
    const int i0 = j * 256 + lane * 8;
    if (i0 < W) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(lw + i0)), w1 = __ldg(reinterpret_cast<const float4*>(lw + i0 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(lb + i0)), b1 = __ldg(reinterpret_cast<const float4*>(lb + i0 + 4));
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) o[u] = (v[j][u] - mean) * rstd * ww[u] + bb[u];
      st8(out + (size_t)row * W + i0, o);
    }
  }
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
    // four independent partial sums (the serial 768-long dependent chain was 0.3 ms of pure latency)
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    int i = 0;
    for (; i + 4 <= W; i += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) a4[u] = fmaf(c[i + u], __ldg(proj + (size_t)(i + u) * E + e), a4[u]);
    }
    for (; i < W; ++i) a4[0] = fmaf(c[i], __ldg(proj + (size_t)i * E + e), a4[0]);
    float acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
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
// image output path / BigGAN latent arithmetic
// ---------------------------------------------------------------------------
__global__ void image_grid_u8_kernel(const float* __restrict__ images, int n, int R, int xmaps, int padding,
                                     int Hg, int Wg, uint8_t* __restrict__ out) {
  const int cell = R + padding;
  const size_t total = (size_t)Hg * Wg;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % Wg), Y = (int)(i / Wg);
    const int gy = Y / cell, gx = X / cell;
    const int py = Y - gy * cell - padding, px = X - gx * cell - padding;
    const int k = gy * xmaps + gx;
    uint8_t v[3] = {0, 0, 0};
    if (py >= 0 && px >= 0 && gx < xmaps && k < n) {
      const float* src = images + (size_t)k * 3 * R * R + (size_t)py * R + px;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        // save_image: mul(255).add_(0.5).clamp_(0, 255).to(uint8)  (the cast truncates)
        const float t = fminf(fmaxf(__fadd_rn(__fmul_rn(src[(size_t)c * R * R], 255.f), 0.5f), 0.f), 255.f);
        v[c] = (uint8_t)t;
      }
    }
    out[i * 3 + 0] = v[0]; out[i * 3 + 1] = v[1]; out[i * 3 + 2] = v[2];
  }
}

__global__ void gather_images_kernel(const float4* __restrict__ images, const int* __restrict__ rows, size_t vec_per_image,
                                     float4* __restrict__ out) {
  const float4* src = images + (size_t)rows[blockIdx.y] * vec_per_image;
  float4* dst = out + (size_t)blockIdx.y * vec_per_image;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < vec_per_image; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __ldg(src + i);
}

// one block per candidate: clip the first dz genes, softmax (fp32, like torch.softmax on a float tensor) over the rest
__global__ void biggan_latent_kernel(const double* __restrict__ x, float* __restrict__ z, float* __restrict__ cls, int dz,
                                     int ncls) {
  __shared__ float red[32];
  const double* xi = x + (size_t)blockIdx.x * (dz + ncls);
  for (int i = threadIdx.x; i < dz; i += blockDim.x)
    z[(size_t)blockIdx.x * dz + i] = fminf(fmaxf((float)xi[i], -2.f), 2.f);
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < ncls; i += blockDim.x) mx = fmaxf(mx, (float)xi[dz + i]);
  mx = warp_max(mx);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (l == 0) red[w] = mx;
  __syncthreads();
  mx = -INFINITY;
  for (int i = 0; i < nw; ++i) mx = fmaxf(mx, red[i]);
  float sum = 0.f;
  for (int i = threadIdx.x; i < ncls; i += blockDim.x) sum += expf((float)xi[dz + i] - mx);
  sum = block_sum(sum, red);
  for (int i = threadIdx.x; i < ncls; i += blockDim.x)
    cls[(size_t)blockIdx.x * ncls + i] = expf((float)xi[dz + i] - mx) / sum;
}

// ---------------------------------------------------------------------------
// discriminator
// ---------------------------------------------------------------------------
template <int C>
__global__ void from_rgb_kernel(const float* __restrict__ images, const float* __restrict__ Wt,
                                const float* __restrict__ bias, __half* __restrict__ out, int P, int R, int out_i8) {
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
      h2[j] = f2h2_sat(a[0], a[1]);
    }
    if (out_i8) {
      const size_t y = pix / R, x = pix - y * R;
      *reinterpret_cast<uint4*>(out + ((((b * R + y) * G + g) * R) + x) * 8) = pk;
    } else {
      *reinterpret_cast<uint4*>(out + pixi * C + g * 8) = pk;
    }
  }
}

// Stride-2 FIR passes of the discriminator (shared pieces: fir_tile.cuh).  Each thread of fir_down_compute owns one
// output column, 8 channels and two vertically adjacent outputs.
__device__ __forceinline__ void fir_down_store(const uint4* tile, __half* __restrict__ out, int b, int ty, int tx, int Ho,
                                               int Wo, int C, int c0) {
  uint4 o[2];
  fir_down_compute(tile, o);
  const int ox = threadIdx.x & 15, g = (threadIdx.x >> 4) & 3, oyp = threadIdx.x >> 6;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int zy = ty * kFdTH + 2 * oyp + k, zx = tx * kFdTW + ox;
    if (zy < Ho && zx < Wo) *reinterpret_cast<uint4*>(out + (((size_t)b * Ho + zy) * Wo + zx) * C + c0 + g * 8) = o[k];
  }
}

template <bool kI8>
__global__ void __launch_bounds__(256, GLASS_FIR_MINB) fir_down_kernel(const __half* __restrict__ x, __half* __restrict__ out, int N,
                                                       int H, int W, int C) {
  __shared__ uint4 tile[kFdUnits];
  const int Ho = H >> 1, Wo = W >> 1;
  int b, ty, tx;
  fd_decode_tile(blockIdx.x, Ho, Wo, b, ty, tx);
  const int c0 = blockIdx.y * kFdC;
  fd_stage_tile<kI8>(tile, x, b, H, W, C, c0, 2 * ty * kFdTH - 1, 2 * tx * kFdTW - 1);
  __syncthreads();
  fir_down_store(tile, out, b, ty, tx, Ho, Wo, C, c0);
}

// fromRGB fused with the first block's projection FIR (fir_tile.cuh: fd_stage_tile_from_rgb): the 32-channel 1024^2
// tensor is never read back for the FIR.
__global__ void __launch_bounds__(256, GLASS_FIR_MINB) from_rgb_fir_kernel(const float* __restrict__ images,
                                                           const __grid_constant__ FrgbConsts k,
                                                           __half* __restrict__ xout, __half* __restrict__ down, int P,
                                                           int R, int C, int c0, int out_i8) {
  __shared__ uint4 tile[kFdUnits];
  const int Ho = R >> 1, Wo = R >> 1;
  int b, ty, tx;
  fd_decode_tile(blockIdx.x, Ho, Wo, b, ty, tx);
  fd_stage_tile_from_rgb(tile, images, k, xout, b, R, C, c0, out_i8, 2 * ty * kFdTH - 1, 2 * tx * kFdTW - 1);
  __syncthreads();
  fir_down_store(tile, down, b, ty, tx, Ho, Wo, C, c0);
}

// ---------------------------------------------------------------------------
// exact polyphase helpers (streaming, HBM-bound)
// ---------------------------------------------------------------------------
// The FIR [1,3,3,1] is the binomial [1,1]*[1,1]*[1,1]: three cascaded adjacent sums per axis (3 adds per output
// instead of 4 multiply-adds), with only three running vectors of vertical state.  Each thread produces two
// adjacent output columns from five loaded columns.  The first level of adjacent sums is taken in fp16 (one
// rounding of <= 1/2 ulp per sum) so that only four vectors per row need converting; everything else is fp32.
__device__ __forceinline__ void adj_sum8(const uint4& p, const uint4& q, float (&v)[8]) {
  const __half2* a = reinterpret_cast<const __half2*>(&p);
  const __half2* b = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __half22float2(__hadd2(a[j], b[j]));
    v[2 * j] = t.x;
    v[2 * j + 1] = t.y;
  }
}
// five columns c[0..4] (x0-1 .. x0+3) -> unnormalised horizontal FIR at x0 and x0+1: h = [1,3,3,1] . c
__device__ __forceinline__ void hfir2(const uint4 (&c)[5], float (&h0)[8], float (&h1)[8]) {
  float p0[8], p1[8], p2[8], p3[8];
  adj_sum8(c[0], c[1], p0);
  adj_sum8(c[1], c[2], p1);
  adj_sum8(c[2], c[3], p2);
  adj_sum8(c[3], c[4], p3);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float s0 = p0[j] + p1[j], s1 = p1[j] + p2[j], s2 = p2[j] + p3[j];
    h0[j] = s0 + s1;
    h1[j] = s1 + s2;
  }
}

// k_upfir: one thread = 8 channels x 2 output columns x kUpRows output rows.  v[Z][X] = sum f f u[Z+jy-1][X+jx-1]
// with f = [1,3,3,1]/4; then + noise + bias, lrelu*sqrt2, * next style.
constexpr int kUpRows = 8;
// Resident blocks per SM.  At one block (the compiler takes 254 registers to hoist the loads of all 11 rows) k_upfir ran
// latency-bound at 12 % occupancy: 545 us; at two blocks (128 registers, 8 B spilled) 364 us at P = 64.  Three or four
// blocks spill heavily (80 / 64 registers) and are slower again (profiles/r02_ab_streaming_occupancy.log).
#ifndef GLASS_POLY_MINB
#define GLASS_POLY_MINB 2
#endif
__global__ void __launch_bounds__(256, GLASS_POLY_MINB) upfir_kernel(
    const __half* __restrict__ u, __half* __restrict__ out, const float* __restrict__ noise, size_t noise_group_stride,
    int noise_group_div, const float* __restrict__ noise_strength, const float* __restrict__ bias,
    const float* __restrict__ out_scale, int out_scale_stride, int P, int Hout, int Wout, int C) {
  const int C8 = C >> 3, Hq = Hout / kUpRows, Wp = Wout >> 1;
  const int Hu = Hout + 2, Wu = Wout + 2;
  const size_t n = (size_t)P * Hq * Wp * C8;
  const float nstr = noise != nullptr ? __ldg(noise_strength) : 0.f;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int xp = (int)((i / C8) % Wp);
    const int zq = (int)((i / ((size_t)C8 * Wp)) % Hq);
    const int b = (int)(i / ((size_t)C8 * Wp * Hq));
    const int Z0 = zq * kUpRows, X0 = 2 * xp;
    float bs[8], sc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bs[j] = __ldg(bias + c8 * 8 + j);
      sc[j] = kSqrt2 * __ldg(out_scale + (size_t)b * out_scale_stride + c8 * 8 + j);
    }
    const float* nz = noise != nullptr ? noise + (size_t)(b / noise_group_div) * noise_group_stride : nullptr;
    const __half* ub = u + ((size_t)b * Hu * Wu) * C + c8 * 8;
    // columns X0-1 .. X0+3: X0-1 >= 0 unless X0 == 0; X0+3 <= Wout+1 = Wu-1 always
    float sa[2][8], sb[2][8], sc3[2][8];      // vertical cascade state: previous h, previous p, previous s
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int j = 0; j < 8; ++j) sa[x][j] = sb[x][j] = sc3[x][j] = 0.f;
#pragma unroll
    for (int r = 0; r < kUpRows + 3; ++r) {
      const int Y = Z0 + r - 1;
      float h[2][8];
      if (Y >= 0 && Y < Hu) {
        const __half* rp = ub + ((size_t)Y * Wu + X0) * C;
        uint4 c[5];
        c[0] = X0 > 0 ? __ldg(reinterpret_cast<const uint4*>(rp - C)) : zero4;
#pragma unroll
        for (int q = 1; q < 5; ++q) c[q] = __ldg(reinterpret_cast<const uint4*>(rp + (size_t)(q - 1) * C));
        hfir2(c, h[0], h[1]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) h[0][j] = h[1][j] = 0.f;
      }
      const int kdone = r - 3;                 // after this row, output row Z0 + kdone is complete
      const int Z = Z0 + kdone;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        float v[8];
        const float nv = (kdone >= 0 && nz != nullptr) ? nstr * __ldg(nz + (size_t)Z * Wout + X0 + x) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float p = sa[x][j] + h[x][j];
          const float q = sb[x][j] + p;
          const float t = sc3[x][j] + q;        // = h[r-3] + 3 h[r-2] + 3 h[r-1] + h[r]
          sa[x][j] = h[x][j];
          sb[x][j] = p;
          sc3[x][j] = q;
          float o = fmaf(t, 1.f / 16.f, bs[j]) + nv;
          o = fmaxf(o, 0.2f * o);
          v[j] = o * sc[j];
        }
        if (kdone >= 0) st8(out + (((size_t)b * Hout + Z) * Wout + X0 + x) * C + c8 * 8, v);
      }
    }
  }
}

// k_blur_s2d: one thread = 8 channels x one space-to-depth cell column (2 output columns) x kBlurCells cells
// vertically (2*kBlurCells output rows).  u[Y][X] = sum f f a[Y+jy-2][X+jx-2], f = [1,3,3,1]/8, for Y,X in [0,H];
// cell (z,w), phase (py,px) holds u[2z+py][2w+px]; positions beyond H are written as zeros.
constexpr int kBlurCells = 4;
__global__ void __launch_bounds__(256, GLASS_POLY_MINB) blur_s2d_kernel(const __half* __restrict__ a, __half* __restrict__ out, int P,
                                                       int H, int W, int C) {
  const int C8 = C >> 3, Hs = (H >> 1) + 1, Ws = (W >> 1) + 1, Hp = (Hs + kBlurCells - 1) / kBlurCells;
  const size_t n = (size_t)P * Hp * Ws * C8;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int w = (int)((i / C8) % Ws);
    const int zp = (int)((i / ((size_t)C8 * Ws)) % Hp);
    const int b = (int)(i / ((size_t)C8 * Ws * Hp));
    const int z0 = zp * kBlurCells;              // output rows Y = 2*z0 .. 2*z0 + 2*kBlurCells - 1
    const __half* ab = a + ((size_t)b * H * W) * C + c8 * 8;
    const int x0 = 2 * w - 2;                    // input columns x0 .. x0+4
    float sa[2][8], sb[2][8], sc3[2][8];
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int j = 0; j < 8; ++j) sa[x][j] = sb[x][j] = sc3[x][j] = 0.f;
#pragma unroll
    for (int r = 0; r < 2 * kBlurCells + 3; ++r) {
      const int yy = 2 * z0 + r - 2;             // input row; output row 2*z0 + (r-3) completes with it
      float h[2][8];
      if (yy >= 0 && yy < H) {
        const __half* rp = ab + ((size_t)yy * W) * C;
        uint4 c[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const int xx = x0 + q;
          c[q] = (xx >= 0 && xx < W) ? __ldg(reinterpret_cast<const uint4*>(rp + (size_t)xx * C)) : zero4;
        }
        hfir2(c, h[0], h[1]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) h[0][j] = h[1][j] = 0.f;
      }
      const int kdone = r - 3;
      const int Y = 2 * z0 + kdone;
      const int z = Y >> 1, py = Y & 1;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        float v[8];
        const bool inside = (Y <= H) && (2 * w + x <= W);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float p = sa[x][j] + h[x][j];
          const float q = sb[x][j] + p;
          const float t = sc3[x][j] + q;
          sa[x][j] = h[x][j];
          sb[x][j] = p;
          sc3[x][j] = q;
          v[j] = inside ? t * (1.f / 64.f) : 0.f;
        }
        if (kdone >= 0 && z < Hs)
          st8(out + (((size_t)b * Hs + z) * Ws + w) * (4 * C) + (py * 2 + x) * C + c8 * 8, v);
      }
    }
  }
}

// k_blur_s2d, shared-memory-tiled version (the product path; blur_s2d_kernel above is kept as the fp32 cross-check,
// GLASS_FLAG_FP32_BLUR).  The streaming version reads its 5 x 11 neighbourhood straight from global memory with a
// serial dependence per row and runs latency-bound at 3.1 TB/s; here a block stages the 19 x 35-pixel x 32-channel
// patch once (every 16-byte load issued up front), then each thread filters one column strip out of shared memory
// with the packed-half2 separable FIR of downconv_tc.cu ((c0+c3) + 3(c1+c2) per axis, 1/64 applied in the vertical
// pass).  Tile layout [row][column][4 groups]: the 8 lanes of an LDS.128 phase (4 groups x 2 columns) read 128
// contiguous bytes, and a pixel's 4 groups are stored as 64 contiguous bytes (two full sectors).
constexpr int kBtRows = 16, kBtCols = 32, kBtRawR = kBtRows + 3, kBtRawC = kBtCols + 3, kBtPitch = 36;
__device__ __forceinline__ uint4 fir4_h2(const uint4& c0, const uint4& c1, const uint4& c2, const uint4& c3, bool scaled) {
  const __half2 k1 = __floats2half2_rn(1.f / 64.f, 1.f / 64.f), k3 = __floats2half2_rn(3.f / 64.f, 3.f / 64.f);
  const __half2 three = __floats2half2_rn(3.f, 3.f);
  uint4 r;
  const __half2* a = reinterpret_cast<const __half2*>(&c0);
  const __half2* b = reinterpret_cast<const __half2*>(&c1);
  const __half2* c = reinterpret_cast<const __half2*>(&c2);
  const __half2* d = reinterpret_cast<const __half2*>(&c3);
  __half2* o = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    o[j] = scaled ? __hfma2(__hadd2(b[j], c[j]), k3, __hmul2(__hadd2(a[j], d[j]), k1))
                  : __hfma2(__hadd2(b[j], c[j]), three, __hadd2(a[j], d[j]));
  return r;
}
#ifndef GLASS_BLUR_MINB
#define GLASS_BLUR_MINB 4      // 64 registers, no spills; 3 -> 4 resident blocks: 1.016 -> 0.997 ms over the four launches
#endif
__global__ void __launch_bounds__(256, GLASS_BLUR_MINB) blur_s2d_tile_kernel(const __half* __restrict__ a, __half* __restrict__ out, int H,
                                                               int W, int C, int tiles_x, int tiles_y) {
  __shared__ uint4 tile[kBtRawR * kBtPitch * 4];
  const int Hs = (H >> 1) + 1, Ws = (W >> 1) + 1;
  // channel chunk fastest: the C/32 blocks of one pixel tile run together, so that whole NHWC pixel rows are consumed
  // while they are in L2 (with the chunk in blockIdx.y every chunk pass re-read the tensor from DRAM: 2x the bytes)
  const int nchunks = C / 32;
  int t = blockIdx.x / nchunks;
  const int c0 = (blockIdx.x - t * nchunks) * 32;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int Y0 = ty * kBtRows, X0 = tx * kBtCols;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  constexpr int kItems = kBtRawR * kBtRawC * 4;                    // (pixel, group) 16-byte pieces
  constexpr int kPer = (kItems + 255) / 256;
  uint4 v[kPer];
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int i = threadIdx.x + k * 256;
    const int g = i & 3, pixel = i >> 2;
    const int r = pixel / kBtRawC, c = pixel - r * kBtRawC;
    const int yy = Y0 - 2 + r, xx = X0 - 2 + c;
    v[k] = zero4;
    if (i < kItems && yy >= 0 && yy < H && xx >= 0 && xx < W)
      v[k] = __ldg(reinterpret_cast<const uint4*>(a + (((size_t)b * H + yy) * W + xx) * C + c0 + g * 8));
  }
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < kItems) {
      const int g = i & 3, pixel = i >> 2;
      const int r = pixel / kBtRawC, c = pixel - r * kBtRawC;
      tile[(r * kBtPitch + c) * 4 + g] = v[k];
    }
  }
  __syncthreads();
  const int g = threadIdx.x & 3, j = (threadIdx.x >> 2) & 31, strip = threadIdx.x >> 7;
  const int i0 = 8 * strip;
  const int X = X0 + j;
  uint4 hw[4];
#pragma unroll
  for (int r = 0; r < 11; ++r) {
    const uint4* rp = tile + ((i0 + r) * kBtPitch + j) * 4 + g;
    const uint4 cur = fir4_h2(rp[0], rp[4], rp[8], rp[12], false);
    hw[r & 3] = cur;
    if (r >= 3) {
      const int Y = Y0 + i0 + r - 3;
      if (Y < 2 * Hs && X < 2 * Ws) {
        uint4 u = fir4_h2(hw[(r - 3) & 3], hw[(r - 2) & 3], hw[(r - 1) & 3], cur, true);
        if (Y > H || X > W) u = zero4;                             // beyond the blurred (H+1) x (W+1) grid
        *reinterpret_cast<uint4*>(out + (((size_t)b * Hs + (Y >> 1)) * Ws + (X >> 1)) * (4 * C) +
                                  ((Y & 1) * 2 + (X & 1)) * C + c0 + g * 8) = u;
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
  // population tiles (bit-identical to vecmat_kernel, see vecmat_tile_kernel) whenever the staged rows fit
  const size_t tile_smem = ((size_t)kVtCand * K + (size_t)K * kVtCols) * sizeof(float);
  if (mode != 2 && K % 16 == 0 && in_stride % 4 == 0 && tile_smem <= 110 * 1024 && P >= 8) {
    static bool configured = false;
    if (!configured) {
      cudaError_t err = cudaFuncSetAttribute(vecmat_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
      if (err != cudaSuccess) return err;
      configured = true;
    }
    dim3 grid((N + kVtCols - 1) / kVtCols, (P + kVtCand - 1) / kVtCand);
    vecmat_tile_kernel<<<grid, 256, tile_smem, s>>>(in, in_stride, Wt, bias, out, out_stride, P, K, N, mode);
    GLASS_RET();
  }
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
cudaError_t k_vecmat_batched(const VecmatBatch& jobs, int in_stride, int P, int mode, cudaStream_t s) {
  if (jobs.n <= 0 || jobs.n > kMaxVecmatJobs) return cudaErrorInvalidValue;
  int maxN = 0, maxK = 0;
  for (int i = 0; i < jobs.n; ++i) { maxN = jobs.job[i].N > maxN ? jobs.job[i].N : maxN; maxK = jobs.job[i].K > maxK ? jobs.job[i].K : maxK; }
  vecmat_batched_kernel<<<dim3((maxN + 127) / 128, P, jobs.n), 128, maxK * sizeof(float), s>>>(jobs, in_stride, mode);
  GLASS_RET();
}
cudaError_t k_style_norm(const float* styles, float* styles_n, float* mscale, int S, int P, const StyleSlices& sl,
                         cudaStream_t s) {
  if (sl.n <= 0 || sl.n > kMaxStyleSlices) return cudaErrorInvalidValue;
  style_norm_kernel<<<dim3(sl.n, P), 128, 0, s>>>(styles, styles_n, mscale, S, sl);
  GLASS_RET();
}
cudaError_t k_range_scan(const __half* x, size_t n, unsigned long long* ctr, cudaStream_t s) {
  range_scan_kernel<<<blocks_for(n), kThreads, 0, s>>>(x, n, ctr);
  GLASS_RET();
}
cudaError_t k_image_grid_u8(const float* images, int n, int R, int nrow, int padding, uint8_t* out, cudaStream_t s) {
  if (n <= 0 || nrow <= 0 || padding < 0) return cudaErrorInvalidValue;
  const int xmaps = n < nrow ? n : nrow, ymaps = (n + xmaps - 1) / xmaps;
  const int Hg = (R + padding) * ymaps + padding, Wg = (R + padding) * xmaps + padding;
  image_grid_u8_kernel<<<blocks_for((size_t)Hg * Wg), kThreads, 0, s>>>(images, n, R, xmaps, padding, Hg, Wg, out);
  GLASS_RET();
}
cudaError_t k_gather_images(const float* images, const int* rows, int n, size_t image_elems, float* out, cudaStream_t s) {
  if (n <= 0 || image_elems % 4 != 0) return cudaErrorInvalidValue;
  const size_t vec = image_elems / 4;
  gather_images_kernel<<<dim3(blocks_for(vec, kThreads, 148 * 4), n), kThreads, 0, s>>>(
      reinterpret_cast<const float4*>(images), rows, vec, reinterpret_cast<float4*>(out));
  GLASS_RET();
}
cudaError_t k_biggan_latent(const double* x, float* z, float* cls, int P, int dz, int ncls, cudaStream_t s) {
  if (P <= 0 || dz <= 0 || ncls <= 0) return cudaErrorInvalidValue;
  biggan_latent_kernel<<<P, 256, 0, s>>>(x, z, cls, dz, ncls);
  GLASS_RET();
}
cudaError_t k_layernorm(const __half* x, const float* w, const float* b, __half* out, int M, int W, cudaStream_t s) {
  if (W > 32 * kMaxPerLane || W % 8 != 0) return cudaErrorInvalidValue;
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
                       int out_i8, cudaStream_t s) {
  const int blocks = blocks_for((size_t)P * R * R * (C / 8), kThreads, 148 * 32);
  if (C == 32) from_rgb_kernel<32><<<blocks, kThreads, 0, s>>>(images, Wt, bias, out, P, R, out_i8);
  else if (C == 64) from_rgb_kernel<64><<<blocks, kThreads, 0, s>>>(images, Wt, bias, out, P, R, out_i8);
  else if (C == 128) from_rgb_kernel<128><<<blocks, kThreads, 0, s>>>(images, Wt, bias, out, P, R, out_i8);
  else return cudaErrorInvalidValue;
  GLASS_RET();
}
cudaError_t k_from_rgb_fir(const float* images, const float* folded_host, __half* xout, __half* down, int P, int R,
                           int C, int out_i8, cudaStream_t s) {
  if (C % kFdC != 0 || (R & 1)) return cudaErrorInvalidValue;
  const int Ho = R / 2;
  const int tiles = ((Ho + kFdTW - 1) / kFdTW) * ((Ho + kFdTH - 1) / kFdTH) * P;
  for (int c0 = 0; c0 < C; c0 += kFdC) {          // one launch per 32-channel chunk (ffhq-f: exactly one)
    FrgbConsts k;
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < kFdC; ++c) k.w[r][c] = folded_host[r * C + c0 + c];
    from_rgb_fir_kernel<<<tiles, 256, 0, s>>>(images, k, xout, down, P, R, C, c0, out_i8);
  }
  GLASS_RET();
}
cudaError_t k_fir_down(const __half* x, __half* out, int N, int H, int W, int C, int in_i8, cudaStream_t s) {
  if (C % kFdC != 0) return cudaErrorInvalidValue;
  const int Ho = H / 2, Wo = W / 2;
  const int tiles = ((Wo + kFdTW - 1) / kFdTW) * ((Ho + kFdTH - 1) / kFdTH) * N;
  dim3 grid(tiles, C / kFdC);
  if (in_i8) fir_down_kernel<true><<<grid, 256, 0, s>>>(x, out, N, H, W, C);
  else fir_down_kernel<false><<<grid, 256, 0, s>>>(x, out, N, H, W, C);
  GLASS_RET();
}
cudaError_t k_upfir(const __half* u, __half* out, const float* noise, size_t noise_group_stride, int noise_group_div,
                    const float* noise_strength, const float* bias, const float* out_scale, int out_scale_stride,
                    int P, int Hout, int Wout, int C, cudaStream_t s) {
  if (C % 8 != 0 || Hout % kUpRows != 0 || (Wout & 1)) return cudaErrorInvalidValue;
  const size_t n = (size_t)P * (Hout / kUpRows) * (Wout / 2) * (C / 8);
  upfir_kernel<<<blocks_for(n, kThreads, 148 * 32), kThreads, 0, s>>>(u, out, noise, noise_group_stride, noise_group_div,
                                                                     noise_strength, bias, out_scale, out_scale_stride,
                                                                     P, Hout, Wout, C);
  GLASS_RET();
}
cudaError_t k_blur_s2d(const __half* a, __half* out, int P, int H, int W, int C, cudaStream_t s, int fp32_variant) {
  if (C % 8 != 0 || (H & 1) || (W & 1)) return cudaErrorInvalidValue;
  if (!fp32_variant && C % 32 == 0) {
    const int tiles_x = (W + 2 + kBtCols - 1) / kBtCols, tiles_y = (H + 2 + kBtRows - 1) / kBtRows;
    blur_s2d_tile_kernel<<<P * tiles_x * tiles_y * (C / 32), 256, 0, s>>>(a, out, H, W, C, tiles_x, tiles_y);
    GLASS_RET();
  }
  const size_t n = (size_t)P * ((H / 2 + 1 + kBlurCells - 1) / kBlurCells) * (W / 2 + 1) * (C / 8);
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
