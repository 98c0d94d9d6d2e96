// Shared declarations for the clip-glass-b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace glass {

constexpr float kSqrt2 = 1.4142135623730951f;
constexpr float kInvSqrt2 = 0.7071067811865476f;

// A/B and work-skipping knobs (GLASS_DEBUG_* environment variables, ConvParams::debug_skip) exist only in builds
// made with -DGLASS_DEBUG.  The product library never reads the environment: nothing outside glass_config can
// change what its kernels compute.
#ifdef GLASS_DEBUG
constexpr bool kDebugBuild = true;
inline const char* debug_env(const char* name) { return getenv(name); }
#else
constexpr bool kDebugBuild = false;
inline const char* debug_env(const char*) { return nullptr; }
#endif

// fp32 pair -> fp16x2, round-to-nearest, SATURATING to +-65504 instead of overflowing to inf (one F2FP.SATFINITE
// instruction, the same cost as the plain conversion).  Every fp16 activation store goes through this: the
// reference runs G/D in fp32, so an out-of-range value must degrade to a clamp, never to inf -> NaN scores.
__device__ __forceinline__ __half2 f2h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return *reinterpret_cast<__half2*>(&r);
}

// Packed fp32 pairs (sm_100: FFMA2 / FMUL2 / FADD2 — two IEEE fp32 operations per issued instruction, each lane
// rounded exactly like the scalar instruction).  The epilogues of the 32/64-channel layers are bound by instruction
// issue and per-warp latency, not by the FMA pipe: half the math instructions per element is what these buy.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// lrelu on a pair: max(t, 0.2 t) per lane (no packed max exists: FMUL2 + 2 FMNMX)
__device__ __forceinline__ f32x2 lrelu2(f32x2 t) {
  float a, b, c, d;
  upk2(t, a, b);
  upk2(mul2(t, pk2(0.2f, 0.2f)), c, d);
  return pk2(fmaxf(a, c), fmaxf(b, d));
}

// Skip-path x2 upsample of one colour (stylegan2/modules.py:580-602 as polyphase weights) and the final
// clip((y + 1) / 2, 0, 1) (utils.py:14-17).  Written with explicit roundings (the compiler may neither contract nor
// split them), so that k_rgb_combine and the image-finishing conv epilogue produce the same bits by construction.
__device__ __forceinline__ float skip_up2(float wy0, float wy1, float wx0, float wx1, float a, float c, float d, float e) {
  const float top = __fmaf_rn(wx1, c, __fmul_rn(wx0, a));
  const float bot = __fmaf_rn(wx1, e, __fmul_rn(wx0, d));
  return __fmaf_rn(wy1, bot, __fmul_rn(wy0, top));
}
__device__ __forceinline__ float image_value(float y) {
  return fminf(fmaxf(__fmul_rn(__fadd_rn(y, 1.f), 0.5f), 0.f), 1.f);
}

// ---------------------------------------------------------------------------
// Implicit-GEMM convolution / GEMM description.
//
// Accumulator space: M = Nimg*H*W "pixels" (NHWC activations, fp16), N = Ntot
// output columns, K = taps*Cin.  Tap t = ky*3+kx reads the input at offset
// (ky-1, kx-1) with zero fill outside the image (taps==1: plain GEMM / 1x1).
// A tile of 128 accumulator rows is a box TN x TH x TW of pixels.
// ---------------------------------------------------------------------------
enum StoreMode : int {
  kStoreRegular = 0,       // out[pix][Ntot]
  kStoreDepthToSpace = 1,  // Ntot = 4*Cout, column (py*2+px)*Cout+o -> out[2y+py][2x+px][o]   (G up-conv)
  kStoreSpaceToDepth = 2,  // out[y/2][x/2][(y&1)*2+(x&1)][Ntot]                               (D conv0 -> conv1 input)
  kStoreSpaceToDepthY = 3  // pixel-pair rows (x_phases == 2): out[y/2][x][(y&1)][Ntot], i.e. space-to-depth of the
                           // underlying full-width image, because a row already holds the two x phases
};
enum Act : int { kActNone = 0, kActLrelu = 1, kActQuickGelu = 2 };

struct EpiParams {
  const float* dmod;        // [Nimg][Cout] demodulation coefficients, or null
  const float* bias;        // [Cout] or null
  const float* noise;       // [groups][Hout*Wout] or null
  const float* noise_strength;  // device scalar
  int noise_group_div;      // images per noise group (config.batch_size)
  size_t noise_group_stride;  // floats between consecutive groups (all noise layers of one group)
  int act;
  int round_fp16_before_act;  // mimic the reference's fp16 op boundaries (CLIP)
  const float* rgb_w;       // [Nimg][3][Cout] per-sample toRGB weights (W*s) or null
  float4* rgb_out;          // [n_tiles][Nimg*H*W] partial toRGB sums
  // Last generator conv, when its n-tile holds all channels (one toRGB slab): the skip-sum / x2 upsample / bias /
  // biggan_norm of k_rgb_combine happen right here and the image is written from the epilogue (no float4 slab round
  // trip, one launch less).  image != null enables it; img_yprev = the previous block's skip sum [Nimg][H/2][W/2].
  float* image;             // [Nimg][3][H][W] fp32 in [0,1], or null
  const float4* img_yprev;
  const float* img_bias;    // toRGB bias [3]
  const float* out_scale;   // next layer's style s[img*out_scale_stride + o], or null
  int out_scale_stride;
  const __half* residual;   // [pix][Ntot] (regular layout) or null
  int res_i8;               // the residual tensor is channel-group-interleaved: [n][y][Ntot/8][x][8]
  float post_scale;         // applied after the residual add
  __half* out;              // null => nothing stored (G's last conv only feeds toRGB)
  int store_mode;
  int Cout;                 // per-phase channel count (== Ntot unless depth-to-space)
  int out_i8;               // store channel-group-interleaved: out[n][y][c/8][x][8] (regular and depth-to-space)
  int x_phases;             // 1, or 2 for pixel-pair rows: columns [k*Cout, (k+1)*Cout) belong to pixel 2x+k
                            // (requires BN == Ntot == 2*Cout, so that an epilogue warp's column half is one pixel)
  int cout_shift;           // log2(Cout) when Cout is a power of two, else -1
  int noise_div_shift;      // log2(noise_group_div) when it is a power of two, else -1
};

struct ConvParams {
  int Nimg, H, W;           // accumulator grid
  int TN, TH, TW;           // tile box, TN*TH*TW == 128
  int tiles_n, tiles_y, tiles_x;
  int in_H, in_W;           // input tensor extent (== H, W except for the exact polyphase forms)
  int Cin;                  // K per tap (multiple of BK)
  int taps;                 // number of filter taps (9, 4 or 1)
  signed char tap_dy[9], tap_dx[9];   // input offset of each tap (3x3: ky-1, kx-1)
  int Ntot;                 // multiple of BN
  int BN, BK;
  int debug_skip;           // bring-up aid (env GLASS_DEBUG_SKIP): bit 0 = epilogue only drains TMEM (results
                            // invalid, timing experiments); bit 1 = MODE 4 with LBO/SBO roles swapped
  int all_valid;            // H % TH == 0 && W % TW == 0 && Nimg % TN == 0: no row of any tile is out of range
  int pow2, sh_n, sh_x, sh_y;  // tile grid is a power of two in every dimension: decode with shifts
  int mode;                 // 0 = streamed taps, 1 = resident taps + halo copies (conv_tc.cu)
  // Structural zeros of the exact polyphase forms (packing.exact_upconv / exact_downconv: 9 of the 16 (tap, phase)
  // weight blocks are non-zero).  MODE 0 skips the zero blocks in the producer and the MMA loop alike:
  //   1 = up-conv: the n-tile lies inside ONE output phase (BN <= skip_ch = Cout); tap (dy,dx) feeds phase (py,px)
  //       only if (dy == 0 || py == 0) && (dx == 0 || px == 0);
  //   2 = down-conv: a K chunk lies inside ONE input phase (BK <= skip_ch = channels per phase); tap (a,b) reads
  //       phase (py,px) only if !(a && py) && !(b && px).
  int skip_mode, skip_ch;
  int rot_div;              // skip_mode 1: n_tile is rotated by m / rot_div so that a CTA's successive tiles cycle
                            // through the phases (phase 0 has four taps, phase 3 one)
  const __half* in;         // [Nimg][H][W][Cin]   (SIMT bring-up path; the TC path reads through TMA)
  const __half* wgt;        // [taps][Ntot][Cin]
  EpiParams epi;
};

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == kActLrelu) {
    v = (v > 0.f ? v : 0.2f * v) * kSqrt2;
  } else if (act == kActQuickGelu) {
    v = __fdividef(v, 1.f + __expf(-1.702f * v));    // (as epilogue_fastN: approximate reciprocal)
  }
  return v;
}

// Epilogue for one accumulator row (pixel) and 16 consecutive columns
// [n0, n0+16).  Shared verbatim by the tcgen05 kernel and the SIMT bring-up
// kernel so that everything after the accumulator is tested once.
__device__ __forceinline__ void epilogue_row16(const ConvParams& p, int img, int y, int x, int n0,
                                               float (&v)[16], float (&rgb)[3]) {
  const EpiParams& e = p.epi;
  const int Cout = e.Cout;
  const int phase = n0 / Cout;
  const int o0 = n0 - phase * Cout;
  int yo = y, xo = x, Ho = p.H, Wo = p.W;
  if (e.store_mode == kStoreDepthToSpace) {
    yo = 2 * y + (phase >> 1);
    xo = 2 * x + (phase & 1);
    Ho = 2 * p.H;
    Wo = 2 * p.W;
  }
  float nz = 0.f;
  if (e.noise != nullptr) {
    nz = __ldg(e.noise_strength) *
         __ldg(e.noise + (size_t)(img / e.noise_group_div) * e.noise_group_stride + (size_t)yo * Wo + xo);
  }
  size_t out_idx;
  size_t half_stride = 8;     // distance between channels [0,8) and [8,16) of the chunk
  if (e.out_i8 && e.store_mode == kStoreSpaceToDepth) {
    // space-to-depth + I8: [n][y/2][(phase*Ntot + o)/8][x/2][8]
    const int k0 = ((y & 1) * 2 + (x & 1)) * p.Ntot + n0;
    out_idx = (((size_t)(img * (p.H >> 1) + (y >> 1)) * (p.Ntot >> 1) + (k0 >> 3)) * (p.W >> 1) + (x >> 1)) * 8;
    half_stride = (size_t)(p.W >> 1) * 8;
  } else if (e.out_i8) {
    // channel-group-interleaved [n][yo][Cout/8][xo][8] (regular and depth-to-space stores)
    out_idx = (((size_t)(img * Ho + yo) * (Cout >> 3) + (o0 >> 3)) * Wo + xo) * 8;
    half_stride = (size_t)Wo * 8;
  } else if (e.store_mode == kStoreRegular) {
    out_idx = ((size_t)(img * p.H + y) * p.W + x) * p.Ntot + n0;
  } else if (e.store_mode == kStoreDepthToSpace) {
    out_idx = ((size_t)(img * Ho + yo) * Wo + xo) * Cout + o0;
  } else {
    out_idx = (((size_t)(img * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1)) * 4 + ((y & 1) * 2 + (x & 1))) *
                  p.Ntot + n0;
  }
  float res[16];
  if (e.residual != nullptr) {
    const __half* rbase = e.res_i8 ? e.residual + (((size_t)(img * p.H + y) * (p.Ntot >> 3) + (n0 >> 3)) * p.W + x) * 8
                                   : e.residual + ((size_t)(img * p.H + y) * p.W + x) * p.Ntot + n0;
    const uint4* rp = reinterpret_cast<const uint4*>(rbase);
    uint4 r0 = __ldg(rp), r1 = __ldg(e.res_i8 ? rp + p.W : rp + 1);
    const __half2* h0 = reinterpret_cast<const __half2*>(&r0);
    const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 a = __half22float2(h0[j]), b = __half22float2(h1[j]);
      res[2 * j] = a.x; res[2 * j + 1] = a.y; res[8 + 2 * j] = b.x; res[8 + 2 * j + 1] = b.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int o = o0 + j;
    float t = v[j];
    if (e.dmod != nullptr) t *= __ldg(e.dmod + (size_t)img * Cout + o);
    t += nz;
    if (e.bias != nullptr) t += __ldg(e.bias + o);
    if (e.round_fp16_before_act) t = __half2float(__float2half_rn(t));
    t = act_apply(t, e.act);
    if (e.rgb_w != nullptr) {
      const float* rw = e.rgb_w + (size_t)img * 3 * Cout + o;
      rgb[0] = fmaf(t, __ldg(rw), rgb[0]);
      rgb[1] = fmaf(t, __ldg(rw + Cout), rgb[1]);
      rgb[2] = fmaf(t, __ldg(rw + 2 * Cout), rgb[2]);
    }
    if (e.residual != nullptr) t += res[j];
    t *= e.post_scale;
    if (e.out_scale != nullptr) t *= __ldg(e.out_scale + (size_t)img * e.out_scale_stride + o);
    v[j] = t;
  }
  if (e.out != nullptr) {
    uint4 w0, w1;
    __half2* h0 = reinterpret_cast<__half2*>(&w0);
    __half2* h1 = reinterpret_cast<__half2*>(&w1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h0[j] = f2h2_sat(v[2 * j], v[2 * j + 1]);
      h1[j] = f2h2_sat(v[8 + 2 * j], v[8 + 2 * j + 1]);
    }
    *reinterpret_cast<uint4*>(e.out + out_idx) = w0;
    *reinterpret_cast<uint4*>(e.out + out_idx + half_stride) = w1;
  }
}

// launchers (conv_tc.cu)
struct TmaMaps {
  CUtensorMap a;   // activations [C, W, H, N]
  CUtensorMap b;   // weights     [Cin, Ntot, taps]
};
cudaError_t launch_conv_tc(const ConvParams& p, const TmaMaps& maps, int num_sms, cudaStream_t s);
cudaError_t launch_conv_simt(const ConvParams& p, cudaStream_t s);

}  // namespace glass
