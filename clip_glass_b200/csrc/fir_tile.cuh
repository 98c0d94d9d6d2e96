// Shared pieces of the stride-2 FIR kernels (kernels.cu: fir_down_kernel, from_rgb_fir_kernel; fir_proj_tc.cu: the
// same passes fused with the D blocks' 1x1 projection GEMM on the tensor cores).
#pragma once
#include "common.cuh"

namespace glass {
namespace {

// FIR (pad 1) sampled at stride 2.  One block = 8x16 outputs x 32 channels: the (18 x 34)-pixel input patch is
// staged once in shared memory, so HBM/L2 see every input byte once and the 16 taps per output come from shared
// memory.  The tile is stored in 16-byte units (one pixel, 8 channels) as planes [channel group g][column parity]
// [row][column/2]: the stride-2 reads of the FIR then touch consecutive units (conflict-free LDS.128), and so do
// the writes of consecutive pixels.  Plane strides are padded so that the parity planes sit 64 B apart modulo the
// 128-byte bank row and the group planes 16 B apart.
#ifndef GLASS_FIR_MINB
#define GLASS_FIR_MINB 4       // 64 registers, no spills; 3 -> 4 resident blocks: from_rgb_fir 1.49 -> 1.43 ms, fir_down<I8> 653 -> 624 us
#endif
constexpr int kFdTH = 8, kFdTW = 16, kFdC = 32;
constexpr int kFdIW = 2 * kFdTW + 2, kFdIH = 2 * kFdTH + 2;     // 34 x 18 input pixels
constexpr int kFdPlane = kFdIH * (kFdIW / 2) + 2;               // 308 units: 16 words (mod 32) between parity planes
constexpr int kFdGroup = 2 * kFdPlane + 1;                      // 617 units: 4 words (mod 32) between channel groups
constexpr int kFdUnits = 4 * kFdGroup;
constexpr int kFdPix = kFdIW * kFdIH;                           // 612 pixels per tile
constexpr int kFdItems = kFdPix * 4;                            // (pixel, 8-channel group) items per tile
constexpr int kFdPerThread = (kFdItems + 255) / 256;
__device__ __forceinline__ int fd_unit(int g, int py, int px) {
  return g * kFdGroup + (px & 1) * kFdPlane + py * (kFdIW / 2) + (px >> 1);
}

// Separable [1,3,3,1]/8 FIR at stride 2 from the staged tile.  Thread mapping (256 threads): output column
// ox = tid & 15, channel group g = (tid >> 4) & 3, output rows 2*oyp and 2*oyp+1 with oyp = tid >> 6.  Six input
// rows are filtered horizontally once ((a+d) + 3(b+c): two adds and one FMA per channel instead of four FMAs) and
// shared by the two outputs o[0], o[1] (8 fp16 channels each).
// Arithmetic: packed half2 (GLASS_FIR_FP32 at compile time restores the fp32 version).  The fp32 version spent most
// of its ~28 instructions per output element converting the 24 loaded vectors to fp32 (these passes are issue-bound
// at 67-85 % issue-slot utilisation, profiles/r02_metrics_p64.txt); in half2 a [1,3,3,1] tap is 3 instructions per
// channel pair ((c0 + c3) + 3 (c1 + c2), un-normalised, at most 8 |a|) and the vertical pass applies the whole 1/64.
// Error: six fp16 roundings of the partial sums, ~1e-3 relative in the worst case, on a tensor that is rounded to
// fp16 anyway; the D parity bounds hold unchanged (tests/test_gpu_parity.py).
__device__ __forceinline__ uint4 fd_fir4_h2(const uint4& c0, const uint4& c1, const uint4& c2, const uint4& c3, bool scaled) {
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

__device__ __forceinline__ void fir_down_compute(const uint4* tile, uint4 (&o)[2]) {
  static_assert(kFdTH == 8 && kFdTW == 16, "thread mapping: 16 columns x 4 groups x 4 row pairs = 256 threads");
  const int ox = threadIdx.x & 15;
  const int g = (threadIdx.x >> 4) & 3;
  const int oyp = threadIdx.x >> 6;                 // input rows 4*oyp .. 4*oyp+5
#ifndef GLASS_FIR_FP32
  uint4 hrow[6];
#pragma unroll
  for (int r = 0; r < 6; ++r)
    hrow[r] = fd_fir4_h2(tile[fd_unit(g, 4 * oyp + r, 2 * ox)], tile[fd_unit(g, 4 * oyp + r, 2 * ox + 1)],
                         tile[fd_unit(g, 4 * oyp + r, 2 * ox + 2)], tile[fd_unit(g, 4 * oyp + r, 2 * ox + 3)], false);
#pragma unroll
  for (int k = 0; k < 2; ++k) o[k] = fd_fir4_h2(hrow[2 * k], hrow[2 * k + 1], hrow[2 * k + 2], hrow[2 * k + 3], true);
#else
  float hrow[6][8];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    float v[4][8];
#pragma unroll
    for (int jx = 0; jx < 4; ++jx) {
      const uint4 q = tile[fd_unit(g, 4 * oyp + r, 2 * ox + jx)];
      const __half2* h2 = reinterpret_cast<const __half2*>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 tt = __half22float2(h2[j]);
        v[jx][2 * j] = tt.x;
        v[jx][2 * j + 1] = tt.y;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) hrow[r][j] = fmaf(3.f, v[1][j] + v[2][j], v[0][j] + v[3][j]);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    __half2* oh = reinterpret_cast<__half2*>(&o[k]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float e0 = fmaf(3.f, hrow[2 * k + 1][2 * j] + hrow[2 * k + 2][2 * j], hrow[2 * k][2 * j] + hrow[2 * k + 3][2 * j]);
      const float e1 = fmaf(3.f, hrow[2 * k + 1][2 * j + 1] + hrow[2 * k + 2][2 * j + 1],
                            hrow[2 * k][2 * j + 1] + hrow[2 * k + 3][2 * j + 1]);
      oh[j] = f2h2_sat(e0 * (1.f / 64.f), e1 * (1.f / 64.f));
    }
  }
#endif
}

// Stage the 18 x 34 x 32-channel input patch of tile (ty, tx) of image b, channels c0..c0+31, zero outside the image.
// kI8: x is channel-group-interleaved ([N][H][C/8][W][8]): consecutive threads take consecutive pixels of one group
// (contiguous 16-byte pieces); NHWC: consecutive threads take the four groups of one pixel (64 contiguous bytes).
// All of a thread's loads are issued before the first one is consumed (ten 16-byte loads in flight per thread).
template <bool kI8>
__device__ __forceinline__ void fd_stage_tile(uint4* tile, const __half* __restrict__ x, int b, int H, int W, int C,
                                              int c0, int iy0, int ix0) {
  uint4 v[kFdPerThread];
#pragma unroll
  for (int k = 0; k < kFdPerThread; ++k) {
    const int i = threadIdx.x + k * 256;
    const int g = kI8 ? i / kFdPix : (i & 3), pix = kI8 ? i - g * kFdPix : (i >> 2);
    const int py = pix / kFdIW, px = pix - py * kFdIW;
    const int yy = iy0 + py, xx = ix0 + px;
    v[k] = make_uint4(0, 0, 0, 0);
    if (i < kFdItems && yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const size_t off = kI8 ? ((((size_t)b * H + yy) * (C >> 3) + (c0 >> 3) + g) * W + xx) * 8
                             : (((size_t)b * H + yy) * W + xx) * C + c0 + g * 8;
      v[k] = __ldg(reinterpret_cast<const uint4*>(x + off));
    }
  }
#pragma unroll
  for (int k = 0; k < kFdPerThread; ++k) {
    const int i = threadIdx.x + k * 256;
    const int g = kI8 ? i / kFdPix : (i & 3), pix = kI8 ? i - g * kFdPix : (i >> 2);
    const int py = pix / kFdIW, px = pix - py * kFdIW;
    if (i < kFdItems) tile[fd_unit(g, py, px)] = v[k];
  }
}

// fromRGB for the same patch: x = lrelu(W.(2 rgb - 1) + b) * sqrt2 = lrelu((2 sqrt2 W).rgb + sqrt2 (b - sum W)),
// computed once from the image (halo pixels are recomputed by the neighbouring block: 19 % extra arithmetic, no extra
// HBM traffic), written to HBM for conv0 (interior pixels only) and into the tile.  One thread computes all 32
// channels of a pixel (index math and the rgb loads are paid once per pixel, not once per 8-channel group); the folded
// constants arrive as a kernel parameter, i.e. through the constant bank / uniform registers, not through shared
// memory.  The pass is issue-bound: 1.95 G -> 1.60 G warp instructions per launch at P = 64, 3.0 -> 1.7 ms (profiles/).
struct FrgbConsts { float w[4][kFdC]; };     // rows: r, g, b weights and the bias, channels c0 .. c0+31
// Two horizontally adjacent pixels per thread and step, packed fp32 pairs (FFMA2 with the folded constant broadcast to
// both lanes: 3 FFMA2 + 1 FMUL2 + 2 FMNMX per channel and pixel PAIR instead of 3 FFMA + 1 FMUL + 1 FMNMX per
// pixel); each lane is the scalar arithmetic, so the values are unchanged.  The 34-pixel tile rows hold whole pairs.
constexpr int kFrPairs = kFdPix / 2;
constexpr int kFrPerThread = (kFrPairs + 255) / 256;
static_assert(kFdIW % 2 == 0, "pixel pairs must not straddle tile rows");
__device__ __forceinline__ void fd_stage_tile_from_rgb(uint4* tile, const float* __restrict__ images,
                                                       const FrgbConsts& k, __half* __restrict__ xout, int b, int R,
                                                       int C, int c0, int out_i8, int iy0, int ix0) {
  const size_t plane = (size_t)R * R;
  const float* img = images + (size_t)b * 3 * plane;
  float rgb[kFrPerThread][2][3];
#pragma unroll
  for (int it = 0; it < kFrPerThread; ++it) {
    const int pix = 2 * (threadIdx.x + it * 256);
    const int py = pix / kFdIW, px = pix - py * kFdIW;
    const int yy = iy0 + py;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int xx = ix0 + px + u;
      rgb[it][u][0] = rgb[it][u][1] = rgb[it][u][2] = 0.f;
      if (pix < kFdPix && yy >= 0 && yy < R && xx >= 0 && xx < R) {
        const float* ip = img + (size_t)yy * R + xx;
        rgb[it][u][0] = __ldg(ip); rgb[it][u][1] = __ldg(ip + plane); rgb[it][u][2] = __ldg(ip + 2 * plane);
      }
    }
  }
#pragma unroll
  for (int it = 0; it < kFrPerThread; ++it) {
    const int pix = 2 * (threadIdx.x + it * 256);
    if (pix >= kFdPix) continue;
    const int py = pix / kFdIW, px = pix - py * kFdIW;
    const int yy = iy0 + py;
    const bool row_in = yy >= 0 && yy < R;
    bool inside[2], mine[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int xx = ix0 + px + u;
      inside[u] = row_in && xx >= 0 && xx < R;
      // interior pixels of the tile (rows/cols 1 .. IH-2 / IW-2) belong to this block
      mine[u] = inside[u] && py >= 1 && py < kFdIH - 1 && px + u >= 1 && px + u < kFdIW - 1;
    }
    const f32x2 r2 = pk2(rgb[it][0][0], rgb[it][1][0]), g2 = pk2(rgb[it][0][1], rgb[it][1][1]);
    const f32x2 b2 = pk2(rgb[it][0][2], rgb[it][1][2]);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 pk[2];
      __half2* h2a = reinterpret_cast<__half2*>(&pk[0]);
      __half2* h2b = reinterpret_cast<__half2*>(&pk[1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float lo[2], hi[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int c = g * 8 + 2 * j + u;
          const f32x2 t = fma2(r2, pk2(k.w[0][c], k.w[0][c]),
                               fma2(g2, pk2(k.w[1][c], k.w[1][c]), fma2(b2, pk2(k.w[2][c], k.w[2][c]), pk2(k.w[3][c], k.w[3][c]))));
          upk2(lrelu2(t), lo[u], hi[u]);
        }
        h2a[j] = f2h2_sat(lo[0], lo[1]);       // pixel px:     channels c, c+1
        h2b[j] = f2h2_sat(hi[0], hi[1]);       // pixel px + 1
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!inside[u]) pk[u] = make_uint4(0, 0, 0, 0);           // outside the image: the FIR's zero padding
        if (mine[u]) {
          const int xx = ix0 + px + u;
          const size_t off = out_i8 ? ((((size_t)b * R + yy) * (C >> 3) + (c0 >> 3) + g) * R + xx) * 8
                                    : (((size_t)b * R + yy) * R + xx) * C + c0 + g * 8;
          *reinterpret_cast<uint4*>(xout + off) = pk[u];
        }
        tile[fd_unit(g, py, px + u)] = pk[u];
      }
    }
  }
}

// tile index -> (image b, tile row ty, tile column tx) for an Ho x Wo output grid
__device__ __forceinline__ void fd_decode_tile(int t, int Ho, int Wo, int& b, int& ty, int& tx) {
  const int tiles_x = (Wo + kFdTW - 1) / kFdTW, tiles_y = (Ho + kFdTH - 1) / kFdTH;
  tx = t % tiles_x; t /= tiles_x;
  ty = t % tiles_y;
  b = t / tiles_y;
}

}  // namespace
}  // namespace glass
