// Implicit-GEMM convolution / GEMM on the 5th-gen tensor cores (sm_100a).
//
//   warp 0      TMA producer (one lane, chosen with elect.sync).  MODE 0: per K step one 4-D box of the NHWC
//               activation tensor (shifted by the filter tap; out-of-image rows/cols are zero-filled by TMA) + one
//               box of the [taps][Ntot][Cin] weights, both landing 128B/64B-swizzled in shared memory.
//               MODE 1/2/4: the taps of the CTA's n-tile stay resident, a stage holds the (haloed) input tile.
//               MODE 6: the input box of a tile pair stays put, the taps stream through a ring (Cfg below).
//   warp 1      tcgen05.mma issuer (one elected lane), fp16 x fp16 -> fp32 accumulators in TMEM, double-buffered
//               (2 x BN columns, 2 x 2 x BN for tile pairs).
//   warps 2..   epilogue (8 warps, 16 for one instance): tcgen05.ld the 128 x BN tile, apply the fused demod / noise /
//               bias / activation / toRGB / residual / style pre-scale and store fp16 (NHWC, depth-to-space,
//               space-to-depth, each also channel-group-interleaved).  Compile-time epilogue specialisations
//               (kEpiSpecs) for the hot layers; common.cuh: epilogue_row16 for tiles that span several images.
//
// Persistent CTAs (one or two per SM), static round-robin over (m_tile, n_tile).
// The SIMT kernel at the bottom computes the same accumulators the slow way and
// shares the epilogue: it exists for bring-up and as the in-library cross-check
// (glass_config.conv_impl = 1); it is never used by the product path.
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "tcgen05.cuh"

namespace glass {

namespace {

#ifndef GLASS_BN32_CTAS
#define GLASS_BN32_CTAS 2
#endif
constexpr int kBlockM = 128;
constexpr int kMaxThreads = 64 + 16 * 32;   // TMA warp, MMA warp, up to 16 epilogue warps

struct TileCoord { int n_tile, tx, ty, tn; };
template <bool kPow2 = false>
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int tile, int n_tiles) {
  TileCoord t;
  if (kPow2 || p.pow2) {
    t.n_tile = tile & (n_tiles - 1);
    int m = tile >> p.sh_n;
    if (!kPow2 && p.skip_mode == 1) t.n_tile = (t.n_tile + m / p.rot_div) & (n_tiles - 1);
    t.tx = m & (p.tiles_x - 1); m >>= p.sh_x;
    t.ty = m & (p.tiles_y - 1);
    t.tn = m >> p.sh_y;
  } else {
    t.n_tile = tile % n_tiles;
    int m = tile / n_tiles;
    if (p.skip_mode == 1) t.n_tile = (t.n_tile + m / p.rot_div) % n_tiles;
    t.tx = m % p.tiles_x; m /= p.tiles_x;
    t.ty = m % p.tiles_y;
    t.tn = m / p.tiles_y;
  }
  return t;
}

// true if the (tap, k-chunk) weight block of this n-tile is structurally zero (ConvParams::skip_mode)
__device__ __forceinline__ bool skip_block(const ConvParams& p, int tap, int kc, int n_tile, int BN, int BK) {
  if (p.skip_mode == 1) {
    const int ph = (n_tile * BN) / p.skip_ch;
    return (p.tap_dy[tap] != 0 && (ph >> 1)) || (p.tap_dx[tap] != 0 && (ph & 1));
  }
  if (p.skip_mode == 2) {
    const int ph = (kc * BK) / p.skip_ch;
    return ((tap >> 1) && (ph >> 1)) || ((tap & 1) && (ph & 1));
  }
  return false;
}

constexpr int kNumParams = 6;   // scale, shift, oscale, rgb0, rgb1, rgb2

// Compile-time epilogue specialisations.  Spec 0 reads every feature switch from EpiParams at run time (any layer);
// the others fix the switches of one hot small-channel layer family, which removes ~150 uniform loads / compares /
// branches per (warp, tile) and shrinks the unrolled chunk bodies from ~2600 to a few hundred SASS instructions
// (the epilogue of these layers is issue/latency-bound).  pick_epi_spec() selects a spec only
// when the layer matches it exactly; everything else runs spec 0.
enum { kStNone = 0, kStRegular = 1, kStD2S = 2, kStS2D = 3 };
struct EpiSpec {
  bool generic, noise, rgb, residual;
  int store;
  bool i8;
  int act;      // kActNone / kActLrelu / kActQuickGelu
  bool round = false;   // round to fp16 before the activation / residual add (CLIP's fp16 op boundaries)
  bool pow2 = true;     // tile grid and channel count are powers of two (decode with shifts); false: plain GEMMs
};
constexpr EpiSpec kEpiSpecs[] = {
    {true, false, false, false, kStNone, false, 0},               // 0: run-time switches
    {false, true, true, false, kStRegular, true, kActLrelu},      // 1: G conv + toRGB, I8 store
    {false, true, false, false, kStD2S, true, kActLrelu},         // 2: G folded up-conv, depth-to-space I8
    {false, true, true, false, kStNone, false, kActLrelu},        // 3: last G conv: toRGB only
    {false, false, false, false, kStS2D, false, kActLrelu},       // 4: D conv0, space-to-depth NHWC
    {false, true, true, false, kStRegular, false, kActLrelu},     // 5: G conv + toRGB, NHWC store
    {false, true, false, false, kStD2S, false, kActLrelu},        // 6: G folded up-conv, depth-to-space NHWC
    {false, false, false, true, kStRegular, true, kActLrelu},     // 7: D conv1 + residual, I8 store
    {false, false, false, true, kStRegular, false, kActLrelu},    // 8: D conv1 + residual, NHWC store
    {false, false, false, false, kStRegular, false, kActNone},    // 9: D projection (1x1, linear)
    {false, false, false, false, kStRegular, false, kActLrelu},   // 10: D conv0 of the exact form, NHWC store
    {false, false, false, false, kStS2D, true, kActLrelu},        // 11: D conv0, space-to-depth I8 (feeds MODE 6)
    {false, false, false, false, kStRegular, true, kActNone},     // 12: D projection (1x1, linear), I8 store
    {false, false, false, false, kStRegular, true, kActLrelu},    // 13: D conv0 feeding the fused-FIR down-conv, I8 store
    // ViT GEMMs (M = 3200 rows: 25 row tiles, N = 768 / 2304 / 3072).  Under the run-time spec these launches were
    // bound by instruction fetch (no_instruction was their top warp stall: ~2600 unrolled SASS instructions per chunk
    // body for a 35-85 us kernel) and by the residual loads issued chunk by chunk (profiles/r02_ncu_vit_gemms.txt).
    {false, false, false, false, kStRegular, false, kActNone, false, false},        // 14: qkv / patch embedding: (+ bias)
    {false, false, false, true, kStRegular, false, kActNone, true, false},          // 15: out / proj: round, + residual
    {false, false, false, false, kStRegular, false, kActQuickGelu, true, false},    // 16: fc: round, QuickGELU
};
constexpr int kNumEpiSpecs = sizeof(kEpiSpecs) / sizeof(kEpiSpecs[0]);
// MODE 0 ("stream"): one pipeline stage per (filter tap, 64-channel chunk): A box + B box per stage.
// MODE 1 ("halo"):   for layers whose whole K per tap is one chunk (Cin == BK in {32,64}): the filter taps of the
//                    n-tile stay resident in shared memory for the whole kernel, and a stage holds the input
//                    tile ONCE per horizontal shift: 3 copies of (TH+2) x TW pixels (dx = -1,0,+1).  The nine
//                    taps are nine descriptor offsets into those copies (a vertical shift is a whole number of
//                    1024-byte swizzle atoms), so L2->SM traffic drops from 9x to 3.75x the tile and the
//                    producer issues 3 TMA ops per tile instead of 18.
constexpr int kHaloTH = 8, kHaloTW = 16;


template <int BN, int BK, int MODE>
struct Cfg {
  static constexpr int kABytes = kBlockM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kCopyBytes = (kHaloTH + 2) * kHaloTW * BK * 2;          // one dx-copy of the halo tile
  // MODE 2 = MODE 1 for 1x1 convs: one resident tap, one un-haloed tile per stage
  // MODE 4 ("I8"): the input tensor is stored channel-group-interleaved, [N][H][C/8][W][8].  One TMA box of
  //   (TW+2)*8 contiguous elements x C/8 groups x (TH+2) rows lands in shared memory as [row][group][pixel][8ch]:
  //   exactly the un-swizzled K-major core-matrix layout (8 pixels x 16 B contiguous; next 8-channel group at
  //   LBO = (TW+2)*16 B; next image row at SBO = (C/8)*(TW+2)*16 B; tile = 8 wide x 16 tall).  A filter tap is just
  //   a different start address (16-byte granularity), so ONE copy of the haloed tile serves all nine taps:
  //   1.4x the tile in L2->SM bytes (MODE 1: 3.75x, MODE 0: 9x) and (TH+2)*C/8 long TMA rows instead of hundreds
  //   of 64/128-byte ones (TMA issues ~0.41 rows/cycle/SM regardless of their length).
  // resident weights are kept in 64-channel (128-byte, swizzled) K chunks: [tap][chunk][BN x kBKc]
  static constexpr int kBKc = BK > 64 ? 64 : BK;
  static constexpr int kKChunks = BK / kBKc;
  static constexpr int kBChunkBytes = BN * kBKc * 2;
  // Tile pairs (kPairM = 2): two horizontally adjacent 8x16 tiles share ONE haloed box (18 pixels wide) and one
  // trip through the producer / MMA / epilogue bookkeeping; their accumulators sit side by side in TMEM.  Halves the
  // per-element bookkeeping of the epilogue-bound small-channel layers and gives every epilogue warp two independent
  // chunks to overlap.
  // MODE 6 ("I8 + streamed taps"): the I8 haloed box of MODE 4 for layers whose nine taps do NOT fit beside it in
  //   shared memory (Cin = 128, BN = 64: 147 KB of weights).  The input box of a tile PAIR stays put while the nine
  //   taps stream through a small ring of 64-channel chunks (one weight pass per 256 pixels).  Against MODE 0 the
  //   L2->SM traffic of the folded D 512^2 down-conv drops from 435 KB to 115 KB per 128 pixels (it ran at the
  //   TMA/L2 delivery limit: 60 GB per launch, profiles/).
  static constexpr bool kI8 = (MODE == 4 || MODE == 6);
  static constexpr int kBStages = (MODE == 6) ? 6 : 0;            // ring of streamed weight chunks
  static constexpr int kPairM = ((MODE == 4 && BK <= 64) || MODE == 6) ? 2 : 1;
  static constexpr int kI8TW = 8, kI8TH = 16;
  static constexpr int kI8RowBytes = (kI8TW * kPairM + 2) * 16;                 // one (row, group): 10 or 18 pixels x 16 B
  static constexpr int kI8StageBytes = (kI8TH + 2) * (BK / 8) * kI8RowBytes;
  static constexpr int kStageBytes = MODE == 0 ? kABytes + kBBytes
                                     : (MODE == 1 ? 3 * kCopyBytes : (kI8 ? kI8StageBytes : kABytes));
  static constexpr int kWBytes = MODE == 0 ? 0   // resident taps, or MODE 6's ring of weight chunks
                                 : (MODE == 6 ? kBStages * kBChunkBytes : ((MODE == 1 || MODE == 4) ? 9 : 1) * kBBytes);
  // Epilogue warps: two per TMEM lane quarter (each owning half of the columns).  Four per quarter (16 warps, 96
  // registers/thread) was measured at P=64: it helps the wide resident-tap instance (G up 64->32 @1024^2: 4.23 ->
  // 3.56 ms) and costs 5-25 % on the streamed large-K instances (spills), so only that instance uses it.
  // The 32->32 I8 instances (G 1024^2 conv, D 1024^2 conv) are bound by per-(warp, tile) bookkeeping: there one
  // warp per lane quarter owns all 32 columns (half the bookkeeping per element, no cross-warp toRGB combine)
  // and GLASS_BN32_CTAS small CTAs per SM supply the warps that hide latency.
  static constexpr bool kSmallN = (MODE == 4 && BK == 32 && BN == 32 && GLASS_BN32_CTAS > 2);   // measured slower
  // (16 warps for the 64-column 64-channel MODE 4 layers, measured at the end of round 2: G15 2.40 -> 2.57 ms, D1:c0
  // 1.25 -> 1.40 ms: 96 registers at 576 threads spill and the four column parts quadruple the toRGB exchange)
  static constexpr int kEpiWarps = (MODE == 4 && BN == 128) ? 16 : (kSmallN ? 4 : 8);
  static constexpr int kParts = kEpiWarps / 4;                      // column parts per lane quarter
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  // double-buffered per-tile epilogue parameters + double-buffered staging of the non-leading parts' toRGB sums
  static constexpr int kParamBytes = 2 * kNumParams * BN * 4 + 2 * (kParts - 1) * kPairM * 128 * 16;
  // the 32-channel MODE-1 layers are bookkeeping/latency-bound, not smem-bound: run two CTAs per SM there
  static constexpr int kMinBlocks = kSmallN ? GLASS_BN32_CTAS : (((MODE == 1 || MODE == 4) && BK == 32 && BN <= 32) ? 2 : 1);
  static constexpr int kBudget = (kMinBlocks == 4 ? 54 : (kMinBlocks == 3 ? 73 : (kMinBlocks == 2 ? 110 : 222))) * 1024 -
                                 kParamBytes - kWBytes;   // of 227 KB/SM (+1 KB reserved per CTA)
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 12 ? 12 : kStagesRaw;
  static_assert(kStages >= 2, "not enough shared memory for a double-buffered pipeline");
  static constexpr int kAccCols = kPairM * BN;                     // TMEM columns of one accumulator stage
  static constexpr int kTmemCols = (2 * kAccCols < 32) ? 32 : 2 * kAccCols;   // power of two for BN in {32,..,256}
  static_assert(kTmemCols <= 512, "TMEM has 512 columns");
  static constexpr int kSmemBytes = kWBytes + kStages * kStageBytes + kParamBytes + 1024 /*align*/ + 256 /*barriers*/;
  // instruction descriptor: D=f32 [4,6)=1, A=B=f16 (0), K-major both, N>>3 [17,23), M>>4 [24,29)
  static constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
};

template <int kThreads>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory"); }

// Fast epilogue for one row and 16 columns with the per-tile parameters staged in shared memory.
//   t = act(acc*scale + shift + nz) ; rgb += t*rgbw ; t = (t + residual) * oscale ; store fp16
// (the sqrt(2) gain of lrelu is folded into scale/shift/nz: lrelu(a)*g == lrelu(a*g) for g > 0)
// kActT: -1 = activation and fp16 pre-rounding from EpiParams at run time; kActNone / kActLrelu = fixed, no rounding
// NT = 2: the two tiles of a pair in one pass -- every parameter vector is read from shared memory ONCE for both
// tiles.  (The broadcast LDS.128 of the staged parameters were 42 % of the shared-memory wavefronts of the 32-channel
// 1024^2 layers, whose LSU data pipe ran at 89 %: profiles/.)
// kRoundT: -1 = fp16 pre-rounding from EpiParams at run time, 0 / 1 = fixed
template <int NT, bool kRgb, int kActT = -1, bool kNz = true, int kRoundT = -1>
__device__ __forceinline__ void epilogue_fastN(const EpiParams& e, const float* __restrict__ par, int BN, int j0,
                                               const uint32_t (*acc)[16], const float* nz, const __half* const* res_ptr,
                                               __half* const* out_ptr, size_t out_half_stride, float (*rgb)[3],
                                               const uint4* const* res_pre, size_t res_half_stride = 8) {
  // All per-element math runs on packed fp32 pairs (common.cuh: FFMA2 / FMUL2 / FADD2): pair k of a tile holds
  // columns 2k, 2k+1 of the chunk.  Each lane is rounded like the scalar instruction, so the values are those of the
  // scalar form; only the toRGB dot product is summed in a different (even / odd column) order.
  const float4* sc = reinterpret_cast<const float4*>(par + 0 * BN + j0);
  const float4* sh = reinterpret_cast<const float4*>(par + 1 * BN + j0);
  const float4* os = reinterpret_cast<const float4*>(par + 2 * BN + j0);
  f32x2 t[NT][8];
  f32x2 nz2[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) nz2[n] = pk2(nz[n], nz[n]);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float4 a = sc[g], b = sh[g];
    const f32x2 a01 = pk2(a.x, a.y), a23 = pk2(a.z, a.w), b01 = pk2(b.x, b.y), b23 = pk2(b.z, b.w);
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const f32x2 v01 = pk2(__uint_as_float(acc[n][4 * g + 0]), __uint_as_float(acc[n][4 * g + 1]));
      const f32x2 v23 = pk2(__uint_as_float(acc[n][4 * g + 2]), __uint_as_float(acc[n][4 * g + 3]));
      t[n][2 * g] = fma2(v01, a01, kNz ? add2(b01, nz2[n]) : b01);
      t[n][2 * g + 1] = fma2(v23, a23, kNz ? add2(b23, nz2[n]) : b23);
    }
  }
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    if (kRoundT > 0 || (kRoundT < 0 && kActT < 0 && e.round_fp16_before_act)) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float lo, hi;
        upk2(t[n][k], lo, hi);
        const float2 r = __half22float2(__floats2half2_rn(lo, hi));
        t[n][k] = pk2(r.x, r.y);
      }
    }
    if (kActT == kActLrelu || (kActT < 0 && e.act == kActLrelu)) {
#pragma unroll
      for (int k = 0; k < 8; ++k) t[n][k] = lrelu2(t[n][k]);
    } else if (kActT == kActQuickGelu || (kActT < 0 && e.act == kActQuickGelu)) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float lo, hi;
        upk2(t[n][k], lo, hi);
        // x * sigmoid(1.702 x) with the approximate reciprocal (MUFU.RCP, <= 2 ulp: the value is rounded to fp16 right
        // after); the IEEE division was 8 of the ~14 instructions per element of the fc GEMM's epilogue
        t[n][k] = pk2(__fdividef(lo, 1.f + __expf(-1.702f * lo)), __fdividef(hi, 1.f + __expf(-1.702f * hi)));
      }
    }
  }
  if (kRgb) {
    const float4* r0 = reinterpret_cast<const float4*>(par + 3 * BN + j0);
    const float4* r1 = reinterpret_cast<const float4*>(par + 4 * BN + j0);
    const float4* r2 = reinterpret_cast<const float4*>(par + 5 * BN + j0);
    f32x2 s[NT][3];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int c = 0; c < 3; ++c) s[n][c] = pk2(rgb[n][c], 0.f);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 a = r0[g], b = r1[g], c = r2[g];
      const f32x2 a01 = pk2(a.x, a.y), a23 = pk2(a.z, a.w), b01 = pk2(b.x, b.y), b23 = pk2(b.z, b.w);
      const f32x2 c01 = pk2(c.x, c.y), c23 = pk2(c.z, c.w);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        s[n][0] = fma2(t[n][2 * g], a01, s[n][0]); s[n][0] = fma2(t[n][2 * g + 1], a23, s[n][0]);
        s[n][1] = fma2(t[n][2 * g], b01, s[n][1]); s[n][1] = fma2(t[n][2 * g + 1], b23, s[n][1]);
        s[n][2] = fma2(t[n][2 * g], c01, s[n][2]); s[n][2] = fma2(t[n][2 * g + 1], c23, s[n][2]);
      }
    }
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float lo, hi;
        upk2(s[n][c], lo, hi);
        rgb[n][c] = lo + hi;
      }
  }
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    if (res_ptr[n] != nullptr || res_pre[n] != nullptr) {
      const uint4* rp = reinterpret_cast<const uint4*>(res_ptr[n]);
      // res_pre: the 16 residual values were fetched before the accumulator wait (their DRAM latency is hidden)
      const uint4 q0 = res_pre[n] != nullptr ? res_pre[n][0] : __ldg(rp);
      // channels [8,16) of the chunk: adjacent in NHWC, one channel-group plane further in an I8 residual tensor
      const uint4 q1 = res_pre[n] != nullptr ? res_pre[n][1]
                                             : __ldg(reinterpret_cast<const uint4*>(res_ptr[n] + res_half_stride));
      const __half2* h0 = reinterpret_cast<const __half2*>(&q0);
      const __half2* h1 = reinterpret_cast<const __half2*>(&q1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = __half22float2(h0[j]), b = __half22float2(h1[j]);
        t[n][j] = add2(t[n][j], pk2(a.x, a.y));
        t[n][4 + j] = add2(t[n][4 + j], pk2(b.x, b.y));
      }
    }
  }
  bool any_out = false;
#pragma unroll
  for (int n = 0; n < NT; ++n) any_out |= out_ptr[n] != nullptr;
  if (any_out) {
    f32x2 osv[8];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 a = os[g];
      osv[2 * g] = pk2(a.x, a.y);
      osv[2 * g + 1] = pk2(a.z, a.w);
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      if (out_ptr[n] == nullptr) continue;
      uint4 w0, w1;
      __half2* h0 = reinterpret_cast<__half2*>(&w0);
      __half2* h1 = reinterpret_cast<__half2*>(&w1);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float lo, hi;
        upk2(mul2(t[n][k], osv[k]), lo, hi);
        h0[k] = f2h2_sat(lo, hi);
        upk2(mul2(t[n][4 + k], osv[4 + k]), lo, hi);
        h1[k] = f2h2_sat(lo, hi);
      }
      // channels [0,8) and [8,16) of the chunk: adjacent in NHWC, one channel-group plane apart in the I8 layout
      *reinterpret_cast<uint4*>(out_ptr[n]) = w0;
      *reinterpret_cast<uint4*>(out_ptr[n] + out_half_stride) = w1;
    }
  }
}

template <bool kRgb, int kActT = -1, bool kNz = true, int kRoundT = -1>
__device__ __forceinline__ void epilogue_fast16(const EpiParams& e, const float* __restrict__ par, int BN, int j0,
                                                const uint32_t (&acc)[16], float nz, const __half* res_ptr,
                                                __half* out_ptr, size_t out_half_stride, float (&rgb)[3],
                                                const uint4* res_pre = nullptr, size_t res_half_stride = 8) {
  const __half* rp[1] = {res_ptr};
  __half* op[1] = {out_ptr};
  const uint4* pre[1] = {res_pre};
  const float nzv[1] = {nz};
  epilogue_fastN<1, kRgb, kActT, kNz, kRoundT>(e, par, BN, j0, &acc, nzv, rp, op, out_half_stride, &rgb, pre, res_half_stride);
}

template <int BN, int BK, int MODE, int EPI>
__global__ void __launch_bounds__((Cfg<BN, BK, MODE>::kThreads), (Cfg<BN, BK, MODE>::kMinBlocks))
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const ConvParams p) {
  using C = Cfg<BN, BK, MODE>;
  constexpr EpiSpec S = kEpiSpecs[EPI];
  constexpr bool kPow2 = !S.generic && S.pow2;   // power-of-two tile grids and channel counts: decode with shifts
  extern __shared__ uint8_t smem_raw[];
  // (offset arithmetic on the extern array keeps the pointers in the shared address space: LDS/STS, not generic LD/ST)
  uint8_t* smem_w = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem = smem_w + C::kWBytes;       // pipeline stages
  float* params = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes);
  float4* rgb_stage = reinterpret_cast<float4*>(params + 2 * kNumParams * BN);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes + C::kParamBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* w_bar = tmem_empty + 2;
  uint64_t* b_full = w_bar + 1;                 // MODE 6: ring of streamed weight chunks (kBStages == 0 otherwise)
  uint64_t* b_empty = b_full + C::kBStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + C::kBStages);
  static_assert((2 * C::kStages + 5 + 2 * C::kBStages) * 8 + 8 <= 256, "barrier block overflows its 256 bytes");
  float* nscale_slot = reinterpret_cast<float*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kchunks = p.Cin / BK;
  const int kiters = p.taps * kchunks;
  const int n_tiles = p.Ntot / BN;
  const int m_tiles = p.tiles_n * p.tiles_y * p.tiles_x;
  const int total_tiles = m_tiles * n_tiles;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], C::kEpiWarps);
    }
    mbar_init(w_bar, 1);
    for (int s = 0; s < C::kBStages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
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
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int bstage = 0;
      uint32_t bphase = 0;
      if (MODE != 0 && MODE != 6) {
        // resident filter taps of this CTA's n-tile (grid is a multiple of n_tiles, so n_tile is fixed)
        const int n_tile = blockIdx.x % n_tiles;
        mbar_expect_tx(w_bar, p.taps * C::kBBytes);
        for (int tap = 0; tap < p.taps; ++tap)
          for (int ch = 0; ch < C::kKChunks; ++ch)
            tma_load_3d(&map_b, smem_w + (tap * C::kKChunks + ch) * C::kBChunkBytes, w_bar, ch * C::kBKc, n_tile * BN,
                        tap);
      }
      if (MODE == 6) {
        // The box of tile i+1 goes into the stage that tile i-1 is still using when the producer starts streaming the
        // taps of tile i (two stages), so it cannot be requested up front without stalling the weight ring for a
        // whole tile; requested only after tile i's taps it would arrive a full TMA round trip late.  Hence: while
        // streaming tile i's taps, poll that stage's barrier without blocking and request the box as soon as tile
        // i-1's MMAs have released it (a few chunks into tile i).
        auto issue_box = [&](int tile) {
          const TileCoord tc = decode_tile<kPow2>(p, tile, n_tiles);
          mbar_expect_tx(&full_bar[stage], C::kI8StageBytes);
          tma_load_4d(&map_a, smem + stage * C::kStageBytes, &full_bar[stage], (tc.tx * p.TW - 1) * 8, 0,
                      tc.ty * p.TH - 1, tc.tn * p.TN);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        };
        const int n_tile = blockIdx.x % n_tiles;
        if ((int)blockIdx.x < total_tiles) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          issue_box(blockIdx.x);
        }
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          bool next_pending = tile + (int)gridDim.x < total_tiles;
          for (int tap = 0; tap < 9; ++tap)
            for (int ch = 0; ch < C::kKChunks; ++ch) {
              if (next_pending && mbar_test(&empty_bar[stage], phase ^ 1)) {
                issue_box(tile + gridDim.x);
                next_pending = false;
              }
              mbar_wait(&b_empty[bstage], bphase ^ 1);
              mbar_expect_tx(&b_full[bstage], C::kBChunkBytes);
              tma_load_3d(&map_b, smem_w + bstage * C::kBChunkBytes, &b_full[bstage], ch * C::kBKc, n_tile * BN, tap);
              if (++bstage == C::kBStages) { bstage = 0; bphase ^= 1; }
            }
          if (next_pending) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            issue_box(tile + gridDim.x);
          }
        }
      }
      for (int tile = blockIdx.x; MODE != 6 && tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile<kPow2>(p, tile, n_tiles);
        const int n_tile = tc.n_tile;
        const int x0 = tc.tx * p.TW, y0 = tc.ty * p.TH, i0 = tc.tn * p.TN;
        if (MODE != 0) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::kStageBytes;
          if (C::kI8) {
            mbar_expect_tx(&full_bar[stage], C::kI8StageBytes);
            // coordinates: ((x0-1)*8 elements, group 0, row y0-1, image); out-of-image parts are zero-filled
            tma_load_4d(&map_a, sa, &full_bar[stage], (x0 - 1) * 8, 0, y0 - 1, i0);
          } else if (p.taps == 9) {
            mbar_expect_tx(&full_bar[stage], 3 * C::kCopyBytes);
#pragma unroll
            for (int c = 0; c < 3; ++c)
              tma_load_4d(&map_a, sa + c * C::kCopyBytes, &full_bar[stage], 0, x0 + c - 1, y0 - 1, i0);
          } else {
            mbar_expect_tx(&full_bar[stage], C::kABytes);
            tma_load_4d(&map_a, sa, &full_bar[stage], 0, x0, y0, i0);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
          continue;
        }
        for (int kit = 0; kit < kiters; ++kit) {
          const int tap = kit / kchunks;
          const int kc = kit - tap * kchunks;
          const int dy = p.tap_dy[tap], dx = p.tap_dx[tap];
          if (skip_block(p, tap, kc, n_tile, BN, BK)) continue;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
          tma_load_4d(&map_a, sa, &full_bar[stage], kc * BK, x0 + dx, y0 + dy, i0);
          tma_load_3d(&map_b, sb, &full_bar[stage], kc * BK, n_tile * BN, tap);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      int bstage = 0;
      uint32_t bphase = 0;
      if (MODE != 0 && MODE != 6) {
        mbar_wait(w_bar, 0);
        tc_fence_after();
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * C::kAccCols;
        if (MODE != 0) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t sw = smem_u32(smem_w);
          if (MODE == 6) {
            // One thread issues 144 MMAs per tile pair: everything but two 32-bit adds per MMA is folded at compile
            // time.  The loops are fully unrolled (18 weight chunks per tile is a multiple of the ring depth, so a
            // chunk's ring slot is a constant), and operand descriptors are a per-stage base plus a constant: the
            // start-address field is the low 14 bits (bytes >> 4) and cannot carry out while the operand lies inside
            // the 227 KB window.  (Built per MMA at run time, the descriptor arithmetic of this single thread was the
            // bottleneck: ~100 cycles per MMA against a 32-cycle tensor-pipe floor.)
            constexpr uint32_t kLbo = C::kI8RowBytes;
            constexpr uint32_t kSbo = (BK / 8) * C::kI8RowBytes;
            constexpr int kPerChunk = C::kBKc / 16;
            constexpr int kRing = C::kBStages > 0 ? C::kBStages : 1;      // (this branch is dead unless MODE == 6)
            static_assert((9 * C::kKChunks) % kRing == 0, "ring slots must be compile-time constants");
            const uint64_t da0 = make_smem_desc_noswz(sa, kLbo, kSbo);
            const uint64_t db0 = make_smem_desc<C::kBKc>(sw);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int ch = 0; ch < C::kKChunks; ++ch) {
                const int slot = (tap * C::kKChunks + ch) % kRing;                  // constant after unrolling
                mbar_wait(&b_full[slot], bphase);
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < C::kPairM; ++h) {
#pragma unroll
                  for (int kk = 0; kk < kPerChunk; ++kk) {
                    const int k = ch * kPerChunk + kk;
                    const uint32_t a_off = (tap / 3) * kSbo + ((tap % 3) + 8 * h) * 16 + k * 2 * kLbo;
                    const uint32_t b_off = slot * C::kBChunkBytes + kk * 32;
                    tc_mma_f16(d_tmem + h * BN, da0 + (uint64_t)(a_off >> 4), db0 + (uint64_t)(b_off >> 4), C::kIdesc,
                               (tap | k) != 0);
                  }
                }
                tc_commit(&b_empty[slot]);
                if (slot == kRing - 1) bphase ^= 1;
              }
            }
            tc_commit(&empty_bar[stage]);
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            tc_commit(&tmem_full[as]);
            continue;
          }
          if (MODE == 4) {
            constexpr uint32_t kLbo = C::kI8RowBytes;                 // next 8-channel group
            constexpr uint32_t kSbo = (BK / 8) * C::kI8RowBytes;      // next image row (= next 8-pixel group)
            const uint32_t lbo = (kDebugBuild && (p.debug_skip & 2)) ? kSbo : kLbo;    // (bring-up knob: swapped roles)
            const uint32_t sbo = (kDebugBuild && (p.debug_skip & 2)) ? kLbo : kSbo;
#pragma unroll
            for (int h = 0; h < C::kPairM; ++h) {                     // tile h of the pair: 8 pixels further right
              for (int tap = 0; tap < 9; ++tap) {
                const int ky = tap / 3, kx = tap - ky * 3;
                const uint32_t a_addr = sa + ky * kSbo + (kx + 8 * h) * 16;   // pixel (ry+ky, rx+kx) of the haloed tile
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  constexpr int kPerChunk = C::kBKc / 16;
                  const int ch = k / kPerChunk, kk = k - ch * kPerChunk;
                  const uint64_t db = make_smem_desc<C::kBKc>(sw + (tap * C::kKChunks + ch) * C::kBChunkBytes);
                  const uint64_t da = make_smem_desc_noswz(a_addr + k * 2 * kLbo, lbo, sbo);
                  tc_mma_f16(d_tmem + h * BN, da, db + (uint64_t)(kk * 2), C::kIdesc, (tap | k) != 0);
                }
              }
            }
            tc_commit(&empty_bar[stage]);
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            tc_commit(&tmem_full[as]);
            continue;
          }
          for (int tap = 0; tap < p.taps; ++tap) {
            // tap (ky,kx): copy kx holds the tile shifted by dx = kx-1; row offset ky*TW pixels shifts by dy = ky-1
            const int ky = tap / 3, kx = tap - ky * 3;
            const uint32_t a_addr = (p.taps == 9) ? sa + kx * C::kCopyBytes + ky * (kHaloTW * BK * 2) : sa;
            const uint64_t da = make_smem_desc<BK>(a_addr);
            const uint64_t db = make_smem_desc<BK>(sw + tap * C::kBBytes);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc_mma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), C::kIdesc, (tap | k) != 0);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
          tc_commit(&tmem_full[as]);
          continue;
        }
        const int n_tile_mma = p.skip_mode == 1 ? decode_tile<kPow2>(p, tile, n_tiles).n_tile : 0;
        uint32_t started = 0;
        for (int kit = 0; kit < kiters; ++kit) {
          if (p.skip_mode != 0) {
            const int tap = kit / kchunks;
            if (skip_block(p, tap, kit - tap * kchunks, n_tile_mma, BN, BK)) continue;
          }
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint64_t da = make_smem_desc<BK>(sa);
          const uint64_t db = make_smem_desc<BK>(sa + C::kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 fp16 = 32 bytes along K inside the swizzle atom
            tc_mma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), C::kIdesc, started | (uint32_t)k);
          }
          started = 1;
          tc_commit(&empty_bar[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[as]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..) =====================
    // kParts warps per TMEM lane quarter; each owns 1/kParts of the tile's columns (of both tiles of a pair).
    const EpiParams& e = p.epi;
    const int ew = warp - 2;
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = ew >> 2;               // column part owned by this warp (0 .. kParts-1)
    const int et = threadIdx.x - 64;        // 0 .. 32*kEpiWarps-1
    const int row = q * 32 + lane;          // accumulator row == pixel within the (first) tile
    constexpr int kPairM = C::kPairM;
    const int geo_w = (kPairM == 2) ? C::kI8TW : p.TW;     // width of ONE 128-pixel tile
    const int thw = p.TH * geo_w;
    const int ri = row / thw;
    const int rr = row - ri * thw;
    const int ry = rr / geo_w;
    const int rx = rr - ry * geo_w;         // tile h of a pair: pixel column rx + 8h
    // every row of a tile belongs to one image (always for I8 tiles and for the specialised layers)
    const bool fast = (C::kI8 || !S.generic) ? true : (p.TN == 1);
    const float gain = (S.generic ? (e.act == kActLrelu) : (S.act == kActLrelu)) ? kSqrt2 : 1.f;
    constexpr int kParts = C::kParts;
    constexpr int kHalf = BN / kParts;      // columns per warp
    constexpr int kChunks = kHalf / 16;
    static_assert(kHalf % 16 == 0, "a warp's column part must be whole 16-column chunks");
    // feature switches: run-time for spec 0, compile-time constants otherwise
    const bool d2s = S.generic ? (e.store_mode == kStoreDepthToSpace) : (S.store == kStD2S);
    const bool s2d = S.generic ? (e.store_mode == kStoreSpaceToDepth) : (S.store == kStS2D);
    const bool s2dy = S.generic ? (e.store_mode == kStoreSpaceToDepthY) : false;
    const bool out_i8 = S.generic ? (e.out_i8 != 0) : S.i8;
    const bool has_out = S.generic ? (e.out != nullptr) : (S.store != kStNone);
    const bool has_noise = S.generic ? (e.noise != nullptr) : S.noise;
    const bool has_res = S.generic ? (e.residual != nullptr) : S.residual;
    const bool all_valid = S.generic ? (p.all_valid != 0) : true;
    const bool paired = S.generic ? (e.x_phases == 2) : false;   // pixel-pair rows: columns' halves are pixels 2x / 2x+1
    const int ppx = paired ? (half * 2) / kParts : 0;           // pixel of this warp's columns
    const int parts_per_sum = paired ? kParts / 2 : kParts;     // warps whose toRGB partial sums belong together
    const bool has_rgb = S.generic ? (e.rgb_w != nullptr) : S.rgb;
    // (kept in shared memory: as a loop-invariant register it is spilled and reloaded from local memory on the
    // critical path of every tile)
    if (et == 0) *nscale_slot = has_noise ? gain * __ldg(e.noise_strength) : 0.f;
    epi_bar_sync<32 * C::kEpiWarps>();
    // All index math below is 32-bit pixel arithmetic (pixel counts stay < 2^31); one 64-bit multiply per tile
    // turns a pixel index into an element offset.  Per-thread row offsets are loop invariants.
    const int W = p.W, H = p.H;
    const int row_reg = ry * W + rx;                                              // NHWC pixel offset inside the tile
    const int row_d2s = (2 * ry) * (2 * W) + 2 * rx;                              // depth-to-space: top-left output pixel
    const int row_s2d = (((ry >> 1) * (W >> 1) + (rx >> 1)) << 2) + ((ry & 1) * 2 + (rx & 1));
    const int cout_sh = e.cout_shift;       // log2(Cout) or -1
    const int ngrp_sh = e.noise_div_shift;  // log2(noise_group_div) or -1
    const bool cout_p2 = kPow2 ? true : (cout_sh >= 0);
    const bool ngrp_p2 = kPow2 ? true : (ngrp_sh >= 0);

    // Noise of the NEXT tile is fetched while the current one is processed: the (L2/DRAM) latency of this
    // scattered 4-byte load would otherwise sit on the critical path of every tile.
    auto fetch_noise = [&](const TileCoord& c2, bool in_range, float (&dst)[kPairM][kChunks]) {
#pragma unroll
      for (int h = 0; h < kPairM; ++h)
#pragma unroll
        for (int c = 0; c < kChunks; ++c) dst[h][c] = 0.f;
      if (!has_noise || !in_range) return;
      const int img2 = c2.tn * p.TN + ri, y2 = c2.ty * p.TH + ry, x2 = c2.tx * p.TW + rx;
      if (!all_valid && !(img2 < p.Nimg && y2 < H && x2 < W)) return;
      const int grp = ngrp_p2 ? (img2 >> ngrp_sh) : (img2 / e.noise_group_div);
      const float* base = e.noise + (size_t)grp * e.noise_group_stride;
      if (d2s) {
        const float* b2 = base + (size_t)(2 * y2) * (2 * W) + 2 * x2;
#pragma unroll
        for (int h = 0; h < kPairM; ++h)
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            const int n0 = c2.n_tile * BN + half * kHalf + c * 16;
            const int ph = cout_p2 ? (n0 >> cout_sh) : (n0 / e.Cout);
            dst[h][c] = __ldg(b2 + 16 * h + (ph >> 1) * (2 * W) + (ph & 1));
          }
      } else if (paired) {
        dst[0][0] = __ldg(base + (size_t)y2 * (2 * W) + 2 * x2 + ppx);
      } else {
#pragma unroll
        for (int h = 0; h < kPairM; ++h) dst[h][0] = __ldg(base + y2 * W + x2 + 8 * h);
      }
    };
    float nz_next[kPairM][kChunks];
    TileCoord tc_next = decode_tile<kPow2>(p, blockIdx.x, n_tiles);
    fetch_noise(tc_next, true, nz_next);
    int it = 0;
    int staged_img = -1, staged_ntile = -1;
    int pbuf = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const TileCoord tc = tc_next;
      const int n_tile = tc.n_tile, tn = tc.tn;
      const int img = tn * p.TN + ri, y = tc.ty * p.TH + ry, x = tc.tx * p.TW + rx;
      const bool valid = all_valid || (img < p.Nimg && y < H && x < W);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * C::kAccCols + half * kHalf;
      float rgb[kPairM][3];
      float nz_cur[kPairM][kChunks];
#pragma unroll
      for (int h = 0; h < kPairM; ++h) {
        rgb[h][0] = rgb[h][1] = rgb[h][2] = 0.f;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) nz_cur[h][c] = nz_next[h][c];
      }
      {
        const int nt = tile + gridDim.x;
        const bool more = nt < total_tiles;
        if (more) tc_next = decode_tile<kPow2>(p, nt, n_tiles);
        if (fast) fetch_noise(tc_next, more, nz_next);
      }

      if (fast) {
        // ---- stage the per-(image, n_tile) parameters once; reuse while they do not change ----
        const int timg = tn;                 // TN == 1
        if (timg != staged_img || n_tile != staged_ntile) {
          pbuf ^= 1;
          float* par = params + pbuf * (kNumParams * BN);
          if (et < BN) {                      // (32*kEpiWarps >= BN for every instantiation)
            const int n = n_tile * BN + et;
            const int o = cout_p2 ? (n & (e.Cout - 1)) : (n % e.Cout);
            const float d = e.dmod != nullptr ? __ldg(e.dmod + (size_t)timg * e.Cout + o) : 1.f;
            const float b = e.bias != nullptr ? __ldg(e.bias + o) : 0.f;
            const float osn = e.out_scale != nullptr ? __ldg(e.out_scale + (size_t)timg * e.out_scale_stride + o) : 1.f;
            par[0 * BN + et] = d * gain;
            par[1 * BN + et] = b * gain;
            par[2 * BN + et] = osn * e.post_scale;
            if (has_rgb) {
              const float* rw = e.rgb_w + (size_t)timg * 3 * e.Cout + o;
              par[3 * BN + et] = __ldg(rw);
              par[4 * BN + et] = __ldg(rw + e.Cout);
              par[5 * BN + et] = __ldg(rw + 2 * e.Cout);
            }
          }
          staged_img = timg;
          staged_ntile = n_tile;
          epi_bar_sync<32 * C::kEpiWarps>();  // all epilogue warps take the same branch (uniform condition)
        }
        const float* par = params + pbuf * (kNumParams * BN);
        // ---- per-tile addresses (tile 0 of a pair; tile h adds 8 pixel columns) ----
        const int n_first = n_tile * BN + half * kHalf;
        const int pix = (img * H + tc.ty * p.TH) * W + tc.tx * p.TW + row_reg;          // NHWC pixel index of this row
        const __half* res_row = has_res ? e.residual + (size_t)pix * p.Ntot + n_first : nullptr;
        __half* out_row = nullptr;           // regular / space-to-depth: contiguous columns
        size_t out_row_hstep = 0;            // element offset of the pair's second tile
        int d2s_pix = 0;
        if (has_out) {
          if (s2d && !out_i8) {
            const int org = ((img * (H >> 1) + ((tc.ty * p.TH) >> 1)) * (W >> 1) + ((tc.tx * p.TW) >> 1)) << 2;
            out_row = e.out + (size_t)(org + row_s2d) * p.Ntot + n_first;
            out_row_hstep = (size_t)16 * p.Ntot;      // 8 pixels right = 4 cells x 4 phases
          } else if (s2dy) {
            const int org = (img * (H >> 1) + ((tc.ty * p.TH + ry) >> 1)) * W + tc.tx * p.TW + rx;
            out_row = e.out + ((size_t)org * 2 + (ry & 1)) * p.Ntot + n_first;
            out_row_hstep = (size_t)16 * p.Ntot;
          } else if (!d2s && !out_i8) {
            out_row = e.out + (size_t)pix * p.Ntot + n_first;
            out_row_hstep = (size_t)8 * p.Ntot;
          }
        }
        if (d2s) d2s_pix = (img * 2 * H + 2 * tc.ty * p.TH) * (2 * W) + 2 * tc.tx * p.TW + row_d2s;
        // I8 layout [n][y][c/8][x][8]: (row index) * (C/8) * Wo + x, in units of 8-channel vectors
        const int i8_groups = e.Cout >> 3;
        const int i8_row = img * H + tc.ty * p.TH + ry;          // regular store: image row of this pixel
        const int i8_x0 = tc.tx * p.TW + rx;
        // residual operand of chunk c of tile h: NHWC [pix][Ntot], or I8 [n][y][Ntot/8][x][8] (EpiParams.res_i8: the D
        // projections store it that way so that both their stores and these loads are 128-byte contiguous per 8 lanes)
        const bool res_i8 = has_res && e.res_i8 != 0;
        const size_t res_stride = res_i8 ? (size_t)W * 8 : 8;      // halfs from channels [0,8) to [8,16) of a chunk
        auto res_addr = [&](int h, int c) -> const __half* {
          if (!has_res) return nullptr;
          if (res_i8)
            return e.residual + (((size_t)i8_row * (p.Ntot >> 3) + ((n_first + c * 16) >> 3)) * W + i8_x0 + 8 * h) * 8;
          return res_row + (size_t)(8 * h) * p.Ntot + c * 16;
        };
        float nscale;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(nscale) : "r"(smem_u32(nscale_slot)));
        // Specialised residual layers fetch the whole tile's residual values NOW, before waiting for the accumulator:
        // loaded chunk by chunk inside the math, every chunk exposed a full DRAM round trip and the epilogue, not the
        // tensor pipe, paced the D conv1 layers (tensor pipe 36 % active in the 512^2 down-conv).
        constexpr bool kResPre = !S.generic && S.residual;
        uint4 resv[kResPre ? kPairM : 1][kResPre ? kChunks : 1][2];
        if (kResPre) {
#pragma unroll
          for (int h = 0; h < kPairM; ++h)
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
              const __half* rp = res_addr(h, c);
              resv[kResPre ? h : 0][kResPre ? c : 0][0] = __ldg(reinterpret_cast<const uint4*>(rp));
              resv[kResPre ? h : 0][kResPre ? c : 0][1] = __ldg(reinterpret_cast<const uint4*>(rp + res_stride));
            }
        }
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
        // output address of chunk c of tile h of the pair (nullptr: nothing stored)
        auto out_addr = [&](int h, int c, size_t& half_stride) -> __half* {
          const int i8_x = i8_x0 + 8 * h;
          const int j0 = half * kHalf + c * 16;
          half_stride = 8;
          if (d2s) {
            // column n -> phase (py,px) and channel o; output pixel (2y+py, 2x+px)
            const int n0 = n_tile * BN + j0;
            const int ph = cout_p2 ? (n0 >> cout_sh) : (n0 / e.Cout);
            const int o0 = n0 - ph * e.Cout;
            if (!has_out) return nullptr;
            if (out_i8) {
              const int yo = 2 * (tc.ty * p.TH + ry) + (ph >> 1), xo = 2 * i8_x + (ph & 1);
              half_stride = (size_t)(2 * W) * 8;
              return e.out + (((size_t)(img * 2 * H + yo) * i8_groups + (o0 >> 3)) * (2 * W) + xo) * 8;
            }
            return e.out + (size_t)(d2s_pix + 16 * h + (ph >> 1) * (2 * W) + (ph & 1)) * e.Cout + o0;
          }
          if (out_i8 && s2d && has_out) {
            // space-to-depth + I8: [n][y/2][(phase*Ntot + o)/8][x/2][8] with 4*Ntot channels per cell
            const int yy = tc.ty * p.TH + ry;
            const int k0 = ((yy & 1) * 2 + (i8_x & 1)) * p.Ntot + n_first + c * 16;
            half_stride = (size_t)(W >> 1) * 8;
            return e.out + (((size_t)(img * (H >> 1) + (yy >> 1)) * (p.Ntot >> 1) + (k0 >> 3)) * (W >> 1) + (i8_x >> 1)) * 8;
          }
          if (out_i8 && has_out) {
            const int o0 = n_first + c * 16;                     // regular store: Ntot == Cout
            half_stride = (size_t)W * 8;
            return e.out + (((size_t)i8_row * i8_groups + (o0 >> 3)) * W + i8_x) * 8;
          }
          if (out_row != nullptr) return out_row + h * out_row_hstep + c * 16;
          return nullptr;
        };
        constexpr int kActT = S.generic ? -1 : S.act;
        constexpr bool kNz = S.generic || S.noise;      // specialised layers without noise: no per-element add
        constexpr int kRoundT = S.generic ? -1 : (S.round ? 1 : 0);
        // specialised pair layers: both tiles of the pair per chunk, parameters read once (epilogue_fastN<2>)
        constexpr bool kPairFused = (kPairM == 2) && !S.generic;
        if constexpr (kPairFused) {
          uint32_t accp[2][2][16];                     // [pipeline buffer][tile of the pair]
          tc_ld16_issue(taddr, accp[0][0]);
          tc_ld16_issue(taddr + BN, accp[0][1]);
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            tc_ld_wait();
            if (c + 1 < kChunks) {
              tc_ld16_issue(taddr + (c + 1) * 16, accp[(c + 1) & 1][0]);
              tc_ld16_issue(taddr + BN + (c + 1) * 16, accp[(c + 1) & 1][1]);
            }
            if (valid && !(kDebugBuild && (p.debug_skip & 1))) {
              const int j0 = half * kHalf + c * 16;
              size_t half_stride = 8, hs1 = 8;
              __half* optr[2] = {out_addr(0, c, half_stride), out_addr(1, c, hs1)};
              const float nzc[2] = {nscale * nz_cur[0][d2s ? c : 0], nscale * nz_cur[1][d2s ? c : 0]};
              const __half* rptr[2];
              const uint4* rpre[2];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                rptr[h] = !kResPre ? res_addr(h, c) : nullptr;
                rpre[h] = kResPre ? resv[kResPre ? h : 0][kResPre ? c : 0] : nullptr;
              }
              if (has_rgb) epilogue_fastN<2, true, kActT, kNz, kRoundT>(e, par, BN, j0, accp[c & 1], nzc, rptr, optr, half_stride, rgb, rpre, res_stride);
              else epilogue_fastN<2, false, kActT, kNz, kRoundT>(e, par, BN, j0, accp[c & 1], nzc, rptr, optr, half_stride, rgb, rpre, res_stride);
            }
          }
        } else {
        uint32_t acc[2][16];
        tc_ld16_issue(taddr, acc[0]);
#pragma unroll
        for (int h = 0; h < kPairM; ++h) {
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            constexpr int kTotal = kPairM * kChunks;
            const int ci = h * kChunks + c;            // position in the LDTM pipeline
            tc_ld_wait();
            if (ci + 1 < kTotal) {
              const int h2 = (ci + 1) / kChunks, c2 = (ci + 1) - h2 * kChunks;
              tc_ld16_issue(taddr + h2 * BN + c2 * 16, acc[(ci + 1) & 1]);
            }
            if (valid && !(kDebugBuild && (p.debug_skip & 1))) {
              const int j0 = half * kHalf + c * 16;
              const float nzc = nscale * nz_cur[h][d2s ? c : 0];
              size_t half_stride = 8;
              __half* optr = out_addr(h, c, half_stride);
              const __half* rptr = !kResPre ? res_addr(h, c) : nullptr;
              const uint4* rpre = kResPre ? resv[kResPre ? h : 0][kResPre ? c : 0] : nullptr;
              if (has_rgb) epilogue_fast16<true, kActT, kNz, kRoundT>(e, par, BN, j0, acc[ci & 1], nzc, rptr, optr, half_stride, rgb[h], rpre, res_stride);
              else epilogue_fast16<false, kActT, kNz, kRoundT>(e, par, BN, j0, acc[ci & 1], nzc, rptr, optr, half_stride, rgb[h], rpre, res_stride);
            }
          }
        }
        }
      } else {
        // ---- generic path (tiles that span several images: 4x4 / 8x8 layers; never a tile pair) ----
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
          float v[16];
          tc_ld16(taddr + c * 16, v);
          if (valid) epilogue_row16(p, img, y, x, n_tile * BN + half * kHalf + c * 16, v, rgb[0]);
        }
      }
      if (has_rgb) {
        // The warps of a lane quarter own different column parts of the same rows: the partial toRGB sums meet in
        // shared memory so that one float4 per pixel goes to HBM.  (Pixel-pair rows: two sum groups, one per pixel.)
        float4* stg = rgb_stage + (it & 1) * ((kParts - 1) * kPairM * 128);
        // finish pixel (tile h of the pair) with its complete toRGB sum
        auto finish_pixel = [&](int h, float r0, float r1, float r2) {
          if (paired) {
            const size_t pix2 = ((size_t)img * H + y) * (2 * W) + 2 * x + ppx;
            e.rgb_out[pix2] = make_float4(r0, r1, r2, 0.f);
          } else if (e.image != nullptr) {
            // final image: y = up2(yprev) + toRGB + bias (modules.py:580-602 polyphase, models.py:1004-1013), then
            // clip((y + 1) / 2, 0, 1) (utils.py:14-17) -- the arithmetic of rgb_combine_kernel
            const int X = x + 8 * h, Y = y;
            const int Wp = W >> 1, zy = Y >> 1, zx = X >> 1;
            const float wy0 = (Y & 1) ? 0.25f : 0.75f, wy1 = 1.f - wy0;
            const float wx0 = (X & 1) ? 0.25f : 0.75f, wx1 = 1.f - wx0;
            // (Fetching these four pixels ahead of the accumulator wait -- into registers, or with cp.async into
            // per-thread shared-memory slots -- was measured neutral / slower: DESIGN.md 7.0.)
            const float4* base = e.img_yprev + ((img * (H >> 1) + zy) * Wp + zx);      // (pixel counts stay < 2^31)
            float4 a = make_float4(0, 0, 0, 0), c = a, d = a;
            const float4 ee = __ldg(base);
            if (zy > 0 && zx > 0) a = __ldg(base - Wp - 1);
            if (zy > 0) c = __ldg(base - Wp);
            if (zx > 0) d = __ldg(base - 1);
            const float r = __fadd_rn(__fadd_rn(__ldg(e.img_bias), r0), skip_up2(wy0, wy1, wx0, wx1, a.x, c.x, d.x, ee.x));
            const float g = __fadd_rn(__fadd_rn(__ldg(e.img_bias + 1), r1), skip_up2(wy0, wy1, wx0, wx1, a.y, c.y, d.y, ee.y));
            const float bl = __fadd_rn(__fadd_rn(__ldg(e.img_bias + 2), r2), skip_up2(wy0, wy1, wx0, wx1, a.z, c.z, d.z, ee.z));
            const size_t plane = (size_t)H * W;
            float* ip = e.image + (size_t)img * 3 * plane + (size_t)(Y * W + X);
            ip[0] = image_value(r);
            ip[plane] = image_value(g);
            ip[2 * plane] = image_value(bl);
          } else {
            const size_t pix = ((size_t)img * H + y) * W + x + 8 * h;
            e.rgb_out[(size_t)n_tile * p.Nimg * H * W + pix] = make_float4(r0, r1, r2, 0.f);
          }
        };
        // Tile pairs with two column parts: part k finishes tile k of the pair (it hands its partial sums of the
        // other tile over and takes the other part's sums of its own), so both warps of a lane quarter carry the
        // same load.  With all pixels finished by the leading part, the other warp idled through the ~140
        // instructions per pixel of the final-image arithmetic of the last generator conv.
        constexpr bool kSplitFinish = (kPairM == 2 && kParts == 2);
        if (kSplitFinish && !paired) {
          const int other = half ^ 1;
          stg[half * 128 + row] = half ? make_float4(rgb[0][0], rgb[0][1], rgb[0][2], 0.f)
                                       : make_float4(rgb[kPairM - 1][0], rgb[kPairM - 1][1], rgb[kPairM - 1][2], 0.f);
          asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "n"(32 * kParts) : "memory");
          if (valid) {
            const float4 t = stg[other * 128 + row];
            const float m0 = half ? rgb[kPairM - 1][0] : rgb[0][0], m1 = half ? rgb[kPairM - 1][1] : rgb[0][1];
            const float m2 = half ? rgb[kPairM - 1][2] : rgb[0][2];
            finish_pixel(half, m0 + t.x, m1 + t.y, m2 + t.z);
          }
        } else {
          const int lead = (half / parts_per_sum) * parts_per_sum;
          if (parts_per_sum > 1) {
            if (half != lead) {
#pragma unroll
              for (int h = 0; h < kPairM; ++h)
                stg[((half - 1) * kPairM + h) * 128 + row] = make_float4(rgb[h][0], rgb[h][1], rgb[h][2], 0.f);
            }
            asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "n"(32 * kParts) : "memory");
          }
          if (half == lead && valid) {
#pragma unroll
            for (int h = 0; h < kPairM; ++h) {
              float r0 = rgb[h][0], r1 = rgb[h][1], r2 = rgb[h][2];
              for (int o = 1; o < parts_per_sum; ++o) {
                const float4 t = stg[((lead + o - 1) * kPairM + h) * 128 + row];
                r0 += t.x; r1 += t.y; r2 += t.z;
              }
              finish_pixel(h, r0, r1, r2);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
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

template <int BN, int BK, int MODE, int EPI = 0>
cudaError_t launch_one(const ConvParams& p, const TmaMaps& maps, int num_sms, cudaStream_t s) {
  using C = Cfg<BN, BK, MODE>;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(conv_tc_kernel<BN, BK, MODE, EPI>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (err != cudaSuccess) return err;
    configured = true;
  }
  const int n_tiles = p.Ntot / BN;
  const int total = p.tiles_n * p.tiles_y * p.tiles_x * n_tiles;
  const int ctas = num_sms * C::kMinBlocks;
  int grid = total < ctas ? total : ctas;
  if (MODE != 0) grid = (grid / n_tiles) * n_tiles;   // keeps tile % n_tiles constant per CTA (resident taps)
  if (grid <= 0) return cudaErrorInvalidValue;
  if (p.skip_mode == 1) {
    ConvParams q = p;
    q.rot_div = grid / n_tiles > 0 ? grid / n_tiles : 1;    // m advances by grid / n_tiles per persistent iteration
    conv_tc_kernel<BN, BK, MODE, EPI><<<grid, C::kThreads, C::kSmemBytes, s>>>(maps.a, maps.b, q);
    return cudaGetLastError();
  }
  conv_tc_kernel<BN, BK, MODE, EPI><<<grid, C::kThreads, C::kSmemBytes, s>>>(maps.a, maps.b, p);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// SIMT bring-up kernel: one thread per (pixel, 16 columns).
// ---------------------------------------------------------------------------
__global__ void conv_simt_kernel(const ConvParams p) {
  const int groups = p.Ntot / 16;
  const size_t total = (size_t)p.Nimg * p.H * p.W * groups;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(idx % groups);
    size_t pix = idx / groups;
    const int x = (int)(pix % p.W); pix /= p.W;
    const int y = (int)(pix % p.H);
    const int img = (int)(pix / p.H);
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    for (int tap = 0; tap < p.taps; ++tap) {
      const int yy = y + p.tap_dy[tap], xx = x + p.tap_dx[tap];
      if (yy < 0 || yy >= p.in_H || xx < 0 || xx >= p.in_W) continue;
      const __half* a = p.in + ((size_t)(img * p.in_H + yy) * p.in_W + xx) * p.Cin;
      const __half* w = p.wgt + ((size_t)tap * p.Ntot + g * 16) * p.Cin;
      for (int c = 0; c < p.Cin; c += 2) {
        const float2 av = __half22float2(*reinterpret_cast<const __half2*>(a + c));
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 wv = __half22float2(*reinterpret_cast<const __half2*>(w + (size_t)j * p.Cin + c));
          acc[j] = fmaf(av.x, wv.x, acc[j]);
          acc[j] = fmaf(av.y, wv.y, acc[j]);
        }
      }
    }
    // the SIMT kernel owns whole rows only when Ntot == 16*groups handled by one thread; toRGB partials
    // are therefore accumulated per 16-column group with atomics into slab 0 (bring-up path only).
    float rgb[3] = {0.f, 0.f, 0.f};
    epilogue_row16(p, img, y, x, g * 16, acc, rgb);
    if (p.epi.rgb_w != nullptr) {
      const size_t opix = ((size_t)img * p.H + y) * p.W + x;
      const int slab = (g * 16) / p.BN;
      float* dst = reinterpret_cast<float*>(p.epi.rgb_out + (size_t)slab * p.Nimg * p.H * p.W + opix);
      atomicAdd(dst + 0, rgb[0]);
      atomicAdd(dst + 1, rgb[1]);
      atomicAdd(dst + 2, rgb[2]);
    }
  }
}

}  // namespace

// The specialised epilogue whose compile-time switches equal this layer's run-time ones, or 0.
static int pick_epi_spec(const ConvParams& p) {
  const EpiParams& e = p.epi;
  static const bool off = debug_env("GLASS_DEBUG_GENERIC_EPI") != nullptr;     // A/B knob: always the run-time spec
  if (off || p.TN != 1 || !p.all_valid || p.debug_skip != 0 || p.skip_mode == 1 || e.x_phases == 2) return 0;
  // power-of-two geometry (tile grid, channel count, noise group): required by the specs that decode with shifts
  const bool pow2 = p.pow2 && e.cout_shift >= 0 && (e.noise == nullptr || e.noise_div_shift >= 0);
  int store = kStNone;
  if (e.out != nullptr) {
    if (e.store_mode == kStoreRegular) store = kStRegular;
    else if (e.store_mode == kStoreDepthToSpace) store = kStD2S;
    else if (e.store_mode == kStoreSpaceToDepth) store = kStS2D;
    else return 0;
  }
  for (int i = 1; i < kNumEpiSpecs; ++i) {
    const EpiSpec& sp = kEpiSpecs[i];
    if (sp.noise == (e.noise != nullptr) && sp.rgb == (e.rgb_w != nullptr) && sp.residual == (e.residual != nullptr) &&
        sp.store == store && (store == kStNone || sp.i8 == (e.out_i8 != 0)) && sp.act == e.act &&
        sp.round == (e.round_fp16_before_act != 0) && (!sp.pow2 || pow2) && (sp.pow2 || !e.res_i8))
      return i;
  }
  return 0;
}

cudaError_t launch_conv_tc(const ConvParams& p, const TmaMaps& maps, int num_sms, cudaStream_t s) {
  const int spec = pick_epi_spec(p);
  static const bool log_spec = debug_env("GLASS_DEBUG_SPEC_LOG") != nullptr;
  if (log_spec)
    fprintf(stderr, "conv_tc: H=%d W=%d Cin=%d taps=%d Ntot=%d BN=%d BK=%d mode=%d spec=%d\n", p.H, p.W, p.Cin, p.taps,
            p.Ntot, p.BN, p.BK, p.mode, spec);
#define GLASS_SPEC(bn, bk, md, sp) \
  if (p.BN == bn && p.BK == bk && p.mode == md && spec == sp) return launch_one<bn, bk, md, sp>(p, maps, num_sms, s);
  GLASS_SPEC(64, 64, 4, 1)
  GLASS_SPEC(64, 64, 4, 2)
  GLASS_SPEC(32, 32, 4, 3)
  GLASS_SPEC(32, 32, 4, 4)
  GLASS_SPEC(64, 64, 4, 4)
  GLASS_SPEC(128, 64, 0, 5)
  GLASS_SPEC(256, 64, 0, 2)
  GLASS_SPEC(256, 64, 0, 6)
  GLASS_SPEC(64, 64, 0, 7)
  GLASS_SPEC(128, 64, 0, 8)
  GLASS_SPEC(256, 64, 0, 8)
  GLASS_SPEC(64, 32, 2, 9)
  GLASS_SPEC(128, 64, 2, 9)
  GLASS_SPEC(256, 64, 0, 9)
  GLASS_SPEC(64, 32, 2, 12)
  GLASS_SPEC(128, 64, 2, 12)
  GLASS_SPEC(256, 64, 0, 12)
  GLASS_SPEC(128, 64, 0, 10)
  GLASS_SPEC(32, 32, 4, 11)
  GLASS_SPEC(32, 32, 4, 13)
  GLASS_SPEC(64, 64, 4, 13)
  GLASS_SPEC(64, 128, 6, 7)
  GLASS_SPEC(32, 128, 4, 7)
  GLASS_SPEC(64, 128, 6, 8)
  GLASS_SPEC(128, 64, 0, 14)
  GLASS_SPEC(128, 64, 0, 15)
  GLASS_SPEC(128, 64, 0, 16)
#undef GLASS_SPEC
#define GLASS_CASE(bn, bk, md) \
  if (p.BN == bn && p.BK == bk && p.mode == md) return launch_one<bn, bk, md>(p, maps, num_sms, s);
  GLASS_CASE(32, 32, 0)
  GLASS_CASE(32, 64, 0)
  GLASS_CASE(64, 64, 0)
  GLASS_CASE(128, 64, 0)
  GLASS_CASE(256, 64, 0)
  GLASS_CASE(64, 32, 0)
  GLASS_CASE(128, 32, 0)
  GLASS_CASE(256, 32, 0)
  GLASS_CASE(32, 32, 1)
  GLASS_CASE(64, 32, 1)
  GLASS_CASE(128, 32, 1)
  GLASS_CASE(32, 64, 1)
  GLASS_CASE(64, 64, 1)
  GLASS_CASE(32, 32, 4)
  GLASS_CASE(64, 32, 4)
  GLASS_CASE(128, 32, 4)
  GLASS_CASE(32, 64, 4)
  GLASS_CASE(64, 64, 4)
  GLASS_CASE(32, 128, 4)
  GLASS_CASE(64, 128, 6)
  GLASS_CASE(32, 32, 2)
  GLASS_CASE(64, 32, 2)
  GLASS_CASE(128, 32, 2)
  GLASS_CASE(32, 64, 2)
  GLASS_CASE(64, 64, 2)
  GLASS_CASE(128, 64, 2)
  GLASS_CASE(256, 64, 2)
#undef GLASS_CASE
  return cudaErrorInvalidValue;
}

cudaError_t launch_conv_simt(const ConvParams& p, cudaStream_t s) {
  const size_t total = (size_t)p.Nimg * p.H * p.W * (p.Ntot / 16);
  int blocks = (int)((total + 127) / 128);
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (p.epi.rgb_w != nullptr) {
    // bring-up path accumulates toRGB partials with atomics: clear the slabs first
    cudaError_t err = cudaMemsetAsync(p.epi.rgb_out, 0,
                                      sizeof(float4) * (size_t)(p.Ntot / p.BN) * p.Nimg * p.H * p.W, s);
    if (err != cudaSuccess) return err;
  }
  conv_simt_kernel<<<blocks, 128, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace glass
