// Discriminator down-conv (stylegan2/modules.py:1204-1254: FilterLayer [1,3,3,1]x[1,3,3,1]/64, pad 2 -> 3x3 conv,
// stride 2) in its EXACT form with the FIR applied INSIDE the kernel — for the 32/64-channel blocks at 1024^2 / 512^2,
// where the FIR-folded 3x3-over-space-to-depth form issues 4x the MACs and a separate blur pass costs more HBM time
// than the inflation.
//
//   warp 0        TMA producer: the haloed raw tile of conv0's output (I8 layout [N][H][C/8][W][8]; 36 rows x 20 pixels x
//                 32 channels per stage, zero-filled outside the image = the FIR's zero padding) and, once, the nine
//                 3x3 taps of the n-tile (resident, [tap][BN][C], swizzled K-major).
//   warp 1        MMA issuer: out[z] = sum_{ky,kx} W[ky,kx] . U[2z+ky, 2z+kx] as 9 x C/16 tcgen05.mma over the blurred tile,
//                 which the blur warps leave in shared memory split by row/column parity so that a stride-2 tap is a
//                 plain descriptor offset (un-swizzled K-major core matrices, 8 pixels x 16 B).
//   warps 2..9    epilogue: TMEM -> bias, lrelu*sqrt2, + projection residual, * 1/sqrt2 -> fp16 store (I8 or NHWC).
//   warps 10..17  blur: U = FIR(a) in packed-half2 arithmetic, (c0+c3)/8 + 3(c1+c2)/8 per axis (separable; ~4
//                 thread-instructions per element against ~10 for an epilogue), written as four parity planes.
//
// Accumulator tile: 8 wide x 16 tall output pixels; channels are processed in 32-channel passes (C = 64: two passes
// accumulate into the same TMEM columns).
//
// kProj (the 32 -> 64 block): the block's projection path (1x1 conv of the FIR-downsampled block input,
// stylegan2/modules.py:1352-1372, the residual operand of this conv) is a SECOND accumulator of the same tile: TMA
// brings the 128 x 32 tile of the downsampled input next to the raw tile, two more MMAs (K = 32) run against the
// resident projection weights, and the epilogue adds the fp32 result -- the projection GEMM launch and the round trip
// of its fp16 output tensor through HBM (2.1 GB written and read back at P = 64) disappear.
#include "common.cuh"
#include "kernels.cuh"
#include "tcgen05.cuh"

namespace glass {
namespace {

constexpr int kDcTW = 8, kDcTH = 16;                    // output tile (z space)
constexpr int kDcRawRows = 2 * kDcTH + 4, kDcRawCols = 2 * kDcTW + 4;     // 36 x 20 raw pixels
constexpr int kDcLbo = 9 * 16;                          // plane: next 8-channel group (9 columns of 16 B)
constexpr int kDcThreads = 64 + 256 + 256;

struct DownParams {
  int N, Ho, Wo, Cout;
  int sh_x, sh_y;         // log2 of the tile grid (Wo / 8, Ho / 16 are powers of two): tile decode with shifts
  const float* bias;
  const __half* residual;
  int res_i8;
  __half* out;
  int out_i8;
  float post_scale;
};

constexpr int kDcXdBytes = 128 * 32 * 2;                 // kProj: the tile of the downsampled block input (128 pixels x 32 ch)
template <int C, int BN, bool kProj = false>
struct DCfg {
  static_assert(!kProj || (C == 32 && BN == 64), "the projection accumulator exists for the 32 -> 64 block");
  static constexpr int kProjBytes = kProj ? 2 * kDcXdBytes + BN * C * 2 : 0;      // two tile stages + the 1x1 weights
  static constexpr int kProjPad = kProj ? 1024 : 0;                               // (alignment of those operands)
  // 8-channel groups per pass.  BN = 64: 32-channel passes.  BN = 128 (the 64 -> 128 block): all 128 output columns of a
  // pixel tile come from ONE blur of that tile (with 64-column n-tiles every tile was loaded and blurred once per
  // n-tile: 2.5 ms, the slowest launch of the step); the nine 128-column taps take 147 KB of shared memory, so the
  // raw / blurred tiles shrink to 16-channel passes (one K = 16 MMA step per tap and pass).
  static constexpr int kG = (BN == 128) ? 2 : 4;
  static constexpr int kPasses = C / (8 * kG);
  static constexpr int kStepsPerPass = kG / 2;             // K = 16 MMA steps per tap and pass
  static constexpr int kRawBytes = kDcRawRows * kG * kDcRawCols * 16;      // 46080 / 23040
  // Blurred tile: four parity planes (row parity py, column parity px), each [row][8-channel group][column] x 16 B.
  // kTrim (BN = 128): the odd planes drop the row / column they never hold (33 x 17 blurred positions: 17 even and
  // 16 odd rows, 9 even and 8 odd columns), which is what lets TWO blurred stages fit beside 147 KB of taps, so that
  // the blur of pass h+1 overlaps the MMAs of pass h (with one stage the 16 worker warps spent most of their stall
  // samples on the pass barrier: profiles/).
  static constexpr bool kTrim = (BN == 128);
  __host__ __device__ static constexpr int plane_cols(int px) { return (kTrim && px) ? 8 : 9; }
  __host__ __device__ static constexpr int plane_rows(int py) { return (kTrim && py) ? kDcTH : kDcTH + 1; }
  __host__ __device__ static constexpr int plane_lbo(int px) { return plane_cols(px) * 16; }       // next channel group
  __host__ __device__ static constexpr int plane_sbo(int px) { return kG * plane_lbo(px); }        // next row
  __host__ __device__ static constexpr int plane_bytes(int py, int px) { return plane_rows(py) * plane_sbo(px); }
  __host__ __device__ static constexpr int plane_off(int py, int px) {
    return (py ? plane_bytes(0, 0) + plane_bytes(0, 1) : 0) + (px ? plane_bytes(py, 0) : 0);
  }
  static constexpr int kBlurBytes = plane_off(1, 1) + plane_bytes(1, 1);       // 39168 / 17952
  static constexpr int kRawStages = 2;
  static constexpr int kBlurStages = (C == 32 || BN == 128) ? 2 : 1;
  static constexpr int kWBytes = 9 * BN * C * 2;
  static constexpr int kTapBytes = BN * C * 2;
  static constexpr int kSmemBytes =
      kWBytes + kRawStages * kRawBytes + kBlurStages * kBlurBytes + kProjBytes + kProjPad + 1024 + 256 + BN * 4;
  static constexpr int kTmemCols = (kProj ? 4 : 2) * BN < 32 ? 32 : (kProj ? 4 : 2) * BN;
  static constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// mbarrier wait for this kernel's polling warps: they share their schedulers with the blur and epilogue warps, and a
// spinning poll loop (try_wait + clock check, ~12 instructions per iteration) took 10 % of all issued instructions.
// Back off between polls; still bounded (a protocol bug must trap, not hang).
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(40);
    if (++spins > 40000000u) {
      printf("glass downconv_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// [1,3,3,1] on 8 channels (packed half2).  kScaled = false: (c0 + c3) + 3 (c1 + c2), 3 instructions per half2 (the
// horizontal pass: the result is at most 8 |a|); kScaled = true: the same times 1/64 (the vertical pass applies the
// whole normalisation of the separable filter, 1/8 per axis, so that the blurred value is back at the scale of a).
template <bool kScaled>
__device__ __forceinline__ uint4 fir4(const uint4& c0, const uint4& c1, const uint4& c2, const uint4& c3) {
  const __half2 k1 = __floats2half2_rn(1.f / 64.f, 1.f / 64.f), k3 = __floats2half2_rn(3.f / 64.f, 3.f / 64.f);
  const __half2 three = __floats2half2_rn(3.f, 3.f);
  uint4 r;
  const __half2* a = reinterpret_cast<const __half2*>(&c0);
  const __half2* b = reinterpret_cast<const __half2*>(&c1);
  const __half2* c = reinterpret_cast<const __half2*>(&c2);
  const __half2* d = reinterpret_cast<const __half2*>(&c3);
  __half2* o = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (kScaled) o[j] = __hfma2(__hadd2(b[j], c[j]), k3, __hmul2(__hadd2(a[j], d[j]), k1));
    else o[j] = __hfma2(__hadd2(b[j], c[j]), three, __hadd2(a[j], d[j]));
  }
  return r;
}

// the same on 4 channels (two half2)
template <bool kScaled>
__device__ __forceinline__ uint2 fir4h(const uint2& c0, const uint2& c1, const uint2& c2, const uint2& c3) {
  const __half2 k1 = __floats2half2_rn(1.f / 64.f, 1.f / 64.f), k3 = __floats2half2_rn(3.f / 64.f, 3.f / 64.f);
  const __half2 three = __floats2half2_rn(3.f, 3.f);
  uint2 r;
  const __half2* a = reinterpret_cast<const __half2*>(&c0);
  const __half2* b = reinterpret_cast<const __half2*>(&c1);
  const __half2* c = reinterpret_cast<const __half2*>(&c2);
  const __half2* d = reinterpret_cast<const __half2*>(&c3);
  __half2* o = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (kScaled) o[j] = __hfma2(__hadd2(b[j], c[j]), k3, __hmul2(__hadd2(a[j], d[j]), k1));
    else o[j] = __hfma2(__hadd2(b[j], c[j]), three, __hadd2(a[j], d[j]));
  }
  return r;
}

template <int C, int BN, bool kProj>
__global__ void __launch_bounds__(kDcThreads, 1)
downconv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                   const __grid_constant__ CUtensorMap map_xd, const __grid_constant__ CUtensorMap map_wp,
                   const DownParams p) {
  using Cf = DCfg<C, BN, kProj>;
  // Worker-warp organisation, chosen by measurement at P = 64 (profiles/): dedicated blur / epilogue warps for the
  // 32-channel block (2.20 ms against 2.6 ms merged), merged roles for the 64-channel block, whose two 32-channel
  // passes per tile double the blur work per epilogue (2.5 ms against 3.3 ms with dedicated warps).
  constexpr bool kSplitRoles = (C == 32);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_w = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* raw = smem_w + Cf::kWBytes;
  uint8_t* blur = raw + Cf::kRawStages * Cf::kRawBytes;
  // kProj: [2 stages][128 pixels][32 ch], 64B-swizzled rows, on a 1024-byte boundary like every swizzled operand
  uint8_t* xd = smem_w + ((Cf::kWBytes + Cf::kRawStages * Cf::kRawBytes + Cf::kBlurStages * Cf::kBlurBytes + 1023) & ~1023);
  uint8_t* wproj = xd + 2 * kDcXdBytes;                      // kProj: [BN][C] K-major, 64B-swizzled
  uint64_t* bars = reinterpret_cast<uint64_t*>(kProj ? xd + Cf::kProjBytes : blur + Cf::kBlurStages * Cf::kBlurBytes);
  uint64_t* raw_full = bars;
  uint64_t* raw_empty = raw_full + 2;
  uint64_t* blur_full = raw_empty + 2;
  uint64_t* blur_empty = blur_full + 2;
  uint64_t* tmem_full = blur_empty + 2;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* w_bar = tmem_empty + 2;
  uint64_t* xd_empty = w_bar + 1;                            // kProj: the tile stage has been read by its MMAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xd_empty + 2);
  float* bias_s = reinterpret_cast<float*>(bars) + 64;        // [BN] bias * sqrt2 * post_scale of this CTA's n-tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.Cout / BN;
  const int tiles_x = p.Wo / kDcTW, tiles_y = p.Ho / kDcTH;
  const int total_tiles = p.N * tiles_y * tiles_x * n_tiles;
  const int n_tile = blockIdx.x % n_tiles;            // grid is a multiple of n_tiles: fixed per CTA (resident taps)

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_w);
    if (kProj) {
      prefetch_tmap(&map_xd);
      prefetch_tmap(&map_wp);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], kSplitRoles ? 8 : 16);
      mbar_init(&blur_full[s], kSplitRoles ? 8 : 16);
      mbar_init(&blur_empty[s], 1);
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kSplitRoles ? 8 : 16);
      mbar_init(&xd_empty[s], 1);
    }
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(Cf::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // (n_tiles is 1 or 2; the tile grid is a power of two: an integer division costs ~20 instructions and eight warps
  // decode every tile)
  auto decode = [&](int tile, int& img, int& ty, int& tx) {
    int m = n_tiles == 1 ? tile : tile >> 1;
    tx = m & (tiles_x - 1); m >>= p.sh_x;
    ty = m & (tiles_y - 1);
    img = m >> p.sh_y;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(w_bar, Cf::kWBytes + (kProj ? BN * C * 2 : 0));
      for (int tap = 0; tap < 9; ++tap) tma_load_3d(&map_w, smem_w + tap * Cf::kTapBytes, w_bar, 0, n_tile * BN, tap);
      if (kProj) tma_load_2d(&map_wp, wproj, w_bar, 0, n_tile * BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int img, ty, tx;
        decode(tile, img, ty, tx);
        for (int h = 0; h < Cf::kPasses; ++h) {
          mbar_wait_backoff(&raw_empty[stage], phase ^ 1);
          if (kProj) mbar_wait_backoff(&xd_empty[stage], phase ^ 1);
          mbar_expect_tx(&raw_full[stage], Cf::kRawBytes + (kProj ? kDcXdBytes : 0));
          // kProj (one pass per tile): the 8 x 16 tile of the downsampled block input rides on the raw tile's barrier
          if (kProj) tma_load_4d(&map_xd, xd + stage * kDcXdBytes, &raw_full[stage], 0, tx * kDcTW, ty * kDcTH, img);
          // (elements, group, row, image): raw columns 2*x0-2 .. +19, groups kG*h .. +kG-1, rows 2*y0-2 .. +35
          tma_load_4d(&map_a, raw + stage * Cf::kRawBytes, &raw_full[stage], (2 * tx * kDcTW - 2) * 8, h * Cf::kG,
                      2 * ty * kDcTH - 2, img);
          if (++stage == Cf::kRawStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      mbar_wait(w_bar, 0);
      tc_fence_after();
      const uint32_t sw = smem_u32(smem_w);
      int bstage = 0, it = 0, rstage = 0;
      uint32_t bphase = 0, rphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        mbar_wait_backoff(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        if (kProj) {
          // projection accumulator (columns 2 BN ..): W_proj . xd over the same 128 pixels, K = C = 32
          mbar_wait_backoff(&raw_full[rstage], rphase);          // (the tile stage shares the raw tile's barrier)
          tc_fence_after();
          const uint64_t dxa = make_smem_desc<32>(smem_u32(xd + rstage * kDcXdBytes));
          const uint64_t dxb = make_smem_desc<32>(smem_u32(wproj));
#pragma unroll
          for (int k = 0; k < 2; ++k)
            tc_mma_f16(tmem_base + 2 * BN + as * BN, dxa + (uint64_t)(k * 2), dxb + (uint64_t)(k * 2), Cf::kIdesc, (uint32_t)k);
          tc_commit(&xd_empty[rstage]);
          if (++rstage == Cf::kRawStages) { rstage = 0; rphase ^= 1; }
        }
        for (int h = 0; h < Cf::kPasses; ++h) {
          mbar_wait_backoff(&blur_full[bstage], bphase);
          tc_fence_after();
          const uint32_t sb = smem_u32(blur + bstage * Cf::kBlurBytes);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap - ky * 3;
            const uint32_t a_addr =
                sb + Cf::plane_off(ky & 1, kx & 1) + (ky >> 1) * Cf::plane_sbo(kx & 1) + (kx >> 1) * 16;
            const uint64_t db = make_smem_desc<(C > 64 ? 64 : C)>(sw + tap * Cf::kTapBytes);
#pragma unroll
            for (int k = 0; k < Cf::kStepsPerPass; ++k) {       // the channels of this pass in K = 16 steps
              const uint64_t da =
                  make_smem_desc_noswz(a_addr + k * 2 * Cf::plane_lbo(kx & 1), Cf::plane_lbo(kx & 1), Cf::plane_sbo(kx & 1));
              tc_mma_f16(d_tmem, da, db + (uint64_t)((h * Cf::kStepsPerPass + k) * 2), Cf::kIdesc,
                         (uint32_t)(h | tap | k));
            }
          }
          tc_commit(&blur_empty[bstage]);
          if (++bstage == Cf::kBlurStages) { bstage = 0; bphase ^= 1; }
        }
        tc_commit(&tmem_full[as]);
      }
    }
  } else if (kSplitRoles && warp < 10) {
    // ===================== epilogue (8 warps: lane quarter x column half) =====================
    // Waiting costs issue slots that the blur warps need (an mbarrier poll is ~12 instructions per iteration): ONE
    // warp polls the mbarrier, the other seven block on a hardware named barrier.
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int ry = row >> 3, rx = row & 7;
    constexpr int kCols = BN / 2, kChunks = kCols / 16;
    static_assert(kCols % 16 == 0, "a warp's column half must be whole 16-column chunks");
    const float s1 = kSqrt2 * p.post_scale;                  // (lrelu(a*sqrt2) + r) * ps == lrelu(a*sqrt2*ps) + r*ps
    if (threadIdx.x - 64 < BN) bias_s[threadIdx.x - 64] = __ldg(p.bias + n_tile * BN + threadIdx.x - 64) * s1;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int img, ty, tx;
      decode(tile, img, ty, tx);
      const int as = it & 1;
      const int y = ty * kDcTH + ry, x = tx * kDcTW + rx;
      const int n0 = n_tile * BN + half * kCols;
      const size_t pix = ((size_t)img * p.Ho + y) * p.Wo + x;
      // residual operand fetched before the accumulator wait (hides its DRAM latency); kProj: it is the second
      // accumulator of this tile instead
      uint4 resv[kChunks][2];
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int nc = n0 + c * 16;
        if (kProj) {
          resv[c][0] = resv[c][1] = make_uint4(0, 0, 0, 0);
        } else if (p.res_i8) {
          const __half* rp = p.residual + ((((size_t)img * p.Ho + y) * (p.Cout >> 3) + (nc >> 3)) * p.Wo + x) * 8;
          resv[c][0] = __ldg(reinterpret_cast<const uint4*>(rp));
          resv[c][1] = __ldg(reinterpret_cast<const uint4*>(rp + (size_t)p.Wo * 8));
        } else {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * p.Cout + nc);
          resv[c][0] = __ldg(rp);
          resv[c][1] = __ldg(rp + 1);
        }
      }
      if (warp == 2) mbar_wait_backoff(&tmem_full[as], (it >> 1) & 1);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + half * kCols;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        float v[16];
        tc_ld16(taddr + c * 16, v);
        const int nc = n0 + c * 16;
        const __half2* r0 = reinterpret_cast<const __half2*>(&resv[c][0]);
        const __half2* r1 = reinterpret_cast<const __half2*>(&resv[c][1]);
        const float4* bs = reinterpret_cast<const float4*>(bias_s + half * kCols + c * 16);
        float t[16];
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          const float4 b = bs[g4];
          const float a0 = fmaf(v[4 * g4], s1, b.x), a1 = fmaf(v[4 * g4 + 1], s1, b.y);
          const float a2 = fmaf(v[4 * g4 + 2], s1, b.z), a3 = fmaf(v[4 * g4 + 3], s1, b.w);
          t[4 * g4] = fmaxf(a0, 0.2f * a0); t[4 * g4 + 1] = fmaxf(a1, 0.2f * a1);
          t[4 * g4 + 2] = fmaxf(a2, 0.2f * a2); t[4 * g4 + 3] = fmaxf(a3, 0.2f * a3);
        }
        uint4 w0, w1;
        __half2* h0 = reinterpret_cast<__half2*>(&w0);
        __half2* h1 = reinterpret_cast<__half2*>(&w1);
        if (kProj) {
          float r[16];                           // projection of this pixel, fp32 (never rounded to fp16)
          tc_ld16(taddr + 2 * BN + c * 16, r);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            h0[j] = f2h2_sat(fmaf(r[2 * j], p.post_scale, t[2 * j]), fmaf(r[2 * j + 1], p.post_scale, t[2 * j + 1]));
            h1[j] = f2h2_sat(fmaf(r[8 + 2 * j], p.post_scale, t[8 + 2 * j]), fmaf(r[8 + 2 * j + 1], p.post_scale, t[8 + 2 * j + 1]));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 a = __half22float2(r0[j]), b = __half22float2(r1[j]);
            h0[j] = f2h2_sat(fmaf(a.x, p.post_scale, t[2 * j]), fmaf(a.y, p.post_scale, t[2 * j + 1]));
            h1[j] = f2h2_sat(fmaf(b.x, p.post_scale, t[8 + 2 * j]), fmaf(b.y, p.post_scale, t[8 + 2 * j + 1]));
          }
        }
        if (p.out_i8) {
          __half* op = p.out + ((((size_t)img * p.Ho + y) * (p.Cout >> 3) + (nc >> 3)) * p.Wo + x) * 8;
          *reinterpret_cast<uint4*>(op) = w0;
          *reinterpret_cast<uint4*>(op + (size_t)p.Wo * 8) = w1;
        } else {
          uint4* op = reinterpret_cast<uint4*>(p.out + pix * p.Cout + nc);
          op[0] = w0;
          op[1] = w1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
  } else if (kSplitRoles) {
    // ===================== blur (8 warps) =====================
    // unit = (column j of the 17 blurred columns, channel group g, row strip s): 17 x 4 x 3 = 204 of 256 threads.
    // Strip s produces blurred rows i in [11 s, 11 s + 11) from raw rows 11 s .. 11 s + 13.
    // (consecutive lanes take consecutive columns of ONE channel group: the 8 lanes of an LDS.128 phase read 128
    // contiguous bytes; with the group index in the low lane bits every phase had a 2-way bank conflict)
    const int bt = threadIdx.x - 320;
    const int j = bt % 17, g = (bt / 17) & 3, s = bt / 68;
    const bool active = s < 3;
    const int i0 = 11 * s;
    int stage = 0, bstage = 0;
    uint32_t phase = 0, bphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int h = 0; h < Cf::kPasses; ++h) {
        if (warp == 10) {                      // one polling warp; the other seven block on a named barrier
          mbar_wait_backoff(&raw_full[stage], phase);
          mbar_wait_backoff(&blur_empty[bstage], bphase ^ 1);
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (active) {
          const uint4* rt = reinterpret_cast<const uint4*>(raw + stage * Cf::kRawBytes);
          // (dedicated blur warps exist only for C == 32: untrimmed planes of equal size)
          uint8_t* bt_base = blur + bstage * Cf::kBlurBytes + ((j & 1) * Cf::plane_bytes(0, 0)) + g * kDcLbo + (j >> 1) * 16;
          uint4 hw[4];                                    // horizontal results of the last four raw rows
#pragma unroll
          for (int r = 0; r < 14; ++r) {
            const uint4* rp = rt + ((i0 + r) * Cf::kG + g) * kDcRawCols + j;
            const uint4 cur = fir4<false>(rp[0], rp[1], rp[2], rp[3]);
            hw[r & 3] = cur;
            if (r >= 3) {
              const int i = i0 + r - 3;                   // blurred row completed by raw row i + 3
              const uint4 u = fir4<true>(hw[(r - 3) & 3], hw[(r - 2) & 3], hw[(r - 1) & 3], cur);
              *reinterpret_cast<uint4*>(bt_base + (i & 1) * 2 * Cf::plane_bytes(0, 0) + (i >> 1) * Cf::plane_sbo(0)) = u;
            }
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&raw_empty[stage]);
          mbar_arrive(&blur_full[bstage]);
        }
        if (++stage == Cf::kRawStages) { stage = 0; phase ^= 1; }
        if (++bstage == Cf::kBlurStages) { bstage = 0; bphase ^= 1; }
      }
    }
  } else {
    // ===================== workers: 16 warps, each blurs AND runs its share of the epilogue =====================
    // With dedicated blur / epilogue warps the blur stage paces the kernel while the epilogue warps sit blocked most
    // of the time (issue slots 45 % busy).  Here every worker warp blurs its part of a pass of tile t and, while the
    // MMAs of that pass run, finishes one 16-column chunk of tile t-1: all 16 warps have work all the time.
    //   blur unit = (column j of 17, 4-channel half-group of 2 kG, row strip): 408 (kG = 4: 3 strips of 11 rows) or 476
    //     (kG = 2: 7 strips of 5 rows) of 512 threads, LDS.64 / STS.64; lanes 2k, 2k+1 take the two halves of one
    //     pixel's 16 bytes, consecutive lane pairs consecutive columns: the 16 lanes of an LDS.64 phase read 128
    //     contiguous bytes (conflict-free), and so do the stores.
    //   epilogue unit = (TMEM lane quarter = warp & 3, BN / 4 columns = kCW chunks of 16 starting at ((warp - 2) >> 2)).
    // Waiting: ONE warp polls each mbarrier (a poll loop costs issue slots), the others block on a named barrier.
    constexpr int kG = Cf::kG;
    constexpr int kStrips = 512 / (34 * kG);                 // 3 / 7
    constexpr int kStripRows = (2 * kDcTH + 1 + kStrips - 1) / kStrips;      // 11 / 5 blurred rows per strip
    constexpr int kCW = BN / 64;                             // 16-column chunks per worker warp
    static_assert(BN == 64 || BN == 128, "16 worker warps = 4 lane quarters x 4 column parts");
    static_assert(Cf::kPasses >= kCW, "one epilogue chunk of the previous tile per pass of the current one");
    const int wt = threadIdx.x - 64;                       // 0 .. 511
    const int hlow = wt & 1, j = (wt >> 1) % 17, g = ((wt >> 1) / 17) % kG, strip = (wt >> 1) / (17 * kG);
    const bool active = strip < kStrips;
    const int i0 = kStripRows * strip;
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int ry = row >> 3, rx = row & 7;
    const float s1 = kSqrt2 * p.post_scale;                  // (lrelu(a*sqrt2) + r) * ps == lrelu(a*sqrt2*ps) + r*ps
    if (wt < BN) bias_s[wt] = __ldg(p.bias + n_tile * BN + wt) * s1;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    const int nc0 = n_tile * BN + part * 16 * kCW;           // first output column of this warp

    auto res_ptr = [&](int img, int y, int x, int nc) -> const __half* {
      if (p.res_i8) return p.residual + ((((size_t)img * p.Ho + y) * (p.Cout >> 3) + (nc >> 3)) * p.Wo + x) * 8;
      return p.residual + (((size_t)img * p.Ho + y) * p.Wo + x) * p.Cout + nc;
    };
    // chunk c of the previous tile of this CTA (it = its iteration index): TMEM -> bias, lrelu, + residual -> store
    auto epilogue = [&](int img, int ty, int tx, int it, int c, const uint4& q0, const uint4& q1) {
      const int as = it & 1;
      const int y = ty * kDcTH + ry, x = tx * kDcTW + rx;
      const int nc = nc0 + c * 16;
      if (c == 0) {
        if (warp == 2) mbar_wait_backoff(&tmem_full[as], (it >> 1) & 1);
        asm volatile("bar.sync 1, 512;" ::: "memory");
        tc_fence_after();
      }
      float v[16];
      tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + part * 16 * kCW + c * 16, v);
      if (c == kCW - 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[as]);         // the accumulator is in registers: the MMAs may go on
      }
      const __half2* r0 = reinterpret_cast<const __half2*>(&q0);
      const __half2* r1 = reinterpret_cast<const __half2*>(&q1);
      const float4* bs = reinterpret_cast<const float4*>(bias_s + part * 16 * kCW + c * 16);
      const f32x2 s2 = pk2(s1, s1), ps2 = pk2(p.post_scale, p.post_scale);
      f32x2 t[8];
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {
        const float4 b = bs[g4];
        t[2 * g4] = lrelu2(fma2(pk2(v[4 * g4], v[4 * g4 + 1]), s2, pk2(b.x, b.y)));
        t[2 * g4 + 1] = lrelu2(fma2(pk2(v[4 * g4 + 2], v[4 * g4 + 3]), s2, pk2(b.z, b.w)));
      }
      uint4 w0, w1;
      __half2* h0 = reinterpret_cast<__half2*>(&w0);
      __half2* h1 = reinterpret_cast<__half2*>(&w1);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 a = __half22float2(r0[k]), b = __half22float2(r1[k]);
        float lo, hi;
        upk2(fma2(pk2(a.x, a.y), ps2, t[k]), lo, hi);
        h0[k] = f2h2_sat(lo, hi);
        upk2(fma2(pk2(b.x, b.y), ps2, t[4 + k]), lo, hi);
        h1[k] = f2h2_sat(lo, hi);
      }
      if (p.out_i8) {
        __half* op = p.out + ((((size_t)img * p.Ho + y) * (p.Cout >> 3) + (nc >> 3)) * p.Wo + x) * 8;
        *reinterpret_cast<uint4*>(op) = w0;
        *reinterpret_cast<uint4*>(op + (size_t)p.Wo * 8) = w1;
      } else {
        uint4* op = reinterpret_cast<uint4*>(p.out + (((size_t)img * p.Ho + y) * p.Wo + x) * p.Cout + nc);
        op[0] = w0;
        op[1] = w1;
      }
    };

    int stage = 0, bstage = 0, it = 0;
    uint32_t phase = 0, bphase = 0;
    int pimg = 0, pty = 0, ptx = 0;                          // the previous tile of this CTA (epilogue pending)
    uint4 pq[kCW][2];                                        // its residual operand, fetched during the tile before
#pragma unroll
    for (int c = 0; c < kCW; ++c) pq[c][0] = pq[c][1] = make_uint4(0, 0, 0, 0);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int img, ty, tx;
      decode(tile, img, ty, tx);
#pragma unroll
      for (int h = 0; h < Cf::kPasses; ++h) {
        if (warp == 2) {
          mbar_wait_backoff(&raw_full[stage], phase);
          mbar_wait_backoff(&blur_empty[bstage], bphase ^ 1);
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (active) {
          const uint2* rt = reinterpret_cast<const uint2*>(raw + stage * Cf::kRawBytes) + hlow;
          // this thread's column in the even-row / odd-row plane of its column parity
          const int px = j & 1;
          const int lbo = px ? Cf::plane_lbo(1) : Cf::plane_lbo(0), sbo = kG * lbo;
          uint8_t* bt_col = blur + bstage * Cf::kBlurBytes + g * lbo + (j >> 1) * 16 + hlow * 8;
          uint8_t* bt_even = bt_col + (px ? Cf::plane_off(0, 1) : Cf::plane_off(0, 0));
          uint8_t* bt_odd = bt_col + (px ? Cf::plane_off(1, 1) : Cf::plane_off(1, 0));
          uint2 hw[4];                                     // horizontal results of the last four raw rows
#pragma unroll
          for (int r = 0; r < kStripRows + 3; ++r) {
            const int i = i0 + r - 3;                      // blurred row completed by raw row i + 3
            if (r >= 3 && i > 2 * kDcTH) break;            // (the last strip is shorter: 33 blurred rows)
            const uint2* rr = rt + (((i0 + r) * kG + g) * kDcRawCols + j) * 2;
            const uint2 cur = fir4h<false>(rr[0], rr[2], rr[4], rr[6]);
            hw[r & 3] = cur;
            if (r >= 3) {
              const uint2 u = fir4h<true>(hw[(r - 3) & 3], hw[(r - 2) & 3], hw[(r - 1) & 3], cur);
              *reinterpret_cast<uint2*>(((i & 1) ? bt_odd : bt_even) + (i >> 1) * sbo) = u;
            }
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&raw_empty[stage]);
          mbar_arrive(&blur_full[bstage]);
        }
        if (++stage == Cf::kRawStages) { stage = 0; phase ^= 1; }
        if (++bstage == Cf::kBlurStages) { bstage = 0; bphase ^= 1; }
        // while the MMAs of this pass run: one chunk of the previous tile
        if (h < kCW && it > 0) epilogue(pimg, pty, ptx, it - 1, h, pq[h < kCW ? h : 0][0], pq[h < kCW ? h : 0][1]);
        if (h == kCW - 1) {
          // residual of THIS tile (consumed one tile later): in flight while the remaining passes and MMAs run
#pragma unroll
          for (int c = 0; c < kCW; ++c) {
            const __half* rp = res_ptr(img, ty * kDcTH + ry, tx * kDcTW + rx, nc0 + c * 16);
            pq[c][0] = __ldg(reinterpret_cast<const uint4*>(rp));
            pq[c][1] = __ldg(reinterpret_cast<const uint4*>(rp + (p.res_i8 ? (size_t)p.Wo * 8 : 8)));
          }
        }
      }
      pimg = img; pty = ty; ptx = tx;
    }
    if (it > 0) {
#pragma unroll
      for (int c = 0; c < kCW; ++c) epilogue(pimg, pty, ptx, it - 1, c, pq[c][0], pq[c][1]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cf::kTmemCols));
  }
}

template <int C, int BN, bool kProj = false>
cudaError_t launch_down(const CUtensorMap& map_a, const CUtensorMap& map_w, const CUtensorMap& map_xd,
                        const CUtensorMap& map_wp, const DownParams& p, int num_sms, cudaStream_t s) {
  using Cf = DCfg<C, BN, kProj>;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(downconv_tc_kernel<C, BN, kProj>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cf::kSmemBytes);
    if (err != cudaSuccess) return err;
    configured = true;
  }
  const int n_tiles = p.Cout / BN;
  const int total = p.N * (p.Ho / kDcTH) * (p.Wo / kDcTW) * n_tiles;
  int grid = total < num_sms ? total : num_sms;
  grid = (grid / n_tiles) * n_tiles;
  if (grid <= 0) return cudaErrorInvalidValue;
  downconv_tc_kernel<C, BN, kProj><<<grid, kDcThreads, Cf::kSmemBytes, s>>>(map_a, map_w, map_xd, map_wp, p);
  return cudaGetLastError();
}

}  // namespace

bool k_downconv_fused_supported(int C, int Cout, int Ho, int Wo) {
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  return (C == 32 || C == 64) && (Cout == 64 || Cout == 128) && Ho % kDcTH == 0 && Wo % kDcTW == 0 &&
         pow2(Ho / kDcTH) && pow2(Wo / kDcTW);
}

bool k_downconv_proj_supported(int C, int Cout) { return C == 32 && Cout == 64; }

// map_xd / map_wp (non-null only where k_downconv_proj_supported): the projection path as a second accumulator -- the
// FIR-downsampled block input [C, Wo, Ho, N] (box C x 8 x 16 x 1, 64B swizzle) and the 1x1 weights [C, Cout] (box
// C x 64); `residual` is ignored then.
cudaError_t k_downconv_fused(const CUtensorMap& map_a, const CUtensorMap& map_w, const CUtensorMap* map_xd,
                             const CUtensorMap* map_wp, int C, int N, int Ho, int Wo, int Cout, const float* bias,
                             const __half* residual, int res_i8, __half* out, int out_i8, float post_scale, int num_sms,
                             cudaStream_t s) {
  auto lg = [](int v) { int s = 0; while ((1 << s) < v) ++s; return ((1 << s) == v) ? s : -1; };
  const int sx = lg(Wo / kDcTW), sy = lg(Ho / kDcTH);
  if (sx < 0 || sy < 0 || Cout / 64 > 2) return cudaErrorInvalidValue;
  DownParams p{N, Ho, Wo, Cout, sx, sy, bias, residual, res_i8, out, out_i8, post_scale};
  if (map_xd != nullptr && map_wp != nullptr) {
    if (!k_downconv_proj_supported(C, Cout)) return cudaErrorInvalidValue;
    return launch_down<32, 64, true>(map_a, map_w, *map_xd, *map_wp, p, num_sms, s);
  }
  // (the two extra tensor maps are kernel parameters of every instance; unused ones just repeat map_a / map_w)
  if (C == 32) return launch_down<32, 64>(map_a, map_w, map_a, map_w, p, num_sms, s);
  if (C == 64 && Cout == 128) return launch_down<64, 128>(map_a, map_w, map_a, map_w, p, num_sms, s);
  if (C == 64) return launch_down<64, 64>(map_a, map_w, map_a, map_w, p, num_sms, s);
  return cudaErrorInvalidValue;
}

// TMA box geometry the kernel instance for (C, Cout) expects: 8-channel groups per raw box, output columns per tap box
void k_downconv_fused_geometry(int C, int Cout, int* box_groups, int* box_cols) {
  const bool wide = (C == 64 && Cout == 128);
  *box_groups = wide ? DCfg<64, 128>::kG : DCfg<32, 64>::kG;
  *box_cols = wide ? 128 : 64;
}

}  // namespace glass
