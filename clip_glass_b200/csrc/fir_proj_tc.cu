// Projection path of a discriminator block (stylegan2/modules.py:1587-1601: FilterLayer pad 1 -> 1x1 conv stride 2,
// no bias, no activation) as ONE kernel: the stride-2 FIR of fir_tile.cuh feeding a tcgen05 GEMM.
//
//   per block: 8 x 16 output pixels (= the M = 128 rows of one MMA tile), all C input channels in chunks of 32:
//     stage the 18 x 34 x 32-channel patch (or compute it from the image: fromRGB)      -> shared tile
//     FIR (fp32, as k_fir_down) -> fp16 A operand [128 pixels x 32 channels] in the un-swizzled K-major layout
//     proj weights of the chunk  -> B operand [Co x 32]
//     2 x tcgen05.mma (K = 16 each), accumulating over the chunks in TMEM [128 lanes x Co columns]
//   epilogue: TMEM -> fp16 -> dR [N][Ho][Wo][Co] (the residual operand of the block's conv1)
//
// Against k_fir_down + a 1x1 conv_tc launch this never writes or re-reads the filtered tensor (1.07 GB each way in the
// 1024^2 block at P = 64).  MEASURED: not faster -- 2.86 ms vs 1.66 + 0.89 ms on the 1024^2 block, 1.30 vs 0.70 + 0.47
// on the 512^2 block, 0.64 vs 0.36 + 0.29 on the 256^2 block (profiles/r01_fir_proj_fusion.txt): a block runs its
// stage / FIR / MMA / store phases serially and only three blocks fit an SM, so the issue slots idle (61 % busy
// against 83 % for k_from_rgb_fir).  Hence opt-in (GLASS_FLAG_PROJ_FUSION) and cross-checked by the GPU tests; making
// it win needs two tiles in flight per block.
// Operands are staged by the block's own threads (like attention_tc.cu): `fence.proxy.async` orders the generic-proxy
// writes before the MMA's reads; a single mbarrier, completed by tcgen05.commit, protects the operand buffers.
#include "fir_tile.cuh"
#include "kernels.cuh"
#include "tcgen05.cuh"

namespace glass {

namespace {

constexpr int kFpLbo = 144;                       // bytes between core matrices adjacent in K (128 + 16: spreads banks)
constexpr int kFpSbo = (kFdC / 8) * kFpLbo;       // 8-row group pitch of a K = 32 operand
constexpr int kFpABytes = (128 / 8) * kFpSbo;     // A: 128 pixels x 32 channels
template <int kCo>
struct FpCfg {
  static constexpr int kBBytes = (kCo / 8) * kFpSbo;                        // B: kCo rows x 32 channels
  static constexpr int kSmemBytes = kFdUnits * 16 + kFpABytes + kBBytes + 64;
  // instruction descriptor: D=f32 [4,6)=1, A=B=f16, both K-major, N>>3 [17,23), M>>4 [24,29)
  static constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kCo >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
};
constexpr int kMaxRgbChunks = 2;                  // fromRGB variant: up to 64 input channels
struct FrgbConstsN { FrgbConsts c[kMaxRgbChunks]; };

// kSrc: 0 = NHWC activations, 1 = I8 activations, 2 = the image (fromRGB computed on the fly; x is written to xout)
template <int kSrc, int kCo, int kRgbChunks>
__global__ void __launch_bounds__(256, (kSrc == 2 ? 3 : 2))
fir_proj_kernel(const __half* __restrict__ x, const float* __restrict__ images, const __grid_constant__ FrgbConstsN frgb,
                __half* __restrict__ xout, int xout_i8, const __half* __restrict__ wproj, __half* __restrict__ out, int H,
                int W, int C) {
  using Cf = FpCfg<kCo>;
  extern __shared__ __align__(128) uint8_t fp_smem[];
  uint4* tile = reinterpret_cast<uint4*>(fp_smem);
  uint8_t* sA = fp_smem + kFdUnits * 16;
  uint8_t* sB = sA + kFpABytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + Cf::kBBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int Ho = H >> 1, Wo = W >> 1;
  int b, ty, tx;
  fd_decode_tile(blockIdx.x, Ho, Wo, b, ty, tx);
  const int iy0 = 2 * ty * kFdTH - 1, ix0 = 2 * tx * kFdTW - 1;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kCo));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nchunks = C / kFdC;
  for (int ch = 0; ch < nchunks; ++ch) {
    const int c0 = ch * kFdC;
    if (kSrc == 2) {
      if (kRgbChunks == 1) fd_stage_tile_from_rgb(tile, images, frgb.c[0], xout, b, H, C, c0, xout_i8, iy0, ix0);
      else fd_stage_tile_from_rgb(tile, images, frgb.c[ch], xout, b, H, C, c0, xout_i8, iy0, ix0);
    } else {
      fd_stage_tile<kSrc == 1>(tile, x, b, H, W, C, c0, iy0, ix0);
    }
    // the MMAs of the previous chunk must have finished reading A and B before they are overwritten
    if (ch > 0) {
      mbar_wait(bar, (uint32_t)((ch - 1) & 1));
      tc_fence_after();
    }
    // B operand: rows n = output channels, K = the chunk's 32 input channels (wproj is [Co][C], K-major)
    for (int i = tid; i < kCo * 4; i += 256) {
      const int g = i & 3, n = i >> 2;
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(wproj + (size_t)n * C + c0 + g * 8));
      *reinterpret_cast<uint4*>(sB + (n >> 3) * kFpSbo + g * kFpLbo + (n & 7) * 16) = w;
    }
    __syncthreads();                               // tile complete
    {
      uint4 o[2];
      fir_down_compute(tile, o);
      const int ox = tid & 15, g = (tid >> 4) & 3, oyp = tid >> 6;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int m = (2 * oyp + k) * kFdTW + ox;  // accumulator row = pixel of the 8 x 16 tile
        *reinterpret_cast<uint4*>(sA + (m >> 3) * kFpSbo + g * kFpLbo + (m & 7) * 16) = o[k];
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();                               // A, B complete; tile free for the next chunk
    tc_fence_after();
    if (warp == 0 && elect_one()) {
#pragma unroll
      for (int k = 0; k < kFdC / 16; ++k) {
        const uint64_t da = make_smem_desc_noswz(smem_u32(sA) + k * 2 * kFpLbo, kFpLbo, kFpSbo);
        const uint64_t db = make_smem_desc_noswz(smem_u32(sB) + k * 2 * kFpLbo, kFpLbo, kFpSbo);
        tc_mma_f16(tmem_base, da, db, Cf::kIdesc, (ch | k) != 0);
      }
      tc_commit(bar);
    }
  }
  mbar_wait(bar, (uint32_t)((nchunks - 1) & 1));
  tc_fence_after();

  // epilogue: warp w reads TMEM lane quarter w & 3 (its 32 pixels), column half w >> 2
  {
    const int q = warp & 3, hh = warp >> 2;
    const int m = q * 32 + (tid & 31);
    const int zy = ty * kFdTH + m / kFdTW, zx = tx * kFdTW + (m % kFdTW);
    const bool valid = zy < Ho && zx < Wo;
    __half* orow = out + (((size_t)b * Ho + zy) * Wo + zx) * kCo + hh * (kCo / 2);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + hh * (kCo / 2);
    constexpr int kChunks = kCo / 2 / 16;
    uint32_t r[2][16];
    tc_ld16_issue(taddr, r[0]);
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      tc_ld_wait();
      if (c + 1 < kChunks) tc_ld16_issue(taddr + (c + 1) * 16, r[(c + 1) & 1]);
      if (valid) {
        uint4 w0, w1;
        __half2* a = reinterpret_cast<__half2*>(&w0);
        __half2* d = reinterpret_cast<__half2*>(&w1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          a[j] = f2h2_sat(__uint_as_float(r[c & 1][2 * j]), __uint_as_float(r[c & 1][2 * j + 1]));
          d[j] = f2h2_sat(__uint_as_float(r[c & 1][8 + 2 * j]), __uint_as_float(r[c & 1][8 + 2 * j + 1]));
        }
        *reinterpret_cast<uint4*>(orow + c * 16) = w0;
        *reinterpret_cast<uint4*>(orow + c * 16 + 8) = w1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kCo));
  }
}

template <int kSrc, int kCo, int kRgbChunks>
cudaError_t launch_fir_proj(const __half* x, const float* images, const FrgbConstsN& frgb, __half* xout, int xout_i8,
                            const __half* wproj, __half* out, int N, int H, int W, int C, cudaStream_t s) {
  using Cf = FpCfg<kCo>;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(fir_proj_kernel<kSrc, kCo, kRgbChunks>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::kSmemBytes);
    if (err != cudaSuccess) return err;
    configured = true;
  }
  const int Ho = H / 2, Wo = W / 2;
  const int tiles = ((Wo + kFdTW - 1) / kFdTW) * ((Ho + kFdTH - 1) / kFdTH) * N;
  fir_proj_kernel<kSrc, kCo, kRgbChunks><<<tiles, 256, Cf::kSmemBytes, s>>>(x, images, frgb, xout, xout_i8, wproj, out, H,
                                                                           W, C);
  return cudaGetLastError();
}

}  // namespace

bool k_fir_proj_supported(int C, int Co, bool from_rgb) {
  if (C % kFdC != 0 || (Co != 64 && Co != 128 && Co != 256)) return false;
  if (from_rgb) return Co == 64 && C <= kFdC * kMaxRgbChunks;
  return true;
}

cudaError_t k_fir_proj(const __half* x, int in_i8, const __half* wproj, __half* out, int N, int H, int W, int C, int Co,
                       cudaStream_t s) {
  if (!k_fir_proj_supported(C, Co, false) || (H & 1) || (W & 1)) return cudaErrorInvalidValue;
  FrgbConstsN none;
  none.c[0].w[0][0] = 0.f;
#define GLASS_FP(src, co) \
  if (in_i8 == src && Co == co) return launch_fir_proj<src, co, 1>(x, nullptr, none, nullptr, 0, wproj, out, N, H, W, C, s);
  GLASS_FP(0, 64) GLASS_FP(0, 128) GLASS_FP(0, 256) GLASS_FP(1, 64) GLASS_FP(1, 128) GLASS_FP(1, 256)
#undef GLASS_FP
  return cudaErrorInvalidValue;
}

cudaError_t k_from_rgb_fir_proj(const float* images, const float* folded_host, __half* xout, int out_i8,
                                const __half* wproj, __half* out, int P, int R, int C, int Co, cudaStream_t s) {
  if (!k_fir_proj_supported(C, Co, true) || (R & 1)) return cudaErrorInvalidValue;
  FrgbConstsN k;
  const int nch = C / kFdC;
  for (int ch = 0; ch < nch; ++ch)
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < kFdC; ++c) k.c[ch].w[r][c] = folded_host[r * C + ch * kFdC + c];
  if (nch == 1) return launch_fir_proj<2, 64, 1>(nullptr, images, k, xout, out_i8, wproj, out, P, R, R, C, s);
  return launch_fir_proj<2, 64, 2>(nullptr, images, k, xout, out_i8, wproj, out, P, R, R, C, s);
}

}  // namespace glass
