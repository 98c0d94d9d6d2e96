// Implicit-GEMM convolution / GEMM on the 5th-gen tensor cores (sm_100a).
//
//   warp 0      TMA producer: per K step one 4-D box of the NHWC activation
//               tensor (shifted by the filter tap; out-of-image rows/cols are
//               zero-filled by TMA) + one box of the [taps][Ntot][Cin] weights,
//               both landing 128B/64B-swizzled in shared memory.
//   warp 1      tcgen05.mma issuer (one elected lane), fp16 x fp16 -> fp32
//               accumulators in TMEM, double-buffered (2 x BN columns).
//   warps 2..5  epilogue: tcgen05.ld the 128 x BN tile, apply the fused
//               demod / noise / bias / activation / toRGB / residual / style
//               pre-scale (common.cuh: epilogue_row16) and store fp16 NHWC.
//
// Persistent CTAs (one per SM), static round-robin over (m_tile, n_tile).
// The SIMT kernel at the bottom computes the same accumulators the slow way and
// shares the epilogue: it exists for bring-up and as the in-library cross-check
// (glass_config.conv_impl = 1); it is never used by the product path.
#include <cstdio>
#include "common.cuh"

namespace glass {

namespace {

constexpr int kBlockM = 128;
constexpr int kNumThreads = 192;
constexpr long long kWaitLimitCycles = 4000000000ll;   // ~2 s at 1.9 GHz

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > kWaitLimitCycles) {
      printf("glass conv_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout [61,64)
template <int BK>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  constexpr uint64_t kSwizzleBytes = BK * 2;                      // 128 or 64
  constexpr uint64_t kLayout = (kSwizzleBytes == 128) ? 2 : 4;    // SWIZZLE_128B : SWIZZLE_64B
  constexpr uint64_t kSbo = (8 * kSwizzleBytes) >> 4;             // 8-row group pitch
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (kSbo << 32) | (1ull << 46) | (kLayout << 61);
}

template <int BN, int BK>
struct Cfg {
  static constexpr int kABytes = kBlockM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBudget = 196 * 1024;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 12 ? 12 : kStagesRaw;
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;   // power of two for BN in {32,64,128,256}
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  // instruction descriptor: D=f32 [4,6)=1, A=B=f16 (0), K-major both, N>>3 [17,23), M>>4 [24,29)
  static constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
};

template <int BN, int BK>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const ConvParams p) {
  using C = Cfg<BN, BK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

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
      mbar_init(&tmem_empty[s], 4);
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
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % n_tiles;
        int m = tile / n_tiles;
        const int tx = m % p.tiles_x; m /= p.tiles_x;
        const int ty = m % p.tiles_y;
        const int tn = m / p.tiles_y;
        const int x0 = tx * p.TW, y0 = ty * p.TH, i0 = tn * p.TN;
        for (int kit = 0; kit < kiters; ++kit) {
          const int tap = kit / kchunks;
          const int kc = kit - tap * kchunks;
          int dy = 0, dx = 0;
          if (p.taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
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
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kit = 0; kit < kiters; ++kit) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint64_t da = make_smem_desc<BK>(sa);
          const uint64_t db = make_smem_desc<BK>(sa + C::kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 fp16 = 32 bytes along K inside the swizzle atom
            tc_mma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), C::kIdesc, (kit | k) != 0);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[as]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;          // accumulator row == pixel within the tile
    const int thw = p.TH * p.TW;
    const int ri = row / thw;
    const int rr = row - ri * thw;
    const int ry = rr / p.TW;
    const int rx = rr - ry * p.TW;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int n_tile = tile % n_tiles;
      int m = tile / n_tiles;
      const int tx = m % p.tiles_x; m /= p.tiles_x;
      const int ty = m % p.tiles_y;
      const int tn = m / p.tiles_y;
      const int img = tn * p.TN + ri, y = ty * p.TH + ry, x = tx * p.TW + rx;
      const bool valid = img < p.Nimg && y < p.H && x < p.W;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      float rgb[3] = {0.f, 0.f, 0.f};
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 16; ++c) {
        float v[16];
        tc_ld16(taddr + c * 16, v);
        if (valid) epilogue_row16(p, img, y, x, n_tile * BN + c * 16, v, rgb);
      }
      if (valid && p.epi.rgb_w != nullptr) {
        const size_t pix = ((size_t)img * p.H + y) * p.W + x;
        p.epi.rgb_out[(size_t)n_tile * p.Nimg * p.H * p.W + pix] = make_float4(rgb[0], rgb[1], rgb[2], 0.f);
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

template <int BN, int BK>
cudaError_t launch_one(const ConvParams& p, const TmaMaps& maps, int num_sms, cudaStream_t s) {
  using C = Cfg<BN, BK>;
  static bool configured = false;
  if (!configured) {
    cudaError_t err =
        cudaFuncSetAttribute(conv_tc_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (err != cudaSuccess) return err;
    configured = true;
  }
  const int total = p.tiles_n * p.tiles_y * p.tiles_x * (p.Ntot / BN);
  const int grid = total < num_sms ? total : num_sms;
  conv_tc_kernel<BN, BK><<<grid, kNumThreads, C::kSmemBytes, s>>>(maps.a, maps.b, p);
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
      int dy = 0, dx = 0;
      if (p.taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
      const int yy = y + dy, xx = x + dx;
      if (yy < 0 || yy >= p.H || xx < 0 || xx >= p.W) continue;
      const __half* a = p.in + ((size_t)(img * p.H + yy) * p.W + xx) * p.Cin;
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

size_t conv_tc_smem_bytes(int BN, int BK) {
  const size_t stage = (size_t)kBlockM * BK * 2 + (size_t)BN * BK * 2;
  size_t stages = (196 * 1024) / stage;
  if (stages > 12) stages = 12;
  return stages * stage + 1024 + 256;
}

cudaError_t launch_conv_tc(const ConvParams& p, const TmaMaps& maps, int num_sms, cudaStream_t s) {
#define GLASS_CASE(bn, bk) \
  if (p.BN == bn && p.BK == bk) return launch_one<bn, bk>(p, maps, num_sms, s);
  GLASS_CASE(32, 32)
  GLASS_CASE(32, 64)
  GLASS_CASE(64, 64)
  GLASS_CASE(128, 64)
  GLASS_CASE(256, 64)
  GLASS_CASE(64, 32)
  GLASS_CASE(128, 32)
  GLASS_CASE(256, 32)
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
