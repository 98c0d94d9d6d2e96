// nn.MultiheadAttention core of the CLIP ViT blocks (clip/model.py:166-187, heads of 64, T <= 64 tokens, no mask)
// on the tensor cores: one CTA = two (image, head) problems stacked along M.
//
//   S = Q K^T      one tcgen05.mma chain  M=128 (2 x 64 query rows)  N=128 (2 x 64 key rows)  K=64
//                  -> TMEM columns [0,128); only the two diagonal 64x64 blocks are used
//   P = softmax    thread r owns accumulator row r (TMEM lane r): tcgen05.ld of its 64 columns, max / exp / sum in
//                  registers (no shuffles), P = fp16(exp / sum) written to shared memory as the next A operand,
//                  zeros in the other problem's key columns and in the padded keys
//   O = P V        M=128, N=64 (head dim), K=128 (2 x 64 keys): rows of problem h only meet V of problem h because
//                  the off-diagonal blocks of P are exact zeros -> TMEM columns [0,64) (S is dead by then)
//
// Operands are staged by the CTA's own threads (16-byte loads from the qkv GEMM output) straight into the un-swizzled
// K-major core-matrix layout the MMA reads (8 rows x 16 bytes per core matrix), so V is transposed on the way in and
// no TMA descriptor is needed; `fence.proxy.async` orders those generic-proxy writes before the MMA's async-proxy
// reads.  Four CTAs fit one SM (55 KB shared memory, 128 TMEM columns each) and overlap each other's phases.
#include "kernels.cuh"
#include "tcgen05.cuh"

namespace glass {

namespace {

constexpr int kAtHd = 64;                 // head dim
constexpr int kAtTok = 64;                // padded tokens per problem
constexpr int kAtLbo = 144;               // bytes between core matrices adjacent in K (128 + 16: spreads banks)
constexpr int kAtSbo64 = (64 / 8) * kAtLbo;     // 8-row group pitch of a K=64 operand (Q, K)
constexpr int kAtSbo128 = (128 / 8) * kAtLbo;   // ... of a K=128 operand (P, V^T)
constexpr int kAtQBytes = (128 / 8) * kAtSbo64;   // Q: 128 rows x K=64
constexpr int kAtKBytes = (128 / 8) * kAtSbo64;   // K: 128 rows x K=64
constexpr int kAtPBytes = (128 / 8) * kAtSbo128;  // P: 128 rows x K=128 (aliases Q and K, which are dead after S)
constexpr int kAtVBytes = (64 / 8) * kAtSbo128;   // V^T: 64 rows (dims) x K=128 (keys of both problems)
static_assert(kAtPBytes == kAtQBytes + kAtKBytes, "P reuses exactly the Q+K staging area");
constexpr int kAtSmemBytes = kAtPBytes + kAtVBytes + 64;
constexpr int kAtTmemCols = 128;
// instruction descriptor: D=f32 [4,6)=1, A=B=f16, both K-major, N>>3 [17,23), M>>4 [24,29)
constexpr uint32_t kAtIdescS = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t kAtIdescO = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// byte offset of element (row, k) of a K-major un-swizzled operand whose 8-row groups are `sbo` bytes apart
__device__ __forceinline__ uint32_t at_off(int row, int k, int sbo) {
  return (uint32_t)((row >> 3) * sbo + (k >> 3) * kAtLbo + (row & 7) * 16 + (k & 7) * 2);
}

__global__ void __launch_bounds__(128, 4)
attention_tc_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int T, int W, int n_problems) {
  extern __shared__ __align__(128) uint8_t at_smem[];
  uint8_t* sQ = at_smem;
  uint8_t* sK = at_smem + kAtQBytes;
  uint8_t* sP = at_smem;                          // alias: written only after the S MMAs have completed
  uint8_t* sV = at_smem + kAtPBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(at_smem + kAtPBytes + kAtVBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int heads = W / kAtHd;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kAtTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }

  // ---- stage Q, K (row = token, K = dim) and V^T (row = dim, K = key) for the CTA's two problems ----
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 2 * kAtTok * 8; i += 128) {       // (problem h, token t, 8-dim group g): 16-byte items
    const int g = i & 7, t = (i >> 3) & (kAtTok - 1), h = i >> 9;
    const int prob = blockIdx.x * 2 + h;
    uint4 q = zero4, k = zero4;
    if (prob < n_problems && t < T) {
      const int b = prob / heads, hd = prob - b * heads;
      const __half* r = qkv + ((size_t)b * T + t) * 3 * W + hd * kAtHd + g * 8;
      q = __ldg(reinterpret_cast<const uint4*>(r));
      k = __ldg(reinterpret_cast<const uint4*>(r + W));
    }
    const uint32_t o = at_off(h * kAtTok + t, g * 8, kAtSbo64);
    *reinterpret_cast<uint4*>(sQ + o) = q;
    *reinterpret_cast<uint4*>(sK + o) = k;
  }
  for (int i = tid; i < 2 * (kAtTok / 2) * 8; i += 128) {  // (problem h, token pair tp, 8-dim group g)
    const int tp = i & 31, g = (i >> 5) & 7, h = i >> 8;   // consecutive threads: consecutive keys (4-byte stores)
    const int prob = blockIdx.x * 2 + h;
    uint4 v0 = zero4, v1 = zero4;
    if (prob < n_problems) {
      const int b = prob / heads, hd = prob - b * heads;
      const __half* r = qkv + ((size_t)b * T + 2 * tp) * 3 * W + 2 * W + hd * kAtHd + g * 8;
      if (2 * tp < T) v0 = __ldg(reinterpret_cast<const uint4*>(r));
      if (2 * tp + 1 < T) v1 = __ldg(reinterpret_cast<const uint4*>(r + 3 * W));
    }
    const __half* a = reinterpret_cast<const __half*>(&v0);
    const __half* c = reinterpret_cast<const __half*>(&v1);
#pragma unroll
    for (int j = 0; j < 8; ++j)       // dim g*8+j, keys (h*64 + 2tp, +1)
      *reinterpret_cast<__half2*>(sV + at_off(g * 8 + j, h * kAtTok + 2 * tp, kAtSbo128)) = __halves2half2(a[j], c[j]);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && elect_one()) {
#pragma unroll
    for (int k = 0; k < kAtHd / 16; ++k) {
      const uint64_t da = make_smem_desc_noswz(smem_u32(sQ) + k * 2 * kAtLbo, kAtLbo, kAtSbo64);
      const uint64_t db = make_smem_desc_noswz(smem_u32(sK) + k * 2 * kAtLbo, kAtLbo, kAtSbo64);
      tc_mma_f16(tmem_base, da, db, kAtIdescS, k != 0);
    }
    tc_commit(&bars[0]);
  }
  mbar_wait(&bars[0], 0);
  tc_fence_after();

  // ---- softmax of row `tid` over its own problem's T keys ----
  const int h = tid >> 6;                                   // problem of this row (warp-uniform)
  const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
  float s[kAtTok];
  {
    uint32_t r[4][16];
#pragma unroll
    for (int c = 0; c < 4; ++c) tc_ld16_issue(lane_addr + h * kAtTok + c * 16, r[c]);
    tc_ld_wait();
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int j = 0; j < 16; ++j) s[c * 16 + j] = __uint_as_float(r[c][j]) * 0.125f;   // head_dim^-0.5 (model.py:181)
  }
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < kAtTok; ++j) m = (j < T) ? fmaxf(m, s[j]) : m;
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kAtTok; ++j) {
    s[j] = (j < T) ? __expf(s[j] - m) : 0.f;
    sum += s[j];
  }
  const float inv = 1.f / sum;
  // P aliases the Q/K staging area: safe, the S MMAs that read it have completed (bars[0]).  The S columns of TMEM
  // are overwritten by O only after the barrier below, when every warp has finished its tcgen05.ld.
#pragma unroll
  for (int g = 0; g < 16; ++g) {                           // 16 groups of 8 keys: [0,8) problem 0, [8,16) problem 1
    uint4 pk = zero4;
    if ((g >> 3) == h) {
      __half2* h2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        h2[j] = __floats2half2_rn(s[(g & 7) * 8 + 2 * j] * inv, s[(g & 7) * 8 + 2 * j + 1] * inv);
    }
    *reinterpret_cast<uint4*>(sP + at_off(tid, g * 8, kAtSbo128)) = pk;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0 && elect_one()) {
#pragma unroll
    for (int k = 0; k < 128 / 16; ++k) {
      const uint64_t da = make_smem_desc_noswz(smem_u32(sP) + k * 2 * kAtLbo, kAtLbo, kAtSbo128);
      const uint64_t db = make_smem_desc_noswz(smem_u32(sV) + k * 2 * kAtLbo, kAtLbo, kAtSbo128);
      tc_mma_f16(tmem_base, da, db, kAtIdescO, k != 0);
    }
    tc_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();

  {
    uint32_t r[4][16];
#pragma unroll
    for (int c = 0; c < 4; ++c) tc_ld16_issue(lane_addr + c * 16, r[c]);
    tc_ld_wait();
    const int t = tid & (kAtTok - 1);
    const int prob = blockIdx.x * 2 + h;
    if (prob < n_problems && t < T) {
      const int b = prob / heads, hd = prob - b * heads;
      __half* o = out + ((size_t)b * T + t) * W + hd * kAtHd;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 w0, w1;
        __half2* a = reinterpret_cast<__half2*>(&w0);
        __half2* d = reinterpret_cast<__half2*>(&w1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          a[j] = __floats2half2_rn(__uint_as_float(r[c][2 * j]), __uint_as_float(r[c][2 * j + 1]));
          d[j] = __floats2half2_rn(__uint_as_float(r[c][8 + 2 * j]), __uint_as_float(r[c][8 + 2 * j + 1]));
        }
        *reinterpret_cast<uint4*>(o + c * 16) = w0;
        *reinterpret_cast<uint4*>(o + c * 16 + 8) = w1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kAtTmemCols));
  }
}

}  // namespace

cudaError_t k_attention_tc(const __half* qkv, __half* out, int P, int T, int W, cudaStream_t s) {
  if (T > kAtTok || W % kAtHd != 0) return cudaErrorInvalidValue;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemBytes);
    if (err != cudaSuccess) return err;
    configured = true;
  }
  const int n_problems = P * (W / kAtHd);
  attention_tc_kernel<<<(n_problems + 1) / 2, 128, kAtSmemBytes, s>>>(qkv, out, T, W, n_problems);
  return cudaGetLastError();
}

}  // namespace glass
