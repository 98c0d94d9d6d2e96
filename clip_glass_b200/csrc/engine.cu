// C ABI + host-side engine: owns packed weights, workspace, TMA descriptors and
// the per-generation launch sequence of the fitness path
//   latents -> mapping -> styles/demod -> 2*nb-1 modulated convs (+fused toRGB)
//   -> skip-sum/upsample -> image -> resize+im2col -> ViT -> cosine
//   [-> fromRGB -> resnet down blocks -> mbstd -> dense -> hinge]
// See include/clipglass_b200.h for the contract and DESIGN.md for the layout.
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/clipglass_b200.h"
#include "common.cuh"
#include "kernels.cuh"

using namespace glass;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_OK(expr)                                                                            \
  do {                                                                                           \
    cudaError_t err__ = (expr);                                                                  \
    if (err__ != cudaSuccess)                                                                    \
      return fail(GLASS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
  } while (0)

struct DevTensor {
  void* ptr = nullptr;
  size_t bytes = 0;
};

struct GLayer {
  int block, l, cin, cout, up, res;   // res = output resolution
};

struct ConvLaunch {
  ConvParams p;
  TmaMaps maps;
  double flops = 0;   // algorithmic (as written by the reference) — for reporting
  // fused-FIR exact down-conv (downconv_tc.cu) instead of a conv_tc launch: maps.a = conv0 output (I8), maps.b = the
  // nine 3x3 taps; p.epi carries bias / residual / out
  bool fused_down = false;
  int fd_C = 0, fd_N = 0, fd_Ho = 0, fd_Wo = 0, fd_Cout = 0;
  // ... with the block's projection as a second accumulator: maps2.a = FIR-downsampled block input, maps2.b = 1x1 weights
  bool fd_proj = false;
  TmaMaps maps2;
};

struct Arena {
  uint8_t* base = nullptr;
  size_t cap = 0, off = 0;
  void* take(size_t bytes) {
    off = (off + 1023) & ~size_t(1023);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

}  // namespace

struct glass_engine {
  glass_config cfg{};
  int num_sms = 148;
  std::map<std::string, DevTensor> tensors;
  bool finalized = false;
  PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;

  // derived architecture
  std::vector<GLayer> glayers;
  std::vector<int> conv_off, rgb_off;
  int S = 0;                 // total style width
  std::vector<int> gch;      // channels, 4x4 first
  int R = 0;                 // output resolution
  size_t noise_per_group = 0;
  std::vector<size_t> noise_layer_off;

  // workspace (sized for cfg.max_population)
  Arena arena;
  float *z32 = nullptr, *wA = nullptr, *wB = nullptr, *styles = nullptr, *noise = nullptr;
  float *styles_n = nullptr, *mscale = nullptr;   // power-of-two normalised conv styles + their scales (k_style_norm)
  std::vector<float*> dmod;
  std::vector<float*> rgbw;
  __half *actA = nullptr, *actB = nullptr, *actC = nullptr;   // actC: intermediate of the exact polyphase forms
  std::vector<int> g_exact;   // per G layer: 1 = exact polyphase up-conv
  std::vector<int> d_exact;   // per D block: 1 = exact polyphase down-conv
  std::vector<int> g_in_i8;   // per G layer: its INPUT activation is stored [N][H][C/8][W][8]
  std::vector<int> d_in_i8;   // per D block: its input activation likewise
  std::vector<int> d_c1_i8;   // per D block: the space-to-depth tensor between conv0 and the folded conv1 likewise
  std::vector<int> d_proj_fused;   // per D block: projection FIR + 1x1 GEMM in one kernel (fir_proj_tc.cu)
  std::vector<int> d_res_i8;       // per D block: the projection output (conv1's residual operand) is stored I8
  std::vector<int> d_c1_proj;      // per D block: the projection is a second accumulator of the fused down-conv
  std::vector<int> d_fused;        // per D block: conv1 runs as the fused-FIR exact down-conv (downconv_tc.cu)
  std::vector<int> g_pair;    // per G layer: 1 = 32-channel conv on horizontally paired pixels
  std::vector<int> d_pair;    // per D block: conv0 likewise
  float4 *slabs = nullptr, *yA = nullptr, *yB = nullptr;
  float* images = nullptr;
  __half *patches = nullptr, *patch_emb = nullptr, *tokens = nullptr, *hbuf = nullptr, *qkv = nullptr, *att = nullptr,
         *fc = nullptr;
  float *features = nullptr, *sim = nullptr, *neg_sim = nullptr, *text = nullptr;
  __half *dXd = nullptr, *dR = nullptr, *dOut = nullptr, *dFin = nullptr, *dFinOut = nullptr, *dDense = nullptr;
  float *dlogits = nullptr, *hinge = nullptr;
  double* x_host_stage = nullptr;   // device staging for f64 latents
  int* gather_rows = nullptr;       // device staging for glass_last_images_gather
  int last_eval_pop = 0;            // candidates whose images the last fused evaluation left in `images`
  bool have_text = false;

  // plan (rebuilt when pop changes)
  int plan_pop = -1;
  std::vector<ConvLaunch> g_convs, c_convs, d_convs;

  std::vector<float> frgb_folded;   // host copy of the folded fromRGB constants [4][C] (k_from_rgb_fir)
  int max_groups = 0;        // noise buffer capacity in minibatch groups
  int64_t launches = 0;
  // CUDA graph of one fitness evaluation (everything after the noise fill; fixed engine-owned pointers), replayed
  // on the caller's stream.  Built on the second evaluation of a plan (the first runs eagerly and configures the
  // kernels' attributes); dropped whenever the plan is rebuilt.
  cudaStream_t cap_stream = nullptr, side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  int graph_evals = 0;       // evaluations since the plan was built
  int64_t graph_launches = 0;   // kernel launches inside the graph
  // timing of tensor-core launches
  bool timing = false;
  std::vector<cudaEvent_t> ev;
  std::vector<const void*> ev_conv;   // which ConvLaunch each event pair timed
  size_t ev_used = 0;
  float last_conv_ms = 0.f;
  int last_conv_launches = 0;
  // debug: fp16 range scan of every G / D activation tensor (glass_set_range_check)
  bool range_check = false;
  unsigned long long* range_ctr = nullptr;
  // debug capture
  bool capture = false;
  std::map<std::string, std::vector<float>> captured;
};

namespace {

DevTensor* find_tensor(glass_engine* e, const std::string& name) {
  auto it = e->tensors.find(name);
  return it == e->tensors.end() ? nullptr : &it->second;
}
template <class T>
T* tptr(glass_engine* e, const std::string& name) {
  DevTensor* t = find_tensor(e, name);
  return t ? reinterpret_cast<T*>(t->ptr) : nullptr;
}

int check_tensor(glass_engine* e, const std::string& name, size_t bytes) {
  DevTensor* t = find_tensor(e, name);
  if (!t) return fail(GLASS_ERR_STATE, "weight tensor '%s' was not set", name.c_str());
  if (t->bytes != bytes)
    return fail(GLASS_ERR_ARG, "weight tensor '%s' has %zu bytes, expected %zu", name.c_str(), t->bytes, bytes);
  return GLASS_OK;
}

int encode_map(glass_engine* e, CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims,
               const uint64_t* strides_bytes, const uint32_t* box, int inner_bytes) {
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  // inner_bytes 128 / 64 select the matching swizzle; 0 = un-swizzled box (conv_tc MODE 4)
  CUtensorMapSwizzle sw = inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                        : inner_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                                             : CU_TENSOR_MAP_SWIZZLE_NONE;
  if (sw == CU_TENSOR_MAP_SWIZZLE_NONE && inner_bytes != 0)
    return fail(GLASS_ERR_ARG, "unsupported TMA inner box of %d bytes", inner_bytes);
  CUresult r = e->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(ptr), gdim, gstr, bdim, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(GLASS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u",
                (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  }
  return GLASS_OK;
}

int pick_bn(int ntot) {
  for (int bn : {256, 128, 64, 32})
    if (ntot % bn == 0) return bn;
  return 0;
}

// Build one implicit-GEMM launch.  gemm=true: plain [M x K] @ [Ntot x K]^T with M = W.
// Tap tables of the exact polyphase forms (packing.py: UP_EXACT_TAPS / DOWN_EXACT_TAPS)
const signed char kUpExactTaps[4][2] = {{-1, -1}, {-1, 0}, {0, -1}, {0, 0}};
const signed char kDownExactTaps[4][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}};

int make_conv(glass_engine* e, ConvLaunch* out, const __half* in, int Nimg, int H, int W, int Cin, const __half* wgt,
              int taps, int Ntot, const EpiParams& epi, bool gemm, const signed char (*table)[2] = nullptr,
              int in_H = 0, int in_W = 0, bool in_i8 = false, int skip_mode = 0, int skip_ch = 0) {
  ConvParams& p = out->p;
  memset(&p, 0, sizeof(p));
  p.Nimg = Nimg; p.H = H; p.W = W; p.Cin = Cin; p.taps = taps; p.Ntot = Ntot;
  p.in_H = in_H > 0 ? in_H : H;
  p.in_W = in_W > 0 ? in_W : W;
  for (int t = 0; t < taps; ++t) {
    if (table != nullptr) { p.tap_dy[t] = table[t][0]; p.tap_dx[t] = table[t][1]; }
    else if (taps == 9) { p.tap_dy[t] = (signed char)(t / 3 - 1); p.tap_dx[t] = (signed char)(t % 3 - 1); }
    else { p.tap_dy[t] = 0; p.tap_dx[t] = 0; }
  }
  p.in = in; p.wgt = wgt; p.epi = epi;
  if (gemm) {
    p.TW = 128; p.TH = 1; p.TN = 1;
  } else {
    p.TW = std::min(W, 16);
    p.TH = std::min(H, 128 / p.TW);
    p.TN = 128 / (p.TW * p.TH);
  }
  p.tiles_x = (W + p.TW - 1) / p.TW;
  p.tiles_y = (H + p.TH - 1) / p.TH;
  p.tiles_n = (Nimg + p.TN - 1) / p.TN;
  p.BN = pick_bn(Ntot);
  if (gemm) {
    // plain GEMMs: 128-column tiles (ViT: 25 row tiles x 6 / 18 / 24 column tiles fill the 148 SMs; the specialised
    // GEMM epilogues of conv_tc.cu are instantiated for this width)
    if (p.BN > 128 && Ntot % 128 == 0) p.BN = 128;
    static const char* cap = debug_env("GLASS_DEBUG_GEMM_BN");      // A/B knob for the plain GEMMs' tile width
    if (cap != nullptr && atoi(cap) >= 32 && Ntot % atoi(cap) == 0) p.BN = atoi(cap);
  }
  // Layers with fewer tiles than SMs (the 4x4 .. 16x16 blocks, the dense layer): narrower column tiles.  A K-streaming
  // CTA is bound by TMA rows per K step (128 activation rows + BN weight rows, ~0.41 rows per cycle): at BN = 256
  // sixteen CTAs each stream 2.4x the rows that 128 CTAs stream at BN = 32.  Cost model: waves x rows per K step.
  // Not for the I8 / pixel-pair layers (tile width tied to their layout), the ViT GEMMs (specialised epilogues at 128
  // columns) and the toRGB layers: their partial toRGB sums are formed per column tile and added up in tile order, so a
  // population-dependent tile width would make a candidate's scores depend on what else is in the launch
  // (tests/test_gpu_parity.py::test_full_size_properties).  GEMM accumulators do not depend on the tile width.
  {
    static const bool keep_wide = debug_env("GLASS_DEBUG_WIDE_SMALL") != nullptr;      // (debug builds: A/B)
    const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
    const bool vit_gemm = gemm && p.BN == 128 && m_tiles >= 8;
    if (!keep_wide && !in_i8 && !vit_gemm && epi.x_phases != 2 && epi.rgb_w == nullptr &&
        m_tiles * (Ntot / p.BN) < e->num_sms) {
      int best = p.BN;
      long best_cost = -1;
      for (int bn = p.BN; bn >= 32 && Ntot % bn == 0; bn /= 2) {
        const long waves = ((long)m_tiles * (Ntot / bn) + e->num_sms - 1) / e->num_sms;
        const long cost = waves * (128 + bn);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
      }
      p.BN = best;
    }
  }
  p.BK = (Cin % 64 == 0) ? 64 : 32;
  if (p.BN == 0 || Cin % 32 != 0 || Ntot % 16 != 0)
    return fail(GLASS_ERR_ARG, "unsupported conv shape Cin=%d Ntot=%d", Cin, Ntot);
  // structural zeros of the exact polyphase forms (common.cuh: ConvParams::skip_mode); only the tensor-core MODE 0
  // path skips them, and only when an n-tile (up) / a K chunk (down) lies inside one phase
  p.skip_mode = 0;
  p.skip_ch = skip_ch;
  p.rot_div = 1;
  if (skip_mode == 1 && e->cfg.conv_impl == 0 && !(e->cfg.flags & GLASS_FLAG_NO_ZERO_SKIP)) {
    while (p.BN > skip_ch) p.BN /= 2;
    if (p.BN >= 32 && skip_ch % p.BN == 0) p.skip_mode = 1;
  } else if (skip_mode == 2 && e->cfg.conv_impl == 0 && !(e->cfg.flags & GLASS_FLAG_NO_ZERO_SKIP) &&
             skip_ch % p.BK == 0) {
    p.skip_mode = 2;
  }
  // Small-channel layers (the whole K of a tap is one chunk) on full 16x8 tiles: resident taps + halo copies.
  p.mode = 0;
  if (in_i8) {
    // channel-group-interleaved input: 8-wide x 16-tall tiles, one un-swizzled haloed box per tile (MODE 4)
    // Cin <= 64: tile pairs (two 8x16 tiles side by side share one box; conv_tc.cu Cfg::kPairM)
    // 128 channels: MODE 6 (tile pairs, streamed taps) or, behind GLASS_DEBUG_C1_MODE4, MODE 4 with nine resident
    // taps of a 32-column n-tile and single 8-wide tiles (A/B knob)
    static const bool c128_mode4 = debug_env("GLASS_DEBUG_C1_MODE4") != nullptr;
    const int tw = (Cin == 128 && c128_mode4) ? 8 : 16;
    if (gemm || table != nullptr || taps != 9 || (Cin != 32 && Cin != 64 && Cin != 128) || H < 16 || W < tw || H % 16 ||
        W % tw)
      return fail(GLASS_ERR_ARG, "I8 input layout needs a 3x3 conv with 32/64/128 channels on a >=16x16 grid");
    p.mode = (Cin == 128 && !c128_mode4) ? 6 : 4;   // 128 channels: the nine taps do not fit beside the box -> streamed
    p.BK = Cin;                        // whole K of a tap in one stage
    p.TW = tw; p.TH = 16; p.TN = 1;
    p.tiles_x = W / tw; p.tiles_y = H / 16; p.tiles_n = Nimg;
    while (p.BN > (Cin == 32 ? 128 : (p.mode == 4 && Cin == 128 ? 32 : 64))) p.BN /= 2;   // resident taps / pair box + >= 2 stages must fit
    if (p.mode == 6 && p.BN != 64) return fail(GLASS_ERR_ARG, "I8 input with 128 channels needs Ntot %% 64 == 0");
  } else if (!gemm && table == nullptr && (Cin == 32 || Cin == 64) && p.TW == 16 && p.TH == 8 && p.TN == 1 &&
             (taps == 9 || taps == 1)) {
    if (taps == 9) {
      p.mode = 1;
      const int bn_cap = (Cin == 64) ? 64 : 128;       // 9 resident taps must leave room for >= 2 stages
      while (p.BN > bn_cap) p.BN /= 2;
    } else {
      p.mode = 2;                                      // 1x1: one resident tap, BN up to 256 (Cin 64) / 128 (Cin 32)
      if (Cin == 32 && p.BN > 128) p.BN = 128;
    }
  }
  {
    auto lg = [](int v) { int s = 0; while ((1 << s) < v) ++s; return ((1 << s) == v) ? s : -1; };
    const int ln = lg(Ntot / p.BN), lx = lg(p.tiles_x), ly = lg(p.tiles_y);
    p.pow2 = (ln >= 0 && lx >= 0 && ly >= 0) ? 1 : 0;
    p.sh_n = ln; p.sh_x = lx; p.sh_y = ly;
    p.all_valid = (H % p.TH == 0 && W % p.TW == 0 && Nimg % p.TN == 0) ? 1 : 0;
    const char* dbg = debug_env("GLASS_DEBUG_SKIP");     // timing experiments only; never set in tests or bench
    p.debug_skip = dbg ? atoi(dbg) : 0;
    p.epi.cout_shift = lg(p.epi.Cout);
    p.epi.noise_div_shift = lg(p.epi.noise_group_div);
  }
  if (e->cfg.conv_impl != 0) return GLASS_OK;   // SIMT bring-up path needs no descriptors
  // activations: [C, W, H, N]; outermost extent rounded up to the box (buffers carry the slack)
  const uint64_t wdecl = gemm ? (uint64_t)p.tiles_x * 128 : (uint64_t)p.in_W;
  const uint64_t ndecl = (uint64_t)p.tiles_n * p.TN;
  uint64_t dims[4] = {(uint64_t)Cin, wdecl, (uint64_t)p.in_H, ndecl};
  uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)Cin * 2 * wdecl, (uint64_t)Cin * 2 * wdecl * p.in_H};
  if (p.mode == 4 || p.mode == 6) {
    // [N][H][G][W][8] seen as dims (W*8, G, H, N); box ((TW+2)*8, G, TH+2, 1), no swizzle
    const uint64_t G = Cin / 8;
    uint64_t d4[4] = {(uint64_t)W * 8, G, (uint64_t)H, (uint64_t)Nimg};
    uint64_t s4[3] = {(uint64_t)W * 16, G * W * 16, (uint64_t)H * G * W * 16};
    uint32_t b4[4] = {(uint32_t)(p.TW + 2) * 8, (uint32_t)G, (uint32_t)p.TH + 2, 1};
    int rc4 = encode_map(e, &out->maps.a, in, 4, d4, s4, b4, 0);
    if (rc4 != GLASS_OK) return rc4;
    uint64_t wd4[3] = {(uint64_t)Cin, (uint64_t)Ntot, (uint64_t)taps};
    uint64_t ws4[2] = {(uint64_t)Cin * 2, (uint64_t)Cin * 2 * Ntot};
    const int bkc = p.BK > 64 ? 64 : p.BK;            // resident weights are loaded in 64-channel swizzled chunks
    uint32_t wb4[3] = {(uint32_t)bkc, (uint32_t)p.BN, 1};
    return encode_map(e, &out->maps.b, wgt, 3, wd4, ws4, wb4, bkc * 2);
  }
  // mode 1 with 3x3 taps loads the tile plus one halo row above and below per horizontal shift
  const uint32_t box_h = (p.mode == 1 && taps == 9) ? (uint32_t)p.TH + 2 : (uint32_t)p.TH;
  uint32_t box[4] = {(uint32_t)p.BK, (uint32_t)p.TW, box_h, (uint32_t)p.TN};
  int rc = encode_map(e, &out->maps.a, in, 4, dims, strides, box, p.BK * 2);
  if (rc != GLASS_OK) return rc;
  uint64_t wd[3] = {(uint64_t)Cin, (uint64_t)Ntot, (uint64_t)taps};
  uint64_t ws[2] = {(uint64_t)Cin * 2, (uint64_t)Cin * 2 * Ntot};
  uint32_t wb[3] = {(uint32_t)p.BK, (uint32_t)p.BN, 1};
  return encode_map(e, &out->maps.b, wgt, 3, wd, ws, wb, p.BK * 2);
}

int run_conv(glass_engine* e, const ConvLaunch& c, cudaStream_t s) {
  cudaError_t err;
  if (c.fused_down) {
    const bool timed = e->timing && e->ev_used + 2 <= e->ev.size();
    if (timed) cudaEventRecord(e->ev[e->ev_used], s);
    err = k_downconv_fused(c.maps.a, c.maps.b, c.fd_proj ? &c.maps2.a : nullptr, c.fd_proj ? &c.maps2.b : nullptr, c.fd_C,
                           c.fd_N, c.fd_Ho, c.fd_Wo, c.fd_Cout, c.p.epi.bias, c.p.epi.residual, c.p.epi.res_i8,
                           c.p.epi.out, c.p.epi.out_i8, c.p.epi.post_scale, e->num_sms, s);
    if (timed) {
      cudaEventRecord(e->ev[e->ev_used + 1], s);
      e->ev_conv[e->ev_used / 2] = &c;
      e->ev_used += 2;
    }
    e->launches++;
    if (err != cudaSuccess) return fail(GLASS_ERR_CUDA, "fused down-conv launch failed: %s", cudaGetErrorString(err));
    return GLASS_OK;
  }
  if (e->cfg.conv_impl == 0) {
    if (e->timing && e->ev_used + 2 <= e->ev.size()) {
      cudaEventRecord(e->ev[e->ev_used], s);
      err = launch_conv_tc(c.p, c.maps, e->num_sms, s);
      cudaEventRecord(e->ev[e->ev_used + 1], s);
      e->ev_conv[e->ev_used / 2] = &c;
      e->ev_used += 2;
    } else {
      err = launch_conv_tc(c.p, c.maps, e->num_sms, s);
    }
  } else {
    err = launch_conv_simt(c.p, s);
  }
  e->launches++;
  if (err != cudaSuccess) return fail(GLASS_ERR_CUDA, "conv launch failed: %s", cudaGetErrorString(err));
  return GLASS_OK;
}

#define LAUNCH(expr)                                                                       \
  do {                                                                                     \
    cudaError_t err__ = (expr);                                                            \
    e->launches++;                                                                         \
    if (err__ != cudaSuccess)                                                              \
      return fail(GLASS_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(err__));      \
  } while (0)
#define RC(expr)                 \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != GLASS_OK) return rc__; \
  } while (0)

int capture_f32(glass_engine* e, const std::string& name, const float* dev, size_t n, cudaStream_t s) {
  if (!e->capture) return GLASS_OK;
  std::vector<float>& v = e->captured[name];
  v.resize(n);
  CUDA_OK(cudaStreamSynchronize(s));
  CUDA_OK(cudaMemcpy(v.data(), dev, n * sizeof(float), cudaMemcpyDeviceToHost));
  return GLASS_OK;
}
// i8_w/i8_c > 0: the device tensor is [rows][C/8][W][8]; captured as NHWC [rows][W][C]
int capture_f16(glass_engine* e, const std::string& name, const __half* dev, size_t n, cudaStream_t s, int i8_w = 0,
                int i8_c = 0) {
  if (!e->capture) return GLASS_OK;
  std::vector<__half> tmp(n);
  CUDA_OK(cudaStreamSynchronize(s));
  CUDA_OK(cudaMemcpy(tmp.data(), dev, n * sizeof(__half), cudaMemcpyDeviceToHost));
  std::vector<float>& v = e->captured[name];
  v.resize(n);
  if (i8_w > 0) {
    const size_t G = i8_c / 8, rows = n / ((size_t)i8_w * i8_c);
    for (size_t r = 0; r < rows; ++r)
      for (size_t g = 0; g < G; ++g)
        for (size_t x = 0; x < (size_t)i8_w; ++x)
          for (size_t j = 0; j < 8; ++j)
            v[(r * i8_w + x) * i8_c + g * 8 + j] = __half2float(tmp[((r * G + g) * i8_w + x) * 8 + j]);
  } else {
    for (size_t i = 0; i < n; ++i) v[i] = __half2float(tmp[i]);
  }
  return GLASS_OK;
}

// ---------------------------------------------------------------------------
// architecture bookkeeping
// ---------------------------------------------------------------------------
void derive_arch(glass_engine* e) {
  const glass_config& c = e->cfg;
  e->gch.assign(c.channels, c.channels + c.num_blocks);
  e->glayers.clear();
  for (int b = 0; b < c.num_blocks; ++b) {
    const int res = 4 << b;
    if (b == 0) {
      e->glayers.push_back({0, 0, e->gch[0], e->gch[0], 0, res});
    } else {
      e->glayers.push_back({b, 0, e->gch[b - 1], e->gch[b], 1, res});
      e->glayers.push_back({b, 1, e->gch[b], e->gch[b], 0, res});
    }
  }
  e->conv_off.clear();
  e->rgb_off.clear();
  int off = 0;
  for (const GLayer& l : e->glayers) { e->conv_off.push_back(off); off += l.cin; }
  for (int b = 0; b < c.num_blocks; ++b) { e->rgb_off.push_back(off); off += e->gch[b]; }
  e->S = off;
  e->R = 4 << (c.num_blocks - 1);
  // Which up/down convs run in the exact polyphase form.  Cost model from profiles/r01_fir_passes_p64.txt: the
  // exact form saves 3/4 of the tensor work of the layer but adds a streaming pass over its output, so it pays
  // where the layer is tensor-bound (wide channels), not where it is HBM/epilogue-bound (32..256 channels).
  const bool folded = (c.flags & GLASS_FLAG_FOLDED_RESAMPLE) != 0;
  const bool exact_all = (c.flags & GLASS_FLAG_EXACT_RESAMPLE) != 0;
  // (debug builds: GLASS_DEBUG_GEXACT="cin_min,in_res_min", GLASS_DEBUG_DEXACT="ci_min" move the thresholds for A/B)
  int g_cin_min = 512, g_res_min = 64, d_ci_min = 128;
  if (const char* v = debug_env("GLASS_DEBUG_GEXACT")) sscanf(v, "%d,%d", &g_cin_min, &g_res_min);
  if (const char* v = debug_env("GLASS_DEBUG_DEXACT")) sscanf(v, "%d", &d_ci_min);
  e->g_exact.clear();
  for (const GLayer& l : e->glayers) {
    const int in_res = l.res / 2;
    const bool possible = !folded && l.up && in_res >= 16;
    e->g_exact.push_back((possible && (exact_all || (l.cin >= g_cin_min && in_res >= g_res_min))) ? 1 : 0);
  }
  e->d_exact.clear();
  for (int b = 0; b + 1 < c.num_blocks; ++b) {
    const bool possible = !folded && (e->R >> b) >= 32;
    e->d_exact.push_back((possible && (exact_all || e->gch[c.num_blocks - 1 - b] >= d_ci_min)) ? 1 : 0);
  }
  // activations feeding a 32/64-channel 3x3 conv: channel-group-interleaved (MODE 4), written by the producing
  // conv's epilogue (not by k_upfir) or by k_from_rgb
  const bool i8_ok = (c.flags & GLASS_FLAG_NO_I8_LAYOUT) == 0 && c.conv_impl == 0;
  e->g_in_i8.clear();
  for (size_t li = 0; li < e->glayers.size(); ++li) {
    const GLayer& l = e->glayers[li];
    const int in_res = l.up ? l.res / 2 : l.res;
    const bool ok = i8_ok && li >= 1 && !e->g_exact[li] && !e->g_exact[li - 1] && (l.cin == 32 || l.cin == 64) &&
                    in_res >= 16 && (l.up ? 4 * l.cout : l.cout) % 32 == 0;
    e->g_in_i8.push_back(ok ? 1 : 0);
  }
  e->d_in_i8.clear();
  for (int b = 0; b + 1 < c.num_blocks; ++b) {
    const int Ci = e->gch[c.num_blocks - 1 - b];
    e->d_in_i8.push_back((i8_ok && (Ci == 32 || Ci == 64) && (e->R >> b) >= 16) ? 1 : 0);
  }
  e->d_c1_i8.clear();
  for (int b = 0; b + 1 < c.num_blocks; ++b) {
    const int Ci = e->gch[c.num_blocks - 1 - b];
    // 4*Ci = 128 channels: conv_tc MODE 6 (haloed I8 box per tile pair, the nine taps streamed through a ring).
    // (MODE 4 with nine resident 128-channel taps only leaves room for BN = 32 and measured slower than the
    // streamed NHWC form: 5.35 vs 4.45 ms at P=64.)  GLASS_FLAG_C1_NHWC keeps the space-to-depth tensor NHWC (cross-check).
    const bool on = (c.flags & GLASS_FLAG_C1_NHWC) == 0;
    e->d_c1_i8.push_back((i8_ok && on && !e->d_exact[b] && Ci == 32 && (e->R >> b) / 2 >= 16 &&
                          e->gch[c.num_blocks - 2 - b] % 64 == 0) ? 1 : 0);
  }
  e->d_proj_fused.clear();
  for (int b = 0; b + 1 < c.num_blocks; ++b) {
    const int Ci = e->gch[c.num_blocks - 1 - b], Co = e->gch[c.num_blocks - 2 - b];
    const bool on = (c.flags & GLASS_FLAG_PROJ_FUSION) != 0 && c.conv_impl == 0;   // opt-in: measured slower
    e->d_proj_fused.push_back((on && k_fir_proj_supported(Ci, Co, false)) ? 1 : 0);
  }
  // The projection output dR is written by a K <= 512, 1x1 GEMM whose epilogue is bound by its own stores: NHWC
  // stores put 16 bytes per lane at the pixel pitch (32 shared/L1 wavefronts per instruction), I8 stores are 128-byte
  // contiguous per 8 lanes (4 wavefronts), and conv1's residual loads coalesce the same way.  (GLASS_DEBUG_RES_I8=0:
  // NHWC, for A/B.)
  e->d_res_i8.clear();
  for (int b = 0; b + 1 < c.num_blocks; ++b) {
    const int Co = e->gch[c.num_blocks - 2 - b], ro = (e->R >> b) / 2;
    const char* env = debug_env("GLASS_DEBUG_RES_I8");
    const bool on = env == nullptr || atoi(env) != 0;
    e->d_res_i8.push_back((i8_ok && on && !e->d_proj_fused[b] && ro >= 16 && ro % 16 == 0 && Co % 16 == 0) ? 1 : 0);
  }
  // 32/64-channel blocks at >= 64x64: exact down-conv with the FIR inside the kernel instead of the folded form (4x the
  // MACs).  Needs conv0's output in the plain I8 layout and the projection residual / block output as planned above.
  e->d_fused.clear();
  for (int b = 0; b + 1 < c.num_blocks; ++b) {
    const int Ci = e->gch[c.num_blocks - 1 - b], Co = e->gch[c.num_blocks - 2 - b], ro = (e->R >> b) / 2;
    // (64-channel block: one 64-column n-tile of taps is resident, so the blur of a pixel tile is repeated per n-tile;
    // its conv1 time equals the folded form's (2.5 vs 2.3 ms) but conv0 stores plain I8 instead of space-to-depth NHWC
    // (1.1 vs 1.3 ms) and the step issues 4x fewer MMAs on that layer: 1.5 ms per step at P = 64, measured)
    static const bool nofused64 = debug_env("GLASS_DEBUG_NOFUSED64") != nullptr;      // (debug builds: A/B)
    const bool on = i8_ok && !folded && !(c.flags & GLASS_FLAG_NO_FUSED_DOWN) && !e->d_exact[b] &&
                    k_downconv_fused_supported(Ci, Co, ro, ro) && ro >= 32 && !(Ci == 64 && nofused64);
    e->d_fused.push_back(on ? 1 : 0);
    if (on) e->d_c1_i8[b] = 0;
  }
  // ... and where that kernel can also carry the block's projection (the 32 -> 64 block): no projection launch, no dR
  e->d_c1_proj.clear();
  for (int b = 0; b + 1 < c.num_blocks; ++b) {
    const int Ci = e->gch[c.num_blocks - 1 - b], Co = e->gch[c.num_blocks - 2 - b];
    e->d_c1_proj.push_back((e->d_fused[b] && !e->d_proj_fused[b] && !(c.flags & GLASS_FLAG_NO_PROJ_ACC) &&
                            k_downconv_proj_supported(Ci, Co)) ? 1 : 0);
  }
  const bool pair_ok = (c.flags & GLASS_FLAG_NO_PAIR_PACK) == 0 && c.conv_impl == 0;
  e->g_pair.clear();
  for (size_t li = 0; li < e->glayers.size(); ++li) {
    const GLayer& l = e->glayers[li];
    e->g_pair.push_back((pair_ok && !e->g_in_i8[li] && !l.up && l.cin == 32 && l.cout == 32 && l.res >= 32) ? 1 : 0);
  }
  e->d_pair.clear();
  for (int b = 0; b + 1 < c.num_blocks; ++b)
    e->d_pair.push_back((pair_ok && !e->d_in_i8[b] && e->gch[c.num_blocks - 1 - b] == 32 && (e->R >> b) >= 32) ? 1 : 0);
  e->noise_layer_off.clear();
  size_t noff = 0;
  for (const GLayer& l : e->glayers) { e->noise_layer_off.push_back(noff); noff += (size_t)l.res * l.res; }
  e->noise_per_group = noff;
}

int validate_weights(glass_engine* e) {
  const glass_config& c = e->cfg;
  const int L = c.latent_size;
  char nm[64];
  for (int i = 0; i < c.mapping_layers; ++i) {
    snprintf(nm, sizeof nm, "g.map.w%d", i); RC(check_tensor(e, nm, (size_t)L * L * 4));
    snprintf(nm, sizeof nm, "g.map.b%d", i); RC(check_tensor(e, nm, (size_t)L * 4));
  }
  RC(check_tensor(e, "g.style.w", (size_t)L * e->S * 4));
  RC(check_tensor(e, "g.style.b", (size_t)e->S * 4));
  RC(check_tensor(e, "g.const", (size_t)16 * e->gch[0] * 4));
  for (size_t li = 0; li < e->glayers.size(); ++li) {
    const GLayer& l = e->glayers[li];
    const int ntot = l.up ? 4 * l.cout : l.cout;
    snprintf(nm, sizeof nm, "g.conv%zu.w", li); RC(check_tensor(e, nm, (size_t)9 * ntot * l.cin * 2));
    if (e->g_exact[li]) { snprintf(nm, sizeof nm, "g.conv%zu.wx", li); RC(check_tensor(e, nm, (size_t)4 * ntot * l.cin * 2)); }
    if (e->g_pair[li]) { snprintf(nm, sizeof nm, "g.conv%zu.wp", li); RC(check_tensor(e, nm, (size_t)9 * 2 * l.cout * 2 * l.cin * 2)); }
    snprintf(nm, sizeof nm, "g.conv%zu.wsq", li); RC(check_tensor(e, nm, (size_t)l.cin * l.cout * 4));
    snprintf(nm, sizeof nm, "g.conv%zu.bias", li); RC(check_tensor(e, nm, (size_t)l.cout * 4));
    snprintf(nm, sizeof nm, "g.conv%zu.nstr", li); RC(check_tensor(e, nm, 4));
  }
  for (int b = 0; b < c.num_blocks; ++b) {
    snprintf(nm, sizeof nm, "g.rgb%d.w", b); RC(check_tensor(e, nm, (size_t)3 * e->gch[b] * 4));
    snprintf(nm, sizeof nm, "g.rgb%d.bias", b); RC(check_tensor(e, nm, 12));
  }
  const int Wd = c.clip_width, T = (c.clip_resolution / c.clip_patch) * (c.clip_resolution / c.clip_patch) + 1;
  RC(check_tensor(e, "c.patch.w", (size_t)Wd * 3 * c.clip_patch * c.clip_patch * 2));
  RC(check_tensor(e, "c.cls", (size_t)Wd * 4));
  RC(check_tensor(e, "c.pos", (size_t)T * Wd * 4));
  for (const char* n : {"c.lnpre.w", "c.lnpre.b", "c.lnpost.w", "c.lnpost.b"}) RC(check_tensor(e, n, (size_t)Wd * 4));
  RC(check_tensor(e, "c.proj", (size_t)Wd * c.clip_embed_dim * 4));
  for (int l = 0; l < c.clip_layers; ++l) {
    auto nmf = [&](const char* suffix) { snprintf(nm, sizeof nm, "c.l%d.%s", l, suffix); return std::string(nm); };
    for (const char* n : {"ln1.w", "ln1.b", "ln2.w", "ln2.b", "out.b", "proj.b"}) RC(check_tensor(e, nmf(n), (size_t)Wd * 4));
    RC(check_tensor(e, nmf("qkv.w"), (size_t)3 * Wd * Wd * 2));
    RC(check_tensor(e, nmf("qkv.b"), (size_t)3 * Wd * 4));
    RC(check_tensor(e, nmf("out.w"), (size_t)Wd * Wd * 2));
    RC(check_tensor(e, nmf("fc.w"), (size_t)4 * Wd * Wd * 2));
    RC(check_tensor(e, nmf("fc.b"), (size_t)4 * Wd * 4));
    RC(check_tensor(e, nmf("proj.w"), (size_t)4 * Wd * Wd * 2));
  }
  if (c.use_discriminator) {
    // D channel order is last-G-layer first
    const int nb = c.num_blocks;
    auto dch = [&](int i) { return e->gch[nb - 1 - i]; };
    RC(check_tensor(e, "d.frgb.w", (size_t)3 * dch(0) * 4));
    RC(check_tensor(e, "d.frgb.b", (size_t)dch(0) * 4));
    for (int b = 0; b < nb - 1; ++b) {
      auto nmf = [&](const char* suffix) { snprintf(nm, sizeof nm, "d.b%d.%s", b, suffix); return std::string(nm); };
      RC(check_tensor(e, nmf("c0.w"), (size_t)9 * dch(b) * dch(b) * 2));
      RC(check_tensor(e, nmf("c0.b"), (size_t)dch(b) * 4));
      RC(check_tensor(e, nmf("c1.w"), (size_t)9 * dch(b + 1) * 4 * dch(b) * 2));
      if (e->d_exact[b]) RC(check_tensor(e, nmf("c1.wx"), (size_t)4 * dch(b + 1) * 4 * dch(b) * 2));
      if (e->d_fused[b]) RC(check_tensor(e, nmf("c1.w9"), (size_t)9 * dch(b + 1) * dch(b) * 2));
      if (e->d_pair[b]) RC(check_tensor(e, nmf("c0.wp"), (size_t)9 * 2 * dch(b) * 2 * dch(b) * 2));
      RC(check_tensor(e, nmf("c1.b"), (size_t)dch(b + 1) * 4));
      RC(check_tensor(e, nmf("proj.w"), (size_t)dch(b + 1) * dch(b) * 2));
    }
    const int C = dch(nb - 1);
    const int cpad = ((C + 1 + 63) / 64) * 64;
    RC(check_tensor(e, "d.fin.w", (size_t)9 * C * cpad * 2));
    RC(check_tensor(e, "d.fin.b", (size_t)C * 4));
    RC(check_tensor(e, "d.dense0.w", (size_t)C * 16 * C * 2));
    RC(check_tensor(e, "d.dense0.b", (size_t)C * 4));
    RC(check_tensor(e, "d.dense1.w", (size_t)C * 4));
    RC(check_tensor(e, "d.dense1.b", 4));
  }
  return GLASS_OK;
}

constexpr size_t kSlackRows = 256;   // rows of slack behind every TMA-read buffer (boxes may overhang)

void layout_workspace(glass_engine* e, Arena& a) {
  const glass_config& c = e->cfg;
  const size_t P = c.max_population;
  const int L = c.latent_size;
  const int nb = c.num_blocks;
  e->x_host_stage = (double*)a.take(P * L * 8);
  e->gather_rows = (int*)a.take(P * 4);
  e->z32 = (float*)a.take(P * L * 4);
  e->wA = (float*)a.take(P * L * 4);
  e->wB = (float*)a.take(P * L * 4);
  e->styles = (float*)a.take(P * e->S * 4);
  e->styles_n = (float*)a.take(P * e->S * 4);
  e->mscale = (float*)a.take(P * e->glayers.size() * 4);
  e->range_ctr = (unsigned long long*)a.take(4 * 8);
  e->max_groups = (int)(P / c.batch_size);
  e->noise = (float*)a.take((size_t)e->max_groups * e->noise_per_group * 4);
  e->dmod.clear();
  for (const GLayer& l : e->glayers) e->dmod.push_back((float*)a.take(P * l.cout * 4));
  e->rgbw.clear();
  for (int b = 0; b < nb; ++b) e->rgbw.push_back((float*)a.take(P * 3 * e->gch[b] * 4));
  size_t act_elems = P * 16 * e->gch[0];
  size_t slab_elems = 0;
  for (const GLayer& l : e->glayers) {
    act_elems = std::max(act_elems, P * (size_t)l.res * l.res * l.cout);
    const int bn = pick_bn(l.cout);
    if (bn) slab_elems = std::max(slab_elems, (size_t)(l.cout / 32) * P * l.res * l.res);   // worst case BN = 32
  }
  // every activation row is at least 32 channels wide; slack covers box overhang in the outermost dim
  const size_t slack = kSlackRows * 512;
  e->actA = (__half*)a.take((act_elems + slack) * 2);
  e->actB = (__half*)a.take((act_elems + slack) * 2);
  {
    size_t c_elems = 0;
    for (size_t li = 0; li < e->glayers.size(); ++li)
      if (e->g_exact[li]) c_elems = std::max(c_elems, P * (size_t)(e->glayers[li].res + 2) * (e->glayers[li].res + 2) * e->glayers[li].cout);
    for (size_t b = 0; b < e->d_exact.size(); ++b)
      if (e->d_exact[b]) {
        const size_t r2 = (size_t)(e->R >> b) / 2 + 1;
        c_elems = std::max(c_elems, P * r2 * r2 * 4 * (size_t)e->gch[nb - 1 - b]);
      }
    e->actC = c_elems ? (__half*)a.take((c_elems + slack) * 2) : nullptr;
  }
  e->slabs = (float4*)a.take(slab_elems * 16);
  e->yA = (float4*)a.take(P * (size_t)e->R * e->R * 16);
  e->yB = (float4*)a.take(P * (size_t)e->R * e->R * 16);
  e->images = (float*)a.take(P * 3 * (size_t)e->R * e->R * 4);
  // CLIP
  const int g = c.clip_resolution / c.clip_patch, T = g * g + 1, Wd = c.clip_width;
  const size_t kdim = (size_t)3 * c.clip_patch * c.clip_patch;
  e->patches = (__half*)a.take((P * g * g + kSlackRows) * kdim * 2);
  e->patch_emb = (__half*)a.take((P * g * g + kSlackRows) * Wd * 2);
  e->tokens = (__half*)a.take((P * T + kSlackRows) * Wd * 2);
  e->hbuf = (__half*)a.take((P * T + kSlackRows) * Wd * 2);
  e->qkv = (__half*)a.take((P * T + kSlackRows) * 3 * Wd * 2);
  e->att = (__half*)a.take((P * T + kSlackRows) * Wd * 2);
  e->fc = (__half*)a.take((P * T + kSlackRows) * 4 * Wd * 2);
  e->features = (float*)a.take(P * c.clip_embed_dim * 4);
  e->sim = (float*)a.take(P * 4);
  e->neg_sim = (float*)a.take(P * 4);
  e->text = (float*)a.take((size_t)c.clip_embed_dim * 4);
  if (c.use_discriminator) {
    const int C0 = e->gch[nb - 1], C1 = nb > 1 ? e->gch[nb - 2] : C0;
    const size_t half_pix = P * (size_t)(e->R / 2) * (e->R / 2);
    e->dXd = (__half*)a.take((half_pix * C0 + slack) * 2);
    // channels at most double per block while pixels quarter, so block 0 bounds these
    e->dR = (__half*)a.take((half_pix * std::max(C1, 2 * C0) + slack) * 2);
    e->dOut = (__half*)a.take((half_pix * std::max(C1, 2 * C0) + slack) * 2);
    const int C = e->gch[0];
    const int cpad = ((C + 1 + 63) / 64) * 64;
    e->dFin = (__half*)a.take((P * 16 * cpad + slack) * 2);
    e->dFinOut = (__half*)a.take((P + kSlackRows) * 16 * C * 2);   // read as a [P x 16C] GEMM operand
    e->dDense = (__half*)a.take((P * C + slack) * 2);
    e->dlogits = (float*)a.take(P * 4);
    e->hinge = (float*)a.take(P * 4);
  }
}

EpiParams epi_default() {
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));
  ep.post_scale = 1.f;
  ep.noise_group_div = 1;
  ep.x_phases = 1;
  return ep;
}

// ---------------------------------------------------------------------------
// plan: all conv/GEMM launches for a population of P candidates
// ---------------------------------------------------------------------------
void drop_graph(glass_engine* e) {
  if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
  e->graph_exec = nullptr;
  e->graph_evals = 0;
}

int build_plan(glass_engine* e, int P) {
  drop_graph(e);
  const glass_config& c = e->cfg;
  e->g_convs.clear();
  e->c_convs.clear();
  e->d_convs.clear();
  char nm[64];
  // ---- G ----
  __half* bufs[2] = {e->actA, e->actB};
  int cur = 0;   // x0 lives in actA
  const size_t nl = e->glayers.size();
  for (size_t li = 0; li < nl; ++li) {
    const GLayer& l = e->glayers[li];
    const bool last_in_block = (li + 1 == nl) || (e->glayers[li + 1].block != l.block);
    const int in_res = l.up ? l.res / 2 : l.res;
    EpiParams ep = epi_default();
    ep.dmod = e->dmod[li];
    snprintf(nm, sizeof nm, "g.conv%zu.bias", li); ep.bias = tptr<float>(e, nm);
    snprintf(nm, sizeof nm, "g.conv%zu.nstr", li); ep.noise_strength = tptr<float>(e, nm);
    ep.noise = e->noise + e->noise_layer_off[li];
    ep.noise_group_div = c.batch_size;
    ep.noise_group_stride = e->noise_per_group;
    ep.act = kActLrelu;
    ep.Cout = l.cout;
    ep.store_mode = l.up ? kStoreDepthToSpace : kStoreRegular;
    if (last_in_block) {
      ep.rgb_w = e->rgbw[l.block];
      ep.rgb_out = e->slabs;
    }
    if (li + 1 < nl) {
      ep.out_scale = e->styles_n + e->conv_off[li + 1];
      ep.out_scale_stride = e->S;
      ep.out = bufs[cur ^ 1];
      ep.out_i8 = e->g_in_i8[li + 1];
    } else {
      ep.out = nullptr;
    }
    ConvLaunch cl;
    if (e->g_exact[li]) {
      // exact polyphase: transposed conv as a 2x2-tap conv on the (H+1)x(W+1) grid -> u (demodulated, fp16) in actC;
      // FIR + noise + bias + activation + next-style pre-scale happen in k_upfir (run_generator)
      EpiParams eu = epi_default();
      eu.dmod = e->dmod[li];
      eu.Cout = l.cout;
      eu.store_mode = kStoreDepthToSpace;
      eu.out = e->actC;
      snprintf(nm, sizeof nm, "g.conv%zu.wx", li);
      RC(make_conv(e, &cl, bufs[cur], P, in_res + 1, in_res + 1, l.cin, tptr<__half>(e, nm), 4, 4 * l.cout, eu, false,
                   kUpExactTaps, in_res, in_res, false, 1, l.cout));
    } else if (e->g_in_i8[li]) {
      snprintf(nm, sizeof nm, "g.conv%zu.w", li);
      RC(make_conv(e, &cl, bufs[cur], P, in_res, in_res, l.cin, tptr<__half>(e, nm), 9, l.up ? 4 * l.cout : l.cout, ep,
                   false, nullptr, 0, 0, true));
    } else if (e->g_pair[li]) {
      // 32-channel layer on paired pixels: [H][W/2][64] view, N = 2*Cout, the epilogue's column halves are pixels
      ep.x_phases = 2;
      snprintf(nm, sizeof nm, "g.conv%zu.wp", li);
      RC(make_conv(e, &cl, bufs[cur], P, in_res, in_res / 2, 2 * l.cin, tptr<__half>(e, nm), 9, 2 * l.cout, ep, false));
    } else {
      snprintf(nm, sizeof nm, "g.conv%zu.w", li);
      RC(make_conv(e, &cl, bufs[cur], P, in_res, in_res, l.cin, tptr<__half>(e, nm), 9, l.up ? 4 * l.cout : l.cout, ep,
                   false));
    }
    // the noise tensor index inside a group is per-layer; stride between groups is noise_per_group
    cl.flops = 2.0 * 9.0 * (double)P * in_res * in_res * l.cin * l.cout;   // as written by the reference
    e->g_convs.push_back(cl);
    cur ^= 1;
  }
  // ---- CLIP ----
  {
    const int g = c.clip_resolution / c.clip_patch, T = g * g + 1, Wd = c.clip_width;
    const int kdim = 3 * c.clip_patch * c.clip_patch;
    const int Mp = P * g * g, M = P * T;
    EpiParams ep = epi_default();
    ep.Cout = Wd; ep.out = e->patch_emb;
    ConvLaunch cl;
    RC(make_conv(e, &cl, e->patches, 1, 1, Mp, kdim, tptr<__half>(e, "c.patch.w"), 1, Wd, ep, true));
    cl.flops = 2.0 * Mp * (double)kdim * Wd;
    e->c_convs.push_back(cl);
    for (int l = 0; l < c.clip_layers; ++l) {
      auto nmf = [&](const char* suffix) { snprintf(nm, sizeof nm, "c.l%d.%s", l, suffix); return std::string(nm); };
      // qkv
      ep = epi_default(); ep.Cout = 3 * Wd; ep.bias = tptr<float>(e, nmf("qkv.b")); ep.out = e->qkv;
      RC(make_conv(e, &cl, e->hbuf, 1, 1, M, Wd, tptr<__half>(e, nmf("qkv.w")), 1, 3 * Wd, ep, true));
      cl.flops = 2.0 * M * (double)Wd * 3 * Wd; e->c_convs.push_back(cl);
      // out proj + residual (in place on tokens: each element read then written by the same thread)
      ep = epi_default(); ep.Cout = Wd; ep.bias = tptr<float>(e, nmf("out.b")); ep.round_fp16_before_act = 1;
      ep.residual = e->tokens; ep.out = e->tokens;
      RC(make_conv(e, &cl, e->att, 1, 1, M, Wd, tptr<__half>(e, nmf("out.w")), 1, Wd, ep, true));
      cl.flops = 2.0 * M * (double)Wd * Wd; e->c_convs.push_back(cl);
      // fc + QuickGELU
      ep = epi_default(); ep.Cout = 4 * Wd; ep.bias = tptr<float>(e, nmf("fc.b")); ep.round_fp16_before_act = 1;
      ep.act = kActQuickGelu; ep.out = e->fc;
      RC(make_conv(e, &cl, e->hbuf, 1, 1, M, Wd, tptr<__half>(e, nmf("fc.w")), 1, 4 * Wd, ep, true));
      cl.flops = 2.0 * M * (double)Wd * 4 * Wd; e->c_convs.push_back(cl);
      // proj + residual
      ep = epi_default(); ep.Cout = Wd; ep.bias = tptr<float>(e, nmf("proj.b")); ep.round_fp16_before_act = 1;
      ep.residual = e->tokens; ep.out = e->tokens;
      RC(make_conv(e, &cl, e->fc, 1, 1, M, 4 * Wd, tptr<__half>(e, nmf("proj.w")), 1, Wd, ep, true));
      cl.flops = 2.0 * M * (double)Wd * 4 * Wd; e->c_convs.push_back(cl);
    }
  }
  // ---- D ----
  if (c.use_discriminator) {
    const int nb = c.num_blocks;
    auto dch = [&](int i) { return e->gch[nb - 1 - i]; };
    {
      // folded fromRGB constants (kernels.cu: from_rgb_fir_kernel) on the host: they travel as kernel parameters
      const int C0 = dch(0);
      std::vector<float> w(3 * (size_t)C0), bs(C0);
      CUDA_OK(cudaMemcpy(w.data(), tptr<float>(e, "d.frgb.w"), w.size() * 4, cudaMemcpyDeviceToHost));
      CUDA_OK(cudaMemcpy(bs.data(), tptr<float>(e, "d.frgb.b"), bs.size() * 4, cudaMemcpyDeviceToHost));
      e->frgb_folded.assign(4 * (size_t)C0, 0.f);
      for (int ch = 0; ch < C0; ++ch) {
        for (int r = 0; r < 3; ++r) e->frgb_folded[r * C0 + ch] = 2.f * kSqrt2 * w[r * C0 + ch];
        e->frgb_folded[3 * C0 + ch] = kSqrt2 * (bs[ch] - w[ch] - w[C0 + ch] - w[2 * C0 + ch]);
      }
    }
    __half* x = e->actA;          // fromRGB output
    __half* outs[2] = {e->dOut, e->actA};
    int res = e->R;
    for (int b = 0; b < nb - 1; ++b) {
      const int Ci = dch(b), Co = dch(b + 1);
      auto nmf = [&](const char* suffix) { snprintf(nm, sizeof nm, "d.b%d.%s", b, suffix); return std::string(nm); };
      ConvLaunch cl;
      // conv0: 3x3 Ci->Ci, bias, lrelu; stored space-to-depth for the folded conv1, or plain NHWC for the blur pass
      EpiParams ep = epi_default();
      ep.Cout = Ci; ep.bias = tptr<float>(e, nmf("c0.b")); ep.act = kActLrelu; ep.out = e->actB;
      ep.store_mode = (e->d_exact[b] || e->d_fused[b]) ? kStoreRegular : kStoreSpaceToDepth;
      ep.out_i8 = e->d_fused[b] ? 1 : e->d_c1_i8[b];
      if (e->d_in_i8[b]) {
        RC(make_conv(e, &cl, x, P, res, res, Ci, tptr<__half>(e, nmf("c0.w")), 9, Ci, ep, false, nullptr, 0, 0, true));
      } else if (e->d_pair[b]) {
        ep.x_phases = 2;
        if (ep.store_mode == kStoreSpaceToDepth) ep.store_mode = kStoreSpaceToDepthY;
        RC(make_conv(e, &cl, x, P, res, res / 2, 2 * Ci, tptr<__half>(e, nmf("c0.wp")), 9, 2 * Ci, ep, false));
      } else {
        RC(make_conv(e, &cl, x, P, res, res, Ci, tptr<__half>(e, nmf("c0.w")), 9, Ci, ep, false));
      }
      cl.flops = 2.0 * 9.0 * (double)P * res * res * Ci * Ci; e->d_convs.push_back(cl);
      // projection: 1x1 on the FIR-downsampled input
      ep = epi_default(); ep.Cout = Co; ep.out = e->dR; ep.out_i8 = e->d_res_i8[b];
      RC(make_conv(e, &cl, e->dXd, P, res / 2, res / 2, Ci, tptr<__half>(e, nmf("proj.w")), 1, Co, ep, false));
      cl.flops = 2.0 * (double)P * (res / 2) * (res / 2) * Ci * Co; e->d_convs.push_back(cl);
      // conv1
      ep = epi_default();
      ep.Cout = Co; ep.bias = tptr<float>(e, nmf("c1.b")); ep.act = kActLrelu; ep.residual = e->dR;
      ep.res_i8 = e->d_res_i8[b];
      ep.post_scale = kInvSqrt2; ep.out = outs[b & 1];
      ep.out_i8 = (b + 1 < nb - 1) ? e->d_in_i8[b + 1] : 0;
      if (e->d_fused[b]) {
        // exact, FIR inside the kernel (downconv_tc.cu): reads conv0's I8 output, no space-to-depth tensor, no blur pass
        memset(&cl.p, 0, sizeof(cl.p));
        cl.p.epi = ep;
        cl.fused_down = true;
        cl.fd_C = Ci; cl.fd_N = P; cl.fd_Ho = res / 2; cl.fd_Wo = res / 2; cl.fd_Cout = Co;
        const uint64_t G = Ci / 8;
        uint64_t d4[4] = {(uint64_t)res * 8, G, (uint64_t)res, (uint64_t)P};
        uint64_t s4[3] = {(uint64_t)res * 16, G * res * 16, (uint64_t)res * G * res * 16};
        int box_groups = 4, box_cols = 64;        // TMA boxes of the kernel instance for this block
        k_downconv_fused_geometry(Ci, Co, &box_groups, &box_cols);
        uint32_t b4[4] = {160, (uint32_t)box_groups, 36, 1};
        RC(encode_map(e, &cl.maps.a, e->actB, 4, d4, s4, b4, 0));
        uint64_t wd[3] = {(uint64_t)Ci, (uint64_t)Co, 9};
        uint64_t ws[2] = {(uint64_t)Ci * 2, (uint64_t)Ci * 2 * Co};
        uint32_t wb[3] = {(uint32_t)Ci, (uint32_t)box_cols, 1};
        RC(encode_map(e, &cl.maps.b, tptr<__half>(e, nmf("c1.w9")), 3, wd, ws, wb, Ci * 2));
        if (e->d_c1_proj[b]) {
          // projection as a second accumulator: tile of the FIR-downsampled block input (NHWC) + the 1x1 weights
          cl.fd_proj = true;
          const int ro = res / 2;
          uint64_t xd4[4] = {(uint64_t)Ci, (uint64_t)ro, (uint64_t)ro, (uint64_t)P};
          uint64_t xs4[3] = {(uint64_t)Ci * 2, (uint64_t)Ci * 2 * ro, (uint64_t)Ci * 2 * ro * ro};
          uint32_t xb4[4] = {(uint32_t)Ci, 8, 16, 1};
          RC(encode_map(e, &cl.maps2.a, e->dXd, 4, xd4, xs4, xb4, Ci * 2));
          uint64_t pd[2] = {(uint64_t)Ci, (uint64_t)Co};
          uint64_t ps[1] = {(uint64_t)Ci * 2};
          uint32_t pb[2] = {(uint32_t)Ci, 64};
          RC(encode_map(e, &cl.maps2.b, tptr<__half>(e, nmf("proj.w")), 2, pd, ps, pb, Ci * 2));
        }
      } else if (e->d_exact[b]) {
        // exact: blurred input (k_blur_s2d -> actC, [(res/2+1)^2][4*Ci]) then a 2x2-tap conv == 3x3 stride 2
        RC(make_conv(e, &cl, e->actC, P, res / 2, res / 2, 4 * Ci, tptr<__half>(e, nmf("c1.wx")), 4, Co, ep, false,
                     kDownExactTaps, res / 2 + 1, res / 2 + 1, false, 2, Ci));
      } else {
        // folded FIR + 3x3 stride 2 == 3x3 over the space-to-depth tensor (4*Ci channels)
        RC(make_conv(e, &cl, e->actB, P, res / 2, res / 2, 4 * Ci, tptr<__half>(e, nmf("c1.w")), 9, Co, ep, false,
                     nullptr, 0, 0, e->d_c1_i8[b] != 0));
      }
      cl.flops = 2.0 * 9.0 * (double)P * (res / 2) * (res / 2) * Ci * Co; e->d_convs.push_back(cl);
      x = outs[b & 1];
      res /= 2;
    }
    const int C = dch(nb - 1);
    const int cpad = ((C + 1 + 63) / 64) * 64;
    ConvLaunch cl;
    EpiParams ep = epi_default();
    ep.Cout = C; ep.bias = tptr<float>(e, "d.fin.b"); ep.act = kActLrelu; ep.out = e->dFinOut;
    RC(make_conv(e, &cl, e->dFin, P, 4, 4, cpad, tptr<__half>(e, "d.fin.w"), 9, C, ep, false));
    cl.flops = 2.0 * 9.0 * (double)P * 16 * (C + 1) * C; e->d_convs.push_back(cl);
    ep = epi_default();
    ep.Cout = C; ep.bias = tptr<float>(e, "d.dense0.b"); ep.act = kActLrelu; ep.out = e->dDense;
    RC(make_conv(e, &cl, e->dFinOut, 1, 1, P, 16 * C, tptr<__half>(e, "d.dense0.w"), 1, C, ep, true));
    cl.flops = 2.0 * (double)P * 16 * C * C; e->d_convs.push_back(cl);
  }
  e->plan_pop = P;
  return GLASS_OK;
}

// ---------------------------------------------------------------------------
// stages
// ---------------------------------------------------------------------------
int fill_noise(glass_engine* e, int P, const glass_noise* nz, cudaStream_t s) {
  const size_t groups = (size_t)P / e->cfg.batch_size;
  const size_t n = groups * e->noise_per_group;
  if (nz != nullptr && nz->noise != nullptr) {
    CUDA_OK(cudaMemcpyAsync(e->noise, nz->noise, n * 4,
                            nz->noise_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  } else {
    const uint64_t first = nz ? nz->first_group : 0ull;
    LAUNCH(k_noise(e->noise, n, nz ? nz->seed : 0ull, first * (e->noise_per_group / 4), s));
  }
  return GLASS_OK;
}

int run_generator(glass_engine* e, const float* z, int P, const glass_noise* nz, float* images_out, cudaStream_t s,
                  bool noise_ready = false) {
  const glass_config& c = e->cfg;
  const int L = c.latent_size;
  char nm[64];
  if (!noise_ready) RC(fill_noise(e, P, nz, s));
  RC(capture_f32(e, "noise", e->noise, (size_t)(P / c.batch_size) * e->noise_per_group, s));
  // mapping (stylegan2/models.py:590-627)
  LAUNCH(k_pixelnorm(z, e->wA, P, L, s));
  float* cur = e->wA;
  float* nxt = e->wB;
  for (int i = 0; i < c.mapping_layers; ++i) {
    snprintf(nm, sizeof nm, "g.map.w%d", i);
    const float* w = tptr<float>(e, nm);
    snprintf(nm, sizeof nm, "g.map.b%d", i);
    LAUNCH(k_vecmat(cur, L, w, tptr<float>(e, nm), nxt, L, P, L, L, 1, s));
    std::swap(cur, nxt);
  }
  RC(capture_f32(e, "w", cur, (size_t)P * L, s));
  // all style affines in one pass (same dlatent for every layer, models.py:427-430)
  LAUNCH(k_vecmat(cur, L, tptr<float>(e, "g.style.w"), tptr<float>(e, "g.style.b"), e->styles, e->S, P, L, e->S, 0, s));
  RC(capture_f32(e, "styles", e->styles, (size_t)P * e->S, s));
  // fp16 range safety: conv styles normalised per (sample, layer) by a power of two, folded back into dmod below
  {
    StyleSlices sl;
    sl.n = (int)e->glayers.size();
    for (int li = 0; li < sl.n; ++li) { sl.off[li] = e->conv_off[li]; sl.cin[li] = e->glayers[li].cin; }
    LAUNCH(k_style_norm(e->styles, e->styles_n, e->mscale, e->S, P, sl, s));
  }
  // demodulation coefficients of every layer (modules.py:945-954): independent GEMVs, kMaxVecmatJobs per launch
  for (size_t l0 = 0; l0 < e->glayers.size(); l0 += kMaxVecmatJobs) {
    VecmatBatch vb;
    vb.n = 0;
    for (size_t li = l0; li < e->glayers.size() && vb.n < kMaxVecmatJobs; ++li) {
      const GLayer& l = e->glayers[li];
      snprintf(nm, sizeof nm, "g.conv%zu.wsq", li);
      vb.job[vb.n++] = VecmatJob{e->styles + e->conv_off[li], tptr<float>(e, nm), e->dmod[li], l.cin, l.cout,
                                 e->mscale + li, (int)e->glayers.size()};
    }
    LAUNCH(k_vecmat_batched(vb, e->S, P, 2, s));
  }
  for (int b = 0; b < c.num_blocks; ++b) {
    snprintf(nm, sizeof nm, "g.rgb%d.w", b);
    LAUNCH(k_rgb_weights(tptr<float>(e, nm), e->styles + e->rgb_off[b], e->S, e->rgbw[b], P, e->gch[b], s));
  }
  LAUNCH(k_const_input(tptr<float>(e, "g.const"), e->styles_n + e->conv_off[0], e->S, e->actA, P, e->gch[0], s));
  RC(capture_f16(e, "x0", e->actA, (size_t)P * 16 * e->gch[0], s));
  float4* ybuf[2] = {e->yA, e->yB};
  int ycur = 0;
  bool have_y = false;
  const size_t nl = e->glayers.size();
  for (size_t li = 0; li < nl; ++li) {
    const GLayer& l = e->glayers[li];
    ConvLaunch& cl = e->g_convs[li];
    // Last conv with a single toRGB slab: skip sum, upsample, bias and biggan_norm run in its epilogue and the image
    // is written from there (EpiParams::image); otherwise (debug captures, SIMT bring-up path, pixel-pair mode,
    // several n-tiles) k_rgb_combine does it below.
    const bool fuse_image = (li + 1 == nl) && have_y && !e->capture && e->cfg.conv_impl == 0 && cl.p.Ntot == cl.p.BN &&
                            cl.p.epi.x_phases == 1 && cl.p.epi.rgb_w != nullptr && !(c.flags & GLASS_FLAG_NO_IMAGE_FUSION);
    if (li + 1 == nl) {
      snprintf(nm, sizeof nm, "g.rgb%d.bias", l.block);
      cl.p.epi.image = fuse_image ? images_out : nullptr;
      cl.p.epi.img_yprev = fuse_image ? ybuf[ycur] : nullptr;
      cl.p.epi.img_bias = fuse_image ? tptr<float>(e, nm) : nullptr;
    }
    RC(run_conv(e, cl, s));
    const __half* layer_out = cl.p.epi.out;
    if (e->g_exact[li]) {
      snprintf(nm, sizeof nm, "u%zu", li);
      RC(capture_f16(e, nm, e->actC, (size_t)P * (l.res + 2) * (l.res + 2) * l.cout, s));
      __half* dst = (cl.p.in == e->actA) ? e->actB : e->actA;
      snprintf(nm, sizeof nm, "g.conv%zu.bias", li);
      const float* bias = tptr<float>(e, nm);
      snprintf(nm, sizeof nm, "g.conv%zu.nstr", li);
      LAUNCH(k_upfir(e->actC, dst, e->noise + e->noise_layer_off[li], e->noise_per_group, c.batch_size,
                     tptr<float>(e, nm), bias, e->styles_n + e->conv_off[li + 1], e->S, P, l.res, l.res, l.cout, s));
      layer_out = dst;
    }
    if (layer_out != nullptr && e->range_check)
      LAUNCH(k_range_scan(layer_out, (size_t)P * l.res * l.res * l.cout, e->range_ctr, s));
    if (layer_out != nullptr) {
      snprintf(nm, sizeof nm, "xs%zu", li);
      const bool i8 = li + 1 < nl && e->g_in_i8[li + 1];
      RC(capture_f16(e, nm, layer_out, (size_t)P * l.res * l.res * l.cout, s, i8 ? l.res : 0, i8 ? l.cout : 0));
    }
    const bool last_in_block = (li + 1 == nl) || (e->glayers[li + 1].block != l.block);
    if (last_in_block && !fuse_image) {
      const bool final_block = (li + 1 == nl);
      snprintf(nm, sizeof nm, "g.rgb%d.bias", l.block);
      const int n_slabs = cl.p.Ntot / cl.p.BN;
      // the last block's skip sum only feeds the image: its fp32 copy is written for debug captures only
      LAUNCH(k_rgb_combine(have_y ? ybuf[ycur] : nullptr, e->slabs, n_slabs, tptr<float>(e, nm),
                           (final_block && !e->capture) ? nullptr : ybuf[ycur ^ 1],
                           final_block ? images_out : nullptr, P, l.res, l.res, s));
      ycur ^= 1;
      have_y = true;
      snprintf(nm, sizeof nm, "rgb%d", l.block);
      RC(capture_f32(e, nm, reinterpret_cast<float*>(ybuf[ycur]), (size_t)P * l.res * l.res * 4, s));
    }
  }
  return GLASS_OK;
}

int run_clip(glass_engine* e, const float* images, int P, float* sim_out, float* neg_sim_out, cudaStream_t s) {
  const glass_config& c = e->cfg;
  if (!e->have_text) return fail(GLASS_ERR_STATE, "glass_set_text_features was not called");
  const int g = c.clip_resolution / c.clip_patch, T = g * g + 1, Wd = c.clip_width;
  char nm[64];
  LAUNCH(k_resize_patches(images, e->patches, P, e->R, c.clip_resolution, c.clip_patch, s));
  size_t ci = 0;
  RC(run_conv(e, e->c_convs[ci++], s));
  LAUNCH(k_embed_lnpre(e->patch_emb, tptr<float>(e, "c.cls"), tptr<float>(e, "c.pos"), tptr<float>(e, "c.lnpre.w"),
                       tptr<float>(e, "c.lnpre.b"), e->tokens, P, T, Wd, s));
  RC(capture_f16(e, "ln_pre", e->tokens, (size_t)P * T * Wd, s));
  for (int l = 0; l < c.clip_layers; ++l) {
    auto nmf = [&](const char* suffix) { snprintf(nm, sizeof nm, "c.l%d.%s", l, suffix); return std::string(nm); };
    LAUNCH(k_layernorm(e->tokens, tptr<float>(e, nmf("ln1.w")), tptr<float>(e, nmf("ln1.b")), e->hbuf, P * T, Wd, s));
    RC(run_conv(e, e->c_convs[ci++], s));   // qkv
    if (c.flags & GLASS_FLAG_SIMT_ATTENTION) LAUNCH(k_attention(e->qkv, e->att, P, T, Wd, s));
    else LAUNCH(k_attention_tc(e->qkv, e->att, P, T, Wd, s));
    RC(run_conv(e, e->c_convs[ci++], s));   // out + residual
    LAUNCH(k_layernorm(e->tokens, tptr<float>(e, nmf("ln2.w")), tptr<float>(e, nmf("ln2.b")), e->hbuf, P * T, Wd, s));
    RC(run_conv(e, e->c_convs[ci++], s));   // fc
    RC(run_conv(e, e->c_convs[ci++], s));   // proj + residual
    snprintf(nm, sizeof nm, "block%d", l);
    RC(capture_f16(e, nm, e->tokens, (size_t)P * T * Wd, s));
  }
  LAUNCH(k_final_cosine(e->tokens, tptr<float>(e, "c.lnpost.w"), tptr<float>(e, "c.lnpost.b"), tptr<float>(e, "c.proj"),
                        e->text, e->features, sim_out, neg_sim_out, P, T, Wd, c.clip_embed_dim, s));
  RC(capture_f32(e, "features", e->features, (size_t)P * c.clip_embed_dim, s));
  return GLASS_OK;
}

int run_discriminator(glass_engine* e, const float* images, int P, float* logits_out, float* hinge_out,
                      cudaStream_t s) {
  const glass_config& c = e->cfg;
  if (!c.use_discriminator) return fail(GLASS_ERR_STATE, "engine was created without a discriminator");
  const int nb = c.num_blocks;
  auto dch = [&](int i) { return e->gch[nb - 1 - i]; };
  char nm[64];
  // fromRGB + the first block's projection FIR in one pass (GLASS_DEBUG_SPLIT_FRGB: the two-kernel route)
  static const bool split_frgb = debug_env("GLASS_DEBUG_SPLIT_FRGB") != nullptr;
  const bool fused_frgb = !split_frgb && dch(0) % 32 == 0;
  // block 0: fromRGB + FIR + projection GEMM in one kernel where the shapes allow it
  const bool frgb_proj = fused_frgb && e->d_proj_fused[0] && k_fir_proj_supported(dch(0), dch(1), true);
  if (frgb_proj)
    LAUNCH(k_from_rgb_fir_proj(images, e->frgb_folded.data(), e->actA, e->d_in_i8[0], tptr<__half>(e, "d.b0.proj.w"), e->dR,
                               P, e->R, dch(0), dch(1), s));
  else if (fused_frgb)
    LAUNCH(k_from_rgb_fir(images, e->frgb_folded.data(), e->actA, e->dXd, P, e->R, dch(0), e->d_in_i8[0], s));
  else
    LAUNCH(k_from_rgb(images, tptr<float>(e, "d.frgb.w"), tptr<float>(e, "d.frgb.b"), e->actA, P, e->R, dch(0),
                      e->d_in_i8[0], s));
  const __half* x = e->actA;
  int res = e->R;
  size_t ci = 0;
  for (int b = 0; b < nb - 1; ++b) {
    // projection path: FIR (+ 1x1 GEMM -> dR when fused; block 0 may already have it from the fromRGB kernel)
    const bool proj_done = (b == 0 && frgb_proj);
    const bool proj_fused = proj_done || (e->d_proj_fused[b] && !(b == 0 && fused_frgb));
    if (!proj_done) {
      if (proj_fused) {
        snprintf(nm, sizeof nm, "d.b%d.proj.w", b);
        LAUNCH(k_fir_proj(x, e->d_in_i8[b], tptr<__half>(e, nm), e->dR, P, res, res, dch(b), dch(b + 1), s));
      } else if (b > 0 || !fused_frgb) {
        LAUNCH(k_fir_down(x, e->dXd, P, res, res, dch(b), e->d_in_i8[b], s));
      }
    }
    RC(run_conv(e, e->d_convs[ci++], s));   // conv0 -> actB (space-to-depth, or NHWC for the exact form)
    if (e->d_exact[b]) LAUNCH(k_blur_s2d(e->actB, e->actC, P, res, res, dch(b), s, (c.flags & GLASS_FLAG_FP32_BLUR) != 0));
    if (proj_fused) {
      ci++;                                 // projection already in dR
    } else if (e->d_c1_proj[b]) {
      // projection = second accumulator of the fused down-conv below: no launch (a zero-length timing entry keeps the
      // per-layer breakdown aligned)
      if (e->timing && e->ev_used + 2 <= e->ev.size()) {
        cudaEventRecord(e->ev[e->ev_used], s);
        cudaEventRecord(e->ev[e->ev_used + 1], s);
        e->ev_conv[e->ev_used / 2] = &e->d_convs[ci];
        e->ev_used += 2;
      }
      ci++;
    } else {
      RC(run_conv(e, e->d_convs[ci++], s));   // projection -> dR
    }
    const ConvLaunch& c1 = e->d_convs[ci++];
    RC(run_conv(e, c1, s));                 // conv1 + residual
    x = c1.p.epi.out;
    res /= 2;
    if (e->range_check) LAUNCH(k_range_scan(x, (size_t)P * res * res * dch(b + 1), e->range_ctr, s));
    snprintf(nm, sizeof nm, "d%d", b);
    const bool i8 = (b + 1 < nb - 1) && e->d_in_i8[b + 1];
    RC(capture_f16(e, nm, x, (size_t)P * res * res * dch(b + 1), s, i8 ? res : 0, i8 ? dch(b + 1) : 0));
  }
  const int C = dch(nb - 1);
  const int cpad = ((C + 1 + 63) / 64) * 64;
  int group = c.mbstd_group_size > 0 ? c.mbstd_group_size : c.batch_size;
  LAUNCH(k_mbstd(x, e->dFin, P, c.batch_size, group, C, cpad, s));
  RC(run_conv(e, e->d_convs[ci++], s));     // final 3x3
  RC(run_conv(e, e->d_convs[ci++], s));     // dense0
  LAUNCH(k_dense1_hinge(e->dDense, tptr<float>(e, "d.dense1.w"), tptr<float>(e, "d.dense1.b"), logits_out, hinge_out, P,
                        C, s));
  return GLASS_OK;
}

int check_pop(glass_engine* e, int pop) {
  if (!e || !e->finalized) return fail(GLASS_ERR_STATE, "engine is not finalized");
  if (pop <= 0 || pop % e->cfg.batch_size != 0)
    return fail(GLASS_ERR_ARG, "population %d is not a positive multiple of batch_size %d (models.py:112,124)", pop,
                e->cfg.batch_size);
  if (pop > e->cfg.max_population)
    return fail(GLASS_ERR_ARG, "population %d exceeds max_population %d", pop, e->cfg.max_population);
  if (pop / e->cfg.batch_size > e->max_groups)
    return fail(GLASS_ERR_ARG, "population %d / batch_size %d needs more noise groups than the %d allocated", pop,
                e->cfg.batch_size, e->max_groups);
  if (e->plan_pop != pop) RC(build_plan(e, pop));
  return GLASS_OK;
}

// The graph body: one evaluation from e->z32 / e->noise to e->neg_sim / e->hinge on stream `s` (in capture).  The
// CLIP tower and the discriminator both depend only on the images, so they are captured as parallel branches
// (CLIP on the side stream): its small GEMMs fill the SMs that the discriminator's kernel tails leave idle.
int evaluate_body(glass_engine* e, int pop, cudaStream_t s) {
  static const bool no_fork = debug_env("GLASS_DEBUG_NO_FORK") != nullptr;
  RC(run_generator(e, e->z32, pop, nullptr, e->images, s, true));
  const bool fork = e->cfg.use_discriminator && !no_fork;
  cudaStream_t cs = fork ? e->side_stream : s;
  if (fork) {
    CUDA_OK(cudaEventRecord(e->ev_fork, s));
    CUDA_OK(cudaStreamWaitEvent(cs, e->ev_fork, 0));
  }
  RC(run_clip(e, e->images, pop, e->sim, e->neg_sim, cs));
  if (e->cfg.use_discriminator) RC(run_discriminator(e, e->images, pop, e->dlogits, e->hinge, s));
  if (fork) {
    CUDA_OK(cudaEventRecord(e->ev_join, cs));
    CUDA_OK(cudaStreamWaitEvent(s, e->ev_join, 0));
  }
  return GLASS_OK;
}

void timing_begin(glass_engine* e) { e->ev_used = 0; }
int timing_end(glass_engine* e, cudaStream_t s) {
  if (!e->timing) return GLASS_OK;
  CUDA_OK(cudaStreamSynchronize(s));
  float total = 0.f;
  for (size_t i = 0; i + 1 < e->ev_used; i += 2) {
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, e->ev[i], e->ev[i + 1]));
    total += ms;
  }
  e->last_conv_ms = total;
  e->last_conv_launches = (int)(e->ev_used / 2);
  return GLASS_OK;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

const char* glass_last_error(void) { return g_last_error.c_str(); }

int glass_create(const glass_config* cfg, glass_engine** out) {
  if (!cfg || !out) return fail(GLASS_ERR_ARG, "null argument");
  if (cfg->num_blocks < 2 || cfg->num_blocks > GLASS_MAX_BLOCKS) return fail(GLASS_ERR_ARG, "num_blocks out of range");
  if (cfg->batch_size <= 0 || cfg->max_population <= 0 || cfg->max_population % cfg->batch_size != 0)
    return fail(GLASS_ERR_ARG, "max_population must be a positive multiple of batch_size");
  for (int i = 0; i < cfg->num_blocks; ++i)
    if (cfg->channels[i] % 32 != 0 || cfg->channels[i] <= 0 || cfg->channels[i] > 512)
      return fail(GLASS_ERR_ARG, "channels must be multiples of 32 in [32,512]");
  if (cfg->latent_size <= 0 || cfg->latent_size > 1024 || cfg->clip_width % 64 != 0)
    return fail(GLASS_ERR_ARG, "unsupported latent_size / clip_width");
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0 || cfg->device >= ndev)
    return fail(GLASS_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                err == cudaSuccess ? "device ordinal out of range" : cudaGetErrorString(err));
  CUDA_OK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return fail(GLASS_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major,
                prop.minor);
  glass_engine* e = new glass_engine();
  e->cfg = *cfg;
  e->num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (err != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    delete e;
    return fail(GLASS_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
  }
  e->encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  derive_arch(e);
  *out = e;
  return GLASS_OK;
}

int glass_set_tensor(glass_engine* e, const char* name, const void* host_data, size_t nbytes) {
  if (!e || !name || !host_data || nbytes == 0) return fail(GLASS_ERR_ARG, "null argument");
  CUDA_OK(cudaSetDevice(e->cfg.device));
  DevTensor& t = e->tensors[name];
  if (t.ptr && t.bytes != nbytes) {
    cudaFree(t.ptr);
    t.ptr = nullptr;
  }
  if (!t.ptr) CUDA_OK(cudaMalloc(&t.ptr, nbytes));
  t.bytes = nbytes;
  CUDA_OK(cudaMemcpy(t.ptr, host_data, nbytes, cudaMemcpyHostToDevice));
  e->plan_pop = -1;
  return GLASS_OK;
}

int glass_finalize(glass_engine* e) {
  if (!e) return fail(GLASS_ERR_ARG, "null engine");
  CUDA_OK(cudaSetDevice(e->cfg.device));
  RC(validate_weights(e));
  Arena probe;
  layout_workspace(e, probe);
  const size_t need = probe.off + 4096;
  void* base = nullptr;
  cudaError_t err = cudaMalloc(&base, need);
  if (err != cudaSuccess)
    return fail(GLASS_ERR_NOMEM, "workspace of %.2f GB for max_population=%d: %s", need / 1e9, e->cfg.max_population,
                cudaGetErrorString(err));
  CUDA_OK(cudaMemset(base, 0, need));
  e->arena.base = (uint8_t*)base;
  e->arena.cap = need;
  e->arena.off = 0;
  layout_workspace(e, e->arena);
  e->ev.resize(512);
  e->ev_conv.assign(256, nullptr);
  for (auto& ev : e->ev) CUDA_OK(cudaEventCreate(&ev));
  CUDA_OK(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&e->side_stream, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  e->finalized = true;
  return GLASS_OK;
}

int glass_set_text_features(glass_engine* e, const float* host_text, int32_t n) {
  if (!e || !e->finalized) return fail(GLASS_ERR_STATE, "engine is not finalized");
  if (n != e->cfg.clip_embed_dim) return fail(GLASS_ERR_ARG, "text feature length %d != embed dim %d", n, e->cfg.clip_embed_dim);
  CUDA_OK(cudaMemcpy(e->text, host_text, (size_t)n * 4, cudaMemcpyHostToDevice));
  e->have_text = true;
  return GLASS_OK;
}

int glass_destroy(glass_engine* e) {
  if (!e) return GLASS_OK;
  cudaSetDevice(e->cfg.device);
  for (auto& kv : e->tensors) cudaFree(kv.second.ptr);
  if (e->arena.base) cudaFree(e->arena.base);
  for (auto& ev : e->ev) cudaEventDestroy(ev);
  drop_graph(e);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  if (e->side_stream) cudaStreamDestroy(e->side_stream);
  delete e;
  return GLASS_OK;
}

int glass_generate(glass_engine* e, const float* z_dev, int32_t pop, const glass_noise* noise, float* images_dev,
                   void* stream) {
  RC(check_pop(e, pop));
  cudaStream_t s = (cudaStream_t)stream;
  timing_begin(e);
  RC(run_generator(e, z_dev, pop, noise, images_dev ? images_dev : e->images, s));
  return timing_end(e, s);
}

int glass_clip_similarity(glass_engine* e, const float* images_dev, int32_t pop, float* sim_dev, void* stream) {
  RC(check_pop(e, pop));
  cudaStream_t s = (cudaStream_t)stream;
  timing_begin(e);
  RC(run_clip(e, images_dev, pop, sim_dev, nullptr, s));
  return timing_end(e, s);
}

int glass_discriminate(glass_engine* e, const float* images_dev, int32_t pop, float* logits_dev, void* stream) {
  RC(check_pop(e, pop));
  cudaStream_t s = (cudaStream_t)stream;
  timing_begin(e);
  RC(run_discriminator(e, images_dev, pop, logits_dev, nullptr, s));
  return timing_end(e, s);
}

int glass_evaluate_device(glass_engine* e, const float* z_dev, int32_t pop, const glass_noise* noise,
                          float* neg_sim_dev, float* hinge_dev, void* stream) {
  RC(check_pop(e, pop));
  if (e->cfg.use_discriminator && hinge_dev == nullptr) return fail(GLASS_ERR_ARG, "hinge output is required");
  cudaStream_t s = (cudaStream_t)stream;
  static const bool no_graph_env = debug_env("GLASS_DEBUG_NO_GRAPH") != nullptr;
  const bool use_graph = !no_graph_env && !(e->cfg.flags & GLASS_FLAG_NO_GRAPH) && e->cfg.conv_impl == 0 &&
                         !e->timing && !e->capture && !e->range_check;
  e->last_eval_pop = pop;
  if (!use_graph || e->graph_evals++ == 0) {
    timing_begin(e);
    RC(run_generator(e, z_dev, pop, noise, e->images, s));
    RC(run_clip(e, e->images, pop, e->sim, neg_sim_dev, s));
    if (e->cfg.use_discriminator) RC(run_discriminator(e, e->images, pop, e->dlogits, hinge_dev, s));
    return timing_end(e, s);
  }
  // per-call inputs go to the engine's fixed buffers on the caller's stream; the graph reads only those
  RC(fill_noise(e, pop, noise, s));
  if (z_dev != e->z32)
    CUDA_OK(cudaMemcpyAsync(e->z32, z_dev, (size_t)pop * e->cfg.latent_size * 4, cudaMemcpyDeviceToDevice, s));
  if (e->graph_exec == nullptr) {
    const int64_t before = e->launches;
    cudaGraph_t graph = nullptr;
    CUDA_OK(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
    int rc = evaluate_body(e, pop, e->cap_stream);
    cudaError_t err = cudaStreamEndCapture(e->cap_stream, &graph);
    e->graph_launches = e->launches - before;
    e->launches = before;
    if (rc != GLASS_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (err != cudaSuccess) return fail(GLASS_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(err));
    err = cudaGraphInstantiate(&e->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (err != cudaSuccess) {
      e->graph_exec = nullptr;
      return fail(GLASS_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(err));
    }
  }
  CUDA_OK(cudaGraphLaunch(e->graph_exec, s));
  e->launches += e->graph_launches;
  if (neg_sim_dev != e->neg_sim)
    CUDA_OK(cudaMemcpyAsync(neg_sim_dev, e->neg_sim, (size_t)pop * 4, cudaMemcpyDeviceToDevice, s));
  if (e->cfg.use_discriminator && hinge_dev != e->hinge)
    CUDA_OK(cudaMemcpyAsync(hinge_dev, e->hinge, (size_t)pop * 4, cudaMemcpyDeviceToDevice, s));
  return GLASS_OK;
}

int glass_evaluate_host(glass_engine* e, const double* x, int32_t pop, const glass_noise* noise, float* neg_sim,
                        float* hinge, void* stream) {
  RC(check_pop(e, pop));
  if (!x || !neg_sim) return fail(GLASS_ERR_ARG, "null argument");
  if (e->cfg.use_discriminator && hinge == nullptr) return fail(GLASS_ERR_ARG, "hinge output is required");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)pop * e->cfg.latent_size;
  CUDA_OK(cudaMemcpyAsync(e->x_host_stage, x, n * 8, cudaMemcpyHostToDevice, s));
  LAUNCH(k_latents_to_f32(e->x_host_stage, e->z32, n, s));
  RC(glass_evaluate_device(e, e->z32, pop, noise, e->neg_sim, e->cfg.use_discriminator ? e->hinge : nullptr, stream));
  CUDA_OK(cudaMemcpyAsync(neg_sim, e->neg_sim, (size_t)pop * 4, cudaMemcpyDeviceToHost, s));
  if (e->cfg.use_discriminator)
    CUDA_OK(cudaMemcpyAsync(hinge, e->hinge, (size_t)pop * 4, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return GLASS_OK;
}

int glass_set_batch_size(glass_engine* e, int32_t batch_size) {
  if (!e || !e->finalized) return fail(GLASS_ERR_STATE, "engine is not finalized");
  if (batch_size <= 0) return fail(GLASS_ERR_ARG, "batch_size must be positive");
  if (batch_size != e->cfg.batch_size) {
    e->cfg.batch_size = batch_size;
    e->plan_pop = -1;
  }
  return GLASS_OK;
}

int64_t glass_launch_count(const glass_engine* e) { return e ? e->launches : 0; }

int glass_set_debug(glass_engine* e, int32_t capture, int32_t timing) {
  if (!e) return fail(GLASS_ERR_ARG, "null engine");
  e->capture = capture != 0;
  e->timing = timing != 0;
  if (!e->capture) e->captured.clear();
  return GLASS_OK;
}

int glass_debug_build(void) { return kDebugBuild ? 1 : 0; }

int glass_last_images_gather(glass_engine* e, const int32_t* rows_host, int32_t n, float* out_dev, void* stream) {
  if (!e || !e->finalized) return fail(GLASS_ERR_STATE, "engine is not finalized");
  if (!rows_host || !out_dev || n <= 0 || n > e->cfg.max_population) return fail(GLASS_ERR_ARG, "bad argument");
  if (e->last_eval_pop <= 0) return fail(GLASS_ERR_STATE, "no fused evaluation has produced images yet");
  for (int i = 0; i < n; ++i)
    if (rows_host[i] < 0 || rows_host[i] >= e->last_eval_pop)
      return fail(GLASS_ERR_ARG, "row %d is outside the last evaluation's population of %d", rows_host[i], e->last_eval_pop);
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_OK(cudaMemcpyAsync(e->gather_rows, rows_host, (size_t)n * 4, cudaMemcpyHostToDevice, s));
  LAUNCH(k_gather_images(e->images, e->gather_rows, n, (size_t)3 * e->R * e->R, out_dev, s));
  CUDA_OK(cudaStreamSynchronize(s));      // rows_host is the caller's
  return GLASS_OK;
}

int glass_image_grid_u8(glass_engine* e, const float* images_dev, int32_t n, int32_t resolution, int32_t nrow,
                        int32_t padding, uint8_t* out_dev, void* stream) {
  if (!images_dev || !out_dev || n <= 0 || resolution <= 0 || nrow <= 0 || padding < 0)
    return fail(GLASS_ERR_ARG, "bad argument");
  cudaError_t err = k_image_grid_u8(images_dev, n, resolution, nrow, padding, out_dev, (cudaStream_t)stream);
  if (e) e->launches++;
  if (err != cudaSuccess) return fail(GLASS_ERR_CUDA, "k_image_grid_u8 failed: %s", cudaGetErrorString(err));
  return GLASS_OK;
}

int glass_biggan_latent(const double* x_host, int32_t pop, int32_t dim_z, int32_t num_classes, float* z_dev,
                        float* cls_dev, void* stream) {
  if (!x_host || !z_dev || !cls_dev || pop <= 0 || dim_z <= 0 || num_classes <= 0)
    return fail(GLASS_ERR_ARG, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  double* stage = nullptr;
  const size_t bytes = (size_t)pop * (dim_z + num_classes) * 8;
  CUDA_OK(cudaMallocAsync((void**)&stage, bytes, s));
  CUDA_OK(cudaMemcpyAsync(stage, x_host, bytes, cudaMemcpyHostToDevice, s));
  cudaError_t err = k_biggan_latent(stage, z_dev, cls_dev, pop, dim_z, num_classes, s);
  cudaFreeAsync(stage, s);
  if (err != cudaSuccess) return fail(GLASS_ERR_CUDA, "k_biggan_latent failed: %s", cudaGetErrorString(err));
  CUDA_OK(cudaStreamSynchronize(s));      // x_host is the caller's
  return GLASS_OK;
}

int glass_set_range_check(glass_engine* e, int32_t enable) {
  if (!e || !e->finalized) return fail(GLASS_ERR_STATE, "engine is not finalized");
  CUDA_OK(cudaSetDevice(e->cfg.device));
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemset(e->range_ctr, 0, 4 * 8));
  e->range_check = enable != 0;
  return GLASS_OK;
}

int glass_range_report(glass_engine* e, int64_t* nonfinite, int64_t* saturated, float* max_abs) {
  if (!e || !e->finalized) return fail(GLASS_ERR_STATE, "engine is not finalized");
  CUDA_OK(cudaSetDevice(e->cfg.device));
  CUDA_OK(cudaDeviceSynchronize());
  unsigned long long h[4];
  CUDA_OK(cudaMemcpy(h, e->range_ctr, sizeof h, cudaMemcpyDeviceToHost));
  if (nonfinite) *nonfinite = (int64_t)h[0];
  if (saturated) *saturated = (int64_t)h[1];
  if (max_abs) { const uint32_t bits = (uint32_t)h[2]; memcpy(max_abs, &bits, 4); }
  return GLASS_OK;
}

int64_t glass_debug_read(glass_engine* e, const char* name, float* host_out, int64_t capacity) {
  if (!e || !name) return fail(GLASS_ERR_ARG, "null argument");
  auto it = e->captured.find(name);
  if (it == e->captured.end()) return fail(GLASS_ERR_ARG, "no captured tensor named '%s'", name);
  const int64_t n = (int64_t)it->second.size();
  if (host_out != nullptr) {
    if (capacity < n) return fail(GLASS_ERR_ARG, "capacity %lld < %lld", (long long)capacity, (long long)n);
    memcpy(host_out, it->second.data(), (size_t)n * 4);
  }
  return n;
}

int glass_last_conv_time(const glass_engine* e, float* ms, int32_t* launches) {
  if (!e) return fail(GLASS_ERR_ARG, "null engine");
  if (ms) *ms = e->last_conv_ms;
  if (launches) *launches = e->last_conv_launches;
  return GLASS_OK;
}

// Per-launch breakdown of the last timed call: fills up to `cap` entries of (ms, algorithmic flops).
int glass_conv_breakdown(glass_engine* e, float* ms, double* flops, int32_t cap) {
  if (!e) return fail(GLASS_ERR_ARG, "null engine");
  std::vector<const ConvLaunch*> all;
  for (auto& c : e->g_convs) all.push_back(&c);
  for (auto& c : e->c_convs) all.push_back(&c);
  for (auto& c : e->d_convs) all.push_back(&c);
  // one entry per PLANNED conv/GEMM, in plan order; launches that the last call did not make (a projection that
  // ran inside k_fir_proj) report 0 ms
  int n = 0;
  for (; n < cap && (size_t)n < all.size(); ++n) {
    ms[n] = 0.f;
    flops[n] = all[n]->flops;
    for (size_t i = 0; i + 1 < e->ev_used; i += 2)
      if (e->ev_conv[i / 2] == all[n]) {
        float t = 0.f;
        cudaEventElapsedTime(&t, e->ev[i], e->ev[i + 1]);
        ms[n] += t;
      }
  }
  return n;
}

}  // extern "C"
