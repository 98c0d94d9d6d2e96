// HBM-bound helper kernels of the fitness path (declarations; kernels.cu).
#pragma once
#include "common.cuh"

namespace glass {

// latent.py:38  f64 -> f32
cudaError_t k_latents_to_f32(const double* x, float* z, size_t n, cudaStream_t s);
// stylegan2/models.py:625-626 pixel norm
cudaError_t k_pixelnorm(const float* z, float* out, int P, int L, cudaStream_t s);
// out[b,n] = f( sum_k g(in[b*in_stride + k]) * Wt[k*N + n] + bias[n] )
//   mode 0: linear; 1: lrelu(0.2)*sqrt2; 2: g = square, f = rsqrt(. + 1e-8) (demodulation, modules.py:945-954)
cudaError_t k_vecmat(const float* in, int in_stride, const float* Wt, const float* bias, float* out, int out_stride,
                     int P, int K, int N, int mode, cudaStream_t s);
// the same for up to kMaxVecmatJobs independent (in, Wt, out) triples in one launch; out row stride = N, no bias
constexpr int kMaxVecmatJobs = 32;
// post_mul (optional): per-sample scalar, out[b,n] *= post_mul[b*post_stride]
struct VecmatJob { const float* in; const float* Wt; float* out; int K, N; const float* post_mul; int post_stride; };
struct VecmatBatch { VecmatJob job[kMaxVecmatJobs]; int n; };
cudaError_t k_vecmat_batched(const VecmatBatch& jobs, int in_stride, int P, int mode, cudaStream_t s);
// fp16 range safety (stylegan2/modules.py:936-958 runs in fp32): every modulated conv's style slice
// s[b, off : off+cin] is divided by m[b,layer] = the power of two >= max|s| before it pre-scales the fp16 activation,
// and the layer's demodulation coefficient is multiplied by m (k_vecmat_batched post_mul).  Demodulation cancels a
// per-sample scalar exactly, and a power of two changes no rounding, so results are bit-identical to the
// un-normalised algebra wherever that one stayed finite.
constexpr int kMaxStyleSlices = 32;
struct StyleSlices { int off[kMaxStyleSlices]; int cin[kMaxStyleSlices]; int n; };
cudaError_t k_style_norm(const float* styles, float* styles_n, float* mscale, int S, int P, const StyleSlices& sl,
                         cudaStream_t s);
// debug: counts non-finite and saturated (|x| >= 65504) fp16 values and tracks max|x| (as float bits) in ctr[0..2]
cudaError_t k_range_scan(const __half* x, size_t n, unsigned long long* ctr, cudaStream_t s);
// x0[b][pix][c] = fp16(const[pix][c] * s[b*stride + c])   (models.py:987 + pre-scale by the first style)
cudaError_t k_const_input(const float* cst, const float* styles, int stride, __half* out, int P, int C, cudaStream_t s);
// wr[b][c][o] = W[c][o] * styles[b*stride + off + o]   (toRGB modulated 1x1 weights, no demod; models.py:848-871)
cudaError_t k_rgb_weights(const float* W, const float* styles, int stride, float* out, int P, int C, cudaStream_t s);
// y[b][Y][X] = up2(yprev)[Y][X] + sum_slabs t + bias   (modules.py:580-602 polyphase; models.py:1004-1013)
// final: also writes image NCHW fp32 = clip((y+1)/2, 0, 1)  (utils.py:14-17)
cudaError_t k_rgb_combine(const float4* yprev, const float4* slabs, int n_slabs, const float* bias, float4* yout,
                          float* image, int P, int H, int W, cudaStream_t s);
// noise: Philox4x32-10 + Box-Muller, n floats
cudaError_t k_noise(float* out, size_t n, uint64_t seed, uint64_t offset, cudaStream_t s);

// Exact up-conv, second half (stylegan2/modules.py:1131-1132 FilterLayer + :414-453 noise + :276-297 bias/act):
// u [P][2H+2][2W+2][C] fp16 (transposed-conv output, demodulated) -> FIR [1,3,3,1]x[1,3,3,1]/16 (pad 1)
// -> + strength*noise + bias -> lrelu*sqrt2 -> * next style -> fp16 NHWC [P][2H][2W][C]
cudaError_t k_upfir(const __half* u, __half* out, const float* noise, size_t noise_group_stride, int noise_group_div,
                    const float* noise_strength, const float* bias, const float* out_scale, int out_scale_stride,
                    int P, int Hout, int Wout, int C, cudaStream_t s);
// Exact down-conv, first half (modules.py:1243-1246 FilterLayer pad 2): a [P][H][W][C] -> blurred (H+1)x(W+1),
// written space-to-depth as [P][H/2+1][W/2+1][4C] (phase-major channels), zeros beyond row/col H
// fp32_variant != 0: the streaming fp32-cascade kernel (cross-check); default: the shared-memory-tiled half2 kernel
cudaError_t k_blur_s2d(const __half* a, __half* out, int P, int H, int W, int C, cudaStream_t s, int fp32_variant = 0);

// ---- image output path (run.py:29-51 save_callback -> generator.py:63-68 -> utils.py:5-7) ----
// torchvision.utils.make_grid (xmaps = min(nrow, n) images per row, `padding` zero pixels around each) fused with
// save_image's mul(255).add(0.5).clamp(0,255).to(uint8) and the CHW -> HWC permute: images [n,3,R,R] fp32 in [0,1]
// -> out [Hg][Wg][3] uint8 with Hg = (R+padding)*ceil(n/xmaps)+padding, Wg = (R+padding)*xmaps+padding.
cudaError_t k_image_grid_u8(const float* images, int n, int R, int nrow, int padding, uint8_t* out, cudaStream_t s);
// out[i] = images[rows[i]] for whole [3,R,R] fp32 images (rows on the device)
cudaError_t k_gather_images(const float* images, const int* rows, int n, size_t image_elems, float* out, cudaStream_t s);
// ---- BigGAN latent arithmetic (latent.py:16-24): x f64 [P, dz + ncls] -> z = clip(x[:, :dz], -2, 2) fp32,
// cls = softmax(x[:, dz:]) fp32 over the ncls "bool" genes ----
cudaError_t k_biggan_latent(const double* x, float* z, float* cls, int P, int dz, int ncls, cudaStream_t s);

// ---- CLIP tower ----
// generator.py:45 (bilinear 1024->224, align_corners=False) fused with the im2col of
// clip/model.py:219 conv1 (k=32,s=32): patches[b*g*g + gy*g+gx][c*p*p + py*p + px] fp16
cudaError_t k_resize_patches(const float* images, __half* patches, int P, int Rin, int Rout, int patch, cudaStream_t s);
// clip/model.py:222-224: cat cls, + pos (fp16 adds), ln_pre -> tokens fp16 [P*T][W]
cudaError_t k_embed_lnpre(const __half* patch_emb, const float* cls, const float* pos, const float* lw, const float* lb,
                          __half* tokens, int P, int T, int W, cudaStream_t s);
// clip/model.py:152-158
cudaError_t k_layernorm(const __half* x, const float* w, const float* b, __half* out, int M, int W, cudaStream_t s);
// nn.MultiheadAttention core (clip/model.py:180-182): qkv [P*T][3W] -> out [P*T][W]; heads of 64
// k_attention_tc (attention_tc.cu): QK^T and PV as tcgen05.mma with the softmax between them out of TMEM -- the
// product path; k_attention is the SIMT bring-up / cross-check version (GLASS_FLAG_SIMT_ATTENTION)
cudaError_t k_attention(const __half* qkv, __half* out, int P, int T, int W, cudaStream_t s);
cudaError_t k_attention_tc(const __half* qkv, __half* out, int P, int T, int W, cudaStream_t s);
// clip/model.py:230-233 + generator.py:51: ln_post(cls) @ proj -> features; cosine vs text
cudaError_t k_final_cosine(const __half* tokens, const float* lw, const float* lb, const float* proj,
                           const float* text, float* features, float* sim, float* neg_sim, int P, int T, int W, int E,
                           cudaStream_t s);

// ---- discriminator ----
// utils.py:19-21 denorm + models.py:1121-1144 fromRGB 1x1 + bias + lrelu*sqrt2 -> NHWC fp16
// out_i8: write the channel-group-interleaved layout [P][R][C/8][R][8] instead of NHWC
cudaError_t k_from_rgb(const float* images, const float* Wt, const float* bias, __half* out, int P, int R, int C,
                       int out_i8, cudaStream_t s);
// k_from_rgb fused with k_fir_down of the first block: writes x AND its FIR/stride-2 copy.  folded_host: HOST array
// [4][C] = {2 sqrt2 W_r, 2 sqrt2 W_g, 2 sqrt2 W_b, sqrt2 (bias - sum W)} (passed to the kernel by value: the weights
// are constant-bank operands of the FFMAs)
cudaError_t k_from_rgb_fir(const float* images, const float* folded_host, __half* xout, __half* down, int P, int R,
                           int C, int out_i8, cudaStream_t s);
// projection path FIR (pad 1) sampled at stride 2 (modules.py:1204-1220, 1243-1246): [N,H,W,C] -> [N,H/2,W/2,C]
// in_i8: x is channel-group-interleaved [N][H][C/8][W][8]; the output is always NHWC
cudaError_t k_fir_down(const __half* x, __half* out, int N, int H, int W, int C, int in_i8, cudaStream_t s);
// The same FIR fused with the block's 1x1 stride-2 projection (modules.py:1587-1601) on the tensor cores
// (fir_proj_tc.cu): x -> dR [N][H/2][W/2][Co]; wproj is [Co][C] fp16.  Co in {64,128,256}.
bool k_fir_proj_supported(int C, int Co, bool from_rgb);
cudaError_t k_fir_proj(const __half* x, int in_i8, const __half* wproj, __half* out, int N, int H, int W, int C, int Co,
                       cudaStream_t s);
// ... and with fromRGB in front (the first block): image -> x (xout) and dR
cudaError_t k_from_rgb_fir_proj(const float* images, const float* folded_host, __half* xout, int out_i8,
                                const __half* wproj, __half* out, int P, int R, int C, int Co, cudaStream_t s);
// D down-conv, exact form with the FIR inside the kernel (downconv_tc.cu): a = conv0 output in the I8 layout
// [N][2Ho][C/8][2Wo][8] (map_a: dims (2Wo*8, C/8, 2Ho, N), box (160, 4, 36, 1), un-swizzled), w9 = [9][Cout][C] fp16
// (map_w: box (C, 64, 1), swizzle C*2 bytes) -> out = (lrelu(conv3x3_s2(FIR(a)) + bias)*sqrt2 + residual) * post_scale
bool k_downconv_fused_supported(int C, int Cout, int Ho, int Wo);
void k_downconv_fused_geometry(int C, int Cout, int* box_groups, int* box_cols);
bool k_downconv_proj_supported(int C, int Cout);
cudaError_t k_downconv_fused(const CUtensorMap& map_a, const CUtensorMap& map_w, const CUtensorMap* map_xd,
                             const CUtensorMap* map_wp, int C, int N, int Ho, int Wo, int Cout, const float* bias,
                             const __half* residual, int res_i8, __half* out, int out_i8, float post_scale, int num_sms,
                             cudaStream_t s);
// modules.py:701-747 incl. the in-place centring; x [P,16,C] -> out [P,16,Cpad] (channel C = std feature)
cudaError_t k_mbstd(const __half* x, __half* out, int P, int batch, int group, int C, int Cpad, cudaStream_t s);
// models.py:1224-1225 last dense + problem.py:23 hinge
cudaError_t k_dense1_hinge(const __half* x, const float* w, const float* b, float* logits, float* hinge, int P, int C,
                           cudaStream_t s);

}  // namespace glass
