/*
 * clipglass_b200.h — C ABI of the B200-native CLIP-GLaSS fitness-evaluation path.
 *
 * Drop-in boundary.  The reference has no FFI: the path is the
 * Python call chain
 *
 *     problem.py:14-29    GenerationProblem._evaluate(x, out)
 *       latent.py:37-38   StyleGAN2LatentSpace.set_from_population   (f64 -> f32, H2D)
 *       generator.py:29-34  Generator.generate(ls, minibatch)        (models.py:108-118 loop over G)
 *       generator.py:43-51  Generator.clip_similarity(generated)     (kornia.resize -> CLIP.encode_image -> cosine)
 *       generator.py:36-38  Generator.discriminate(generated, minibatch) (models.py:120-130 loop over D)
 *
 * Each entry point below replaces one of those calls and is what a ctypes stub
 * in the reference's generator.py / problem.py binds.
 * Plain pointers and sizes only; no torch types.  All functions return 0 on
 * success or a negative glass_status; glass_last_error() gives the message.
 * There is no CPU fallback: every compute entry point fails with
 * GLASS_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef CLIPGLASS_B200_H_
#define CLIPGLASS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct glass_engine glass_engine;   /* opaque; owns device weights + workspace */

typedef enum glass_status {
  GLASS_OK = 0,
  GLASS_ERR_ARG = -1,       /* bad argument (mirrors the reference's AssertionError, models.py:112,124) */
  GLASS_ERR_CUDA = -2,      /* CUDA runtime / driver failure, or no sm_100 device */
  GLASS_ERR_STATE = -3,     /* call order (weights missing, not finalized, ...) */
  GLASS_ERR_NOMEM = -4
} glass_status;

#define GLASS_MAX_BLOCKS 12

/* Architecture + run shape.  Mirrors what the reference spreads over
 * config.py:74-94 (StyleGAN2_ffhq_d), the G/D pickles and clip/model.py:363-392. */
typedef struct glass_config {
  int32_t num_blocks;                     /* resolutions 4 .. 4*2^(num_blocks-1) */
  int32_t channels[GLASS_MAX_BLOCKS];     /* per block, 4x4 first (G order)      */
  int32_t latent_size;                    /* config.dim_z = 512                  */
  int32_t mapping_layers;                 /* 8                                   */
  int32_t batch_size;                     /* config.batch_size: noise scope (modules.py:426-452)
                                             and MinibatchStd scope (modules.py:726) */
  int32_t mbstd_group_size;               /* 4 (models.py:1047)                  */
  int32_t use_discriminator;              /* config.use_discriminator            */
  int32_t clip_width, clip_layers, clip_patch, clip_resolution, clip_embed_dim;
  int32_t max_population;                 /* workspace is sized for this many candidates per call */
  int32_t device;                         /* CUDA ordinal                        */
  int32_t conv_impl;                      /* 0 = tcgen05 tensor-core path (product);
                                             1 = SIMT bring-up kernels (same epilogues; tests only) */
  int32_t flags;                          /* GLASS_FLAG_* */
} glass_config;

/* Up/down convs exist in two forms: FIR-folded 3x3 (4x the MACs, everything in one tensor-core launch) and
 * exact polyphase (2x2-tap conv + one streaming FIR pass).  By default the engine picks per layer with a
 * measured cost model (exact where the layer is tensor-bound: Cin >= 512 at >= 64x64 in G, Cin >= 128 in D). */
#define GLASS_FLAG_FOLDED_RESAMPLE 1   /* folded form everywhere */
#define GLASS_FLAG_EXACT_RESAMPLE 2    /* exact form wherever it is defined (inputs >= 16x16) */
/* The 32-channel 3x3 convs normally run on horizontally paired pixels ([H][W/2][64] view of the same bytes:
 * 128-byte TMA rows, 256-pixel tiles, 2x the MACs of a layer that is not math-bound). */
#define GLASS_FLAG_NO_PAIR_PACK 4      /* keep them on single pixels */
/* Activations that feed a 32/64-channel 3x3 conv are normally stored channel-group-interleaved
 * ([N][H][C/8][W][8]) so that one un-swizzled haloed TMA box per tile serves all nine taps (conv_tc MODE 4). */
#define GLASS_FLAG_NO_I8_LAYOUT 8      /* keep every activation NHWC (MODE 1 / pixel pairs for those layers) */
/* The ViT attention core (QK^T, softmax, PV) runs as tcgen05.mma out of TMEM (attention_tc.cu). */
#define GLASS_FLAG_SIMT_ATTENTION 16   /* use the scalar shared-memory kernel instead (cross-check) */
/* glass_evaluate_device / glass_evaluate_host replay one CUDA graph per launch plan (built on the second evaluation;
 * CLIP and the discriminator as parallel branches).  The facade calls always launch eagerly. */
#define GLASS_FLAG_NO_GRAPH 32         /* launch every kernel eagerly */
/* Opt-in: the projection path of the D blocks with 64/128/256 output channels (FIR + 1x1 stride-2 conv) as one kernel
 * (fir_proj_tc.cu: FIR in fp32 -> fp16 A operand in shared memory -> tcgen05 GEMM).  Measured SLOWER than k_fir_down
 * + a 1x1 conv_tc launch at P=64 (2.86 vs 2.55 ms on the 1024^2 block: its serial stage/FIR/MMA/store phases leave
 * three resident blocks per SM idle too often), so it is off by default and kept as a cross-checked variant. */
#define GLASS_FLAG_PROJ_FUSION 64
/* Cross-check variant: keep the space-to-depth tensor between conv0 and the folded conv1 of the D 1024^2 block NHWC
 * (conv_tc MODE 0 instead of the I8 layout + MODE 6 streamed taps).  Same products, different accumulation order. */
#define GLASS_FLAG_C1_NHWC 128
/* The exact polyphase weight tensors have 9 non-zero (tap, phase) blocks of 16; the tensor-core path skips the zero
 * blocks (no TMA load, no MMA).  Cross-check variant: multiply them like any other block. */
#define GLASS_FLAG_NO_ZERO_SKIP 256
/* The D down-convs of the 32/64-channel blocks (1024^2 -> 512^2 -> 256^2) normally run in their exact form with the
 * FIR applied inside the kernel (downconv_tc.cu).  Cross-check variant: the FIR-folded 3x3 over the space-to-depth
 * tensor (conv_tc MODE 6 / MODE 0), 4x the MACs. */
#define GLASS_FLAG_NO_FUSED_DOWN 512
/* The last generator conv normally finishes the image in its own epilogue (skip sum + x2 upsample + toRGB bias +
 * biggan_norm, what k_rgb_combine does for the other blocks).  Cross-check variant: write the toRGB slab and run
 * k_rgb_combine for the last block too. */
#define GLASS_FLAG_NO_IMAGE_FUSION 1024
/* The blur pass of the exact down-convs (128..512-channel D blocks) normally runs shared-memory-tiled in packed-half2
 * arithmetic.  Cross-check variant: the streaming kernel with an fp32 cascade. */
#define GLASS_FLAG_FP32_BLUR 2048
/* Keep the projection path of the 32 -> 64 discriminator block as its own 1x1 GEMM launch (default: a second
 * accumulator of that block's fused-FIR down-conv kernel; cross-check / A-B variant). */
#define GLASS_FLAG_NO_PROJ_ACC 4096

/* -- lifetime ------------------------------------------------------------- */
/* Replaces Generator.__init__ (generator.py:12-27): allocate the engine. */
int glass_create(const glass_config* cfg, glass_engine** out);
/* Upload one packed weight tensor from HOST memory (names and layouts are
 * documented in clip_glass_b200/packing.py; produced from the reference's
 * G.pth / D.pth / ViT-B-32.pt state_dict layouts). */
int glass_set_tensor(glass_engine* e, const char* name, const void* host_data, size_t nbytes);
/* After all tensors are set: build TMA descriptors and the launch plan. */
int glass_finalize(glass_engine* e);
/* generator.py:23-24 caches text_features [1,E]; fp32 here. */
int glass_set_text_features(glass_engine* e, const float* host_text, int32_t n);
int glass_destroy(glass_engine* e);
/* Change the minibatch size used for the noise / MinibatchStd scope of later
 * calls (`minibatch` argument of Generator.generate / discriminate,
 * generator.py:29-38; minibatch=None in the reference == batch_size = pop).
 * pop / batch_size may not exceed max_population / creation batch_size. */
int glass_set_batch_size(glass_engine* e, int32_t batch_size);

/* -- the hot path ---------------------------------------------------------- */
/* Noise for NoiseInjectionWrapper (modules.py:414-453).  Either explicit
 * tensors — `noise` points to HOST or DEVICE fp32 laid out as
 * [n_groups][sum over noise layers of H*W] (groups = population/batch_size,
 * layers in forward order) — or, when `noise` is NULL, drawn on the device
 * from `seed` (Philox4x32-10 + Box-Muller), one independent draw per group,
 * which is what "fresh normal_() per minibatch forward" means. */
typedef struct glass_noise {
  const float* noise;       /* NULL => use seed */
  int32_t noise_on_device;  /* 1 if `noise` is a device pointer */
  uint64_t seed;
  uint64_t first_group;     /* global index of this call's first minibatch group: the seeded stream is
                               indexed by (seed, global group, element), so a population shard evaluated
                               on another rank draws exactly the noise the single-GPU run would */
} glass_noise;

/* problem.py:14-29.  x: HOST float64 [pop, latent_size] exactly as pymoo
 * passes it.  Outputs (HOST): neg_sim[pop] (= -sim, fp32; the reference's
 * value is fp16-rounded), hinge[pop] (relu(1-D), only if the engine was
 * created with use_discriminator; may be NULL otherwise).  pop must be a
 * multiple of batch_size (models.py:112,124).  `stream` is a cudaStream_t
 * (0 = default stream).  Synchronous: returns when outputs are on the host. */
int glass_evaluate_host(glass_engine* e, const double* x, int32_t pop,
                        const glass_noise* noise, float* neg_sim, float* hinge,
                        void* stream);

/* Same, inputs/outputs resident in device memory (z fp32 [pop, latent]);
 * asynchronous on `stream`. */
int glass_evaluate_device(glass_engine* e, const float* z_dev, int32_t pop,
                          const glass_noise* noise, float* neg_sim_dev, float* hinge_dev,
                          void* stream);

/* Generator.generate (generator.py:29-34): latents -> images in [0,1],
 * fp32 NCHW [pop,3,R,R] written to DEVICE memory `images_dev`. */
int glass_generate(glass_engine* e, const float* z_dev, int32_t pop,
                   const glass_noise* noise, float* images_dev, void* stream);
/* Generator.clip_similarity (generator.py:43-51), txt2img branch:
 * images fp32 NCHW in [0,1] (device) -> sim[pop] fp32 (device). */
int glass_clip_similarity(glass_engine* e, const float* images_dev, int32_t pop,
                          float* sim_dev, void* stream);
/* Generator.discriminate (generator.py:36-38): images in [0,1] (device) ->
 * D logits [pop] fp32 (device); denorm x*2-1 is applied inside. */
int glass_discriminate(glass_engine* e, const float* images_dev, int32_t pop,
                       float* logits_dev, void* stream);

/* -- image output path ----------------------------------- */
/* run.py:29-51 save_callback calls generator.generate on candidates the last _evaluate already rendered.  Copies the
 * images of rows `rows_host[0..n)` of the LAST glass_evaluate_host / glass_evaluate_device call (fp32 NCHW in [0,1],
 * the images that were scored) into DEVICE memory out_dev [n,3,R,R].  GLASS_ERR_STATE before the first evaluation. */
int glass_last_images_gather(glass_engine* e, const int32_t* rows_host, int32_t n, float* out_dev, void* stream);
/* utils.py:5-7 save_grid = torchvision make_grid (nrow images per row, `padding` zero pixels) + save_image's
 * mul(255).add(0.5).clamp(0,255).to(uint8): images_dev [n,3,R,R] fp32 -> out_dev uint8 HWC
 * [(R+padding)*ceil(n/min(n,nrow))+padding][(R+padding)*min(n,nrow)+padding][3], ready for the JPEG encoder.
 * `e` may be NULL (it only counts the launch). */
int glass_image_grid_u8(glass_engine* e, const float* images_dev, int32_t n, int32_t resolution, int32_t nrow,
                        int32_t padding, uint8_t* out_dev, void* stream);

/* -- BigGAN latent arithmetic (latent.py:16-24) ------------ */
/* x: HOST float64 [pop, dim_z + num_classes] as pymoo passes the mixed real/bool population.  z_dev [pop, dim_z] =
 * clip(x[:, :dim_z], -2, 2); cls_dev [pop, num_classes] = softmax(x[:, dim_z:], dim=1) (fp32, DEVICE).  The generator
 * itself (pytorch_pretrained_biggan 0.1.1) is not vendored by the reference and is not built here. */
int glass_biggan_latent(const double* x_host, int32_t pop, int32_t dim_z, int32_t num_classes, float* z_dev,
                        float* cls_dev, void* stream);

/* -- img2txt path (BASELINE config 5) ---------------------- */
/* GPT-2 greedy decode + CLIP text tower (clip_glass_b200/csrc/text_engine.cu).  A separate engine: it shares no
 * weights with the StyleGAN2 path.  Either half may be left out (layers = 0). */
typedef struct glass_text_engine glass_text_engine;
typedef struct glass_text_config {
  /* gpt2/config.py:6-24 */
  int32_t gpt2_vocab, gpt2_positions, gpt2_embd, gpt2_layers, gpt2_heads;
  float gpt2_eps;
  /* config.py:8-9,15: latent length (dim_z = 20), number of init tokens ("the picture of" = 3), max_tokens_len = 30 */
  int32_t dim_z, n_init, max_tokens_len;
  /* clip/model.py:363-392, text half */
  int32_t text_width, text_heads, text_layers, text_context, text_vocab, text_embed_dim;
  int32_t max_population;
  int32_t device;
  int32_t flags;
} glass_text_config;
/* glass_text_config.flags: GLASS_FLAG_NO_GRAPH (32) = launch eagerly; cross-check variant of the decode-step GEMMs: */
#define GLASS_TEXT_FLAG_NO_SPLIT_K 1   /* one CTA per n-tile over the whole K, epilogue-fused bias / GELU / residual */
int glass_text_create(const glass_text_config* cfg, glass_text_engine** out);
/* Packed tensors from HOST memory; names / layouts in clip_glass_b200/text_packing.py (from the reference's
 * gpt2-pytorch_model.bin and ViT-B-32.pt state_dict layouts). */
int glass_text_set_tensor(glass_text_engine* e, const char* name, const void* host_data, size_t nbytes);
int glass_text_finalize(glass_text_engine* e);
/* generator.py:26-27 caches image_features [1,E] (CLIP.encode_image of the target picture); fp32 here. */
int glass_text_set_image_features(glass_text_engine* e, const float* host_image, int32_t n);
/* models.py:45-60 GPT2.generate up to the token level: z HOST int64 [pop, dim_z] (latent.py:55-56) -> tokens HOST int64
 * [pop, dim_z + n_init + max_tokens_len] = cat(z, init tokens, generated tokens), exactly what gpt2/sample.py:21-37
 * returns with sample=False.  A latent token outside [0, vocab) is GLASS_ERR_ARG (the reference raises IndexError). */
int glass_text_generate(glass_text_engine* e, const int64_t* z_host, int32_t pop, int64_t* tokens_host, void* stream);
/* generator.py:53-59 after clip.tokenize: clip tokens HOST int64 [pop, context] -> sim HOST fp32 [pop] (cosine of
 * CLIP.encode_text against the cached image features); features_host [pop, E] optional (may be NULL). */
int glass_text_similarity(glass_text_engine* e, const int64_t* clip_tokens_host, int32_t pop, float* sim_host,
                          float* features_host, void* stream);
int64_t glass_text_launch_count(const glass_text_engine* e);
/* bench.py roofline: with timing enabled every tensor-core GEMM launch is bracketed by CUDA events;
 * glass_text_gemm_time returns (and resets) their summed device time, their count and their algorithmic bytes
 * (operands read once + outputs written once). */
int glass_text_set_timing(glass_text_engine* e, int32_t enable);
int glass_text_gemm_time(glass_text_engine* e, float* ms, int32_t* launches, double* bytes);
const char* glass_text_last_error(void);
int glass_text_destroy(glass_text_engine* e);

/* -- GPU-resident genetic operators ------------------------- */
/* The reference's search loop (run.py:59-76: pymoo get_algorithm("ga" | "nsga2") + minimize) runs its operators
 * (operators.py:66-78: real_sbx / real_pm, int_sbx / int_pm) on the host and ships the whole population through
 * _evaluate every generation (latent.py:38 H2D, problem.py:20,24 D2H).  These entry points are the same operators on
 * DEVICE buffers, all asynchronous on `stream`, so that a generation is
 *     glass_ga_uniform -> glass_ga_permutations -> glass_ga_tournament -> glass_ga_offspring -> glass_ga_dedup_append
 *     -> glass_evaluate_device -> glass_ga_survive -> glass_ga_gather
 * without a host round trip (clip_glass_b200/device_ga.py drives it).  pymoo 0.4.2.1 is not vendored by the reference:
 * the arithmetic follows clip_glass_b200/ga.py (parity with pymoo itself unpinned, see that file's header); given the
 * same uniform draws the device operators reproduce ga.py's children to the last bit of pow(), and its survivors,
 * ranks and crowding distances exactly.  Errors: glass_ga_last_error(). */
typedef struct glass_ga_params {
  double sbx_eta, sbx_prob, sbx_prob_var;   /* operators.py:69: eta 3, prob 1; pymoo's per-variable probability 0.5 */
  double pm_eta, pm_prob;                   /* operators.py:70: eta 3, prob 0.5; pm_prob < 0 => 1 / n_var           */
  int32_t n_var;
  int32_t integer;                          /* operators.py:75-77 int_sbx / int_pm: rint + clip after each operator */
} glass_ga_params;
/* out_dev[0..n): uniforms in [0,1) with 53 random bits, Philox4x32-10 keyed by `seed`; element i depends on
 * (seed, offset + i/2) only (offset counts PAIRS), so any split of a request reproduces the same stream. */
int glass_ga_uniform(uint64_t seed, uint64_t offset, double* out_dev, int64_t n, void* stream);
/* n_perm random permutations of 0..n-1 from n_perm*n uniform keys: out[p] = stable argsort(keys[p]).  n <= 4096. */
int glass_ga_permutations(const double* keys_dev, int32_t n, int32_t n_perm, int32_t* out_dev, void* stream);
/* Binary tournament: pair t = (pairs[2t], pairs[2t+1]); lower rank wins, then larger crowding distance, then the first. */
int glass_ga_tournament(const int32_t* pairs_dev, const int32_t* rank_dev, const double* crowd_dev, int32_t n_select,
                        int32_t* selected_dev, void* stream);
/* Number of uniforms one glass_ga_offspring call consumes: SBX do / u / swap [M][V], SBX keep [M], PM do / u [2M][V]. */
int64_t glass_ga_rand_count(int32_t n_matings, int32_t n_var);
/* x_dev f64 [n][V] population, parents_dev int32 [M][2] row indices, bounds_dev f64 [4][V] (operator lower / upper
 * bound, then the variable's own lower / upper bound for the integer rounding), rand_dev the uniforms in the layout
 * above.  out_dev f64 [2M][V]: the first children of all matings, then the second children (pymoo's reshape). */
int glass_ga_offspring(const glass_ga_params* params, const double* x_dev, const int32_t* parents_dev,
                       const double* bounds_dev, const double* rand_dev, int32_t n_matings, double* out_dev,
                       void* stream);
/* eliminate_duplicates=True (run.py:66): candidates equal (max |diff| <= eps) to a population row, an accepted
 * offspring or an earlier candidate are dropped; the others are appended to off_dev [n_off][V] at *n_have_dev (device
 * counter, advanced) until n_off rows exist; z32_dev (optional) receives the same rows as fp32 (latent.py:38).
 * workspace_dev: glass_ga_dedup_workspace(n_cand) bytes. */
int64_t glass_ga_dedup_workspace(int32_t n_cand);
int glass_ga_dedup_append(const double* cand_dev, int32_t n_cand, const double* x_dev, int32_t n_x, double* off_dev,
                          int32_t n_off, int32_t* n_have_dev, int32_t n_var, double eps, int32_t eliminate,
                          float* z32_dev, void* workspace_dev, void* stream);
/* Rows [*n_have, n_off) repeat the last accepted row (no-op when the offspring buffer is full). */
int glass_ga_pad(double* off_dev, int32_t n_off, const int32_t* n_have_dev, int32_t n_var, float* z32_dev,
                 void* stream);
int glass_ga_cast_f32(const double* x_dev, float* z_dev, int64_t n, void* stream);
/* Survival of the merged population.  f_dev fp32 column-major: objective k of candidate j at f_dev[k*ld + j] (the
 * layout glass_evaluate_device writes: one array per objective).  nsga2 != 0: fast non-dominated sort + crowding
 * distance, the last front truncated by crowding; nsga2 == 0: the n_survive smallest f_dev[j] (stable).  Outputs
 * (n_survive entries): candidate index, rank, crowding distance (GA: -F), in survivor order.  n <= 4096; one CTA.
 * workspace_dev: glass_ga_survive_workspace(n) bytes. */
int64_t glass_ga_survive_workspace(int32_t n);
int glass_ga_survive(const float* f_dev, int32_t ld, int32_t n, int32_t n_obj, int32_t n_survive, int32_t nsga2,
                     int32_t* idx_dev, int32_t* rank_dev, double* crowd_dev, void* workspace_dev, void* stream);
/* x_out[r] = x_all[idx[r]], f_out[k*ld_out + r] = f_all[k*ld_in + idx[r]] for r < n_out (n_obj <= n_var). */
int glass_ga_gather(const double* x_all_dev, const float* f_all_dev, int32_t ld_in, const int32_t* idx_dev,
                    int32_t n_out, int32_t n_var, int32_t n_obj, double* x_out_dev, float* f_out_dev, int32_t ld_out,
                    void* stream);
const char* glass_ga_last_error(void);

/* -- introspection ---------------------------------------------------------- */
const char* glass_last_error(void);
/* Number of kernels launched by this engine since creation (bench.py's gpu_launches). */
int64_t glass_launch_count(const glass_engine* e);
/* Copy a named intermediate of the last call to HOST as fp32 (tests only;
 * names documented in engine.cu: "w", "styles", "act:<layer>", "rgb:<block>",
 * "tokens", "features", ...).  Returns element count or a negative status. */
int64_t glass_debug_read(glass_engine* e, const char* name, float* host_out, int64_t capacity);
/* capture != 0: keep host copies of intermediates for glass_debug_read (slow;
 * tests only).  timing != 0: bracket every tensor-core launch with CUDA events. */
int glass_set_debug(glass_engine* e, int32_t capture, int32_t timing);
/* 1 if the library was built with -DGLASS_DEBUG (GLASS_DEBUG_* environment knobs honoured: A/B and work-skipping
 * experiments), 0 for the product build, which never reads the environment.  bench.py refuses a debug build. */
int glass_debug_build(void);
/* fp16 range evidence (tests only; forces eager launches): with enable != 0 every G / D activation tensor written by
 * later calls is scanned; glass_range_report returns the totals since the last glass_set_range_check: values that
 * are inf/NaN, values saturated at +-65504 by the epilogues' satfinite conversion, and the largest finite |x|. */
int glass_set_range_check(glass_engine* e, int32_t enable);
int glass_range_report(glass_engine* e, int64_t* nonfinite, int64_t* saturated, float* max_abs);
/* Per-launch device time (ms) and algorithmic FLOPs (as the reference writes
 * the op) of the tensor-core launches of the last timed call, in launch order
 * (G convs, CLIP GEMMs, D convs).  Returns the number of entries written. */
int glass_conv_breakdown(glass_engine* e, float* ms, double* flops, int32_t cap);
/* Time (ms, CUDA events on `stream`) spent in the tensor-core conv/GEMM
 * kernels during the last evaluate call, and their launch count. */
int glass_last_conv_time(const glass_engine* e, float* ms, int32_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* CLIPGLASS_B200_H_ */
