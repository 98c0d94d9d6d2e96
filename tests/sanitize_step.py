"""One small invocation of every CUDA path, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tests/sanitize_step.py

Tiny StyleGAN2 + CLIP configuration (64x64 images, P = 8) through the fused evaluate (eager and graph replay), the
facade calls, the exact-resampling variant, the image-output kernels, the BigGAN latent kernel and the tiny img2txt
engine (GPT-2 decode + CLIP text tower).  Prints the scores so that a sanitizer run can also be checked for value
changes.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_glass_b200 import text_weights as TW, weights as W      # noqa: E402
from clip_glass_b200.engine import GlassEngine                    # noqa: E402
from clip_glass_b200.text_engine import TextEngine                # noqa: E402


def full_size():
    """--full: ONE eager full-size evaluation (ffhq-f 1024^2 + ViT-B/32, P = 4): the kernels the tiny configuration never
    reaches -- the fused-FIR down-convs (32 -> 64 and the 128-column 64 -> 128 instance), the I8 / tile-pair conv modes,
    the image-finishing epilogue, the specialised GEMM epilogues need M % 128 == 0 and stay on the run-time spec here."""
    gan, clip = W.FFHQ, W.VIT_B32
    eng = GlassEngine(gan, clip, W.make_generator_weights(gan, 1000), W.make_discriminator_weights(gan, 1001),
                      W.make_clip_visual_weights(clip, 1002), batch_size=4, max_population=4, flags=32)
    eng.set_text_features(torch.randn(1, 512, generator=torch.Generator().manual_seed(5)))
    f, h = eng.evaluate(W.make_latents(4, 512, 50), seed=3)
    print("full size neg_sim", f, "hinge", h)
    eng.close()
    torch.cuda.synchronize()
    print("sanitize_step: full-size evaluation done")


def main():
    if "--full" in sys.argv:
        return full_size()
    gan, clip = W.TINY_GAN, W.TINY_CLIP
    P, B = 8, 4
    text = torch.randn(1, 512, generator=torch.Generator().manual_seed(5))
    x = W.make_latents(P, 512, 4)
    for flags in (0, 2):
        eng = GlassEngine(gan, clip, W.make_generator_weights(gan, 1), W.make_discriminator_weights(gan, 2),
                          W.make_clip_visual_weights(clip, 3), batch_size=B, max_population=P, flags=flags)
        eng.set_text_features(text)
        for _ in range(3 if flags == 0 else 1):                   # eager, graph capture, graph replay
            f, h = eng.evaluate(x, seed=7)
        print("flags", flags, "neg_sim", f[:3], "hinge", h[:3])
        if flags == 0:
            z = torch.from_numpy(x).float().cuda()
            img = eng.generate(z, seed=7)
            eng.clip_similarity(img)
            eng.discriminate(img)
            eng.image_grid_u8(img, nrow=3)
            eng.last_images([1, 5])
        eng.close()
    spec_g, spec_t = TW.TINY_GPT2, TW.TINY_CLIP_TEXT
    te = TextEngine(spec_g, TW.make_gpt2_weights(spec_g, 1), spec_t, TW.make_clip_text_weights(spec_t, 2),
                    init_tokens=[5, 6, 7], max_population=P)
    te.set_image_features(text)
    z = TW.make_token_latents(P, 20, spec_g.vocab, 3)
    for _ in range(3):
        toks = te.generate_tokens(z)
    ct = np.zeros((P, spec_t.context), dtype=np.int64)
    ct[:, 0] = spec_t.vocab - 2
    ct[:, 1:5] = toks[:, 23:27] % 4000 + 256
    ct[:, 5] = spec_t.vocab - 1
    for _ in range(3):
        sim = te.text_similarity(ct)
    print("tokens", toks[0, 23:29], "sim", sim[:3])
    te.close()
    torch.cuda.synchronize()
    print("sanitize_step: done")


if __name__ == "__main__":
    main()
