"""Run N img2txt evaluations at the BASELINE config-5 shape (for ncu / timing).
    python tests/profile_text.py --pop 64 --evals 2
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_glass_b200 import text_weights as TW                  # noqa: E402
from clip_glass_b200.models import standin_clip_tokens          # noqa: E402
from clip_glass_b200.text_engine import TextEngine              # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pop", type=int, default=64)
    ap.add_argument("--evals", type=int, default=2)
    args = ap.parse_args()
    g, t = TW.GPT2_SMALL, TW.CLIP_TEXT_B32
    eng = TextEngine(g, TW.make_gpt2_weights(g, 1000), t, TW.make_clip_text_weights(t, 1001), init_tokens=[1169, 4286, 286],
                     max_population=args.pop)
    eng.set_image_features(torch.randn(1, 512, generator=torch.Generator().manual_seed(6)))
    z = TW.make_token_latents(args.pop, 20, g.vocab, 50)
    for i in range(args.evals):
        t0 = time.perf_counter()
        toks = eng.generate_tokens(z)
        t1 = time.perf_counter()
        gen = [s[20:] for s in toks.tolist()]
        sim = eng.text_similarity(standin_clip_tokens(gen, t))
        t2 = time.perf_counter()
        print(f"eval {i}: generate {1e3 * (t1 - t0):.2f} ms, similarity {1e3 * (t2 - t1):.2f} ms")
    eng.close()


if __name__ == "__main__":
    main()
