"""Run N device-resident fitness evaluations of the BASELINE workload (for ncu / timing).
    python tests/profile_step.py --pop 64 --evals 2 [--timing]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_glass_b200 import packing, weights as W            # noqa: E402
from clip_glass_b200.engine import GlassEngine               # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pop", type=int, default=64)
    ap.add_argument("--evals", type=int, default=2)
    ap.add_argument("--nod", action="store_true")
    ap.add_argument("--timing", action="store_true")
    ap.add_argument("--flags", type=int, default=0)
    args = ap.parse_args()
    gan, clip = W.FFHQ, W.VIT_B32
    eng = GlassEngine(gan, clip, W.make_generator_weights(gan, 1000),
                      None if args.nod else W.make_discriminator_weights(gan, 1001),
                      W.make_clip_visual_weights(clip, 1002), batch_size=4, max_population=args.pop, flags=args.flags)
    eng.set_text_features(torch.randn(1, 512, generator=torch.Generator().manual_seed(5)))
    z = torch.from_numpy(W.make_latents(args.pop, 512, 50)).float().cuda()
    if args.timing:
        eng.set_debug(timing=True)
    step_ms = []
    best = None
    for i in range(args.evals):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        eng.evaluate_device(z, seed=1 + i)
        t1.record()
        torch.cuda.synchronize()
        step_ms.append(t0.elapsed_time(t1))
        if args.timing and i > 0:        # per-layer minimum over the evaluations after the first (warm-up)
            bd = eng.conv_breakdown()
            best = bd if best is None else [(min(a[0], b[0]), a[1]) for a, b in zip(best, bd)]
    print("step ms per eval:", " ".join(f"{m:.2f}" for m in step_ms), f"(lib {os.environ.get('CLIPGLASS_LIB', 'default')})")
    if args.timing:
        layers = packing.g_layers(gan)
        names = [f"G{li}:{'up' if l['up'] else 'cv'}{l['cin']}->{l['cout']}@{l['res']}" for li, l in enumerate(layers)]
        names += ["C:patch"] + [f"C{l}:{n}" for l in range(clip.layers) for n in ("qkv", "out", "fc", "proj")]
        if not args.nod:
            ch = list(gan.channels)
            for b in range(gan.num_blocks - 1):
                r = gan.resolution >> b
                names += [f"D{b}:c0 {ch[b]}@{r}", f"D{b}:proj {ch[b]}->{ch[b+1]}@{r//2}", f"D{b}:c1 {ch[b]}->{ch[b+1]}@{r//2}"]
            names += ["D:fin", "D:dense0"]
        bd = best if best is not None else eng.conv_breakdown()
        tot = sum(m for m, _ in bd)
        print(f"pop={args.pop} flags={args.flags} conv launches={len(bd)} total conv ms={tot:.3f}")
        for n, (ms, fl) in zip(names, bd):
            print(f"{n:28s} {ms:9.4f} ms  {fl/1e9:10.2f} GFLOP  {fl/ms/1e9 if ms > 0 else 0:9.1f} TFLOP/s")
    eng.close()


if __name__ == "__main__":
    main()
