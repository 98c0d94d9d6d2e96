"""Layer-by-layer GPU diagnostic: engine intermediates vs the CPU emulation
(tests/emulate.py, identical rounding points) and final scores vs the oracle.

    python -m tests.gpu_diag --config tiny --impl 0      # tcgen05 path
    python -m tests.gpu_diag --config tiny --impl 1      # SIMT bring-up path

Prints one line per tensor: name, max |err|, max |ref|, relative.  Used by
tests/test_gpu_parity.py (in a subprocess with a timeout, so that a trapped or
hung kernel cannot take the test session down) and by hand when bringing up a
kernel.
"""
from __future__ import annotations

import argparse
import json
import sys
import time

import numpy as np
import torch

from clip_glass_b200 import packing, weights as W
from clip_glass_b200.engine import GlassEngine
from oracle import evaluate_oracle
from tests import emulate as E
from tests.fixtures import build_inputs, load_golden


def err_line(name, got, ref):
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    if got.shape != ref.shape:
        return dict(name=name, shape_mismatch=[int(got.size), int(ref.size)])
    d = np.abs(got - ref)
    scale = float(np.abs(ref).max()) + 1e-30
    bad = int((~np.isfinite(got)).sum())
    return dict(name=name, max_err=float(d.max()), ref_max=scale, rel=float(d.max() / scale),
                mean_err=float(d.mean()), nonfinite=bad)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="tiny")
    ap.add_argument("--impl", type=int, default=0)
    ap.add_argument("--no-capture", action="store_true")
    ap.add_argument("--flags", type=int, default=0, help="glass_config.flags (1 = folded up/down convs everywhere, 2 = exact everywhere)")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()

    inp = build_inputs(args.config)
    gold = load_golden(args.config)
    gan, clip, P, B = inp["gan"], inp["clip"], inp["pop"], inp["batch"]
    text = torch.from_numpy(gold["text_features"])
    pk = {}
    pk.update(packing.pack_generator(inp["g_sd"], gan))
    pk.update(packing.pack_discriminator(inp["d_sd"], gan))
    pk.update(packing.pack_clip_visual(inp["c_sd"], clip))

    t0 = time.time()
    eng = GlassEngine(gan, clip, inp["g_sd"], inp["d_sd"], inp["c_sd"], batch_size=B, max_population=P,
                      conv_impl=args.impl, flags=args.flags)
    eng.set_text_features(text)
    eng.set_debug(capture=not args.no_capture)
    print(f"engine ready in {time.time() - t0:.1f}s", flush=True)

    z = evaluate_oracle.latents_from_population(inp["x"])
    zc = z.cuda()
    lines = []
    images = eng.generate(zc, noise=inp["noise"])
    torch.cuda.synchronize()
    print("generate done", flush=True)

    cap = {}
    exact = (args.flags & 2) != 0      # tiny config: the default cost model keeps every layer folded
    emu_img = E.emu_generator(pk, gan, z, inp["noise"], B, capture=cap, exact=exact)
    layers = packing.g_layers(gan)
    if not args.no_capture:
        lines.append(err_line("w", eng.debug_read("w"), cap["w"]))
        lines.append(err_line("styles", eng.debug_read("styles"), cap["styles"]))
        lines.append(err_line("x0", eng.debug_read("x0"), cap["x0"]))
        for li in range(len(layers) - 1):
            if f"u{li}" in cap:
                lines.append(err_line(f"u{li}", eng.debug_read(f"u{li}"), cap[f"u{li}"]))
            lines.append(err_line(f"xs{li}", eng.debug_read(f"xs{li}"), cap[f"xs{li}"]))
        for b in range(gan.num_blocks):
            got = eng.debug_read(f"rgb{b}").reshape(-1, 4)[:, :3]
            lines.append(err_line(f"rgb{b}", got, cap[f"rgb{b}"].reshape(-1, 3)))
    lines.append(err_line("images_vs_emulation", images.cpu().numpy(), emu_img))

    sim = eng.clip_similarity(images)
    torch.cuda.synchronize()
    print("clip done", flush=True)
    ccap = {}
    feats, emu_sim = E.emu_clip(pk, clip, images.cpu(), text, capture=ccap)
    if not args.no_capture:
        lines.append(err_line("ln_pre", eng.debug_read("ln_pre"), ccap["ln_pre"]))
        for l in range(clip.layers):
            lines.append(err_line(f"block{l}", eng.debug_read(f"block{l}"), ccap[f"block{l}"]))
        lines.append(err_line("features", eng.debug_read("features"), feats))
    lines.append(err_line("sim_vs_emulation", sim.cpu().numpy(), emu_sim))

    logits = eng.discriminate(images)
    torch.cuda.synchronize()
    print("discriminate done", flush=True)
    dcap = {}
    emu_logits = E.emu_discriminator(pk, gan, images.cpu(), B, capture=dcap, exact=exact)
    if not args.no_capture:
        for b in range(gan.num_blocks - 1):
            lines.append(err_line(f"d{b}", eng.debug_read(f"d{b}"), dcap[f"d{b}"]))
    lines.append(err_line("logits_vs_emulation", logits.cpu().numpy(), emu_logits))

    # end to end through the host entry point, against the oracle (fp32 CLIP arithmetic)
    eng.set_debug(capture=False)
    neg_sim, hinge = eng.evaluate(inp["x"], noise=inp["noise"])
    ref = evaluate_oracle.evaluate(inp["x"], inp["g_sd"], inp["d_sd"], W.clip_as_built(inp["c_sd"]), text, gan, clip,
                                   B, True, noise=inp["noise"], clip_mode="fp32", return_images=True)
    lines.append(err_line("images_vs_oracle", images.cpu().numpy(), ref["images"].numpy()))
    rel = np.abs(neg_sim - ref["F"][:, 0]) / np.abs(ref["F"][:, 0])
    lines.append(dict(name="neg_sim_vs_oracle", max_rel=float(rel.max()), got=neg_sim.tolist(),
                      ref=ref["F"][:, 0].tolist()))
    lines.append(dict(name="hinge_vs_oracle", max_abs=float(np.abs(hinge - ref["F"][:, 1]).max()),
                      got=hinge.tolist(), ref=ref["F"][:, 1].tolist()))
    relg = np.abs(-neg_sim - gold["sim_fp16"].astype(np.float32)) / np.abs(gold["sim_fp16"].astype(np.float32))
    lines.append(dict(name="sim_vs_reference_fixture_fp16", max_rel=float(relg.max())))
    lines.append(dict(name="launches", count=eng.launch_count))
    for ln in lines:
        print(json.dumps(ln), flush=True)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(lines, f, indent=1)
    eng.close()


if __name__ == "__main__":
    main()
