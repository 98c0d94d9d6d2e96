"""Achieved parity errors of the CUDA path against the committed reference fixtures (tests/golden/*.npz, written by
oracle/make_golden.py from the UNMODIFIED reference modules).

    python -m tests.parity_report [--out profiles/r02_parity_full.json]

One record per fixture: max relative error of the CLIP cosine against the fp32 oracle value ("before the final
cast") and against the reference's own fp16-rounded value, max absolute / relative error of the D logit and hinge,
image error, and the fp16 range counters (non-finite / saturated activations).  Plus the benchmarked launch plan
(max_population = 64, P = 64, 16 noise groups): golden candidates placed as groups 0, 7 and 15 among filler latents.
tests/test_gpu_parity.py asserts on the same records; this script commits them as a table.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from clip_glass_b200 import weights as W                     # noqa: E402
from tests.fixtures import FULL_FIXTURES, build_inputs, load_golden   # noqa: E402


def score_errors(neg_sim, hinge, gold, rows=None):
    """Error record of scores against a fixture (``rows``: fixture rows the scores correspond to)."""
    rows = np.arange(len(neg_sim)) if rows is None else np.asarray(rows)
    sim32 = gold["sim_oracle_fp32"][rows]
    ref16 = gold["sim_fp16"].astype(np.float32)[rows]
    rec = dict(
        sim_rel_vs_oracle_fp32=float((np.abs(-neg_sim - sim32) / np.abs(sim32)).max()),
        sim_rel_vs_reference_fp16=float((np.abs(-neg_sim - ref16) / np.abs(ref16)).max()),
        sim_min=float(np.abs(sim32).min()),
    )
    if hinge is not None:
        hg = gold["F"][rows, 1]
        dis = gold["dis"].reshape(-1)[rows]
        rec.update(
            hinge_abs=float(np.abs(hinge - hg).max()),
            hinge_rel=float((np.abs(hinge - hg) / np.maximum(np.abs(hg), 1e-6)).max()),
            logit_abs_max=float(np.abs(dis).max()),
            hinge_err_over_max_logit=float(np.abs(hinge - hg).max() / np.abs(dis).max()),
        )
    return rec


def fixture_record(name, flags=0, with_images=False, use_d=True, range_check=False):
    """Evaluate one fixture through the C ABI on a fresh engine sized for it."""
    from clip_glass_b200.engine import GlassEngine
    inp, gold = build_inputs(name), load_golden(name)
    eng = GlassEngine(inp["gan"], inp["clip"], inp["g_sd"], inp["d_sd"] if use_d else None, inp["c_sd"],
                      batch_size=inp["batch"], max_population=inp["pop"], flags=flags)
    eng.set_text_features(torch.from_numpy(gold["text_features"]))
    rec = dict(fixture=name, pop=inp["pop"], flags=flags, use_discriminator=use_d)
    if range_check:
        eng.set_range_check(True)
    neg_sim, hinge = eng.evaluate(inp["x"], noise=inp["noise"])
    if range_check:
        rec["range"] = eng.range_report()
        eng.set_range_check(False)
    if use_d:
        rec.update(score_errors(neg_sim, hinge, gold))
    else:
        ref = -gold["F_nod"].astype(np.float32)
        rec["sim_rel_vs_reference_fp16"] = float((np.abs(-neg_sim - ref) / np.abs(ref)).max())
        rec["sim_rel_vs_oracle_fp32"] = float((np.abs(-neg_sim - gold["sim_oracle_fp32"]) /
                                              np.abs(gold["sim_oracle_fp32"])).max())
    rec["finite"] = bool(np.isfinite(neg_sim).all() and (hinge is None or np.isfinite(hinge).all()))
    if with_images:
        z = torch.from_numpy(inp["x"]).float().cuda()
        images = eng.generate(z, noise=inp["noise"])
        pool = max(1, inp["gan"].resolution // 64)
        small = torch.nn.functional.avg_pool2d(images, pool).cpu().numpy()
        rec["image64_abs"] = float(np.abs(small - gold["images_64"]).max())
        rec["image_mean_abs"] = float(np.abs(images.mean(dim=(1, 2, 3)).cpu().numpy() - gold["image_mean"]).max())
    eng.close()
    return rec


def embedded_p64_record(name="full_c", P=64, groups_at=((0, 0), (7, 1), (15, 0))):
    """The benchmarked launch plan: max_population = P = 64 (16 noise groups of 4).  Golden minibatch groups of
    fixture ``name`` are placed at the given (slot, fixture group) positions among filler latents / filler noise;
    their scores must equal the fixture's."""
    from clip_glass_b200.engine import GlassEngine
    inp, gold = build_inputs(name), load_golden(name)
    B = inp["batch"]
    eng = GlassEngine(inp["gan"], inp["clip"], inp["g_sd"], inp["d_sd"], inp["c_sd"], batch_size=B, max_population=P)
    eng.set_text_features(torch.from_numpy(gold["text_features"]))
    x = W.make_latents(P, inp["gan"].latent_size, 4242)
    noise = W.make_noise(inp["gan"], P // B, 4243)
    rows_out, rows_fix = [], []
    for slot, g in groups_at:
        x[slot * B:(slot + 1) * B] = inp["x"][g * B:(g + 1) * B]
        noise[slot] = inp["noise"][g]
        rows_out += list(range(slot * B, (slot + 1) * B))
        rows_fix += list(range(g * B, (g + 1) * B))
    neg_sim, hinge = eng.evaluate(x, noise=noise)           # first evaluation of the plan: eager launches
    neg_sim2, hinge2 = eng.evaluate(x, noise=noise)         # second: CUDA-graph replay (what bench.py times)
    rec = dict(fixture=name, plan=f"max_population={P}, P={P}, batch_size={B}", placed=[list(t) for t in groups_at])
    rec.update(score_errors(neg_sim2[rows_out], hinge2[rows_out], gold, rows_fix))
    rec["graph_equals_eager"] = bool(np.array_equal(neg_sim, neg_sim2) and np.array_equal(hinge, hinge2))
    rec["all_finite"] = bool(np.isfinite(neg_sim2).all() and np.isfinite(hinge2).all())
    # the same candidates evaluated alone (P = 8 plan) give bit-identical scores
    xs = np.concatenate([x[s * B:(s + 1) * B] for s, _ in groups_at[:2]])
    ns = [noise[s] for s, _ in groups_at[:2]]
    a_sim, a_h = eng.evaluate(xs, noise=ns)
    rec["equals_small_plan_bitwise"] = bool(np.array_equal(a_sim, neg_sim2[rows_out[:2 * B]]) and
                                            np.array_equal(a_h, hinge2[rows_out[:2 * B]]))
    eng.close()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "parity_full.json"))
    args = ap.parse_args()
    recs = [fixture_record("tiny", with_images=True), fixture_record("tiny_stress", range_check=True)]
    for name in FULL_FIXTURES:
        recs.append(fixture_record(name, with_images=True, range_check=True))
    recs.append(fixture_record("full_b", use_d=False))
    for flags in (1, 2):
        recs.append(fixture_record("full_b", flags=flags))
    recs.append(embedded_p64_record())
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(recs, f, indent=1)
    for r in recs:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
