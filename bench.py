#!/usr/bin/env python
"""bench.py — candidates evaluated per second through the fitness path.

One "step" = one generation's ``GenerationProblem._evaluate`` over a synthetic
population (problem.py:14-29): latents -> StyleGAN2 ffhq-config-f G -> CLIP
ViT-B/32 -> cosine -> StyleGAN2 D hinge.  Workload = BASELINE.json configs[1]
(StyleGAN2_ffhq_d, pop 64 per GPU, batch_size 4); at N GPUs the population is
N*64 sharded by whole minibatches with one all-gather of F (configs[3] at N=8):
weak scaling.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  See the module-level contract in the task
statement; keys: metric value unit n_gpus steps warmup ms_per_step
higher_is_better scaling vs_baseline dtype data config clocks e2e gpu_launches
roofline cpu_baseline.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "candidate latents evaluated/sec StyleGAN2_ffhq+CLIP (problem._evaluate path)"
UNIT = "candidates/s"
# algorithmic work per candidate as the reference writes the ops (BASELINE.md §2 / SURVEY.md §8d)
GFLOP_PER_CAND = {"_d": 317.1, "_nod": 159.6}


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(tflops_burst=d.get("bf16_tflops"), tflops_sustained=d.get("bf16_tflops_sustained"),
                    hbm_gbs=d.get("hbm_gbs"), source="measured (MEASURED_PEAKS.json)")
    return dict(tflops_burst=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt, self.proc = index, [], threading.Event(), None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self._stop_evt.is_set():
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx or None,
                    reasons=sorted(reasons), samples=len(sm))


def conv_launch_names(use_d):
    """Short names of the tensor-core launches of one step, in launch order (G convs, CLIP GEMMs, D convs)."""
    names = [f"G{i}" for i in range(17)] + ["Cpatch"] + [f"C{l}{n}" for l in range(12) for n in ("qkv", "out", "fc", "proj")]
    if use_d:
        for b in range(8):
            names += [f"D{b}c0", f"D{b}proj", f"D{b}c1"]
        names += ["Dfin", "Ddense0"]
    return names


def best_cpu_threads():
    """Threads for the CPU arm.  One 4-candidate minibatch does not scale past ~32 threads: measured on the
    128-core GPU box (tests/cpu_threads_probe.py): 8 -> 0.35, 16 -> 0.38, 32 -> 0.38, 64 -> 0.32, 128 -> 0.05
    candidates/s.  Using every core would make the baseline 7x slower than it can be, so cap at 32."""
    return min(os.cpu_count() or 1, 32)


def cpu_kind():
    """"reference": the reference's own stylegan2 / clip modules (oracle/_ref, vendored by oracle/build_ref.py in the
    build container) run the path; "port": the oracle restatement (when oracle/_ref is absent)."""
    from oracle import reference_modules
    return "reference" if reference_modules.reference_available(reference_modules.VENDORED_ROOT) else "port"


def cpu_oracle_step(pop, batch, use_d, seed, threads):
    """One bounded sample of the SAME workload on the host cores (full ffhq-config-f G + ViT-B/32 + D, fp32 G/D,
    CLIP fp16 as built): problem.py:14-29 over the reference's own modules from oracle/_ref, or over the oracle
    port when those are absent."""
    import torch
    from clip_glass_b200 import weights as W
    from oracle import evaluate_oracle, reference_modules
    torch.set_num_threads(threads)
    st = cpu_oracle_step.__dict__.setdefault("state", {})
    if not st:
        st["kind"] = cpu_kind()
        g = W.make_generator_weights(W.FFHQ, 1000)
        d = W.make_discriminator_weights(W.FFHQ, 1001)
        c = W.make_clip_visual_weights(W.VIT_B32, 1002)
        st["t"] = torch.randn(1, 512, generator=torch.Generator().manual_seed(5)).half()
        if st["kind"] == "reference":
            reference_modules.use_reference_root(reference_modules.VENDORED_ROOT)
            st["G"], st["D"] = reference_modules.build_reference_gan(W.FFHQ, g, d)
            st["C"] = reference_modules.build_reference_clip(W.VIT_B32, c)
        else:
            st["g"], st["d"], st["c"] = g, d, W.clip_as_built(c)
    x = W.make_latents(pop, 512, seed)
    noise = W.make_noise(W.FFHQ, pop // batch, seed + 1)
    t0 = time.perf_counter()
    if st["kind"] == "reference":
        reference_modules.reference_evaluate(x, st["G"], st["D"], st["C"], st["t"], batch, use_d, noise)
    else:
        evaluate_oracle.evaluate(x, st["g"], st["d"], st["c"], st["t"], W.FFHQ, W.VIT_B32, batch, use_d, noise=noise)
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port; /root/reference cannot travel to the GPU box),
    timed on the host cores with every thread it can use.  Rank 0 only."""
    if rank != 0:
        return
    threads = best_cpu_threads()
    use_d = args.variant == "_d"
    sample = args.cpu_sample
    budget_s = 270.0
    t_begin = time.perf_counter()
    times = []
    done_w = 0
    for i in range(args.warmup):
        if time.perf_counter() - t_begin > budget_s * 0.4 and done_w >= 1:
            break
        cpu_oracle_step(sample, args.batch, use_d, 10 + i, threads)
        done_w += 1
    for i in range(args.steps):
        if times and time.perf_counter() - t_begin + times[-1] > budget_s:
            break
        times.append(cpu_oracle_step(sample, args.batch, use_d, 100 + i, threads))
    ms = 1e3 * sum(times) / len(times)
    value = sample / (ms / 1e3)
    line = dict(
        impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=len(times),
        warmup=done_w, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="fp32 (G/D), fp16-as-built (CLIP)", data="synthetic",
        config=dict(workload=f"StyleGAN2_ffhq{args.variant} ffhq-config-f 1024^2 + CLIP ViT-B/32, batch_size {args.batch}",
                    population_per_step=sample, note="bounded sample of the same workload on host cores"),
        cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind=cpu_kind(),
                          sample=f"{sample} candidates per step x {len(times)} steps (problem.py:14-29 over "
                                 f"{'the reference modules in oracle/_ref' if cpu_kind() == 'reference' else 'the oracle port'}; "
                                 f"{threads} of {os.cpu_count()} host threads, the fastest setting measured; "
                                 "steps stop early once ~4.5 min of wall time is used)"),
        e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
    )
    emit(line)


def gpt2_reference_step(pop, seed, threads):
    """--workload gpt2, CPU arm: models.py:45-60 + generator.py:53-59 over the reference's own gpt2 / clip modules
    (oracle/_ref) or the oracle port, on the host cores."""
    import torch
    from clip_glass_b200 import text_weights as TW
    from clip_glass_b200.models import INIT_TEXT_TOKENS, standin_clip_tokens
    from oracle import gpt2_oracle, reference_modules
    torch.set_num_threads(threads)
    st = gpt2_reference_step.__dict__.setdefault("state", {})
    init = INIT_TEXT_TOKENS["the picture of"]
    if not st:
        st["g"] = TW.make_gpt2_weights(TW.GPT2_SMALL, 1000)
        st["t"] = TW.text_as_built(TW.make_clip_text_weights(TW.CLIP_TEXT_B32, 1001))
        st["img"] = torch.randn(1, 512, generator=torch.Generator().manual_seed(6)).half()
        st["kind"] = "port"
        root = reference_modules.VENDORED_ROOT
        if os.path.exists(os.path.join(root, "gpt2", "model.py")) and reference_modules.reference_available(root):
            import importlib
            sys.path.insert(0, root)
            try:
                for k in [k for k in sys.modules if k == "gpt2" or k.startswith("gpt2.")]:
                    del sys.modules[k]
                gm = importlib.import_module("gpt2.model")
                gs = importlib.import_module("gpt2.sample")
                gc = importlib.import_module("gpt2.config")
                gu = importlib.import_module("gpt2.utils")
            finally:
                sys.path.remove(root)
            reference_modules.use_reference_root(root)
            model = gu.load_weight(gm.GPT2LMHeadModel(gc.GPT2Config()), {k: v.clone() for k, v in st["g"].items()}).eval()
            _, clip_mod = reference_modules._import_reference()
            sp = TW.CLIP_TEXT_B32
            clip = clip_mod.CLIP(sp.embed_dim, 64, 1, 64, 32, sp.context, sp.vocab, sp.width, sp.heads, sp.layers)
            clip_mod.convert_weights(clip)
            clip.load_state_dict(st["t"], strict=False)
            st.update(kind="reference", model=model, sample=gs.sample_sequence, clip=clip.eval())
    z = TW.make_token_latents(pop, 20, TW.GPT2_SMALL.vocab, seed)
    t0 = time.perf_counter()
    with torch.no_grad():
        if st["kind"] == "reference":
            ctx = torch.cat((torch.tensor(z).long(), torch.tensor(init).long().repeat(pop, 1)), dim=1)
            toks = np.asarray(st["sample"](model=st["model"], length=30, context=ctx, start_token=None, batch_size=pop,
                                           temperature=0.7, top_k=40, device="cpu", sample=False))
        else:
            toks = gpt2_oracle.gpt2_generate_tokens(st["g"], TW.GPT2_SMALL, z, init, 30)
        gen = gpt2_oracle.parse_out_tokens(toks, 20, TW.GPT2_SMALL.vocab - 1)
        ct = torch.tensor(standin_clip_tokens(gen, TW.CLIP_TEXT_B32))
        feats = st["clip"].encode_text(ct) if st["kind"] == "reference" else \
            gpt2_oracle.clip_encode_text(st["t"], TW.CLIP_TEXT_B32, ct, mode="as_built")
        torch.cosine_similarity(feats.float(), st["img"].float())
    return time.perf_counter() - t0, st["kind"]


GPT2_METRIC = "candidate token-latents evaluated/sec GPT2 img2txt + CLIP text tower (problem._evaluate path, config 5)"


def run_gpt2(args, rank, world, local_rank):
    """--workload gpt2 (BASELINE config 5: GPT2 image-to-text, pop 64 per GPU): one step = one generation's
    _evaluate: 30-step greedy decode of P x 53 tokens + CLIP text tower + cosine."""
    threads = best_cpu_threads()
    P_local = args.pop_per_gpu
    if args.impl == "reference":
        if rank != 0:
            return
        times, kind = [], "port"
        gpt2_reference_step(args.cpu_sample_gpt2, 1, threads)
        for i in range(max(1, min(args.steps, 5))):
            t, kind = gpt2_reference_step(args.cpu_sample_gpt2, 10 + i, threads)
            times.append(t)
        ms = 1e3 * sum(times) / len(times)
        v = args.cpu_sample_gpt2 / (ms / 1e3)
        emit(dict(impl="reference", metric=GPT2_METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=len(times), warmup=1,
                  ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="fp32 (GPT-2), fp16-as-built (CLIP)",
                  data="synthetic", config=dict(workload="GPT2 img2txt (config 5): GPT-2 small greedy decode 23+30 tokens + CLIP text tower",
                                                population_per_step=args.cpu_sample_gpt2),
                  cpu_baseline=dict(value=v, unit=UNIT, cores=threads, kind=kind,
                                    sample=f"{args.cpu_sample_gpt2} candidates per step x {len(times)} steps"),
                  e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0)))
        return
    import torch
    import torch.distributed as tdist
    from clip_glass_b200 import text_weights as TW
    from clip_glass_b200.config import make_namespace
    from clip_glass_b200.models import standin_clip_tokens
    from clip_glass_b200.problem import GenerationProblem
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    P = P_local * world
    img = torch.randn(1, 512, generator=torch.Generator().manual_seed(6))
    config = make_namespace("GPT2", device=f"cuda:{local_rank}", target="synthetic", pop_size=P, batch_size=P_local,
                            max_population=P_local, synthetic_seed=1000, image_features=img, clip_token_map="standin")
    problem = GenerationProblem(config)
    gen = problem.generator
    eng = gen.engine
    stream = torch.cuda.current_stream()
    W_ = max(args.warmup, 3)
    zs = [TW.make_token_latents(P, 20, TW.GPT2_SMALL.vocab, 50 + i) for i in range(4)]
    lo, hi = rank * P_local, (rank + 1) * P_local

    def step_local(i):
        toks = eng.generate_tokens(zs[i % 4][lo:hi])
        return eng.text_similarity(standin_clip_tokens(gen.model.parse_out_tokens(toks), gen.text_spec))

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    for i in range(W_):
        step_local(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = eng.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        ev[i][0].record(stream)
        step_local(W_ + i)
        ev[i][1].record(stream)
    barrier()
    launches = eng.launch_count - l0
    ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device="cuda")
    if world > 1:
        tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
    ms_dev = float(ms.mean())
    e2e = []
    for i in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        out = {}
        problem._evaluate(zs[i % 4].astype(np.float64), out)
        torch.cuda.synchronize()
        e2e.append((time.perf_counter() - t0) * 1e3)
    e2e_ms = torch.tensor(e2e, dtype=torch.float64, device="cuda")
    if world > 1:
        tdist.all_reduce(e2e_ms, op=tdist.ReduceOp.MAX)
    if rank == 0:
        sampler.stop()
    eng.set_timing(True)
    step_local(0)
    g_ms, g_n, g_bytes = eng.gemm_time()
    eng.set_timing(False)
    peaks = measured_peaks()
    achieved = g_bytes / (g_ms * 1e-3) / 1e9
    roofline = dict(bound="hbm", kernel="gemm_tc_kernel (split-fp16 tcgen05 GEMMs of the GPT-2 decode + CLIP text GEMMs)",
                    achieved=achieved, peak=peaks["hbm_gbs"], unit="GB/s", frac=achieved / peaks["hbm_gbs"], traffic=None,
                    peak_source=peaks["source"], launches_per_step=g_n, kernel_ms_per_step=g_ms,
                    share_of_step=g_ms / ms_dev, algorithmic_bytes_per_step=g_bytes,
                    note="decode steps have M = P rows: every GEMM is a pass over its (hi + lo) fp16 weights")
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gpt2_reference_step(args.cpu_sample_gpt2, 1, threads)
        ts = [gpt2_reference_step(args.cpu_sample_gpt2, 2 + i, threads) for i in range(2)]
        v = args.cpu_sample_gpt2 / statistics.median(t for t, _ in ts)
        cpu = dict(value=v, unit=UNIT, cores=threads, kind=ts[0][1],
                   sample=f"{args.cpu_sample_gpt2} candidates x 2 timed repetitions, median (models.py:45-60 + generator.py:53-59)")
    if rank == 0:
        em = float(e2e_ms.mean())
        emit(dict(metric=GPT2_METRIC, value=P / (ms_dev / 1e3), unit=UNIT, n_gpus=world, steps=args.steps, warmup=W_,
                  ms_per_step=ms_dev, higher_is_better=True, scaling="weak", vs_baseline=None,
                  dtype="split fp16 (22-bit) operands, fp32 accumulate for GPT-2; fp16 as built for the CLIP text tower",
                  data="synthetic",
                  config=dict(workload="GPT2 img2txt (config 5): GPT-2 small greedy decode 23+30 tokens + CLIP text tower",
                              population=P, population_per_gpu=P_local, parallelism=f"population-sharded dp{world}",
                              weights="seeded random (no checkpoints offline)",
                              text_round_trip="token-level stand-in for BPE decode + clip.tokenize (no vocabulary files on the box)",
                              l2="decode streams ~0.5 GB of weights per step (> 126 MB L2)"),
                  clocks=sampler.summary(), gpu_launches=int(launches),
                  e2e=dict(value=P / (em / 1e3), unit=UNIT, ms_per_step=em, h2d_bytes_per_step=int(P_local * (20 + 77) * 8),
                           d2h_bytes_per_step=int(P_local * (53 * 8 + 4)),
                           api="GenerationProblem._evaluate(x) -> out['F'] (glass_text_generate + glass_text_similarity)"),
                  roofline=roofline, cpu_baseline=cpu, flags=dict(engine_flags=0, debug_build=False, glass_debug_env=[])))
    if world > 1:
        tdist.destroy_process_group()


_JSON_FD = None


def _guard_stdout():
    """Route fd 1 to stderr for the whole run (NCCL and torchrun children print banners on stdout) and keep the
    original stdout for the ONE JSON line of the contract."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pop-per-gpu", type=int, default=64)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--variant", default="_d", choices=["_d", "_nod"])
    ap.add_argument("--cpu-sample", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--flags", type=int, default=0, help="glass_config.flags (tuning experiments; 0 = product default)")
    ap.add_argument("--workload", default="stylegan2", choices=["stylegan2", "gpt2"],
                    help="stylegan2 = BASELINE configs 2/4 (the headline metric); gpt2 = config 5 (img2txt path)")
    ap.add_argument("--cpu-sample-gpt2", type=int, default=16)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.workload == "gpt2":
        if args.impl == "ours":
            from clip_glass_b200 import _lib as _l
            if sorted(k for k in os.environ if k.startswith("GLASS_DEBUG_")) or os.environ.get("CLIPGLASS_LIB") or \
                    _l.load_library().glass_debug_build():
                raise SystemExit("bench.py: refusing to time a debug configuration")
        run_gpt2(args, rank, world, local_rank)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as tdist
    from clip_glass_b200 import weights as W
    from clip_glass_b200.config import make_namespace
    from clip_glass_b200.problem import GenerationProblem

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the fitness path has no CPU fallback)")
    from clip_glass_b200 import _lib
    dbg_env = sorted(k for k in os.environ if k.startswith("GLASS_DEBUG_"))
    if dbg_env or os.environ.get("CLIPGLASS_LIB") or _lib.load_library().glass_debug_build():
        raise SystemExit(f"bench.py: refusing to time a debug configuration (GLASS_DEBUG_* set: {dbg_env}; "
                         f"CLIPGLASS_LIB={os.environ.get('CLIPGLASS_LIB')!r}; "
                         f"debug build: {bool(_lib.load_library().glass_debug_build())})")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W_ = max(args.warmup, 3)
    use_d = args.variant == "_d"
    P_local = args.pop_per_gpu
    P = P_local * world
    cfg_name = "StyleGAN2_ffhq_d" if use_d else "StyleGAN2_ffhq_nod"
    text = torch.randn(1, 512, generator=torch.Generator().manual_seed(5))
    config = make_namespace(cfg_name, device=f"cuda:{local_rank}", target="synthetic", pop_size=P,
                            batch_size=args.batch, max_population=P_local, synthetic_seed=1000,
                            text_features=text, noise_seed=7, engine_flags=args.flags)
    problem = GenerationProblem(config)
    eng = problem.generator.engine
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    from clip_glass_b200 import dist as gdist
    bounds = gdist.shard_bounds(P, args.batch, world)
    s0, e0 = bounds[rank]

    # ---------------- device-resident step: inputs already in HBM -----------------
    z_all = [torch.from_numpy(W.make_latents(P, 512, 50 + i)[s0:e0]).float().cuda() for i in range(4)]

    pending = []          # (work handle, gathered tensor) of the previous step's all-gather

    def step_device(i):
        """One generation: evaluate the local shard, then the ONE collective of the path (all-gather of F).  The
        gather only feeds the host GA, so it is issued asynchronously on NCCL's stream and waited for when the NEXT
        step has been enqueued: generation k+1's kernels do not queue behind generation k's collective."""
        neg_sim, hinge = eng.evaluate_device(z_all[i % 4], seed=1 + i, first_group=s0 // args.batch)
        local = torch.stack([neg_sim, hinge], 0) if hinge is not None else neg_sim[None]
        if world > 1:
            out = torch.empty((world,) + tuple(local.shape), device="cuda")
            work = tdist.all_gather_into_tensor(out, local.contiguous(), async_op=True)
            while pending:
                w, _ = pending.pop()
                w.wait()                               # stream-level wait for the previous generation's gather
            pending.append((work, out))
            return out
        return local

    def drain():
        while pending:
            pending.pop()[0].wait()

    def timed(step_fn, n_steps, first):
        """EXACTLY n_steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps
        (outside the events).  Returns per-step ms (max over ranks)."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        barrier()
        for i in range(n_steps):
            flush.zero_()
            ev[i][0].record(stream)
            step_fn(first + i)
            ev[i][1].record(stream)
        drain()
        barrier()
        ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device="cuda")
        timed.per_rank = None
        if world > 1:
            # the contract: time K steps, take the MAX over ranks of that time (not the mean of per-step maxima,
            # which would add every step's slowest rank together)
            allr = torch.empty(world, n_steps, dtype=torch.float64, device="cuda")
            tdist.all_gather_into_tensor(allr, ms)
            timed.per_rank = allr.mean(1).cpu().numpy().tolist()      # skew between ranks vs collective latency
            slowest = int(torch.argmax(allr.sum(1)).item())
            ms = allr[slowest]
        return ms.cpu().numpy()

    for i in range(W_):
        step_device(i)
    drain()
    barrier()

    # ---------------- N > 1: the gathered F equals what one GPU computes (checked on hardware) ----------------
    parity_check = None
    if world > 1:
        full = step_device(1000)
        drain()
        torch.cuda.synchronize()
        peer = (rank + 1) % world                       # every rank re-evaluates its neighbour's shard
        ps, pe = bounds[peer]
        zp = torch.from_numpy(W.make_latents(P, 512, 50 + 1000 % 4)[ps:pe]).float().cuda()
        n2, h2 = eng.evaluate_device(zp, seed=1 + 1000, first_group=ps // args.batch)
        mine = torch.stack([n2, h2], 0) if h2 is not None else n2[None]
        ok = torch.tensor([int(torch.equal(full[peer], mine))], device="cuda")
        tdist.all_reduce(ok, op=tdist.ReduceOp.MIN)
        parity_check = dict(what="after an NCCL all-gather of a full generation, every rank re-evaluates its "
                                 "neighbour rank's shard locally (same seed, first_group) and compares with the gathered rows",
                            bitwise_equal_on_all_ranks=bool(ok.item()), ranks=world, rows_checked_per_rank=pe - ps)
        if not ok.item():
            raise SystemExit("bench.py: gathered fitness differs from the single-GPU evaluation of the same shard")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = eng.launch_count
    ms_dev = timed(step_device, args.steps, W_)
    per_rank_ms = timed.per_rank
    launches = eng.launch_count - launches0

    # ---------------- e2e: host f64 population -> host F through the plugin API ----------------
    xs = [W.make_latents(P, 512, 80 + i) for i in range(4)]
    pinned = [torch.from_numpy(x).pin_memory().numpy() for x in xs]

    def step_e2e(i):
        out = {}
        problem._evaluate(pinned[i % 4], out)
        return out

    for i in range(2):
        step_e2e(i)
    e2e_wall = []
    barrier()
    for i in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        if world > 1:
            tdist.barrier()
        t0 = time.perf_counter()
        out = step_e2e(i)
        torch.cuda.synchronize()
        e2e_wall.append((time.perf_counter() - t0) * 1e3)
    e2e_ms = torch.tensor(e2e_wall, dtype=torch.float64, device="cuda")
    if world > 1:
        tdist.all_reduce(e2e_ms, op=tdist.ReduceOp.MAX)
    e2e_ms = e2e_ms.cpu().numpy()
    if rank == 0:
        sampler.stop()

    # ---------------- roofline of the dominant kernel (conv_tc), measured live ----------------
    roofline = None
    eng.set_debug(capture=False, timing=True)
    for i in range(2):
        eng.evaluate_device(z_all[i % 4], seed=100 + i)
    torch.cuda.synchronize()
    bd = eng.conv_breakdown()
    eng.set_debug(capture=False, timing=False)
    peaks = measured_peaks()
    if bd:
        conv_ms = sum(m for m, _ in bd)
        conv_flops = sum(f for _, f in bd)
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12
        peak = peaks["tflops_sustained"] or peaks["tflops_burst"]
        traffic = None
        tpath = os.path.join(REPO, "profiles", "conv_tc_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("bytes_per_launch")
        # exemplars: the most tensor-bound and the most HBM-bound launch of the step, each against its own roof
        names = conv_launch_names(use_d)
        detail = {}
        for key, bound in (("G8", "tensor"), ("G16", "hbm")):
            if key in names:
                i = names.index(key)
                if i < len(bd) and bd[i][0] > 0:
                    m, f = bd[i]
                    if bound == "tensor":
                        a = f / (m * 1e-3) / 1e12
                        detail[key] = dict(layer="G 3x3 512->512 @64^2", bound="tensor", achieved=a, unit="TFLOP/s",
                                           peak=peak, frac=a / peak, ms=m)
                    else:
                        # algorithmic bytes: fp16 NHWC input read once + one float4 toRGB partial per pixel written
                        byts = P_local * 1024 * 1024 * (32 * 2 + 16)
                        a = byts / (m * 1e-3) / 1e9
                        detail[key] = dict(layer="G 3x3 32->32 @1024^2 (+fused toRGB)", bound="hbm", achieved=a,
                                           unit="GB/s", peak=peaks["hbm_gbs"], frac=a / peaks["hbm_gbs"], ms=m)
        roofline = dict(bound="tensor", kernel="conv_tc_kernel (all tcgen05 conv/GEMM launches of one step)",
                        achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, traffic=traffic,
                        peak_source=peaks["source"] + ", sustained bf16 figure (kernel timed inside a long step)",
                        launches_per_step=len(bd), kernel_ms_per_step=conv_ms,
                        share_of_step=conv_ms / float(ms_dev.mean()),
                        algorithmic_gflop_per_step=conv_flops / 1e9, exemplars=detail)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = best_cpu_threads()
        cpu_oracle_step(args.cpu_sample, args.batch, use_d, 3, threads)            # warm
        ts = [cpu_oracle_step(args.cpu_sample, args.batch, use_d, 4 + i, threads) for i in range(2)]
        v = args.cpu_sample / statistics.median(ts)
        cpu_baseline = dict(value=v, unit=UNIT, cores=threads, kind=cpu_kind(),
                            sample=f"{args.cpu_sample} candidates (one minibatch) x 2 timed repetitions, median; "
                                   f"problem.py:14-29 over {'the reference modules (oracle/_ref)' if cpu_kind() == 'reference' else 'the oracle port'} "
                                   f"on torch CPU; {threads} of {os.cpu_count()} host "
                                   "threads (more threads are slower for a 4-candidate minibatch, see best_cpu_threads)")

    if rank == 0:
        ms = float(ms_dev.mean())
        e2e = float(e2e_ms.mean())
        line = dict(
            metric=METRIC, value=P / (ms / 1e3), unit=UNIT, n_gpus=world, steps=args.steps, warmup=W_,
            ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="fp16 operands, fp32 accumulate (tcgen05 kind::f16)", data="synthetic",
            config=dict(workload=f"{cfg_name} ffhq-config-f 1024^2 + CLIP ViT-B/32", population=P,
                        population_per_gpu=P_local, batch_size=args.batch, parallelism=f"population-sharded dp{world}",
                        weights="seeded random (no checkpoints offline)", noise="device Philox, fresh per minibatch group",
                        l2="per-step working set (>8 GB activations) >> 126 MB L2, plus a 256 MB flush between timed steps"),
            clocks=sampler.summary() if rank == 0 else None,
            e2e=dict(value=P / (e2e / 1e3), unit=UNIT, ms_per_step=e2e,
                     h2d_bytes_per_step=int(P_local * 512 * 8), d2h_bytes_per_step=int(P_local * (2 if use_d else 1) * 4),
                     api="GenerationProblem._evaluate(x: float64 ndarray) -> out['F'] (glass_evaluate_host)"),
            gpu_launches=int(launches),
            flags=dict(engine_flags=args.flags, debug_build=False, glass_debug_env=[]),
            parity_check=parity_check,
            per_rank_ms_per_step=per_rank_ms,
            gflop_per_candidate=GFLOP_PER_CAND[args.variant],
            roofline=roofline, cpu_baseline=cpu_baseline,
        )
        emit(line)
    if world > 1:
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
