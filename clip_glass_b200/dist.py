"""Population sharding over ranks.

Candidates are independent through G and CLIP, and coupled only inside one
reference minibatch (shared noise draw, modules.py:426-452; MinibatchStd
groups, modules.py:726).  So the population is cut into contiguous blocks that
are whole multiples of ``batch_size``, each rank evaluates its block, and ONE
all-gather of the [n_obj, P_local] fp32 fitnesses per generation rebuilds F
on every rank — 4 KB at P=512.  No other collective is on the data path: the
column count comes from the config (``n_obj``), not from a second collective,
and with NCCL the gather runs straight from the engine's device outputs (one
D2H copy of the gathered F at the end instead of one per rank-local array).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as tdist


def shard_bounds(pop: int, batch_size: int, world: int) -> List[Tuple[int, int]]:
    """[start, end) per rank; every boundary is a multiple of batch_size; sizes differ by
    at most one minibatch (ranks beyond the number of minibatches get empty shards)."""
    assert pop % batch_size == 0, f"population {pop} is not a multiple of batch_size {batch_size}"
    groups = pop // batch_size
    base, extra = divmod(groups, world)
    bounds, g0 = [], 0
    for r in range(world):
        g1 = g0 + base + (1 if r < extra else 0)
        bounds.append((g0 * batch_size, g1 * batch_size))
        g0 = g1
    return bounds


def init_from_env(device: str = "cuda", timeout_s: float = 600.0) -> Tuple[int, int, int]:
    """``torchrun`` launch of the driver (SURVEY.md §5 failure detection): when WORLD_SIZE > 1 and no process group
    exists yet, create one — NCCL for a CUDA device (bound to LOCAL_RANK's GPU), gloo otherwise — with a collective
    timeout, so that a rank that died or hung makes the per-generation all-gather of F raise on the others instead of
    blocking the search forever (NCCL's asynchronous error handling tears the communicator down).  Returns
    (rank, world, local_rank); (0, 1, 0) for a plain single-process launch, in which nothing is initialised."""
    import datetime
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, 0
    rank, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not tdist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "1")
        timeout = datetime.timedelta(seconds=float(timeout_s))
        if str(device).startswith("cuda"):
            torch.cuda.set_device(local_rank)
            tdist.init_process_group("nccl", timeout=timeout, device_id=torch.device("cuda", local_rank))
        else:
            tdist.init_process_group("gloo", timeout=timeout)
    return rank, world, local_rank


def _world():
    if tdist.is_available() and tdist.is_initialized():
        return tdist.get_rank(), tdist.get_world_size()
    return 0, 1


def _gather_blocks(send: torch.Tensor, world: int) -> torch.Tensor:
    """The one collective: all ranks' [k, longest] blocks -> [world, k, longest]."""
    recv = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=send.device)
    if tdist.get_backend() == "gloo":
        # (all_gather_into_tensor is not implemented by every gloo build)
        tdist.all_gather(list(recv.unbind(0)), send)
    else:
        tdist.all_gather_into_tensor(recv, send)
    return recv


def all_gather_fitness(local, bounds, n_obj: int, device: Optional[torch.device] = None) -> np.ndarray:
    """All-gather of per-rank fitness blocks.  ``local``: [P_r, n_obj] ndarray, or a tuple of ``n_obj`` device
    (or host) tensors of length P_r.  Returns the full [P, n_obj] fp32 ndarray on every rank."""
    rank, world = _world()
    cols = [torch.as_tensor(c) for c in (local if isinstance(local, (tuple, list)) else np.asarray(local).T)]
    assert len(cols) == n_obj or (len(cols) == 0 and bounds[rank][0] == bounds[rank][1]), (len(cols), n_obj)
    if world == 1:
        return torch.stack([c.float().cpu() for c in cols], 1).numpy()
    longest = max(e - s for s, e in bounds)
    if device is None:
        device = (torch.device("cuda", torch.cuda.current_device()) if tdist.get_backend() == "nccl"
                  else torch.device("cpu"))
    send = torch.zeros(n_obj, longest, dtype=torch.float32, device=device)
    for j, c in enumerate(cols):
        if c.numel():
            send[j, :c.numel()].copy_(c.to(device=device, dtype=torch.float32), non_blocking=True)
    recv = _gather_blocks(send, world).cpu().numpy()                # [world, n_obj, longest]
    return np.concatenate([recv[r, :, : e - s].T for r, (s, e) in enumerate(bounds)], axis=0)


def sharded_evaluate(x: np.ndarray, batch_size: int, n_obj: int,
                     evaluate_local: Callable[[np.ndarray, int], Tuple]):
    """Evaluate ``x`` [P, n_var] over all ranks.  ``evaluate_local(x_shard, first_group)`` returns
    (neg_sim[P_r], hinge[P_r] | None) as host arrays or device tensors; ``first_group`` is the global index of
    the shard's first minibatch (so that seeded noise is identical to the single-rank run).  ``n_obj`` = 2 when
    the hinge column exists (config.problem_args["n_obj"] == 2 and config.use_discriminator), else 1."""
    rank, world = _world()
    if world == 1:
        neg_sim, hinge = evaluate_local(x, 0)
        as_np = lambda t: t.float().cpu().numpy() if isinstance(t, torch.Tensor) else t
        return as_np(neg_sim), (as_np(hinge) if hinge is not None else None)
    bounds = shard_bounds(x.shape[0], batch_size, world)
    s, e = bounds[rank]
    cols: list = []
    if e > s:
        neg_sim, hinge = evaluate_local(x[s:e], s // batch_size)
        cols = [neg_sim] + ([hinge] if n_obj == 2 else [])
        assert n_obj == 1 or hinge is not None, "n_obj == 2 needs the discriminator's hinge column"
    full = all_gather_fitness(tuple(cols), bounds, n_obj)
    return full[:, 0], (full[:, 1] if n_obj == 2 else None)
