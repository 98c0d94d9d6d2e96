"""Population sharding over ranks.

Candidates are independent through G and CLIP, and coupled only inside one
reference minibatch (shared noise draw, modules.py:426-452; MinibatchStd
groups, modules.py:726).  So the population is cut into contiguous blocks that
are whole multiples of ``batch_size``, each rank evaluates its block, and ONE
all-gather of the [P_local, n_obj] fp32 fitnesses per generation rebuilds F
on every rank — 4 KB at P=512.  No other collective is on the data path.
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


def _world():
    if tdist.is_available() and tdist.is_initialized():
        return tdist.get_rank(), tdist.get_world_size()
    return 0, 1


def all_gather_fitness(local: np.ndarray, bounds, device: Optional[torch.device] = None) -> np.ndarray:
    """All-gather of per-rank [P_r, k] fp32 blocks (padded to the largest shard)."""
    rank, world = _world()
    if world == 1:
        return local
    k = local.shape[1]
    longest = max(e - s for s, e in bounds)
    backend = tdist.get_backend()
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if backend == "nccl" else torch.device("cpu"))
    send = torch.zeros(longest, k, dtype=torch.float32, device=dev)
    if local.shape[0]:
        send[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float32)).to(dev)
    recv = torch.empty(world * longest, k, dtype=torch.float32, device=dev)
    tdist.all_gather_into_tensor(recv, send)
    recv = recv.reshape(world, longest, k).cpu().numpy()
    return np.concatenate([recv[r, : e - s] for r, (s, e) in enumerate(bounds)], axis=0)


def sharded_evaluate(x: np.ndarray, batch_size: int,
                     evaluate_local: Callable[[np.ndarray, int], Tuple[np.ndarray, Optional[np.ndarray]]]):
    """Evaluate ``x`` [P, n_var] over all ranks.  ``evaluate_local(x_shard, first_group)``
    returns (neg_sim[P_r], hinge[P_r] | None); ``first_group`` is the global index of the
    shard's first minibatch (so that seeded noise is identical to the single-rank run)."""
    rank, world = _world()
    if world == 1:
        return evaluate_local(x, 0)
    bounds = shard_bounds(x.shape[0], batch_size, world)
    s, e = bounds[rank]
    if e > s:
        neg_sim, hinge = evaluate_local(x[s:e], s // batch_size)
        cols = [neg_sim] + ([hinge] if hinge is not None else [])
        local = np.stack(cols, axis=1).astype(np.float32)
        have_hinge = hinge is not None
    else:
        local, have_hinge = None, None
    # ranks with empty shards need the column count
    ncol = torch.tensor([0 if local is None else local.shape[1]], dtype=torch.int64)
    if tdist.get_backend() == "nccl":
        ncol = ncol.cuda()
    tdist.all_reduce(ncol, op=tdist.ReduceOp.MAX)
    k = int(ncol.item())
    if local is None:
        local = np.zeros((0, k), dtype=np.float32)
    full = all_gather_fitness(local, bounds)
    return full[:, 0], (full[:, 1] if k == 2 else None)
