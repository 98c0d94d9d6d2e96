"""Mirror of the reference's utils.py helpers that sit on the path.

``biggan_norm`` / ``biggan_denorm`` (utils.py:14-21) are fused into the CUDA
kernels (k_rgb_combine writes clip((y+1)/2); k_from_rgb applies x*2-1); the
torch versions here exist for callers that hold images as tensors.
"""
import contextlib

import torch


try:                                    # NVTX is optional instrumentation: never let it break the path
    from torch.cuda import nvtx as _nvtx
    _nvtx.range_push("glass.import")
    _nvtx.range_pop()
except Exception:                       # pragma: no cover - depends on the torch build
    _nvtx = None


@contextlib.contextmanager
def nvtx_range(name: str):
    """NVTX range around a host-side phase (SURVEY.md §5 tracing): shows up in nsys / ncu timelines as
    ``glass.<phase>``; a no-op stub call when no profiler is attached."""
    if _nvtx is None:
        yield
        return
    _nvtx.range_push(name)
    try:
        yield
    finally:
        _nvtx.range_pop()


def biggan_norm(images):
    return ((images + 1) / 2.0).clip(0, 1)


def biggan_denorm(images):
    return images * 2 - 1


def freeze_model(model):
    for p in model.parameters():
        p.requires_grad = False


def save_grid(images, path):
    import torchvision
    torchvision.utils.save_image(torchvision.utils.make_grid(images), path)
