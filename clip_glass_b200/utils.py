"""Mirror of the reference's utils.py helpers that sit on the path.

``biggan_norm`` / ``biggan_denorm`` (utils.py:14-21) are fused into the CUDA
kernels (k_rgb_combine writes clip((y+1)/2); k_from_rgb applies x*2-1); the
torch versions here exist for callers that hold images as tensors.
"""
import torch


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
