"""GlassEngine: Python handle on one C-ABI engine (one GPU).

Holds the packed weights of StyleGAN2 G (+ D) and the CLIP visual tower on the
device and exposes the calls that the reference's ``Generator`` façade makes
(generator.py:29-60).  PyTorch is used only to own device buffers and streams;
all arithmetic happens in libclipglass_b200.so.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import packing
from ._lib import GlassConfig, GlassError, GlassNoise, check, load_library
from .weights import ClipSpec, GanSpec


def flatten_noise(noise: Sequence[Sequence[torch.Tensor]]) -> np.ndarray:
    """[groups][layers] of [1,1,H,W] -> fp32 [groups, sum H*W] (the layout
    glass_noise documents)."""
    rows = [np.concatenate([np.asarray(t, dtype=np.float32).reshape(-1) for t in grp]) for grp in noise]
    return np.ascontiguousarray(np.stack(rows))


class GlassEngine:
    def __init__(self, gan: GanSpec, clip: ClipSpec, g_sd: Dict[str, torch.Tensor],
                 d_sd: Optional[Dict[str, torch.Tensor]], clip_sd: Dict[str, torch.Tensor],
                 batch_size: int, max_population: int, device: int = 0, conv_impl: int = 0, flags: int = 0):
        self.lib = load_library()
        self.gan, self.clip = gan, clip
        self.batch_size = batch_size
        self.max_population = max_population
        self.device = device
        self.use_discriminator = d_sd is not None
        cfg = GlassConfig()
        cfg.num_blocks = gan.num_blocks
        for i, c in enumerate(list(gan.channels)[::-1]):     # 4x4 first
            cfg.channels[i] = c
        cfg.latent_size = gan.latent_size
        cfg.mapping_layers = gan.mapping_layers
        cfg.batch_size = batch_size
        cfg.mbstd_group_size = gan.mbstd_group_size
        cfg.use_discriminator = int(self.use_discriminator)
        cfg.clip_width, cfg.clip_layers = clip.width, clip.layers
        cfg.clip_patch, cfg.clip_resolution, cfg.clip_embed_dim = clip.patch, clip.resolution, clip.embed_dim
        cfg.max_population = max_population
        cfg.device = device
        cfg.conv_impl = conv_impl
        cfg.flags = flags
        self._h = ctypes.c_void_p()
        check(self.lib.glass_create(ctypes.byref(cfg), ctypes.byref(self._h)))
        packed = {}
        packed.update(packing.pack_generator(g_sd, gan))
        if d_sd is not None:
            packed.update(packing.pack_discriminator(d_sd, gan))
        packed.update(packing.pack_clip_visual(clip_sd, clip))
        for name, arr in packed.items():
            arr = np.ascontiguousarray(arr)
            check(self.lib.glass_set_tensor(self._h, name.encode(), arr.ctypes.data, arr.nbytes))
        check(self.lib.glass_finalize(self._h))
        self.noise_per_group = sum(s * s for s in gan.noise_shapes())

    # -- lifetime --------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.glass_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_text_features(self, text_features) -> None:
        """generator.py:23-24: the cached ``CLIP.encode_text`` result [1,E]."""
        t = np.ascontiguousarray(np.asarray(torch.as_tensor(text_features).float().cpu()).reshape(-1), dtype=np.float32)
        check(self.lib.glass_set_text_features(self._h, t.ctypes.data, t.size))

    def set_batch_size(self, batch_size: int) -> None:
        """Noise / MinibatchStd scope of later calls (the reference's ``minibatch`` argument)."""
        if batch_size != self.batch_size:
            check(self.lib.glass_set_batch_size(self._h, int(batch_size)))
            self.batch_size = int(batch_size)

    # -- helpers -----------------------------------------------------------
    def _noise_arg(self, noise, seed, pop, keep, first_group=0):
        nz = GlassNoise()
        nz.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        nz.first_group = int(first_group)
        nz.noise = None
        nz.noise_on_device = 0
        if noise is not None:
            if isinstance(noise, torch.Tensor) and noise.is_cuda:
                assert noise.dtype == torch.float32 and noise.is_contiguous()
                assert noise.numel() == (pop // self.batch_size) * self.noise_per_group
                nz.noise = noise.data_ptr()
                nz.noise_on_device = 1
                keep.append(noise)
            else:
                arr = noise if isinstance(noise, np.ndarray) else flatten_noise(noise)
                arr = np.ascontiguousarray(arr, dtype=np.float32)
                assert arr.size == (pop // self.batch_size) * self.noise_per_group, (arr.shape, pop)
                nz.noise = arr.ctypes.data
                keep.append(arr)
        return nz

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- the hot path ------------------------------------------------------
    def evaluate(self, x: np.ndarray, noise=None, seed: int = 0, first_group: int = 0):
        """problem.py:14-29 through ``glass_evaluate_host``: host f64 in, host fp32 out."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        pop = x.shape[0]
        keep: list = []
        nz = self._noise_arg(noise, seed, pop, keep, first_group)
        neg_sim = np.empty(pop, dtype=np.float32)
        hinge = np.empty(pop, dtype=np.float32) if self.use_discriminator else None
        check(self.lib.glass_evaluate_host(
            self._h, x.ctypes.data, pop, ctypes.byref(nz), neg_sim.ctypes.data,
            hinge.ctypes.data if hinge is not None else None, self._stream()))
        return neg_sim, hinge

    def evaluate_device(self, z: torch.Tensor, noise=None, seed: int = 0, first_group: int = 0):
        """Device-resident variant (z fp32 cuda [P,L]); returns cuda tensors, asynchronous."""
        assert z.is_cuda and z.dtype == torch.float32 and z.is_contiguous()
        pop = z.shape[0]
        keep: list = []
        nz = self._noise_arg(noise, seed, pop, keep, first_group)
        neg_sim = torch.empty(pop, dtype=torch.float32, device=z.device)
        hinge = torch.empty(pop, dtype=torch.float32, device=z.device) if self.use_discriminator else None
        check(self.lib.glass_evaluate_device(
            self._h, z.data_ptr(), pop, ctypes.byref(nz), neg_sim.data_ptr(),
            hinge.data_ptr() if hinge is not None else None, self._stream()))
        if keep and not (isinstance(keep[0], torch.Tensor)):
            torch.cuda.current_stream(self.device).synchronize()   # host noise buffer must outlive the copy
        return neg_sim, hinge

    def generate(self, z: torch.Tensor, noise=None, seed: int = 0) -> torch.Tensor:
        """generator.py:29-34: images fp32 [P,3,R,R] in [0,1] on the device."""
        assert z.is_cuda and z.dtype == torch.float32 and z.is_contiguous()
        pop = z.shape[0]
        keep: list = []
        nz = self._noise_arg(noise, seed, pop, keep)
        R = self.gan.resolution
        images = torch.empty(pop, 3, R, R, dtype=torch.float32, device=z.device)
        check(self.lib.glass_generate(self._h, z.data_ptr(), pop, ctypes.byref(nz), images.data_ptr(), self._stream()))
        if keep and not (isinstance(keep[0], torch.Tensor)):
            torch.cuda.current_stream(self.device).synchronize()
        return images

    def clip_similarity(self, images: torch.Tensor) -> torch.Tensor:
        assert images.is_cuda and images.dtype == torch.float32 and images.is_contiguous()
        sim = torch.empty(images.shape[0], dtype=torch.float32, device=images.device)
        check(self.lib.glass_clip_similarity(self._h, images.data_ptr(), images.shape[0], sim.data_ptr(), self._stream()))
        return sim

    def discriminate(self, images: torch.Tensor) -> torch.Tensor:
        assert images.is_cuda and images.dtype == torch.float32 and images.is_contiguous()
        out = torch.empty(images.shape[0], dtype=torch.float32, device=images.device)
        check(self.lib.glass_discriminate(self._h, images.data_ptr(), images.shape[0], out.data_ptr(), self._stream()))
        return out.unsqueeze(1)

    # -- image output path -------------------------------------------------
    def last_images(self, rows) -> torch.Tensor:
        """Images [n,3,R,R] (fp32, [0,1], device) of rows ``rows`` of the last fused evaluation — the images that were
        scored (run.py:29-51 save_callback would render them again)."""
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        R = self.gan.resolution
        out = torch.empty(len(rows), 3, R, R, dtype=torch.float32, device=torch.device("cuda", self.device))
        check(self.lib.glass_last_images_gather(self._h, rows.ctypes.data, len(rows), out.data_ptr(), self._stream()))
        return out

    def image_grid_u8(self, images: torch.Tensor, nrow: int = 8, padding: int = 2) -> np.ndarray:
        """utils.py:5-7: make_grid + save_image's uint8 conversion, on the device; returns the HWC uint8 ndarray."""
        assert images.is_cuda and images.dtype == torch.float32 and images.is_contiguous() and images.shape[1] == 3
        n, _, R, R2 = images.shape
        assert R == R2
        xmaps = min(nrow, n)
        ymaps = (n + xmaps - 1) // xmaps
        Hg, Wg = (R + padding) * ymaps + padding, (R + padding) * xmaps + padding
        out = torch.empty(Hg, Wg, 3, dtype=torch.uint8, device=images.device)
        check(self.lib.glass_image_grid_u8(self._h, images.data_ptr(), n, R, nrow, padding, out.data_ptr(), self._stream()))
        return out.cpu().numpy()

    # -- introspection -----------------------------------------------------
    @property
    def launch_count(self) -> int:
        return int(self.lib.glass_launch_count(self._h))

    def set_debug(self, capture: bool = False, timing: bool = False) -> None:
        check(self.lib.glass_set_debug(self._h, int(capture), int(timing)))

    def set_range_check(self, enable: bool = True) -> None:
        """Scan every G / D fp16 activation tensor of later calls (resets the counters; eager launches)."""
        check(self.lib.glass_set_range_check(self._h, int(enable)))

    def range_report(self) -> dict:
        """{'nonfinite', 'saturated', 'max_abs'} over the activations scanned since ``set_range_check``."""
        bad, sat, mx = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_float()
        check(self.lib.glass_range_report(self._h, ctypes.byref(bad), ctypes.byref(sat), ctypes.byref(mx)))
        return dict(nonfinite=int(bad.value), saturated=int(sat.value), max_abs=float(mx.value))

    def debug_read(self, name: str) -> np.ndarray:
        n = self.lib.glass_debug_read(self._h, name.encode(), None, 0)
        check(int(n))
        out = np.empty(int(n), dtype=np.float32)
        check(int(self.lib.glass_debug_read(self._h, name.encode(), out.ctypes.data, int(n))))
        return out

    def conv_breakdown(self):
        """[(ms, algorithmic_flops)] of the tensor-core launches of the last timed call."""
        cap = 512
        ms = np.zeros(cap, dtype=np.float32)
        fl = np.zeros(cap, dtype=np.float64)
        n = self.lib.glass_conv_breakdown(self._h, ms.ctypes.data, fl.ctypes.data, cap)
        check(int(n))
        return [(float(ms[i]), float(fl[i])) for i in range(int(n))]
