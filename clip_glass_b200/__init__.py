"""clip-glass-b200: B200-native population fitness evaluation for CLIP-GLaSS.

The hot path (latents -> StyleGAN2 G -> CLIP ViT-B/32 -> cosine [-> StyleGAN2 D
hinge]) runs as hand-written sm_100a CUDA kernels behind a C-ABI shared
library (include/clipglass_b200.h); this package is the host-side mirror of
the reference's plugin surface (problem.py / generator.py / latent.py /
operators.py / config.py) on top of it.
"""
__version__ = "0.1.0"
