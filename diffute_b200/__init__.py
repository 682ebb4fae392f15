"""diffute_b200 — B200-native (sm_100a) engine for the DiffUTE sampling hot path.

Drop-in for the diffusers objects chenhaoxing/DiffUTE calls while sampling
(app.ipynb:545-553, 772-819): UNet2DConditionModel, AutoencoderKL,
DDIMScheduler / DDPMScheduler, plus a DiffUTEPipeline for the loop itself.
All device arithmetic is hand-written CUDA behind a C-ABI shared library
(include/diffute_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
