"""CPU oracle for the DiffUTE sampling hot path (TEST INFRASTRUCTURE ONLY).

This package is a plain-PyTorch fp32 restatement of the third-party diffusers
modules that chenhaoxing/DiffUTE calls on its sampling path
(reference call sites: app.ipynb:772-819, train_diffute_v1.py:875-913):

    UNet2DConditionModel.forward      -> oracle.unet.UNetOracle
    AutoencoderKL.encode/decode       -> oracle.vae.VAEOracle
    DDIMScheduler / DDPMScheduler     -> oracle.schedulers
    text_editing's cv2 / PIL / numpy glue (app.ipynb:663-771, :821-841) -> oracle.glue  (pinned to cv2 4.13 / PIL)
    the ancestral step's Gaussian noise stream                          -> oracle.philox (Random123 known answers)

diffusers itself is not vendored under /root/reference, is not pinned by the
reference (requirements.txt:1-8; floor "0.15.0.dev0" at train_diffute_v1.py:63)
and is not installable offline, so the algorithm is restated from its published
architecture.  PARITY PINS: exact parameter totals
(UNet 865,925,124; VAE 83,653,863) and six upstream scheduler known-answer
constants (tests/test_oracle_*.py).  UNet/VAE *outputs* have no golden vectors
in the reference (it has no tests at all): "parity unpinned" against the
reference itself.  Independent evidence instead: tests/test_oracle_crosscheck.py
checks this package primitive by primitive, block by block and end to end
against tests/refmath.py, a second restatement written in a different form
(numpy float64, flat functional walk, explicit im2col/softmax/normalisation);
they agree to <= 5e-6 (fp32 vs fp64 noise).  The same file holds a test that
activates when the real package is importable and compares against diffusers'
own UNet2DConditionModel / AutoencoderKL / DDIMScheduler on the same weights:

    python -m pytest tests/test_oracle_crosscheck.py -k real_diffusers

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (diffute_b200/) never does.
"""
from .unet import UNetOracle, SD2_INPAINT_UNET_CONFIG  # noqa: F401
from .vae import VAEOracle, SD2_VAE_CONFIG, DiagonalGaussian  # noqa: F401
from .schedulers import DDIMOracle, DDPMOracle, SD2_SCHEDULER_CONFIG  # noqa: F401
from .sampling import sample_loop  # noqa: F401
