"""GPU parity of the reference's pre-/post-processing kernels (dfu_glue_preprocess / dfu_glue_composite, through the
C-ABI) against oracle/glue.py, which tests/test_glue_oracle.py pins to cv2 / PIL: bit-exact for the uint8 path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from diffute_b200 import glue
    from oracle import glue as G
    return glue, G


CASES = [  # (h, w, bbox, window or None)
    (600, 800, (300, 250, 420, 290), None),          # 6 * 40 = 240 < 256: 256 window, 2x upscale
    (300, 260, (10, 20, 200, 60), None),             # window = 256 at the origin
    (1100, 1300, (100, 100, 900, 330), None),        # 6 * 230 >= 1000: window = short side, clipped
    (512, 512, (400, 450, 500, 500), None),          # 384 window
    (1400, 1500, (200, 300, 700, 400), (150, 100, 1024)),   # exactly 2x decimation: OpenCV's area fast path
    (700, 900, (100, 100, 300, 150), (50, 60, 512)),        # identity resize
    (640, 480, (380, 500, 470, 560), (300, 420, 384)),      # window clipped at both borders: 180 x 220 crop, non-square
    (130, 97, (0, 0, 96, 129), (0, 0, 97)),                 # box = whole image (inclusive corners), tiny photograph
    (520, 520, (3, 3, 4, 4), (0, 0, 2)),                    # 2 x 2 window: every destination pixel clamps
]


@pytest.mark.parametrize("h,w,bbox,window", CASES)
def test_preprocess_bit_exact(mods, h, w, bbox, window):
    glue, G = mods
    rng = np.random.default_rng(h * 7 + w)
    image = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    win = window if window is not None else G.crop_window(bbox, h, w, np.random.RandomState(0))
    assert win == (window if window is not None else glue.crop_window(bbox, h, w, np.random.RandomState(0)))
    pre = glue.preprocess(image, bbox, window=win)
    torch.cuda.synchronize()
    img_c, mim_c, msk_c = G.preprocess(image, bbox, win)
    assert np.array_equal(pre.image[0].cpu().numpy(), img_c)
    assert np.array_equal(pre.masked_image[0].cpu().numpy(), mim_c)
    assert np.array_equal(pre.mask[0].cpu().numpy(), msk_c.astype(np.float32))
    assert np.array_equal(pre.mask_latents[0, 0].cpu().numpy(), msk_c[0, ::8, ::8].astype(np.float32))
    assert set(np.unique(msk_c).tolist()) <= {0, 1}


@pytest.mark.parametrize("h,w,bbox,window", CASES)
@pytest.mark.parametrize("wrap", [False, True])
def test_composite_matches_oracle(mods, h, w, bbox, window, wrap):
    glue, G = mods
    rng = np.random.default_rng(h * 11 + w)
    image = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    win = window if window is not None else G.crop_window(bbox, h, w, np.random.RandomState(0))
    decoded = (rng.random((3, 512, 512), dtype=np.float32) * 2.4 - 1.2).astype(np.float32)   # leaves [-1, 1]: clamp / wrap
    pre = glue.preprocess(image, bbox, window=win)
    got = glue.composite(torch.from_numpy(decoded).cuda()[None], pre, wrap=wrap).cpu().numpy()
    ref = G.composite(decoded, image, bbox, win, wrap=wrap)
    assert got.shape == ref.shape == (h, w, 3) and got.dtype == np.uint8
    x1, y1, x2, y2 = bbox
    outside = np.ones((h, w), bool)
    outside[y1:y2, x1:x2] = False
    assert np.array_equal(got[outside], image[outside])          # nothing but the text box changes
    # the float resize is a true fma on the device and an fma emulated in float64 in the oracle: identical except, at
    # most, a rounding tie once in 2^29 pixels
    assert (got != ref).mean() <= 1e-6, (got != ref).mean()


def test_rejects_bad_geometry(mods):
    glue, G = mods
    from diffute_b200._lib import DfuError
    image = np.zeros((64, 64, 3), np.uint8)
    with pytest.raises(ValueError):
        glue.preprocess(image, (1, 1, 5, 5), window=(70, 0, 32))
    with pytest.raises(ValueError):
        glue.preprocess(image.astype(np.float32), (1, 1, 5, 5))
    from diffute_b200 import ops
    with pytest.raises(DfuError):
        ops.glue_preprocess(torch.zeros((64, 64, 3), dtype=torch.uint8, device="cuda"), (40, 0, 32, 32), (1, 1, 5, 5))


def test_text_editing_end_to_end(mods):
    """The reference's text_editing (app.ipynb:653-856) on the engine: photograph -> window -> 512 x 512 tensors ->
    4 DDIM steps -> decode -> composite, against the same chain on the CPU oracle."""
    glue, G = mods
    from diffute_b200 import arch, synthetic
    from diffute_b200.pipeline import DiffUTEPipeline
    from oracle import DDIMOracle, UNetOracle, VAEOracle, sample_loop
    usd = synthetic.make_state_dict(arch.unet_param_shapes())
    vsd = synthetic.make_state_dict(arch.vae_param_shapes())
    pipe = DiffUTEPipeline.from_synthetic("fp16x2", "fp16x2", state_dicts=(usd, vsd))
    rng = np.random.default_rng(3)
    h, w, bbox = 420, 560, (200, 180, 330, 215)
    image = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    inp = synthetic.make_inputs(1, 512, 512)
    edited, mask255 = glue.text_editing(pipe, None, image, 4, *bbox, glyph_embeds=inp["glyph_embeds"],
                                        latents=inp["latents"], posterior_noise=inp["posterior_noise"])
    assert edited.shape == (h, w, 3) and edited.dtype == np.uint8 and mask255.max() == 255
    assert np.array_equal(mask255 // 255, G.generate_mask(h, w, bbox))
    uo, vo = UNetOracle(), VAEOracle()
    uo.load_state_dict(usd)
    vo.load_state_dict(vsd)
    win = G.crop_window(bbox, h, w)
    _, mim_c, msk_c = G.preprocess(image, bbox, win)
    dec = sample_loop(uo, vo, DDIMOracle(), torch.from_numpy(mim_c)[None], torch.from_numpy(msk_c.astype(np.float32))[None],
                      inp["glyph_embeds"], inp["latents"], 4, posterior_noise=inp["posterior_noise"])
    ref = G.composite(dec[0].numpy(), image, bbox, win)
    diff = np.abs(edited.astype(int) - ref.astype(int))
    print(f"text_editing vs oracle chain: pixels differing {(diff > 0).mean():.2e}, max |diff| {diff.max()}")
    assert diff.max() <= 1 and (diff > 0).mean() < 2e-3   # decoded RGB agrees to ~1e-5: only rounding ties move


@pytest.mark.parametrize("h,w", [(60, 200), (60, 680), (60, 1240), (384, 384), (500, 384), (384, 200), (7, 5)])
def test_glyph_processor_bit_exact(mods, h, w):
    """TrOCRProcessor's image side (app.ipynb:773-774) on the GPU == the numpy restatement of PIL's resample + the
    PIL-backend ViTImageProcessor arithmetic (pinned to Pillow / transformers in tests/test_glue_oracle.py)."""
    glue, G = mods
    from diffute_b200.glyph_encoder import TrOCRGlyphProcessor
    rng = np.random.default_rng(h * 13 + w)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(2)]
    pv = TrOCRGlyphProcessor()(images=imgs, return_tensors="pt").pixel_values
    assert pv.shape == (2, 3, 384, 384) and pv.dtype == torch.float32
    for i, im in enumerate(imgs):
        assert np.array_equal(pv[i].cpu().numpy(), G.vit_pixel_values(im)), (h, w, i)


def test_round_trip_properties(mods):
    """Size-independent properties at a full-size photograph: (1) with a 512-pixel window both resizes are identities, so
    pre-processing followed by compositing of the untouched crop returns the photograph bit for bit; (2) at other scales
    the text box is reproduced to within the two bilinear resamplings of a smooth image; (3) the noise stream is a pure
    function of (seed, step, element): a prefix of a longer draw equals the shorter draw, steps differ."""
    glue, G = mods
    from diffute_b200 import ops
    rng = np.random.default_rng(9)
    h, w = 2160, 3840
    image = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    bbox = (1700, 1000, 2100, 1085)                      # 6 * 85 = 510 < 512 -> 512 window: identity resize
    pre = glue.preprocess(image, bbox)
    assert pre.window[2] == 512
    out = glue.composite(pre.image, pre).cpu().numpy()   # "decoded" = the normalised crop itself
    assert np.array_equal(out, image)
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = np.stack([(127 + 100 * np.sin(xx / 97.0) * np.cos(yy / 53.0)), (xx * 255.0 / w), (yy * 255.0 / h)], -1)
    smooth = np.clip(np.rint(smooth), 0, 255).astype(np.uint8)
    box2 = (900, 700, 1500, 830)                         # 6 * 130 = 780 < 784 -> 784 window: down to 512 and back up
    pre2 = glue.preprocess(smooth, box2)
    assert pre2.window[2] == 784
    out2 = glue.composite(pre2.image, pre2).cpu().numpy().astype(int)
    x1, y1, x2, y2 = box2
    assert np.abs(out2[y1:y2, x1:x2] - smooth[y1:y2, x1:x2].astype(int)).max() <= 2
    outside = np.ones((h, w), bool)
    outside[y1:y2, x1:x2] = False
    assert np.array_equal(out2[outside], smooth[outside].astype(int))
    a = ops.philox_normal(77, 3, 1 << 20)
    b = ops.philox_normal(77, 3, 1000)
    assert torch.equal(a[:1000], b) and not torch.equal(b, ops.philox_normal(77, 4, 1000))
    assert abs(a.mean().item()) < 5e-3 and abs(a.std().item() - 1.0) < 5e-3


def test_request_queue_batches_equal_single_requests(mods):
    """Five requests with different photographs / boxes / glyph embeddings served as one batch of 3 and one of 2
    (text_editing_batch) against the same requests served one by one: the edited photographs agree to within one grey
    level on a handful of rounding ties (different GEMM tilings at batch 3 vs 1 move the decoded image by ~1e-5)."""
    glue, G = mods
    from diffute_b200 import synthetic
    from diffute_b200.pipeline import DiffUTEPipeline
    pipe = DiffUTEPipeline.from_synthetic("fp16x2", "fp16x2")
    rng = np.random.default_rng(12)
    reqs = []
    for i, (h, w, bbox) in enumerate([(420, 560, (200, 180, 330, 215)), (600, 800, (300, 250, 420, 290)),
                                      (512, 512, (100, 300, 400, 360)), (700, 650, (50, 60, 250, 100)),
                                      (380, 900, (500, 200, 760, 240))]):
        inp = synthetic.make_inputs(1, 512, 512, seed=40 + i)
        reqs.append(dict(instance_image=rng.integers(0, 256, (h, w, 3), dtype=np.uint8), bbox=bbox,
                         glyph_embeds=inp["glyph_embeds"], latents=inp["latents"]))
    kw = dict(sample_posterior=False)
    batched = glue.text_editing_batch(pipe, reqs, 3, max_batch=3, **kw)
    assert len(batched) == 5
    for r, (img_b, mask_b) in zip(reqs, batched):
        img_1, mask_1 = glue.text_editing(pipe, None, r["instance_image"], 3, *r["bbox"], glyph_embeds=r["glyph_embeds"],
                                          latents=r["latents"], **kw)
        assert img_b.shape == r["instance_image"].shape and np.array_equal(mask_b, mask_1)
        d = np.abs(img_b.astype(int) - img_1.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3, (d.max(), (d > 0).mean())
