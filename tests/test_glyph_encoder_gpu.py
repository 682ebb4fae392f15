"""f3: the TrOCR ViT glyph encoder on the native kernels against the real third-party implementation the reference
calls (transformers' ViTModel — `VisionEncoderDecoderModel.from_pretrained('trocr-large-printed').encoder`,
app.ipynb:546-548, :773-776) on the CPU, same randomly initialised weights."""
import pytest
import torch


def _vit(cfg_over=None):
    from transformers import ViTConfig, ViTModel
    from diffute_b200.glyph_encoder import TROCR_LARGE_VIT_CONFIG
    c = dict(TROCR_LARGE_VIT_CONFIG)
    c.update(cfg_over or {})
    torch.manual_seed(7)
    m = ViTModel(ViTConfig(**c), add_pooling_layer=False).eval()
    with torch.no_grad():  # default init leaves LayerNorm at (1, 0) and cls/pos tiny: make every parameter matter
        for k, p in m.named_parameters():
            if "layernorm" in k:
                p.add_(0.1 * torch.randn_like(p))
            elif "cls_token" in k or "position_embeddings" in k:
                p.copy_(0.5 * torch.randn_like(p))
            elif k.endswith(".bias"):
                p.copy_(0.05 * torch.randn_like(p))
    return c, m


def test_vit_inventory_matches_transformers():
    from diffute_b200.glyph_encoder import vit_param_shapes
    for over in ({}, {"qkv_bias": True, "num_hidden_layers": 2}):
        c, m = _vit({"num_hidden_layers": 2, **over})
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == vit_param_shapes(c)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp16x2", 2e-4), ("fp16", 8e-3)])
def test_trocr_large_encoder_parity(precision, tol):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from diffute_b200.glyph_encoder import TrOCRGlyphEncoder
    c, m = _vit()
    g = torch.Generator().manual_seed(3)
    pv = torch.rand((2, 3, 384, 384), generator=g) * 2 - 1          # TrOCRProcessor output range (Normalize(0.5, 0.5))
    with torch.no_grad():
        ref = m(pv).last_hidden_state
    # the reference's checkpoint is a VisionEncoderDecoderModel: encoder keys carry an `encoder.` prefix
    sd = {"encoder." + k: v for k, v in m.state_dict().items()}
    sd["decoder.dummy"] = torch.zeros(1)
    enc = TrOCRGlyphEncoder(sd, c, precision=precision)
    out = enc(pv.cuda())
    assert out.last_hidden_state.shape == (2, 577, 1024) and out[0] is out.last_hidden_state
    err = ((out.last_hidden_state.cpu() - ref).abs().max() / ref.abs().max()).item()
    print(f"TrOCR-large ViT encoder {precision}: last_hidden_state maxrel {err:.3e}")
    assert err < tol, err
    again = enc(pv.cuda()).last_hidden_state      # static buffers: a second call is bit-identical
    assert torch.equal(again, out.last_hidden_state)


@pytest.mark.gpu
def test_pipeline_with_native_glyph_encoder():
    """DiffUTEPipeline(text=glyph images) through the native encoder equals passing its embeddings explicitly."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from diffute_b200 import synthetic
    from diffute_b200.glyph_encoder import TrOCRGlyphEncoder
    from diffute_b200.pipeline import DiffUTEPipeline
    c, m = _vit({"qkv_bias": True})
    enc = TrOCRGlyphEncoder(m.state_dict(), c, precision="fp16x2")

    class Proc:  # stands in for TrOCRProcessor: images already are [B,3,384,384] tensors in [-1,1]
        def __call__(self, images, return_tensors="pt"):
            return type("R", (), {"pixel_values": images})()

    pipe = DiffUTEPipeline.from_synthetic("fp16x2", "fp16x2")
    pipe.glyph_encoder, pipe.glyph_processor = enc, Proc()
    inp = synthetic.make_inputs(1, 64, 64)
    pv = torch.rand((1, 3, 384, 384), generator=torch.Generator().manual_seed(4)) * 2 - 1
    with torch.no_grad():
        ref_emb = m(pv).last_hidden_state
    kw = dict(masked_image=inp["masked_image"], mask_image=inp["mask"], latents=inp["latents"],
              posterior_noise=inp["posterior_noise"], num_inference_steps=2)
    a = pipe(text=pv, **kw).images.cpu()
    b = pipe(glyph_embeds=ref_emb, **kw).images.cpu()
    err = ((a - b).abs().max() / b.abs().max()).item()
    print(f"pipeline via native glyph encoder vs transformers embeddings: decoded RGB maxrel {err:.3e}")
    assert err < 1e-3
    # and from the uint8 glyph image itself (what draw_text renders, app.ipynb:347-368): the GPU TrOCRGlyphProcessor +
    # native encoder against transformers' PIL-backend image processor + ViTModel
    import numpy as np
    import transformers
    from PIL import Image
    from diffute_b200.glyph_encoder import TrOCRGlyphProcessor
    if not hasattr(transformers, "ViTImageProcessorPil"):
        return
    glyph = np.random.default_rng(2).integers(0, 256, (60, 440, 3), dtype=np.uint8)
    pipe.glyph_processor = TrOCRGlyphProcessor()
    a2 = pipe(text=[glyph], **kw).images.cpu()
    proc = transformers.ViTImageProcessorPil(do_resize=True, size={"height": 384, "width": 384}, resample=2, do_rescale=True,
                                             rescale_factor=1 / 255, do_normalize=True, image_mean=[0.5] * 3,
                                             image_std=[0.5] * 3)
    pv_ref = torch.from_numpy(proc(images=[Image.fromarray(glyph)], return_tensors="np").pixel_values)
    assert torch.equal(pipe.glyph_processor(images=[glyph]).pixel_values.cpu(), pv_ref)
    with torch.no_grad():
        emb = m(pv_ref).last_hidden_state
    b2 = pipe(glyph_embeds=emb, **kw).images.cpu()
    err2 = ((a2 - b2).abs().max() / b2.abs().max()).item()
    print(f"glyph image -> GPU processor -> native encoder -> pipeline vs transformers chain: decoded RGB maxrel {err2:.3e}")
    assert err2 < 1e-3
