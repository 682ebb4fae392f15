"""The reference's `text_editing` around the sampling engine: window choice, masks, resize /
normalise, and the compositing of the decoded crop back into the photograph -- with the pixel work on the GPU.

Drop-in for /root/reference/app.ipynb:653-856 (`text_editing(text, instance_image, slider_step, x0, y0, x1, y1)`):
the photograph goes to the device once as uint8, `dfu_glue_preprocess` produces the three 512 x 512 tensors the
reference builds with albumentations / cv2 / PIL on the host (bit-exact with those libraries; tests/test_glue_gpu.py),
`DiffUTEPipeline` samples, and `dfu_glue_composite` resizes the result back and pastes the text box; only the final
uint8 image returns to the host.  Glyph rendering (PIL + arialuni.ttf) stays reference-side: pass the rendered glyph
image(s) as `text`, or the TrOCR embedding as `glyph_embeds`.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

# 6 * char_height < bound  ->  preset window side (app.ipynb:678-693)
_CROP_PRESETS: Tuple[Tuple[int, int], ...] = ((128, 128), (256, 256), (384, 384), (512, 512), (640, 640), (784, 784),
                                               (1000, 1000))


def crop_window(bbox: Sequence[int], h: int, w: int, rng=None) -> Tuple[int, int, int]:
    """(x0, y0, x1, y1) -> (x_s, y_s, crop_scale) exactly as app.ipynb:668-726 chooses them (including its comparison of
    the y anchor with the image width).  `rng.randint` is used only when the box is at least as large as the window."""
    x1, y1, x2, y2 = (int(v) for v in np.int32(bbox))
    box_h, box_w = y2 - y1, x2 - x1
    side = 6 * box_h
    for bound, preset in _CROP_PRESETS:
        if side < bound:
            side = max(preset, box_w)
            break
    short = min(h, w)
    crop = min(side, short) if box_w < side else short
    rnd = rng if rng is not None else np.random

    def anchor(lo, hi):
        if hi - lo >= crop:
            return int(rnd.randint(lo, max(0, hi - crop - 1)))
        if hi - crop > 0:
            return hi - crop
        return lo if lo + crop < w else 0

    return anchor(x1, x2), anchor(y1, y2), crop


class Preprocessed:
    """Device tensors of one request: image / masked_image [1,3,512,512] in [-1,1], mask [1,1,512,512] and its
    latent-grid reduction, plus the geometry `composite` needs."""

    def __init__(self, image_u8, bbox, window, image, masked_image, mask, mask_latents):
        self.image_u8, self.bbox, self.window = image_u8, bbox, window
        self.image, self.masked_image, self.mask, self.mask_latents = image, masked_image, mask, mask_latents


def preprocess(instance_image, bbox: Sequence[int], window: Optional[Tuple[int, int, int]] = None, device="cuda",
               out_size: int = 512, vae_scale_factor: int = 8, rng=None) -> Preprocessed:
    """instance_image: uint8 [h, w, 3] (numpy or torch).  One H2D copy of the photograph, one kernel."""
    img = torch.as_tensor(np.ascontiguousarray(instance_image) if isinstance(instance_image, np.ndarray) else instance_image)
    if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3:
        raise ValueError("instance_image must be uint8 [h, w, 3]")
    img = img.to(device).contiguous()
    h, w = int(img.shape[0]), int(img.shape[1])
    box = tuple(int(v) for v in np.int32(bbox))
    if window is None:
        window = crop_window(box, h, w, rng)
    x_s, y_s, cs = window
    if not (0 <= x_s < w and 0 <= y_s < h and cs > 0):
        raise ValueError(f"window {window} lies outside the {w} x {h} image")
    cw, ch = min(cs, w - x_s), min(cs, h - y_s)  # numpy slicing clips the window at the image border
    image, masked, mask, mask_lat = ops.glue_preprocess(img, (x_s, y_s, cw, ch), box, out_size, vae_scale_factor)
    return Preprocessed(img, box, (x_s, y_s, cs), image[None], masked[None], mask[None], mask_lat[None])


def composite(decoded: torch.Tensor, pre: Preprocessed, wrap: bool = False) -> torch.Tensor:
    """decoded [1,3,S,S] or [3,S,S] fp32 in [-1,1] (vae.decode(...).sample) -> uint8 [h, w, 3] on the device."""
    d = decoded[0] if decoded.dim() == 4 else decoded
    h, w = int(pre.image_u8.shape[0]), int(pre.image_u8.shape[1])
    x_s, y_s, cs = pre.window
    r_w, r_h = min(cs, w - x_s), min(cs, h - y_s)
    x1, y1, x2, y2 = pre.bbox
    if min(x1, y1, x2, y2) < 0:
        raise ValueError("negative box corners are numpy from-the-end indices in the reference; not supported")
    return ops.glue_composite(d.to(torch.float32).contiguous(), pre.image_u8, (x_s, y_s), (r_w, r_h), (x1, y1, x2, y2), wrap)


@torch.no_grad()
def text_editing(pipe, text, instance_image, slider_step: int, x0, y0, x1, y1, glyph_embeds: Optional[torch.Tensor] = None,
                 generator: Optional[torch.Generator] = None, rng=None, wrap: bool = False, **pipe_kwargs):
    """app.ipynb:653 `text_editing`: returns (edited photograph uint8 [h, w, 3] numpy, mask * 255 uint8 [h, w]).

    `text`: the rendered glyph image(s) for the attached TrOCR encoder (what `draw_text` + `processor` produce), or None
    when `glyph_embeds` [1, 577, 1024] is given.  `pipe`: a DiffUTEPipeline."""
    pre = preprocess(instance_image, (x0, y0, x1, y1), device=pipe.device, rng=rng)
    out = pipe(masked_image=pre.masked_image, mask_image=pre.mask, text=text if glyph_embeds is None else None,
               glyph_embeds=glyph_embeds, num_inference_steps=int(slider_step), generator=generator, **pipe_kwargs)
    edited = composite(out.images, pre, wrap=wrap)
    h, w = int(pre.image_u8.shape[0]), int(pre.image_u8.shape[1])
    bx0, by0, bx1, by1 = pre.bbox
    mask = np.zeros((h, w), np.uint8)  # generate_mask(...) * 255, returned for display only
    mask[max(by0, 0):max(min(by1, h - 1) + 1, 0), max(bx0, 0):max(min(bx1, w - 1) + 1, 0)] = 255
    return edited.cpu().numpy(), mask


@torch.no_grad()
def text_editing_batch(pipe, requests, slider_step: int, max_batch: int = 8, wrap: bool = False, **pipe_kwargs):
    """A queue of `text_editing` requests served in UNet batches ("batching ... across a request queue"):
    requests = [dict(instance_image=uint8 [h, w, 3], bbox=(x0, y0, x1, y1), glyph_embeds=[1, 577, 1024] (or text=glyph
    image), latents=optional [1, 4, 64, 64])].  Photographs may differ in size; every request is pre-processed by its own
    kernel launch, the sampling loop runs once per group of up to `max_batch` requests (one K/V projection of the
    stacked glyph embeddings, one captured step graph at that batch), and each result is composited into its own
    photograph.  Returns a list of (edited uint8 [h, w, 3] numpy, mask * 255) in request order.  At batch 8 one B200
    serves 2.1x the images per second of one-by-one calls (bench.py configs)."""
    results = [None] * len(requests)
    for g0 in range(0, len(requests), max_batch):
        group = requests[g0:g0 + max_batch]
        pres = [preprocess(r["instance_image"], r["bbox"], device=pipe.device, rng=r.get("rng")) for r in group]
        embeds = torch.cat([r["glyph_embeds"].to(pipe.device, torch.float32) if r.get("glyph_embeds") is not None
                            else pipe.encode_glyph([r["text"]] if not isinstance(r["text"], (list, tuple)) else r["text"])
                            for r in group], 0)
        kw = dict(pipe_kwargs)
        if all(r.get("latents") is not None for r in group):
            kw["latents"] = torch.cat([r["latents"] for r in group], 0)
        out = pipe(masked_image=torch.cat([p.masked_image for p in pres], 0), mask_image=torch.cat([p.mask for p in pres], 0),
                   glyph_embeds=embeds, num_inference_steps=int(slider_step), **kw)
        for i, (r, pre) in enumerate(zip(group, pres)):
            edited = composite(out.images[i], pre, wrap=wrap)
            h, w = int(pre.image_u8.shape[0]), int(pre.image_u8.shape[1])
            bx0, by0, bx1, by1 = pre.bbox
            mask = np.zeros((h, w), np.uint8)
            mask[max(by0, 0):max(min(by1, h - 1) + 1, 0), max(bx0, 0):max(min(bx1, w - 1) + 1, 0)] = 255
            results[g0 + i] = (edited.cpu().numpy(), mask)
    return results
