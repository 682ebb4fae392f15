"""The pre-/post-processing kernels of one text_editing request and the Philox noise kernel, launched a few times, for
  ncu --clock-control none -k regex:'glue_|pil_|philox' --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --csv ...
Sizes as in bench.py's text_editing_ddpm record: a 1440 x 1080 photograph, a 680 x 60 glyph image."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from diffute_b200 import glue, ops
from diffute_b200.glyph_encoder import TrOCRGlyphProcessor

rng = np.random.default_rng(5)
photo = torch.from_numpy(rng.integers(0, 256, (1080, 1440, 3), dtype=np.uint8)).cuda()
glyph = rng.integers(0, 256, (60, 680, 3), dtype=np.uint8)
box = (500, 400, 860, 470)
dec = torch.rand((1, 3, 512, 512), device="cuda") * 2 - 1
proc = TrOCRGlyphProcessor()
for _ in range(3):
    pre = glue.preprocess(photo, box)
    out = glue.composite(dec, pre)
    pv = proc(images=[glyph]).pixel_values
    z = ops.philox_normal(1234, 3, 4 * 64 * 64 * 8)
torch.cuda.synchronize()
print("done", out.shape, pv.shape, float(z.std()))
