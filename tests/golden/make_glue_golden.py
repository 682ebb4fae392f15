"""Generates tests/golden/glue_golden.json with the REAL third-party libraries the reference calls for its glue
(cv2.resize at app.ipynb:332-341 via albumentations and at :839; PIL.ImageDraw.rectangle at :370-378): SHA-256 of
their outputs on seeded inputs, so tests/test_glue_oracle.py can pin oracle/glue.py where cv2 is not importable.
Run here (cv2 4.13.0, PIL): python tests/golden/make_glue_golden.py"""
import hashlib, json, os
import numpy as np
import cv2
from PIL import Image, ImageDraw

def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()

def u8_src(seed, shape, binary):
    r = np.random.default_rng(seed)
    a = r.integers(0, 256, shape, dtype=np.uint8)
    return (a > 200).astype(np.uint8) if binary else a

def f32_src(seed):
    r = np.random.default_rng(seed)
    return (r.random((512, 512, 3), dtype=np.float32) * 280 - 10).astype(np.float32)

CASES_U8 = [  # (seed, src shape, (dw, dh), binary)
    (1, (128, 128, 3), (512, 512), False), (2, (384, 384, 3), (512, 512), False), (3, (1024, 1024, 3), (512, 512), False),
    (4, (512, 512, 3), (512, 512), False), (5, (784, 784, 3), (512, 512), False), (6, (300, 417, 3), (512, 512), False),
    (7, (640, 333), (512, 512), True), (8, (1000, 1000), (512, 512), True), (9, (97, 131, 3), (64, 48), False),
    (10, (2, 3, 3), (512, 512), False), (11, (1, 1), (512, 512), False), (12, (1024, 700, 3), (512, 512), False),
]
CASES_F32 = [(21, (300, 300)), (22, (640, 640)), (23, (1000, 784)), (24, (256, 256)), (25, (512, 512)), (26, (128, 97)),
             (27, (3, 5))]
CASES_RECT = [((40, 30), (5, 6, 20, 17)), ((40, 30), (0, 0, 39, 29)), ((40, 30), (-3, 10, 12, 50)), ((16, 16), (7, 7, 7, 7))]

CASES_PIL = [(31, (60, 280, 3), (384, 384)), (32, (60, 640, 3), (384, 384)), (33, (60, 1500, 3), (384, 384)),
             (34, (384, 200, 3), (384, 384)), (35, (500, 384, 3), (384, 384)), (36, (7, 5, 3), (33, 21)), (37, (384, 384, 3), (384, 384))]

import PIL
out = {"cv2": cv2.__version__, "pil": PIL.__version__, "u8": [], "f32": [], "rect": [], "pil_bilinear": []}
for seed, shape, (ow, oh) in CASES_PIL:
    src = u8_src(seed, shape, False)
    out["pil_bilinear"].append({"seed": seed, "shape": list(shape), "dsize": [ow, oh],
                                "sha256": sha(np.array(Image.fromarray(src).resize((ow, oh), resample=Image.BILINEAR)))})
for seed, shape, (dw, dh), binary in CASES_U8:
    src = u8_src(seed, shape, binary)
    out["u8"].append({"seed": seed, "shape": list(shape), "dsize": [dw, dh], "binary": binary,
                      "sha256": sha(cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR))})
for seed, (dw, dh) in CASES_F32:
    ref = cv2.resize(f32_src(seed), (dw, dh))
    out["f32"].append({"seed": seed, "dsize": [dw, dh], "sha256": sha(ref), "sha256_rounded": sha(np.rint(ref).astype(np.int32))})
for (w, h), box in CASES_RECT:
    m = Image.new("L", (w, h), 0)
    ImageDraw.Draw(m).rectangle(box, fill=1)
    out["rect"].append({"size": [w, h], "box": list(box), "sha256": sha(np.array(m))})
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "glue_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out["u8"]), len(out["f32"]), len(out["rect"]), len(out["pil_bilinear"]))
