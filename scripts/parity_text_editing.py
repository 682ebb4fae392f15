"""The reference's serving call end to end at full size, engine vs CPU oracle chain: a 1440 x 1080 uint8 photograph and a
text box -> window / masks / cv2-exact resize / normalise -> 50 steps at 512 x 512 (benched precision mode) -> decode ->
resize back + paste -> uint8 photograph.  Reports how many output pixels differ and by how much.
Usage: python scripts/parity_text_editing.py [steps]  -> gpurun_out/parity_text_editing.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from diffute_b200 import arch, glue, synthetic
from diffute_b200.pipeline import DiffUTEPipeline
from oracle import DDIMOracle, UNetOracle, VAEOracle, sample_loop
from oracle import glue as G

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
torch.set_num_threads(min(os.cpu_count(), int(os.environ.get("DFU_CPU_THREADS", "32"))))
usd = synthetic.make_state_dict(arch.unet_param_shapes())
vsd = synthetic.make_state_dict(arch.vae_param_shapes())
rng = np.random.default_rng(5)
h, w, bbox = 1080, 1440, (500, 400, 860, 470)
yy, xx = np.mgrid[0:h, 0:w]
photo = np.stack([127 + 90 * np.sin(xx / 61.0) * np.cos(yy / 47.0), xx * 255.0 / w, yy * 255.0 / h], -1)
photo = np.clip(np.rint(photo + rng.normal(0, 12, photo.shape)), 0, 255).astype(np.uint8)
inp = synthetic.make_inputs(1, 512, 512)
pipe = DiffUTEPipeline.from_synthetic("fp16", "fp16x2", state_dicts=(usd, vsd), vae_encoder_precision="fp16")
t0 = time.time()
edited, _ = glue.text_editing(pipe, None, photo, steps, *bbox, glyph_embeds=inp["glyph_embeds"], latents=inp["latents"],
                              sample_posterior=False)
t_gpu = time.time() - t0
uo, vo = UNetOracle(), VAEOracle()
uo.load_state_dict(usd); vo.load_state_dict(vsd)
win = G.crop_window(bbox, h, w)
t0 = time.time()
_, mim_c, msk_c = G.preprocess(photo, bbox, win)
dec = sample_loop(uo, vo, DDIMOracle(), torch.from_numpy(mim_c)[None], torch.from_numpy(msk_c.astype(np.float32))[None],
                  inp["glyph_embeds"], inp["latents"], steps)
ref = G.composite(dec[0].numpy(), photo, bbox, win)
t_cpu = time.time() - t0
d = np.abs(edited.astype(int) - ref.astype(int))
x1, y1, x2, y2 = bbox
box = d[y1:y2, x1:x2]
res = {"photo": [h, w], "bbox": list(bbox), "window": list(win), "steps": steps, "precision": "mixed",
       "pixels_changed_by_the_edit": int((ref != photo).any(-1).sum()),
       "box_values_differing": float((box > 0).mean()), "box_max_abs_diff_grey_levels": int(box.max()),
       "outside_box_identical": bool((d.sum(-1)[np.where(np.ones((h, w), bool))].reshape(h, w)[:y1] == 0).all()
                                     and np.array_equal(np.delete(edited, np.s_[y1:y2], 0), np.delete(ref, np.s_[y1:y2], 0))),
       "first_call_gpu_seconds": t_gpu, "cpu_oracle_seconds": t_cpu}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/parity_text_editing.json", "w"), indent=1)
print(json.dumps(res))
