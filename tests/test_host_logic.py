"""CPU: host-side logic of the product — C-ABI surface, packing, tap tables, schedulers' coefficient math,
checkpoint IO, sharding.  No kernel is launched here (no GPU in this container)."""
import json
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from diffute_b200 import _lib
    L = _lib.lib()  # builds with nvcc if needed; loads without a GPU
    hdr = open(os.path.join(ROOT, "include", "diffute_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(dfu_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    bound = set(_lib.SIGNATURES) | {"dfu_version", "dfu_last_error", "dfu_num_sms", "dfu_gemm", "dfu_gemm_workspace",
                                     "dfu_gemm_plan", "dfu_gemm_stats"}
    assert set(names) == bound, set(names) ^ bound
    assert L.dfu_version() >= 100


def test_gemm_rejects_bad_descriptors_without_gpu():
    import ctypes as C
    from diffute_b200 import _lib
    L = _lib.lib()
    d = _lib.Gemm()
    assert L.dfu_gemm(C.byref(d), None) == -1          # empty problem
    assert b"empty" in L.dfu_last_error()
    d.m, d.n, d.ngroups, d.npass = 128, 100, 1, 1        # n not a multiple of 32
    assert L.dfu_gemm(C.byref(d), None) == -1
    d.n, d.npass = 128, 2
    assert L.dfu_gemm(C.byref(d), None) == -1
    assert L.dfu_gemm(None, None) == -1


def test_struct_layout_matches_header(tmp_path):
    """ctypes mirrors of DfuGemmOperand / DfuGemm have the layout a C compiler gives the header's structs."""
    import ctypes as C
    import subprocess
    from diffute_b200 import _lib
    src = tmp_path / "lay.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "diffute_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(DfuGemmOperand), sizeof(DfuGemm), '
                   'offsetof(DfuGemmOperand,b), offsetof(DfuGemmOperand,tap_dn), offsetof(DfuGemm,g), '
                   'offsetof(DfuGemm,conv), offsetof(DfuGemm,sync_words));return 0;}')
    exe = tmp_path / "lay"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_lib.GemmOperand), C.sizeof(_lib.Gemm), _lib.GemmOperand.b.offset, _lib.GemmOperand.tap_dn.offset,
            _lib.Gemm.g.offset, _lib.Gemm.conv.offset, _lib.Gemm.sync_words.offset]
    assert got == want, (got, want)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "diffute_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu\\n", sizeof(DfuPackJob), offsetof(DfuPackJob,dst), '
                   'offsetof(DfuPackJob,rows), offsetof(DfuPackJob,planes));return 0;}')
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_lib.PackJob), _lib.PackJob.dst.offset, _lib.PackJob.rows.offset, _lib.PackJob.planes.offset]
    assert got == want, (got, want)


def test_weight_packing_and_geglu_interleave():
    from diffute_b200 import ops
    w = torch.randn(8, 6, 3, 3)
    p = ops.pack_conv_weight(w, 2)
    assert p.shape == (16, 54) and p.dtype == torch.float16
    rec = p.float().reshape(2, 8, 9, 6).sum(0)
    assert torch.allclose(rec, w.permute(0, 2, 3, 1).reshape(8, 9, 6), atol=1e-6)
    assert torch.equal(p[:8].float(), w.permute(0, 2, 3, 1).reshape(8, 54).half().float())
    g = torch.arange(64.0)[:, None].repeat(1, 2)
    gi = ops.geglu_interleave(g)
    assert gi[:16, 0].tolist() == list(range(16)) and gi[16:32, 0].tolist() == list(range(32, 48))
    assert gi[32:48, 0].tolist() == list(range(16, 32)) and gi[48:, 0].tolist() == list(range(48, 64))
    x = torch.randn(1000) * 3
    s = ops.split_f16(x, 2)
    assert ((s.float().sum(0) - x).abs() <= 1e-6 * x.abs() + 1e-7).all()   # ~22-bit split; fp16 subnormal floor


def test_stride2_tap_tables():
    from diffute_b200 import ops
    # UNet Downsample2D (pad 1): input row 2*yo + ky - 1
    t = ops.taps_3x3_s2(3, 1)
    for i, (dn, dy, dx) in enumerate(t):
        ky, kx = divmod(i, 3)
        py, px = dn // 3 // 2, dn // 3 % 2
        for yo in (0, 5):
            assert 2 * (yo + dy) + py == 2 * yo + ky - 1
            assert 2 * (yo + dx) + px == 2 * yo + kx - 1
    # VAE encoder (pad (0,1,0,1)): input row 2*yo + ky
    for i, (dn, dy, dx) in enumerate(ops.taps_3x3_s2(1, 0)):
        ky, kx = divmod(i, 3)
        assert 2 * dy + dn // 2 == ky and 2 * dx + dn % 2 == kx
    assert ops.taps_3x3_s1()[0] == (0, -1, -1) and ops.taps_3x3_s1()[8] == (0, 1, 1)


def test_scheduler_tables_and_coefficients_match_oracle():
    from diffute_b200.schedulers import DDIMScheduler, DDPMScheduler
    from oracle.schedulers import DDIMOracle, DDPMOracle
    s, o = DDIMScheduler(), DDIMOracle()
    assert torch.equal(s.alphas_cumprod, o.alphas_cumprod)
    s.set_timesteps(50)
    o.set_timesteps(50)
    assert s.timesteps.tolist() == o.timesteps.tolist() == list(range(981, 0, -20))
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_golden.json")))["ddim_sd2_coeffs"]
    for t, (gx, ge) in zip(s.timesteps.tolist(), gold):
        cx, ce = s.collapsed_coefficients(t)
        ox, oe = o.collapsed_coeffs(t)
        assert abs(cx - ox) < 1e-12 and abs(ce - oe) < 1e-12
        assert abs(cx - gx) < 1e-12 and abs(ce - ge) < 1e-12
    # general coefficient form reproduces the oracle's step (CPU arithmetic only) incl. clip / v-prediction / eta
    for cfg in (dict(), dict(prediction_type="v_prediction"), dict(clip_sample=True), dict(prediction_type="sample")):
        s, o = DDIMScheduler(**cfg), DDIMOracle(**cfg)
        s.set_timesteps(20)
        o.set_timesteps(20)
        x, m, n = torch.randn(2, 4, 8, 8).double(), torch.randn(2, 4, 8, 8).double(), torch.randn(2, 4, 8, 8).double()
        o.alphas_cumprod = o.alphas_cumprod.double()
        o.final_alpha_cumprod = o.final_alpha_cumprod.double()
        for t in (951, 501, 1):
            for eta in (0.0, 0.7):
                a0, a1, p0, d0, d1, sg, clip = s.step_coefficients(t, eta)
                x0 = a0 * x + a1 * m
                if clip:
                    x0 = x0.clamp(-1, 1)
                got = p0 * x0 + d0 * x + d1 * m + sg * n
                ref = o.step(m, t, x, eta=eta, variance_noise=n).prev_sample
                assert torch.allclose(got, ref, atol=1e-9), (cfg, t, eta)
    s, o = DDPMScheduler(), DDPMOracle()
    s.set_timesteps(50)
    o.set_timesteps(50)
    o.alphas_cumprod = o.alphas_cumprod.double()
    x, m, n = torch.randn(2, 4, 8, 8).double(), torch.randn(2, 4, 8, 8).double(), torch.randn(2, 4, 8, 8).double()
    for t in (980, 500, 0):
        a0, a1, p0, d0, d1, sg, clip = s.step_coefficients(t)
        got = p0 * (a0 * x + a1 * m) + d0 * x + d1 * m + sg * n
        ref = o.step(m, t, x, noise=n).prev_sample
        assert torch.allclose(got, ref, atol=1e-9)
    with pytest.raises(ValueError):
        DDIMScheduler().step_coefficients(981)       # set_timesteps not called
    with pytest.raises(ValueError):
        DDIMScheduler().set_timesteps(2000)


def test_scheduler_from_pretrained_pndm_style_config(tmp_path):
    """The SD2-inpainting scheduler folder holds a PNDM config that DDIM/DDPM reinterpret."""
    from diffute_b200.schedulers import DDIMScheduler, DDPMScheduler
    os.makedirs(tmp_path / "scheduler")
    cfg = {"_class_name": "PNDMScheduler", "_diffusers_version": "0.8.0", "beta_end": 0.012,
           "beta_schedule": "scaled_linear", "beta_start": 0.00085, "num_train_timesteps": 1000,
           "set_alpha_to_one": False, "skip_prk_steps": True, "steps_offset": 1, "trained_betas": None,
           "clip_sample": False, "prediction_type": "epsilon"}
    json.dump(cfg, open(tmp_path / "scheduler" / "scheduler_config.json", "w"))
    s = DDIMScheduler.from_pretrained(str(tmp_path), subfolder="scheduler")
    assert s.config.steps_offset == 1 and s.config["prediction_type"] == "epsilon" and s.init_noise_sigma == 1.0
    d = DDPMScheduler.from_pretrained(str(tmp_path), subfolder="scheduler")
    assert d.num_train_timesteps == 1000 and len(d) == 1000


def test_checkpoint_roundtrip_and_legacy_vae_keys(tmp_path):
    from diffute_b200 import checkpoint
    sd = {"encoder.mid_block.attentions.0.query.weight": torch.randn(4, 4),
          "encoder.mid_block.attentions.0.proj_attn.bias": torch.randn(4),
          "decoder.conv_in.weight": torch.randn(2, 2, 3, 3)}
    for safe in (True, False):
        checkpoint.save_diffusers_folder(str(tmp_path / f"m{safe}"), "vae", {"scaling_factor": 0.18215}, sd,
                                         "AutoencoderKL", safe_serialization=safe)
        cfg, back = checkpoint.load_diffusers_folder(str(tmp_path / f"m{safe}"), "vae")
        assert cfg["scaling_factor"] == 0.18215 and cfg["_class_name"] == "AutoencoderKL"
        assert all(torch.equal(back[k], sd[k]) for k in sd)
    r = checkpoint.remap_legacy_vae_keys(sd)
    assert "encoder.mid_block.attentions.0.to_q.weight" in r
    assert "encoder.mid_block.attentions.0.to_out.0.bias" in r and "decoder.conv_in.weight" in r
    with pytest.raises(FileNotFoundError):
        checkpoint.load_diffusers_folder(str(tmp_path / "nope"), "unet")


def test_synthetic_is_deterministic_and_order_independent():
    from diffute_b200 import arch, synthetic
    shapes = arch.vae_param_shapes()
    a = synthetic.make_state_dict(shapes)
    b = synthetic.make_state_dict(dict(reversed(list(shapes.items()))))
    assert all(torch.equal(a[k], b[k]) for k in shapes)
    i1, i2 = synthetic.make_inputs(2, 64, 64), synthetic.make_inputs(2, 64, 64)
    assert all(torch.equal(i1[k], i2[k]) for k in i1)
    m = i1["mask"]
    assert m[0, 0, 16:32, 8:56].min() == 1 and m.sum() == 2 * 16 * 48
    assert (i1["masked_image"][m.expand(-1, 3, -1, -1) > 0.5] == -1).all()


def test_engine_fails_loudly_without_cuda():
    """No CPU fallback: constructing the engine / launching a kernel without a GPU must raise, not degrade."""
    if torch.cuda.is_available():
        pytest.skip("needs a CPU-only box")
    from diffute_b200 import arch, synthetic
    from diffute_b200.vae import AutoencoderKL
    with pytest.raises(Exception):
        AutoencoderKL(synthetic.make_state_dict(arch.vae_param_shapes()), device="cuda")
    assert "oracle" not in "".join(open(os.path.join(ROOT, "diffute_b200", f)).read()
                                   for f in os.listdir(os.path.join(ROOT, "diffute_b200"))
                                   if f.endswith(".py")).replace("the oracle", "").replace("CPU oracle", "")


def test_attention_streamk_plan_invariants():
    """dfu_attention_plan (host arithmetic only): every CTA gets a non-empty contiguous range of key-block units, the
    unit -> CTA map inverts the range map, no item is cut into more pieces than the merge kernel holds (8), an explicit
    kv_splits = 1 never cuts, and the UNet shapes get the distribution DESIGN.md describes."""
    import ctypes as C
    from diffute_b200 import _lib
    L = _lib.lib()
    out = (C.c_int32 * 5)()
    shapes = [(1, 5, 4096, 4096), (1, 5, 4096, 577), (1, 10, 1024, 1024), (1, 10, 1024, 577), (1, 20, 256, 256),
              (1, 20, 256, 577), (1, 20, 64, 64), (1, 20, 64, 577), (8, 5, 9216, 9216), (2, 5, 4096, 577),
              (1, 2, 64, 4096), (1, 2, 200, 130), (3, 7, 130, 1500), (1, 1, 1, 1), (1, 1, 128, 65)]
    for B, h, Nq, Nk in shapes:
        for ks in (0, 1, 2, 3, 5, 64):
            assert L.dfu_attention_plan(B, h, Nq, Nk, ks, out) == 0
            G, items, nblk, pieces, ok = list(out)
            assert ok == 1, (B, h, Nq, Nk, ks)
            assert items == B * h * -(-Nq // 128) and nblk == -(-Nk // 64)
            assert items <= G <= items * nblk and pieces <= 8, (B, h, Nq, Nk, ks, G, pieces)
            if ks == 1:
                assert G == items and pieces == 1
    # level-0 self-attention at 64x64 latents: two CTAs per SM when a device is visible, else the 148-SM default
    L.dfu_attention_plan(1, 5, 4096, 4096, 0, out)
    assert out[0] in (2 * 148, 2 * max(L.dfu_num_sms(), 1))
    L.dfu_attention_plan(1, 5, 4096, 577, 0, out)   # 577 glyph tokens: never cut
    assert out[0] == out[1] == 160 and out[3] == 1


def test_gemm_plan_tilings_without_gpu():
    """dfu_gemm_plan (host arithmetic only): the tiling the library would launch for representative shapes — stage ring
    within the shared-memory budget, FP16X2 counting k-blocks once (its three products share a stage), explicit
    tilings honoured, illegal ones refused, block_n multiples of 16 accepted (32 for GEGLU)."""
    import ctypes as C
    from diffute_b200 import _lib
    L = _lib.lib()
    out = (C.c_int32 * 8)()

    def plan(m, n, k, npass=1, conv=0, hw=None, epi=0, tune=(0, 0, 0), taps=1, kernel=0):
        d = _lib.Gemm()
        d.m, d.n, d.ngroups, d.npass, d.epi = m, n, 1, npass, epi
        d.g[0].ntaps, d.g[0].k_per_tap = taps, k
        if conv:
            d.conv, d.B, d.H, d.W = 1, 1, hw, hw
            d.g[0].a_mode, d.g[0].a_c = 1, k
        d.block_n, d.splits, d.stages, d.kernel = (*tune, kernel)
        rc = L.dfu_gemm_plan(C.byref(d), out)
        return rc, list(out)

    for m, n, k, taps, conv, hw in [(4096, 320, 320, 1, 0, None), (4096, 320, 320, 9, 1, 64), (1024, 640, 640, 9, 1, 32),
                                    (256, 1280, 1280, 9, 1, 16), (64, 1280, 1280, 9, 1, 8), (4096, 2560, 320, 1, 0, None),
                                    (262144, 128, 128, 9, 1, 512)]:
        for npass in (1, 3):
            rc, (bn, sp, st, tm, tn, kb, kern, _) = plan(m, n, k, npass, conv, hw, taps=taps, kernel=1)
            assert rc == 0, L.dfu_last_error()
            assert kb == taps * k // 64                      # k-blocks counted once, also for the 3-pass mode
            assert bn % 16 == 0 and n % bn == 0 and tn == n // bn and tm == -(-m // 128)
            assert 1 <= sp <= kb and 2 <= st <= 12
            stage = (2 if npass == 3 else 1) * (16384 + bn * 128)
            assert st * stage + 1024 <= 224 * 1024           # ring + alignment slack inside the dynamic smem budget
    assert plan(4096, 320, 320, tune=(80, 1, 4))[1][:3] == [80, 1, 4]     # block_n = 80 is a legal UMMA N
    assert plan(4096, 320, 320, tune=(160, 9, 3))[1][1] == 5              # more splits than k-blocks: clamped
    assert plan(4096, 320, 320, tune=(48, 1, 3))[0] == -1                  # does not divide n
    assert plan(4096, 2560, 320, epi=2, tune=(80, 1, 3))[0] == -1          # GEGLU pairs value/gate columns in 32s
    assert plan(4096, 320, 320, npass=3, tune=(160, 1, 12))[1][2] < 12    # 12 double-size stages do not fit: clamped


def test_crop_window_product_equals_literal_restatement():
    """diffute_b200.glue.crop_window (table-driven) against the literal restatement of app.ipynb:668-726 in the oracle,
    including the branches that consult the random generator and the inputs on which the reference raises."""
    import numpy as np
    from diffute_b200 import glue as P
    from oracle import glue as G
    rng = np.random.default_rng(0)
    raised = 0
    for _ in range(4000):
        h, w = (int(v) for v in rng.integers(64, 2000, 2))
        x0 = int(rng.integers(0, w - 2)); x1 = int(rng.integers(x0 + 1, w))
        y0 = int(rng.integers(0, h - 2)); y1 = int(rng.integers(y0 + 1, h))
        try:
            ref = G.crop_window((x0, y0, x1, y1), h, w, np.random.RandomState(5))
        except ValueError:  # np.random.randint(low, high <= low): the reference fails on such boxes
            raised += 1
            try:
                P.crop_window((x0, y0, x1, y1), h, w, np.random.RandomState(5))
            except ValueError:
                continue
            raise AssertionError("product did not raise where the reference does")
        assert P.crop_window((x0, y0, x1, y1), h, w, np.random.RandomState(5)) == ref, (h, w, x0, y0, x1, y1)
    assert raised > 0


def test_every_entry_point_is_documented():
    """INTEGRATION.md section 3 names every extern "C" entry point of the header (directly, or through its
    `name(+_workspace)` / `dfu_trace_set_{a,b}` shorthands)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = sorted(set(re.findall(r"\b(dfu_[a-z0-9_]+)\s*\(", open(os.path.join(root, "include", "diffute_b200.h")).read())))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    expanded = set(re.findall(r"dfu_[a-z0-9_]+", doc))
    for base, suffix in re.findall(r"(dfu_[a-z0-9_]+)\(\+(_[a-z]+)\)", doc):
        expanded.add(base + suffix)
    for prefix, alts in re.findall(r"(dfu_[a-z0-9_]+_)\{([a-z0-9_,]+)\}", doc):
        expanded.update(prefix + a for a in alts.split(","))
    missing = [n for n in names if n not in expanded]
    assert not missing, missing
