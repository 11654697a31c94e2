"""GPU parity of the full training step against the CPU oracle (which is pinned to the reference by
tests/test_oracle_golden.py).  Runs the drop-in modules of vae_gan_mark_b200 through the C ABI kernels.

Tolerances (bf16 storage / bf16 tensor-core inputs, fp32 accumulation): relative L2 error per tensor
<= 2e-2 on activations and losses (north_star), <= 5e-2 on per-parameter gradients, whose error compounds
through ~40 bf16 layers; the per-tensor numbers are printed so regressions are visible.
"""
import os

import pytest
import torch

from oracle import models as om
from oracle.step import LossWeights as OLW, deterministic_state, make_optimizers, synthetic_batch, train_step

pytestmark = pytest.mark.gpu

ACT_TOL, GRAD_TOL = 2e-2, 5e-2


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def build_pair(family, h, w, z):
    from vae_gan_mark_b200 import modules as M
    if family == "base":
        og = om.VAEGAN(4, z, 64, 3, patch_hw=(h, w))
        mg = M.VAEGAN(4, z, 64, 3, patch_shape=(w, h), text_embedder=om.hash_sentence_embedding)
    elif family == "v2":
        og = om.VAEGAN_UNet_SpatialFiLM(4, z, patch_hw=(h, w))
        mg = M.VAEGAN_UNet_SpatialFiLM(4, z, patch_shape=(w, h))
    else:
        og = om.VAEGAN_UNet_CharEmb(4, z, patch_hw=(h, w), repaired=True)
        mg = M.VAEGAN_UNet_CharEmb(4, z, patch_shape=(w, h))
    od, md = om.Discriminator(3), M.Discriminator(3)
    sg, sd = deterministic_state(og, 1234), deterministic_state(od, 4321)
    og.load_state_dict(sg, strict=True); od.load_state_dict(sd, strict=True)
    mg.load_state_dict(sg, strict=True); md.load_state_dict(sd, strict=True)      # same keys/shapes: drop-in contract
    for g in (og, mg):
        g.train()
        if hasattr(g, "char_text_encoder_module"):
            g.char_text_encoder_module.rnn.dropout = 0.0   # GRU dropout draws from different RNGs on CPU and CUDA
    od.train(); md.train()
    return og, od, mg.cuda(), md.cuda()


CASES = [("base", 32, 32, 4, 128), ("v2", 32, 64, 2, 128), ("v2", 32, 32, 3, 32), ("unet", 32, 32, 2, 128),
         ("base", 64, 64, 16, 128)]


@pytest.mark.parametrize("family,h,w,batch,z", CASES)
def test_train_step_matches_oracle(family, h, w, batch, z):
    from vae_gan_mark_b200.train import LossWeights, VAEGANTrainer
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    og, od, mg, md = build_pair(family, h, w, z)
    wts = OLW.for_family(family)
    opt_g, opt_d = make_optimizers(og, od)
    ru, en, mask, texts = synthetic_batch(batch, h, w, step=0)
    ref = train_step(og, od, opt_g, opt_d, (ru, en, mask, texts), wts, seed=10_000)

    grads = {}
    trainer = VAEGANTrainer(mg, md, LossWeights(wts.recon, wts.kl, wts.gan),
                            grad_hook=lambda which, params: grads.setdefault(which, [p.grad.clone() if p.grad is not None else None for p in params]))
    torch.manual_seed(10_000)
    mg.__dict__["eps_fn"] = lambda shape: torch.randn(shape)
    enc = getattr(mg, "style_vae_encoder_module", None) or mg.encoder
    enc.__dict__["eps_fn"] = mg.__dict__["eps_fn"]
    out = trainer.step(ru.cuda(), en.cuda(), mask.cuda(), texts)
    torch.cuda.synchronize()

    report = {}
    for k in ("loss_G", "loss_D", "recon", "kl", "gan", "d_real", "d_fake"):
        report[k] = abs(float(out[k]) - ref.losses[k]) / max(abs(ref.losses[k]), 1e-6)
    report["fake"] = rel(out["fake"], ref.recon)
    report["mu"] = rel(out["mu"], ref.mu)
    report["logvar"] = rel(out["logvar"], ref.logvar)
    report["grad_norm"] = abs(float(out["grad_norm_sq"]) ** 0.5 - ref.grad_norm) / ref.grad_norm
    print(family, h, w, {k: f"{v:.2e}" for k, v in report.items()})
    gerr = {}
    for (name, _), g in zip(md.named_parameters(), grads["D"]):
        gerr["D." + name] = rel(g, ref.d_grads[name])
    for (name, p), g in zip(mg.named_parameters(), grads["G"]):
        if name in ref.g_grads and g is not None:
            gerr["G." + name] = rel(g, ref.g_grads[name])
    worst = sorted(((k, v) for k, v in gerr.items() if not ref_is_noise(k, ref)), key=lambda kv: -kv[1])[:14]
    print("worst grads:", [(k, f"{v:.2e}") for k, v in worst])
    assert set(n for n, p in mg.named_parameters() if p.grad is not None) >= set(ref.g_grads), "missing G gradients"
    for k, v in report.items():
        assert v <= ACT_TOL, (k, v)
    for k, v in gerr.items():
        # gradients that are exactly-zero-in-theory (conv bias before BatchNorm) are pure rounding noise on both sides
        if ref_is_noise(k, ref):
            continue
        assert v <= GRAD_TOL, (k, v)
    # parameters after the step: Adam moves every weight by ~lr, so compare the update direction loosely
    for (name, p), (_, q) in zip(mg.named_parameters(), og.named_parameters()):
        assert rel(p, q) <= 1e-3, name


def ref_is_noise(key, ref):
    which, name = key.split(".", 1)
    g = (ref.d_grads if which == "D" else ref.g_grads)[name]
    return float(g.abs().max()) < 1e-7
