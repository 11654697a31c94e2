"""Whole-model eval() forward on the GPU against the fixtures recorded from the REFERENCE's own modules.

The reference's ``val_loop`` (vae-gan.py:325-377; vae-gan-v2.py:560-640) runs ``model.eval()`` + G forward + L1 on every
validation batch and drives ReduceLROnPlateau / best-model selection with the result.  In eval mode BatchNorm uses its
running statistics, spectral norm does no power iteration, and the reparameterisation noise is still sampled.

The fixtures (tests/golden/*.pt, key "eval") were recorded by tests/golden/make_golden.py after the reference had trained
for the fixture's steps; the same state is reproduced here by running the oracle (pinned to those fixtures by
tests/test_oracle_golden.py) for the same steps on the CPU and loading its state_dict into the drop-in modules.
Tolerance: 2e-2 (north_star's bf16 bound) on the reconstructed image (norm, sum, first 16 pixels), mu and D's patch logits.
A second eval forward after further training steps of a captured CUDA graph checks that eval never sees stale weight
operands (the weights are updated through raw pointers inside the graph).
"""
import os

import pytest
import torch

from oracle import models as om
from oracle.step import LossWeights as OLW, deterministic_state, make_optimizers, synthetic_batch, train_step

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-2


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def build(family, h, w, z):
    from vae_gan_mark_b200 import modules as M
    if family == "base":
        return om.VAEGAN(4, z, 64, 3, patch_hw=(h, w)), M.VAEGAN(4, z, 64, 3, patch_shape=(w, h), text_embedder=om.hash_sentence_embedding)
    if family == "v2":
        return om.VAEGAN_UNet_SpatialFiLM(4, z, patch_hw=(h, w)), M.VAEGAN_UNet_SpatialFiLM(4, z, patch_shape=(w, h))
    if family == "oldv":
        return om.VAEGAN_UNet_SpatialFiLM_OldV(4, z, patch_hw=(h, w)), M.VAEGAN_UNet_SpatialFiLM_OldV(4, z, patch_shape=(w, h))
    return om.VAEGAN_UNet_CharEmb(4, z, patch_hw=(h, w), repaired=True), M.VAEGAN_UNet_CharEmb(4, z, patch_shape=(w, h))


@pytest.mark.parametrize("case", ["base_32x32_b4", "base_64x64_b16", "v2_32x64_b2", "unet_32x32_b2", "oldv_32x64_b2",
                                  "v2_128x128_b8", "unet_256x256_b2"])
def test_eval_forward_matches_reference_fixture(case):
    from vae_gan_mark_b200 import modules as M
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    gold = torch.load(os.path.join(GOLD, case + ".pt"), weights_only=False)
    family, h, w, batch, z = gold["family"], gold["h"], gold["w"], gold["batch"], gold["z"]
    og, mg = build(family, h, w, z)
    od, md = om.Discriminator(3), M.Discriminator(3)
    og.load_state_dict(deterministic_state(og, 1234)); od.load_state_dict(deterministic_state(od, 4321))
    og.train(); od.train()
    if gold.get("gru_dropout") is not None:
        og.char_text_encoder_module.rnn.dropout = gold["gru_dropout"]
    opt_g, opt_d = make_optimizers(og, od)
    wts = OLW.for_family(family)
    for step in range(len(gold["steps"])):          # the training steps the reference had taken when "eval" was recorded
        train_step(og, od, opt_g, opt_d, synthetic_batch(batch, h, w, step=step), wts, seed=10_000 + step, keep_grads=False)
    mg.load_state_dict(og.state_dict(), strict=True); md.load_state_dict(od.state_dict(), strict=True)
    mg, md = mg.cuda().eval(), md.cuda().eval()
    ru, en, mask, texts = synthetic_batch(batch, h, w, step=7)
    eps_fn = lambda shape: torch.randn(shape)
    mg.__dict__["eps_fn"] = eps_fn
    (getattr(mg, "style_vae_encoder_module", None) or mg.encoder).__dict__["eps_fn"] = eps_fn
    sd_before = {k: v.clone() for k, v in list(mg.state_dict().items()) + list(md.state_dict().items())}
    with torch.no_grad():
        torch.manual_seed(77)
        fake, mu, logvar = mg(ru.cuda(), mask.cuda(), texts)
        d_out = md(en.cuda())
    torch.cuda.synchronize()
    want = gold["eval"]
    f = fake.double().cpu().flatten()
    rs = want["recon_sum"]
    errs = {"|fake|": abs(float(f.norm()) - float(rs[0])) / float(rs[0]), "sum fake": abs(float(f.sum()) - float(rs[1])) / abs(float(rs[1])),
            "fake[:16]": float((f[:16] - rs[2:18]).abs().max()), "mu": rel(mu, want["mu"]),
            "D(en)": float((d_out.double().cpu() - want["d_out"].double()).norm() / want["d_out"].double().norm().clamp_min(1e-12))}
    print(case, {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        assert v <= TOL, (k, v)
    # eval must not touch any state: running statistics, num_batches_tracked, spectral-norm u / v
    for k, v in list(mg.state_dict().items()) + list(md.state_dict().items()):
        assert torch.equal(v, sd_before[k]), k


def test_eval_between_graph_replays_sees_current_weights():
    """train (captured graph) -> eval -> train -> eval, the reference's epoch loop (vae-gan.py:560-600): the second eval
    must use the weights of the second training phase, not operands cached by the first eval."""
    from vae_gan_mark_b200 import layers as L, modules as M
    from vae_gan_mark_b200.train import LossWeights, VAEGANTrainer
    h = w = 32
    batch = 4
    torch.manual_seed(0)
    mg = M.VAEGAN_UNet_SpatialFiLM(4, 128, patch_shape=(w, h)).cuda().train()
    md = M.Discriminator(3).cuda().train()
    tr = VAEGANTrainer(mg, md, LossWeights.for_family("v2", perceptual=False), lr_g=1e-3, lr_d=1e-3)
    ru, en, mask, texts = synthetic_batch(batch, h, w, step=0)
    ru, en, mask = ru.cuda(), en.cuda(), mask.cuda()
    eps = torch.randn(batch, 128, 1, 1).cuda()
    enc = mg.style_vae_encoder_module

    def evaluate():
        mg.eval()
        enc.__dict__["eps_fn"] = lambda shape: eps          # the same noise for every evaluation (training draws its own)
        with torch.no_grad():
            out = mg(ru, mask, texts)[0].clone()
        del enc.__dict__["eps_fn"]
        mg.train()
        return out

    tr.capture(ru, en, mask, texts)
    for _ in range(3):
        tr.replay()
    e1 = evaluate()
    for _ in range(20):
        tr.replay()
    e2 = evaluate()
    L.bump_weight_epoch()            # force every cached operand to be rebuilt from the current weights
    e2_fresh = evaluate()
    torch.cuda.synchronize()
    moved, stale = float((e2 - e1).abs().max()), float((e2 - e2_fresh).abs().max())
    print(f"eval output moved by {moved:.3e} over 20 steps; with freshly rebuilt operands it differs by {stale:.3e}")
    assert moved > 1e-3, "20 training steps at lr 1e-3 must change the eval output"
    assert stale <= 1e-2 * moved, (stale, moved)      # (not bit-equal: the heads GEMM combines its K splits with atomics)
