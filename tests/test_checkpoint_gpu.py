"""Checkpoint / resume (vae-gan-v2.py:802-808, 966-972): VAEGANTrainer.checkpoint() has the
reference's dictionary layout, its optimiser entries load into a stock torch.optim.Adam over the reference (oracle)
modules, and resuming from it continues the run: [2 steps] == [1 step, save, new trainer, load, 1 step] within the
run-to-run noise of the fp32 atomics (high-accuracy mode: 1e-4 on every loss term, 1e-3 on the squared gradient
norm; in bf16 mode two executions of the same step already differ by ~1e-3)."""
import io

import pytest
import torch

from oracle import models as om
from oracle.step import deterministic_state, make_optimizers, synthetic_batch

pytestmark = pytest.mark.gpu


def build():
    from vae_gan_mark_b200 import modules as M
    from vae_gan_mark_b200.train import LossWeights, VAEGANTrainer
    torch.manual_seed(0)
    G = M.VAEGAN_UNet_SpatialFiLM(4, 32, patch_shape=(64, 32)).cuda().train()
    D = M.Discriminator(3).cuda().train()
    G.load_state_dict(deterministic_state(G, 5)); D.load_state_dict(deterministic_state(D, 6))
    G.char_text_encoder_module.rnn.dropout = 0.0
    return G, D, VAEGANTrainer(G, D, LossWeights.for_family("v2", perceptual=False))


def run_step(tr, G, step):
    ru, en, mask, texts = synthetic_batch(4, 32, 64, step=step)
    eps = torch.randn(4, 32, 1, 1, generator=torch.Generator().manual_seed(100 + step))
    G.style_vae_encoder_module.eps_fn = lambda shape: eps
    out = tr.step(ru.cuda(), en.cuda(), mask.cuda(), texts)
    return {k: float(v) for k, v in out.items() if v.numel() == 1}


@pytest.fixture(autouse=True)
def fp32_mode():
    import vae_gan_mark_b200 as vg
    vg.set_precision("fp32")
    yield
    vg.set_precision("bf16")


def test_checkpoint_layout_and_resume():
    G, D, tr = build()
    run_step(tr, G, 0)
    buf = io.BytesIO()
    torch.save(tr.checkpoint(epoch=3), buf)
    straight = run_step(tr, G, 1)

    buf.seek(0)
    ck = torch.load(buf, map_location="cuda", weights_only=False)
    assert {"model_state_dict", "disc_state_dict", "opt_G_state_dict", "opt_D_state_dict", "epoch"} <= set(ck)
    # the optimiser entries are torch.optim.Adam state_dicts: they load into a stock Adam over the reference modules
    og = om.VAEGAN_UNet_SpatialFiLM(4, 32, patch_hw=(32, 64)).cuda()
    od = om.Discriminator(3).cuda()
    og.load_state_dict(ck["model_state_dict"]); od.load_state_dict(ck["disc_state_dict"])
    opt_g, opt_d = make_optimizers(og, od)
    opt_g.load_state_dict(ck["opt_G_state_dict"]); opt_d.load_state_dict(ck["opt_D_state_dict"])
    st = opt_g.state[next(iter(og.parameters()))]
    assert float(st["step"]) == 1.0 and st["exp_avg"].shape == next(iter(og.parameters())).shape

    G2, D2, tr2 = build()
    tr2.load_checkpoint(ck)
    assert tr2.opt_G.step_count == 1 and tr2.opt_D.step_count == 1
    resumed = run_step(tr2, G2, 1)
    for k in straight:
        tol = 1e-3 if k == "grad_norm_sq" else 1e-4
        assert abs(straight[k] - resumed[k]) <= tol * max(1.0, abs(straight[k])), (k, straight[k], resumed[k])


def test_lr_and_kl_weight_changes_reach_a_captured_graph():
    """Row f4 plumbing: the learning rates (ReduceLROnPlateau, vae-gan-v2.py:944-953) and the KL weight (annealed per
    epoch, vae-gan-v2.py:1002-1004) live in device scalars, so a step captured once as a CUDA graph follows them
    without re-capturing.  bf16 mode (the mode the graph is used in)."""
    import vae_gan_mark_b200 as vg
    vg.set_precision("bf16")
    G, D, tr = build()
    ru, en, mask, texts = synthetic_batch(4, 32, 64, step=0)
    tr.capture(ru.cuda(), en.cuda(), mask.cuda(), texts, warmup=2)
    wg, wd = G.style_vae_encoder_module.e_conv2[0].weight, D.body[2].weight_orig
    w = tr.w

    def kl_part(out):       # loss_G - recon - gan terms == kl_weight * kl  (vae-gan-v2.py:733-737)
        return float(out["loss_G"]) - w.recon * float(out["recon"]) - w.gan * float(out["gan"]), float(out["kl"])

    tr.opt_G.set_lr(0.0); tr.opt_D.set_lr(0.0)
    g0, d0 = wg.detach().clone(), wd.detach().clone()
    part, kl = kl_part(tr.replay())
    torch.cuda.synchronize()
    assert torch.equal(wg, g0) and torch.equal(wd, d0), "lr = 0 must freeze the weights inside the captured graph"
    assert abs(part - w.kl * kl) <= 1e-5 * max(1.0, abs(kl))
    tr.opt_G.set_lr(1e-3)
    tr.replay()
    torch.cuda.synchronize()
    assert not torch.equal(wg, g0) and torch.equal(wd, d0)
    step_g = float((wg - g0).abs().max())
    assert 1e-5 <= step_g <= 3e-3, step_g             # an Adam step moves a weight by about lr
    tr.opt_G.set_lr(0.0)
    for value in (0.0, 0.5):
        tr.set_kl_weight(value)
        part, kl = kl_part(tr.replay())
        assert abs(part - value * kl) <= 1e-5 * max(1.0, abs(kl)), (value, part, kl)
    sched_metric = [1.0, 1.0, 1.0]
    from vae_gan_mark_b200.train import ReduceLROnPlateau
    tr.opt_D.set_lr(1e-4)
    sch = ReduceLROnPlateau(tr.opt_D, mode="min", factor=0.5, patience=1)
    for m in sched_metric:
        sch.step(m)
    assert abs(tr.opt_D.lr - 5e-5) < 1e-12 and abs(float(tr.opt_D.state[3]) - 5e-5) < 1e-11
