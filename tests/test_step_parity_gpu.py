"""GPU parity of the full training step against the CPU oracle (which is pinned to the reference by
tests/test_oracle_golden.py).  Runs the drop-in modules of vae_gan_mark_b200 through the C ABI kernels.

Tolerances.  The CUDA path stores activations in bf16 and feeds bf16 to the tensor cores (fp32 accumulation,
fp32 master weights); the oracle is fp32.
  * losses, reconstructed image: relative error <= 2e-2 (north_star's bf16 bound); the adversarial terms of the tiny
    test images are means over a handful of patch logits and get 2e-2 x sqrt(64 / n_logits) (see the test body);
  * mu / logvar (10 bf16 conv+BatchNorm layers with batch statistics over as few as 16 samples): <= 6e-2;
  * per-parameter gradients: with random inputs and random weights the true gradients are small residuals of
    heavily cancelling sums, so ANY bf16 evaluation deviates by tens of percent from fp32 -- the reference's own
    modules under torch.autocast(bfloat16) on the same GPU are used as the calibration: our relative L2 error must
    be <= max(5e-2, 2.5 x the autocast run's error) per tensor (the same rule is applied to the scalar quantities).  The backward kernels themselves are pinned
    tightly (<= 1e-2, typically 2e-3) op by op in tests/test_ops_gpu.py, where no such cancellation occurs.
"""
import os

import pytest
import torch

from oracle import models as om
from oracle.step import LossWeights as OLW, deterministic_state, make_optimizers, synthetic_batch, train_step

pytestmark = pytest.mark.gpu

ACT_TOL, LATENT_TOL, GRAD_TOL, CAL = 2e-2, 6e-2, 5e-2, 2.5


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
    elif family == "oldv":
        og = om.VAEGAN_UNet_SpatialFiLM_OldV(4, z, patch_hw=(h, w))
        mg = M.VAEGAN_UNet_SpatialFiLM_OldV(4, z, patch_shape=(w, h))
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
         ("base", 64, 64, 16, 128), ("oldv", 32, 64, 2, 128), ("oldv", 64, 64, 5, 64),
         # the benchmark's own image sizes (BASELINE configs[1], configs[2]); also held to the fixtures recorded from the
         # reference itself (tests/golden/v2_128x128_b8.pt, unet_256x256_b2.pt)
         ("v2", 128, 128, 8, 128), ("unet", 256, 256, 2, 128)]
# fixtures recorded from the reference WITHOUT GRU dropout (the CPU and CUDA dropout masks cannot be lined up, so the step
# tests run with rnn.dropout = 0): the base family has no GRU, the two big ones were recorded with dropout switched off
GOLDEN = {("v2", 128, 128, 8, 128): "v2_128x128_b8", ("unet", 256, 256, 2, 128): "unet_256x256_b2",
          ("base", 64, 64, 16, 128): "base_64x64_b16", ("base", 32, 32, 4, 128): "base_32x32_b4"}


@pytest.mark.parametrize("family,h,w,batch,z", CASES)
def test_train_step_matches_oracle(family, h, w, batch, z):
    from vae_gan_mark_b200.train import LossWeights, VAEGANTrainer
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    og, od, mg, md = build_pair(family, h, w, z)
    wts = OLW.for_family(family)
    opt_g, opt_d = make_optimizers(og, od)
    ru, en, mask, texts = synthetic_batch(batch, h, w, step=0)
    ref = train_step(og, od, opt_g, opt_d, (ru, en, mask, texts), wts, seed=10_000)

    # calibration: the same oracle modules under torch bf16 autocast on this GPU (cuDNN/cuBLAS)
    import copy
    cg, cd = build_pair(family, h, w, z)[:2]
    cg, cd = cg.cuda(), cd.cuda()
    cg.__dict__["_cuda_eps"] = True
    torch.manual_seed(10_000)
    eps_cpu = torch.randn(batch, z, 1, 1)
    orig_randn_like = torch.randn_like
    try:
        torch.randn_like = lambda t, **k: eps_cpu.to(t.device, t.dtype) if tuple(t.shape) == tuple(eps_cpu.shape) else orig_randn_like(t, **k)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            cal = train_step(cg, cd, *make_optimizers(cg, cd), (ru.cuda(), en.cuda(), mask.cuda(), texts), wts)
    finally:
        torch.randn_like = orig_randn_like
    cal_err = {"D." + k: rel(v, ref.d_grads[k]) for k, v in cal.d_grads.items()}
    cal_err.update({"G." + k: rel(v, ref.g_grads[k]) for k, v in cal.g_grads.items() if k in ref.g_grads})
    cal_rep = {k: abs(cal.losses[k] - ref.losses[k]) / max(abs(ref.losses[k]), 1e-6) for k in ref.losses}
    cal_rep.update({"fake": rel(cal.recon, ref.recon), "mu": rel(cal.mu, ref.mu), "logvar": rel(cal.logvar, ref.logvar),
                    "grad_norm": abs(cal.grad_norm - ref.grad_norm) / ref.grad_norm})
    cal_median = sorted(cal_err.values())[len(cal_err) // 2]
    print("autocast calibration:", {"fake": f"{rel(cal.recon, ref.recon):.2e}", "mu": f"{rel(cal.mu, ref.mu):.2e}",
                                    "median grad err": f"{sorted(cal_err.values())[len(cal_err) // 2]:.2e}"})

    grads = {}
    trainer = VAEGANTrainer(mg, md, LossWeights(wts.recon, wts.kl, wts.gan),
                            grad_hook=lambda which, params: grads.setdefault(which, [p.grad.clone() if p.grad is not None else None for p in params]))
    before = {"G": [p.detach().clone() for p in trainer.opt_G.params], "D": [p.detach().clone() for p in trainer.opt_D.params]}
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
    print("median grad err:", f"{sorted(gerr.values())[len(gerr) // 2]:.2e}")
    # The adversarial scalars are means over only batch x (h/16 - 1) x (w/16 - 1) patch logits (6 at 32x64, batch 2):
    # independent bf16 rounding errors of the logits average out as 1/sqrt(n), so below 64 logits their bound widens
    # accordingly (2e-2 is the bound for the 7x7 maps of the 128x128 workload and larger).
    n_logits = batch * max(1, h // 16 - 1) * max(1, w // 16 - 1)
    few = max(1.0, (64.0 / n_logits) ** 0.5)
    for k, v in report.items():
        base = LATENT_TOL if k in ("mu", "logvar", "kl", "grad_norm") else ACT_TOL
        if k in ("d_fake", "d_real", "gan", "loss_D", "loss_G"):
            base *= few
        assert v <= max(base, CAL * cal_rep.get(k, 0.0)), (k, v, cal_rep.get(k))
    bad = []
    for k, v in gerr.items():
        # gradients that are exactly-zero-in-theory (conv bias before BatchNorm) are pure rounding noise on both sides
        if ref_is_noise(k, ref):
            continue
        # (both runs are noisy -- split-K atomics make them non-deterministic -- so the autocast run's median error is
        # also accepted as a floor)
        if v > max(GRAD_TOL, CAL * cal_err.get(k, 0.0), 1.5 * cal_median):
            bad.append((k, f"{v:.2e}", f"autocast {cal_err.get(k, 0.0):.2e}"))
    assert not bad, bad
    # ---- the fixture recorded from the REFERENCE's own modules (tests/golden/make_golden.py), where one exists ----
    gname = GOLDEN.get((family, h, w, batch, z))
    if gname is not None:
        gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", gname + ".pt"), weights_only=False)["steps"][0]
        for k, v in gold["losses"].items():
            base = LATENT_TOL if k == "kl" else ACT_TOL
            if k in ("d_fake", "d_real", "gan", "loss_D", "loss_G"):
                base *= few
            e = abs(float(out[k]) - v) / max(abs(v), 1e-6)
            assert e <= max(base, CAL * cal_rep.get(k, 0.0)), ("reference golden", k, float(out[k]), v)
        assert rel(out["mu"], gold["mu"]) <= max(LATENT_TOL, CAL * cal_rep["mu"])
        assert rel(out["logvar"], gold["logvar"]) <= max(LATENT_TOL, CAL * cal_rep["logvar"])
        rs = gold["recon_sum"]                      # [l2 norm, sum, first 16 values] of the reconstructed image
        f = out["fake"].detach().double().cpu().flatten()
        assert abs(float(f.norm()) - float(rs[0])) <= ACT_TOL * float(rs[0])
        assert abs(float(f.sum()) - float(rs[1])) <= ACT_TOL * abs(float(rs[1]))
        assert float((f[:16] - rs[2:18]).abs().max()) <= 1.5 * ACT_TOL      # 16 single pixels in [0, 1], absolute
    cal_grads = {"D": [cal.d_grads.get(n) for n, _ in md.named_parameters()],
                 "G": [cal.g_grads.get(n) for n, _ in mg.named_parameters()]}
    check_post_step_state(trainer, mg, md, og, od, grads, before, ref, cal_grads)


def sign_flips(net_params, grads, rgrads, signs=None):
    """(flipped, total) over the gradient elements whose oracle value is above 4 x the rms difference between ``grads`` and
    the oracle's gradient of that tensor: there ``signs`` (default: ``grads`` itself) must have the oracle's sign."""
    flips = total = 0
    for i, ((name, _), g) in enumerate(zip(net_params, grads)):
        if g is None or name not in rgrads:
            continue
        gr = rgrads[name].double()
        gd = g.detach().double().cpu()
        noise = float((gd - gr).pow(2).mean().sqrt())
        sel = gr.abs() > max(4.0 * noise, 1e-7)
        sg = gd if signs is None else signs[i].detach().double().cpu()
        flips += int((torch.sign(sg[sel]) != torch.sign(gr[sel])).sum())
        total += int(sel.sum())
    return flips, total


def check_post_step_state(trainer, mg, md, og, od, grads, before, ref, cal_grads, lr=1e-4):
    """State after one full step (vae-gan.py:404-424): every buffer and every parameter.

    * BatchNorm ``num_batches_tracked`` exactly, ``running_mean`` / ``running_var`` within 2e-2 (statistics of bf16
      activations); spectral-norm ``weight_u`` / ``weight_v`` after the THREE discriminator calls of the step (the third
      one after ``opt_D.step()``) within 1e-2.
    * The optimiser kernels exactly: from the parameters before the step and OUR gradients (captured after each backward,
      before clipping), ``torch.nn.utils.clip_grad_norm_`` + ``torch.optim.Adam(lr, betas=(0.5, 0.999))`` must reproduce
      our updated parameters to 1e-7 absolute (1e-3 of one Adam step of size lr) -- a wrong sign, a missed clip or a wrong
      bias correction is 1e-4 away (asserted: 2e-7, i.e. one ulp of the largest parameters).
    * Against the oracle: the sign of every parameter's update (Adam's first step is -lr * sign(g)) must agree wherever the
      oracle's gradient element is above the noise between the two gradient evaluations."""
    from torch.nn.utils import clip_grad_norm_
    so_g, so_d = og.state_dict(), od.state_dict()
    for ours, theirs, tag in ((mg.state_dict(), so_g, "G"), (md.state_dict(), so_d, "D")):
        for k, v in theirs.items():
            leaf = k.rsplit(".", 1)[-1]
            if leaf == "num_batches_tracked":
                assert int(ours[k]) == int(v), (tag, k, int(ours[k]), int(v))
            elif leaf in ("running_mean", "running_var"):
                assert rel(ours[k], v) <= 2e-2, (tag, k, rel(ours[k], v))
            elif leaf in ("weight_u", "weight_v"):
                assert rel(ours[k], v) <= 1e-2, (tag, k, rel(ours[k], v))
    for which, opt, clip in (("D", trainer.opt_D, 0.0), ("G", trainer.opt_G, trainer.clip_norm)):
        ps, gs = [], []
        for p0, g in zip(before[which], grads[which]):
            q = torch.nn.Parameter(p0.clone())
            q.grad = g.clone() if g is not None else None
            ps.append(q)
        if clip > 0:
            clip_grad_norm_([q for q in ps if q.grad is not None], max_norm=clip)
        torch.optim.Adam(ps, lr=lr, betas=(0.5, 0.999)).step()
        worst = 0.0
        for q, p in zip(ps, opt.params):
            worst = max(worst, float((q.detach() - p.detach()).abs().max()))
        print(f"Adam/clip kernels vs torch on our gradients ({which}): max |dp| = {worst:.2e}")
        assert worst <= 2e-7, (which, worst)      # (one fp32 ulp of a BatchNorm weight near 1.0 is 1.2e-7)
    # bf16 operand shadows written by the Adam kernel == bf16(updated master weight), bit for bit
    from vae_gan_mark_b200 import layers as L
    mine, checked = {id(p) for p in trainer.opt_G.params}, 0
    for operand, refs, _ in L.SHADOWS.values():
        ps = [r() for r in refs]
        if any(q is None or id(q) not in mine for q in ps):
            continue
        rows = 0
        for q in ps:
            o = q.shape[0]
            assert torch.equal(operand[rows:rows + o], q.detach().permute(0, 2, 3, 1).reshape(o, -1).to(torch.bfloat16))
            rows += o
            checked += 1
    print(f"bf16 operand shadows checked: {checked}")
    assert checked > 0
    # sign of the update against the oracle (Adam's first step is -lr * sign(g)): the parameters must have moved the way
    # the oracle's gradient says wherever that gradient is above the noise; calibrated like the gradient bound itself by
    # the reference's own modules under torch bf16 autocast (batch-2 U-Net gradients are ~70 % noise in ANY bf16 evaluation)
    flips = total = cflips = ctotal = 0
    for which, opt, net in (("D", trainer.opt_D, md), ("G", trainer.opt_G, mg)):
        rgrads = ref.d_grads if which == "D" else ref.g_grads
        named = list(net.named_parameters())
        moved = [-(p.detach() - p0) if g is not None else None for (_, p), p0, g in zip(named, before[which], grads[which])]
        f, t = sign_flips(named, grads[which], rgrads, signs=moved)      # -(dp) has the sign of the gradient the update followed
        flips, total = flips + f, total + t
        f, t = sign_flips(named, cal_grads[which], rgrads)
        cflips, ctotal = cflips + f, ctotal + t
    frac, cfrac = flips / max(total, 1), cflips / max(ctotal, 1)
    print(f"update-sign agreement with the oracle: {total - flips} / {total} elements above the gradient noise "
          f"(flipped {frac:.2%}; torch bf16 autocast: {cfrac:.2%})")
    assert total > 0 and frac <= max(1e-2, 1.5 * cfrac), (flips, total, cfrac)


def ref_is_noise(key, ref):
    which, name = key.split(".", 1)
    g = (ref.d_grads if which == "D" else ref.g_grads)[name]
    return float(g.abs().max()) < 1e-7
