"""Pin oracle/ against the fixtures recorded from the reference's own modules (tests/golden/make_golden.py).

CPU only.  Tolerance: fp32, rtol 1e-5 / atol 1e-6 on every recorded quantity -- the oracle runs the same
torch ops in the same order as the reference, so it is expected to agree to rounding.
"""
import os

import pytest
import torch

from oracle import models as om
from oracle.step import LossWeights, deterministic_state, make_optimizers, synthetic_batch, train_step

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["base_32x32_b4", "base_64x64_b16", "v2_32x64_b2", "v2_32x32_b3_z32", "unet_32x32_b2", "oldv_32x64_b2",
         "v2_128x128_b8", "unet_256x256_b2"]      # the last two: the benchmark's own image sizes, one step each


def summarize(t, n=6):
    f = t.detach().double().flatten()
    head = torch.zeros(n, dtype=torch.float64)
    head[:min(n, f.numel())] = f[:n]
    return torch.cat([torch.stack([f.norm(), f.sum()]), head])


def build_oracle(family, h, w, z):
    if family == "base":
        G = om.VAEGAN(4, z, 64, 3, patch_hw=(h, w))
    elif family == "v2":
        G = om.VAEGAN_UNet_SpatialFiLM(4, z, patch_hw=(h, w))
    elif family == "oldv":
        G = om.VAEGAN_UNet_SpatialFiLM_OldV(4, z, patch_hw=(h, w))
    else:
        G = om.VAEGAN_UNet_CharEmb(4, z, patch_hw=(h, w), repaired=True)
    return G, om.Discriminator(3)


def close(a, b, rtol=1e-5, atol=1e-6):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    scale = max(float(b.abs().max()), 1e-30)
    return bool(((a - b).abs() <= atol * max(scale, 1.0) + rtol * b.abs().clamp_min(scale * 1e-3)).all())


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(case):
    gold = torch.load(os.path.join(GOLD, case + ".pt"), weights_only=False)
    family, h, w, batch, z = gold["family"], gold["h"], gold["w"], gold["batch"], gold["z"]
    torch.set_num_threads(8)
    G, D = build_oracle(family, h, w, z)
    # strict=True load proves the state_dict key/shape contract
    G.load_state_dict(deterministic_state(G, 1234), strict=True)
    D.load_state_dict(deterministic_state(D, 4321), strict=True)
    G.train(); D.train()
    if gold.get("gru_dropout") is not None:
        G.char_text_encoder_module.rnn.dropout = gold["gru_dropout"]
    opt_G, opt_D = make_optimizers(G, D)
    wts = LossWeights.for_family(family)
    for step, rec in enumerate(gold["steps"]):
        out = train_step(G, D, opt_G, opt_D, synthetic_batch(batch, h, w, step=step), wts, seed=10_000 + step)
        for k, v in rec["losses"].items():
            assert close(out.losses[k], v), (case, step, k, out.losses[k], v)
        assert close(out.grad_norm, rec["grad_norm"], rtol=1e-4)
        assert close(out.mu, rec["mu"]) and close(out.logvar, rec["logvar"])
        assert close(summarize(out.recon, 16), rec["recon_sum"])
        if rec["recon_img"] is not None:
            assert close(out.recon, rec["recon_img"])
        assert set(out.d_grads) == set(rec["d_grads"]) and set(out.g_grads) == set(rec["g_grads"])
        for k, v in rec["d_grads"].items():
            assert close(summarize(out.d_grads[k]), v, rtol=2e-4, atol=1e-5), (case, step, "D", k)
        for k, v in rec["g_grads"].items():
            assert close(summarize(out.g_grads[k]), v, rtol=2e-4, atol=1e-5), (case, step, "G", k)
        for k, v in rec["G_state"].items():
            assert close(summarize(G.state_dict()[k].float()), v, rtol=1e-4, atol=1e-5), (case, step, k)
        for k, v in rec["D_state"].items():
            assert close(summarize(D.state_dict()[k].float()), v, rtol=1e-4, atol=1e-5), (case, step, k)
    G.eval(); D.eval()
    ru, en, mask, texts = synthetic_batch(batch, h, w, step=7)
    with torch.no_grad():
        torch.manual_seed(77)
        fake, mu, _ = G(ru, mask, texts)
        assert close(summarize(fake, 16), gold["eval"]["recon_sum"])
        assert close(mu, gold["eval"]["mu"])
        assert close(D(en), gold["eval"]["d_out"])


def test_unet_shipped_forward_is_unrunnable():
    """the reference's own vae-gan-unet.py forward raises; so does the faithful oracle."""
    gold = torch.load(os.path.join(GOLD, "unet_32x32_b2.pt"), weights_only=False)
    assert gold["shipped_forward_error"] is not None
    G = om.VAEGAN_UNet_CharEmb(4, 128, patch_hw=(32, 32), repaired=False)
    ru, en, mask, texts = synthetic_batch(2, 32, 32)
    with pytest.raises(RuntimeError):
        G(ru, mask, texts)


def test_oracle_vgg_features_match_torchvision():
    """oracle.models.VGGPerceptual.features restates torchvision's vgg16().features[:16] (what vae-gan.py:303-304
    builds): same state_dict keys, and identical outputs for the same (random) weights."""
    import torch
    torchvision = pytest.importorskip("torchvision")
    from oracle import models as om
    torch.manual_seed(0)
    tv = torchvision.models.vgg16(weights=None).features[:16].eval()
    ours = om.VGGPerceptual()
    assert list(tv.state_dict().keys()) == list(ours.features.state_dict().keys())
    ours.features.load_state_dict(tv.state_dict())
    x = torch.rand(2, 3, 32, 48)
    assert torch.equal(ours.features(x), tv(x))
    fake, real = torch.rand(2, 3, 32, 32), torch.rand(2, 3, 32, 32)
    norm = lambda t: (t - ours.mean) / ours.std   # noqa: E731
    want = torch.nn.functional.l1_loss(tv(norm(fake)), tv(norm(real)))
    assert torch.allclose(ours(fake, real), want, rtol=0, atol=0)
