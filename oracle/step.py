"""Oracle restatement of the per-batch training step body (CPU, fp32).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Reference anchors: vae-gan.py:399-428, vae-gan-v2.py:696-748, vae-gan-unet.py:574-625 (the three
step bodies are the same code), hinge_loss at vae-gan.py:313-320, KL at vae-gan.py:420.
The VGG perceptual term (vae-gan.py:300-311,422) is NOT restated: its ImageNet weights are
unobtainable offline, so every comparison runs with perceptual weight 0 ("parity unpinned").
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from torch.nn.utils import clip_grad_norm_

TEXT_POOL = ["SALE", "New arrivals 2024", "Buy 1 get 1 FREE!", "50% off", "Limited time offer - today",
             "Free shipping on orders over $25", "Best price", "Hello, world", "Summer collection",
             "Subscribe & save", "Open 24/7", "Click here", "Black Friday deals start now", "Top rated",
             "Only 3 left in stock", "Thank you"]


@dataclass
class LossWeights:
    recon: float
    kl: float
    gan: float
    perc: float = 0.0      # PERC_WEIGHT (vae-gan.py:38); only used when train_step is given a perceptual module

    @staticmethod
    def for_family(family: str) -> "LossWeights":
        # vae-gan.py:35-38 ; vae-gan-v2.py:42-45 ; vae-gan-unet.py:43-46   (perceptual weight forced to 0)
        # vae-gan-oldv.py:40-44: KL 0.001, GAN 0.07
        return {"base": LossWeights(1.0, 0.005, 0.1), "v2": LossWeights(1.0, 0.001, 0.15),
                "unet": LossWeights(1.0, 0.001, 0.15), "oldv": LossWeights(1.0, 0.001, 0.07)}[family]


def hinge_loss(preds, target):
    """vae-gan.py:313-320."""
    if target == 1:
        return F.relu(1.0 - preds).mean()
    if target == 0:
        return F.relu(1.0 + preds).mean()
    return -preds.mean()


def kl_term(mu, logvar):
    """vae-gan.py:420."""
    return torch.mean(-0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp(), dim=[1, 2, 3]))


def synthetic_batch(batch: int, h: int, w: int, seed: int = 4321, step: int = 0):
    """uniform [0,1) images (what T.ToTensor yields), 0/1 mask, cycled strings."""
    g = torch.Generator().manual_seed(seed + 7919 * step)
    ru = torch.rand(batch, 3, h, w, generator=g)
    en = torch.rand(batch, 3, h, w, generator=g)
    mask = (torch.rand(batch, 1, h, w, generator=g) > 0.5).float()
    texts = [TEXT_POOL[(i + step) % len(TEXT_POOL)] for i in range(batch)]
    return ru, en, mask, texts


def deterministic_state(module: torch.nn.Module, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Fill every tensor of ``module.state_dict()`` from a generator seeded by (seed, key name).

    Independent of construction order and of torch's default-init RNG consumption, so the
    reference modules, this oracle and the CUDA drop-ins can all be given bit-identical weights
    with a plain ``load_state_dict``.  Scales mimic PyTorch's defaults (uniform +-1/sqrt(fan_in)).
    """
    out = {}
    for key, ref in module.state_dict().items():
        g = torch.Generator().manual_seed((seed * 1_000_003 + zlib.crc32(key.encode())) % (2 ** 63))
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t = torch.zeros_like(ref)
        elif leaf == "running_mean":
            t = 0.05 * torch.randn(ref.shape, generator=g)
        elif leaf == "running_var":
            t = 1.0 + 0.1 * torch.rand(ref.shape, generator=g)
        elif leaf in ("weight_u", "weight_v"):
            t = F.normalize(torch.randn(ref.shape, generator=g), dim=0, eps=1e-12)
        elif ref.dim() == 1 and leaf == "weight":          # BN / IN affine scale
            t = 1.0 + 0.1 * torch.randn(ref.shape, generator=g)
        elif ref.dim() == 1:                                # biases, BN/IN shift
            t = 0.05 * torch.randn(ref.shape, generator=g)
        else:
            fan_in = ref[0].numel() if ref.dim() > 1 else ref.numel()
            if "embedding" in key:
                t = torch.randn(ref.shape, generator=g)
                t[0].zero_()                                # padding_idx row
            else:
                bound = (3.0 / fan_in) ** 0.5
                t = (torch.rand(ref.shape, generator=g) * 2 - 1) * bound
        out[key] = t.to(ref.dtype)
    return out


@dataclass
class StepResult:
    losses: Dict[str, float]
    recon: torch.Tensor
    mu: torch.Tensor
    logvar: torch.Tensor
    d_grads: Dict[str, torch.Tensor]   # right after loss_D.backward()
    g_grads: Dict[str, torch.Tensor]   # right after loss_G.backward(), before clipping
    grad_norm: float


def train_step(G, D, opt_G, opt_D, batch, weights: LossWeights, seed: Optional[int] = None,
               clip_norm: float = 1.0, keep_grads: bool = True, perceptual=None) -> StepResult:
    """One iteration of the reference's hot loop (vae-gan.py:399-428) with perceptual weight 0.

    The D gradients are snapshotted right after ``loss_D.backward()`` because the reference's
    ``loss_G.backward()`` later accumulates wasted D-wgrads on top of them.
    """
    ru, en, mask, texts = batch
    if seed is not None:
        torch.manual_seed(seed)
    fake, mu, logvar = G(ru, mask, texts)

    opt_D.zero_grad()
    loss_d_real = hinge_loss(D(en), 1)
    loss_d_fake = hinge_loss(D(fake.detach()), 0)
    loss_d = (loss_d_real + loss_d_fake) * 0.5
    loss_d.backward()
    d_grads = {k: p.grad.detach().clone() for k, p in D.named_parameters()} if keep_grads else {}
    opt_D.step()

    opt_G.zero_grad()
    fake_preds = D(fake)
    recon = F.l1_loss(fake, en)
    kl = kl_term(mu, logvar)
    gan = hinge_loss(fake_preds, None)
    loss_g = weights.recon * recon + weights.kl * kl + weights.gan * gan
    perc = None
    if perceptual is not None and weights.perc != 0.0:          # vae-gan.py:422-423
        perc = perceptual(fake, en)
        loss_g = loss_g + weights.perc * perc
    loss_g.backward()
    g_grads = ({k: p.grad.detach().clone() for k, p in G.named_parameters() if p.grad is not None}
               if keep_grads else {})
    gn = clip_grad_norm_(G.parameters(), max_norm=clip_norm)
    opt_G.step()

    return StepResult(
        losses={"loss_G": float(loss_g), "loss_D": float(loss_d), "recon": float(recon), "kl": float(kl),
                "gan": float(gan), "d_real": float(loss_d_real), "d_fake": float(loss_d_fake),
                **({"perc": float(perc)} if perc is not None else {})},
        recon=fake.detach(), mu=mu.detach(), logvar=logvar.detach(), d_grads=d_grads, g_grads=g_grads,
        grad_norm=float(gn))


def make_optimizers(G, D, lr_g: float = 1e-4, lr_d: float = 1e-4):
    """vae-gan.py:541-542."""
    return (torch.optim.Adam(G.parameters(), lr=lr_g, betas=(0.5, 0.999)),
            torch.optim.Adam(D.parameters(), lr=lr_d, betas=(0.5, 0.999)))
