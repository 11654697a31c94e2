"""The per-batch VAE-GAN training step (reference: vae-gan.py:399-428, vae-gan-v2.py:696-748,
vae-gan-unet.py:574-625) on the CUDA kernels of this package.

``VAEGANTrainer.step`` reproduces the reference's step body operation for operation -- G forward, D step
(real + detached fake, hinge, backward, Adam), G step (D(fake), L1 + KL + hinge-G, backward,
clip_grad_norm_, Adam) -- with three deliberate differences that do not change any result:
  * the discriminator's weight gradients that the reference computes (and throws away) during
    ``loss_G.backward()`` are skipped;
  * loss scalars stay on the device (the reference's six ``.item()`` syncs per step are left to the caller);
  * the VGG perceptual term is not part of this path (its ImageNet weights are unavailable offline; weight 0).
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional

import torch

from . import layers as L
from . import ops
from .ops import F32


@dataclass
class LossWeights:
    recon: float = 1.0
    kl: float = 0.001
    gan: float = 0.15

    @staticmethod
    def for_family(family: str) -> "LossWeights":
        # vae-gan.py:35-38 ; vae-gan-v2.py:42-45 ; vae-gan-unet.py:43-46
        return {"base": LossWeights(1.0, 0.005, 0.1), "v2": LossWeights(1.0, 0.001, 0.15),
                "unet": LossWeights(1.0, 0.001, 0.15)}[family]


class FusedAdam:
    """Adam(lr, betas=(0.5, 0.999), eps=1e-8) with optional global-norm clipping, on our kernels
    (vae-gan.py:424,541-542).  State layout mirrors torch.optim.Adam (step, exp_avg, exp_avg_sq per parameter)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1e-4, betas=(0.5, 0.999), eps=1e-8):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        self.lr, self.betas, self.eps = lr, betas, eps
        self.step_count = 0
        self.exp_avg = [torch.zeros_like(p, dtype=F32) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p, dtype=F32) for p in self.params]
        dev = self.params[0].device
        self.norm_sq = torch.zeros((), dtype=F32, device=dev)

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            p.grad = None

    def grad_norm_sq(self) -> torch.Tensor:
        first = True
        for p in self.params:
            if p.grad is not None:
                ops.sumsq(p.grad, self.norm_sq, zero_first=first)
                first = False
        if first:
            self.norm_sq.zero_()
        return self.norm_sq

    def step(self, max_norm: float = 0.0):
        """One update; when ``max_norm`` > 0 gradients are first scaled by min(1, max_norm/(||g||+1e-6)) like
        torch.nn.utils.clip_grad_norm_ (the scaled gradients are written back, as the reference's in-place clip does)."""
        self.step_count += 1
        nrm = self.grad_norm_sq() if max_norm > 0 else None
        for p, m, v in zip(self.params, self.exp_avg, self.exp_avg_sq):
            if p.grad is None:
                continue
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            ops.adam_step(p.data, g, m, v, self.lr, self.betas[0], self.betas[1], self.eps, self.step_count, nrm,
                          max_norm, write_back_grad=max_norm > 0)
        L.bump_weight_epoch()

    def state_dict(self) -> Dict:
        """torch.optim.Adam-compatible state_dict (checkpoint contract of vae-gan.py:449-456)."""
        state = {i: {"step": torch.tensor(float(self.step_count)), "exp_avg": m, "exp_avg_sq": v}
                 for i, (m, v) in enumerate(zip(self.exp_avg, self.exp_avg_sq))}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: Dict):
        for i, st in sd["state"].items():
            self.exp_avg[int(i)].copy_(st["exp_avg"])
            self.exp_avg_sq[int(i)].copy_(st["exp_avg_sq"])
            self.step_count = int(st["step"])
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps = g["lr"], tuple(g["betas"]), g["eps"]


@contextlib.contextmanager
def frozen(module: torch.nn.Module):
    """Temporarily stop gradient computation for a module's parameters (D during the G step)."""
    flags = [p.requires_grad for p in module.parameters()]
    for p in module.parameters():
        p.requires_grad_(False)
    try:
        yield
    finally:
        for p, f in zip(module.parameters(), flags):
            p.requires_grad_(f)


class VAEGANTrainer:
    def __init__(self, G: torch.nn.Module, D: torch.nn.Module, weights: LossWeights, lr_g=1e-4, lr_d=1e-4,
                 clip_norm: float = 1.0, grad_hook=None):
        self.G, self.D, self.w, self.clip_norm = G, D, weights, clip_norm
        self.opt_G = FusedAdam(G.parameters(), lr=lr_g)
        self.opt_D = FusedAdam(D.parameters(), lr=lr_d)
        self.grad_hook = grad_hook        # called as grad_hook("D"|"G", params) after each backward (DP allreduce)

    def step(self, ru, en, mask, texts, kl_weight: Optional[float] = None) -> Dict[str, torch.Tensor]:
        G, D, w = self.G, self.D, self.w
        klw = w.kl if kl_weight is None else kl_weight
        fake, mu, logvar = G(ru, mask, texts)
        kl = G.__dict__["_last_kl"]

        # ---- discriminator step (vae-gan.py:408-414) ----
        self.opt_D.zero_grad()
        loss_d_real = L.hinge_loss(D(en), 1)
        loss_d_fake = L.hinge_loss(D(fake.detach()), 0)
        loss_d = (loss_d_real + loss_d_fake) * 0.5
        loss_d.backward()
        if self.grad_hook is not None:
            self.grad_hook("D", self.opt_D.params)
        self.opt_D.step()

        # ---- generator step (vae-gan.py:417-424) ----
        self.opt_G.zero_grad()
        with frozen(D):
            fake_preds = D(fake)
            recon = L.l1_loss(fake, en)
            gan = L.hinge_loss(fake_preds, None)
            loss_g = w.recon * recon + klw * kl + w.gan * gan
            loss_g.backward()
        if self.grad_hook is not None:
            self.grad_hook("G", self.opt_G.params)
        self.opt_G.step(max_norm=self.clip_norm)
        return {"loss_G": loss_g.detach(), "loss_D": loss_d.detach(), "recon": recon.detach(), "kl": kl.detach(),
                "gan": gan.detach(), "d_real": loss_d_real.detach(), "d_fake": loss_d_fake.detach(),
                "grad_norm_sq": self.opt_G.norm_sq, "fake": fake.detach(), "mu": mu.detach(), "logvar": logvar.detach()}
