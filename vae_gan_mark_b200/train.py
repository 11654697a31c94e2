"""The per-batch VAE-GAN training step (reference: vae-gan.py:399-428, vae-gan-v2.py:696-748,
vae-gan-unet.py:574-625) on the CUDA kernels of this package.

``VAEGANTrainer.step`` reproduces the reference's step body operation for operation -- G forward, D step
(real + detached fake, hinge, backward, Adam), G step (D(fake), L1 + KL + hinge-G, backward,
clip_grad_norm_, Adam) -- with three deliberate differences that do not change any result:
  * the discriminator's weight gradients that the reference computes (and throws away) during
    ``loss_G.backward()`` are skipped;
  * loss scalars stay on the device (the reference's six ``.item()`` syncs per step are left to the caller);
  * the VGG perceptual term is not part of this path (its ImageNet weights are unavailable offline; weight 0).

``VAEGANTrainer.capture`` records the whole step (about 900 kernel launches) into one CUDA graph; ``replay`` then
runs a step with a single graph launch, which removes the host launch overhead that otherwise leaves the GPU idle
for ~20% of the step.
"""
from __future__ import annotations

import contextlib
import ctypes as C
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional

import torch

from . import _lib
from .dispatch import launch_op
from . import layers as L
from . import ops
from .ops import F32


@dataclass
class LossWeights:
    recon: float = 1.0
    kl: float = 0.001
    gan: float = 0.15
    perc: float = 0.0       # PERC_WEIGHT; needs a VGGPerceptual module (VAEGANTrainer(perceptual=...))

    @staticmethod
    def for_family(family: str, perceptual: bool = True) -> "LossWeights":
        """The reference's loss weights incl. PERC_WEIGHT (vae-gan.py:35-38 base 0.05; vae-gan-v2.py:42-45 and
        vae-gan-unet.py:43-46: 0.1; vae-gan-oldv.py:42-45: 0.2).  ``perceptual=False`` drops the VGG term (weight 0) --
        the configuration every parity test and the headline bench use, because the ImageNet weights of the reference's
        VGG16 cannot be obtained offline.  VAEGANTrainer refuses a non-zero ``perc`` without a VGGPerceptual module and
        a VGGPerceptual module with ``perc == 0``, so the term can never be dropped or added silently."""
        w = {"base": LossWeights(1.0, 0.005, 0.1, 0.05), "v2": LossWeights(1.0, 0.001, 0.15, 0.1),
             "unet": LossWeights(1.0, 0.001, 0.15, 0.1), "oldv": LossWeights(1.0, 0.001, 0.07, 0.2)}[family]
        if not perceptual:
            w.perc = 0.0
        return w


# The optimiser launches as dispatcher ops (dispatch.py).  ``table`` is the device copy of the VgAdamTensor table: the
# parameters, gradients, moments and operand shadows a launch reads / writes are reached through the raw pointers in it, so
# the schema names the table itself as the mutated argument.
@launch_op("vg_multi_sumsq(Tensor table, int count, Tensor(a!) out, Tensor(b!) scratch) -> ()")
def _multi_sumsq(table, count, out, scratch):
    _lib.call("vg_multi_sumsq", C.c_void_p(table.data_ptr()), count, C.c_void_p(out.data_ptr()), C.c_void_p(scratch.data_ptr()),
              scratch.numel(), ops.stream())


@launch_op("vg_adam_prepare(Tensor(a!) state, float beta1, float beta2) -> ()")
def _adam_prepare(state, beta1, beta2):
    _lib.call("vg_adam_prepare", C.c_void_p(state.data_ptr()), C.c_float(beta1), C.c_float(beta2), ops.stream())


@launch_op("vg_multi_adam(Tensor(a!) table, int count, float lr, float beta1, float beta2, float eps, Tensor state, "
           "Tensor? gnorm_sq, float max_norm, bool write_back_grad) -> ()")
def _multi_adam(table, count, lr, beta1, beta2, eps, state, gnorm_sq, max_norm, write_back_grad):
    _lib.call("vg_multi_adam", C.c_void_p(table.data_ptr()), count, C.c_float(lr), C.c_float(beta1), C.c_float(beta2),
              C.c_float(eps), C.c_void_p(state.data_ptr()), C.c_void_p(gnorm_sq.data_ptr() if gnorm_sq is not None else 0),
              C.c_float(max_norm), int(write_back_grad), ops.stream())


class FusedAdam:
    """Adam(lr, betas=(0.5, 0.999), eps=1e-8) with optional global-norm clipping, on our kernels
    (vae-gan.py:424,541-542): one multi-tensor launch for the norm, one for the update.  The step counter and bias
    corrections live on the device, so the update can be captured in a CUDA graph.  State layout mirrors
    torch.optim.Adam (step, exp_avg, exp_avg_sq per parameter)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1e-4, betas=(0.5, 0.999), eps=1e-8):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        self.lr, self.betas, self.eps = lr, betas, eps
        self.exp_avg = [torch.zeros_like(p, dtype=F32) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p, dtype=F32) for p in self.params]
        dev = self.params[0].device
        self.norm_sq = torch.zeros((), dtype=F32, device=dev)
        self.state = torch.zeros(4, dtype=F32, device=dev)          # {step, 1-b1^t, sqrt(1-b2^t), lr}
        self.state[3] = lr
        self._norm_scratch = torch.zeros(1024, dtype=F32, device=dev)
        self.shadows: List[Optional[torch.Tensor]] = [None] * len(self.params)   # bf16 operand copies the update keeps current
        self._table_host = torch.zeros((len(self.params), 6), dtype=torch.int64).pin_memory() if dev.type == "cuda" \
            else torch.zeros((len(self.params), 6), dtype=torch.int64)
        self._table_host_graph = torch.zeros_like(self._table_host)
        if dev.type == "cuda":
            self._table_host_graph = self._table_host_graph.pin_memory()
        self._table = torch.zeros((len(self.params), 6), dtype=torch.int64, device=dev)

    def attach_operand_shadows(self) -> int:
        """Give every eligible conv / conv-transpose weight (memory order [O][kh][kw][I], I % 64 == 0; the two full-kernel
        heads as one [2z][K] operand) a bf16 shadow that ``vg_multi_adam`` rewrites together with the fp32 master weight:
        the next step's forward finds its tensor-core operand ready instead of converting the weight again.  Returns the
        number of shadowed tensors."""
        count = 0
        for i, p in enumerate(self.params):
            if self.shadows[i] is None and L.shadow_eligible(p):
                o, c, kh, kw = p.shape
                sh = torch.empty((o, kh * kw * c), dtype=torch.bfloat16, device=p.device)
                L.register_shadow(p, sh)
                self.shadows[i] = sh
                count += 1
        return count

    def attach_heads_shadow(self, w_mu: torch.nn.Parameter, w_lv: torch.nn.Parameter) -> bool:
        """mu_head / logvar_head run as ONE GEMM with N = 2z: their operand is one [2z][kh*kw*C] matrix, so their two
        shadows are the halves of one buffer."""
        idx = {id(p): i for i, p in enumerate(self.params)}
        if id(w_mu) not in idx or id(w_lv) not in idx or not (L.shadow_eligible(w_mu) and L.shadow_eligible(w_lv)):
            return False
        z, c, kh, kw = w_mu.shape
        buf = torch.empty((2 * z, kh * kw * c), dtype=torch.bfloat16, device=w_mu.device)
        # the pair's entry serves HeadsFn; the single-tensor entries created by attach_operand_shadows (if any) are replaced
        L.SHADOWS.pop(w_mu.data_ptr(), None)
        L.SHADOWS.pop(w_lv.data_ptr(), None)
        L.register_shadow((w_mu, w_lv), buf)
        self.shadows[idx[id(w_mu)]], self.shadows[idx[id(w_lv)]] = buf[:z], buf[z:]
        return True

    @property
    def step_count(self) -> int:
        return int(self.state[0].item())

    def set_lr(self, lr: float) -> None:
        """Change the learning rate.  The kernels read it from the device state block, so this also takes effect in an
        already captured CUDA graph of the step (LR schedulers of vae-gan-lr-sh.py / vae-gan-v2.py)."""
        self.lr = float(lr)
        self.state[3] = self.lr

    @property
    def param_groups(self):
        """Read-only torch.optim-style view (one group) for code that inspects ``optimizer.param_groups[0]['lr']``."""
        return [{"lr": self.lr, "betas": self.betas, "eps": self.eps, "params": self.params}]

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            p.grad = None

    def _upload_table(self):
        # the pinned staging table is read by an async H2D copy: do not rewrite it before the previous copy has run
        ev = self.__dict__.get("_table_event")
        if ev is not None and not torch.cuda.is_current_stream_capturing():
            ev.synchronize()
        t = self._table_host
        if torch.cuda.is_current_stream_capturing():
            # a captured copy re-reads its host source on every replay: give the graph a table of its own that eager
            # steps never overwrite
            t = self._table_host_graph
        for i, (p, m, v) in enumerate(zip(self.params, self.exp_avg, self.exp_avg_sq)):
            g = p.grad
            if g is not None and g.stride() != p.stride():
                # the kernels walk p, g, m, v as flat arrays: the gradient must share the parameter's memory order
                # (OIHW, or channels_last for conv weights -- see weights_channels_last)
                g = torch.empty_like(p).copy_(g)
                p.grad = g
            t[i, 0], t[i, 1] = p.data_ptr(), (g.data_ptr() if g is not None else 0)
            t[i, 2], t[i, 3], t[i, 4] = m.data_ptr(), v.data_ptr(), p.numel()
            t[i, 5] = self.shadows[i].data_ptr() if self.shadows[i] is not None else 0
        self._table.copy_(t, non_blocking=True)
        if not torch.cuda.is_current_stream_capturing():
            self._table_event = torch.cuda.Event()
            self._table_event.record()

    def grad_norm_sq(self) -> torch.Tensor:
        self._upload_table()
        _multi_sumsq(self._table, len(self.params), self.norm_sq, self._norm_scratch)
        return self.norm_sq

    def step(self, max_norm: float = 0.0):
        """One update; when ``max_norm`` > 0 gradients are first scaled by min(1, max_norm/(||g||+1e-6)) like
        torch.nn.utils.clip_grad_norm_ (the scaled gradients are written back, as the reference's in-place clip does)."""
        if max_norm > 0:
            self.grad_norm_sq()
        else:
            self._upload_table()
        _adam_prepare(self.state, self.betas[0], self.betas[1])
        _multi_adam(self._table, len(self.params), -1.0, self.betas[0], self.betas[1], self.eps, self.state,   # lr < 0: state[3]
                    self.norm_sq if max_norm > 0 else None, max_norm, max_norm > 0)
        L.bump_weight_epoch()

    def state_dict(self) -> Dict:
        """torch.optim.Adam-compatible state_dict (checkpoint contract of vae-gan.py:449-456)."""
        step = torch.tensor(float(self.step_count))
        state = {i: {"step": step.clone(), "exp_avg": m, "exp_avg_sq": v}
                 for i, (m, v) in enumerate(zip(self.exp_avg, self.exp_avg_sq))}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: Dict):
        step = 0
        for i, st in sd["state"].items():
            self.exp_avg[int(i)].copy_(st["exp_avg"])
            self.exp_avg_sq[int(i)].copy_(st["exp_avg_sq"])
            step = int(st["step"])
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps = g["lr"], tuple(g["betas"]), g["eps"]
        b1, b2 = self.betas
        self.state.copy_(torch.tensor([float(step), 1.0 - b1 ** step, (1.0 - b2 ** step) ** 0.5, self.lr]))


class ReduceLROnPlateau:
    """torch.optim.lr_scheduler.ReduceLROnPlateau for ``FusedAdam`` (the reference builds one per optimiser:
    vae-gan-lr-sh.py:751-760, vae-gan-v2.py:944-953, and steps it with the epoch's validation loss, :632-633 / :796-797).
    Same algorithm, defaults and ``state_dict`` keys as the torch class (threshold_mode 'rel' | 'abs', cooldown, eps), so
    the reference's ``scheduler_G_state_dict`` checkpoints load; the new rate goes to ``FusedAdam.set_lr`` and therefore
    into a captured graph without re-capturing."""

    def __init__(self, optimizer: FusedAdam, mode="min", factor=0.1, patience=10, threshold=1e-4, threshold_mode="rel",
                 cooldown=0, min_lr=0.0, eps=1e-8):
        assert factor < 1.0 and mode in ("min", "max") and threshold_mode in ("rel", "abs")
        self.optimizer = optimizer
        self.mode, self.factor, self.patience, self.threshold, self.threshold_mode = mode, factor, patience, threshold, threshold_mode
        self.cooldown, self.min_lrs, self.eps = cooldown, [min_lr], eps
        self.best = float("inf") if mode == "min" else -float("inf")
        self.num_bad_epochs, self.cooldown_counter, self.last_epoch = 0, 0, 0
        self._last_lr = [optimizer.lr]

    def _is_better(self, a: float) -> bool:
        if self.mode == "min":
            return a < (self.best * (1.0 - self.threshold) if self.threshold_mode == "rel" else self.best - self.threshold)
        return a > (self.best * (self.threshold + 1.0) if self.threshold_mode == "rel" else self.best + self.threshold)

    def step(self, metrics) -> None:
        current = float(metrics)
        self.last_epoch += 1
        if self._is_better(current):
            self.best, self.num_bad_epochs = current, 0
        else:
            self.num_bad_epochs += 1
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.num_bad_epochs = 0
        if self.num_bad_epochs > self.patience:
            old = self.optimizer.lr
            new = max(old * self.factor, self.min_lrs[0])
            if old - new > self.eps:
                self.optimizer.set_lr(new)
            self.cooldown_counter, self.num_bad_epochs = self.cooldown, 0
        self._last_lr = [self.optimizer.lr]

    def get_last_lr(self):
        return self._last_lr

    def state_dict(self) -> Dict:
        return {k: v for k, v in self.__dict__.items() if k != "optimizer"}

    def load_state_dict(self, sd: Dict) -> None:
        self.__dict__.update({k: v for k, v in sd.items() if k != "optimizer"})
        if self._last_lr:
            self.optimizer.set_lr(self._last_lr[0])


def kl_anneal_weight(epoch: int, start: float = 1e-7, target: float = 0.001, anneal_epochs: int = 20) -> float:
    """KL weight of an epoch (vae-gan-v2.py:48-49,1001-1004; vae-gan-oldv.py:1043-1045): linear from ``start`` to
    ``target`` over ``anneal_epochs`` epochs, then ``target``.  Feed it to ``VAEGANTrainer.set_kl_weight``."""
    if epoch < anneal_epochs:
        return start + (target - start) * (epoch / max(1, anneal_epochs - 1))
    return target


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


def weights_channels_last(module: torch.nn.Module) -> int:
    """Store the 4-D conv / conv-transpose weights of ``module`` in channels_last memory order ([O][kh][kw][I]).

    Shapes, dtypes and state_dict keys do not change (``load_state_dict`` / ``state_dict`` are oblivious to strides),
    but this is the order the tensor core consumes and produces: the forward operand becomes a plain fp32 -> bf16
    conversion, and the weight-gradient kernel's [O][kh][kw][I] result IS the gradient -- autograd adopts it without
    a re-layout copy.  Spectral-norm weights keep the OIHW order their u / v vectors are defined in.  Returns the
    number of tensors converted."""
    count = 0
    for name, p in module.named_parameters():
        if p.dim() == 4 and name.endswith("weight") and p.shape[2] * p.shape[3] > 1 and p.shape[1] > 1:
            if not p.data.is_contiguous(memory_format=torch.channels_last):
                p.data = p.data.contiguous(memory_format=torch.channels_last)
                if p.grad is not None:
                    p.grad = None
                count += 1
    return count


class VAEGANTrainer:
    def __init__(self, G: torch.nn.Module, D: torch.nn.Module, weights: LossWeights, lr_g=1e-4, lr_d=1e-4,
                 clip_norm: float = 1.0, grad_hook=None, channels_last_weights: bool = True, perceptual=None):
        """``perceptual``: an optional ``modules.VGGPerceptual`` (frozen); its loss enters loss_G with weight
        ``weights.perc`` (vae-gan.py:422-423)."""
        self.G, self.D, self.w, self.clip_norm = G, D, weights, clip_norm
        self.perceptual = perceptual
        if perceptual is None and weights.perc != 0.0:
            raise ValueError(f"LossWeights.perc = {weights.perc} but no VGGPerceptual module was given: pass "
                             "perceptual=modules.VGGPerceptual() (with the VGG16 weights loaded) or use "
                             "LossWeights.for_family(family, perceptual=False)")
        if perceptual is not None and weights.perc == 0.0:
            raise ValueError("a VGGPerceptual module was given but LossWeights.perc is 0: the perceptual term would be "
                             "dropped silently (reference PERC_WEIGHT: 0.05 base, 0.1 v2 / unet, 0.2 oldv)")
        if channels_last_weights:
            weights_channels_last(G)
            weights_channels_last(D)
        self.opt_G = FusedAdam(G.parameters(), lr=lr_g)
        self.opt_D = FusedAdam(D.parameters(), lr=lr_d)
        if channels_last_weights and next(G.parameters()).is_cuda:
            # bf16 operand shadows written by the Adam kernel (generator only: D's convs are spectrally normalised, their
            # operand W / sigma changes with every call)
            for m in G.modules():
                mu, lv = getattr(m, "mu_head", None), getattr(m, "logvar_head", None)
                if mu is not None and lv is not None:
                    self.opt_G.attach_heads_shadow(mu.weight, lv.weight)
            self.opt_G.attach_operand_shadows()
        # KL weight as a device scalar: the reference anneals it per epoch (vae-gan-v2.py:1002-1004); a captured graph
        # reads the current value, see set_kl_weight
        self.kl_weight = torch.full((), float(weights.kl), dtype=F32, device=self.opt_G.state.device)
        self.grad_hook = grad_hook        # called as grad_hook("D"|"G", params) after each backward (DP allreduce)
        # data parallel: SMs the persistent tensor-core grids may occupy during loss_G.backward() (0 = all).  The conv
        # grids otherwise own every SM, so NCCL's CTAs only get in at kernel boundaries; leaving a few SMs free lets the
        # bucketed all-reduces of the big generator run concurrently with the rest of the backward pass
        self.backward_sm_limit = 0
        self._graph = None
        self.sched_G = self.sched_D = None

    def attach_schedulers(self, sched_G: Optional["ReduceLROnPlateau"] = None, sched_D: Optional["ReduceLROnPlateau"] = None):
        """LR schedulers whose state travels with the checkpoint as ``scheduler_G_state_dict`` / ``scheduler_D_state_dict``
        (vae-gan-lr-sh.py:643-644,784-787; vae-gan-v2.py:807-808,974-977)."""
        self.sched_G, self.sched_D = sched_G, sched_D

    def step(self, ru, en, mask, texts, kl_weight: Optional[float] = None) -> Dict[str, torch.Tensor]:
        G, D, w = self.G, self.D, self.w
        klw = self.kl_weight if kl_weight is None else kl_weight
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
            perc = None
            if self.perceptual is not None and w.perc != 0.0:
                perc = self.perceptual(fake, en)
                loss_g = loss_g + w.perc * perc
            if self.backward_sm_limit:
                _lib.call("vg_set_conv_sm_limit", int(self.backward_sm_limit))
            loss_g.backward()
        _lib.call("vg_set_conv_sm_limit", 0)      # (only ever raised by the optional side-stream text path)
        if self.grad_hook is not None:
            self.grad_hook("G", self.opt_G.params)
        self.opt_G.step(max_norm=self.clip_norm)
        out = {"loss_G": loss_g.detach(), "loss_D": loss_d.detach(), "recon": recon.detach(), "kl": kl.detach(),
               "gan": gan.detach(), "d_real": loss_d_real.detach(), "d_fake": loss_d_fake.detach(),
               "grad_norm_sq": self.opt_G.norm_sq, "fake": fake.detach(), "mu": mu.detach(), "logvar": logvar.detach()}
        if perc is not None:
            out["perc"] = perc.detach()
        return out

    def set_kl_weight(self, value: float) -> None:
        """KL annealing (vae-gan-v2.py:1002-1004): takes effect in eager steps and in an already captured graph."""
        self.w.kl = float(value)
        self.kl_weight.fill_(float(value))

    # ------------------------------------------------------------------ checkpoint / resume
    def checkpoint(self, **extra) -> Dict:
        """The reference's checkpoint dictionary (vae-gan-v2.py:802-808; vae-gan.py:449-456): module state_dicts plus
        torch.optim.Adam-compatible optimiser state_dicts.  ``extra`` entries (epoch, scheduler states, ...) are passed
        through.  Save with ``torch.save``; a checkpoint written by the reference loads with ``load_checkpoint``."""
        ck = {"model_state_dict": self.G.state_dict(), "disc_state_dict": self.D.state_dict(),
              "opt_G_state_dict": self.opt_G.state_dict(), "opt_D_state_dict": self.opt_D.state_dict()}
        if self.sched_G is not None:
            ck["scheduler_G_state_dict"] = self.sched_G.state_dict()
        if self.sched_D is not None:
            ck["scheduler_D_state_dict"] = self.sched_D.state_dict()
        ck.update(extra)
        return ck

    def load_checkpoint(self, ck: Dict, strict: bool = False) -> None:
        """Inverse of ``checkpoint`` (the reference resumes with strict=False, vae-gan-v2.py:968-972).  Keeps the memory
        order of the parameters, invalidates the cached operand layouts, and drops a captured graph's claim to be
        current -- call ``capture`` again before ``replay``."""
        self.G.load_state_dict(ck["model_state_dict"], strict=strict)
        self.D.load_state_dict(ck["disc_state_dict"], strict=strict)
        opt_g = ck.get("opt_G_state_dict", ck.get("optG_state_dict"))
        opt_d = ck.get("opt_D_state_dict", ck.get("optD_state_dict"))
        if opt_g is not None:
            self.opt_G.load_state_dict(opt_g)
        if opt_d is not None:
            self.opt_D.load_state_dict(opt_d)
        # like the reference, the schedulers are restored after the optimisers (their last rate wins)
        if self.sched_G is not None and "scheduler_G_state_dict" in ck:
            self.sched_G.load_state_dict(ck["scheduler_G_state_dict"])
        if self.sched_D is not None and "scheduler_D_state_dict" in ck:
            self.sched_D.load_state_dict(ck["scheduler_D_state_dict"])
        L.bump_weight_epoch()
        self._graph = None

    # ------------------------------------------------------------------ CUDA graph
    def capture(self, ru, en, mask, texts, warmup: int = 3, kl_weight: Optional[float] = None):
        """Record one full step into a CUDA graph.  ``texts`` is tokenised / embedded once here (the graph takes the
        resulting device tensor as a static input; call ``set_texts`` to change it between replays)."""
        G = self.G
        self._static = [ru.clone(), en.clone(), mask.clone()]
        self._static_text = self._encode_texts(texts)
        L.bump_weight_epoch()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.step(*self._static, self._static_text, kl_weight)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        L.bump_weight_epoch()
        self._graph = torch.cuda.CUDAGraph()
        n0 = _lib.lib().vg_launch_count()
        with torch.cuda.graph(self._graph):
            self._static_out = self.step(*self._static, self._static_text, kl_weight)
        self.launches_per_step = int(_lib.lib().vg_launch_count() - n0)   # our kernels recorded in the graph
        return self._static_out

    def _encode_texts(self, texts, to_device: bool = True):
        """Host-side part of the text path (code units of the strings / sentence embedding), hoisted out of the graph."""
        G = self.G
        if torch.is_tensor(texts):
            return texts
        enc = getattr(G, "char_text_encoder_module", None)
        if enc is not None:
            # UTF-32 code units (host, C speed); the lookup-table tokeniser kernel runs inside the step / the graph
            dev = next(G.parameters()).device
            host = enc.codepoints(texts, 60, pin=dev.type == "cuda")
            return host.to(dev, non_blocking=True) if to_device else host
        te = G.text_encoder
        with torch.no_grad():
            return te._embed(texts).to(next(G.parameters()).device, F32).clone()

    def set_texts(self, texts):
        """New strings for the next replays: host -> code units -> one pinned H2D copy into the graph's static buffer."""
        self._static_text.copy_(self._encode_texts(texts, to_device=False), non_blocking=True)

    def replay(self, ru=None, en=None, mask=None) -> Dict[str, torch.Tensor]:
        """Run one captured step; new inputs (device or pinned-host tensors) are copied into the graph's static buffers."""
        for dst, src in zip(self._static, (ru, en, mask)):
            if src is not None and src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        # the captured Adam kernels write the weights through raw pointers (no autograd version bump): cached bf16
        # operands of an eager forward before this replay (a validation pass) are stale from here on
        L.bump_weight_epoch()
        return self._static_out
