"""Teacher-forcing harness for the layer-wise parity tests (test infrastructure, CPU side).

One real training step of the ORACLE (oracle/step.py, pinned to the reference by tests/test_oracle_golden.py) is run with
hooks that record, for every layer, the activation that entered it and the gradient that arrived at its output.  Each
layer ("unit") can then be re-evaluated in isolation -- by the oracle's own torch module here, and by the CUDA kernels in
tests/test_layer_parity_gpu.py -- on exactly those tensors, so that a kernel is checked at the benchmark's shapes on the
data of a real step without the errors of the layers before it (which is what makes whole-step gradient comparisons in
bf16 so loose: one flipped ReLU mask upstream moves every gradient downstream).

Units (reference file:line in the oracle's docstrings):
  conv      nn.Conv2d / nn.ConvTranspose2d (incl. spectral-norm convs of D, first call of the step)
  norm      BatchNorm2d -> ReLU [-> MaxPool2d(2)]   /   InstanceNorm2d -> LeakyReLU(0.2)
  film      gamma * x + beta of SpatialFiLMLayer (vae-gan-v2.py:146-149) and the bilinear resize in front of it (:138-140)
  text      the whole CharacterTokenEncoder (vae-gan-v2.py:65-114)
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import models as om
from oracle.step import LossWeights, deterministic_state, make_optimizers, synthetic_batch, train_step


def build_oracle(family: str, h: int, w: int, z: int):
    if family == "base":
        G = om.VAEGAN(4, z, 64, 3, patch_hw=(h, w))
    elif family == "v2":
        G = om.VAEGAN_UNet_SpatialFiLM(4, z, patch_hw=(h, w))
    elif family == "oldv":
        G = om.VAEGAN_UNet_SpatialFiLM_OldV(4, z, patch_hw=(h, w))
    else:
        G = om.VAEGAN_UNet_CharEmb(4, z, patch_hw=(h, w), repaired=True)
    D = om.Discriminator(3)
    for m in (G, D):
        m.train()
    if hasattr(G, "char_text_encoder_module"):
        G.char_text_encoder_module.rnn.dropout = 0.0      # GRU dropout draws from different RNGs on CPU and CUDA
    return G, D


class Tape:
    """Per module name: a list (one entry per call) of {"x": input, "gy": grad at the output, "gx": grad at the input,
    "extra": further positional inputs}."""

    def __init__(self):
        self.rec: Dict[str, List[dict]] = {}
        self.handles = []

    def watch(self, root: nn.Module, prefix: str, types, want_x=True, want_gx=False):
        for name, m in root.named_modules():
            if isinstance(m, types):
                self.handles.append(m.register_forward_hook(self._hook(prefix + name, want_x, want_gx)))

    def _hook(self, key, want_x, want_gx):
        def fn(mod, inp, out):
            if not torch.is_grad_enabled():
                return
            e: dict = {}
            self.rec.setdefault(key, []).append(e)
            if want_x:
                e["x"] = inp[0].detach().clone() if torch.is_tensor(inp[0]) else inp[0]
                e["extra"] = [t.detach().clone() if torch.is_tensor(t) else t for t in inp[1:]]
            if torch.is_tensor(out) and out.requires_grad:
                out.register_hook(lambda g, e=e: e.__setitem__("gy", g.detach().clone()))
            if want_gx and torch.is_tensor(inp[0]) and inp[0].requires_grad:
                inp[0].register_hook(lambda g, e=e: e.__setitem__("gx", g.detach().clone()))
        return fn

    def close(self):
        for h in self.handles:
            h.remove()
        self.handles = []


@dataclass
class Unit:
    kind: str                       # conv | norm | film | upsample | text
    net: str                        # "G" | "D"
    name: str                       # module name inside the net (the conv / the norm layer / the FiLM layer / text encoder)
    x: object = None                # input activation (NCHW fp32) or the list of strings
    gy: Optional[torch.Tensor] = None       # gradient at the unit's (full-resolution) output
    gpool: Optional[torch.Tensor] = None    # norm units with a fused max-pool: gradient at the pooled output
    x2: Optional[torch.Tensor] = None       # film: x_main (x is the text map)
    act: int = 0                    # 0 none, 1 ReLU, 2 LeakyReLU(0.2) -- the activation that belongs to the unit
    pool: bool = False
    meta: dict = field(default_factory=dict)


def _seq_next(root: nn.Module, name: str, offset: int = 1):
    """(name, module) of the sibling `offset` places after `name` inside its nn.Sequential parent, or (None, None)."""
    if "." not in name:
        return None, None
    parent_name, leaf = name.rsplit(".", 1)
    parent = root.get_submodule(parent_name)
    if not isinstance(parent, nn.Sequential) or not leaf.isdigit():
        return None, None
    i = int(leaf) + offset
    if i >= len(parent):
        return None, None
    return f"{parent_name}.{i}", parent[i]


def record_step(family: str, h: int, w: int, batch: int, z: int, seed: int = 10_000):
    """Run one oracle step with the tape on.  Returns (units, step result, G state before, D state before, batch)."""
    G, D = build_oracle(family, h, w, z)
    sg, sd = deterministic_state(G, 1234), deterministic_state(D, 4321)
    G.load_state_dict(sg, strict=True)
    D.load_state_dict(sd, strict=True)
    tape = Tape()
    convs = (nn.Conv2d, nn.ConvTranspose2d)
    tape.watch(G, "G.", convs, want_gx=True)
    tape.watch(G, "G.", (nn.BatchNorm2d,))
    tape.watch(G, "G.", (nn.ReLU,), want_x=False)
    tape.watch(G, "G.", (nn.MaxPool2d,))
    tape.watch(G, "G.", (om.SpatialFiLMLayer,))
    tape.watch(G, "G.", (om.CharacterTokenEncoder, om.CharacterTokenEncoderOldV))
    tape.watch(D, "D.", convs)
    tape.watch(D, "D.", (nn.InstanceNorm2d,))
    tape.watch(D, "D.", (nn.LeakyReLU,), want_x=False)
    data = synthetic_batch(batch, h, w, step=0)
    res = train_step(G, D, *make_optimizers(G, D), data, LossWeights.for_family(family), seed=seed)
    tape.close()

    units: List[Unit] = []
    rec = tape.rec

    def first(key):
        return rec[key][0] if key in rec and rec[key] else None

    # ---- generator ----
    for name, m in G.named_modules():
        key = "G." + name
        e = first(key)
        if e is None:
            continue
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)) and "gy" in e:
            units.append(Unit("conv", "G", name, e["x"], e["gy"]))
        elif isinstance(m, nn.BatchNorm2d):
            relu_name, relu = _seq_next(G, name)
            assert isinstance(relu, nn.ReLU), f"{name}: BatchNorm2d not followed by ReLU"
            gy = first("G." + relu_name).get("gy")
            if gy is None:
                continue
            u = Unit("norm", "G", name, e["x"], gy, act=1)
            # encoder blocks: e_convK.4 -> ReLU -> poolK (vae-gan-v2.py:157-163)
            parent = name.rsplit(".", 2)[0] if name.count(".") >= 2 else ""
            blk = name.split(".")[-2]
            if blk.startswith("e_conv") and name.endswith(".4"):
                pool_key = f"G.{parent}.pool{blk[len('e_conv'):]}" if parent else f"G.pool{blk[len('e_conv'):]}"
                pe = first(pool_key)
                if pe is not None and "gy" in pe:
                    # the ReLU output feeds the pool AND (v2 / oldv) the decoder: split the recorded total gradient
                    y = pe["x"].clone().requires_grad_()          # the pool's input = this unit's full-resolution output
                    F.max_pool2d(y, 2).backward(pe["gy"])
                    u.gpool, u.pool = pe["gy"], True
                    u.gy = gy - y.grad
            units.append(u)
        elif isinstance(m, om.SpatialFiLMLayer) and "gy" in e:
            units.append(Unit("film", "G", name, e["extra"][0], e["gy"], x2=e["x"]))
            ce = first(f"G.{name}.param_predictor.0")
            if ce is not None and "gx" in ce:
                units.append(Unit("upsample", "G", name, e["extra"][0], ce["gx"], meta={"size": tuple(e["x"].shape[2:])}))
        elif isinstance(m, (om.CharacterTokenEncoder, om.CharacterTokenEncoderOldV)) and "gy" in e:
            units.append(Unit("text", "G", name, e["x"], e["gy"]))
    # ---- discriminator: first call of the step, D(en) (vae-gan.py:409) ----
    for name, m in D.named_modules():
        e = first("D." + name)
        if e is None:
            continue
        if isinstance(m, nn.Conv2d):
            nxt_name, nxt = _seq_next(D, name)
            if isinstance(nxt, nn.LeakyReLU):       # body.0: SN conv -> LeakyReLU, fused into the conv epilogue
                gy = first("D." + nxt_name).get("gy")
                units.append(Unit("conv", "D", name, e["x"], gy, act=2))
            elif "gy" in e:
                units.append(Unit("conv", "D", name, e["x"], e["gy"]))
        elif isinstance(m, nn.InstanceNorm2d):
            act_name, act = _seq_next(D, name)
            assert isinstance(act, nn.LeakyReLU)
            units.append(Unit("norm", "D", name, e["x"], first("D." + act_name)["gy"], act=2, meta={"per_sample": True}))
    return units, res, sg, sd, data


def ref_unit(u: Unit, G: nn.Module, D: nn.Module, round_bf16: bool):
    """Evaluate the unit with the oracle's own module (a fresh copy holding the pre-step state) on the recorded tensors.
    Returns {"y": ..., "pool": ..., "dx": ..., "dx2": ..., "grads": {param name: grad}}.
    ``round_bf16``: the recorded tensors are first rounded to bf16 (the CUDA path stores activations in bf16; both sides
    then see identical inputs)."""
    r = (lambda t: t.detach().bfloat16().float()) if round_bf16 else (lambda t: t.detach().clone())
    net = G if u.net == "G" else D
    out: dict = {"grads": {}}
    if u.kind == "conv":
        m = copy.deepcopy(net.get_submodule(u.name)).train()
        x = r(u.x).requires_grad_()
        y = m(x)
        if u.act == 2:
            y = F.leaky_relu(y, 0.2)
        y.backward(r(u.gy))
        out.update(y=y.detach(), dx=x.grad)
        out["grads"] = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    elif u.kind == "norm":
        m = copy.deepcopy(net.get_submodule(u.name)).train()
        x = r(u.x).requires_grad_()
        y = m(x)
        y = F.relu(y) if u.act == 1 else F.leaky_relu(y, 0.2)
        if u.pool:
            p = F.max_pool2d(y, 2)
            torch.autograd.backward([y, p], [r(u.gy), r(u.gpool)])
            out["pool"] = p.detach()
        else:
            y.backward(r(u.gy))
        out.update(y=y.detach(), dx=x.grad)
        out["grads"] = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    elif u.kind == "film":
        m = net.get_submodule(u.name)
        with torch.no_grad():
            t = F.interpolate(u.x, size=u.x2.shape[2:], mode="bilinear", align_corners=False)
            gb = r(copy.deepcopy(m.param_predictor).train()(t))
        gb.requires_grad_()
        x = r(u.x2).requires_grad_()
        n = m.num_features_main
        y = gb[:, :n] * x + gb[:, n:]
        y.backward(r(u.gy))
        out.update(y=y.detach(), dx=gb.grad, dx2=x.grad, gb=gb.detach())
    elif u.kind == "upsample":
        x = r(u.x).requires_grad_()
        y = F.interpolate(x, size=u.meta["size"], mode="bilinear", align_corners=False)
        y.backward(r(u.gy))
        out.update(y=y.detach(), dx=x.grad)
    elif u.kind == "text":
        m = copy.deepcopy(net.get_submodule(u.name)).train()
        y = m(u.x)
        y.backward(u.gy)
        out.update(y=y.detach())
        out["grads"] = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    else:
        raise ValueError(u.kind)
    return out


def fresh_oracle(family, h, w, z, sg, sd):
    G, D = build_oracle(family, h, w, z)
    G.load_state_dict(sg, strict=True)
    D.load_state_dict(sd, strict=True)
    return G, D


def act_ambiguity_mask(u: Unit, G: nn.Module, D: nn.Module, rel_band: float = 1e-2) -> torch.Tensor:
    """For a conv unit whose activation is fused into the conv (D's first layer: SN conv -> LeakyReLU): boolean mask of
    the output elements whose pre-activation lies within ``rel_band`` x rms of zero.  The CUDA path rounds the weight
    operand W / sigma to bf16, which moves every pre-activation by ~1e-3 of its rms; elements that close to zero can land on
    either side, so their activation slope (1 vs 0.2) is ambiguous at bf16 resolution.  The test zeroes the incoming
    gradient there (for both sides), which removes them from every gradient of the unit."""
    net = G if u.net == "G" else D
    with torch.no_grad():
        pre = copy.deepcopy(net.get_submodule(u.name)).train()(u.x.detach().clone())
    return pre.abs() < rel_band * pre.pow(2).mean().sqrt()


def pool_tie_mask(y: torch.Tensor) -> torch.Tensor:
    """Boolean NCHW mask of the 2x2 windows whose two largest values become EQUAL when rounded to bf16: there the
    arg-max -- and with it the routing of the pooled gradient -- is ambiguous at bf16 resolution (both routings are valid
    sub-gradients of the bf16 computation), so those windows are left out of the input-gradient comparison."""
    n, c, h, w = y.shape
    win = y.bfloat16().float().reshape(n, c, h // 2, 2, w // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(n, c, h // 2, w // 2, 4)
    top2 = win.topk(2, dim=-1).values
    tie = (top2[..., 0] == top2[..., 1]) & (top2[..., 0] > 0)
    return tie.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
