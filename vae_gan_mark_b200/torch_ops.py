"""``torch.library`` registration of the tensor-core and normalisation primitives (namespace ``vaegan``).

The drop-in modules call the same kernels through ``torch.autograd.Function`` objects (layers.py), which carry
per-layer caches (operand layouts, spectral-norm state) that a stateless op cannot.  This module exposes the
stateless core of the path as dispatcher ops, for callers that want to build their own graphs out of them or trace them
(``register_fake`` gives shape / dtype inference, ``register_autograd`` wires the backward ops):

    torch.ops.vaegan.conv2d(x, weight, bias, stride, pad_h, pad_w, act)         NHWC activations, OIHW fp32 weight
    torch.ops.vaegan.conv2d_dgrad / conv2d_wgrad                                its two gradients
    torch.ops.vaegan.conv_transpose2d(x, weight, bias, stride, pad, out_h, out_w, act)
    torch.ops.vaegan.batch_norm_act(x, gamma, beta, eps, act) -> (y, mean_rstd)  training-mode statistics
    torch.ops.vaegan.batch_norm_act_backward
    torch.ops.vaegan.film(gb, x)
    torch.ops.vaegan.upsample_bilinear2d(t, h, w)                               F.interpolate(bilinear) of an NHWC map
    torch.ops.vaegan.channel_gate(x, scale)                                     x * scale[c] (GatedSkipConnection)
    torch.ops.vaegan.reparam_kl(heads, bias_mu, bias_lv, eps) -> (mu, logvar, z, kl)
    torch.ops.vaegan.l1_loss(a, b) / hinge_loss(p, mode)                         fp32 scalars, mode 1 real / 0 fake / 2 G

Activations are NHWC tensors of the package's activation dtype (bf16, or fp32 in the high-accuracy mode); ``act`` is
0 none / 1 ReLU / 2 LeakyReLU(0.2).  Everything runs on the C ABI of libvaegan_b200.so; there is no CPU implementation
(the fake implementations only describe shapes).  Reference call sites: see INTEGRATION.md.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor
from torch.library import custom_op, register_autograd

from . import layers as L
from . import ops
from .conv import ConvLinear, new_act
from .ops import F32


def _op(weight: Tensor, stride: int, pad_h: int, pad_w: int, in_hw: Optional[Tuple[int, int]] = None,
        transposed: bool = False) -> ConvLinear:
    if transposed:     # ConvTranspose2d weight is IOHW = the OIHW weight of the adjoint conv
        return ConvLinear(weight.shape[1], weight.shape[0], weight.shape[2], weight.shape[3], stride, (pad_h, pad_w), in_hw)
    return ConvLinear(weight.shape[1], weight.shape[0], weight.shape[2], weight.shape[3], stride, (pad_h, pad_w), in_hw)


def _hi(x: Tensor) -> bool:
    return x.dtype == F32


# ---------------------------------------------------------------------------------------------------------------
# Conv2d
# ---------------------------------------------------------------------------------------------------------------
@custom_op("vaegan::conv2d", mutates_args=())
def conv2d(x: Tensor, weight: Tensor, bias: Optional[Tensor], stride: int, pad_h: int, pad_w: int, act: int) -> Tensor:
    op = _op(weight, stride, pad_h, pad_w)
    return op.forward(x, op.prep_fwd(weight, None, _hi(x)), bias, act)


@conv2d.register_fake
def _(x, weight, bias, stride, pad_h, pad_w, act):
    n, h, w, _ = x.shape
    oh = (h + 2 * pad_h - weight.shape[2]) // stride + 1
    ow = (w + 2 * pad_w - weight.shape[3]) // stride + 1
    return x.new_empty((n, oh, ow, weight.shape[0]))


@custom_op("vaegan::conv2d_dgrad", mutates_args=())
def conv2d_dgrad(dy: Tensor, weight: Tensor, stride: int, pad_h: int, pad_w: int, in_h: int, in_w: int) -> Tensor:
    op = _op(weight, stride, pad_h, pad_w)
    return op.backward_data(dy, op.prep_bwd(weight, None, _hi(dy)), (in_h, in_w))


@conv2d_dgrad.register_fake
def _(dy, weight, stride, pad_h, pad_w, in_h, in_w):
    return dy.new_empty((dy.shape[0], in_h, in_w, weight.shape[1]))


@custom_op("vaegan::conv2d_wgrad", mutates_args=())
def conv2d_wgrad(dy: Tensor, x: Tensor, kh: int, kw: int, stride: int, pad_h: int, pad_w: int) -> Tensor:
    op = ConvLinear(x.shape[3], dy.shape[3], kh, kw, stride, (pad_h, pad_w))
    return op.backward_weight(dy, x).contiguous()


@conv2d_wgrad.register_fake
def _(dy, x, kh, kw, stride, pad_h, pad_w):
    return dy.new_empty((dy.shape[3], x.shape[3], kh, kw), dtype=torch.float32)


@custom_op("vaegan::act_backward", mutates_args=())
def act_backward(y: Tensor, dy: Tensor, act: int) -> Tensor:
    dx = new_act(*dy.shape, dy.device, dy.dtype)
    ops.act_bwd(y, dy.contiguous(), dx, act)
    return dx


@act_backward.register_fake
def _(y, dy, act):
    return torch.empty_like(dy)


@custom_op("vaegan::channel_sum", mutates_args=())
def channel_sum(dy: Tensor) -> Tensor:
    return ops.norm_stats(dy.contiguous(), per_sample=False)[0, 0].clone()


@channel_sum.register_fake
def _(dy):
    return dy.new_empty((dy.shape[3],), dtype=torch.float32)


def _conv2d_setup(ctx, inputs, output):
    x, weight, bias, stride, pad_h, pad_w, act = inputs
    ctx.save_for_backward(x, weight, output if act else None)
    ctx.geom, ctx.act, ctx.has_bias = (stride, pad_h, pad_w), act, bias is not None


def _conv2d_backward(ctx, dy):
    x, weight, y = ctx.saved_tensors
    stride, pad_h, pad_w = ctx.geom
    g = torch.ops.vaegan.act_backward(y, dy, ctx.act) if ctx.act else dy
    dx = torch.ops.vaegan.conv2d_dgrad(g, weight, stride, pad_h, pad_w, x.shape[1], x.shape[2]) if ctx.needs_input_grad[0] else None
    dw = (torch.ops.vaegan.conv2d_wgrad(g, x, weight.shape[2], weight.shape[3], stride, pad_h, pad_w)
          if ctx.needs_input_grad[1] else None)
    db = torch.ops.vaegan.channel_sum(g) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
    return dx, dw, db, None, None, None, None


register_autograd("vaegan::conv2d", _conv2d_backward, setup_context=_conv2d_setup)


# ---------------------------------------------------------------------------------------------------------------
# ConvTranspose2d = data gradient of the adjoint conv (forward) / forward conv (input gradient)
# ---------------------------------------------------------------------------------------------------------------
@custom_op("vaegan::conv_transpose2d", mutates_args=())
def conv_transpose2d(x: Tensor, weight: Tensor, bias: Optional[Tensor], stride: int, pad: int, out_h: int, out_w: int,
                     act: int) -> Tensor:
    op = _op(weight, stride, pad, pad, (out_h, out_w), transposed=True)
    return op.backward_data(x, op.prep_bwd(weight, None, _hi(x)), (out_h, out_w), bias, act)


@conv_transpose2d.register_fake
def _(x, weight, bias, stride, pad, out_h, out_w, act):
    return x.new_empty((x.shape[0], out_h, out_w, weight.shape[1]))


@custom_op("vaegan::conv_transpose2d_dgrad", mutates_args=())
def conv_transpose2d_dgrad(dy: Tensor, weight: Tensor, stride: int, pad: int) -> Tensor:
    op = _op(weight, stride, pad, pad, (dy.shape[1], dy.shape[2]), transposed=True)
    return op.forward(dy, op.prep_fwd(weight, None, _hi(dy)))


@conv_transpose2d_dgrad.register_fake
def _(dy, weight, stride, pad):
    kh, kw = weight.shape[2], weight.shape[3]
    return dy.new_empty((dy.shape[0], (dy.shape[1] + 2 * pad - kh) // stride + 1, (dy.shape[2] + 2 * pad - kw) // stride + 1,
                         weight.shape[0]))


@custom_op("vaegan::conv_transpose2d_wgrad", mutates_args=())
def conv_transpose2d_wgrad(dy: Tensor, x: Tensor, kh: int, kw: int, stride: int, pad: int) -> Tensor:
    op = ConvLinear(dy.shape[3], x.shape[3], kh, kw, stride, (pad, pad), (dy.shape[1], dy.shape[2]))
    return op.backward_weight(x, dy).contiguous()        # operands swapped: [C_in][C_out][kh][kw]


@conv_transpose2d_wgrad.register_fake
def _(dy, x, kh, kw, stride, pad):
    return dy.new_empty((x.shape[3], dy.shape[3], kh, kw), dtype=torch.float32)


def _convT_setup(ctx, inputs, output):
    x, weight, bias, stride, pad, out_h, out_w, act = inputs
    ctx.save_for_backward(x, weight, output if act else None)
    ctx.geom, ctx.act, ctx.has_bias = (stride, pad), act, bias is not None


def _convT_backward(ctx, dy):
    x, weight, y = ctx.saved_tensors
    stride, pad = ctx.geom
    g = torch.ops.vaegan.act_backward(y, dy, ctx.act) if ctx.act else dy
    dx = torch.ops.vaegan.conv_transpose2d_dgrad(g, weight, stride, pad) if ctx.needs_input_grad[0] else None
    dw = (torch.ops.vaegan.conv_transpose2d_wgrad(g, x, weight.shape[2], weight.shape[3], stride, pad)
          if ctx.needs_input_grad[1] else None)
    db = torch.ops.vaegan.channel_sum(g) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
    return dx, dw, db, None, None, None, None, None


register_autograd("vaegan::conv_transpose2d", _convT_backward, setup_context=_convT_setup)


# ---------------------------------------------------------------------------------------------------------------
# BatchNorm2d (training statistics) + activation
# ---------------------------------------------------------------------------------------------------------------
@custom_op("vaegan::batch_norm_act", mutates_args=())
def batch_norm_act(x: Tensor, gamma: Tensor, beta: Tensor, eps: float, act: int) -> Tuple[Tensor, Tensor]:
    n, h, w, c = x.shape
    mr = ops.norm_finalize(ops.norm_stats(x, False), n * h * w, eps)
    y = new_act(n, h, w, c, x.device, x.dtype)
    ops.norm_apply(x, mr, gamma, beta, act, y)
    return y, mr


@batch_norm_act.register_fake
def _(x, gamma, beta, eps, act):
    return torch.empty_like(x), x.new_empty((1, 2, x.shape[3]), dtype=torch.float32)


@custom_op("vaegan::batch_norm_act_backward", mutates_args=())
def batch_norm_act_backward(x: Tensor, dy: Tensor, mean_rstd: Tensor, gamma: Tensor, beta: Tensor,
                            act: int) -> Tuple[Tensor, Tensor, Tensor]:
    n, h, w, c = x.shape
    dx = new_act(n, h, w, c, x.device, x.dtype)
    dgamma = torch.empty(c, dtype=F32, device=x.device)
    dbeta = torch.empty(c, dtype=F32, device=x.device)
    ops.norm_backward(x, dy.contiguous(), None, mean_rstd, False, gamma, beta, act, dx, dgamma, dbeta)
    return dx, dgamma, dbeta


@batch_norm_act_backward.register_fake
def _(x, dy, mean_rstd, gamma, beta, act):
    c = x.shape[3]
    return torch.empty_like(x), x.new_empty((c,), dtype=torch.float32), x.new_empty((c,), dtype=torch.float32)


def _bn_setup(ctx, inputs, output):
    x, gamma, beta, eps, act = inputs
    ctx.save_for_backward(x, gamma, beta, output[1])
    ctx.act = act


def _bn_backward(ctx, dy, dmr):
    x, gamma, beta, mr = ctx.saved_tensors
    dx, dgamma, dbeta = torch.ops.vaegan.batch_norm_act_backward(x, dy, mr, gamma, beta, ctx.act)
    return dx, dgamma, dbeta, None, None


register_autograd("vaegan::batch_norm_act", _bn_backward, setup_context=_bn_setup)


# ---------------------------------------------------------------------------------------------------------------
# FiLM
# ---------------------------------------------------------------------------------------------------------------
@custom_op("vaegan::film", mutates_args=())
def film(gb: Tensor, x: Tensor) -> Tensor:
    y = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    ops.film_fwd(gb.contiguous(), x, y)
    return y


@film.register_fake
def _(gb, x):
    return x.new_empty(x.shape)


@custom_op("vaegan::film_backward", mutates_args=())
def film_backward(gb: Tensor, x: Tensor, dy: Tensor) -> Tuple[Tensor, Tensor]:
    gb = gb.contiguous()
    dgb = torch.empty_like(gb)
    dx = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    ops.film_bwd(gb, x, dy.contiguous(), dgb, dx)
    return dgb, dx


@film_backward.register_fake
def _(gb, x, dy):
    return torch.empty_like(gb), x.new_empty(x.shape)


def _film_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _film_backward(ctx, dy):
    gb, x = ctx.saved_tensors
    return torch.ops.vaegan.film_backward(gb, x, dy)


register_autograd("vaegan::film", _film_backward, setup_context=_film_setup)


# ---------------------------------------------------------------------------------------------------------------
# Bilinear resize of an NHWC map and the per-channel skip gate (vae-gan-oldv.py:165-176, 226-231)
# ---------------------------------------------------------------------------------------------------------------
@custom_op("vaegan::upsample_bilinear2d", mutates_args=())
def upsample_bilinear2d(t: Tensor, h: int, w: int) -> Tensor:
    return L.upsample2d_fwd(t, h, w)


@upsample_bilinear2d.register_fake
def _(t, h, w):
    return t.new_empty((t.shape[0], h, w, t.shape[3]))


@custom_op("vaegan::upsample_bilinear2d_backward", mutates_args=())
def upsample_bilinear2d_backward(dy: Tensor, h0: int, w0: int) -> Tensor:
    return L.upsample2d_bwd(dy, h0, w0)


@upsample_bilinear2d_backward.register_fake
def _(dy, h0, w0):
    return dy.new_empty((dy.shape[0], h0, w0, dy.shape[3]))


def _up_setup(ctx, inputs, output):
    ctx.hw0 = (inputs[0].shape[1], inputs[0].shape[2])


def _up_backward(ctx, dy):
    return torch.ops.vaegan.upsample_bilinear2d_backward(dy, ctx.hw0[0], ctx.hw0[1]), None, None


register_autograd("vaegan::upsample_bilinear2d", _up_backward, setup_context=_up_setup)


@custom_op("vaegan::channel_gate", mutates_args=())
def channel_gate(x: Tensor, scale: Tensor) -> Tensor:
    y = new_act(*x.shape, x.device, x.dtype)
    ops.channel_scale_fwd(x, scale.contiguous(), y)
    return y


@channel_gate.register_fake
def _(x, scale):
    return torch.empty_like(x)


@custom_op("vaegan::channel_gate_backward", mutates_args=())
def channel_gate_backward(x: Tensor, scale: Tensor, dy: Tensor) -> Tuple[Tensor, Tensor]:
    c = x.shape[3]
    dx = new_act(*x.shape, x.device, x.dtype)
    ds = torch.empty(2 * c, dtype=F32, device=x.device)
    ops.channel_scale_bwd(x, L.grad_in(dy, x.dtype), scale.contiguous(), dx, ds)
    return dx, ds[:c].clone()


@channel_gate_backward.register_fake
def _(x, scale, dy):
    return torch.empty_like(x), x.new_empty((x.shape[3],), dtype=torch.float32)


def _gate_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _gate_backward(ctx, dy):
    x, scale = ctx.saved_tensors
    return torch.ops.vaegan.channel_gate_backward(x, scale, dy)


register_autograd("vaegan::channel_gate", _gate_backward, setup_context=_gate_setup)


# ---------------------------------------------------------------------------------------------------------------
# Reparameterisation + KL, L1 and hinge losses (vae-gan.py:133-136, 313-320, 419-420); all fp32
# ---------------------------------------------------------------------------------------------------------------
@custom_op("vaegan::reparam_kl", mutates_args=())
def reparam_kl(heads: Tensor, bias_mu: Tensor, bias_lv: Tensor, eps: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """heads: fp32 [B, 2z] = (mu | logvar) pre-bias; eps: fp32 [B, z].  Returns mu, logvar, z = mu + eps*exp(logvar/2)
    (all [B, z]) and kl = mean_b(-0.5 mean_c(1 + logvar - mu^2 - exp(logvar)))."""
    b = heads.shape[0]
    return ops.reparam_kl_fwd(heads.reshape(b, -1).contiguous(), bias_mu.contiguous(), bias_lv.contiguous(),
                              eps.reshape(b, -1).contiguous())


@reparam_kl.register_fake
def _(heads, bias_mu, bias_lv, eps):
    b, z = heads.shape[0], bias_mu.shape[0]
    f = dict(dtype=torch.float32)
    return heads.new_empty((b, z), **f), heads.new_empty((b, z), **f), heads.new_empty((b, z), **f), heads.new_empty((), **f)


@custom_op("vaegan::reparam_kl_backward", mutates_args=())
def reparam_kl_backward(mu: Tensor, lv: Tensor, eps: Tensor, dmu: Tensor, dlv: Tensor, dz: Tensor, dkl: Tensor) -> Tensor:
    return ops.reparam_kl_bwd(mu, lv, eps.reshape(mu.shape).contiguous(), dz.contiguous(), dmu.contiguous(),
                              dlv.contiguous(), dkl.contiguous())


@reparam_kl_backward.register_fake
def _(mu, lv, eps, dmu, dlv, dz, dkl):
    return mu.new_empty((mu.shape[0], 2 * mu.shape[1]))


def _rk_setup(ctx, inputs, output):
    ctx.save_for_backward(output[0], output[1], inputs[3])
    ctx.heads_shape = inputs[0].shape


def _rk_backward(ctx, dmu, dlv, dz, dkl):
    mu, lv, eps = ctx.saved_tensors
    dheads = torch.ops.vaegan.reparam_kl_backward(mu, lv, eps, dmu, dlv, dz, dkl)
    z = mu.shape[1]
    return dheads.view(ctx.heads_shape), dheads[:, :z].sum(0), dheads[:, z:].sum(0), None


register_autograd("vaegan::reparam_kl", _rk_backward, setup_context=_rk_setup)


@custom_op("vaegan::l1_loss", mutates_args=())
def l1_loss(a: Tensor, b: Tensor) -> Tensor:
    return ops.l1_fwd(a.contiguous(), b.contiguous())


@l1_loss.register_fake
def _(a, b):
    return a.new_empty((), dtype=torch.float32)


@custom_op("vaegan::l1_loss_backward", mutates_args=())
def l1_loss_backward(a: Tensor, b: Tensor, gout: Tensor) -> Tensor:
    a = a.contiguous()
    da = torch.empty_like(a)
    ops.l1_bwd(a, b.contiguous(), gout.contiguous(), da)
    return da


@l1_loss_backward.register_fake
def _(a, b, gout):
    return torch.empty_like(a)


def _l1_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _l1_backward(ctx, g):
    a, b = ctx.saved_tensors
    return torch.ops.vaegan.l1_loss_backward(a, b, g), None


register_autograd("vaegan::l1_loss", _l1_backward, setup_context=_l1_setup)


@custom_op("vaegan::hinge_loss", mutates_args=())
def hinge_loss(p: Tensor, mode: int) -> Tensor:
    """mode 1: mean(relu(1 - p)) (real), 0: mean(relu(1 + p)) (fake), 2: -mean(p) (generator); vae-gan.py:313-320."""
    return ops.hinge_fwd(p.contiguous(), mode)


@hinge_loss.register_fake
def _(p, mode):
    return p.new_empty((), dtype=torch.float32)


@custom_op("vaegan::hinge_loss_backward", mutates_args=())
def hinge_loss_backward(p: Tensor, mode: int, gout: Tensor) -> Tensor:
    p = p.contiguous()
    dp = torch.empty_like(p)
    ops.hinge_bwd(p, mode, gout.contiguous(), dp)
    return dp


@hinge_loss_backward.register_fake
def _(p, mode, gout):
    return torch.empty_like(p)


def _hinge_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0])
    ctx.mode = inputs[1]


def _hinge_backward(ctx, g):
    (p,) = ctx.saved_tensors
    return torch.ops.vaegan.hinge_loss_backward(p, ctx.mode, g), None


register_autograd("vaegan::hinge_loss", _hinge_backward, setup_context=_hinge_setup)
