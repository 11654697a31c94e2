"""Autograd-level building blocks: each ``torch.autograd.Function`` here runs hand-written CUDA kernels
(ops.py / conv.py) in both directions.  Activations between them are NHWC bf16 views; parameters stay fp32 in
PyTorch layouts so that state_dicts match the reference modules.
"""
from __future__ import annotations

import weakref
from typing import Dict, Optional, Tuple

import torch
from torch.autograd import Function

from . import ops
from .conv import PAIRS, ConvLinear, _hi_launch, _hi_wgrad_split, _operand, _vtaps, fprop, new_act, pad_channels, wgrad
from .ops import BF16, F32, round_up

_WEIGHT_EPOCH = 0


def bump_weight_epoch() -> None:
    """Invalidate cached bf16 weight re-layouts (call after updating parameters through raw pointers)."""
    global _WEIGHT_EPOCH
    _WEIGHT_EPOCH += 1


# bf16 operand shadows maintained by the optimiser (train.FusedAdam writes them in the same kernel that updates the fp32
# master weights): {data_ptr of the parameter (or tuple of data_ptrs): [bf16 operand, parameter(s), version at last sync]}.
# The forward operand of a conv whose weight lives in [Cout][kh][kw][Cin] order with Cin % 64 == 0 is exactly the bf16
# copy of that memory, so ``WeightCache.get("fwd", ...)`` returns the shadow instead of converting the weight every step.
SHADOWS: Dict[object, list] = {}


def shadow_eligible(p: torch.Tensor) -> bool:
    """A 4-D weight whose memory order is [O][kh][kw][I] (channels_last, or any order when the kernel is 1x1) with
    I % 64 == 0: its forward operand is the plain bf16 copy of its memory."""
    if p.dim() != 4 or p.shape[1] % 64 != 0 or p.dtype != F32:
        return False
    if p.shape[2] * p.shape[3] == 1:
        return p.is_contiguous() or p.is_contiguous(memory_format=torch.channels_last)
    return p.is_contiguous(memory_format=torch.channels_last)


def register_shadow(params, operand: torch.Tensor) -> None:
    """``params``: one parameter, or the (mu_head, logvar_head) pair whose operands are the two halves of ``operand``."""
    ps = params if isinstance(params, (tuple, list)) else (params,)
    for k in [k for k, e in SHADOWS.items() if any(r() is None for r in e[1])]:      # parameters that no longer exist
        del SHADOWS[k]
    key = ps[0].data_ptr() if len(ps) == 1 else tuple(q.data_ptr() for q in ps)
    SHADOWS[key] = [operand, tuple(weakref.ref(q) for q in ps), None]
    sync_shadow(key)


def sync_all_shadows() -> None:
    """After parameters were written behind autograd's back (``p.data`` broadcasts of the data-parallel setup)."""
    for key in list(SHADOWS):
        if all(r() is not None for r in SHADOWS[key][1]):
            sync_shadow(key)


def sync_shadow(key) -> None:
    """(Re)fill a shadow from its fp32 parameter(s) -- at registration and whenever a parameter was modified by anything
    other than the optimiser kernel (load_state_dict, an eager copy_: both bump the tensor's version counter)."""
    operand, refs, _ = SHADOWS[key]
    ps = tuple(r() for r in refs)
    rows = 0
    for q in ps:
        o, i, kh, kw = q.shape
        ops.strided_copy(q.detach().permute(0, 2, 3, 1), operand[rows:rows + o].view(o, kh, kw, i))
        rows += o
    SHADOWS[key][2] = tuple(q._version for q in ps)


def _shadow_for(params) -> Optional[torch.Tensor]:
    if not SHADOWS or ops.act_dtype() != BF16:
        return None
    key = params[0].data_ptr() if len(params) == 1 else tuple(q.data_ptr() for q in params)
    ent = SHADOWS.get(key)
    if ent is None or any(r() is not b for r, b in zip(ent[1], params)):
        return None
    if ent[2] != tuple(q._version for q in params):
        sync_shadow(key)
    return ent[0]


class WeightCache:
    """bf16 GEMM layouts of one fp32 parameter, rebuilt only when the parameter changed."""

    def __init__(self):
        self.key = None
        self.store: Dict[str, object] = {}

    def get(self, name: str, param, build):
        params = param if isinstance(param, (tuple, list)) else (param,)
        if name == "fwd":
            sh = _shadow_for(params)
            if sh is not None:
                return sh
        key = tuple((p.data_ptr(), p._version) for p in params) + (_WEIGHT_EPOCH,)
        if key != self.key:
            self.key, self.store = key, {}
        if name not in self.store:
            self.store[name] = build()
        return self.store[name]

    def clear(self):
        self.key, self.store = None, {}


def grad_in(t: Optional[torch.Tensor], dtype=None) -> Optional[torch.Tensor]:
    """Coerce an incoming autograd gradient to a valid NHWC view of the activation dtype (copying only if it is not
    one already); when the channel count is not a multiple of 64 the copy is padded so the tensor pipe can read it."""
    if t is None:
        return None
    dtype = dtype or ops.act_dtype()
    c = t.shape[3]
    need_pad = c % 64 != 0 and t.stride(2) < round_up(c, 64)
    if t.dtype == dtype and ops.nhwc_ok(t) and not need_pad:
        return t
    out = new_act(t.shape[0], t.shape[1], t.shape[2], c, t.device, dtype)
    ops.strided_copy(t, out)
    return out


def _write_param_grad(view_oihw: torch.Tensor, like: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 gradient, given as an OIHW-shaped (permuted) view of the kernel's [O][kh][kw][I] result, in the memory
    order of the parameter ``like``.  For channels_last parameters the view already IS that order: no copy."""
    if like is not None and view_oihw.stride() == like.stride():
        return view_oihw
    out = torch.empty_like(like, dtype=F32) if like is not None else torch.empty(view_oihw.shape, dtype=F32,
                                                                                 device=view_oihw.device)
    ops.strided_copy(view_oihw, out)
    return out


def _bias_grad(dy: torch.Tensor) -> torch.Tensor:
    """Per-channel sum of an NHWC bf16 gradient (fp32)."""
    return ops.norm_stats(dy, per_sample=False)[0, 0].clone()


# ------------------------------------------------------------------------------------------------
# convolutions on the tensor pipe
# ------------------------------------------------------------------------------------------------
class Conv2dFn(Function):
    """y = act(conv2d(x, w) + b).  Replaces nn.Conv2d forward/backward (cuDNN) of the reference layers."""

    @staticmethod
    def forward(ctx, x, weight, bias, op: ConvLinear, cache: WeightCache, act: int, out, out_kind: int, sn, stats=None):
        """``stats``: optional fp32 [1,2,cout] filled by the epilogue with the batch statistics of y (bf16 mode)."""
        scale = sn.sigma if sn is not None else None
        hi = x.dtype == F32
        wf = (cache.get("fwd_hi" if hi else "fwd", weight, lambda: op.prep_fwd(weight.detach(), scale, hi)) if sn is None
              else op.prep_fwd(weight.detach(), scale, hi))
        y = op.forward(x, wf, bias.detach() if bias is not None else None, act, out, out_kind, stats=stats)
        ctx.op, ctx.cache, ctx.act, ctx.sn, ctx.has_bias = op, cache, act, sn, bias is not None
        ctx.wf = wf if (sn is not None and not hi) else None     # W / sigma of THIS call, reused by its data gradient
        ctx.in_hw = (x.shape[1], x.shape[2])
        ctx.save_for_backward(x, weight, y if act else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        op: ConvLinear = ctx.op
        hi = x.dtype == F32
        dy = grad_in(dy, x.dtype)
        if ctx.act:
            g = new_act(*dy.shape, dy.device, dy.dtype)
            ops.act_bwd(y, dy, g, ctx.act)
            dy = g
        dx = dw = db = None
        sn = ctx.sn
        scale = sn.sigma if sn is not None else None
        if ctx.needs_input_grad[0]:
            if not hi and op.prefer_mn(x.shape[0] * x.shape[1] * x.shape[2]):
                # small GEMM, big weight: the forward operand doubles as the (MN-major) operand of the data gradient
                wb = {"mn": ctx.wf if sn is not None else
                      ctx.cache.get("fwd", weight, lambda: op.prep_fwd(weight.detach(), None, False))}
            else:
                wb = (ctx.cache.get("bwd_hi" if hi else "bwd", weight, lambda: op.prep_bwd(weight.detach(), scale, hi))
                      if sn is None else op.prep_bwd(weight.detach(), scale, hi))
            dx = op.backward_data(dy, wb, ctx.in_hw)
        if ctx.needs_input_grad[1]:
            gview = op.backward_weight(dy, x)
            dw = _write_param_grad(gview, weight)
            if sn is not None:
                dw = sn.backward(dw, weight.detach())
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _bias_grad(dy)
        return dx, dw, db, None, None, None, None, None, None, None


class ConvTranspose2dFn(Function):
    """y = act(conv_transpose2d(x, w) + b), computed as the data-gradient primitive of the adjoint conv ``op``
    (cin(op) = C_out of the transpose, cout(op) = C_in of the transpose; the IOHW weight is op's OIHW)."""

    @staticmethod
    def forward(ctx, x, weight, bias, op: ConvLinear, cache: WeightCache, act: int, out, out_hw, stats=None):
        hi = x.dtype == F32
        if not hi and op.prefer_mn(x.shape[0] * x.shape[1] * x.shape[2]):
            wb = {"mn": cache.get("fwd", weight, lambda: op.prep_fwd(weight.detach(), None, False))}
        else:
            wb = cache.get("bwd_hi" if hi else "bwd", weight, lambda: op.prep_bwd(weight.detach(), None, hi))
        y = op.backward_data(x, wb, out_hw, bias.detach() if bias is not None else None, act, out, stats=stats)
        ctx.op, ctx.cache, ctx.act, ctx.has_bias = op, cache, act, bias is not None
        ctx.save_for_backward(x, weight, y if act else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        op: ConvLinear = ctx.op
        hi = x.dtype == F32
        dy = grad_in(dy, x.dtype)
        if ctx.act:
            g = new_act(*dy.shape, dy.device, dy.dtype)
            ops.act_bwd(y, dy, g, ctx.act)
            dy = g
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            wf = ctx.cache.get("fwd_hi" if hi else "fwd", weight, lambda: op.prep_fwd(weight.detach(), None, hi))
            dx = op.forward(dy, wf)
        if ctx.needs_input_grad[1]:
            dw = _write_param_grad(op.backward_weight(x, dy), weight)     # operands swapped: [cout(op)=C_in][cin(op)=C_out]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _bias_grad(dy)
        return dx, dw, db, None, None, None, None, None, None


class HeadsFn(Function):
    """mu_head and logvar_head (full-spatial-kernel Conv2d, vae-gan.py:59-60, vae-gan-v2.py:168-169) as ONE split-K
    GEMM with N = 2z over the flattened (h, w, c) feature map; fp32 result [B,1,1,2z] WITHOUT bias (the bias is
    added in the fused reparameterisation kernel)."""

    @staticmethod
    def forward(ctx, feat, w_mu, w_lv, op: ConvLinear, cache_f: WeightCache, cache_b: WeightCache):
        z, c, kh, kw = w_mu.shape
        hi = feat.dtype == F32

        def build():
            tmp = torch.empty((2 * z, kh, kw, c), dtype=F32, device=feat.device)
            ops.strided_copy(w_mu.detach().permute(0, 2, 3, 1), tmp[:z])
            ops.strided_copy(w_lv.detach().permute(0, 2, 3, 1), tmp[z:])
            if op.flat:
                return _operand(tmp.view(2 * z, 1, kh * kw * c), kh * kw * c, hi)
            return _operand(tmp.view(2 * z, kh * kw, c), op.cin_p, hi)
        wf = cache_f.get("fwd_hi" if hi else "fwd", (w_mu, w_lv), build)
        if not feat.is_contiguous():
            feat = ops.dense_nhwc(feat)
        y = op.forward(feat, wf, None, 0, None, 2)
        ctx.op, ctx.cache_b = op, cache_b
        ctx.wf = wf if (op.shuffle and not hi and 2 * z == op.cout_p) else None
        ctx.save_for_backward(feat, w_mu, w_lv)
        return y

    @staticmethod
    def backward(ctx, dy):
        feat, w_mu, w_lv = ctx.saved_tensors
        op: ConvLinear = ctx.op
        z, c, kh, kw = w_mu.shape
        hi = feat.dtype == F32
        dy = grad_in(dy, feat.dtype)
        dx = dmu = dlv = None
        if ctx.needs_input_grad[0]:
            def build():
                tmp = torch.empty((kh, kw, c, 2 * z), dtype=F32, device=dy.device)
                ops.strided_copy(w_mu.detach().permute(2, 3, 1, 0), tmp[..., :z])
                ops.strided_copy(w_lv.detach().permute(2, 3, 1, 0), tmp[..., z:])
                return {"shuffle": _operand(tmp.view(kh * kw * c, 1, 2 * z), op.cout_p, hi)}
            if not hi and ctx.wf is not None and op.prefer_mn(feat.shape[0]):
                # a handful of pixels against a (2z x h*w*c) weight: the forward operand doubles as the MN-major operand of
                # the data gradient -- no transposed copy of the two head weights per step (2 x 134 MB at 256x256)
                wb = {"mn": ctx.wf}
            else:
                wb = ctx.cache_b.get("bwd_hi" if hi else "bwd", (w_mu, w_lv), build)
            dx = op.backward_data(dy, wb, (feat.shape[1], feat.shape[2]))
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            g = op.backward_weight(dy, feat)                 # view [2z, c, kh, kw]
            dmu, dlv = _write_param_grad(g[:z], w_mu), _write_param_grad(g[z:], w_lv)
        return dx, dmu, dlv, None, None, None


class CopyIntoFn(Function):
    """Copy an NHWC tensor into a channel slice of a concat buffer (only needed when the producer could not
    write there directly)."""

    @staticmethod
    def forward(ctx, src, dst):
        ops.strided_copy(src, dst)
        return dst.detach()

    @staticmethod
    def backward(ctx, g):
        return g, None


# ------------------------------------------------------------------------------------------------
# few-channel image-side conv (im2col + GEMM) and few-output-channel conv (CUDA cores)
# ------------------------------------------------------------------------------------------------
class ImageConvFn(Function):
    """Conv2d whose input is a 3/4-channel NCHW fp32 image (or a channel-concatenation of several): the
    encoder's first conv on cat(image, mask) (vae-gan-v2.py:318-319) and D's first conv (vae-gan.py:153).
    im2col into a [pixels][64] matrix, then a plain tensor-core GEMM with bias/activation fused."""

    @staticmethod
    def forward(ctx, weight, bias, geom, cache: WeightCache, act: int, sn, stats, *images):
        kh, kw, stride, pad = geom
        n, _, h, w = images[0].shape
        cin = sum(t.shape[1] for t in images)
        cout = weight.shape[0]
        k = kh * kw * cin
        kpad = round_up(k, 64)
        dt = ops.act_dtype()
        hi = dt == F32
        src = torch.empty((n, h, w, 8), dtype=dt, device=weight.device)
        c0 = 0
        for t in images:
            ops.strided_copy(t.detach().permute(0, 2, 3, 1), src[..., c0:c0 + t.shape[1]])
            c0 += t.shape[1]
        oh, ow = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
        col = torch.empty((n, oh, ow, kpad), dtype=dt, device=weight.device)
        ops.im2col(src, cin, kh, kw, stride, pad, col)
        scale = sn.sigma if sn is not None else None

        def build():
            tmp = torch.empty((cout, kh, kw, cin), dtype=F32, device=weight.device)
            ops.strided_copy(weight.detach().permute(0, 2, 3, 1), tmp, scale, scale_inverse=scale is not None)
            return _operand(tmp.view(cout, 1, k), kpad, hi)
        wf = cache.get("fwd_hi" if hi else "fwd", weight, build) if sn is None else build()
        y = new_act(n, oh, ow, cout, weight.device, dt)
        vt, wk = _vtaps([(0, 0, 0, 0)], kpad, kpad, hi)
        b_ = bias.detach() if bias is not None else None
        if hi:
            _hi_launch(ops.split3(col), vt, 1, kpad, wf, cout, (n, oh, ow), y, wk, b_, act)
        else:
            fprop(col, vt, 1, kpad, wf, cout, (n, oh, ow), y, bias=b_, act=act, stats=stats)
        ctx.geom, ctx.act, ctx.sn, ctx.cin, ctx.kpad, ctx.has_bias = geom, act, sn, cin, kpad, bias is not None
        ctx.img_shapes = [tuple(t.shape) for t in images]
        ctx.save_for_backward(col, weight, y if act else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        col, weight, y = ctx.saved_tensors
        kh, kw, stride, pad = ctx.geom
        cout, cin, kpad = weight.shape[0], ctx.cin, ctx.kpad
        k = kh * kw * cin
        hi = col.dtype == F32
        dy = grad_in(dy, col.dtype)
        if ctx.act:
            g = new_act(*dy.shape, dy.device, dy.dtype)
            ops.act_bwd(y, dy, g, ctx.act)
            dy = g
        n, oh, ow, _ = dy.shape
        cout_p = round_up(cout, 64)
        dw = db = None
        sn = ctx.sn
        if ctx.needs_input_grad[0]:
            dwm = torch.empty((cout, kpad), dtype=F32, device=dy.device)
            if hi:
                wgrad(ops.split3(dy), cout, ops.split3(col), [(0, 0, 0, 0)], 1, kpad, (n, oh, ow), dwm,
                      pairs=[(pa * cout_p, pb * kpad) for (pa, pb) in PAIRS], ksplit=_hi_wgrad_split(n, oh, ow))
            else:
                wgrad(dy, cout, col, [(0, 0, 0, 0)], 1, kpad, (n, oh, ow), dwm)
            dw = _write_param_grad(dwm[:, :k].view(cout, kh, kw, cin).permute(0, 3, 1, 2), weight)
            if sn is not None:
                dw = sn.backward(dw, weight.detach())
        if ctx.has_bias and ctx.needs_input_grad[1]:
            db = _bias_grad(dy)
        dimgs = [None] * len(ctx.img_shapes)
        if any(ctx.needs_input_grad[7:]):
            scale = sn.sigma if sn is not None else None
            tmp = torch.zeros((kpad, cout), dtype=F32, device=dy.device)                 # rows (r, q, ci), zero padded
            ops.strided_copy(weight.detach().permute(2, 3, 1, 0), tmp[:k].view(kh, kw, cin, cout), scale,
                             scale_inverse=scale is not None)
            wd = _operand(tmp.view(kpad, 1, cout), cout_p, hi)
            dcol = torch.empty((n, oh, ow, kpad), dtype=dy.dtype, device=dy.device)
            vt, wk = _vtaps([(0, 0, 0, 0)], cout_p, cout_p, hi)
            if hi:
                _hi_launch(ops.split3(dy), vt, 1, cout_p, wd, kpad, (n, oh, ow), dcol, wk, None, 0)
            else:
                fprop(pad_channels(dy, cout_p), vt, 1, cout_p, wd, kpad, (n, oh, ow), dcol)
            n_, _, h, w = ctx.img_shapes[0]
            dsrc = torch.empty((n_, cin, h, w), dtype=F32, device=dy.device)
            ops.col2im(dcol, n_, h, w, cin, kh, kw, stride, pad, dsrc)
            c0 = 0
            for i, shp in enumerate(ctx.img_shapes):
                if ctx.needs_input_grad[7 + i]:
                    dimgs[i] = dsrc[:, c0:c0 + shp[1]]
                c0 += shp[1]
        return (dw, db, None, None, None, None, None, *dimgs)


class SmallOutConvFn(Function):
    """Stride-1 Conv2d with <= 4 output channels on CUDA cores (HBM-bound): final_image_conv
    (vae-gan-v2.py:232), decode.15 (vae-gan.py:81), D's patch head (vae-gan.py:157).  Output fp32 NHWC."""

    @staticmethod
    def forward(ctx, x, weight, bias, pad: int):
        cout, cin, kh, kw = weight.shape
        n, h, w, _ = x.shape
        wt = torch.empty((cout, kh, kw, cin), dtype=F32, device=x.device)
        ops.strided_copy(weight.detach().permute(0, 2, 3, 1), wt)
        oh, ow = h + 2 * pad - kh + 1, w + 2 * pad - kw + 1
        out = torch.empty((n, oh, ow, cout), dtype=F32, device=x.device)
        ops.smalln_fwd(x, wt, bias.detach() if bias is not None else None, kh, kw, pad, out)
        ctx.pad, ctx.has_bias, ctx.w_param = pad, bias is not None, weight
        ctx.save_for_backward(x, wt)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, wt = ctx.saved_tensors
        cout, kh, kw, cin = wt.shape
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = new_act(x.shape[0], x.shape[1], x.shape[2], cin, x.device, x.dtype)
            ops.smalln_dgrad(dy, wt, kh, kw, ctx.pad, dx)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dwt = torch.empty_like(wt)
            db_ = torch.empty(cout, dtype=F32, device=x.device) if ctx.has_bias else None
            ops.smalln_wgrad(dy, x, kh, kw, ctx.pad, dwt, db_)
            dw = _write_param_grad(dwt.permute(0, 3, 1, 2), ctx.w_param)
            db = db_
        return dx, dw, db, None


# ------------------------------------------------------------------------------------------------
# normalisation + activation (+ pool)
# ------------------------------------------------------------------------------------------------
class NormActFn(Function):
    """BatchNorm2d (training or eval) + ReLU [+ MaxPool2d(2)]  or  InstanceNorm2d(affine) + LeakyReLU(0.2).
    Returns (y, pooled-or-None).  ``out`` lets the full-resolution result land in a slice of a concat buffer."""

    @staticmethod
    def forward(ctx, x, gamma, beta, per_sample: bool, act: int, pool: bool, out, eps: float, bn_state, pool_out=None,
                virt_h: int = 0, sums=None):
        """``virt_h`` > 0: the h (= 3) rows of x stand for virt_h rows whose interior rows are all equal (row-class
        FiLM maps); the statistics weight them accordingly and the backward expects per-class gradient sums.
        ``sums``: fp32 [1,2,c] batch statistics of x already produced by the epilogue of the convolution that wrote x
        (``stats`` of Conv2dFn / ConvTranspose2dFn): the statistics pass over x is skipped."""
        n, h, w, c = x.shape
        ctx.virt_h = virt_h
        g, b = (gamma.detach() if gamma is not None else None), (beta.detach() if beta is not None else None)
        if bn_state is not None and not bn_state["training"]:
            mr = torch.stack([bn_state["running_mean"], torch.rsqrt(bn_state["running_var"] + eps)]).unsqueeze(0).contiguous()
            ctx.eval_mode = True
        else:
            if sums is None or virt_h or per_sample:
                sums = ops.norm_stats_rows(x, virt_h) if virt_h else ops.norm_stats(x, per_sample)
            rows = h * w if per_sample else n * (virt_h or h) * w
            rm = rv = nbt = None
            if bn_state is not None:
                rm, rv, nbt = bn_state["running_mean"], bn_state["running_var"], bn_state["num_batches_tracked"]
            mr = ops.norm_finalize(sums, rows, eps, 0.1, rm, rv, nbt)
            ctx.eval_mode = False
        y = out if out is not None else new_act(n, h, w, c, x.device, x.dtype)
        pooled = None
        if pool:
            pooled = pool_out if pool_out is not None else new_act(n, h // 2, w // 2, c, x.device, x.dtype)
        ops.norm_apply(x, mr, g, b, act, y, pooled)
        ctx.per_sample, ctx.act, ctx.pool = per_sample, act, pool
        ctx.save_for_backward(x, gamma, beta, mr)
        y_ret = y.detach() if out is not None else y
        p_ret = pooled.detach() if (pool and pool_out is not None) else pooled
        return y_ret, p_ret

    @staticmethod
    def backward(ctx, dy, dpool):
        x, gamma, beta, mr = ctx.saved_tensors
        assert not ctx.eval_mode, "backward through eval-mode normalisation is not supported"
        dy, dpool = grad_in(dy, x.dtype), grad_in(dpool, x.dtype)
        n, h, w, c = x.shape
        dx = new_act(n, h, w, c, x.device, x.dtype)
        dgamma = torch.empty(c, dtype=F32, device=x.device) if gamma is not None else None
        dbeta = torch.empty(c, dtype=F32, device=x.device) if beta is not None else None
        if ctx.virt_h and not dy.is_contiguous():
            dy = ops.dense_nhwc(dy)
        ops.norm_backward(x, dy, dpool, mr, ctx.per_sample, gamma, beta, ctx.act, dx, dgamma, dbeta, virt_h=ctx.virt_h)
        return dx, dgamma, dbeta, None, None, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# FiLM, upsample, layout glue
# ------------------------------------------------------------------------------------------------
class FiLMFn(Function):
    """y = gamma * x + beta with (gamma | beta) = gb[..., :C] | gb[..., C:]   (vae-gan-v2.py:146-149)."""

    @staticmethod
    def forward(ctx, gb, x):
        n, h, w, c = x.shape
        y = torch.empty((n, h, w, c), dtype=x.dtype, device=x.device)
        ops.film_fwd(gb, x, y)
        ctx.save_for_backward(gb, x)
        return y

    @staticmethod
    def backward(ctx, dy):
        gb, x = ctx.saved_tensors
        dy = grad_in(dy, x.dtype)
        if not dy.is_contiguous():
            dy = ops.dense_nhwc(dy)
        dgb = torch.empty_like(gb)
        dx = torch.empty(x.shape, dtype=x.dtype, device=x.device)
        ops.film_bwd(gb, x, dy, dgb, dx)
        return dgb, dx


class MaxPool2x2Fn(Function):
    """nn.MaxPool2d(2, 2) on an NHWC activation (VGG16 features of the perceptual loss, vae-gan.py:300-311)."""

    @staticmethod
    def forward(ctx, x):
        n, h, w, c = x.shape
        y = torch.empty((n, h // 2, w // 2, c), dtype=x.dtype, device=x.device)
        ops.maxpool_fwd(x, y)
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = grad_in(dy, x.dtype)
        if not dy.is_contiguous():
            dy = ops.dense_nhwc(dy)
        dx = torch.empty(x.shape, dtype=x.dtype, device=x.device)
        ops.maxpool_bwd(x, dy, dx)
        return dx


class FiLMRowsFn(Function):
    """FiLM with a 3-row parameter map gb3 [B,3,w,2C] (first | interior | last row class); exact restatement of
    FiLMFn for maps that are row-constant away from the border.  d gb3 is the gradient summed over each class."""

    @staticmethod
    def forward(ctx, gb3, x):
        n, h, w, c = x.shape
        if not gb3.is_contiguous():
            gb3 = ops.dense_nhwc(gb3)
        y = torch.empty((n, h, w, c), dtype=x.dtype, device=x.device)
        ops.film_rows_fwd(gb3, x, y)
        ctx.save_for_backward(gb3, x)
        return y

    @staticmethod
    def backward(ctx, dy):
        gb3, x = ctx.saved_tensors
        dy = grad_in(dy, x.dtype)
        if not dy.is_contiguous():
            dy = ops.dense_nhwc(dy)
        dgb3 = torch.empty_like(gb3)
        dx = torch.empty(x.shape, dtype=x.dtype, device=x.device)
        ops.film_rows_bwd(gb3, x, dy, dgb3, dx)
        return dgb3, dx


class UpsampleWFn(Function):
    """F.interpolate(t, size=(h, w), mode='bilinear', align_corners=False) for a (1 x w0) map (vae-gan-v2.py:138-140)."""

    @staticmethod
    def forward(ctx, t, h: int, w: int):
        n, _, w0, c = t.shape
        y = torch.empty((n, h, w, c), dtype=t.dtype, device=t.device)
        ops.upsample_w_fwd(t, y)
        ctx.w0, ctx.dt = w0, t.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = grad_in(dy, ctx.dt)
        if not dy.is_contiguous():
            dy = ops.dense_nhwc(dy)
        n, h, w, c = dy.shape
        dt = torch.empty((n, 1, ctx.w0, c), dtype=F32, device=dy.device)
        ops.upsample_w_bwd(dy, dt)
        if ctx.dt == F32:
            return dt, None, None
        out = torch.empty((n, 1, ctx.w0, c), dtype=BF16, device=dy.device)
        ops.strided_copy(dt, out)
        return out, None, None


def upsample2d_fwd(t: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """Separable bilinear resize [n,h0,w0,c] -> [n,h,w,c] (align_corners=False): H pass on the narrow map, then W pass."""
    n, h0, w0, c = t.shape
    if not t.is_contiguous():
        t = ops.dense_nhwc(t)
    th = t
    if h != h0:
        th = torch.empty((n, h, w0, c), dtype=t.dtype, device=t.device)
        ops.upsample_h_fwd(t, th)
    y = torch.empty((n, h, w, c), dtype=t.dtype, device=t.device)
    ops.upsample_w_fwd(th.view(n * h, 1, w0, c), y.view(n * h, 1, w, c))
    return y


def upsample2d_bwd(dy: torch.Tensor, h0: int, w0: int) -> torch.Tensor:
    """Adjoint of upsample2d_fwd: dy [n,h,w,c] (dense) -> gradient [n,h0,w0,c] in dy's dtype (fp32 accumulation)."""
    if not dy.is_contiguous():
        dy = ops.dense_nhwc(dy)
    n, h, w, c = dy.shape
    dth = torch.empty((n * h, 1, w0, c), dtype=F32, device=dy.device)
    ops.upsample_w_bwd(dy.view(n * h, 1, w, c), dth)
    dt = dth.view(n, h, w0, c)
    if h != h0:
        dt = torch.empty((n, h0, w0, c), dtype=F32, device=dy.device)
        ops.upsample_h_bwd(dth.view(n, h, w0, c), dt)
    if dy.dtype == F32:
        return dt
    out = torch.empty((n, h0, w0, c), dtype=BF16, device=dy.device)
    ops.strided_copy(dt, out)
    return out


class Upsample2DFn(Function):
    """F.interpolate(t, size=(h, w), mode='bilinear', align_corners=False) of a multi-row NHWC map [n,h0,w0,c]
    (the 4-row text map of vae-gan-oldv.py:165-176 and its (1, W/8) resize at :286-291).  Bilinear interpolation is
    separable: the H pass runs on the narrow map, the W pass then writes the full-size tensor once."""

    @staticmethod
    def forward(ctx, t, h: int, w: int):
        ctx.h0, ctx.w0, ctx.dt = t.shape[1], t.shape[2], t.dtype
        return upsample2d_fwd(t, h, w)

    @staticmethod
    def backward(ctx, dy):
        return upsample2d_bwd(grad_in(dy, ctx.dt), ctx.h0, ctx.w0), None, None


class ChannelGateFn(Function):
    """skip * sigmoid(alpha) (GatedSkipConnection, vae-gan-oldv.py:226-231) written straight into the skip half of a
    concat buffer; ``scale`` = sigmoid(alpha) as fp32 [C] (the sigmoid of C numbers and its gradient stay in autograd).
    The backward is one pass over (x, dy): dx = dy * scale and dscale = sum over pixels of dy * x."""

    @staticmethod
    def forward(ctx, x, scale, out):
        n, h, w, c = x.shape
        scale = scale.detach().contiguous()
        y = out if out is not None else new_act(n, h, w, c, x.device, x.dtype)
        ops.channel_scale_fwd(x, scale, y)
        ctx.save_for_backward(x, scale)
        return y.detach() if out is not None else y

    @staticmethod
    def backward(ctx, dy):
        x, scale = ctx.saved_tensors
        n, h, w, c = x.shape
        dy = grad_in(dy, x.dtype)
        dx = new_act(n, h, w, c, x.device, x.dtype)
        ds = torch.empty(2 * c, dtype=F32, device=x.device)
        ops.channel_scale_bwd(x, dy, scale, dx, ds)
        return dx, ds[:c], None


class EmbeddingFn(Function):
    """nn.Embedding(padding_idx) lookup (vae-gan-v2.py:83,104): a row gather forward; backward = per-vocabulary-row sum of
    the token gradients in a fixed order (no index sort, no atomics), padding row zero."""

    @staticmethod
    def forward(ctx, idx, weight, padding_idx: int):
        idx = idx.contiguous()
        ctx.save_for_backward(idx)
        ctx.vocab, ctx.pad = weight.shape[0], padding_idx
        return ops.embedding_fwd(idx, weight.detach().contiguous())

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        return None, ops.embedding_bwd(idx, g.float(), ctx.vocab, ctx.pad), None


class SeqPoolFn(Function):
    """nn.AdaptiveAvgPool1d(w) over the time axis of a sequence (B, L, C) -- or of an NHWC activation [B,1,L,C] -- written
    as the NHWC text map [B,1,w,C] (vae-gan-v2.py:107-113: adaptive_pool(out.permute(0,2,1)).unsqueeze(2), plus the change
    to this package's activation layout, in one pass).  ``out_dtype`` None = the activation dtype."""

    @staticmethod
    def forward(ctx, seq, w: int, out_dtype=None):
        b, l, c = seq.shape[0], seq.shape[-2], seq.shape[-1]
        src = seq.detach()
        if src.stride(-1) != 1 or src.stride(0) != l * src.stride(-2):
            src = src.contiguous()
        out = new_act(b, 1, w, c, seq.device, out_dtype or ops.act_dtype())
        ops.seqpool_fwd(src, out)
        ctx.shape, ctx.dt = tuple(seq.shape), seq.dtype
        return out

    @staticmethod
    def backward(ctx, dy):
        dy = grad_in(dy, dy.dtype if dy.dtype in (BF16, F32) else None)
        dseq = torch.empty(ctx.shape, dtype=ctx.dt, device=dy.device)
        ops.seqpool_bwd(dy, dseq)
        return dseq, None, None


class _gru_gemm_steps:
    """bf16 mode: let the GRU's split-bf16 GEMMs accumulate their whole K extent (<= 48 steps) in one TMEM pass."""

    def __enter__(self):
        from . import conv
        self.prev = conv.HI_MAX_STEPS
        if ops.act_dtype() == BF16:
            conv.HI_MAX_STEPS = 64
        return self

    def __exit__(self, *exc):
        from . import conv
        conv.HI_MAX_STEPS = self.prev
        return False


class GRULayerFn(Function):
    """One bidirectional, batch_first GRU layer with hidden size 256 (torch.nn.GRU semantics; the text encoder of
    vae-gan-v2.py:84-89,105).  The recurrence -- the part that costs the stock path ~1000 launches per training step --
    is one cluster kernel per direction pair (vg_gru.cu).  The time-parallel GEMMs around it (x W_ih^T for all t, dx,
    dW_ih, dW_hh) run on the tcgen05 implicit-GEMM kernels as 1x1 convolutions over the [B,1,T,C] view of the sequence,
    always with split-bf16 operands (three bf16 planes per fp32 value, fp32 accumulation: ~1e-6 relative, tighter than
    the TF32 GEMMs cuDNN's GRU uses) -- they are ~1 GFLOP, so the 6x tensor work is free and no library GEMM is left
    on the path."""

    @staticmethod
    def forward(ctx, x, cache, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r):
        b, t, i = x.shape
        h = w_hh_f.shape[1]
        op = ConvLinear(i, 6 * h, 1, 1)

        def build():
            w_ih = torch.cat([w_ih_f.detach(), w_ih_r.detach()], 0).contiguous()       # [6H, I]
            w4 = w_ih.view(6 * h, i, 1, 1)
            return {"wf": op.prep_fwd(w4, None, True), "wb": op.prep_bwd(w4, None, True),
                    "b_ih": torch.cat([b_ih_f.detach(), b_ih_r.detach()], 0).contiguous(),
                    "w_hh": torch.stack([w_hh_f.detach(), w_hh_r.detach()], 0).contiguous(),   # [2, 3H, H]
                    "b_hh": torch.stack([b_hh_f.detach(), b_hh_r.detach()], 0).contiguous()}
        pk = cache.get("gru", (w_ih_f, w_ih_r, b_ih_f, b_ih_r, w_hh_f, w_hh_r, b_hh_f, b_hh_r), build)
        x4 = x.detach().float().reshape(b, 1, t, i).contiguous()
        with _gru_gemm_steps():
            xproj = op.forward(x4, pk["wf"], pk["b_ih"])                # fp32 [B,1,T,6H] == [B, T, 2, 3H]
        out = torch.empty((b, t, 2 * h), dtype=F32, device=x.device)
        gates = torch.empty((2, b, t, 4, h), dtype=F32, device=x.device)
        ops.gru_seq_fwd(xproj.view(b, t, 2, 3 * h), pk["w_hh"], pk["b_hh"], out, gates)
        ctx.save_for_backward(x4, pk["w_hh"], out, gates)
        ctx.op, ctx.wb, ctx.dims = op, pk["wb"], (b, t, i, h)
        return out

    @staticmethod
    def backward(ctx, dout):
        x4, w_hh, out, gates = ctx.saved_tensors
        b, t, i, h = ctx.dims
        op: ConvLinear = ctx.op
        dout = dout.contiguous().float()
        dgx = torch.empty((b, t, 2, 3 * h), dtype=F32, device=dout.device)
        dgh = torch.empty((2, b, t, 3 * h), dtype=F32, device=dout.device)
        ops.gru_seq_bwd(dout, out, gates, w_hh, dgx, dgh)
        dgx4 = dgx.view(b, 1, t, 6 * h)
        hprev = torch.zeros((2, b, t, h), dtype=F32, device=dout.device)  # h_{t-1} of each direction's recurrence
        hprev[0, :, 1:] = out[:, :-1, :h]
        hprev[1, :, :-1] = out[:, 1:, h:]
        with _gru_gemm_steps():
            dx = op.backward_data(dgx4, ctx.wb, (1, t)).view(b, t, i) if ctx.needs_input_grad[0] else None
        dw_ih = op.backward_weight(dgx4, x4).reshape(6 * h, i)                                # [6H, I]
        op_hh = ConvLinear(h, 3 * h, 1, 1)
        dw_hh = [op_hh.backward_weight(dgh[d].view(b, 1, t, 3 * h), hprev[d].view(b, 1, t, h)).reshape(3 * h, h)
                 for d in (0, 1)]
        # bias gradients = column sums over all (batch, time) rows: the row-parallel statistics kernel
        db_ih = ops.norm_stats(dgx4, per_sample=False)[0, 0]
        db_hh = [ops.norm_stats(dgh[d].view(b, 1, t, 3 * h), per_sample=False)[0, 0] for d in (0, 1)]
        return (dx, None, dw_ih[:3 * h], dw_hh[0], db_ih[:3 * h], db_hh[0], dw_ih[3 * h:], dw_hh[1], db_ih[3 * h:], db_hh[1])


class ToNHWCFn(Function):
    """NCHW fp32 -> NHWC bf16 (module boundary / stock-torch text encoder output)."""

    @staticmethod
    def forward(ctx, t):
        n, c, h, w = t.shape
        out = new_act(n, h, w, c, t.device)
        ops.strided_copy(t.permute(0, 2, 3, 1), out)
        return out

    @staticmethod
    def backward(ctx, dy):
        n, h, w, c = dy.shape
        out = torch.empty((n, c, h, w), dtype=F32, device=dy.device)
        ops.strided_copy(dy.permute(0, 3, 1, 2), out)
        return out


class CatSlicesFn(Function):
    """torch.cat([a, b], dim=C) where both already sit in adjacent channel slices of one buffer: no copy in
    either direction (vae-gan-v2.py:251-274, vae-gan-unet.py:232-248)."""

    @staticmethod
    def forward(ctx, a, b, buf):
        ca, cb = a.shape[3], b.shape[3]
        assert (a.data_ptr() == buf.data_ptr() and b.data_ptr() == buf.data_ptr() + buf.element_size() * ca
                and buf.shape[3] == ca + cb)
        ctx.ca = ca
        return buf.detach()

    @staticmethod
    def backward(ctx, d):
        d = grad_in(d, d.dtype if d.dtype in (BF16, F32) else None)
        return d[..., :ctx.ca], d[..., ctx.ca:], None


class ZTextCatFn(Function):
    """cat([z.expand(-1,-1,1,w0), text], C) as an NHWC bf16 [B,1,w0,z+ct] tensor (vae-gan-v2.py:249-251; with w0 = 1
    it is cat([z, spatial_broadcast(text)]) of vae-gan.py:143-145).  z: fp32 [B, zc]; text: NHWC bf16 [B,1,w0,ct]."""

    @staticmethod
    def forward(ctx, z, text):
        b, zc = z.shape
        _, _, w0, ct = text.shape
        out = new_act(b, 1, w0, zc + ct, z.device, text.dtype)
        ops.strided_copy(z.view(b, 1, 1, zc).expand(b, 1, w0, zc), out[..., :zc])
        ops.strided_copy(text, out[..., zc:])
        ctx.zc = zc
        return out

    @staticmethod
    def backward(ctx, d):
        d = grad_in(d, d.dtype if d.dtype in (BF16, F32) else None)
        zc = ctx.zc
        dz = ops.norm_stats(d[..., :zc], per_sample=True)[:, 0, :].contiguous()   # sum over the w0 columns, fp32
        return dz, d[..., zc:]


# ------------------------------------------------------------------------------------------------
# reparameterisation + KL, output sigmoid, losses
# ------------------------------------------------------------------------------------------------
class ReparamKLFn(Function):
    """(mu, logvar, z, kl) from the raw head GEMM result (vae-gan.py:133-136,420).  ``eps`` comes from
    torch.randn_like on the host side so the RNG stream matches the reference."""

    @staticmethod
    def forward(ctx, heads, bias_mu, bias_lv, eps):
        b = heads.shape[0]
        h2 = heads.reshape(b, -1)
        mu, lv, z, kl = ops.reparam_kl_fwd(h2, bias_mu.detach(), bias_lv.detach(), eps)
        ctx.save_for_backward(mu, lv, eps)
        ctx.heads_shape = heads.shape
        return mu, lv, z, kl

    @staticmethod
    def backward(ctx, dmu, dlv, dz, dkl):
        mu, lv, eps = ctx.saved_tensors
        c = lambda t: t.contiguous() if t is not None else None
        dheads = ops.reparam_kl_bwd(mu, lv, eps, c(dz), c(dmu), c(dlv), c(dkl))
        z = mu.shape[1]
        db_mu = torch.empty(z, dtype=F32, device=mu.device)
        db_lv = torch.empty(z, dtype=F32, device=mu.device)
        ops.colsum_f32(dheads[:, :z], db_mu)
        ops.colsum_f32(dheads[:, z:], db_lv)
        return dheads.view(ctx.heads_shape), db_mu, db_lv, None


class SigmoidOutFn(Function):
    """NHWC fp32 pre-activation -> NCHW fp32 sigmoid image (vae-gan.py:82)."""

    @staticmethod
    def forward(ctx, pre):
        n, h, w, c = pre.shape
        y = torch.empty((n, c, h, w), dtype=F32, device=pre.device)
        ops.sigmoid_fwd(pre, y)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        n, c, h, w = y.shape
        dpre = torch.empty((n, h, w, c), dtype=F32, device=y.device)
        ops.sigmoid_bwd(y, dy.contiguous(), dpre)
        return dpre


class L1LossFn(Function):
    """nn.L1Loss() (mean) -- vae-gan.py:419,537."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        ctx.save_for_backward(a, b)
        return ops.l1_fwd(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = torch.empty_like(a)
        ops.l1_bwd(a, b, g.contiguous(), da)
        return da, None


class HingeLossFn(Function):
    """hinge_loss(preds, target) -- vae-gan.py:313-320.  mode: 1 real, 0 fake, 2 generator (target None)."""

    @staticmethod
    def forward(ctx, p, mode: int):
        p = p.contiguous()
        ctx.save_for_backward(p)
        ctx.mode = mode
        return ops.hinge_fwd(p, mode)

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        dp = torch.empty_like(p)
        ops.hinge_bwd(p, ctx.mode, g.contiguous(), dp)
        return dp, None


def l1_loss(a, b):
    return L1LossFn.apply(a, b)


def hinge_loss(preds, target):
    """Same call convention as the reference's hinge_loss(preds, target) with target in {1, 0, None}."""
    return HingeLossFn.apply(preds, 1 if target == 1 else (0 if target == 0 else 2))
